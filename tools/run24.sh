cd /root/repo
timeout 300 python tools/tma_check.py humanoid 2>&1 | grep -v EXACT | tail -12
SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 120 python profiles/phase_timeline.py 6 > gpurun_out/r2_tl_c2_tma2.txt 2>&1; tail -26 gpurun_out/r2_tl_c2_tma2.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -8
