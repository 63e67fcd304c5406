cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_train_loop.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/tma_check.py humanoid 2>&1 | grep -v EXACT | tail -6
SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 120 python profiles/phase_timeline.py 6 > gpurun_out/r2_tl_c2_tma3.txt 2>&1; tail -26 gpurun_out/r2_tl_c2_tma3.txt
timeout 120 python profiles/phase_timeline.py 8 > gpurun_out/r2_tl_c1_adamfast.txt 2>&1; tail -14 gpurun_out/r2_tl_c1_adamfast.txt
