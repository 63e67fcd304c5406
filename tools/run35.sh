cd /root/repo
timeout 300 python tools/tma_check.py humanoid ragged8 2>&1 | grep -v "TMA operand" | awk '{print $1,$2,$3,$4,$5,$6}' | tail -32
SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 120 python profiles/phase_timeline.py 6 > gpurun_out/r2_tl_c2_tma5.txt 2>&1; tail -26 gpurun_out/r2_tl_c2_tma5.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
