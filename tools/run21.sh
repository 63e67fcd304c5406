cd /root/repo
timeout 300 python tools/tma_check.py ragged8 > gpurun_out/r2_tma_check_ragged.txt 2>&1; echo "rc=$?"; tail -40 gpurun_out/r2_tma_check_ragged.txt
timeout 300 python tools/tma_check.py humanoid > gpurun_out/r2_tma_check_hum.txt 2>&1; echo "rc=$?"; tail -30 gpurun_out/r2_tma_check_hum.txt
SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 120 python profiles/phase_timeline.py 6 > gpurun_out/r2_tl_c2_tma.txt 2>&1; tail -28 gpurun_out/r2_tl_c2_tma.txt
