cd /root/repo
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv
SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 60 python profiles/phase_timeline.py 6 2>&1 | tail -4
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bf16" 2>&1 | tail -3
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv
