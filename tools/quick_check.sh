cd /root/repo
SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 60 python profiles/phase_timeline.py 6 2>&1 | tail -9
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_train_loop.py -x -q -m gpu 2>&1 | tail -3
