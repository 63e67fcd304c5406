cd /root/repo
R=$PWD/torch-actor-critic_b200
for lib in libsacb200_old.so libsacb200.so libsacb200_old.so libsacb200.so; do echo "== $lib"; SACB200_LIB=$R/$lib timeout 120 python tools/burst_time.py | tail -1; done
SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 120 python profiles/phase_timeline.py 10 2>&1 | grep -E "adam|SUM|per update"
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4
