cd /root/repo
R=$PWD/torch-actor-critic_b200
for i in 1 2; do SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 120 python profiles/phase_timeline.py 6 2>&1 | tail -25 | awk '{printf "%s %s/%s | ", $1, $(NF-3), $(NF-1)} END {print ""}'; done
for lib in libsacb200_old.so libsacb200.so; do echo "== $lib"; SACB200_LIB=$R/$lib timeout 120 python tools/burst_time.py | tail -1; done
