cd /root/repo
R=$PWD/torch-actor-critic_b200
for lib in libsacb200_old.so libsacb200.so libsacb200_v1.so libsacb200_v2.so libsacb200_v3.so libsacb200_old.so; do echo "== $lib"; SACB200_LIB=$R/$lib timeout 120 python tools/burst_time.py | tail -1; done
