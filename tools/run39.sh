cd /root/repo
R=$PWD/torch-actor-critic_b200
for lib in libsacb200_old.so libsacb200.so; do echo "== $lib"; SACB200_LIB=$R/$lib timeout 120 python tools/burst_time.py; done
