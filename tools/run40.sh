cd /root/repo
R=$PWD/torch-actor-critic_b200
for lib in libsacb200_old.so libsacb200.so; do echo "== $lib"; SACB200_LIB=$R/$lib timeout 120 python profiles/phase_timeline.py 20 2>&1 | tail -11; done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tma_operands and humanoid" 2>&1 | grep -E "AssertionError|assert|Error" | head -8
