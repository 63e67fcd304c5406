cd /root/repo
R=$PWD/torch-actor-critic_b200
for i in 1 2; do
for lib in libsacb200_old.so libsacb200.so; do
echo "== $lib"
SACB200_LIB=$R/$lib timeout 120 python profiles/phase_timeline.py 10 2>&1 | grep -E "^F|SUM|per update" | awk '{printf "%s %s | ", $1, $(NF-3)} END {print ""}'
SACB200_LIB=$R/$lib timeout 300 python bench.py --no-cpu --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench value', d['value'], 'e2e', d['e2e']['value'])"
done
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tma" 2>&1 | tail -3
