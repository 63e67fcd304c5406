cd /root/repo
timeout 300 python bench.py --no-cpu --no-extra 2>gpurun_out/qb.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e']['value'], d['clocks'], d['config']['ms_repeats_min_max'])"
tail -2 gpurun_out/qb.err
