cd /root/repo
D=gpurun_out/r2c; mkdir -p $D
timeout 600 python bench.py > $D/bench_n1_c1_c2_c5.json 2> $D/bench_n1.err
SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 100 python profiles/phase_timeline.py 8 > $D/phase_timeline_c2_tma.txt 2>&1
timeout 200 python bench.py --workload c2 --no-cpu --no-extra > $D/bench_c2_as_main_workload.json 2>/dev/null
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3 > $D/pytest_gpu_1xB200.txt
python - <<PY
import json
for f in ("bench_n1_c1_c2_c5","bench_c2_as_main_workload"):
    d=json.loads(open("$D/"+f+".json").read().strip().splitlines()[-1])
    print(f, d["value"], d["e2e"]["value"], d["roofline"]["frac"], (d["config"].get("extra_workloads") or {}).get("c2",{}).get("value"), (d["config"].get("extra_workloads") or {}).get("c2",{}).get("roofline"), (d["config"].get("extra_workloads") or {}).get("c5",{}).get("value"))
PY
tail -3 $D/phase_timeline_c2_tma.txt; cat $D/pytest_gpu_1xB200.txt
