python -m pytest tests -m gpu -q --deselect tests/test_multi_gpu.py 2>&1 | grep -v "^$" | tail -12
python bench.py --steps 500 --warmup 50 > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; tail -3 gpurun_out/r2_bench13.err
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench13.json"))
print("c1", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d["config"]["extra_workloads"].items(): print(k, {a:b for a,b in v.items() if a in ("value","ms_per_step","error","roofline","cpu_baseline","gpu_launches_per_update","finite_losses")})
PY
