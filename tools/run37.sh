cd /root/repo
mkdir -p gpurun_out/r2b
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r2b/pytest_gpu_1xB200.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2b/smoke.txt
timeout 900 python bench.py > gpurun_out/r2b/bench_n1.json 2> gpurun_out/r2b/bench_n1.err; cut -c1-600 gpurun_out/r2b/bench_n1.json; tail -2 gpurun_out/r2b/bench_n1.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2b/bench_n1.json").read().strip().splitlines()[-1])
print("value",d["value"],"e2e",d["e2e"]["value"],"roof",d["roofline"]["frac"])
print("extra",json.dumps(d["config"].get("extra_workloads"))[:900])
PY
