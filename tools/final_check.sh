cd /root/repo
timeout 600 python bench.py > gpurun_out/final_bench_n1.json 2> gpurun_out/final_bench_n1.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1000 --warmup 100 --no-extra > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err
python - <<PY
import json
for f in ("gpurun_out/final_bench_n1.json","gpurun_out/final_bench_n2.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, d["value"], d["ms_per_step"], d["config"]["timing"], d["config"]["ms_repeats_min_max"], "e2e", d["e2e"]["value"], d["clocks"], (d["config"].get("extra_workloads") or {}).get("c2",{}).get("value"))
PY
