cd /root/repo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 1000 --warmup 100 --no-extra > gpurun_out/r2_bench_n8b.json 2> gpurun_out/r2_bench_n8b.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n8b.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["config"]["timing"], d["config"]["ms_repeats_min_max"], d["e2e"]["value"], d["config"]["replicas_identical"])
PY
