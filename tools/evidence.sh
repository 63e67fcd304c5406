set -x
mkdir -p gpurun_out/v6
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/v6/pytest_gpu.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v6/smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/v6/bench_n1.json 2> gpurun_out/v6/bench_n1.err
timeout 600 python bench.py --impl reference --steps 300 --warmup 20 > gpurun_out/v6/bench_ref.json 2> gpurun_out/v6/bench_ref.err
timeout 300 python bench.py --no-cpu --workload c2 --precision 3 > gpurun_out/v6/bench_c2_bf16.json 2>/dev/null
timeout 300 python bench.py --no-cpu --workload c2 --precision 1 > gpurun_out/v6/bench_c2_tf32x3.json 2>/dev/null
timeout 200 python profiles/phase_timeline.py 8 > gpurun_out/v6/phase_timeline_c1.txt
SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 200 python profiles/phase_timeline.py 6 > gpurun_out/v6/phase_timeline_c2_bf16.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/v6/ncu_launches.csv python bench.py --steps 100 --warmup 50 --no-cpu > gpurun_out/v6/ncu_launches.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/v6/ncu_launches_phased.csv python bench.py --steps 6 --warmup 3 --mode 1 --no-cpu > gpurun_out/v6/ncu_launches_phased.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sac_phase -s 4 -c 1 -o gpurun_out/v6/ncu_persistent -f python bench.py --steps 100 --warmup 50 --no-cpu > gpurun_out/v6/ncu_full.log 2>&1
ls -la gpurun_out/v6
cat gpurun_out/v6/pytest_gpu.txt gpurun_out/v6/smoke.txt | tail -8
cut -c1-300 gpurun_out/v6/bench_n1.json; cut -c1-400 gpurun_out/v6/bench_ref.json
