# round-2 (second half) evidence on ONE B200 (run under gpurun); outputs -> gpurun_out/r2b/
set -x
cd /root/repo
D=gpurun_out/r2b; mkdir -p $D
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -6 > $D/pytest_gpu_1xB200.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.txt 2>&1
timeout 900 python bench.py > $D/bench_n1_c1_c2_c5.json 2> $D/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $D/bench_reference_arm.json 2> $D/bench_ref.err
timeout 300 python bench.py --workload c2 --no-cpu --no-extra > $D/bench_c2_as_main_workload.json 2>/dev/null
SACB_TMA=0 timeout 300 python bench.py --workload c2 --no-cpu --no-extra > $D/bench_c2_staged_operands_SACB_TMA0.json 2>/dev/null
SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 200 python profiles/phase_timeline.py 8 > $D/phase_timeline_c2_tma.txt 2>&1
SACB_TMA=0 SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 200 python profiles/phase_timeline.py 8 > $D/phase_timeline_c2_staged.txt 2>&1
timeout 200 python profiles/phase_timeline.py 8 > $D/phase_timeline_c1_fold.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $D/ncu_launches_c2.csv python bench.py --workload c2 --steps 100 --warmup 50 --no-cpu --no-extra > $D/ncu_launches_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sac_phase -s 2 -c 1 -o $D/ncu_c2_tma -f python bench.py --workload c2 --steps 50 --warmup 50 --no-cpu --no-extra > $D/ncu_full_c2.log 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "bf16 and ragged8 and 0 or tma_operands and ragged8" > $D/sanitizer_memcheck_tma.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "tma_operands and ragged8" > $D/sanitizer_racecheck_tma.log 2>&1
tail -3 $D/pytest_gpu_1xB200.txt $D/smoke.txt $D/sanitizer_memcheck_tma.log $D/sanitizer_racecheck_tma.log
for f in bench_n1_c1_c2_c5 bench_reference_arm bench_c2_as_main_workload bench_c2_staged_operands_SACB_TMA0; do python - <<PY
import json
d=json.loads(open("$D/$f.json").read().strip().splitlines()[-1])
print("$f", d.get("value"), d.get("e2e",{}).get("value"), (d.get("roofline") or {}).get("frac"), json.dumps((d.get("config") or {}).get("extra_workloads",{}).get("c2",{}).get("value")))
PY
done
tail -4 $D/phase_timeline_c2_tma.txt
ls -la $D
