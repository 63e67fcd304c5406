# round-2 evidence on ONE B200 (run under gpurun); outputs -> gpurun_out/r2ev/
set -x
D=gpurun_out/r2ev; mkdir -p $D
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6 > $D/pytest_gpu_1xB200.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $D/smoke.txt 2>&1
timeout 900 python bench.py > $D/bench_n1.json 2> $D/bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $D/bench_reference_arm.json 2> $D/bench_ref.err
SACB_FOLD=0 timeout 300 python bench.py --no-cpu --no-extra > $D/bench_n1_first_program.json 2>/dev/null
timeout 200 python profiles/phase_timeline.py 8 > $D/phase_timeline_fold.txt 2>&1
SACB_FOLD=0 timeout 200 python profiles/phase_timeline.py 8 > $D/phase_timeline_first_program.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/ncu_launches_fold.csv python bench.py --steps 100 --warmup 50 --no-cpu --no-extra > $D/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sac_phase -s 4 -c 1 -o $D/ncu_persistent_fold -f python bench.py --steps 100 --warmup 50 --no-cpu --no-extra > $D/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 150 --csv --log-file $D/ncu_launches_visual_c5.csv python tools/visual_bench.py 2 1 > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_visual.py -q -m gpu -k "(golden and small and 0-1-1) or (golden and ragged and 0-1-1) or visual_small or pi_from_states and small-0" > $D/sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -q -m gpu -k "golden and small and 0-1-1" > $D/sanitizer_racecheck.log 2>&1
tail -3 $D/pytest_gpu_1xB200.txt $D/smoke.txt $D/sanitizer_memcheck.log $D/sanitizer_racecheck.log
cut -c1-400 $D/bench_n1.json; cut -c1-300 $D/bench_reference_arm.json
ls -la $D
