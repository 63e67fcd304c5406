python -m pytest tests/test_gpu_visual.py tests/test_gpu_parity.py -m gpu -q -x -k "visual or golden or burst" 2>&1 | tail -4
python tools/visual_bench.py 10 1
python profiles/phase_timeline.py 8 2>&1 | tail -12
ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 160 --csv --log-file gpurun_out/r2_visual_launches2.csv python tools/visual_bench.py 2 1 > /dev/null 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r2_visual_launches2.csv")) if len(r)>5 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    agg[r[4]][0]+=1; agg[r[4]][1]+=float(r[-1].replace(",",""))
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda x:-x[1][1])[:10]: print(f"{v[1]/1e3:9.1f} us {v[0]:4d}x {100*v[1]/tot:5.1f}%  {k[:80]}")
print("total us", tot/1e3)
PY
