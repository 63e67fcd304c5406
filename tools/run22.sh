cd /root/repo
for cta in 3 100; do
SACB_WORKLOAD=c2 SACB_PRECISION=3 SACB200_LIB=$PWD/torch-actor-critic_b200/libsacb200_trace.so SACB_TRACE_CTA=$cta timeout 120 python profiles/phase_timeline.py 3 > gpurun_out/r2_tl_c2_trace.txt 2> gpurun_out/r2_trace_c2_cta$cta.txt
python - <<PY
ev=[l.split() for l in open("gpurun_out/r2_trace_c2_cta$cta.txt").read().split("\n")[1:] if l.strip() and l.split()[0].isdigit()]
idx=[i for i,e in enumerate(ev) if e[0]=="1001"]
s=idx[1] if len(idx)>1 else 0
print("CTA $cta", ";".join(f"{a} {b}" for a,b in ev[s:s+330]))
PY
done
tail -5 gpurun_out/r2_tl_c2_trace.txt
