for cta in 3 16 140; do
SACB200_LIB=$PWD/torch-actor-critic_b200/libsacb200_trace.so SACB_TRACE_CTA=$cta python profiles/phase_timeline.py 3 > gpurun_out/r2_tl_trace7.txt 2> gpurun_out/r2_trace7_cta$cta.txt
done
tail -3 gpurun_out/r2_tl_trace7.txt
python - <<PY
for cta in (3,16,140):
    ev=[l.split() for l in open(f"gpurun_out/r2_trace7_cta{cta}.txt").read().split("\n")[1:] if l.strip()]
    idx=[i for i,e in enumerate(ev) if e[0]=="1002"]
    s=idx[1] if len(idx)>1 else 0
    print("CTA",cta, ";".join(f"{a} {b}" for a,b in ev[s:s+135]))
PY
SACB_WORKLOAD=c2 SACB_PRECISION=3 python profiles/phase_timeline.py 6 > gpurun_out/r2_tl_c2_bf16.txt 2>&1
cat gpurun_out/r2_tl_c2_bf16.txt
