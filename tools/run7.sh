for cta in 0 100; do
  SACB200_LIB=$PWD/torch-actor-critic_b200/libsacb200_trace.so SACB_TRACE_CTA=$cta python profiles/phase_timeline.py 3 > /dev/null 2> gpurun_out/r2_trace2_cta$cta.txt
done
sed -n 1,140p gpurun_out/r2_trace2_cta0.txt | tr '\n' ';' | head -c 6000
