python tools/fold_debug.py small 2>&1 | tail -12
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -15
for cta in 0 70 100 140; do
  SACB200_LIB=$PWD/torch-actor-critic_b200/libsacb200_trace.so SACB_TRACE_CTA=$cta python profiles/phase_timeline.py 3 > /dev/null 2> gpurun_out/r2_trace_cta$cta.txt
done
wc -l gpurun_out/r2_trace_cta*.txt
