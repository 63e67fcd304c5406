python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8
python profiles/phase_timeline.py 8 2>&1 | tail -14
SACB200_LIB=$PWD/torch-actor-critic_b200/libsacb200_trace.so SACB_TRACE_CTA=0 python profiles/phase_timeline.py 3 > /dev/null 2> gpurun_out/r2_trace3_cta0.txt
sed -n 1,140p gpurun_out/r2_trace3_cta0.txt | tr '\n' ';' | head -c 3500
