python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
python profiles/phase_timeline.py 8 2>&1 | tail -12
SACB200_LIB=$PWD/torch-actor-critic_b200/libsacb200_trace.so SACB_TRACE_CTA=3 python profiles/phase_timeline.py 3 > /dev/null 2> gpurun_out/r2_trace5_cta3.txt
sed -n 1,140p gpurun_out/r2_trace5_cta3.txt | tr '\n' ';' | head -c 2400
