SACB_GRID=64 SACB200_LIB=$PWD/torch-actor-critic_b200/libsacb200_trace.so SACB_TRACE_CTA=3 python profiles/phase_timeline.py 3 > gpurun_out/r2_tl_grid64.txt 2> gpurun_out/r2_trace_grid64_cta3.txt
tail -3 gpurun_out/r2_tl_grid64.txt
sed -n 1,400p gpurun_out/r2_trace_grid64_cta3.txt | tr '\n' ';' | head -c 5000
