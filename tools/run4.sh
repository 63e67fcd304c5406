python tools/fold_debug.py small > gpurun_out/r2_folddbg.log 2>&1
python tools/fold_debug.py c1 >> gpurun_out/r2_folddbg.log 2>&1
python profiles/phase_timeline.py 8 > gpurun_out/r2_timeline_fold.txt 2>&1
tail -30 gpurun_out/r2_folddbg.log; cat gpurun_out/r2_timeline_fold.txt | tail -20
