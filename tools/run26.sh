cd /root/repo
mkdir -p gpurun_out/r2c2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sac_phase -s 2 -c 1 -o gpurun_out/r2c2/ncu_c2_tma -f python bench.py --workload c2 --steps 50 --warmup 50 --no-cpu --no-extra > gpurun_out/r2c2/ncu_full.log 2>&1
tail -3 gpurun_out/r2c2/ncu_full.log
ls -la gpurun_out/r2c2
