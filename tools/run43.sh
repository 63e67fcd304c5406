cd /root/repo
for i in 1 2; do SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 120 python profiles/phase_timeline.py 6 2>&1 | tail -25 | awk '{printf "%s %s/%s | ", $1, $(NF-3), $(NF-1)} END {print ""}'; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv
