cd /root/repo
for v in 1 0 1 0; do
SACB_TMA_PREB=$v SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 120 python profiles/phase_timeline.py 10 2>&1 | grep -E "SUM|per update"
done
