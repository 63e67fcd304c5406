cd /root/repo
for b in 256 512; do
SACB_BATCH=$b SACB_WORKLOAD=c2 SACB_PRECISION=3 timeout 120 python profiles/phase_timeline.py 6 2>&1 | grep -E "adam|SUM|per update|C.dW1|T.L1 "
done
