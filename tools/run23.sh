cd /root/repo
SACB_WORKLOAD=c2 SACB_PRECISION=3 SACB_DUMP_PHASES=2,8,9,10,15,17,18,19,20 timeout 120 python profiles/phase_timeline.py 4 2>&1 | tail -12 | cut -c1-1500
