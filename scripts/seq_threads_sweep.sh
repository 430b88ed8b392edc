# c2 frame loop: host threads of the sequence entry point x work items per warp (python bench.py, device-resident leg)
for t in ${THREADS:-1 2 3 4}; do for ipw in ${IPW:-4}; do echo "SEQ_THREADS=$t ITEMS_PER_WARP=$ipw"; CVGS_B200_SEQ_THREADS=$t CVGS_TMA_ITEMS_PER_WARP=$ipw python bench.py --no-baselines --steps 50 --warmup 5 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(' value', round(d['value']), 'us/launch', round(d['roofline']['us_per_launch'],3), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']))
"; done; done
