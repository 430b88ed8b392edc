for sl in 2 3 4; do for ipw in 6 8 12; do echo -n "SLOTS=$sl IPW=$ipw "; CVGS_TMA_SLOTS=$sl CVGS_TMA_ITEMS_PER_WARP=$ipw python bench.py --no-baselines --steps 50 --warmup 5 2>/dev/null | python -c "
import sys, json
for line in sys.stdin:
    line=line.strip()
    if line.startswith('{'):
        d=json.loads(line); print(' value', round(d['value']), 'us/launch', round(d['roofline']['us_per_launch'],3))
"; done; done
