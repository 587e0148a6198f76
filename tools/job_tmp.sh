run() { python bench.py --steps 5 --warmup 3 --no-cpu-baseline --graphs 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['roofline']['stages']; print('$1', 'views/s', round(d['value'],1), 'fwd', s['blend_fwd']['ms'], 'bwd', s['blend_bwd']['ms'])"; }
run base
BRS_BWD_PAD_SMEM=22000 run bwd4cta
BRS_BWD_PAD_SMEM=40000 run bwd3cta
BRS_FWD_PAD_SMEM=8000 run fwd6cta
BRS_FWD_PAD_SMEM=8000 BRS_BWD_PAD_SMEM=22000 run both
