timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python tools/_dbg_stack.py 2>&1 | grep -c "px 0 max"
for c in A C; do python tools/stage_times.py $c --bwd 2>&1 | tail -2 | head -1; done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['roofline']['stages']; print('views/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'pre', s['preprocess']['ms'], 'fwd', s['blend_fwd']['ms'], 'bwd', s['blend_bwd']['ms'])"
