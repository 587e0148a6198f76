#!/bin/bash
# Quick GPU check: parity tests + one bench line (no ncu).  Usage: bash tools/gpu_quick.sh <tag> [notests]
TAG=${1:-q}
mkdir -p gpurun_out
if [ "$2" != "notests" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  tail -3 gpurun_out/${TAG}_pytest.log
fi
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_ours.json').read().strip().splitlines()[-1])
print('views/s', round(d['value'],1), 'ms/view', round(d['ms_per_view'],4), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])
print(' '.join(f"{k}={v['ms']}" for k,v in d['roofline']['stages'].items()))
PY
