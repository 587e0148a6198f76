for s in 1 2 3 4; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams $s > gpurun_out/r01k_s$s.json 2> gpurun_out/r01k_s$s.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r01k_s$s.json').read().strip().splitlines()[-1])
print('streams $s views/s', round(d['value'],1), 'ms/view', round(d['ms_per_view'],4), 'e2e', round(d['e2e']['value'],1), 'loss', d['loss'])
PY
done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
