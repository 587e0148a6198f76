# bench with different stream counts (diagnostic)
for s in ${STREAMS:-1 2 4 6}; do
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams $s > gpurun_out/${TAG:-sj}_s$s.json 2> gpurun_out/${TAG:-sj}_s$s.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG:-sj}_s$s.json').read().strip().splitlines()[-1])
print('streams $s views/s', round(d['value'],1), 'ms/view', round(d['ms_per_view'],4), 'e2e', round(d['e2e']['value'],1), 'loss', d['loss'])
PY
done
