#!/bin/bash
# Multi-GPU job (run under gpurun --gpus N): bench at N for several (graphs, streams) settings + the step timeline.
N=${1:-8}; TAG=${2:-r2m}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
mkdir -p gpurun_out
port=29510
for cfg in ${CFGS:-0:4 1:4 1:8}; do
  g=${cfg%%:*}; s=${cfg##*:}; port=$((port+1))
  timeout 600 $RUN $port bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --streams $s --graphs $g > gpurun_out/${TAG}_bench_n${N}_g${g}_s$s.json 2> gpurun_out/${TAG}_bench_n${N}_g${g}_s$s.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_n${N}_g${g}_s$s.json').read().strip().splitlines()[-1])
    print('N=$N graphs $g streams $s views/s', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'e2e ms', round(d['e2e']['ms_per_step'],3), 'parity', d['grad_parity_rel_l2'], d.get('cuda_graphs'))
    print('  collective', d['collective']); print('  e2e breakdown', {k:v for k,v in d['e2e']['breakdown'].items() if k!='note'})
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/${TAG}_bench_n${N}_g${g}_s$s.err').read()[-2000:])
PY
done
if [ "${TIMELINE:-1}" = "1" ]; then
timeout 300 $RUN 29531 tools/step_timeline.py --streams ${TL_STREAMS:-4} > gpurun_out/${TAG}_timeline_n${N}.md 2> gpurun_out/${TAG}_timeline.err
grep -v "memsets" gpurun_out/${TAG}_timeline_n${N}.md | cut -c1-200
fi
