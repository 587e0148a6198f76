#!/bin/bash
# Round-2 GPU job: parity tests, both bench arms, sort vs CUB.  Usage: bash tools/gpu_r2.sh <tag> [steps...]
# steps: tests bench ref cub configs launches ncu   (default: tests bench ref cub)
TAG=${1:-r02}; shift
STEPS=${@:-tests bench ref cub}
K='regex:blend_|sweep_kernel|emit_kernel|preprocess_|scan_kernel|tile_ranges|fine_|tile_offsets|radix_|bin_|depth_'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
for s in $STEPS; do
case $s in
tests)
  timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  tail -5 gpurun_out/${TAG}_pytest.log ;;
bench)
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench_ours.json').read().strip().splitlines()[-1])
    print('views/s', round(d['value'],1), 'ms/view', round(d['ms_per_view'],4), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'], 'grad_parity', d.get('grad_parity_rel_l2'))
    print(' '.join(f"{k}={v['ms']}({v['frac']})" for k,v in d['roofline']['stages'].items()))
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/${TAG}_bench_ours.err').read()[-1500:])
PY
  ;;
ref)
  timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
  tail -c 900 gpurun_out/${TAG}_bench_reference.json; echo; tail -3 gpurun_out/${TAG}_bench_reference.err ;;
cub)
  timeout 300 bloomscene_b200/_build/sort_vs_cub 50 > gpurun_out/${TAG}_sort_vs_cub.jsonl 2> gpurun_out/${TAG}_sort_vs_cub.err
  cat gpurun_out/${TAG}_sort_vs_cub.jsonl; tail -2 gpurun_out/${TAG}_sort_vs_cub.err ;;
configs)
  timeout 1500 python tools/bench_configs.py > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err
  cat gpurun_out/${TAG}_configs.jsonl | cut -c1-600; tail -3 gpurun_out/${TAG}_configs.err ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 1200 --launch-count 400 --csv \
    --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu1.log ;;
ncu)
  timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip ${NCU_SKIP:-13} --launch-count ${NCU_COUNT:-13} \
    -f -o gpurun_out/${TAG}_full python tests/profile_step.py --impl ours --config C --iters 2 > gpurun_out/${TAG}_ncu2.log 2>&1
  tail -2 gpurun_out/${TAG}_ncu2.log ;;
esac
done
ls -la gpurun_out | grep ${TAG} | tail -12
