#!/bin/bash
# One gpurun job: GPU parity tests, bench (ours), ncu launch list + full capture of one config-C view.
# Usage (from the repo root on the GPU box): bash tools/gpu_job.sh <tag> [skip-tests|ncu-only|quick]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if [ "$2" != "skip-tests" ] && [ "$2" != "ncu-only" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  tail -5 gpurun_out/${TAG}_pytest.log
fi
if [ "$2" != "ncu-only" ]; then
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err
  tail -c 600 gpurun_out/${TAG}_bench_ours.json
fi
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:blend_|sweep_kernel|emit_kernel|preprocess_|scan_kernel|tile_ranges|fine_|tile_offsets' --launch-skip 28 --launch-count 56 --csv \
  --log-file gpurun_out/${TAG}_launches_ours_configC.csv python tests/profile_step.py --impl ours --config C --iters 3 > gpurun_out/${TAG}_ncu1.log 2>&1
[ "$2" == "quick" ] && { ls -la gpurun_out; exit 0; }
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:blend_|sweep_kernel|emit_kernel|preprocess_|scan_kernel|tile_ranges|fine_|tile_offsets' --launch-skip 28 --launch-count 28 \
  -f -o gpurun_out/${TAG}_full python tests/profile_step.py --impl ours --config C --iters 2 > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out
