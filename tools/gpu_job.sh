#!/bin/bash
# One gpurun job: GPU parity tests, bench (ours + reference arm), ncu launch list of the bench command,
# ncu --set full capture of one config-C view (every native kernel).
# Usage (from the repo root on the GPU box): bash tools/gpu_job.sh <tag> [noref]
TAG=${1:-r01}
K='regex:blend_|sweep_kernel|emit_kernel|preprocess_|scan_kernel|tile_ranges|fine_|tile_offsets'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
tail -3 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_ours.json 2> gpurun_out/${TAG}_bench_ours.err
tail -c 400 gpurun_out/${TAG}_bench_ours.json; echo
if [ "$2" != "noref" ]; then
  timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
  tail -c 300 gpurun_out/${TAG}_bench_reference.json; echo
fi
# launch list of the bench command itself (steady state: skip the first step's launches)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" --launch-skip 1600 --launch-count 400 --csv \
  --log-file gpurun_out/${TAG}_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu1.log 2>&1
# full capture: one forward+backward of config C, every native kernel once
timeout 1500 ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 23 --launch-count 23 \
  -f -o gpurun_out/${TAG}_full python tests/profile_step.py --impl ours --config C --iters 2 > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out | tail -12
