timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python tools/bench_fused.py 2>&1 | tee gpurun_out/r2t_bench_fused.jsonl
python tools/bench_configs.py --views 120 --configs B,D > gpurun_out/r2t_configs_BD_120.jsonl 2>gpurun_out/r2t_configs.err; cut -c1-700 gpurun_out/r2t_configs_BD_120.jsonl
