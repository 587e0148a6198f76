timeout 1500 python tools/bench_configs.py --configs A,B,C,D --views 120 --stacks 8,16 > gpurun_out/r2F_configs.jsonl 2> gpurun_out/r2F_configs.err
cut -c1-2200 gpurun_out/r2F_configs.jsonl; tail -3 gpurun_out/r2F_configs.err
