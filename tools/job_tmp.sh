timeout 900 python -m pytest tests -m gpu -x -q -k "stacked or render_views" 2>&1 | tail -3
timeout 900 python tools/bench_configs.py --configs B,D --views 120 --stacks 8,16 --no-ref 2>&1 | cut -c1-1800
