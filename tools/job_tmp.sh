timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for c in A B C; do python tools/stage_times.py $c --bwd 2>&1 | tail -2; done
