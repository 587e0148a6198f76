timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in A C D; do python tools/stage_times.py $c --bwd 2>&1 | tail -2 | head -1; done
