timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "forward_modes" 2>&1 | grep -E "ERROR SUMMARY|passed|failed|Invalid" | head
