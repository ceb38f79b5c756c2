timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_spartan.py -m gpu -x -q 2>&1 | grep -v "^$" | grep "Error\|errs\|passed\|failed" | head -8 | cut -c1-900
