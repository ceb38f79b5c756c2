for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -40 | cut -c1-400; done
