for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q -k "sharded_spartan" 2>&1 | grep -v "^$" | tail -25 | cut -c1-600; done
