timeout 1200 python -m pytest tests/test_gpu_spartan.py -m gpu -x -q 2>&1 | tail -4 | cut -c1-400
python __graft_entry__.py smoke 2>&1 | tail -2
