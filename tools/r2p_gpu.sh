timeout 600 python -m pytest tests/test_gpu_sumcheck.py -m gpu -x -q 2>&1 | tail -4
python tools/mid_trace.py 20 2>&1 | tail -10
python tools/sc_round_profile.py 20 2>&1 | tail -12
