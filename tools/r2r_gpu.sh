timeout 900 python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_spartan.py -m gpu -x -q 2>&1 | tail -4
python tools/mid_trace.py 20 2>&1 | tail -9
python tools/tail_trace.py 16 2>&1 | tail -11
python tools/sc_round_profile.py 20 2>&1 | tail -12
