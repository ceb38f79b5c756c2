timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/nn_snark_time.py 32 256 2>&1 | grep snark_prove | cut -c1-420
