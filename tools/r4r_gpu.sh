timeout 900 python -m pytest tests/test_gpu_neutronnova_snark.py tests/test_gpu_neutronnova.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -4 | cut -c1-400
python tools/nn_snark_time.py 32 256 2>&1 | grep snark_prove | cut -c1-520
