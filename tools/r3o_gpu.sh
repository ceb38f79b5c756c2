timeout 1200 python -m pytest tests/test_gpu_neutronnova.py tests/test_gpu_neutronnova_snark.py tests/test_gpu_multirank.py tests/test_gpu_small_value.py -m gpu -x -q 2>&1 | tail -3
python tools/nn_snark_time.py 32 256 2>&1 | grep snark_prove | cut -c1-420
