SP2_NN_PIPE=1 timeout 1200 python -m pytest tests/test_gpu_neutronnova.py tests/test_gpu_neutronnova_snark.py -m gpu -x -q 2>&1 | tail -3
timeout 1200 python -m pytest tests/test_gpu_neutronnova.py tests/test_gpu_neutronnova_snark.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do for p in 1 0; do SP2_NN_PIPE=$p python tools/nn_snark_time.py 32 256 2>&1 | grep snark_prove | cut -c1-100 | sed "s/^/pipe=$p /"; done; done
