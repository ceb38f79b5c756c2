SP2_NN_PIPE=1 timeout 1200 python -m pytest tests/test_gpu_neutronnova.py tests/test_gpu_neutronnova_snark.py -m gpu -x -q 2>&1 | tail -3
for p in 1 0; do SP2_NN_PIPE=$p python tools/nn_snark_time.py 32 256 2>&1 | tail -2 | cut -c1-420 | sed "s/^/pipe=$p /"; done
