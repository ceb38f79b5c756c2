mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_neutronnova.py tests/test_gpu_neutronnova_snark.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -6
for p in 1 0; do SP2_NN_PIPE=$p python tools/nn_snark_time.py 32 256 2>&1 | tail -2 | sed "s/^/pipe=$p /"; done
