set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_msm_var.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2d_msmvar.log
timeout 900 python -m pytest tests/test_gpu_neutronnova_snark.py -m gpu -q 2>&1 | grep -E "^E   |AssertionError|passed|failed|parity|comm_" | head -40 > gpurun_out/r2d_snark.log
cat gpurun_out/r2d_msmvar.log | tail -25
cat gpurun_out/r2d_snark.log
