set -x
mkdir -p gpurun_out
./tools/microbench/field_mul > gpurun_out/r2b_microbench_field_mul.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2b_multirank.log
cat gpurun_out/r2b_microbench_field_mul.txt
tail -30 gpurun_out/r2b_multirank.log
