set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
nproc >> gpurun_out/r2a_gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r2a_gpu.txt
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -60 > gpurun_out/r2a_multirank.log
timeout 1200 python -m pytest tests -m gpu -q --deselect tests/test_gpu_multirank.py > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_bench_ref.json 2>> gpurun_out/r2a_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2a_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^(k_quad_persist|k_msm_gather|k_msm_final|k_abc|k_abc_long|k_cubic_tail|k_quad_tail|k_spmv3|k_hyrax_bind)' -s 4 -c 22 -o /tmp/r2a_prove_kernels python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2a_ncu1.log 2>&1
ncu -i /tmp/r2a_prove_kernels.ncu-rep --page raw --csv > gpurun_out/r2a_prove_kernels_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^(k_nn_outer_round|k_nn_inner_round|k_nifs_round|k_nifs_fold_v|k_fold_vectors|k_nifs_round0_small)' -c 16 -o /tmp/r2a_nn_kernels python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_ncu2.log 2>&1
ncu -i /tmp/r2a_nn_kernels.ncu-rep --page raw --csv > gpurun_out/r2a_nn_kernels_raw.csv 2>/dev/null
ls -la gpurun_out | tail -20
cat gpurun_out/r2a_multirank.log | tail -40
tail -5 gpurun_out/r2a_tests.log
