# r4e: host-side timeline of the prove (SP2_PROVE_TRACE=1) + NeutronNova with the hoisted delta MSM
SP2_PROVE_TRACE=1 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4e_bench.json 2> gpurun_out/r4e_trace.txt
grep "sp2 prove" gpurun_out/r4e_trace.txt | tail -26
timeout 900 python -m pytest tests/test_gpu_neutronnova_snark.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
python tools/nn_snark_time.py 32 2>&1 | grep snark_prove | cut -c1-420
