# final numbers of the round: bench (both arms), host trace, launch list
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2zz_bench.json 2> gpurun_out/r2zz_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2zz_bench_reference.json 2>> gpurun_out/r2zz_bench.err
SP2_NO_GATES=1 timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2zz_launches_full_prove.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/r2zz_ncu_bench.log 2>&1
SP2_PROVE_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep "sp2 prove" | tail -17 > gpurun_out/r2zz_prove_host_trace.txt
python tools/nn_snark_time.py 32 256 2>&1 | tail -2 > gpurun_out/r2zz_neutronnova.log
tail -2 gpurun_out/r2zz_bench.err | cut -c1-300
