# r4b: early comm_LZ / delta on side streams — parity, then same-box A/B against the round-start library
timeout 900 python -m pytest tests/test_gpu_spartan.py tests/test_gpu_verifier.py tests/test_gpu_sumcheck.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -5
for i in 1 2; do for lib in new old; do
  if [ $lib = old ]; then export SP2_LIB_PATH=$PWD/libold_r4a.so; else unset SP2_LIB_PATH; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4b_bench_$lib.json 2> gpurun_out/r4b_bench_$lib.err
  python - <<PY
import json
for l in open("gpurun_out/r4b_bench_$lib.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$lib", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done; done
unset SP2_LIB_PATH
