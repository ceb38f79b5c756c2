# r4a: lane-quad MSM trees — parity tests, then same-box A/B of the prove against the previous library
timeout 900 python -m pytest tests/test_gpu_msm.py tests/test_gpu_hyrax_family.py tests/test_gpu_spartan.py tests/test_gpu_verifier.py -m gpu -x -q 2>&1 | tail -5
for i in 1 2; do for lib in new old; do
  if [ $lib = old ]; then export SP2_LIB_PATH=$PWD/libold_r4a.so; else unset SP2_LIB_PATH; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4a_bench_$lib.json 2> gpurun_out/r4a_bench_$lib.err
  python - <<PY
import json
for l in open("gpurun_out/r4a_bench_$lib.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$lib", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done; done
unset SP2_LIB_PATH
timeout 900 python -m pytest tests/test_gpu_neutronnova_snark.py tests/test_gpu_neutronnova.py -m gpu -x -q 2>&1 | tail -3
python tools/nn_snark_time.py 32 2>&1 | grep snark_prove | cut -c1-420
