# A/B: quadratic prover's pipelined kernel starting one round earlier (2^18-entry tables going in)
SP2_LIB_PATH=$PWD/libalt_q18.so timeout 900 python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_spartan.py -m gpu -x -q 2>&1 | tail -3 | cut -c1-300
for i in 1 2; do for lib in alt base; do
  if [ $lib = alt ]; then export SP2_LIB_PATH=$PWD/libalt_q18.so; else unset SP2_LIB_PATH; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4x_bench.json 2> gpurun_out/r4x_bench.err
  python - <<PY
import json
for l in open("gpurun_out/r4x_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$lib", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done; done
