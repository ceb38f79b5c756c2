# r4c: whole GPU suite after the PCS restructure (sharded path without the spinner), bench A/B
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for i in 1 2; do for lib in new old; do
  if [ $lib = old ]; then export SP2_LIB_PATH=$PWD/libold_r4a.so; else unset SP2_LIB_PATH; fi
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4c_bench_$lib.json 2> gpurun_out/r4c_bench_$lib.err
  python - <<PY
import json
for l in open("gpurun_out/r4c_bench_$lib.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$lib", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done; done
unset SP2_LIB_PATH
