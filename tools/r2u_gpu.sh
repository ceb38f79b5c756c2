mkdir -p gpurun_out
for i in 1 2; do
for v in prev cur; do
if [ $v = prev ]; then export SP2_LIB_PATH=$PWD/spartan2_b200/libprev.so; else unset SP2_LIB_PATH; fi
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_$v.json 2> gpurun_out/r2u_bench.err
python - <<PY
import json
for l in open("gpurun_out/r2u_bench_$v.json"):
    if l.startswith("{"):
        d=json.loads(l); print("$v", round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()}, d["clocks"])
PY
done
done
