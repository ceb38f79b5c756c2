for i in 1 2; do for m in 0 1; do
SP2_MID_PLAIN=$m timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4t_bench.json 2> gpurun_out/r4t_bench.err
python - <<PY
import json
for l in open("gpurun_out/r4t_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("plain=$m", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done; done
tail -2 gpurun_out/r4t_bench.err | cut -c1-300
