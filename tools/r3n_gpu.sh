for i in 1 2 3; do for f in 1 0; do SP2_MID_FINISH=$f timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r3n_bench.json 2> gpurun_out/r3n_bench.err
python - <<PY
import json
for l in open("gpurun_out/r3n_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("finish=$f", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done; done
