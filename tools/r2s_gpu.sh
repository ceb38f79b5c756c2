mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python - <<PY
import json
for l in open("gpurun_out/r2s_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],4), round(d["e2e"].get("ms_per_step",0),4), {k:round(v,3) for k,v in d["phase_ms"].items()}, d["gpu_launches"], round(d["roofline"]["frac"],3))
PY
done
python tools/neutronnova_path.py 2>&1 | tail -5
