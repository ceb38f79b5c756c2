timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | cut -c1-400
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4u_bench.json 2> gpurun_out/r4u_bench.err
python - <<PY
import json
for l in open("gpurun_out/r4u_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["gpu_launches"], {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done
python tools/nn_snark_time.py 32 2>&1 | grep snark_prove | cut -c1-200
