mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for p in 1 0; do
SP2_MID_PIPE=$p timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2q_bench_mid$p.json 2> gpurun_out/r2q_bench.err
python - <<PY
import json
for l in open("gpurun_out/r2q_bench_mid$p.json"):
    if l.startswith("{"):
        d=json.loads(l); print("mid=$p", round(d["ms_per_step"],4), round(d["e2e"].get("ms_per_step",0),4), {k:round(v,3) for k,v in d["phase_ms"].items()}, d["gpu_launches"], d["roofline"]["frac"])
PY
done
