mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_spartan.py -m gpu -x -q 2>&1 | tail -4
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
python - <<PY
import json
for l in open("gpurun_out/r2x_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()}, d["roofline"].get("int_pipe"))
PY
done
python tools/sc_round_profile.py 20 2>&1 | tail -8
