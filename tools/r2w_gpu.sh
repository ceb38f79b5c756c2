mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_r1cs.py tests/test_gpu_spartan.py tests/test_gpu_verifier.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err
python - <<PY
import json
for l in open("gpurun_out/r2w_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
