timeout 1200 python -m pytest tests/test_gpu_spartan.py tests/test_gpu_verifier.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3 | cut -c1-400
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4s_bench.json 2> gpurun_out/r4s_bench.err
python - <<PY
import json
for l in open("gpurun_out/r4s_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["roofline"]["traffic"])
PY
done
