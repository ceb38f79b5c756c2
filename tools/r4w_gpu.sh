timeout 1500 python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_multirank.py tests/test_gpu_spartan.py tests/test_gpu_neutronnova.py -m gpu -x -q 2>&1 | tail -4 | cut -c1-400
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r4w_bench.json 2> gpurun_out/r4w_bench.err
python - <<PY
import json
for l in open("gpurun_out/r4w_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print(round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), d["tables_bench"])
PY
python tools/sc_round_profile.py 20 2>&1 | head -7
