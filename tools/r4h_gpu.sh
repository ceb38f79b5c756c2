# r4h: commitment on the second side stream, outer sum-check pre-enqueued behind the host-opened gate
timeout 900 python -m pytest tests/test_gpu_spartan.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -5 | cut -c1-300
for i in 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4h_bench.json 2> gpurun_out/r4h_bench.err
  python - <<PY
import json
for l in open("gpurun_out/r4h_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("new", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done
SP2_PROVE_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep "sp2 prove" | tail -16
