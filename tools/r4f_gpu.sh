# r4f: PCS tail split over two streams, eq(r_x) beside outer->inner, long-column chunks beside k_abc
timeout 900 python -m pytest tests/test_gpu_spartan.py tests/test_gpu_verifier.py tests/test_gpu_multirank.py tests/test_gpu_r1cs.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4f_bench.json 2> gpurun_out/r4f_bench.err
  python - <<PY
import json
for l in open("gpurun_out/r4f_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("new", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done
SP2_PROVE_TRACE=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras 2>&1 >/dev/null | grep "sp2 prove" | tail -11
