# r4d: head hashing on the helper thread — parity + bench (phase times are the indicator: commit_transcript was 0.201 ms)
timeout 900 python -m pytest tests/test_gpu_spartan.py tests/test_gpu_verifier.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2 3; do
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4d_bench.json 2> gpurun_out/r4d_bench.err
  python - <<PY
import json
for l in open("gpurun_out/r4d_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("new", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done
