# r4n: derive via the pinned mailbox: A/B, then the order-dependent multirank failure with its message
for i in 1 2; do for d in 0 1; do
  SP2_NO_DERIVE=$d timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4n_bench_$d.json 2> gpurun_out/r4n_bench_$d.err
  python - <<PY
import json
for l in open("gpurun_out/r4n_bench_$d.json"):
    if l.startswith("{"):
        d=json.loads(l); print("no_derive=$d", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()}, round(d["roofline"]["frac"],3), round(d["roofline"]["ms"],4))
PY
done; done
tail -2 gpurun_out/r4n_bench_0.err | cut -c1-500
timeout 900 python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_spartan.py tests/test_gpu_verifier.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_spartan.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | grep -v "^$" | grep "Error\|errs\|passed\|failed" | head -12 | cut -c1-700
