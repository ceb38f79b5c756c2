# r4l: t(1) derived from the claim in the streaming rounds (+ the one-product first inner round): parity, A/B against SP2_NO_DERIVE=1
timeout 1500 python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_spartan.py tests/test_gpu_verifier.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -5 | cut -c1-300
for i in 1 2; do for d in 0 1; do
  SP2_NO_DERIVE=$d timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r4l_bench_$d.json 2> gpurun_out/r4l_bench_$d.err
  python - <<PY
import json
for l in open("gpurun_out/r4l_bench_$d.json"):
    if l.startswith("{"):
        d=json.loads(l); print("no_derive=$d", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()}, round(d["roofline"]["frac"],3), round(d["roofline"]["ms"],4))
PY
done; done
tail -3 gpurun_out/r4l_bench_0.err
