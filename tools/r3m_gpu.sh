timeout 1200 python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_spartan.py tests/test_gpu_multirank.py tests/test_gpu_verifier.py -m gpu -x -q 2>&1 | tail -4
for f in 1 0; do SP2_MID_FINISH=$f timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r3m_bench.json 2> gpurun_out/r3m_bench.err
python - <<PY
import json
for l in open("gpurun_out/r3m_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); print("finish=$f", round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()}, d["gpu_launches"])
PY
done
