mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_spartan.py -m gpu -x -q 2>&1 | tail -5
python tools/tail_trace.py 16 2>&1 | tail -12
for p in 1 0; do
SP2_TAIL_PIPE=$p python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_pipe$p.json 2> gpurun_out/r2l_bench.err
python - <<PY
import json
for l in open("gpurun_out/r2l_bench_pipe$p.json"):
    if l.startswith("{"):
        d=json.loads(l); print("pipe=$p", round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done
