mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_spartan.py -m gpu -x -q 2>&1 | tail -15
for p in 1 0; do
SP2_MID_PIPE=$p timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_bench_mid$p.json 2> gpurun_out/r2n_bench.err
python - <<PY
import json
for l in open("gpurun_out/r2n_bench_mid$p.json"):
    if l.startswith("{"):
        d=json.loads(l); print("mid=$p", round(d["ms_per_step"],4), {k:round(v,3) for k,v in d["phase_ms"].items()}, d.get("parity"), d["gpu_launches"])
PY
done
tail -3 gpurun_out/r2n_bench.err
