set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sumcheck.py tests/test_gpu_spartan.py -m gpu -q -x 2>&1 | grep -E "^E   |Error|passed|failed|assert" | head -30 > gpurun_out/r2h_tests.log
cat gpurun_out/r2h_tests.log
for v in 1 0; do
SP2_TAIL_PIPE=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2h_bench_pipe$v.json 2> gpurun_out/r2h_bench.err
python - <<PY
import json
b=json.load(open('gpurun_out/r2h_bench_pipe$v.json'))
print("pipe=$v", round(b['ms_per_step'],4), round(b['e2e']['ms_per_step'],4), {k: round(x,3) for k,x in b['phase_ms'].items()})
PY
done
