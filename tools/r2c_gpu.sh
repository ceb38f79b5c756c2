set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_neutronnova_snark.py tests/test_gpu_msm.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2c_snark.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_neutronnova_snark.py > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
tail -40 gpurun_out/r2c_snark.log
tail -8 gpurun_out/r2c_tests.log
python - <<'PY'
import json
b=json.load(open('gpurun_out/r2c_bench.json'))
print(b['ms_per_step'], b['e2e']['ms_per_step'], b['phase_ms'], b['neutronnova']['prove_ms'])
PY
