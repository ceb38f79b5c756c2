set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_verifier.py -m gpu -q -x 2>&1 | grep -E "^E   |Error|passed|failed|parity" | head -30 > gpurun_out/r2g_tests.log
cat gpurun_out/r2g_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2g_bench2.json 2> gpurun_out/r2g_bench2.err; echo "bench2 rc=$?"
tail -5 gpurun_out/r2g_bench2.err
python - <<'PY'
import json
try:
    b=json.load(open('gpurun_out/r2g_bench2.json'))
    print({k: b[k] for k in ('value','ms_per_step','scaling','n_gpus','phase_ms','parity')})
    print('single', b.get('single_gpu'))
    print('nn5', json.dumps(b.get('neutronnova_config5'))[:1500])
except Exception as e:
    print("ERR", e); print(open('gpurun_out/r2g_bench2.json').read()[:2000])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2g_bench2_ref.json 2>> gpurun_out/r2g_bench2.err
head -c 600 gpurun_out/r2g_bench2_ref.json
