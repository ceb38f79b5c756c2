mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2v_bench_2gpu.json 2> gpurun_out/r2v_bench_2gpu.err
tail -3 gpurun_out/r2v_bench_2gpu.err
python - <<'PY'
import json
for l in open("gpurun_out/r2v_bench_2gpu.json"):
    if l.startswith("{"):
        d=json.loads(l); print({k:d[k] for k in ("metric","value","n_gpus","ms_per_step","scaling")}, d.get("parity"), d.get("single_gpu"), d["config"].get("workload"))
PY
