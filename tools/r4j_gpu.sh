# multi-GPU bench of the sharded partition: $1 = number of GPUs
N=$1
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r4j_bench_${N}gpu.json 2> gpurun_out/r4j_bench_${N}gpu.err
tail -2 gpurun_out/r4j_bench_${N}gpu.err
python - <<PY
import json
for l in open("gpurun_out/r4j_bench_${N}gpu.json"):
    if l.startswith("{"):
        d=json.loads(l); print({k:d[k] for k in ("value","n_gpus","ms_per_step","scaling")}, d.get("parity")); print(" sharded phases", {k:round(v,3) for k,v in d["phase_ms"].items()}); print(" single", d["single_gpu"]["ms_per_step"], {k:round(v,3) for k,v in d["single_gpu"]["phase_ms"].items()}); print(" nn5", d.get("neutronnova_config5"))
PY
