mkdir -p gpurun_out
for p in 0 1; do
SP2_TAIL_PIPE=$p python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2k_bench_pipe$p.json 2> gpurun_out/r2k_bench.err
python - <<PY
import json
for l in open("gpurun_out/r2k_bench_pipe$p.json"):
    if l.startswith("{"):
        d=json.loads(l); print("pipe=$p", round(d["ms_per_step"],4), round(d["e2e"]["ms_per_step"],4) if "ms_per_step" in d["e2e"] else d["e2e"], {k:round(v,3) for k,v in d["phase_ms"].items()})
PY
done
SP2_TAIL_PIPE=0 python tools/sc_round_profile.py 20 2>&1 | tail -12
SP2_TAIL_PIPE=0 python tools/sc_clocks.py 2>&1 | tail -6
