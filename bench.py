#!/usr/bin/env python
"""bench.py — the Spartan prover hot path on B200.

  python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port, all host threads)

One JSON line on stdout (rank 0).  A "step" is one pass of the prover hot path over one synthetic instance of
the workload named in config.workload.  `value` = field-ops/s with inputs resident in HBM; `e2e` = the same
through the C-ABI with HOST buffers (pinned), copies inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sha256_r1cs_prove_field_ops_per_sec"
UNIT = "field-ops/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def rand_fe(rng, n):
    a = rng.integers(0, 2**64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64(0x7fffffffffffffff)
    return a


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device; self.proc = None; self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True); self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# workload: the two sum-checks of SpartanSNARK::prove at the 2 KiB SHA-256 shape (N = M = 2^20)
# ------------------------------------------------------------------------------------------------
class SumcheckWorkload:
    """outer cubic sum-check over Az,Bz,Cz (2^l) + inner quadratic sum-check over (ABC, z) (2^m).
    Algorithmic bytes / field-ops per SURVEY.md §8(d): cubic 368*T B and 7*T ops, quad 256*T B and 4*T ops."""

    def __init__(self, l, m, seed=0xDEADBEEF):
        self.l, self.m = l, m
        self.N, self.M = 1 << l, 1 << m
        self.name = "spartan_sumchecks_N2^%d_M2^%d" % (l, m)
        self.field_ops = 7 * self.N + 4 * self.M
        self.bytes_cubic = 368 * self.N
        self.bytes_quad = 256 * self.M
        rng = np.random.default_rng(seed)
        self.taus = rand_fe(rng, l)
        self.rng = rng

    def host_tables(self, alloc):
        A, B, C, X, Y = alloc((self.N, 4)), alloc((self.N, 4)), alloc((self.N, 4)), alloc((self.M, 4)), alloc((self.M, 4))
        for t in (A, B, C, X, Y):
            t[:] = rand_fe(self.rng, t.shape[0])
        return A, B, C, X, Y


def run_cuda(args):
    import torch
    import spartan2_b200 as sp
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ctx = sp.Context(local)
    wl = SumcheckWorkload(args.log_n, args.log_n, seed=0xDEADBEEF + rank)
    hbm_peak, peak_kind = peaks()
    A, B, C, X, Y = wl.host_tables(ctx.pinned_empty)
    zero = np.zeros((1, 4), dtype=np.uint64)
    # pristine copies in HBM + working copies (the provers bind in place)
    pr = [ctx.upload(t) for t in (A, B, C, X, Y)]
    wk = [ctx.alloc(t.nbytes) for t in (A, B, C, X, Y)]
    flush = ctx.alloc(512 << 20)                     # > 126 MB L2

    def restore():
        for d, s in zip(wk, pr):
            d.copy_from(s)
        ctx.check(ctx.L.sp2_dev_memset(ctx.h, flush.ptr, 0, 512 << 20))   # flush L2 between timed iterations
        ctx.synchronize()

    def barrier():
        ctx.synchronize(); torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    def step_device():
        ts = sp.TranscriptState()
        ctx.timer_start()
        # device time of both sum-checks (kernels only; the tiny state download is after the timer)
        sp.SumcheckProof.prove_cubic_with_three_inputs(ctx, zero, wl.taus, wk[0], wk[1], wk[2], ts)
        t1 = ctx.timer_stop()
        ctx.timer_start()
        sp.SumcheckProof.prove_quad(ctx, zero, wl.m, wk[3], wk[4], ts)
        t2 = ctx.timer_stop()
        return t1, t2

    def step_e2e():
        ts = sp.TranscriptState()
        t0 = time.perf_counter()
        sp.SumcheckProof.prove_cubic_with_three_inputs(ctx, zero, wl.taus, A, B, C, ts)
        sp.SumcheckProof.prove_quad(ctx, zero, wl.m, X, Y, ts)
        return (time.perf_counter() - t0) * 1e3

    for _ in range(max(args.warmup, 3)):
        restore(); step_device()
    l0 = ctx.launch_count()
    sampler = ClockSampler(local); sampler.start()
    barrier()
    t_c, t_q = [], []
    for _ in range(args.steps):
        restore()
        a, b = step_device()
        t_c.append(a); t_q.append(b)
    barrier()
    launches = ctx.launch_count() - l0
    ms_dev = float(np.mean(t_c) + np.mean(t_q))
    # e2e: host (pinned) tables -> H2D -> prove -> results back on the host, wall clock around the ABI calls
    step_e2e()
    barrier()
    e2e_ms = float(np.mean([step_e2e() for _ in range(args.steps)]))
    barrier()
    clocks = sampler.stop()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_dev, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, e2e_ms = float(t[0]), float(t[1])
    out = None
    if rank == 0:
        ach = wl.bytes_cubic / (float(np.mean(t_c)) * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": world * wl.field_ops / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (256-bit prime field, Montgomery)", "data": "synthetic",
            "config": {"workload": wl.name, "phases": ["outer_sumcheck", "inner_sumcheck"], "l2": "flushed between timed iterations (512 MiB memset)",
                       "field_ops_per_step": wl.field_ops, "replicas": world},
            "e2e": {"value": world * wl.field_ops / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(A.nbytes * 3 + X.nbytes * 2), "d2h_bytes_per_step": int((wl.l * 4 + wl.l + 3 + wl.m * 3 + wl.m + 2) * 32)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_cubic_round (outer sum-check, all rounds)", "achieved": ach, "peak": hbm_peak,
                         "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None, "peak_kind": peak_kind,
                         "ms": float(np.mean(t_c)), "algorithmic_bytes": wl.bytes_cubic},
            "phase_ms": {"outer_sumcheck": float(np.mean(t_c)), "inner_sumcheck": float(np.mean(t_q))},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(sample_log_n=min(args.log_n, 16), threads=1)
        print(json.dumps(out), flush=True)
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def cpu_baseline(sample_log_n, threads, reps=1):
    """The oracle port of the reference's CPU algorithm on a bounded sample of the same workload."""
    from oracle import pyoracle as orc
    orc.lib(native=True)
    orc.set_threads(threads)
    wl = SumcheckWorkload(sample_log_n, sample_log_n)
    A, B, C, X, Y = wl.host_tables(lambda s: np.zeros(s, dtype=np.uint64))
    zero = np.zeros((1, 4), dtype=np.uint64)
    best = None
    for _ in range(reps + 1):
        t = orc.Transcript(b"bench")
        t0 = time.perf_counter()
        orc.sumcheck_cubic_prove(zero, wl.taus, A, B, C, t)
        orc.sumcheck_quad_prove(zero, wl.m, X, Y, t)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": wl.field_ops / best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%s (oracle/oracle.c, -march=native, %d thread(s)), %.1f ms" % (wl.name, threads, best * 1e3)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as orc
    orc.lib(native=True)
    threads = orc.max_threads()
    orc.set_threads(threads)
    wl = SumcheckWorkload(args.log_n, args.log_n)
    A, B, C, X, Y = wl.host_tables(lambda s: np.zeros(s, dtype=np.uint64))
    zero = np.zeros((1, 4), dtype=np.uint64)

    def step():
        t = orc.Transcript(b"bench")
        t0 = time.perf_counter()
        orc.sumcheck_cubic_prove(zero, wl.taus, A, B, C, t)
        orc.sumcheck_quad_prove(zero, wl.m, X, Y, t)
        return (time.perf_counter() - t0) * 1e3
    for _ in range(args.warmup):
        step()
    ms = float(np.mean([step() for _ in range(args.steps)]))
    v = wl.field_ops / (ms * 1e-3)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64x4 (256-bit prime field, Montgomery)",
        "data": "synthetic", "config": {"workload": wl.name, "phases": ["outer_sumcheck", "inner_sumcheck"], "field_ops_per_step": wl.field_ops},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "full workload, oracle/oracle.c (C restatement of the reference's rayon prover; no Rust toolchain), OpenMP %d threads" % threads},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--log-n", type=int, default=20, help="log2 of the padded constraint count (2 KiB SHA-256: 20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
