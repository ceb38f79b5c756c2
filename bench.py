#!/usr/bin/env python
"""bench.py — SpartanSNARK::prove of the SHA-256 circuit on B200 (BASELINE.json config 2).

  python bench.py --gpus N --steps K --warmup W            # CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port, all host threads)

One JSON line on stdout (rank 0).  A "step" is one `prove` (src/spartan.rs:219-466) of sha256_spartan after prep_prove,
as the reference's bench times it (benches/sha256_spartan.rs:224-244):
  --gpus 1   BASELINE config 2: the 2 KiB all-zero message (benches/sha256_spartan.rs:167,200), N = M = 2^20, one GPU;
  --gpus N>1 BASELINE config 4: ONE proof of the 8 KiB message (N = M = 2^22) with its hypercube sharded across the N
             GPUs (north_star's partition; "scaling": "strong"), plus BASELINE config 5 (sha256_neutronnova, 256 step
             circuits, instance-sharded) as the `neutronnova_config5` object.  `--replicas` keeps the old one-proof-per-
             GPU mode (weak scaling, no data-path exchange).
The reference arm (`--impl reference`) runs the same config as the CUDA arm at the same --gpus.
`value` = field-ops/s with the witness and prep state resident in HBM (device timeline of the prove, CUDA events);
`e2e` = the same prove through the C ABI from HOST buffers, wall clock, copies and host syncs included.
field-op = one 256-bit modular multiplication-equivalent of the reference's algorithm, MSM work
excluded from the count (its time is included in the step).
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
WIDTH = 2048          # DEFAULT_COMMITMENT_WIDTH (src/lib.rs:63)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured"
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
        self.device = device; self.proc = None; self.lines = []; self.first = 0

    def mark(self):
        """the timed region starts here (the sampler is started before the warm-up steps: nvidia-smi takes up to a second to come up)"""
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
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
        window = "timed steps"
        lines = self.lines[self.first:]
        if not lines:
            lines = self.lines; window = "warm-up + timed steps (the timed region was shorter than one nvidia-smi period)"
        for ln in lines:
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
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def persist_end(l, tail_len, mid_len):
    """first round NOT run by the persistent multi-CTA kernel (sumcheck.cu: sumcheck_cubic_enqueue / mid_plan): the single-CTA tail
    takes tables of <= tail_len entries; with the pipelined multi-CTA kernels enabled, rounds >= 2 whose input tables have <= mid_len
    entries go to k_cubic_mid_pipe (when there are at least two of them before the tail)"""
    r_end = 1
    while r_end <= l and ((4 if r_end > 1 else 2) << (l - r_end)) > tail_len:
        r_end += 1
    if mid_len:
        rm = 2
        while rm <= l and (4 << (l - rm)) > mid_len:
            rm += 1
        if rm + 1 < r_end:
            return rm
    return r_end


class Workload:
    """sha256_spartan: Sha256Circuit(vec![0u8; msg_len]) -> padded R1CS, witness, prover randomness (seeded)."""

    def __init__(self, msg_len, seed=0xDEADBEEF, tail_len=1024, mid_len=0):
        from spartan2_b200.frontend import Sha256Circuit
        self.msg_len = msg_len
        self.circ = c = Sha256Circuit(b"\x00" * msg_len, width=WIDTH)
        self.name = "sha256_spartan_%dB_zero_message" % msg_len
        self.A, self.B, self.C = c.matrices()
        self.W, self.X = c.witness()
        self.N, self.M = c.num_cons, c.num_vars
        self.rows = self.M // WIDTH; self.cached_len = c.num_precommitted; self.cached_rows = self.cached_len // WIDTH
        rng = np.random.default_rng(seed)
        self.blinds = rand_fe(rng, self.rows); self.blind_eval = rand_fe(rng, 1); self.d_vec = rand_fe(rng, WIDTH)
        self.r_delta = rand_fe(rng, 1); self.r_beta = rand_fe(rng, 1)
        self.vk = bytes(32)
        # general-coefficient entries (one modmul each in SpMV / ABC; everything else is adds)
        is_gen = np.array([abs(v) != 1 for v in c.coef_values], dtype=bool)
        gen = [int(is_gen[co].sum()) for (co, _, _) in c.raw]
        self.nnz = sum(c.nnz); self.nnz_general = sum(gen)
        N, M = self.N, self.M
        # outer cubic 7N; eq table N; ABC 2N + general nnz; inner 4M + manual round 0 (3M);
        # Hyrax bind M; incremental SpMV: general nnz of the filtered columns (bounded by nnz_general; counted as 0)
        self.field_ops = 7 * N + N + 2 * N + self.nnz_general + 7 * M + M
        self.bytes_outer = 368 * N
        # the persistent kernel k_cubic_persist runs rounds 1..R-1 of the outer sum-check (tables > tail_len entries going in,
        # sumcheck.cu: sumcheck_cubic_enqueue): round 1 reads 2.5 tables, round i >= 2 reads 3 * T/2^(i-2) entries and
        # writes half as many (bind fused into the evaluation)  — SURVEY.md §8(d) accounting, 32 B per entry
        l = N.bit_length() - 1
        r_end = persist_end(l, tail_len, mid_len)
        self.persist_rounds = r_end - 1
        self.bytes_persist = (80 * N if r_end > 1 else 0) + sum(144 * (N >> (i - 2)) for i in range(2, r_end))

    def describe(self):
        c = self.circ
        return {"workload": self.name, "num_cons": c.num_cons_unpadded, "N": self.N, "M": self.M, "nnz": self.nnz, "nnz_general": self.nnz_general,
                "hyrax_rows": self.rows, "hyrax_width": WIDTH, "field_ops_per_step": self.field_ops,
                "phases": ["commit_rest+transcript", "matrix_vector_multiply", "outer_sumcheck", "prepare_poly_ABC", "inner_sumcheck", "pcs_prove", "ipa_response"]}


def default_msg_len(args, world):
    """--gpus 1: BASELINE config 2 (2 KiB message); --gpus N > 1: config 4 (8 KiB message, N = M = 2^22), sharded."""
    if args.msg_len:
        return args.msg_len
    return 2048 if (world == 1 or args.replicas) else 8192


def run_cuda(args):
    import ctypes as C
    import torch
    import spartan2_b200 as sp
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    ctx = sp.Context(local)
    # N > 1: ONE proof, its 2^l hypercube split across the GPUs (strong scaling);
    # --replicas: every GPU proves its own instance (independent proofs: weak scaling, no data-path exchange)
    sharded = world > 1 and not args.replicas
    wl = Workload(default_msg_len(args, world), seed=0xDEADBEEF + (0 if sharded else rank), tail_len=int(ctx.L.sp2_sc_tail_len()), mid_len=int(ctx.L.sp2_sc_mid_len()))
    hbm_peak, peak_kind = peaks()
    pts = ctx.test_points(WIDTH + 3, seed=7)
    K = sp.CommitmentKey(ctx, pts[:WIDTH], pts[WIDTH:WIDTH + 1], pts[WIDTH + 1:WIDTH + 2], pts[WIDTH + 2:WIDTH + 3])
    comm = None

    def allgather_bytes(b):
        import torch.distributed as dist
        out = [None] * world
        dist.all_gather_object(out, b)
        return out
    if sharded:
        comm = sp.Comm(ctx, rank, world, allgather_bytes)
        S = sp.SplitR1CSShape(ctx, *wl.circ.dims(), wl.A, wl.B, wl.C, rank=rank, nranks=world)
    else:
        S = sp.SplitR1CSShape(ctx, *wl.circ.dims(), wl.A, wl.B, wl.C)
    t0 = time.perf_counter()
    prep = sp.SpartanSNARK.prep_prove(ctx, S, K, wl.W[:wl.cached_len], wl.blinds[:wl.cached_rows], is_small=True)
    prep_ms = (time.perf_counter() - t0) * 1e3
    # the circuit allocates nothing in `synthesize` (benches/sha256_spartan.rs:140-148): the rest section is pure zero
    # padding, which the ABI takes as NULL
    assert not wl.W[wl.cached_len:].any()
    W_rest = None
    flush = ctx.alloc(512 << 20)                     # > 126 MB L2

    def barrier():
        ctx.synchronize(); torch.cuda.synchronize()
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    def step(shape, prp, cm):
        ctx.check(ctx.L.sp2_dev_memset(ctx.h, flush.ptr, 0, 512 << 20))   # flush L2 between timed iterations
        ctx.synchronize()
        if cm is not None:
            import torch.distributed as dist
            dist.barrier()
        t0 = time.perf_counter()
        proof = sp.SpartanSNARK.prove(ctx, shape, K, prp, wl.vk, wl.X, W_rest, wl.blinds, wl.blind_eval, wl.d_vec, wl.r_delta, wl.r_beta, comm=cm)
        wall = (time.perf_counter() - t0) * 1e3
        return proof, wall

    W_ = max(args.warmup, 3)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(W_):
        step(S, prep, comm)
    l0 = ctx.launch_count()
    barrier()
    sampler.mark()
    dev, wall, phases, persist = [], [], [], []
    for _ in range(args.steps):
        proof, w = step(S, prep, comm)
        dev.append(proof.phase_ms["total"]); wall.append(w); phases.append(proof.phase_ms)
        ms = C.c_float()
        if not sharded and ctx.L.sp2_last_cubic_persist_ms(ctx.h, C.byref(ms)) == 0:
            persist.append(float(ms.value))
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - l0
    ms_dev, ms_wall = float(np.mean(dev)), float(np.mean(wall))
    ph = {k: float(np.mean([p[k] for p in phases])) for k in phases[0]}
    parity = {}
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_dev, ms_wall] + [ph[k] for k in sorted(ph)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)                       # max over ranks, per phase too
        ms_dev, ms_wall = float(t[0]), float(t[1])
        ph = {k: float(t[2 + i]) for i, k in enumerate(sorted(ph))}
    single = None
    if sharded:
        # every rank must hold the same proof; rank 0 also proves the same instance on ONE GPU (same protocol) for the
        # strong-scaling reference point and compares the two proofs bit for bit
        import hashlib
        blob = hashlib.sha256(b"".join(getattr(proof, f).tobytes() for f in sp.SpartanProof.FIELDS)).hexdigest()
        parity["all_ranks_identical"] = len(set(allgather_bytes(blob))) == 1
        prep.free(); S.free()
        if rank == 0:
            S1 = sp.SplitR1CSShape(ctx, *wl.circ.dims(), wl.A, wl.B, wl.C)
            prep1 = sp.SpartanSNARK.prep_prove(ctx, S1, K, wl.W[:wl.cached_len], wl.blinds[:wl.cached_rows], is_small=True)
            d1, w1, p1s = [], [], []
            for i in range(3 + max(3, args.steps // 2)):
                p1, w = step(S1, prep1, None)
                if i >= 3:
                    d1.append(p1.phase_ms["total"]); w1.append(w); p1s.append(p1.phase_ms)
            parity["sharded_equals_single_gpu_proof"] = all(np.array_equal(getattr(proof, f), getattr(p1, f)) for f in sp.SpartanProof.FIELDS)
            single = {"ms_per_step": float(np.mean(d1)), "e2e_ms_per_step": float(np.mean(w1)),
                      "phase_ms": {k: float(np.mean([p[k] for p in p1s])) for k in p1s[0]},
                      "note": "the same instance, keys, randomness and timing protocol on rank 0's GPU alone (strong-scaling reference point)"}
            prep1.free(); S1.free()
    if rank == 0:
        ach_phase = wl.bytes_outer / (ph["outer_sumcheck"] * 1e-3) / 1e9
        traffic = ncu_traffic("k_cubic_persist")
        if persist:
            k_ms = float(np.mean(persist))
            roof = {"bound": "hbm", "kernel": "k_cubic_persist (rounds 1-%d of the outer sum-check, bind fused into the next evaluation, one cooperative launch)" % wl.persist_rounds,
                    "achieved": wl.bytes_persist / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "traffic": traffic, "peak_kind": peak_kind,
                    "ms": k_ms, "algorithmic_bytes": wl.bytes_persist, "share_of_step": k_ms / ms_dev,
                    "note": "the %d streaming rounds of the outer sum-check (tables > 2^17 entries; the later rounds run in k_cubic_mid_pipe / k_cubic_tail_pipe, "
                            "bound by the transcript); bound by the integer pipe, not by HBM: see int_pipe, and tables_bench for the same kernel on 2^24-entry tables" % wl.persist_rounds}
            # the binding resource: 256-bit Montgomery multiplications on the IMAD pipe; peak = Fq::mul throughput measured by
            # tools/microbench/field_mul.cu on this pool's B200 (profiles/r2_b_microbench_field_mul.txt); field-ops per SURVEY §8(d): 7 T (1 - 2^-rounds)
            fops = 7 * wl.N * (1.0 - 0.5 ** wl.persist_rounds)
            roof["int_pipe"] = {"achieved": fops / (k_ms * 1e-3) / 1e9, "peak": 90.2, "unit": "G field-mul/s", "frac": fops / (k_ms * 1e-3) / 1e9 / 90.2,
                                "field_ops": fops, "peak_kind": "microbenchmark (Fq::mul, 4 independent chains per thread, 148 SMs)"}
        else:
            # sharded: the outer sum-check is one launch per round on every rank (k_cubic_round, the in-kernel exchange needs the
            # last-CTA form); the phase is reported against the aggregate HBM bandwidth of the GPUs it runs on
            roof = {"bound": "hbm", "kernel": "outer sum-check phase (all %d rounds, k_cubic_round per round on each of %d GPUs + gather + tail)" % (proof.l, world),
                    "achieved": ach_phase, "peak": hbm_peak * world, "unit": "GB/s", "traffic": None, "peak_kind": peak_kind + " x %d GPUs" % world,
                    "ms": ph["outer_sumcheck"], "algorithmic_bytes": wl.bytes_outer}
        roof["frac"] = roof["achieved"] / roof["peak"]
        par = "single GPU" if world == 1 else ("one proof, hypercube sharded across %d GPUs (rows/columns i mod %d; per-round sums exchanged in-kernel over NVLink)" % (world, world)
                                               if sharded else "1 proof per GPU (replicas)")
        jobs = 1 if sharded or world == 1 else world
        out = {
            "metric": METRIC, "value": jobs * wl.field_ops / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": W_, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "u32x8 limbs (256-bit prime field, Montgomery)", "data": "synthetic", "config": wl.describe(),
            "protocol": {"l2": "flushed between timed iterations (512 MiB memset)", "parallelism": par, "prep_prove_ms_untimed": prep_ms,
                         "timing": "CUDA events on the library's stream around one prove, mean of the timed steps, max over ranks"},
            "e2e": {"value": jobs * wl.field_ops / (ms_wall * 1e-3), "unit": UNIT, "ms_per_step": ms_wall,
                    "h2d_bytes_per_step": int(wl.X.nbytes + wl.d_vec.nbytes + wl.blinds.nbytes + 16 * 32 + 64 * 21),
                    "d2h_bytes_per_step": int(sum(getattr(proof, f).nbytes for f in sp.SpartanProof.FIELDS))},
            "gpu_launches": int(launches),
            "roofline": roof,
            "roofline_phase": {"phase": "outer_sumcheck (all %d rounds)" % proof.l, "achieved": ach_phase, "unit": "GB/s",
                               "frac": ach_phase / (hbm_peak * (world if sharded else 1)), "ms": ph["outer_sumcheck"], "algorithmic_bytes": wl.bytes_outer},
            "phase_ms": ph, "prove_ms": ms_dev, "clocks": clocks,
        }
        if single is not None:
            out["single_gpu"] = single
    else:
        out = None
    if world == 1 and not args.no_extras:
        out["tables_bench"] = tables_bench(ctx, sp, hbm_peak)
        out["neutronnova"] = neutronnova_bench(ctx, sp, hbm_peak, with_cpu=not args.no_cpu_baseline)
    if world == 1 and not args.no_cpu_baseline:
        # the oracle port, one thread, on the full workload; its proof is compared with the device's bit for bit
        cb, oproof = cpu_prove(wl, pts, threads=1, steps=1, warmup=0, want_proof=True)
        parity["fields_compared"] = len(sp.SpartanProof.FIELDS)
        parity["bit_exact_vs_oracle_prover"] = all(np.array_equal(getattr(proof, f).reshape(-1), getattr(oproof, f).reshape(-1)) for f in sp.SpartanProof.FIELDS)
        out["cpu_baseline"] = cb
        if not args.no_extras:
            out["cpu_baseline_config1"] = cpu_prove(Workload(1024), pts, threads=1, steps=1, warmup=0)
    if sharded and not args.no_extras:
        nn5 = neutronnova_sharded_bench(ctx, sp, comm, rank, world, local, allgather_bytes, n=256)
        if rank == 0:
            out["neutronnova_config5"] = nn5
    if rank == 0:
        out["parity"] = parity
        print(json.dumps(out), flush=True)
    if comm is not None:
        comm.free()
    ctx.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def neutronnova_sharded_bench(ctx, sp, comm, rank, world, local, allgather_bytes, n=256):
    """BASELINE config 5: sha256_neutronnova with n = 256 step circuits, the instances sharded across the GPUs — the full non-ZK
    prove (sp2_neutronnova_prep_prove_sharded + _prep_commit + sp2_neutronnova_snark_prove_sharded: every rank rerandomises its own
    instances, local NIFS rounds with the per-round sums and the bulk exchanges over NVLink peer memory, replicated sum-checks and
    PCS); wall clock of the C-ABI call, max over ranks; rank 0 also runs all n instances on one GPU and compares the proofs."""
    import hashlib
    import torch
    import torch.distributed as dist
    from spartan2_b200 import neutronnova as nn
    from spartan2_b200 import _fq as fq
    from spartan2_b200.frontend import Sha256Circuit
    one = fq.from_int(1); dev = torch.device("cuda", local)
    nl = n // world
    fields = ["comm_W_steps", "comm_W_core", "nifs_polys", "outer_polys", "claims_outer", "inner_polys", "eval_W", "blind_eval_W", "delta", "beta", "z_vec", "z_delta", "z_beta"]

    def z_of(c):
        W, X = c.witness()
        return np.concatenate([W, one, X], axis=0)
    core = Sha256Circuit(bytes(64), kind="compression")
    A, B, Cm = core.matrices()
    M = core.num_vars; rows = M // WIDTH; pre_rows = core.num_precommitted // WIDTH
    S = sp.SplitR1CSShape(ctx, *core.dims(), A, B, Cm)
    pts = ctx.test_points(WIDTH + 3, seed=7)
    K = sp.CommitmentKey(ctx, pts[:WIDTH], pts[WIDTH:WIDTH + 1], pts[WIDTH + 1:WIDTH + 2], pts[WIDTH + 2:WIDTH + 3])
    rng = np.random.default_rng(0xDEADBEEF)                  # same randomness on every rank
    b_old_s, b_old_c = rand_fe(rng, n * pre_rows), rand_fe(rng, pre_rows)
    rnd = (rand_fe(rng, n * rows), rand_fe(rng, rows), rand_fe(rng, 2), rand_fe(rng, WIDTH), rand_fe(rng, 1), rand_fe(rng, 1))
    mine = [z_of(Sha256Circuit(bytes([i % 256]) * 64, kind="compression")) for i in range(rank * nl, (rank + 1) * nl)]
    t0 = time.perf_counter()
    prover = nn.NeutronNovaProver(ctx, S, mine, z_of(core), rank=rank, nranks=world, allgather=nn.torch_allgather(world, dev), comm=comm,
                                  allgather_bytes=allgather_bytes)
    prover.commit(K, b_old_s[rank * nl * pre_rows:(rank + 1) * nl * pre_rows], b_old_c)
    ctx.synchronize(); prep_ms = (time.perf_counter() - t0) * 1e3
    walls, phs = [], []
    for it in range(8):
        dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        v, ph = prover.snark_prove(bytes(32), *rnd)
        if it >= 3:
            walls.append((time.perf_counter() - t0) * 1e3); phs.append(ph)
    t = torch.tensor([float(np.mean(walls))] + [float(np.mean([p[k] for p in phs])) for k in sorted(phs[0])], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    blob = hashlib.sha256(b"".join(np.ascontiguousarray(v[k]).tobytes() for k in fields)).hexdigest()
    same = len(set(allgather_bytes(blob))) == 1
    out = None
    prover.free()
    if rank == 0:
        allz = [z_of(Sha256Circuit(bytes([i % 256]) * 64, kind="compression")) for i in range(n)]
        single = nn.NeutronNovaProver(ctx, S, allz, z_of(core)); single.commit(K, b_old_s, b_old_c)
        w1, p1 = [], []
        for it in range(6):
            t0 = time.perf_counter()
            v1, ph1 = single.snark_prove(bytes(32), *rnd)
            if it >= 3:
                w1.append((time.perf_counter() - t0) * 1e3); p1.append(ph1)
        eq = all(np.array_equal(v[k], v1[k]) for k in fields) and bool(v["outer_ok"] and v["inner_ok"])
        single.free()
        out = {"workload": "sha256_neutronnova_%d_steps (N = M = 2^15 per instance), %d instances per GPU, full non-ZK prove incl. the commitment half" % (n, nl),
               "prove_ms": float(t[0]), "phase_ms": {k: float(t[1 + i]) for i, k in enumerate(sorted(phs[0]))}, "prep_prove_ms_untimed": prep_ms,
               "single_gpu": {"prove_ms": float(np.mean(w1)), "phase_ms": {k: float(np.mean([p[k] for p in p1])) for k in p1[0]}},
               "parity": {"all_ranks_identical": same, "sharded_equals_single_gpu": eq},
               "timing": "host wall clock of the C-ABI call (every phase ends in a host wait), mean of 5 after 3 warm-ups, max over ranks"}
    S.free(); K.free()
    return out


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the capture named in it); None if absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return float(json.load(open(p))[kernel]["dram_bytes"])
    except Exception:
        return None


def tables_bench(ctx, sp, hbm_peak, num_vars=24):
    """The reference's pure-table sum-check benchmark (src/sumcheck.rs:1450-1553) at num_vars = 24 on device-resident
    tables: the SAME kernels as the prove, at a size where streaming rather than per-round latency dominates."""
    import ctypes as C
    rng = np.random.default_rng(0xDEADBEEF)
    n = 1 << num_vars; chunk = 1 << 20
    host = [rand_fe(rng, chunk) for _ in range(3)]
    tabs = [ctx.alloc(n * 32) for _ in range(3)]
    small = [ctx.upload(h) for h in host]

    def fill():
        for t, sbuf in zip(tabs, small):
            for i in range(n // chunk):
                ctx.check(ctx.L.sp2_dev_copy(ctx.h, C.c_void_p(t.ptr.value + i * chunk * 32), sbuf.ptr, C.c_uint64(chunk * 32)))
    taus = rand_fe(rng, num_vars); zero = np.zeros((1, 4), dtype=np.uint64)
    tot, per = [], []
    for it in range(4):
        fill()
        ctx.synchronize()
        ts = sp.TranscriptState()
        ctx.timer_start()
        sp.SumcheckProof.prove_cubic_with_three_inputs(ctx, zero, taus, tabs[0], tabs[1], tabs[2], ts)
        ms = ctx.timer_stop()
        k = C.c_float(); ctx.check(ctx.L.sp2_last_cubic_persist_ms(ctx.h, C.byref(k)))
        if it:
            tot.append(ms); per.append(float(k.value))
    for t in tabs + small:
        t.free()
    l = num_vars
    r_end = persist_end(l, int(ctx.L.sp2_sc_tail_len()), int(ctx.L.sp2_sc_mid_len()))
    b_persist = 80 * n + sum(144 * (n >> (i - 2)) for i in range(2, r_end))
    k_ms, t_ms = float(np.mean(per)), float(np.mean(tot))
    return {"workload": "prove_cubic_with_three_inputs, 3 uniform random tables of 2^%d entries, device-resident (1.5 GiB > L2)" % num_vars,
            "kernel": "k_cubic_persist (rounds 1-%d)" % (r_end - 1), "ms": k_ms, "algorithmic_bytes": b_persist,
            "achieved": b_persist / (k_ms * 1e-3) / 1e9, "unit": "GB/s", "frac": b_persist / (k_ms * 1e-3) / 1e9 / hbm_peak,
            "sumcheck_total_ms": t_ms, "field_ops_per_sec": 7 * n / (t_ms * 1e-3)}


def neutronnova_bench(ctx, sp, hbm_peak, n=32, with_cpu=True):
    """BASELINE config 3: sha256_neutronnova, 32 step circuits (2048 B total), multi-fold + Spartan prove on one GPU — the FULL
    non-ZK prove of the library (sp2_neutronnova_prep_commit + sp2_neutronnova_snark_prove: rerandomisation, commit_zeros, the
    instance transcript, NIFS, both batched sum-checks, the commitment / witness folds and PCS::prove), wall clock of the C-ABI
    call from host buffers (= e2e: every input arrives from and every output returns to host memory inside the call).  The
    oracle's C driver of the same protocol (oracle.c: orc_neutronnova_prove) runs once on one host thread: the CPU baseline,
    and every proof field is compared bit for bit; the oracle's verifier checks the device-made proof."""
    import ctypes as C
    from spartan2_b200 import neutronnova as nn
    from spartan2_b200.frontend import Sha256Circuit
    from spartan2_b200 import _fq as fq
    one = fq.from_int(1)
    circs = [Sha256Circuit(bytes([i % 256]) * 64, kind="compression") for i in range(n)] + [Sha256Circuit(bytes(64), kind="compression")]
    zs = []
    for c in circs:
        W, X = c.witness(); zs.append(np.concatenate([W, one, X], axis=0))
    c0 = circs[0]
    A, B, Cm = c0.matrices(); d = c0.dims()
    N = c0.num_cons; M = c0.num_vars; rows = M // WIDTH; pre_rows = c0.num_precommitted // WIDTH
    S = sp.SplitR1CSShape(ctx, *d, A, B, Cm)
    pts = ctx.test_points(WIDTH + 3, seed=7)
    K = sp.CommitmentKey(ctx, pts[:WIDTH], pts[WIDTH:WIDTH + 1], pts[WIDTH + 1:WIDTH + 2], pts[WIDTH + 2:WIDTH + 3])
    rng = np.random.default_rng(0xDEADBEEF)
    b_old_s, b_old_c = rand_fe(rng, n * pre_rows), rand_fe(rng, pre_rows)
    rnd = (rand_fe(rng, n * rows), rand_fe(rng, rows), rand_fe(rng, 2), rand_fe(rng, WIDTH), rand_fe(rng, 1), rand_fe(rng, 1))
    vk = bytes(32)
    t0 = time.perf_counter()
    prover = nn.NeutronNovaProver(ctx, S, zs[:n], zs[n])
    comm_s, comm_c = prover.commit(K, b_old_s, b_old_c)
    prep_ms = (time.perf_counter() - t0) * 1e3
    walls, phs, r0 = [], [], []
    l0 = None
    for it in range(8):
        if it == 3:
            l0 = ctx.launch_count()
        t0 = time.perf_counter()
        v, ph = prover.snark_prove(vk, *rnd)
        if it >= 3:
            walls.append((time.perf_counter() - t0) * 1e3); phs.append(ph)
            ms = C.c_float()
            if ctx.L.sp2_neutronnova_last_round0_ms(prover.h, C.byref(ms)) == 0:
                r0.append(float(ms.value))
        assert v["outer_ok"] and v["inner_ok"]
    launches = (ctx.launch_count() - l0) // 5
    hot = []
    for it in range(5):                                # HOT LOOPS A-C alone (round 1's config-3 number), for continuity
        t0 = time.perf_counter(); prover.prove(sp.Keccak256Transcript(b"neutronnova_prove")); hot.append((time.perf_counter() - t0) * 1e3)
    is_gen = np.array([abs(x) != 1 for x in c0.coef_values], dtype=bool)
    gen = sum(int(is_gen[co].sum()) for (co, _, _) in c0.raw)
    # field-ops: NIFS rounds 3 n_t N + folds (sum over rounds ~ 6 n N), witness fold n M, pow-cubic 2 branches 12 N,
    # ABC 2 x (2N + general nnz), inner 2 x 4 (2M), W = W_fold + c W_core and the Hyrax bind 2 M
    field_ops = 6 * n * N + n * M + 12 * N + 2 * (2 * N + gen) + 16 * M + 2 * M
    ms = float(np.mean(walls)); phm = {k: float(np.mean([p[k] for p in phs])) for k in phs[0]}
    out = {"workload": "sha256_neutronnova_%d_steps (N = M = 2^%d per instance; %d + %d commitment rows of %d)" % (n, N.bit_length() - 1, pre_rows, rows - pre_rows, WIDTH),
           "metric": METRIC, "value": field_ops / (ms * 1e-3), "unit": UNIT, "field_ops_per_step": field_ops,
           "prove_ms": ms, "phase_ms": phm, "hot_loops_only_ms": float(np.mean(hot[2:])), "prep_prove_ms_untimed": prep_ms, "gpu_launches": int(launches),
           "e2e": {"value": field_ops / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                   "h2d_bytes_per_step": int(sum(x.nbytes for x in rnd) + 32),
                   "d2h_bytes_per_step": int(sum(np.asarray(v[k]).nbytes for k in ("comm_W_steps", "comm_W_core", "nifs_polys", "outer_polys", "claims_outer", "inner_polys",
                                                                                    "eval_W", "delta", "beta", "z_vec", "z_delta", "z_beta"))),
                   "note": "the prove takes host buffers and returns the proof in host memory: wall clock of the C-ABI call"},
           "protocol": "non-ZK variant: round polynomials absorbed directly instead of committed in the reference's in-circuit verifier (process_round, out of scope); "
                       "checker oracle/oracle.c orc_neutronnova_prove / _verify"}
    if r0:
        k_ms = float(np.mean(r0)); by = 2 * n * N * 8
        out["roofline"] = {"bound": "hbm", "kernel": "k_nifs_round0_small (NIFS round 0: i64 Az/Bz layers of all %d instances, prove_helper_small)" % n,
                           "achieved": by / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": by / (k_ms * 1e-3) / 1e9 / hbm_peak, "ms": k_ms,
                           "algorithmic_bytes": by, "traffic": None, "share_of_step": k_ms / ms}
    by_outer = 768 * N
    out["roofline_phase"] = {"phase": "outer_sumcheck_batched (pow-cubic, 2 branches, %d rounds, one launch per round)" % (N.bit_length() - 1), "algorithmic_bytes": by_outer,
                             "ms": phm["outer_sumcheck_batched"], "achieved": by_outer / (phm["outer_sumcheck_batched"] * 1e-3) / 1e9, "unit": "GB/s",
                             "frac": by_outer / (phm["outer_sumcheck_batched"] * 1e-3) / 1e9 / hbm_peak,
                             "note": "latency-bound: 2^15-entry tables, ~40 us per round against ~1 us of streaming"}
    if with_cpu:
        from oracle import pyoracle as orc
        orc.lib(native=True); orc.set_threads(1)
        O = orc.Shape(*d, A, B, Cm)
        keys = orc.Keys(pts[:WIDTH], pts[WIDTH:WIDTH + 1], pts[WIDTH + 1:WIDTH + 2], pts[WIDTH + 2:WIDTH + 3])
        t0 = time.perf_counter()
        P = orc.neutronnova_prove(O, keys, vk, np.stack(zs[:n]), zs[n], comm_s, b_old_s, comm_c, b_old_c, orc.NnRand(*rnd))
        cpu_wall = (time.perf_counter() - t0) * 1e3
        cpu_ms = float(sum(P.phase_ms.values()))            # prove-phase work (the per-step SpMV / i64 conversion of prep_prove is excluded)
        fields = ["comm_W_steps", "comm_W_core", "nifs_polys", "outer_polys", "claims_outer", "inner_polys", "eval_W", "blind_eval_W", "delta", "beta", "z_vec", "z_delta", "z_beta"]
        exact = all(np.array_equal(np.asarray(v[k]).reshape(-1), getattr(P, k).reshape(-1)) for k in fields)
        V = orc.NnProof(n, N, M, WIDTH)
        for k in fields:
            getattr(V, k)[...] = np.asarray(v[k]).reshape(getattr(V, k).shape)
        accepted = orc.neutronnova_verify(O, keys, vk, np.zeros((1, 4), dtype=np.uint64), np.zeros((1, 4), dtype=np.uint64), V) == 0
        out["cpu_baseline"] = {"value": field_ops / (cpu_ms * 1e-3), "unit": UNIT, "cores": 1, "kind": "port", "ms_per_step": cpu_ms, "wall_ms_incl_prep_work": cpu_wall,
                               "phase_ms": P.phase_ms, "host": host_cpu(), "sample": "full workload, 1 prove, oracle/oracle.c orc_neutronnova_prove, 1 thread"}
        out["parity"] = {"fields_compared": len(fields), "bit_exact_vs_oracle_prover": bool(exact), "oracle_verifier_accepts_device_proof": bool(accepted)}
    prover.free(); S.free(); K.free()
    return out


def cpu_prove(wl, pts, threads, steps, warmup, want_proof=False):
    """The oracle port of the reference's prover on the same workload (oracle/oracle.c, -march=native)."""
    from oracle import pyoracle as orc
    orc.lib(native=True)
    orc.set_threads(threads)
    O = orc.Shape(*wl.circ.dims(), wl.A, wl.B, wl.C)
    keys = orc.Keys(pts[:WIDTH], pts[WIDTH:WIDTH + 1], pts[WIDTH + 1:WIDTH + 2], pts[WIDTH + 2:WIDTH + 3])
    comm_pre = orc.hyrax_commit(pts[:WIDTH], pts[WIDTH:WIDTH + 1], wl.W[:wl.cached_len], wl.blinds[:wl.cached_rows], is_small=True)
    rnd = orc.Rand(wl.blinds, wl.blind_eval, wl.d_vec, wl.r_delta, wl.r_beta)
    cached = orc.spartan_prep_cached(O, wl.W)        # prep_prove's cached Az/Bz/Cz (spartan.rs:184-187): prove adds the remaining columns only
    times, ph = [], None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        p = orc.spartan_prove(O, keys, wl.vk, wl.X, wl.W, comm_pre, rnd, cached=cached)
        if i >= warmup:
            times.append((time.perf_counter() - t0) * 1e3); ph = p.phase_ms
    ms = float(np.mean(times))
    res = _cpu_result(wl, ms, threads, steps, ph)
    return (res, p) if want_proof else res


def host_cpu():
    model = "unknown"
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                model = ln.split(":", 1)[1].strip(); break
    except OSError:
        pass
    return {"model": model, "nproc": os.cpu_count()}


def _cpu_result(wl, ms, threads, steps, ph):
    return {"value": wl.field_ops / (ms * 1e-3), "unit": UNIT, "cores": threads, "kind": "port", "ms_per_step": ms, "host": host_cpu(),
            "phase_ms": dict(zip(["commit_rest+transcript", "matrix_vector_multiply", "outer_sumcheck", "prepare_poly_ABC", "inner_sumcheck", "pcs_prove"], ph)),
            "sample": "full workload %s, %d prove(s); oracle/oracle.c = C restatement of the reference's rayon prover (no Rust toolchain), OpenMP %d thread(s)" % (wl.name, steps, threads)}


def oracle_points(orc, n, seed=7):
    rng = np.random.default_rng(seed)
    order = 0xffffffff00000001000000000000000000000000ffffffffffffffffffffffff
    g = np.concatenate([orc.to_mont([3], orc.FP), orc.to_mont([0x5a6dd32df58708e64e97345cbe66600decd9d538a351bb3c30b4954925b1f02d], orc.FP)], axis=1)
    deltas = [orc.scalar_mul(g, orc.to_mont([int.from_bytes(rng.bytes(32), "little") % order])) for _ in range(8)]
    cur = orc.scalar_mul(g, orc.to_mont([int.from_bytes(rng.bytes(32), "little") % order]))
    out = np.zeros((n, 8), dtype=np.uint64)
    for i in range(n):
        out[i] = cur[0]
        cur = orc.point_add(cur, deltas[int(rng.integers(0, 8))])
    return out


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import pyoracle as orc
    orc.lib(native=True)
    threads = orc.max_threads()
    wl = Workload(default_msg_len(args, args.gpus))          # the CUDA arm's config at this --gpus
    pts = oracle_points(orc, WIDTH + 3)
    r = cpu_prove(wl, pts, threads, args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference", "reference_kind": "oracle_port (oracle/oracle.c: C/OpenMP restatement of the reference's rayon prover; the Rust crate cannot be built in this image)",
        "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong" if (args.gpus > 1 and not args.replicas) else "weak", "vs_baseline": None,
        "dtype": "u64x4 limbs (256-bit prime field, Montgomery)", "data": "synthetic", "config": wl.describe(), "phase_ms": r["phase_ms"],
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"], "host": r["host"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--msg-len", type=int, default=0, help="SHA-256 message bytes (default: 2048 = BASELINE config 2 at --gpus 1, 8192 = config 4 at --gpus N > 1; config 1: 1024)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the 2^24 pure-table sum-check leg and the config-3 NeutronNova leg (N = 1 only)")
    ap.add_argument("--replicas", action="store_true", help="N > 1: one independent proof per GPU (weak scaling) instead of ONE proof sharded across the GPUs (the default)")
    ap.add_argument("--sharded", action="store_true", help="(default for N > 1; kept for compatibility)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
