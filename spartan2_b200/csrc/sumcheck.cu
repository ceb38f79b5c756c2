// Sum-check provers on the device — whole loops, no host round trip per round.
//
// Restates reference src/sumcheck.rs:
//   prove_cubic_with_three_inputs (:502-571) with EqSumCheckInstance (:934-1405): split-eq (Gruen)
//     weights, round polynomial s(X) = l(X) * eval_eq_left * t(X); bound (:1399-1405)
//   prove_quad (:190-247) with compute_eval_points_quad (:128-174)
// and the per-round transcript step (absorb b"p" / squeeze b"c", :536-548) of
// src/provider/keccak.rs:70-99.
//
// B200 design (not the reference's rayon fold/reduce):
//   * one launch per round while the tables are large; the launch for round i first binds the tables
//     to challenge r_{i-1} (in place: the thread that reads T[j], T[j+L/4], T[j+L/2], T[j+3L/4]
//     writes T'[j], T'[j+L/4]) and evaluates round i on the freshly bound values while they are
//     still in registers — every table element is read once and written once per round
//     (368*T bytes in total for three tables of length T);
//   * per-thread 544-bit delayed-reduction accumulators, warp-shuffle + block reduction, one
//     partial triple per CTA; the LAST CTA to finish (atomic ticket) sums the partials, does the
//     round's scalar algebra, hashes the transcript (Keccak-f in registers, the two challenge halves
//     on two warps) and publishes the challenge in device memory for the next launch: rounds are
//     chained on the stream without host syncs;
//   * once a table fits one SM's reach (<= SC_TAIL_LEN entries) a single CTA runs all remaining
//     rounds in one launch (no relaunch, no ticket, warm instruction cache);
//   * t(0), t(1), t(inf) are all summed directly.  The reference derives t(1) from the running
//     claim with one field inversion per round (derive_from_claim, :1277-1324) and needs a
//     fallback when tau_i * eval_eq_left = 0 (:1327-1396); a serial 256-bit inversion costs
//     ~100 us on a GPU thread, a third fused sum costs ~2% more streaming work, needs no fallback
//     and yields the same polynomial.
// Every emitted value is the canonical representative of the same field element the reference
// computes, hence bit-identical.
// The field multiplication is an out-of-line call in this translation unit (field.cuh): measured on B200 the streaming rounds
// are no slower (they are bound by the integer pipe either way) and the latency-bound rounds gain from the smaller code
// (k_cubic_persist 398 -> 263 KB; outer sum-check of the 2^20 prove 0.661 -> 0.625 ms).
#ifndef SP2_SC_INLINE_MUL
#define SP2_FQ_OUTLINE 1
#endif
#include <stdlib.h>
#include <string.h>
#include <utility>
#include "ctx.cuh"
#include "host_transcript.h"
#include "keccak.cuh"
#include "polys.cuh"
#include "sumcheck.cuh"

using namespace sp2;

namespace sp2 {

struct FinSmem {
  u64 m[2][34];               // padded squeeze input for the lo / hi hash (2 Keccak blocks each)
  u64 dg[8];                  // lo || hi digests
  fe g[8];
  fe ch;
  fe tq[3];                   // the round's t(0), tb, t(inf) for the off-path claim update (cubic_bound)
  fe lt[SC_DERIVE_MAX];       // derived t(1): (1 - tau_i) / tau_i per streaming round ...
  fe ct;                      // ... and t_{i-1}(r_{i-1}) / tau_i of the round being evaluated: t(1) = ct - lt_i t(0)
  u32 tinvw[SC_DERIVE_MAX * 8];   // the host's tau inverses (ScTinvMail), fetched once
  fe red[3 * 32];
  int is_last;
};

__device__ __forceinline__ fe ld_state(const fe *p) {   // device-produced scalars: bypass L1
  fe r; u64 a, b, c, d;
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  r.v[0] = (u32)a; r.v[1] = (u32)(a >> 32); r.v[2] = (u32)b; r.v[3] = (u32)(b >> 32);
  r.v[4] = (u32)c; r.v[5] = (u32)(c >> 32); r.v[6] = (u32)d; r.v[7] = (u32)(d >> 32);
  return r;
}
__device__ __forceinline__ u32 ld_volatile_u32(const u32 *p) { u32 v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
// out-of-line multiply for the serial scalar tail: keeps the finaliser's code (and its cold
// instruction-cache footprint) small
__device__ __noinline__ void fq_mul_ni(fe *out, const fe *a, const fe *b) { *out = Fq::mul(*a, *b); }
__device__ __forceinline__ fe mul_ni(const fe &a, const fe &b) { fe o; fq_mul_ni(&o, &a, &b); return o; }

__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#ifdef SP2_TAIL_TRACE
__device__ long long g_tail_trace[2][40][10];
#define TT(kern, slot_) do { g_tail_trace[kern][round1][slot_] = (long long)gtimer(); } while (0)
#else
#define TT(kern, slot_) do { } while (0)
#endif
// Bounded spin: waits until pred() holds; gives up after SC_WAIT_NS of %globaltimer (sampled every 256 polls so the
// timer read stays off the fast path) and raises *err instead of hanging — a peer that failed or died, or a grid that
// lost a CTA, must come back to the host as SP2_ERR_INTERNAL, not as a wedged GPU.
template <class Pred>
__device__ __forceinline__ bool spin_until(Pred pred, u32 *err) {
  unsigned long long t0 = 0; u32 polls = 0;
  while (!pred()) {
    if ((++polls & 255u) == 0) {
      const unsigned long long t = gtimer();
      if (t0 == 0) t0 = t;
      else if (t - t0 > SC_WAIT_NS) { atomicExch(err, 1u); return false; }
      if (*(volatile u32 *)err) return false;             // another wait of this call already expired: do not queue up behind it
    }
  }
  return true;
}
// publish this CTA's partial sums and elect the last CTA of the grid
template <int NV>
__device__ __forceinline__ bool publish_and_elect(ScState *st, fe (&x)[NV], FinSmem &sm) {
  const int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (nblocks == 1) return true;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) stg_fe(&st->partial[3 * bid + k], x[k]);
    __threadfence();
    u32 t = atomicAdd(&st->ticket, 1u);
    sm.is_last = (t == (u32)nblocks - 1);
  }
  __syncthreads();
  if (!sm.is_last) return false;
  __threadfence();
  if (threadIdx.x == 0) st->gt[1] = gtimer();
  x[0] = Fq::zero(); x[1] = Fq::zero(); if (NV > 2) x[NV - 1] = Fq::zero();
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
#pragma unroll
    for (int k = 0; k < NV; k++) x[k] = Fq::add(x[k], ld_state(&st->partial[3 * b + k]));
  }
  block_sum_fq<NV>(x, sm.red);
  if (threadIdx.x == 0) st->ticket = 0;
  return true;
}

// Cross-GPU sum of the round's partial sums, executed by the last CTA of every rank's round kernel: write the local
// sums into every peer's mailbox (peer stores over NVLink), publish with a system-scope fence + flag, wait for
// all peers' flags, add.  x valid in warp 0 on entry and on exit.
template <int NV>
__device__ __forceinline__ void exchange_sums(const DevComm &dc, int slot, fe (&x)[NV], FinSmem &sm) {
  const int tid = threadIdx.x;
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) sm.g[k] = x[k];
  }
  __syncthreads();
  if (tid < dc.n) {
    MailBox *mb = dc.peer[tid];
#pragma unroll
    for (int k = 0; k < NV; k++) stg_fe(&mb->sums[slot][dc.rank][k], sm.g[k]);
    __threadfence_system();
    *(volatile u32 *)&mb->flag[slot][dc.rank] = dc.epoch;
    // wait for rank `tid`'s contribution to arrive in the local mailbox (bounded: see spin_until)
    MailBox *me = dc.peer[dc.rank];
    const volatile u32 *fl = &me->flag[slot][tid]; const u32 want = dc.epoch;
    spin_until([=] { return *fl == want; }, &me->err);
    __threadfence_system();
  }
  __syncthreads();
  if (tid < 32) {
    const MailBox *me = dc.peer[dc.rank];
#pragma unroll
    for (int k = 0; k < NV; k++) {
      fe acc = Fq::zero();
      for (int q = 0; q < dc.n; q++) {
        fe v; u64 a, b, c, d;
        asm volatile("ld.volatile.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(&me->sums[slot][q][k]));
        v.v[0] = (u32)a; v.v[1] = (u32)(a >> 32); v.v[2] = (u32)b; v.v[3] = (u32)(b >> 32);
        v.v[4] = (u32)c; v.v[5] = (u32)(c >> 32); v.v[6] = (u32)d; v.v[7] = (u32)(d >> 32);
        acc = Fq::add(acc, v);
      }
      x[k] = acc;
    }
  }
}

// All-gather of the shards into every rank's gather area (global index order i = (j << k) | rank), by peer stores.
__global__ void __launch_bounds__(256) k_shard_gather(DevComm dc, const fe *A, const fe *B, const fe *C, u64 len_local, int ntab) {
  const fe *T[3] = {A, B, C};
  for (int t = 0; t < ntab; t++)
    for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < len_local; j += (u64)gridDim.x * blockDim.x) {
      const fe v = ldg_fe(T[t] + j);
      for (int q = 0; q < dc.n; q++) stg_fe(&dc.peer[q]->gather[t][(j << dc.k) | (u64)dc.rank], v);
    }
}
// publish / wait for the gather (one CTA, after k_shard_gather on the same stream)
__global__ void k_shard_barrier(DevComm dc, int slot) {
  const int tid = threadIdx.x;
  if (tid < dc.n) {
    __threadfence_system();
    *(volatile u32 *)&dc.peer[tid]->flag[slot][dc.rank] = dc.epoch;
    MailBox *me = dc.peer[dc.rank];
    const volatile u32 *fl = &me->flag[slot][tid]; const u32 want = dc.epoch;
    spin_until([=] { return *fl == want; }, &me->err);
    __threadfence_system();
  }
}

// Transcript step shared by both provers.  On entry lanes [0, ncoef) of warp 0 hold the canonical
// (non-Montgomery) transcript coefficients in `canon` (UniPoly::to_transcript_bytes: all but the
// linear term, each to_repr() little-endian, univariate.rs:182-190).  Returns the challenge to all.
#define SC_STAMP(k) do { if (threadIdx.x == 0) st->clk[k] = clock64(); } while (0)
__device__ __forceinline__ fe sc_squeeze(ScState *st, FinSmem &sm, const fe &canon, int ncoef) {
  const int tid = threadIdx.x;
  SC_STAMP(2);
  unsigned char *mb = (unsigned char *)sm.m[0];
  const int plen = 1 + 32 * ncoef;                 // b"p" || coefficients
  const int mlen = plen + 4 + 2 + 64 + 1;          // || "NoDS" || round_le16 || state || b"c"
  const int total = mlen + 1;                      // || 0x00 / 0x01 (compute_updated_state, keccak.rs:33-54)
  const int nblocks = total / 136 + 1;
  if (tid < 34) sm.m[0][tid] = 0;
  __syncthreads();
  if (tid < ncoef) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const u32 w = canon.v[k];
      unsigned char *q = mb + 1 + 32 * tid + 4 * k;
      q[0] = (unsigned char)w; q[1] = (unsigned char)(w >> 8); q[2] = (unsigned char)(w >> 16); q[3] = (unsigned char)(w >> 24);
    }
  }
  if (tid == 32) {
    const u32 round = st->ts.round;
    mb[0] = 'p';
    mb[plen + 0] = 'N'; mb[plen + 1] = 'o'; mb[plen + 2] = 'D'; mb[plen + 3] = 'S';
    mb[plen + 4] = (unsigned char)(round & 0xff); mb[plen + 5] = (unsigned char)(round >> 8);
    mb[mlen - 1] = 'c';
    mb[total] ^= 0x01;                             // Keccak (not SHA-3) padding
    mb[nblocks * 136 - 1] ^= 0x80;
    st->ts.round = round + 1;
  }
  if (tid >= 64 && tid < 128) mb[plen + 6 + (tid - 64)] = st->ts.state[tid - 64];
  __syncthreads();
  if (tid < 34) sm.m[1][tid] = sm.m[0][tid];
  __syncthreads();
  if (tid == 32) ((unsigned char *)sm.m[1])[mlen] = 0x01;
  __syncwarp();
  SC_STAMP(3);
  if (st->flags & 1u) {
    if (tid < 64) {                                // warp 0 -> lo, warp 1 -> hi: one Keccak lane per GPU lane
      const int lane = tid & 31, w = tid >> 5;
      const KeccakLane kl = keccak_lane_init(lane);
      u64 s = 0;
      for (int blk = 0; blk < nblocks; blk++) {
        if (lane < 17) s ^= sm.m[w][blk * 17 + lane];
        s = keccak_f_warp(s, kl, lane);
      }
      if (lane < 4) sm.dg[w * 4 + lane] = s;
    }
  } else if (tid == 0 || tid == 32) {
    u64 out[4];
    keccak256_padded(sm.m[tid >> 5], nblocks, out);
#pragma unroll
    for (int i = 0; i < 4; i++) sm.dg[(tid >> 5) * 4 + i] = out[i];
  }
  __syncthreads();
  SC_STAMP(4);
  if (tid >= 64 && tid < 128) st->ts.state[tid - 64] = ((unsigned char *)sm.dg)[tid - 64];
  if (tid < 2) {                                   // from_uniform: lo * R^2 + hi * R^3 (Montgomery form of lo + 2^256 hi)
    fe h;
#pragma unroll
    for (int i = 0; i < 4; i++) { h.v[2 * i] = (u32)sm.dg[4 * tid + i]; h.v[2 * i + 1] = (u32)(sm.dg[4 * tid + i] >> 32); }
    sm.g[6 + tid] = mul_ni(h, tid == 0 ? Fq::cst_r2() : Fq::cst_r3());
  }
  __syncthreads();
  if (tid == 0) sm.ch = Fq::add(sm.g[6], sm.g[7]);
  __syncthreads();
  SC_STAMP(5);
  return sm.ch;
}

// ---------------------------------------------------------------------------------------------
// cubic: s(X) = l(X) * p * t(X),  l(X) = (1-tau) + (2tau-1) X,  t(X) = t0 + tb X + tinf X^2
// x = (t(0), t(1), t(inf)) valid in every lane of warp 0.
// ---------------------------------------------------------------------------------------------
// The finaliser is split so that only what the NEXT ROUND'S PAIR WORK needs (the challenge) sits on the critical
// path: cubic_finalize_pre ends with r known to every thread of the CTA; cubic_bound (the eq-prefix update, three
// serial multiplications that only the next round's FINALISER reads) and cubic_claims run afterwards on a thread that
// has no pair work, overlapped with the next round (persistent kernels: after the grid release; tail kernels: the
// scalar warp at the start of the next round).
// derive: x[1] was not summed; t(1) = (t_{i-1}(r_{i-1}) - (1 - tau_i) t(0)) / tau_i  (derive_from_claim, sumcheck.rs:1277-1324, with the
// eq prefix divided out on both sides and the inversion done once per prove on the host)
__device__ __forceinline__ u32 ld_sys_word(const u32 *p) { u32 v; asm volatile("ld.volatile.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ fe sm_tinv(const FinSmem &sm, int i) { fe r; for (int k = 0; k < 8; k++) r.v[k] = sm.tinvw[8 * i + k]; return r; }
// CTA 0, warp 0, once per kernel, while the other CTAs are still arriving at the first derived round's barrier: wait (bounded) for the
// host's mailbox, fetch every inverse in one PCIe round trip, and prepare lt_i = (1 - tau_i) / tau_i and ct = claim / tau_1
__device__ __forceinline__ void derive_prefetch(ScState *st, FinSmem &sm, const ScTinvMail *mail, u32 mail_epoch, int first_round1, int derive_rounds) {
  const int lane = threadIdx.x;
  if (lane == 0) spin_until([=] { return ld_sys_word(&mail->flag) == mail_epoch; }, &st->err);
  __syncwarp();
  sm.tinvw[lane] = ld_sys_word(&mail->tinv[0][0] + lane); sm.tinvw[lane + 32] = ld_sys_word(&mail->tinv[0][0] + lane + 32);
  __syncwarp();
  if (lane < derive_rounds) sm.lt[lane] = mul_ni(Fq::sub(Fq::one(), ld_state(&st->taus[lane])), sm_tinv(sm, lane));
  if (lane == 31) sm.ct = mul_ni(ld_state(&st->tclaim), sm_tinv(sm, first_round1 - 1));
  __syncwarp();
}
__device__ __forceinline__ fe cubic_finalize_pre(ScState *st, int round1, const fe (&x)[3], FinSmem &sm, bool derive = false) {
  const int tid = threadIdx.x, i = round1 - 1;
  fe canon = Fq::zero();
  SC_STAMP(1);
  if (tid < 32) {
    fe x1 = x[1];
    if (derive) {
      x1 = Fq::sub(sm.ct, mul_ni(sm.lt[i], x[0]));                          // one multiplication on the critical path
    }
    const fe tb = Fq::sub(Fq::sub(x1, x[0]), x[2]);
    if (tid == 0) { sm.tq[0] = x[0]; sm.tq[1] = tb; sm.tq[2] = x[2]; }
    if (tid < 6) {            // L0*t0, L0*tb, SL*t0, L0*tinf, SL*tb, SL*tinf
      const bool use_sl = (tid == 2) | (tid == 4) | (tid == 5);
      const fe v = fe_sel(tid == 0 || tid == 2, x[0], fe_sel(tid == 1 || tid == 4, tb, x[2]));
      sm.g[tid] = mul_ni(ld_state(use_sl ? &st->SL : &st->L0), v);
    }
    __syncwarp();
    if (tid < 4) {            // coefficients low -> high (UniPoly, univariate.rs:102-118)
      const fe co = tid == 0 ? sm.g[0] : tid == 1 ? Fq::add(sm.g[1], sm.g[2]) : tid == 2 ? Fq::add(sm.g[3], sm.g[4]) : sm.g[5];
      stg_fe(&st->polys[4 * i + tid], co);
      canon = Fq::from_mont(co);
    }
    // transcript lanes: coefficient 0, 2, 3 -> slots 0, 1, 2
    const int src = tid == 0 ? 0 : tid + 1;
#pragma unroll
    for (int k = 0; k < 8; k++) canon.v[k] = __shfl_sync(0xffffffffu, canon.v[k], src & 31);
  }
  const fe r = sc_squeeze(st, sm, canon, 3);
  if (tid == 0) stg_fe(&st->r[i], r);
  return r;
}
// bound(): p <- p * (1 - tau - r + 2 r tau) = p * l(r)   (sumcheck.rs:1399-1405), and the next round's L0 / SL.  One thread.
__device__ __forceinline__ void cubic_bound(ScState *st, int round1, int l, const fe &r, FinSmem *dsm = nullptr) {
  const int i = round1 - 1;
  // dsm: the next round derives its t(1) from t_i(r_i) = t(0) + r (tb + r t(inf)), pre-divided by its tau
  if (dsm) dsm->ct = mul_ni(Fq::add(dsm->tq[0], mul_ni(r, Fq::add(dsm->tq[1], mul_ni(r, dsm->tq[2])))), sm_tinv(*dsm, round1));
  const fe tau = ld_state(&st->taus[i]);
  const fe l0 = Fq::sub(Fq::one(), tau), sl = Fq::sub(tau, l0);
  const fe pn = mul_ni(ld_state(&st->p), Fq::add(l0, mul_ni(sl, r)));
  stg_fe(&st->p, pn);
  if (round1 < l) {
    const fe tn = ld_state(&st->taus[i + 1]);
    const fe l0n = Fq::sub(Fq::one(), tn), sln = Fq::sub(tn, l0n);
    stg_fe(&st->L0, mul_ni(pn, l0n));
    stg_fe(&st->SL, mul_ni(pn, sln));
  }
}
// final claims: bind the length-2 tables (threads 64, 96, 128 of the CTA)
__device__ __forceinline__ void cubic_claims(ScState *st, const fe *A, const fe *B, const fe *C, const fe &r) {
  const int tid = threadIdx.x;
  if (tid == 64 || tid == 96 || tid == 128) {
    const fe *T = tid == 64 ? A : tid == 96 ? B : C;
    const fe lo = ld_state(T), hi = ld_state(T + 1);
    stg_fe(&st->claims[(tid - 64) / 32], Fq::add(lo, mul_ni(Fq::sub(hi, lo), r)));
  }
}
__device__ __forceinline__ void cubic_finalize(ScState *st, int round1, int l, const fe *A, const fe *B, const fe *C,
                                               const fe (&x)[3], FinSmem &sm) {
  const fe r = cubic_finalize_pre(st, round1, x, sm);
  if (threadIdx.x == 32) cubic_bound(st, round1, l, r);
  if (round1 == l) cubic_claims(st, A, B, C, r);
  SC_STAMP(6);
}

// one (fused bind +) evaluation of a pair; accumulates E * (t(0), t(1), t(inf)) terms
template <bool FUSED>
__device__ __forceinline__ void cubic_pair(fe *A, fe *B, fe *C, u64 id, u64 P, const fe &r, const fe &w,
                                           Fq::acc &acc0, Fq::acc &acc1, Fq::acc &acci, bool skip1 = false) {
  fe a0, a1, b0, b1, c0, c1;
  if (FUSED) {
    const fe a00 = ldg_fe(A + id), a01 = ldg_fe(A + id + P), a10 = ldg_fe(A + id + 2 * P), a11 = ldg_fe(A + id + 3 * P);
    const fe b00 = ldg_fe(B + id), b01 = ldg_fe(B + id + P), b10 = ldg_fe(B + id + 2 * P), b11 = ldg_fe(B + id + 3 * P);
    const fe c00 = ldg_fe(C + id), c01 = ldg_fe(C + id + P), c10 = ldg_fe(C + id + 2 * P), c11 = ldg_fe(C + id + 3 * P);
    a0 = bind_pair(a00, a10, r); a1 = bind_pair(a01, a11, r);
    stg_fe(A + id, a0); stg_fe(A + id + P, a1);
    b0 = bind_pair(b00, b10, r); b1 = bind_pair(b01, b11, r);
    stg_fe(B + id, b0); stg_fe(B + id + P, b1);
    c0 = bind_pair(c00, c10, r); c1 = bind_pair(c01, c11, r);
    stg_fe(C + id, c0); stg_fe(C + id + P, c1);
  } else {
    a0 = ldg_fe(A + id); a1 = ldg_fe(A + id + P);
    b0 = ldg_fe(B + id); b1 = ldg_fe(B + id + P);
    c0 = ldg_fe(C + id);
    if (!skip1) c1 = ldg_fe(C + id + P);        // (t(1) derived: round 1 reads 2.5 tables, as the reference does)
  }
  Fq::mul_acc(acc0, w, Fq::sub(Fq::mul(a0, b0), c0));
  if (!skip1) Fq::mul_acc(acc1, w, Fq::sub(Fq::mul(a1, b1), c1));     // (skip1: t(1) is derived from the claim, warp-uniform)
  Fq::mul_acc(acci, w, Fq::mul(Fq::sub(a1, a0), Fq::sub(b1, b0)));
}

// generic weights: el[id >> sh] * er[id & mask] (first-half rounds on tiny tables) or er[id]
// (second half, sumcheck.rs:1107-1142)
template <bool FUSED>
__device__ __forceinline__ void cubic_generic(fe *A, fe *B, fe *C, u64 P, const fe &r, const fe *el, const fe *er, u32 sh,
                                              u64 first, u64 stride, fe (&x)[3], int shard_k = 0, int shard_rank = 0, bool skip1 = false) {
  // sh: GLOBAL width of x_in; weights are indexed by the global pair id
  Fq::acc acc0 = Fq::acc_zero(), acc1 = Fq::acc_zero(), acci = Fq::acc_zero();
  const u64 mask = ((u64)1 << sh) - 1;
  for (u64 id = first; id < P; id += stride) {
    const u64 gid = (id << shard_k) | (u64)shard_rank;
    fe w = ldg_fe_ro(er + (el ? (gid & mask) : gid));
    if (el) w = Fq::mul(ldg_fe_ro(el + (gid >> sh)), w);
    cubic_pair<FUSED>(A, B, C, id, P, r, w, acc0, acc1, acci, skip1);
  }
  x[0] = Fq::acc_reduce(acc0); x[1] = Fq::acc_reduce(acc1); x[2] = Fq::acc_reduce(acci);
}

// MODE 0: two-level split-eq.  Thread owns one x_in (coalesced across the warp) and walks x_out:
//         acc += el[x_out] * v(x_out, x_in), then one multiply by er[x_in]  (sumcheck.rs:1045-1100,
//         with the roles of the inner/outer sums swapped so the flush happens once per thread).
// MODE 1: generic.
#ifndef SC_CUBIC_MINB
#define SC_CUBIC_MINB 1
#endif
template <bool FUSED, int MODE>
__global__ void __launch_bounds__(SC_THREADS, SC_CUBIC_MINB)
k_cubic_round(ScState *st, fe *A, fe *B, fe *C, u64 P, int round1, int l, const fe *el, const fe *er, u32 out_len, u32 sh,
              DevComm dc) {
  // P, A, B, C are LOCAL (this rank's shard); weights are indexed by the GLOBAL pair id (id << k) | rank
  __shared__ FinSmem sm;
  if (threadIdx.x == 0) atomicMin(&st->gt[0], gtimer());
  fe r;
  if (FUSED) r = ld_state(&st->r[round1 - 2]);
  fe x[3];
  if (MODE == 0) {
    Fq::acc acc0 = Fq::acc_zero(), acc1 = Fq::acc_zero(), acci = Fq::acc_zero();
    // sh is the LOCAL width of x_in (global width minus k)
    const u64 xi = (u64)blockIdx.x * SC_THREADS + threadIdx.x;
    for (u32 xo = blockIdx.y; xo < out_len; xo += gridDim.y)
      cubic_pair<FUSED>(A, B, C, ((u64)xo << sh) | xi, P, r, ldg_fe_ro(el + xo), acc0, acc1, acci);
    const fe wr = ldg_fe_ro(er + ((xi << dc.k) | (u64)dc.rank));
    x[0] = Fq::mul(wr, Fq::acc_reduce(acc0));
    x[1] = Fq::mul(wr, Fq::acc_reduce(acc1));
    x[2] = Fq::mul(wr, Fq::acc_reduce(acci));
  } else {
    cubic_generic<FUSED>(A, B, C, P, r, el, er, sh, (u64)blockIdx.x * SC_THREADS + threadIdx.x, (u64)gridDim.x * SC_THREADS, x, dc.k, dc.rank);
  }
  block_sum_fq<3>(x, sm.red);
  if (!publish_and_elect<3>(st, x, sm)) return;
  if (dc.n > 1) exchange_sums<3>(dc, round1, x, sm);
  cubic_finalize(st, round1, l, A, B, C, x, sm);
  if (threadIdx.x == 0) { st->gt[4] = st->gt[2]; st->gt[2] = gtimer(); st->gt[3] = st->gt[0]; st->gt[0] = ~0ull; }
}


// ---- small rounds: three work items per pair ("roles") ---------------------------------------------------
// When a round has fewer pairs than the GPU has threads, one thread per pair is a ~12-multiplication serial
// chain on an otherwise idle machine.  Small rounds therefore split every pair into three independent items:
//   role 0: binds the low entries  (a0, b0, c0), stores them, contributes E*(a0 b0 - c0) to t(0)
//   role 1: binds the high entries (a1, b1, c1), stores them, contributes E*(a1 b1 - c1) to t(1)
//   role 2: binds the DIFFERENCES (bind is linear: a1-a0 = bind(A[j+P]-A[j], A[j+3P]-A[j+2P])), contributes to t(inf)
// Items of one pair read the same source entries, so binding cannot be in place here: small rounds ping-pong
// between the table and a scratch copy (src -> dst).
template <bool FUSED>
__device__ __forceinline__ void cubic_roles(const fe *sA, const fe *sB, const fe *sC, fe *dA, fe *dB, fe *dC, u64 P, const fe &r,
                                            const fe *el, const fe *er, u32 sh, int role, u64 first, u64 stride, fe (&x)[3]) {
  // the role is uniform across a warp (role = warp index mod 3): no divergence, one accumulator per thread
  Fq::acc acc = Fq::acc_zero();
  const u64 mask = ((u64)1 << sh) - 1;
  for (u64 id = first; id < P; id += stride) {
    fe w = ldg_fe_ro(er + (el ? (id & mask) : id));
    if (el) w = Fq::mul(ldg_fe_ro(el + (id >> sh)), w);
    if (role < 2) {
      const u64 o = role ? P : 0;
      fe a, b, c;
      if (FUSED) {
        a = bind_pair(ldg_fe(sA + id + o), ldg_fe(sA + id + o + 2 * P), r); stg_fe(dA + id + o, a);
        b = bind_pair(ldg_fe(sB + id + o), ldg_fe(sB + id + o + 2 * P), r); stg_fe(dB + id + o, b);
        c = bind_pair(ldg_fe(sC + id + o), ldg_fe(sC + id + o + 2 * P), r); stg_fe(dC + id + o, c);
      } else {
        a = ldg_fe(sA + id + o); b = ldg_fe(sB + id + o); c = ldg_fe(sC + id + o);
      }
      Fq::mul_acc(acc, w, Fq::sub(Fq::mul(a, b), c));
    } else {
      fe da, db;
      if (FUSED) {
        da = bind_pair(Fq::sub(ldg_fe(sA + id + P), ldg_fe(sA + id)), Fq::sub(ldg_fe(sA + id + 3 * P), ldg_fe(sA + id + 2 * P)), r);
        db = bind_pair(Fq::sub(ldg_fe(sB + id + P), ldg_fe(sB + id)), Fq::sub(ldg_fe(sB + id + 3 * P), ldg_fe(sB + id + 2 * P)), r);
      } else {
        da = Fq::sub(ldg_fe(sA + id + P), ldg_fe(sA + id));
        db = Fq::sub(ldg_fe(sB + id + P), ldg_fe(sB + id));
      }
      Fq::mul_acc(acc, w, Fq::mul(da, db));
    }
  }
  const fe v = Fq::acc_reduce(acc);
  x[0] = role == 0 ? v : Fq::zero(); x[1] = role == 1 ? v : Fq::zero(); x[2] = role == 2 ? v : Fq::zero();
}

template <bool FUSED>
__global__ void __launch_bounds__(SC_ROLE_THREADS, 1)
k_cubic_round_roles(ScState *st, const fe *sA, const fe *sB, const fe *sC, fe *dA, fe *dB, fe *dC, u64 P, int round1, int l,
                    const fe *el, const fe *er, u32 sh) {
  __shared__ FinSmem sm;
  if (threadIdx.x == 0) atomicMin(&st->gt[0], gtimer());
  fe r;
  if (FUSED) r = ld_state(&st->r[round1 - 2]);
  fe x[3];
  { const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, trios = SC_ROLE_THREADS / 96;
    cubic_roles<FUSED>(sA, sB, sC, dA, dB, dC, P, r, el, er, sh, w % 3, ((u64)blockIdx.x * trios + w / 3) * 32 + lane, (u64)gridDim.x * trios * 32, x); }
  block_sum_fq<3>(x, sm.red);
  if (!publish_and_elect<3>(st, x, sm)) return;
  cubic_finalize(st, round1, l, FUSED ? dA : sA, FUSED ? dB : sB, FUSED ? dC : sC, x, sm);
  if (threadIdx.x == 0) { st->gt[4] = st->gt[2]; st->gt[2] = gtimer(); st->gt[3] = st->gt[0]; st->gt[0] = ~0ull; }
}

// all remaining rounds [round_first, l] in one CTA (tables of <= SC_TAIL_LEN entries going in); ping-pongs between
// (A,B,C) and the scratch copies (A2,B2,C2)
// The CTA has SC_TAIL_THREADS role threads plus one SCALAR WARP (warp SC_TAIL_THREADS/32) that has no pair work: at
// the start of round i+1 its lane 0 runs round i's cubic_bound while the role warps already evaluate round i+1.
__global__ void __launch_bounds__(SC_TAIL_THREADS + 32, 1)
k_cubic_tail(ScState *st, fe *A, fe *B, fe *C, fe *A2, fe *B2, fe *C2, int round_first, int l, const fe *eq_left, const fe *eq_right) {
  __shared__ FinSmem sm;
  const int first_half = l / 2, second_half = l - first_half;
  fe *sA = A, *sB = B, *sC = C, *dA = A2, *dB = B2, *dC = C2;
  const int warp = threadIdx.x >> 5, role = warp % 3;
  const bool scalar_warp = warp >= SC_TAIL_THREADS / 32;
  const u64 slot = (u64)(warp / 3) * 32 + (threadIdx.x & 31), nslots = (SC_TAIL_THREADS / 96) * 32;
  fe r = Fq::zero();
  if (round_first > 1) r = ld_state(&st->r[round_first - 2]);
  for (int round1 = round_first; round1 <= l; round1++) {
    const u64 P = (u64)1 << (l - round1);
    const fe *el = nullptr, *er; u32 sh = 0;
    if (round1 < first_half) {
      const int kl = first_half - round1;
      el = eq_left + (((size_t)1 << kl) - 1); er = eq_right + (((size_t)1 << second_half) - 1); sh = (u32)second_half;
    } else {
      er = eq_right + (((size_t)1 << (l - round1)) - 1);
    }
    fe x[3] = {Fq::zero(), Fq::zero(), Fq::zero()};
    if (threadIdx.x == 0) st->clk[7] = st->clk[0];
    SC_STAMP(0);
    if (threadIdx.x == 0) TT(1, 0);
    if (scalar_warp) {
      if (round1 > round_first && threadIdx.x == SC_TAIL_THREADS) cubic_bound(st, round1 - 1, l, r);
    } else if (round1 > 1) {
      cubic_roles<true>(sA, sB, sC, dA, dB, dC, P, r, el, er, sh, role, slot, nslots, x);
    } else {
      cubic_roles<false>(sA, sB, sC, dA, dB, dC, P, Fq::zero(), el, er, sh, role, slot, nslots, x);
    }
    if (round1 > 1) { fe *t; t = sA; sA = dA; dA = t; t = sB; sB = dB; dB = t; t = sC; sC = dC; dC = t; }   // the bound tables are now the source
    if (threadIdx.x == 0) TT(1, 1);
    __syncthreads();
    block_sum_fq<3>(x, sm.red);
    if (threadIdx.x == 0) TT(1, 2);
    r = cubic_finalize_pre(st, round1, x, sm);
    if (threadIdx.x == 0) TT(1, 4);
    if (round1 == l) cubic_claims(st, sA, sB, sC, r);
    SC_STAMP(6);
  }
}

// evaluation_points_zero_check_round0 (sumcheck.rs:1163-1271): round 1 of a zero-check — t(0) = 0 on a satisfied instance, so only
// t(inf) = sum_x E(x) (A1 - A0)(B1 - B0) is summed (no C reads: half the traffic of a full first round)
__global__ void __launch_bounds__(SC_THREADS) k_zero_check_round0(const fe *A, const fe *B, u64 P, const fe *el, const fe *er, u32 sh, fe *partials) {
  __shared__ fe red[32];
  Fq::acc acc = Fq::acc_zero();
  const u64 mask = ((u64)1 << sh) - 1;
  for (u64 id = (u64)blockIdx.x * SC_THREADS + threadIdx.x; id < P; id += (u64)gridDim.x * SC_THREADS) {
    fe w = ldg_fe_ro(er + (el ? (id & mask) : id));
    if (el) w = Fq::mul(ldg_fe_ro(el + (id >> sh)), w);
    const fe a0 = ldg_fe(A + id), b0 = ldg_fe(B + id);
    Fq::mul_acc(acc, w, Fq::mul(Fq::sub(ldg_fe(A + id + P), a0), Fq::sub(ldg_fe(B + id + P), b0)));
  }
  fe x[1] = {Fq::acc_reduce(acc)};
  block_sum_fq<1>(x, red);
  if (threadIdx.x == 0) stg_fe(partials + blockIdx.x, x[0]);
}
// s(X) = l(X) t(X) with t(0) = t(1) = 0: eval_0 = 0, eval_2 = 2 l(2) t_inf, eval_3 = 6 l(3) t_inf  (the values derive_from_claim /
// its fallback produce from (t_0, t_inf, claim) = (0, t_inf, 0), sumcheck.rs:1244-1270)
__global__ void __launch_bounds__(256) k_zero_check_finish(const ScState *st, const fe *partials, u32 nb, fe *out3) {
  __shared__ fe red[32];
  fe x[1] = {Fq::zero()};
  for (u32 b = threadIdx.x; b < nb; b += blockDim.x) x[0] = Fq::add(x[0], ldg_fe(partials + b));
  block_sum_fq<1>(x, red);
  if (threadIdx.x == 0) {
    const fe tau = ldg_fe(&st->taus[0]);
    const fe l0 = Fq::sub(Fq::one(), tau), sl = Fq::sub(tau, l0);
    const fe l2 = Fq::add(l0, Fq::dbl(sl)), l3 = Fq::add(l2, sl);
    const fe t2 = Fq::dbl(x[0]), t6 = Fq::add(Fq::dbl(t2), t2);
    stg_fe(out3, Fq::zero()); stg_fe(out3 + 1, Fq::mul(l2, t2)); stg_fe(out3 + 2, Fq::mul(l3, t6));
  }
}

// init: the split-eq prefix tables and the round-1 constants
__global__ void __launch_bounds__(1024) k_cubic_init(ScState *st, int l, fe *eq_left, fe *eq_right) {
  const int first_half = l / 2, second_half = l - first_half;
  if (blockIdx.x == 0) {
    eq_prefix_block(st->taus + 1, first_half > 0 ? first_half - 1 : 0, eq_left);
    if (threadIdx.x == 0) {
      const fe tau = ldg_fe(&st->taus[0]);
      const fe l0 = Fq::sub(Fq::one(), tau);
      st->gt[0] = ~0ull;
      stg_fe(&st->p, Fq::one());
      stg_fe(&st->L0, l0);
      stg_fe(&st->SL, Fq::sub(tau, l0));
    }
  } else {
    eq_prefix_block(st->taus + first_half, second_half, eq_right);
  }
}

// ---------------------------------------------------------------------------------------------
// quadratic: eval0 = sum a_lo b_lo, tinf = sum (a_hi-a_lo)(b_hi-b_lo)   (sumcheck.rs:128-174)
// ---------------------------------------------------------------------------------------------
// split like the cubic finaliser: quad_finalize_pre leaves r with every thread (and the round's (e0, b, tinf) in
// sm.g[0..2] for quad_claim); the running-claim update (two serial multiplications, read by the next FINALISER only)
// and the final claims run off the critical path
__device__ __forceinline__ fe quad_finalize_pre(ScState *st, int round1, const fe (&x)[2], FinSmem &sm) {
  const int tid = threadIdx.x, i = round1 - 1;
  fe canon = Fq::zero();
  SC_STAMP(1);
  if (tid < 32) {
    // from_evals([e0, claim-e0, 2claim-3e0+2tinf]) = [e0, claim - 2 e0 - tinf, tinf]  (sumcheck.rs:205-216)
    const fe e0 = x[0], ti = x[1];
    const fe b = Fq::sub(Fq::sub(ld_state(&st->claim), Fq::dbl(e0)), ti);
    if (tid < 3) stg_fe(&st->polys[4 * i + tid], fe_sel(tid == 0, e0, fe_sel(tid == 1, b, ti)));
    if (tid < 2) canon = Fq::from_mont(fe_sel(tid == 0, e0, ti));
    if (tid == 0) { sm.g[0] = e0; sm.g[1] = b; sm.g[2] = ti; }
  }
  const fe r = sc_squeeze(st, sm, canon, 2);      // (its barriers publish sm.g[0..2] to the CTA)
  if (tid == 0) { stg_fe(&st->r[i], r); __threadfence(); *(volatile u32 *)&st->r_ready = (u32)round1; }
  return r;
}
// claim <- poly(r); one thread; (e0, b, tinf) by value (copied out of sm.g before the next round overwrites it)
__device__ __forceinline__ void quad_claim(ScState *st, const fe &e0, const fe &b, const fe &ti, const fe &r) {
  stg_fe(&st->claim, Fq::add(e0, mul_ni(r, Fq::add(b, mul_ni(r, ti)))));
}
__device__ __forceinline__ void quad_claims(ScState *st, const fe *A, const fe *B, const fe &r) {
  const int tid = threadIdx.x;
  if (tid == 64 || tid == 96) {
    const fe *T = tid == 64 ? A : B;
    const fe lo = ld_state(T), hi = ld_state(T + 1);
    stg_fe(&st->claims[(tid - 64) / 32], Fq::add(lo, mul_ni(Fq::sub(hi, lo), r)));
  }
}
__device__ __forceinline__ void quad_finalize(ScState *st, int round1, int rounds, const fe *A, const fe *B,
                                              const fe (&x)[2], FinSmem &sm) {
  const fe r = quad_finalize_pre(st, round1, x, sm);
  if (threadIdx.x == 0) quad_claim(st, sm.g[0], sm.g[1], sm.g[2], r);
  if (round1 == rounds) quad_claims(st, A, B, r);
  SC_STAMP(6);
}

// `nvalid`: entries at index >= nvalid are not materialised and read as zero — the zero-suffix awareness of
// the reference's MultilinearPolynomial{lo_eff, hi_eff} (multilinear.rs:36-43, 106-163; sumcheck.rs:136), used by
// the Spartan inner sum-check whose virtual 2M-entry tables are non-zero only in the first M + num_extra slots
// (spartan.rs:330-384 does that round by hand).
__device__ __forceinline__ fe ldg_fe_valid(const fe *T, u64 idx, u64 nvalid) { return idx < nvalid ? ldg_fe(T + idx) : Fq::zero(); }
template <bool FUSED>
__device__ __forceinline__ void quad_body(fe *A, fe *B, u64 P, const fe &r, u64 first, u64 stride, u64 nvalid, fe (&x)[2]) {
  Fq::acc acc0 = Fq::acc_zero(), acci = Fq::acc_zero();
  for (u64 id = first; id < P; id += stride) {
    fe a0, a1, b0, b1;
    if (FUSED) {
      const fe a00 = ldg_fe(A + id), a01 = ldg_fe(A + id + P), a10 = ldg_fe_valid(A, id + 2 * P, nvalid), a11 = ldg_fe_valid(A, id + 3 * P, nvalid);
      const fe b00 = ldg_fe(B + id), b01 = ldg_fe(B + id + P), b10 = ldg_fe_valid(B, id + 2 * P, nvalid), b11 = ldg_fe_valid(B, id + 3 * P, nvalid);
      a0 = bind_pair(a00, a10, r); a1 = bind_pair(a01, a11, r);
      stg_fe(A + id, a0); stg_fe(A + id + P, a1);
      b0 = bind_pair(b00, b10, r); b1 = bind_pair(b01, b11, r);
      stg_fe(B + id, b0); stg_fe(B + id + P, b1);
    } else {
      a0 = ldg_fe(A + id); b0 = ldg_fe(B + id);
      if (id + P >= nvalid) {
        // both high entries are unmaterialised zeros: (0 - a0)(0 - b0) = a0 b0 — ONE product feeds both sums (the first round of the
        // Spartan inner sum-check, where the whole upper half but a few entries is zero: spartan.rs:330-384)
        Fq::mul_acc2(acc0, acci, a0, b0);
        continue;
      }
      a1 = ldg_fe(A + id + P); b1 = ldg_fe(B + id + P);
    }
    Fq::mul_acc(acc0, a0, b0);
    Fq::mul_acc(acci, Fq::sub(a1, a0), Fq::sub(b1, b0));
  }
  x[0] = Fq::acc_reduce(acc0); x[1] = Fq::acc_reduce(acci);
}

template <bool FUSED>
__global__ void __launch_bounds__(SC_THREADS, 2)
k_quad_round(ScState *st, fe *A, fe *B, u64 P, int round1, int rounds, u64 nvalid, DevComm dc) {
  __shared__ FinSmem sm;
  fe r;
  if (FUSED) r = ld_state(&st->r[round1 - 2]);
  fe x[2];
  quad_body<FUSED>(A, B, P, r, (u64)blockIdx.x * SC_THREADS + threadIdx.x, (u64)gridDim.x * SC_THREADS, nvalid, x);
  block_sum_fq<2>(x, sm.red);
  if (!publish_and_elect<2>(st, x, sm)) return;
  if (dc.n > 1) exchange_sums<2>(dc, round1, x, sm);
  quad_finalize(st, round1, rounds, A, B, x, sm);
}

// small rounds of the quadratic prover: role 0 binds/stores the low entries and sums a0*b0, role 1 binds/stores the
// high entries, role 2 binds the differences and sums da*db (see cubic_roles)
template <bool FUSED>
__device__ __forceinline__ void quad_roles(const fe *sA, const fe *sB, fe *dA, fe *dB, u64 P, const fe &r, int role, u64 first, u64 stride,
                                           u64 nvalid, fe (&x)[2]) {
  Fq::acc acc = Fq::acc_zero();
  for (u64 id = first; id < P; id += stride) {
    if (role < 2) {
      const u64 o = role ? P : 0;
      if (FUSED) {
        const fe a = bind_pair(ldg_fe(sA + id + o), ldg_fe_valid(sA, id + o + 2 * P, nvalid), r); stg_fe(dA + id + o, a);
        const fe b = bind_pair(ldg_fe(sB + id + o), ldg_fe_valid(sB, id + o + 2 * P, nvalid), r); stg_fe(dB + id + o, b);
        if (role == 0) Fq::mul_acc(acc, a, b);
      } else if (role == 0) {
        Fq::mul_acc(acc, ldg_fe(sA + id), ldg_fe(sB + id));
      }
    } else {
      fe da, db;
      if (FUSED) {
        da = bind_pair(Fq::sub(ldg_fe(sA + id + P), ldg_fe(sA + id)), Fq::sub(ldg_fe_valid(sA, id + 3 * P, nvalid), ldg_fe_valid(sA, id + 2 * P, nvalid)), r);
        db = bind_pair(Fq::sub(ldg_fe(sB + id + P), ldg_fe(sB + id)), Fq::sub(ldg_fe_valid(sB, id + 3 * P, nvalid), ldg_fe_valid(sB, id + 2 * P, nvalid)), r);
      } else {
        da = Fq::sub(ldg_fe_valid(sA, id + P, nvalid), ldg_fe(sA + id));
        db = Fq::sub(ldg_fe_valid(sB, id + P, nvalid), ldg_fe(sB + id));
      }
      Fq::mul_acc(acc, da, db);
    }
  }
  const fe v = Fq::acc_reduce(acc);
  x[0] = role == 0 ? v : Fq::zero(); x[1] = role == 2 ? v : Fq::zero();
}

template <bool FUSED>
__global__ void __launch_bounds__(SC_ROLE_THREADS, 1)
k_quad_round_roles(ScState *st, const fe *sA, const fe *sB, fe *dA, fe *dB, u64 P, int round1, int rounds, u64 nvalid) {
  __shared__ FinSmem sm;
  fe r;
  if (FUSED) r = ld_state(&st->r[round1 - 2]);
  fe x[2];
  { const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, trios = SC_ROLE_THREADS / 96;
    quad_roles<FUSED>(sA, sB, dA, dB, P, r, w % 3, ((u64)blockIdx.x * trios + w / 3) * 32 + lane, (u64)gridDim.x * trios * 32, nvalid, x); }
  block_sum_fq<2>(x, sm.red);
  if (!publish_and_elect<2>(st, x, sm)) return;
  quad_finalize(st, round1, rounds, FUSED ? dA : sA, FUSED ? dB : sB, x, sm);
}

__global__ void __launch_bounds__(SC_TAIL_THREADS + 32, 1)
k_quad_tail(ScState *st, fe *A, fe *B, fe *A2, fe *B2, int round_first, int rounds, u64 nvalid) {
  __shared__ FinSmem sm;
  fe *sA = A, *sB = B, *dA = A2, *dB = B2;
  const int warp = threadIdx.x >> 5, role = warp % 3;
  const bool scalar_warp = warp >= SC_TAIL_THREADS / 32;     // see k_cubic_tail
  const u64 slot = (u64)(warp / 3) * 32 + (threadIdx.x & 31), nslots = (SC_TAIL_THREADS / 96) * 32;
  fe r = Fq::zero();
  if (round_first > 1) r = ld_state(&st->r[round_first - 2]);
  for (int round1 = round_first; round1 <= rounds; round1++) {
    const u64 P = (u64)1 << (rounds - round1);
    fe x[2] = {Fq::zero(), Fq::zero()};
    if (threadIdx.x == 0) st->clk[7] = st->clk[0];
    SC_STAMP(0);
    // only the first two launches can see unmaterialised entries; afterwards the bound table is dense
    const u64 nv = round1 <= 2 ? nvalid : ~0ull;
    if (scalar_warp) {
      if (round1 > round_first && threadIdx.x == SC_TAIL_THREADS) quad_claim(st, sm.g[0], sm.g[1], sm.g[2], r);
    } else if (round1 > 1) {
      quad_roles<true>(sA, sB, dA, dB, P, r, role, slot, nslots, nv, x);
    } else {
      quad_roles<false>(sA, sB, dA, dB, P, Fq::zero(), role, slot, nslots, nv, x);
    }
    if (round1 > 1) { fe *t; t = sA; sA = dA; dA = t; t = sB; sB = dB; dB = t; }
    __syncthreads();
    block_sum_fq<2>(x, sm.red);
    r = quad_finalize_pre(st, round1, x, sm);
    if (round1 == rounds) quad_claims(st, sA, sB, r);
    SC_STAMP(6);
  }
}

// ---- pipelined single-CTA tails ------------------------------------------------------------------------------------------
// In the tail a round costs ~27k cycles, of which the Keccak squeeze is ~12k and the pair work ~9k — and the pair work waits
// for the challenge.  It does not have to: binding is linear in r, so the sums of round i+1, taken over the table bound to
// r_i, are QUADRATIC POLYNOMIALS in r_i whose coefficients depend on the unbound table only.  The pipelined tails compute,
// during round i's squeeze, the coefficient sums of round i+1 (Karatsuba form: sum x y, sum (x+dx)(y+dy), sum dx dy for each
// bilinear sum); when r_i arrives a finaliser warp evaluates the round-(i+1) sums with two multiplications and goes straight
// to the next squeeze, while the role warps bind the table to r_i and take the coefficients of round i+2 under that squeeze.
// Only the transcript, the round algebra and the challenge stay on a round's critical path.  Same sums, same field
// elements: bit-identical (tests/test_gpu_sumcheck.py, every l).  CTA = 12 role warps + 4 finaliser warps (128 registers).
constexpr int TP_ROLE = SC_TAIL_THREADS, TP_FIN = 128, TP_THREADS = TP_ROLE + TP_FIN;
__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct TailSmem {
  u64 m[2][34];               // squeeze inputs of the lo / hi hash, kept across rounds: only the coefficients, the round counter and the state change
  u64 dg[8];                  // lo || hi digests
  fe g[8];
  fe ch;                      // the round's challenge
  fe coef[2][9];              // coefficient sums of the next round (double-buffered)
  fe x[3];                    // sums of a directly evaluated round
  fe claim;                   // quadratic prover: the running claim
  fe red[TP_ROLE / 32][3];    // per-warp partial sums
  fe gat[12];                 // multi-CTA rounds: the first round's gathered sums
  u32 round;                  // transcript round counter (written back at the end)
};
// sum of NV values per role thread over the TP_ROLE role threads -> out[0..NV) (shared); all role threads must call
template <int NV>
__device__ __forceinline__ void role_sum(fe (&x)[NV], fe *red, fe *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = TP_ROLE / 32;
  warp_sum_fq_cols<NV>(x);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) red[k * NW + warp] = x[k];
  }
  bar_sync_n(1, TP_ROLE);
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) x[k] = lane < NW ? red[k * NW + lane] : Fq::zero();
    warp_sum_fq_cols<NV>(x);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NV; k++) out[k] = x[k];
    }
  }
  bar_sync_n(1, TP_ROLE);
}
// The finaliser group's transcript (threads [TP_ROLE, TP_THREADS), ft = tid - TP_ROLE, barrier 2).  The squeeze input
//   b"p" || coefficients || "NoDS" || round_le16 || state || b"c" || {0,1}  + Keccak padding          (keccak.rs:33-54, 72-105)
// is kept in shared memory for the whole kernel: tp_msg_init lays down the constant bytes, the state and the counter once,
// and a round only stores its coefficients before the two permutations and the new state / counter after them.
struct TpMsg { int plen, mlen, total, nblocks; };
__device__ __forceinline__ TpMsg tp_msg(int ncoef) {
  TpMsg g; g.plen = 1 + 32 * ncoef; g.mlen = g.plen + 4 + 2 + 64 + 1; g.total = g.mlen + 1; g.nblocks = g.total / 136 + 1; return g;
}
__device__ __forceinline__ void tp_put_round(TailSmem &ts, const TpMsg &g) {
#pragma unroll
  for (int w = 0; w < 2; w++) {
    unsigned char *mb = (unsigned char *)ts.m[w];
    mb[g.plen + 4] = (unsigned char)(ts.round & 0xff); mb[g.plen + 5] = (unsigned char)(ts.round >> 8);
  }
}
__device__ __forceinline__ void tp_msg_init(TailSmem &ts, const ScState *st, int ncoef, int ft) {
  const TpMsg g = tp_msg(ncoef);
  if (ft < 68) ((u64 *)ts.m)[ft] = 0;
  bar_sync_n(2, TP_FIN);
  if (ft < 2) {
    unsigned char *mb = (unsigned char *)ts.m[ft];
    mb[0] = 'p';
    mb[g.plen + 0] = 'N'; mb[g.plen + 1] = 'o'; mb[g.plen + 2] = 'D'; mb[g.plen + 3] = 'S';
    mb[g.mlen - 1] = 'c';
    mb[g.mlen] = (unsigned char)ft;                // 0x00 -> lo half, 0x01 -> hi half
    mb[g.total] ^= 0x01;                           // Keccak (not SHA-3) padding
    mb[g.nblocks * 136 - 1] ^= 0x80;
  }
  if (ft == 32) ts.round = st->ts.round;
  if (ft >= 64) {
    const unsigned char b = st->ts.state[ft - 64];
    ((unsigned char *)ts.m[0])[g.plen + 6 + ft - 64] = b; ((unsigned char *)ts.m[1])[g.plen + 6 + ft - 64] = b;
  }
  bar_sync_n(2, TP_FIN);
  if (ft == 32) tp_put_round(ts, g);
  bar_sync_n(2, TP_FIN);
}
// transcript hand-off back to global memory (one thread group, after the last round)
__device__ __forceinline__ void tp_msg_store(const TailSmem &ts, ScState *st, int ncoef, int ft) {
  const TpMsg g = tp_msg(ncoef);
  if (ft == 32) st->ts.round = ts.round;
  if (ft >= 64) st->ts.state[ft - 64] = ((const unsigned char *)ts.m[0])[g.plen + 6 + ft - 64];
}
// one absorb-and-squeeze; lanes ft < ncoef hold the canonical coefficients.  The challenge is returned to fin warp 0 (every lane)
// and left in ts.ch for everybody else (visible after the CTA barrier that ends the round).  Fin warps 2, 3 run the two hashes: they sit
// on the SM sub-partitions that the lowest role warps (the only ones with work in the last rounds) use least.
struct TpNoSide { __device__ __forceinline__ void operator()() const {} };
template <class Side0 = TpNoSide, class Side1 = TpNoSide>
__device__ __forceinline__ fe tp_squeeze(TailSmem &ts, const fe &canon, int ncoef, int ft, Side0 side0 = Side0(), Side1 side1 = Side1()) {
  const TpMsg g = tp_msg(ncoef);
  if (ft < ncoef) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const u32 w = canon.v[k];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        unsigned char *q = (unsigned char *)ts.m[h] + 1 + 32 * ft + 4 * k;
        q[0] = (unsigned char)w; q[1] = (unsigned char)(w >> 8); q[2] = (unsigned char)(w >> 16); q[3] = (unsigned char)(w >> 24);
      }
    }
  }
  bar_sync_n(2, TP_FIN);
  if (ft >= 64) {
    const int lane = ft & 31, w = (ft >> 5) - 2;
    const KeccakLane kl = keccak_lane_init(lane);
    u64 s = 0;
    for (int blk = 0; blk < g.nblocks; blk++) {
      if (lane < 17) s ^= ts.m[w][blk * 17 + lane];
      s = keccak_f_warp(s, kl, lane);
    }
    if (lane < 4) ts.dg[w * 4 + lane] = s;
  } else if (ft >= 32) {
    side1();                                       // fin warp 1: work that may run beside the two permutations
  } else {
    side0();                                       // fin warp 0: scalars of the NEXT round that do not need this round's challenge
  }
  bar_sync_n(2, TP_FIN);
  fe ch = Fq::zero();
  if (ft < 32) {                                   // from_uniform: lo * R^2 + hi * R^3 (Montgomery form of lo + 2^256 hi): even lanes lo, odd lanes hi
    const int half = ft & 1;
    fe h;
#pragma unroll
    for (int i = 0; i < 4; i++) { h.v[2 * i] = (u32)ts.dg[4 * half + i]; h.v[2 * i + 1] = (u32)(ts.dg[4 * half + i] >> 32); }
    const fe m = Fq::mul(h, fe_sel(half != 0, Fq::cst_r3(), Fq::cst_r2()));
    ch = Fq::add(m, shfl_xor_fe(m, 1));
    if (ft == 0) ts.ch = ch;
  } else if (ft >= 64) {                           // the next round's header, off the challenge's path: state <- lo || hi
    const unsigned char b = ((const unsigned char *)ts.dg)[ft - 64];
    ((unsigned char *)ts.m[0])[g.plen + 6 + ft - 64] = b; ((unsigned char *)ts.m[1])[g.plen + 6 + ft - 64] = b;
  } else if (ft == 32) {
    ts.round++;
    tp_put_round(ts, g);
  }
  return ch;
}
// c0 + r (mid + r c2) with mid = c22 - c0 - c2 (Karatsuba form of the middle coefficient)
#ifdef SP2_FIN_INLINE_MUL
#define FIN_MUL Fq::mul_inl
#else
#define FIN_MUL Fq::mul
#endif
__device__ __forceinline__ fe eval_karatsuba(const fe &c00, const fe &c22, const fe &cdd, const fe &r) {
  const fe mid = Fq::sub(Fq::sub(c22, c00), cdd);
  return Fq::add(c00, FIN_MUL(r, Fq::add(mid, FIN_MUL(r, cdd))));
}

// ---- the finaliser's scalar work: fin warp 0, every lane calls, __syncwarp only ---------------------------------------------------
// Everything a round needs from the previous challenge r is a polynomial of degree <= 2 in r whose coefficients are ready before r is:
// the sums (coefficient form), and for the cubic prover L0_i = L0_(i-1) l0_i + SL_(i-1) l0_i r, SL_i likewise, p_i = L0_(i-1) + SL_(i-1) r
// (l_i(X) = l0_i + sl_i X; sumcheck.rs:1399-1405), for the quadratic prover the claim e0 + r (b + r t_inf).  One SIMD evaluation (two dependent
// multiplications) gives all of them, one more multiplication the round polynomial: three multiplications and a from_mont between the
// challenge and the next absorb.
struct FinCubic {
  fe lin[3][3];                                     // (c00, c22, cdd) of L0, SL, p of the coming round (eval_karatsuba form)
  fe h[6];
  fe l0[SC_MAX_ROUNDS], sl[SC_MAX_ROUNDS];          // l_i(X) = l0[i-1] + sl[i-1] X
};
__device__ __forceinline__ void fin_cubic_init(FinCubic &fc, const ScState *st, int l, int ft) {   // the whole finaliser group; barrier afterwards
  if (ft < l) { const fe tau = ld_state(&st->taus[ft]); const fe l0 = Fq::sub(Fq::one(), tau); fc.l0[ft] = l0; fc.sl[ft] = Fq::sub(tau, l0); }
  if (ft >= 64 && ft < 67) {
    const fe v = ld_state(ft == 64 ? &st->L0 : ft == 65 ? &st->SL : &st->p);
    fc.lin[ft - 64][0] = v; fc.lin[ft - 64][1] = v; fc.lin[ft - 64][2] = Fq::zero();
  }
}
// the round's message: sums from ts.x (direct) or from the coefficient triples at r; leaves t0, t1, tinf, L0, SL, p in ts.g[0..6); returns
// the canonical transcript coefficients in lanes 0..2
__device__ __forceinline__ fe fin_cubic_msg(TailSmem &ts, FinCubic &fc, ScState *st, int round1, bool direct, const fe *coef, const fe &r, int lane) {
  fe c00 = Fq::zero(), c22 = Fq::zero(), cdd = Fq::zero();
  if (lane == 0) TT(0, 0);
  if (lane < 3) {
    if (direct) { c00 = ts.x[lane]; c22 = c00; }
    else { c00 = coef[3 * lane]; c22 = coef[3 * lane + 1]; cdd = coef[3 * lane + 2]; }
  } else if (lane < 6) { c00 = fc.lin[lane - 3][0]; c22 = fc.lin[lane - 3][1]; cdd = fc.lin[lane - 3][2]; }
  if (lane == 0) TT(0, 1);
  const fe val = eval_karatsuba(c00, c22, cdd, r);
  if (lane < 6) ts.g[lane] = val;
  __syncwarp();
  if (lane == 0) TT(0, 2);
  const fe t0 = ts.g[0], t1 = ts.g[1], tinf = ts.g[2];
  const fe tb = Fq::sub(Fq::sub(t1, t0), tinf);
  {  // L0*t0, L0*tb, SL*t0, L0*tinf, SL*tb, SL*tinf
    const bool use_sl = (lane == 2) | (lane == 4) | (lane == 5);
    const fe v = fe_sel(lane == 0 || lane == 2, t0, fe_sel(lane == 1 || lane == 4, tb, tinf));
    const fe pr = FIN_MUL(fe_sel(use_sl, ts.g[4], ts.g[3]), v);
    if (lane < 6) fc.h[lane] = pr;
  }
  __syncwarp();
  if (lane == 0) TT(0, 3);
  fe canon = Fq::zero();
  if (lane < 4) {                                   // lanes 0, 1, 2: coefficients 0, 2, 3 (the transcript's); lane 3: the linear one
    const int idx = lane == 0 ? 0 : lane == 1 ? 2 : lane == 2 ? 3 : 1;
    // (h[0] | h[1] + h[2] | h[3] + h[4] | h[5]) as x + y with y = 0 for the outer two
    const fe co = Fq::add(fc.h[idx == 0 ? 0 : idx == 1 ? 1 : idx == 2 ? 3 : 5], fe_sel(idx == 1 || idx == 2, fc.h[idx == 1 ? 2 : 4], Fq::zero()));
    stg_fe(&st->polys[4 * (round1 - 1) + idx], co);
    canon = Fq::from_mont(co);
  }
  __syncwarp();
  if (lane == 0) TT(0, 4);
  return canon;
}
// beside the squeeze: the coming round's L0, SL, p as polynomials in this round's challenge
__device__ __forceinline__ void fin_cubic_next(TailSmem &ts, FinCubic &fc, int round1, int l, int lane) {
  if (round1 < l) {
    const fe m = Fq::mul((lane & 1) ? ts.g[4] : ts.g[3], (lane & 2) ? fc.sl[round1] : fc.l0[round1]);   // L0 l0', SL l0', L0 sl', SL sl'
    if (lane < 4) fc.h[lane] = m;
    __syncwarp();
    if (lane < 2) { const fe a = fc.h[2 * lane], b = fc.h[2 * lane + 1]; fc.lin[lane][0] = a; fc.lin[lane][1] = Fq::add(a, b); fc.lin[lane][2] = Fq::zero(); }
  }
  if (lane == 2) { fc.lin[2][0] = ts.g[3]; fc.lin[2][1] = Fq::add(ts.g[3], ts.g[4]); fc.lin[2][2] = Fq::zero(); }
  __syncwarp();
}
// hand-off after the last round of a kernel: p (and L0, SL of the round after it) back to the state
__device__ __forceinline__ void fin_cubic_store(FinCubic &fc, ScState *st, int last, int l, const fe &r, int lane) {
  if (lane < 3) {
    const fe v = eval_karatsuba(fc.lin[lane][0], fc.lin[lane][1], fc.lin[lane][2], r);
    if (lane == 2) stg_fe(&st->p, v);
    else if (last < l) stg_fe(lane == 0 ? &st->L0 : &st->SL, v);
  }
}
struct FinQuad { fe lin[3]; };                      // the running claim as a polynomial in the coming challenge
__device__ __forceinline__ void fin_quad_init(FinQuad &fq, const ScState *st, int ft) {
  if (ft == 64) { const fe v = ld_state(&st->claim); fq.lin[0] = v; fq.lin[1] = v; fq.lin[2] = Fq::zero(); }
}
__device__ __forceinline__ fe fin_quad_msg(TailSmem &ts, FinQuad &fq, ScState *st, int round1, bool direct, const fe *coef, const fe &r, int lane) {
  fe c00 = Fq::zero(), c22 = Fq::zero(), cdd = Fq::zero();
  if (lane < 2) {
    if (direct) { c00 = ts.x[lane]; c22 = c00; }
    else { c00 = coef[3 * lane]; c22 = coef[3 * lane + 1]; cdd = coef[3 * lane + 2]; }
  } else if (lane == 2) { c00 = fq.lin[0]; c22 = fq.lin[1]; cdd = fq.lin[2]; }
  const fe val = eval_karatsuba(c00, c22, cdd, r);
  if (lane < 3) ts.g[lane] = val;                   // e0, tinf, claim
  __syncwarp();
  // from_evals([e0, claim-e0, 2claim-3e0+2tinf]) = [e0, claim - 2 e0 - tinf, tinf]  (sumcheck.rs:205-216)
  const fe e0 = ts.g[0], ti = ts.g[1];
  const fe b = Fq::sub(Fq::sub(ts.g[2], Fq::dbl(e0)), ti);
  if (lane < 3) stg_fe(&st->polys[4 * (round1 - 1) + lane], fe_sel(lane == 0, e0, fe_sel(lane == 1, b, ti)));
  fe canon = Fq::zero();
  if (lane < 2) canon = Fq::from_mont(fe_sel(lane == 0, e0, ti));
  __syncwarp();
  if (lane == 0) { fq.lin[0] = e0; fq.lin[1] = Fq::add(Fq::add(e0, b), ti); fq.lin[2] = ti; }      // claim <- e0 + r (b + r tinf)
  __syncwarp();
  return canon;
}
__device__ __forceinline__ void fin_quad_store(FinQuad &fq, ScState *st, const fe &r, int lane) {
  if (lane == 0) stg_fe(&st->claim, eval_karatsuba(fq.lin[0], fq.lin[1], fq.lin[2], r));
}

__global__ void __launch_bounds__(TP_THREADS, 1)
k_quad_tail_pipe(ScState *st, fe *A, fe *B, fe *A2, fe *B2, int round_first, int rounds, u64 nvalid) {
  __shared__ TailSmem ts;
  __shared__ FinQuad fq;
  const int tid = threadIdx.x, ft = tid - TP_ROLE, lane = tid & 31;
  const bool is_role = tid < TP_ROLE;
  fe *sA = A, *sB = B, *dA = A2, *dB = B2;
  const int warp = tid >> 5, role = warp % 3;
  const u64 slot = (u64)(warp / 3) * 32 + lane, nslots = (TP_ROLE / 96) * 32;
  fe r = Fq::zero();
  if (round_first > 1) r = ld_state(&st->r[round_first - 2]);
  if (!is_role) { fin_quad_init(fq, st, ft); tp_msg_init(ts, st, 2, ft); }
  fe rf = r;                                        // fin warp 0: the previous challenge, in registers
  bool have_coef = false; int cur = 0;
  __syncthreads();
  for (int round1 = round_first; round1 <= rounds; round1++) {
    const u64 P = (u64)1 << (rounds - round1);                 // pairs of this round
    const bool direct = !have_coef;
    const bool want_next = round1 + 1 <= rounds && round1 + 1 > 2;   // rounds 1, 2 may see unmaterialised entries: evaluated directly
    const u64 nv = round1 <= 2 ? nvalid : ~0ull;
    if (is_role) {
      if (direct) {
        fe x[2];
        if (round1 > 1) quad_roles<true>(sA, sB, dA, dB, P, r, role, slot, nslots, nv, x);
        else quad_roles<false>(sA, sB, dA, dB, P, Fq::zero(), role, slot, nslots, nv, x);
        role_sum<2>(x, &ts.red[0][0], ts.x);
        __threadfence_block();
        bar_arrive_n(3, TP_THREADS);                            // the finalisers may start on ts.x
      } else {
        // bind the table to r: dst[j] = src[j] + r (src[j + 2P] - src[j]), j < 2P, both tables
        for (u64 q = (u64)tid; q < 4 * P; q += TP_ROLE) {
          const u64 j = q >> 1;
          const fe *S = (q & 1) ? sB : sA; fe *D = (q & 1) ? dB : dA;
          stg_fe(D + j, bind_pair(ldg_fe(S + j), ldg_fe(S + j + 2 * P), r));
        }
        bar_sync_n(1, TP_ROLE);
      }
      const fe *cA = round1 > 1 ? dA : sA, *cB = round1 > 1 ? dB : sB;     // the table this round's pairs live in (length 2P)
      if (want_next) {
        // coefficients of round1 + 1 (pairs (j, j + P/2) of the table bound to THIS round's challenge):
        //   even warps: sum a0 b0, sum a2 b2, sum (a2-a0)(b2-b0);  odd warps: the same for u = a1 - a0, u + du = a3 - a2
        const u64 Hn = P / 2;
        const int grp = warp & 1, pair = warp >> 1;
        fe xs[3] = {Fq::zero(), Fq::zero(), Fq::zero()};
        if ((u64)pair * 32 < Hn) {                              // warp-uniform: in the last rounds most warps have no pairs
          Fq::acc c0 = Fq::acc_zero(), c2 = Fq::acc_zero(), cd = Fq::acc_zero();
          for (u64 j = (u64)pair * 32 + lane; j < Hn; j += (TP_ROLE / 64) * 32) {
            fe x0, x2, y0, y2;
            if (grp == 0) { x0 = ldg_fe(cA + j); x2 = ldg_fe(cA + j + P); y0 = ldg_fe(cB + j); y2 = ldg_fe(cB + j + P); }
            else {
              x0 = Fq::sub(ldg_fe(cA + j + Hn), ldg_fe(cA + j)); x2 = Fq::sub(ldg_fe(cA + j + P + Hn), ldg_fe(cA + j + P));
              y0 = Fq::sub(ldg_fe(cB + j + Hn), ldg_fe(cB + j)); y2 = Fq::sub(ldg_fe(cB + j + P + Hn), ldg_fe(cB + j + P));
            }
            Fq::mul_acc(c0, x0, y0); Fq::mul_acc(c2, x2, y2); Fq::mul_acc(cd, Fq::sub(x2, x0), Fq::sub(y2, y0));
          }
          xs[0] = Fq::acc_reduce(c0); xs[1] = Fq::acc_reduce(c2); xs[2] = Fq::acc_reduce(cd);
          warp_sum_fq_cols<3>(xs);
        }
        if (lane == 0) { ts.red[warp][0] = xs[0]; ts.red[warp][1] = xs[1]; ts.red[warp][2] = xs[2]; }
        bar_sync_n(1, TP_ROLE);
        if (tid < 6) {                                          // coefficient tid % 3 of group tid / 3: the six warps of that group
          const int g = tid / 3, c = tid % 3;
          fe acc = ts.red[g][c];
#pragma unroll
          for (int q = 1; q < TP_ROLE / 64; q++) acc = Fq::add(acc, ts.red[2 * q + g][c]);
          ts.coef[cur ^ 1][3 * g + c] = acc;
        }
      }
    } else {
      if (direct) bar_sync_n(3, TP_THREADS);
      fe canon = Fq::zero();
      if (ft < 32) canon = fin_quad_msg(ts, fq, st, round1, direct, ts.coef[cur], rf, ft);
      rf = tp_squeeze(ts, canon, 2, ft);
      if (ft == 0) { stg_fe(&st->r[round1 - 1], rf); __threadfence(); *(volatile u32 *)&st->r_ready = (u32)round1; }
    }
    __syncthreads();
    r = ts.ch;
    if (round1 > 1) { fe *t; t = sA; sA = dA; dA = t; t = sB; sB = dB; dB = t; }
    have_coef = want_next; cur ^= 1;
  }
  if (!is_role) { tp_msg_store(ts, st, 2, ft); if (ft < 32) fin_quad_store(fq, st, rf, ft); }
  quad_claims(st, sA, sB, r);
}

__global__ void __launch_bounds__(TP_THREADS, 1)
k_cubic_tail_pipe(ScState *st, fe *A, fe *B, fe *C, fe *A2, fe *B2, fe *C2, int round_first, int l, const fe *eq_left, const fe *eq_right) {
  __shared__ TailSmem ts;
  __shared__ FinCubic fc;
  const int first_half = l / 2, second_half = l - first_half;
  const int tid = threadIdx.x, ft = tid - TP_ROLE, lane = tid & 31;
  const bool is_role = tid < TP_ROLE;
  fe *sA = A, *sB = B, *sC = C, *dA = A2, *dB = B2, *dC = C2;
  const int warp = tid >> 5, role = warp % 3, trio = warp / 3;
  const u64 slot = (u64)trio * 32 + lane, nslots = (TP_ROLE / 96) * 32;
  fe r = Fq::zero();
  if (round_first > 1) r = ld_state(&st->r[round_first - 2]);
  if (!is_role) { fin_cubic_init(fc, st, l, ft); tp_msg_init(ts, st, 3, ft); }
  fe rf = r;                                        // fin warp 0: the previous challenge, in registers
  bool have_coef = false; int cur = 0;
  // split-eq weights of a round (EqSumCheckInstance::poly_eqs_first_half / poly_eq_right_last_half, sumcheck.rs:1007-1023)
  auto weights = [&](int round1, const fe *&el, const fe *&er, u32 &sh) {
    el = nullptr; sh = 0;
    if (round1 < first_half) { const int kl = first_half - round1; el = eq_left + (((size_t)1 << kl) - 1); er = eq_right + (((size_t)1 << second_half) - 1); sh = (u32)second_half; }
    else er = eq_right + (((size_t)1 << (l - round1)) - 1);
  };
  __syncthreads();
  for (int round1 = round_first; round1 <= l; round1++) {
    const u64 P = (u64)1 << (l - round1);
    const bool direct = !have_coef;
    const bool want_next = round1 + 1 <= l;
    if (tid == 0) TT(0, 0);
    if (tid == TP_ROLE) TT(0, 5);
    if (is_role) {
      if (direct) {
        const fe *el, *er; u32 sh; weights(round1, el, er, sh);
        fe x[3];
        if (round1 > 1) cubic_roles<true>(sA, sB, sC, dA, dB, dC, P, r, el, er, sh, role, slot, nslots, x);
        else cubic_roles<false>(sA, sB, sC, dA, dB, dC, P, Fq::zero(), el, er, sh, role, slot, nslots, x);
        role_sum<3>(x, &ts.red[0][0], ts.x);
        __threadfence_block();
        bar_arrive_n(3, TP_THREADS);
      } else {
        // bind the three tables to r: role k binds table k (both halves of the bound table)
        const fe *S = role == 0 ? sA : role == 1 ? sB : sC; fe *D = role == 0 ? dA : role == 1 ? dB : dC;
        for (u64 j = slot; j < 2 * P; j += nslots) stg_fe(D + j, bind_pair(ldg_fe(S + j), ldg_fe(S + j + 2 * P), r));
        bar_sync_n(1, TP_ROLE);
      }
      if (tid == 0) TT(0, 1);
      const fe *cA = round1 > 1 ? dA : sA, *cB = round1 > 1 ? dB : sB, *cC = round1 > 1 ? dC : sC;
      if (want_next) {
        // coefficient sums of round1 + 1 over the pairs (j, j + P/2) of this round's table, weighted by the NEXT round's eq:
        //   role 0: t(0) from the low entries, role 1: t(1) from the high entries, role 2: t(inf) from the differences
        const u64 Hn = P / 2;
        fe xs[3] = {Fq::zero(), Fq::zero(), Fq::zero()};
        if ((u64)trio * 32 < Hn) {                              // warp-uniform: in the last rounds most warps have no pairs
          const fe *el, *er; u32 sh; weights(round1 + 1, el, er, sh);
          const u64 mask = ((u64)1 << sh) - 1;
          Fq::acc c0 = Fq::acc_zero(), c2 = Fq::acc_zero(), cd = Fq::acc_zero();
          for (u64 j = slot; j < Hn; j += nslots) {
            fe w = ldg_fe_ro(er + (el ? (j & mask) : j));
            if (el) w = Fq::mul(ldg_fe_ro(el + (j >> sh)), w);
            fe x0, x2, y0, y2;
            if (role < 2) {
              const u64 o = role ? Hn : 0;
              x0 = ldg_fe(cA + j + o); x2 = ldg_fe(cA + j + o + P); y0 = ldg_fe(cB + j + o); y2 = ldg_fe(cB + j + o + P);
              Fq::mul_acc(c0, w, Fq::sub(Fq::mul(x0, y0), ldg_fe(cC + j + o)));
              Fq::mul_acc(c2, w, Fq::sub(Fq::mul(x2, y2), ldg_fe(cC + j + o + P)));
            } else {
              x0 = Fq::sub(ldg_fe(cA + j + Hn), ldg_fe(cA + j)); x2 = Fq::sub(ldg_fe(cA + j + P + Hn), ldg_fe(cA + j + P));
              y0 = Fq::sub(ldg_fe(cB + j + Hn), ldg_fe(cB + j)); y2 = Fq::sub(ldg_fe(cB + j + P + Hn), ldg_fe(cB + j + P));
              Fq::mul_acc(c0, w, Fq::mul(x0, y0));
              Fq::mul_acc(c2, w, Fq::mul(x2, y2));
            }
            Fq::mul_acc(cd, w, Fq::mul(Fq::sub(x2, x0), Fq::sub(y2, y0)));
          }
          xs[0] = Fq::acc_reduce(c0); xs[1] = Fq::acc_reduce(c2); xs[2] = Fq::acc_reduce(cd);
          if (tid == 0) TT(0, 2);
          warp_sum_fq_cols<3>(xs);
        }
        if (lane == 0) { ts.red[warp][0] = xs[0]; ts.red[warp][1] = xs[1]; ts.red[warp][2] = xs[2]; }
        bar_sync_n(1, TP_ROLE);
        if (tid < 9) {                                          // coefficient tid % 3 of role tid / 3: the four warps of that role
          const int k = tid / 3, c = tid % 3;
          fe acc = ts.red[k][c];
#pragma unroll
          for (int q = 1; q < TP_ROLE / 96; q++) acc = Fq::add(acc, ts.red[3 * q + k][c]);
          ts.coef[cur ^ 1][3 * k + c] = acc;
        }
      }
      if (tid == 0) TT(0, 3);
    } else {
      if (direct) bar_sync_n(3, TP_THREADS);
      fe canon = Fq::zero();
      if (ft < 32) canon = fin_cubic_msg(ts, fc, st, round1, direct, ts.coef[cur], rf, ft);
      if (tid == TP_ROLE) TT(0, 7);
      rf = tp_squeeze(ts, canon, 3, ft, [&] { fin_cubic_next(ts, fc, round1, l, ft); });
      if (ft == 0) stg_fe(&st->r[round1 - 1], rf);
      if (tid == TP_ROLE) TT(0, 8);
    }
    __syncthreads();
    if (tid == 0) TT(0, 4);
    r = ts.ch;
    if (round1 > 1) { fe *t; t = sA; sA = dA; dA = t; t = sB; sB = dB; dB = t; t = sC; sC = dC; dC = t; }
    have_coef = want_next; cur ^= 1;
  }
  if (!is_role) { tp_msg_store(ts, st, 3, ft); if (ft < 32) fin_cubic_store(fc, st, l, l, rf, ft); }
  cubic_claims(st, sA, sB, sC, r);
}

// ---- persistent multi-round kernels ------------------------------------------------------------------------
// All multi-CTA rounds of a sum-check in ONE cooperative launch (one CTA per SM): per round every CTA computes its
// partial sums, arrives at a grid barrier (monotonic counter), CTA 0 sums the partials, runs the finaliser (so the
// finaliser's code and the transcript state stay warm on one SM) and releases the grid by publishing the round
// number; the others spin on it.  Saves, per round, the launch gap (measured 5 us), the last-CTA election and a cold
// finaliser (+5 us) of the one-launch-per-round scheme.

// returns true on CTA 0 with the grid-wide sums in x (warp 0); other CTAs return false after the release
template <int NV>
__device__ __forceinline__ bool persist_gather(ScState *st, u32 seq, fe (&x)[NV], FinSmem &sm) {
  const u32 G = gridDim.x, tid = threadIdx.x;
  if (blockIdx.x != 0) {
    if (tid == 0) {
#pragma unroll
      for (int k = 0; k < NV; k++) stg_fe(&st->partial[3 * blockIdx.x + k], x[k]);
    }
    __syncthreads();                                  // every thread's table stores of this round precede the fence
    if (tid == 0) { __threadfence(); atomicAdd(&st->arrived, 1u); }
    return false;
  }
  if (tid == 0) { const u32 want = (G - 1) * seq; spin_until([=] { return ld_volatile_u32(&st->arrived) >= want; }, &st->err); __threadfence(); }
  __syncthreads();
  fe y[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) y[k] = Fq::zero();
  for (u32 b = 1 + tid; b < G; b += blockDim.x) {
#pragma unroll
    for (int k = 0; k < NV; k++) y[k] = Fq::add(y[k], ld_state(&st->partial[3 * b + k]));
  }
  block_sum_fq<NV>(y, sm.red);
  if (tid < 32) {
#pragma unroll
    for (int k = 0; k < NV; k++) x[k] = Fq::add(x[k], y[k]);
  }
  return true;
}
__device__ __forceinline__ void persist_release(ScState *st, u32 seq) {       // CTA 0, after the finaliser
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); *(volatile u32 *)&st->released = seq; }
}
__device__ __forceinline__ void persist_wait(ScState *st, u32 seq) {          // other CTAs
  if (threadIdx.x == 0) { spin_until([=] { return ld_volatile_u32(&st->released) >= seq; }, &st->err); __threadfence(); }
  __syncthreads();
}

struct PersistCubic {
  ScState *st; fe *A, *B, *C, *A2, *B2, *C2; int l, round_first, round_end; const fe *eq_left, *eq_right;
  const ScTinvMail *mail; u32 mail_epoch;      // tau inverses for the derived t(1) (nullptr: ScState::derive_rounds must be 0)
};

#ifndef SC_PERSIST_MINB
#define SC_PERSIST_MINB 1
#endif
__global__ void __launch_bounds__(SC_THREADS, SC_PERSIST_MINB) k_cubic_persist(PersistCubic a) {
  __shared__ FinSmem sm;
  ScState *st = a.st;
  const int l = a.l, first_half = l / 2, second_half = l - first_half;
  const u32 G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
  fe *sA = a.A, *sB = a.B, *sC = a.C, *dA = a.A2, *dB = a.B2, *dC = a.C2;
  u32 seq = 0;
  const int derive_rounds = a.mail ? (int)ld_volatile_u32(&st->derive_rounds) : 0;   // set before the launch (k_gate_taus)
  for (int round1 = a.round_first; round1 < a.round_end; round1++) {
    const bool fused = round1 > 1;
    const bool derive = round1 <= derive_rounds;                          // t(1) from the claim: two sums instead of three
    const u64 P = (u64)1 << (l - round1);
    const u64 len_in = fused ? 4 * P : 2 * P;
    const bool in_first = round1 < first_half;
    const fe *el = nullptr, *er; u32 out_len = 1, sh = 0;
    if (in_first) {
      const int kl = first_half - round1;
      el = a.eq_left + (((size_t)1 << kl) - 1); out_len = 1u << kl;
      er = a.eq_right + (((size_t)1 << second_half) - 1); sh = (u32)second_half;
    } else {
      er = a.eq_right + (((size_t)1 << (l - round1)) - 1);
    }
    if (bid == 0 && tid == 0) st->prof[round1 - 1][0] = gtimer();
    fe r = Fq::zero();
    if (fused) r = ld_state(&st->r[round1 - 2]);
    fe x[3];
    const bool roles = len_in <= SC_ROLE_LEN;
    if (roles) {
      // 256 threads = 2 role trios (warps 0..5); warps 6, 7 idle
      const int w = tid >> 5, lane = tid & 31;
      if (w < 6) {
        const u64 first = ((u64)bid * 2 + w / 3) * 32 + lane, stride = (u64)G * 64;
        if (fused) cubic_roles<true>(sA, sB, sC, dA, dB, dC, P, r, el, er, sh, w % 3, first, stride, x);
        else cubic_roles<false>(sA, sB, sC, dA, dB, dC, P, r, el, er, sh, w % 3, first, stride, x);
      } else { x[0] = Fq::zero(); x[1] = Fq::zero(); x[2] = Fq::zero(); }
    } else {
      const u64 in_len = (u64)1 << sh;
      if (in_first && in_len >= SC_THREADS && in_len / SC_THREADS <= G) {
        const u32 gx = (u32)(in_len / SC_THREADS); u32 gy = G / gx; if (gy > out_len) gy = out_len;
        const u32 bx = bid % gx, by = bid / gx;
        Fq::acc acc0 = Fq::acc_zero(), acc1 = Fq::acc_zero(), acci = Fq::acc_zero();
        const u64 xi = (u64)bx * SC_THREADS + tid;
        if (by < gy) {
          if (fused) { for (u32 xo = by; xo < out_len; xo += gy) cubic_pair<true>(sA, sB, sC, ((u64)xo << sh) | xi, P, r, ldg_fe_ro(el + xo), acc0, acc1, acci, derive); }
          else { for (u32 xo = by; xo < out_len; xo += gy) cubic_pair<false>(sA, sB, sC, ((u64)xo << sh) | xi, P, r, ldg_fe_ro(el + xo), acc0, acc1, acci, derive); }
        }
        const fe wr = ldg_fe_ro(er + xi);
        x[0] = Fq::mul(wr, Fq::acc_reduce(acc0)); x[1] = Fq::mul(wr, Fq::acc_reduce(acc1)); x[2] = Fq::mul(wr, Fq::acc_reduce(acci));
      } else {
        if (fused) cubic_generic<true>(sA, sB, sC, P, r, el, er, sh, (u64)bid * SC_THREADS + tid, (u64)G * SC_THREADS, x, 0, 0, derive);
        else cubic_generic<false>(sA, sB, sC, P, r, el, er, sh, (u64)bid * SC_THREADS + tid, (u64)G * SC_THREADS, x, 0, 0, derive);
      }
    }
    block_sum_fq<3>(x, sm.red);
    seq++;
    const bool swapped = roles && fused;
    if (swapped) { fe *t; t = sA; sA = dA; dA = t; t = sB; sB = dB; dB = t; t = sC; sC = dC; dC = t; }
    if (bid == 0 && tid == 0) st->prof[round1 - 1][1] = gtimer();
    if (bid == 0 && derive && round1 == a.round_first && tid < 32) derive_prefetch(st, sm, a.mail, a.mail_epoch, a.round_first, derive_rounds);
    if (persist_gather<3>(st, seq, x, sm)) {
      if (tid == 0) st->prof[round1 - 1][2] = gtimer();
      const fe rn = cubic_finalize_pre(st, round1, x, sm, derive);
      if (tid == 0) st->prof[round1 - 1][3] = gtimer();
      persist_release(st, seq);
      // off the critical path: the grid is already streaming the next round (warps 6, 7 have no pair work in role rounds)
      if (tid == SC_THREADS - 32) cubic_bound(st, round1, l, rn, round1 + 1 <= derive_rounds ? &sm : nullptr);
      if (round1 == l) cubic_claims(st, sA, sB, sC, rn);
    } else {
      persist_wait(st, seq);
    }
  }
}

struct PersistQuad { ScState *st; fe *A, *B, *A2, *B2; int rounds, round_first, round_end; u64 nvalid; };

__global__ void __launch_bounds__(SC_THREADS, SC_PERSIST_MINB) k_quad_persist(PersistQuad a) {
  __shared__ FinSmem sm;
  ScState *st = a.st;
  const u32 G = gridDim.x, bid = blockIdx.x, tid = threadIdx.x;
  fe *sA = a.A, *sB = a.B, *dA = a.A2, *dB = a.B2;
  u32 seq = 0;
  for (int round1 = a.round_first; round1 < a.round_end; round1++) {
    const bool fused = round1 > 1;
    const u64 P = (u64)1 << (a.rounds - round1);
    const u64 len_in = fused ? 4 * P : 2 * P;
    const u64 nv = round1 <= 2 ? a.nvalid : ~0ull;
    fe r = Fq::zero();
    if (fused) r = ld_state(&st->r[round1 - 2]);
    fe x[2];
    const bool roles = len_in <= SC_ROLE_LEN;
    if (roles) {
      const int w = tid >> 5, lane = tid & 31;
      if (w < 6) {
        const u64 first = ((u64)bid * 2 + w / 3) * 32 + lane, stride = (u64)G * 64;
        if (fused) quad_roles<true>(sA, sB, dA, dB, P, r, w % 3, first, stride, nv, x);
        else quad_roles<false>(sA, sB, dA, dB, P, r, w % 3, first, stride, nv, x);
      } else { x[0] = Fq::zero(); x[1] = Fq::zero(); }
    } else {
      if (fused) quad_body<true>(sA, sB, P, r, (u64)bid * SC_THREADS + tid, (u64)G * SC_THREADS, nv, x);
      else quad_body<false>(sA, sB, P, r, (u64)bid * SC_THREADS + tid, (u64)G * SC_THREADS, nv, x);
    }
    block_sum_fq<2>(x, sm.red);
    seq++;
    if (roles && fused) { fe *t; t = sA; sA = dA; dA = t; t = sB; sB = dB; dB = t; }
    if (persist_gather<2>(st, seq, x, sm)) {
      const fe rn = quad_finalize_pre(st, round1, x, sm);
      persist_release(st, seq);
      if (tid == SC_THREADS - 32) quad_claim(st, sm.g[0], sm.g[1], sm.g[2], rn);
      if (round1 == a.rounds) quad_claims(st, sA, sB, rn);
    } else {
      persist_wait(st, seq);
    }
  }
}

// ---- pipelined MULTI-CTA rounds ("mid" rounds: tables of 2^11 .. 2^17 entries) ------------------------------------------
// Between the streaming rounds (bound by the integer pipe) and the single-CTA tail sit the rounds whose tables are too long for
// one SM and too short to amortise a grid barrier: in k_*_persist such a round costs ~21 us — ~5.5 us of pair work, ~5.5 us until
// every CTA's partial sums are in, ~9 us of finaliser — of which only the finaliser has to be serial.  The mid kernels run the
// pipelined tail (above) on G = 2^k CTAs at once:
//   * the hypercube is split CYCLICALLY across the CTAs (CTA c owns the indices = c mod G, stored densely in ITS shared memory):
//     both entries of every bind pair share their low bits, so binding and pairing never leave the CTA — the multi-GPU split
//     (sumcheck.cuh) applied to SMs.  Only <= 12 partial sums per CTA and round cross the chip;
//   * the role warps of every CTA take, as soon as r_(i-1) is published, the bind to r_(i-1) and the COEFFICIENT sums of round
//     i+1 (quadratics in r_i, see the tails), publish them and arrive on a per-round counter;
//   * the finaliser group (CTA 0) evaluates round i from the coefficients gathered one round earlier, absorbs, squeezes and
//     publishes r_i; its second warp gathers the next coefficients while the Keccak warps hash.
// A mid round costs what the transcript costs (~11 us).  Every wait is bounded (spin_until).  Same sums, same field elements:
// bit-identical (tests/test_gpu_sumcheck.py).
constexpr int MP_LOG_LEN_IN = 17;                  // the first mid round reads tables of <= 2^17 entries
#ifndef SP2_MP_LOG_LEN_IN_QUAD
#define SP2_MP_LOG_LEN_IN_QUAD 17
#endif
constexpr int MP_LOG_LEN_IN_QUAD = SP2_MP_LOG_LEN_IN_QUAD;   // ... of the quadratic prover (two tables: 2^18 entries would fit the shared memory)
constexpr int MP_LOG_CTAS = 7;                     // <= 128 CTAs
constexpr int MP_LOG_LOCAL = 8;                    // >= 256 entries per CTA going into the first bind (else fewer CTAs)
// round_last: the last round on the cyclic layout; finish: role CTA 0 then runs the remaining rounds (<= SC_TAIL_LEN entries) on the whole table
// in natural order — the single-CTA tail without a second launch, a direct first round or pair work beside the finaliser — and leaves the claims
struct MidCubic { ScState *st; const fe *A, *B, *C; fe *oA, *oB, *oC; int l, round_first, round_last, k; const fe *eq_left, *eq_right; int finish; };
struct MidQuad { ScState *st; const fe *A, *B; fe *oA, *oB; int rounds, round_first, round_last, k, finish; };

__device__ __forceinline__ void st_volatile_u32(u32 *p, u32 v) { asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
// role warps: wait until the challenge of `round` is published
__device__ __forceinline__ void mid_wait_released(ScState *st, int round) {
  if (threadIdx.x == 0) { const u32 want = (u32)round; spin_until([=] { return ld_volatile_u32(&st->mid_released) >= want; }, &st->err); __threadfence(); }
  bar_sync_n(1, TP_ROLE);
}
// The CTAs publish their partial sums by ATOMIC adds of 16-bit limb halves into u32 accumulators (G * 4 * 2^16 < 2^32): the chip-wide
// reduction happens in the L2 atomic units and the finaliser reads 16 words per value instead of G field elements.  The accumulators of the
// first round live in the state's head (zeroed with it), those of the later rounds behind the partials (zeroed by the finaliser CTA before
// it releases the first challenge).
constexpr int MP_ACC_WORDS = 9 * 16;
constexpr int MP_ARRIVE_FIRST_COEF = SC_MAX_ROUNDS + 1;   // arrival counter of the first round's coefficient sums (its direct sums use the round's own)
__device__ __forceinline__ u32 *mid_acc(ScState *st, int round) { return (u32 *)st->partial + (size_t)round * MP_ACC_WORDS; }
// role warps: per-warp partial sums (`xs`, identical in every lane) -> values [base + 3 g + c] (or [base + g] from c = 0 alone when only_first);
// the warps of group g are g, g + NG, g + 2 NG, ...  (cubic: NG = 3 roles; quadratic: NG = 2 halves)
template <int NG>
__device__ __forceinline__ void mid_publish_acc(TailSmem &ts, const fe (&xs)[3], u32 *acc, int base, bool only_first) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (lane == 0) { ts.red[warp][0] = xs[0]; ts.red[warp][1] = xs[1]; ts.red[warp][2] = xs[2]; }
  bar_sync_n(1, TP_ROLE);
  if (tid < 3 * NG * 16) {
    const int v = tid >> 4, col = tid & 15, g = v / 3, c = v % 3;
    if (!only_first || c == 0) {
      u32 sum = 0;
#pragma unroll
      for (int q = 0; q < TP_ROLE / 32 / NG; q++) { const u32 w = ts.red[NG * q + g][c].v[col >> 1]; sum += (col & 1) ? (w >> 16) : (w & 0xffffu); }
      if (sum) atomicAdd(acc + (base + (only_first ? g : v)) * 16 + col, sum);
    }
    __threadfence();
  }
  bar_sync_n(1, TP_ROLE);
}
// fin warp 1: wait for the G arrivals of `round`, then lane v < nval rebuilds value v from its 16 column sums
__device__ __forceinline__ void mid_gather_acc(ScState *st, int round, int G, const u32 *accs, int nval, fe *out) {
  const int lane = threadIdx.x & 31;
  if (lane == 0) { const u32 want = (u32)G; const u32 *ctr = &st->mid_arrive[round]; spin_until([=] { return ld_volatile_u32(ctr) >= want; }, &st->err); __threadfence(); }
  __syncwarp();
  if (lane < nval) {
    const u32 *acc = accs + lane * 16;
    u32 w[16];
#pragma unroll
    for (int q = 0; q < 4; q++) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4 * q]), "=r"(w[4 * q + 1]), "=r"(w[4 * q + 2]), "=r"(w[4 * q + 3]) : "l"(acc + 4 * q));
    fe x; u64 carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { carry += (u64)w[2 * i] + ((u64)w[2 * i + 1] << 16); x.v[i] = (u32)carry; carry >>= 32; }
    u32 top = Fq::fold_top(x, (u32)carry);
    top = Fq::fold_top(x, top);
    cond_sub_p<FqParams>(x, top);
    cond_sub_p<FqParams>(x, 0);
    out[lane] = x;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(TP_THREADS, 1) k_cubic_mid_pipe(MidCubic a) {
  extern __shared__ __align__(32) unsigned char mp_dyn[];
  __shared__ TailSmem ts;
  __shared__ FinCubic fc;
  ScState *st = a.st;
  // CTAs [0, G): role warps only; CTA G: the finaliser group only (it shares its SM with nobody: no issue-slot contention)
  const int l = a.l, first = a.round_first, last = a.round_last, G = (int)gridDim.x - 1;
  const int last_all = a.finish ? l : last;           // the kernel's last round
  int k = a.k, cta = (int)blockIdx.x - 1;             // the role CTAs' layout: cyclic on 2^k CTAs; (0, 0) = natural order on role CTA 0 after round `last`
  const int first_half = l / 2, second_half = l - first_half;
  const int tid = threadIdx.x, ft = tid - TP_ROLE, lane = tid & 31;
  const bool is_role = tid < TP_ROLE;
  const int warp = tid >> 5, role = warp % 3, trio = warp / 3;
  const u64 slot = (u64)trio * 32 + lane, nslots = (TP_ROLE / 96) * 32;
  // split-eq weights of a round (EqSumCheckInstance::poly_eqs_first_half / poly_eq_right_last_half, sumcheck.rs:1007-1023)
  auto weights = [&](int round1, const fe *&el, const fe *&er, u32 &sh) {
    el = nullptr; sh = 0;
    if (round1 < first_half) { const int kl = first_half - round1; el = a.eq_left + (((size_t)1 << kl) - 1); er = a.eq_right + (((size_t)1 << second_half) - 1); sh = (u32)second_half; }
    else er = a.eq_right + (((size_t)1 << (l - round1)) - 1);
  };
  if (is_role) {
    if (cta < 0) return;
    const u64 n0 = ((u64)2 << (l - first)) >> k;               // local length of the first bound table
    fe *X = (fe *)mp_dyn, *Y = X + 3 * n0;
    fe *cur = X, *nxt = Y; u64 cs = n0, ns = n0 / 2;            // table t of a buffer starts at t * stride
    for (int round1 = first; round1 <= last_all; round1++) {
      if (round1 == last + 1) {
        // the rounds on <= SC_TAIL_LEN entries: role CTA 0 alone, the whole table (written back in natural order by every CTA before its arrival
        // of round `last`) in its shared memory
        if (cta != 0) return;
        if (tid == 0) { const u32 want = (u32)G; const u32 *ctr = &st->mid_arrive[last]; spin_until([=] { return ld_volatile_u32(ctr) >= want; }, &st->err); __threadfence(); }
        bar_sync_n(1, TP_ROLE);
        k = 0;
        const u64 nt = (u64)2 << (l - last);
        cur = X; nxt = X + 3 * nt; cs = nt; ns = nt / 2;
        const fe *S = role == 0 ? a.oA : role == 1 ? a.oB : a.oC;
        for (u64 j = slot; j < nt; j += nslots) cur[role * cs + j] = ldg_fe(S + j);
        bar_sync_n(1, TP_ROLE);
      }
      const u64 P = (u64)1 << (l - round1), Pl = P >> k, Hl = Pl >> 1;
      const bool want_next = round1 < last_all;
      if (tid == 0 && cta == 0) TT(1, 0);
      // what does not depend on the challenge goes before the wait: the eq weight of the (single) pair of the coefficient pass
      const bool single = want_next && Hl <= nslots;
      fe wpre = Fq::zero();
      const u64 pslot = Hl <= 32 ? (u64)lane : slot;               // (see the coefficient pass)
      if (single && pslot < Hl) {
        const fe *el, *er; u32 sh; weights(round1 + 1, el, er, sh);
        const u64 g = (pslot << k) | (u64)cta, mask = ((u64)1 << sh) - 1;
        wpre = ldg_fe_ro(er + (el ? (g & mask) : g));
        if (el) wpre = Fq::mul(ldg_fe_ro(el + (g >> sh)), wpre);
      }
      if (round1 > first) mid_wait_released(st, round1 - 1);
      const fe r = ld_state(&st->r[round1 - 2]);
      if (tid == 0 && cta == 0) TT(1, 1);
      // bind to r: role t binds table t
      if (round1 == first) {
        const fe *S = role == 0 ? a.A : role == 1 ? a.B : a.C; fe *D = cur + role * cs;
        for (u64 j0 = slot; j0 < 2 * Pl; j0 += 4 * nslots) {    // strided reads of the natural-order table: four pairs in flight per thread
          fe lo[4], hi[4];
#pragma unroll
          for (int q = 0; q < 4; q++) {
            const u64 j = j0 + q * nslots, g = (j << k) | (u64)cta;
            if (j < 2 * Pl) { lo[q] = ldg_fe(S + g); hi[q] = ldg_fe(S + g + 2 * P); }
          }
#pragma unroll
          for (int q = 0; q < 4; q++) { const u64 j = j0 + q * nslots; if (j < 2 * Pl) D[j] = bind_pair(lo[q], hi[q], r); }
        }
      } else {
        const fe *S = cur + role * cs; fe *D = nxt + role * ns;
        for (u64 j = slot; j < 2 * Pl; j += nslots) D[j] = bind_pair(S[j], S[j + 2 * Pl], r);
        { fe *t = cur; cur = nxt; nxt = t; const u64 u = cs; cs = ns; ns = u; }
      }
      bar_sync_n(1, TP_ROLE);
      if (tid == 0 && cta == 0) TT(1, 2);
      const fe *cA = cur, *cB = cur + cs, *cC = cur + 2 * cs;
      if (round1 == first) {
        // the round's own sums, directly: role 0 -> t(0) over the low entries, role 1 -> t(1) over the high entries, role 2 -> t(inf)
        fe xs[3] = {Fq::zero(), Fq::zero(), Fq::zero()};
        if ((u64)trio * 32 < Pl) {
          const fe *el, *er; u32 sh; weights(round1, el, er, sh);
          const u64 mask = ((u64)1 << sh) - 1;
          Fq::acc d = Fq::acc_zero();
          for (u64 j = slot; j < Pl; j += nslots) {
            const u64 g = (j << k) | (u64)cta;
            fe w = ldg_fe_ro(er + (el ? (g & mask) : g));
            if (el) w = Fq::mul(ldg_fe_ro(el + (g >> sh)), w);
            if (role < 2) { const u64 o = role ? Pl : 0; Fq::mul_acc(d, w, Fq::sub(Fq::mul(cA[j + o], cB[j + o]), cC[j + o])); }
            else Fq::mul_acc(d, w, Fq::mul(Fq::sub(cA[j + Pl], cA[j]), Fq::sub(cB[j + Pl], cB[j])));
          }
          xs[0] = Fq::acc_reduce(d);
          warp_sum_fq_cols<3>(xs);
        }
        mid_publish_acc<3>(ts, xs, st->mid_acc0, 0, true);
        if (tid == 0) atomicAdd(&st->mid_arrive[round1], 1u);       // the finaliser starts on the round's own sums; its coefficients follow
        if (tid == 0 && cta == 0) TT(1, 3);
      }
      if (want_next) {
        // coefficient sums of round1 + 1 over the pairs (j, j + Pl/2) of this round's local table, weighted by the NEXT round's eq
        // (with <= 32 local pairs the three coefficient sums of a pair go to three different trios: one product chain per thread)
        fe xs[3] = {Fq::zero(), Fq::zero(), Fq::zero()};
        const bool split = Hl <= 32;
        const int cmask = split ? (trio < 3 ? 1 << trio : 0) : 7;
        const u64 jslot = split ? (u64)lane : slot;
        if (cmask && (split || (u64)trio * 32 < Hl)) {
          const fe *el, *er; u32 sh; weights(round1 + 1, el, er, sh);
          const u64 mask = ((u64)1 << sh) - 1;
          Fq::acc c0 = Fq::acc_zero(), c2 = Fq::acc_zero(), cd = Fq::acc_zero();
          for (u64 j = jslot; j < Hl; j += nslots) {
            fe w = wpre;
            if (!single) {
              const u64 g = (j << k) | (u64)cta;
              w = ldg_fe_ro(er + (el ? (g & mask) : g));
              if (el) w = Fq::mul(ldg_fe_ro(el + (g >> sh)), w);
            }
            fe x0, x2, y0, y2;
            if (role < 2) {
              const u64 o = role ? Hl : 0;
              x0 = cA[j + o]; x2 = cA[j + o + Pl]; y0 = cB[j + o]; y2 = cB[j + o + Pl];
              if (cmask & 1) Fq::mul_acc(c0, w, Fq::sub(Fq::mul(x0, y0), cC[j + o]));
              if (cmask & 2) Fq::mul_acc(c2, w, Fq::sub(Fq::mul(x2, y2), cC[j + o + Pl]));
            } else {
              x0 = Fq::sub(cA[j + Hl], cA[j]); x2 = Fq::sub(cA[j + Pl + Hl], cA[j + Pl]);
              y0 = Fq::sub(cB[j + Hl], cB[j]); y2 = Fq::sub(cB[j + Pl + Hl], cB[j + Pl]);
              if (cmask & 1) Fq::mul_acc(c0, w, Fq::mul(x0, y0));
              if (cmask & 2) Fq::mul_acc(c2, w, Fq::mul(x2, y2));
            }
            if (cmask & 4) Fq::mul_acc(cd, w, Fq::mul(Fq::sub(x2, x0), Fq::sub(y2, y0)));
          }
          if (cmask & 1) xs[0] = Fq::acc_reduce(c0);
          if (cmask & 2) xs[1] = Fq::acc_reduce(c2);
          if (cmask & 4) xs[2] = Fq::acc_reduce(cd);
          warp_sum_fq_cols<3>(xs);
        }
        if (tid == 0 && cta == 0) TT(1, 4);
        mid_publish_acc<3>(ts, xs, round1 == first ? st->mid_acc0 : mid_acc(st, round1), round1 == first ? 3 : 0, false);
      }
      if (round1 == last) {
        // last round on the cyclic layout: the bound table goes back to global memory in natural order (for role CTA 0, or for the tail kernel)
        const fe *S = cur + role * cs; fe *D = role == 0 ? a.oA : role == 1 ? a.oB : a.oC;
        for (u64 j = slot; j < 2 * Pl; j += nslots) stg_fe(D + ((j << k) | (u64)cta), S[j]);
        __threadfence();
        bar_sync_n(1, TP_ROLE);
      }
      if ((want_next || round1 == last) && tid == 0) atomicAdd(&st->mid_arrive[(round1 == first && want_next) ? MP_ARRIVE_FIRST_COEF : round1], 1u);
      if (tid == 0 && cta == 0) TT(1, 5);
    }
    if (a.finish) {
      // final claims: bind the three length-2 tables to the last challenge (role CTA 0; threads 0, 32, 64)
      mid_wait_released(st, l);
      if (lane == 0 && warp < 3) {
        const fe r = ld_state(&st->r[l - 1]);
        const fe lo = cur[warp * cs], hi = cur[warp * cs + 1];
        stg_fe(&st->claims[warp], Fq::add(lo, Fq::mul(Fq::sub(hi, lo), r)));
      }
    }
    return;
  }
  if (cta >= 0) return;
  // ---- finaliser group ----
  fin_cubic_init(fc, st, l, ft);
  for (int q = ft; q < (last_all - first + 1) * MP_ACC_WORDS; q += TP_FIN) mid_acc(st, first)[q] = 0;
  __threadfence();                                              // (ordered before the first release: the role CTAs add after it)
  tp_msg_init(ts, st, 3, ft);
  int cur = 0;
  fe rf = Fq::zero();                               // fin warp 0: the previous challenge, in registers
  for (int round1 = first; round1 <= last_all; round1++) {
    const bool want_next = round1 < last_all;
    const int Gr = round1 <= last ? G : 1;              // CTAs that arrive in this round
    if (ft == 0) st->prof[round1 - 1][0] = gtimer();
    if (ft == 0) TT(1, 6);
    if (round1 == first) {
      if (ft >= 32 && ft < 64) mid_gather_acc(st, round1, G, st->mid_acc0, 3, ts.gat);                    // the round's own sums
      bar_sync_n(2, TP_FIN);
      if (ft < 3) ts.x[ft] = ts.gat[ft];
      __syncwarp();
    }
    if (ft == 0) st->prof[round1 - 1][1] = gtimer();
    fe canon = Fq::zero();
    if (ft < 32) canon = fin_cubic_msg(ts, fc, st, round1, round1 == first, ts.coef[cur], rf, ft);
    fe *coef_next = ts.coef[cur ^ 1];
    if (ft == 0) st->prof[round1 - 1][2] = gtimer();
    if (ft == 0) TT(1, 7);
    rf = tp_squeeze(ts, canon, 3, ft, [&] { fin_cubic_next(ts, fc, round1, l, ft); },
                    [&] { if (want_next) mid_gather_acc(st, round1 == first ? MP_ARRIVE_FIRST_COEF : round1, Gr, round1 == first ? st->mid_acc0 + 3 * 16 : mid_acc(st, round1), 9, coef_next); });
    if (ft == 0) { stg_fe(&st->r[round1 - 1], rf); __threadfence(); st_volatile_u32(&st->mid_released, (u32)round1); st->prof[round1 - 1][3] = gtimer(); }
    if (ft == 0) TT(1, 8);
    cur ^= 1;
  }
  // hand-off to the next kernel: transcript, and the eq prefix of the round after the last
  tp_msg_store(ts, st, 3, ft);
  if (ft < 32) fin_cubic_store(fc, st, last_all, l, rf, ft);
}

__global__ void __launch_bounds__(TP_THREADS, 1) k_quad_mid_pipe(MidQuad a) {
  extern __shared__ __align__(32) unsigned char mp_dyn[];
  __shared__ TailSmem ts;
  __shared__ FinQuad fq;
  ScState *st = a.st;
  const int rounds = a.rounds, first = a.round_first, last = a.round_last, G = (int)gridDim.x - 1;
  const int last_all = a.finish ? rounds : last;      // the kernel's last round (see MidCubic)
  int k = a.k, cta = (int)blockIdx.x - 1;             // block 0: finaliser only
  const int tid = threadIdx.x, ft = tid - TP_ROLE, lane = tid & 31;
  const bool is_role = tid < TP_ROLE;
  const int warp = tid >> 5;
  if (is_role) {
    if (cta < 0) return;
    const u64 n0 = ((u64)2 << (rounds - first)) >> k;          // local length of the first bound table
    fe *X = (fe *)mp_dyn, *Y = X + 2 * n0;
    fe *cur = X, *nxt = Y; u64 cs = n0, ns = n0 / 2;
    const int grp = warp & 1, pair = warp >> 1;                 // even warps: the low halves, odd warps: the differences
    for (int round1 = first; round1 <= last_all; round1++) {
      if (round1 == last + 1) {
        // the rounds on <= SC_TAIL_LEN entries: role CTA 0 alone, the whole tables (natural order) in its shared memory
        if (cta != 0) return;
        if (tid == 0) { const u32 want = (u32)G; const u32 *ctr = &st->mid_arrive[last]; spin_until([=] { return ld_volatile_u32(ctr) >= want; }, &st->err); __threadfence(); }
        bar_sync_n(1, TP_ROLE);
        k = 0;
        const u64 nt = (u64)2 << (rounds - last);
        cur = X; nxt = X + 2 * nt; cs = nt; ns = nt / 2;
        for (u64 q = (u64)tid; q < 2 * nt; q += TP_ROLE) cur[(q & 1) * cs + (q >> 1)] = ldg_fe(((q & 1) ? a.oB : a.oA) + (q >> 1));
        bar_sync_n(1, TP_ROLE);
      }
      const u64 P = (u64)1 << (rounds - round1), Pl = P >> k, Hl = Pl >> 1;
      const bool want_next = round1 < last_all;
      if (round1 > first) mid_wait_released(st, round1 - 1);
      const fe r = ld_state(&st->r[round1 - 2]);
      if (round1 == first) {
        for (u64 q0 = (u64)tid; q0 < 4 * Pl; q0 += 4 * TP_ROLE) {   // strided reads of the natural-order tables: four pairs in flight per thread
          fe lo[4], hi[4];
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const u64 q = q0 + (u64)u * TP_ROLE, j = q >> 1, g = (j << k) | (u64)cta;
            const fe *S = (q & 1) ? a.B : a.A;
            if (q < 4 * Pl) { lo[u] = ldg_fe(S + g); hi[u] = ldg_fe(S + g + 2 * P); }
          }
#pragma unroll
          for (int u = 0; u < 4; u++) {
            const u64 q = q0 + (u64)u * TP_ROLE, j = q >> 1;
            if (q < 4 * Pl) cur[(q & 1) * cs + j] = bind_pair(lo[u], hi[u], r);
          }
        }
      } else {
        for (u64 q = (u64)tid; q < 4 * Pl; q += TP_ROLE) {
          const u64 j = q >> 1;
          const fe *S = cur + (q & 1) * cs; fe *D = nxt + (q & 1) * ns;
          D[j] = bind_pair(S[j], S[j + 2 * Pl], r);
        }
        { fe *t = cur; cur = nxt; nxt = t; const u64 u = cs; cs = ns; ns = u; }
      }
      bar_sync_n(1, TP_ROLE);
      const fe *cA = cur, *cB = cur + cs;
      if (round1 == first) {
        // the round's own sums, directly: even warps e0 = sum a0 b0, odd warps t(inf) = sum (a1 - a0)(b1 - b0)
        fe xs[3] = {Fq::zero(), Fq::zero(), Fq::zero()};
        if ((u64)pair * 32 < Pl) {
          Fq::acc d = Fq::acc_zero();
          for (u64 j = (u64)pair * 32 + lane; j < Pl; j += (TP_ROLE / 64) * 32) {
            if (grp == 0) Fq::mul_acc(d, cA[j], cB[j]);
            else Fq::mul_acc(d, Fq::sub(cA[j + Pl], cA[j]), Fq::sub(cB[j + Pl], cB[j]));
          }
          xs[0] = Fq::acc_reduce(d);
          warp_sum_fq_cols<3>(xs);
        }
        mid_publish_acc<2>(ts, xs, st->mid_acc0, 0, true);
        if (tid == 0) atomicAdd(&st->mid_arrive[round1], 1u);       // the finaliser starts on the round's own sums; its coefficients follow
      }
      if (want_next) {
        // (with <= 32 local pairs the three coefficient sums of a pair go to three different warp pairs: one product per thread)
        fe xs[3] = {Fq::zero(), Fq::zero(), Fq::zero()};
        const bool split = Hl <= 32;
        const int cmask = split ? (pair < 3 ? 1 << pair : 0) : 7;
        if (cmask && (split || (u64)pair * 32 < Hl)) {
          Fq::acc c0 = Fq::acc_zero(), c2 = Fq::acc_zero(), cd = Fq::acc_zero();
          for (u64 j = split ? (u64)lane : (u64)pair * 32 + lane; j < Hl; j += (TP_ROLE / 64) * 32) {
            fe x0, x2, y0, y2;
            if (grp == 0) { x0 = cA[j]; x2 = cA[j + Pl]; y0 = cB[j]; y2 = cB[j + Pl]; }
            else {
              x0 = Fq::sub(cA[j + Hl], cA[j]); x2 = Fq::sub(cA[j + Pl + Hl], cA[j + Pl]);
              y0 = Fq::sub(cB[j + Hl], cB[j]); y2 = Fq::sub(cB[j + Pl + Hl], cB[j + Pl]);
            }
            if (cmask & 1) Fq::mul_acc(c0, x0, y0);
            if (cmask & 2) Fq::mul_acc(c2, x2, y2);
            if (cmask & 4) Fq::mul_acc(cd, Fq::sub(x2, x0), Fq::sub(y2, y0));
          }
          if (cmask & 1) xs[0] = Fq::acc_reduce(c0);
          if (cmask & 2) xs[1] = Fq::acc_reduce(c2);
          if (cmask & 4) xs[2] = Fq::acc_reduce(cd);
          warp_sum_fq_cols<3>(xs);
        }
        mid_publish_acc<2>(ts, xs, round1 == first ? st->mid_acc0 : mid_acc(st, round1), round1 == first ? 3 : 0, false);
      }
      if (round1 == last) {
        // last round on the cyclic layout: the bound tables go back to global memory in natural order
        for (u64 q = (u64)tid; q < 4 * Pl; q += TP_ROLE) {
          const u64 j = q >> 1;
          stg_fe(((q & 1) ? a.oB : a.oA) + ((j << k) | (u64)cta), cur[(q & 1) * cs + j]);
        }
        __threadfence();
        bar_sync_n(1, TP_ROLE);
      }
      if ((want_next || round1 == last) && tid == 0) atomicAdd(&st->mid_arrive[(round1 == first && want_next) ? MP_ARRIVE_FIRST_COEF : round1], 1u);
    }
    if (a.finish) {
      // final claims: bind the two length-2 tables to the last challenge (role CTA 0; threads 0, 32)
      mid_wait_released(st, rounds);
      if (lane == 0 && warp < 2) {
        const fe r = ld_state(&st->r[rounds - 1]);
        const fe lo = cur[warp * cs], hi = cur[warp * cs + 1];
        stg_fe(&st->claims[warp], Fq::add(lo, Fq::mul(Fq::sub(hi, lo), r)));
      }
    }
    return;
  }
  if (cta >= 0) return;
  // ---- finaliser group ----
  fin_quad_init(fq, st, ft);
  for (int q = ft; q < (last_all - first + 1) * MP_ACC_WORDS; q += TP_FIN) mid_acc(st, first)[q] = 0;
  __threadfence();                                              // (ordered before the first release: the role CTAs add after it)
  tp_msg_init(ts, st, 2, ft);
  int cur = 0;
  fe rf = Fq::zero();                               // fin warp 0: the previous challenge, in registers
  for (int round1 = first; round1 <= last_all; round1++) {
    const bool want_next = round1 < last_all;
    const int Gr = round1 <= last ? G : 1;              // CTAs that arrive in this round
    if (round1 == first) {
      if (ft >= 32 && ft < 64) mid_gather_acc(st, round1, G, st->mid_acc0, 3, ts.gat);                    // the round's own sums (0, 1)
      bar_sync_n(2, TP_FIN);
      if (ft < 2) ts.x[ft] = ts.gat[ft];
      __syncwarp();
    }
    fe canon = Fq::zero();
    if (ft < 32) canon = fin_quad_msg(ts, fq, st, round1, round1 == first, ts.coef[cur], rf, ft);
    fe *coef_next = ts.coef[cur ^ 1];
    rf = tp_squeeze(ts, canon, 2, ft, TpNoSide(), [&] { if (want_next) mid_gather_acc(st, round1 == first ? MP_ARRIVE_FIRST_COEF : round1, Gr, round1 == first ? st->mid_acc0 + 3 * 16 : mid_acc(st, round1), 6, coef_next); });
    if (ft == 0) { stg_fe(&st->r[round1 - 1], rf); __threadfence(); st_volatile_u32(&st->mid_released, (u32)round1); st_volatile_u32(&st->r_ready, (u32)round1); }
    cur ^= 1;
  }
  tp_msg_store(ts, st, 2, ft);
  if (ft < 32) fin_quad_store(fq, st, rf, ft);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// SP2_NO_DERIVE=1: three direct sums in every round (measurement switch)
static bool derive_enabled() { static int v = -1; if (v < 0) { const char *e = getenv("SP2_NO_DERIVE"); v = (e && e[0] == '1') ? 0 : 1; } return v == 1; }
uint32_t sumcheck_cubic_persist_rounds(sp2_ctx *ctx, uint32_t l);
int sc_state_upload(sp2_ctx *ctx, ScState **d_st, const uint64_t *claim, const uint64_t *taus, uint32_t l,
                    const sp2_transcript_state *ts) {
  void *d, *h;
  SP2_TRY(scratch(ctx, 14, sizeof(ScState), &d));
  SP2_TRY(pinned(ctx, sizeof(ScState), &h));
  ScState *hs = (ScState *)h;
  const size_t head = offsetof(ScState, r);
  memset(hs, 0, head);
  hs->ts.round = ts->round; hs->ts.pending_len = 0; memcpy(hs->ts.state, ts->state, 64);
  memcpy(&hs->claim, claim, sizeof(fe));
  hs->l = l;
  { static int kflag = -1; if (kflag < 0) { const char *e = getenv("SP2_KECCAK_THREAD"); kflag = (e && e[0] == '1') ? 0 : 1; } hs->flags = (u32)kflag; }
  if (taus) memcpy(hs->taus, taus, (size_t)l * sizeof(fe));
  // standalone cubic prover: the host has the taus right here — the streaming rounds derive t(1) from the claim (sumcheck.cuh) with
  // inverses from one batch inversion, unless a tau of those rounds is zero (then, as everywhere else, three direct sums)
  if (taus && derive_enabled()) {
    uint32_t n = std::min<uint32_t>(SC_DERIVE_MAX, sumcheck_cubic_persist_rounds(ctx, l));
    uint64_t pre[SC_DERIVE_MAX][4], acc[4], inv[4];
    for (uint32_t i = 0; i < n; i++) { const uint64_t *t = taus + 4 * i; if (!(t[0] | t[1] | t[2] | t[3])) n = 0; }
    if (n) {
      memcpy(acc, taus, 32);
      for (uint32_t i = 1; i < n; i++) { memcpy(pre[i], acc, 32); sp2h::mont_mul(acc, taus + 4 * i, sp2h::FQ_MOD, sp2h::FQ_INV, acc); }
      sp2h::fq_inv(acc, inv);
      for (uint32_t i = n; i-- > 1;) {
        uint64_t ti[4]; sp2h::mont_mul(inv, pre[i], sp2h::FQ_MOD, sp2h::FQ_INV, ti); memcpy(hs->mail.tinv[i], ti, 32);
        sp2h::mont_mul(inv, taus + 4 * i, sp2h::FQ_MOD, sp2h::FQ_INV, inv);
      }
      memcpy(hs->mail.tinv[0], inv, 32);
      hs->mail.n = n; hs->mail.flag = 1; hs->derive_rounds = n;
      memcpy(&hs->tclaim, claim, sizeof(fe));
    }
  }
  SP2_CUDA_OK(cudaMemcpyAsync(d, hs, head, cudaMemcpyHostToDevice, ctx->stream));
  *d_st = (ScState *)d;
  return SP2_OK;
}

int sc_state_download(sp2_ctx *ctx, ScState *d_st, sp2_transcript_state *ts, uint64_t *polys, int ncoef, uint64_t *r,
                      uint64_t *claims, int nclaims, uint32_t l) {
  void *h;
  SP2_TRY(pinned(ctx, sizeof(ScState), &h));
  ScState *hs = (ScState *)h;
  const size_t upto = offsetof(ScState, partial);
  SP2_CUDA_OK(cudaMemcpyAsync(hs, d_st, upto, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (hs->err) return set_error(ctx, SP2_ERR_INTERNAL, "sum-check: a device-side wait (grid barrier) did not complete within 2 s");
  ts->round = (uint16_t)hs->ts.round; memcpy(ts->state, hs->ts.state, 64);
  for (uint32_t i = 0; i < l; i++) memcpy(polys + (size_t)i * ncoef * 4, &hs->polys[4 * i], (size_t)ncoef * sizeof(fe));
  memcpy(r, hs->r, (size_t)l * sizeof(fe));
  memcpy(claims, hs->claims, (size_t)nclaims * sizeof(fe));
  return SP2_OK;
}

static bool use_persistent() { static int v = -1; if (v < 0) { const char *e = getenv("SP2_NO_PERSIST"); v = (e && e[0] == '1') ? 0 : 1; } return v == 1; }
// SP2_TAIL_PIPE=0: the non-pipelined single-CTA tails (measurement switch)
static bool use_tail_pipe() { static int v = -1; if (v < 0) { const char *e = getenv("SP2_TAIL_PIPE"); v = (e && e[0] == '0') ? 0 : 1; } return v == 1; }
// SP2_MID_PIPE=0: every multi-CTA round stays in the persistent kernels (measurement switch)
static bool use_mid_pipe() { static int v = -1; if (v < 0) { const char *e = getenv("SP2_MID_PIPE"); v = (e && e[0] == '0') ? 0 : 1; } return v == 1 && use_tail_pipe(); }
// SP2_MID_FINISH=0: the rounds on <= SC_TAIL_LEN entries go to the single-CTA tail kernels in a second launch (measurement switch)
static bool use_mid_finish() { static int v = -1; if (v < 0) { const char *e = getenv("SP2_MID_FINISH"); v = (e && e[0] == '0') ? 0 : 1; } return v == 1; }
// plan of the pipelined multi-CTA rounds [*first, round_end - 1] on 2^*k CTAs (k_*_mid_pipe); false when fewer than two such rounds exist
static bool mid_plan(sp2_ctx *ctx, uint32_t l, uint32_t min_first, uint32_t round_end, int ntab, uint32_t *first, uint32_t *k, size_t *smem) {
  if (!use_mid_pipe()) return false;
  uint32_t rm = min_first;
  while (rm <= l && (4ull << (l - rm)) > (1ull << (ntab == 2 ? MP_LOG_LEN_IN_QUAD : MP_LOG_LEN_IN))) rm++;
  if (rm + 1 >= round_end) return false;
  int kk = (int)(l - rm) + 1 - MP_LOG_LOCAL;                    // log2(length of the first bound table) - log2(local length)
  if (kk > MP_LOG_CTAS) kk = MP_LOG_CTAS;
  while (kk > 0 && (1 << kk) + 1 > ctx->num_sms) kk--;
  while (kk > 0 && (((u64)1 << (l - (round_end - 2))) >> kk) < 2) kk--;   // the coefficient rounds need two local pairs
  if (kk < 1) return false;
  const u64 n0 = ((u64)2 << (l - rm)) >> kk;
  u64 cap = n0 + n0 / 2;
  if (use_mid_finish()) { const u64 nt = (u64)2 << (l - (round_end - 1)); cap = std::max<u64>(cap, nt + nt / 2); }   // role CTA 0 holds the whole table of the last rounds
  *smem = (size_t)ntab * cap * sizeof(fe);
  if (*smem > 200 * 1024) return false;
  *first = rm; *k = (uint32_t)kk;
  return true;
}
// launch the pipelined multi-CTA rounds [first, round_end) of a cubic / quadratic sum-check (tables src -> natural-order tables dst)
static int launch_mid_cubic(sp2_ctx *ctx, ScState *st, uint32_t l, fe *const *src, fe *const *dst, uint32_t first, uint32_t round_end, uint32_t k, size_t smem,
                            const fe *eq_left, const fe *eq_right) {
  static bool attr = false;
  if (!attr) { SP2_CUDA_OK(cudaFuncSetAttribute((const void *)k_cubic_mid_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
  MidCubic ma{st, src[0], src[1], src[2], dst[0], dst[1], dst[2], (int)l, (int)first, (int)round_end - 1, (int)k, eq_left, eq_right, use_mid_finish() ? 1 : 0};
  void *args[] = {&ma};
  SP2_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_cubic_mid_pipe, dim3((1u << k) + 1), dim3(TP_THREADS), args, smem, ctx->stream));
  ctx->launches++;
  return SP2_OK;
}
static int launch_mid_quad(sp2_ctx *ctx, ScState *st, uint32_t rounds, fe *const *src, fe *const *dst, uint32_t first, uint32_t round_end, uint32_t k, size_t smem) {
  static bool attr = false;
  if (!attr) { SP2_CUDA_OK(cudaFuncSetAttribute((const void *)k_quad_mid_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
  MidQuad ma{st, src[0], src[1], dst[0], dst[1], (int)rounds, (int)first, (int)round_end - 1, (int)k, use_mid_finish() ? 1 : 0};
  void *args[] = {&ma};
  SP2_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_quad_mid_pipe, dim3((1u << k) + 1), dim3(TP_THREADS), args, smem, ctx->stream));
  ctx->launches++;
  return SP2_OK;
}
static DevComm comm_none() { DevComm d; memset(&d, 0, sizeof(d)); d.n = 1; return d; }

// all-gather the shards (len_local entries per table) into every rank's gather area and return the local copy
static int shard_gather(sp2_ctx *ctx, const DevComm &dc, fe *const *src, int ntab, u64 len_local, fe **out) {
  if ((len_local << dc.k) > SC_ROLE_LEN) return set_error(ctx, SP2_ERR_INTERNAL, "shard gather: table too large");
  unsigned nb = (unsigned)((len_local + 255) / 256); if (nb > (unsigned)ctx->num_sms * 4) nb = ctx->num_sms * 4; if (nb == 0) nb = 1;
  k_shard_gather<<<nb, 256, 0, ctx->stream>>>(dc, src[0], src[1], ntab > 2 ? src[2] : src[1], len_local, ntab);
  SP2_LAUNCH_CHECK();
  k_shard_barrier<<<1, 32, 0, ctx->stream>>>(dc, SC_MAX_ROUNDS + 1);
  SP2_LAUNCH_CHECK();
  for (int t = 0; t < ntab; t++) out[t] = dc.peer[dc.rank]->gather[t];
  return SP2_OK;
}

// The prover enqueues the cubic sum-check while a kernel in front of it waits for the HOST (prover.cu: k_gate_taus).  Nothing on the
// enqueue path may then need the device to go idle: (a) with CUDA's lazy module loading the first launch of a kernel can require a
// context synchronisation — the kernels are loaded beforehand; (b) scratch growth is an allocation — the slots are reserved beforehand.
template <class K> static void preload_one(K k) { cudaFuncAttributes a; cudaFuncGetAttributes(&a, (const void *)k); }
void sumcheck_cubic_preload() {
  static unsigned done_mask = 0;                     // per device (the loading is per CUDA context)
  int dev = 0; cudaGetDevice(&dev);
  if (done_mask >> dev & 1u) return;
  preload_one(k_cubic_init); preload_one(k_cubic_persist); preload_one(k_cubic_mid_pipe); preload_one(k_cubic_tail_pipe); preload_one(k_cubic_tail);
  preload_one(k_cubic_round<true, 0>); preload_one(k_cubic_round<true, 1>); preload_one(k_cubic_round<false, 0>); preload_one(k_cubic_round<false, 1>);
  preload_one(k_cubic_round_roles<true>); preload_one(k_cubic_round_roles<false>);
  preload_one(k_shard_gather); preload_one(k_shard_barrier);
  cudaFuncSetAttribute((const void *)k_cubic_mid_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  done_mask |= 1u << dev;
}
int sumcheck_cubic_reserve(sp2_ctx *ctx, uint32_t l) {
  const int first_half = (int)l / 2, second_half = (int)l - first_half;
  const size_t nleft = (size_t)1 << (first_half > 0 ? first_half : 1), nright = (size_t)2 << second_half;
  void *p;
  SP2_TRY(scratch(ctx, 13, (nleft + nright) * sizeof(fe), &p));
  SP2_TRY(scratch(ctx, 7, 3 * (SC_ROLE_LEN / 2) * sizeof(fe), &p));
  return SP2_OK;
}

uint32_t sumcheck_cubic_persist_rounds(sp2_ctx *ctx, uint32_t l) {
  if (!use_persistent()) return 0;
  uint32_t round_end = 1;
  while (round_end <= l && ((round_end > 1 ? 4ull : 2ull) << (l - round_end)) > SC_TAIL_LEN) round_end++;
  uint32_t mid_first = 0, mid_k = 0; size_t mid_smem = 0;
  const bool mid = mid_plan(ctx, l, 2, round_end, 3, &mid_first, &mid_k, &mid_smem);
  const uint32_t persist_end = mid ? mid_first : round_end;
  return persist_end > 1 ? persist_end - 1 : 0;
}

// enqueue the whole cubic sum-check on ctx->stream (state already on the device, taus in st->taus)
int sumcheck_cubic_enqueue(sp2_ctx *ctx, ScState *st, uint32_t l, fe *A, fe *B, fe *C, const DevComm *dcp, const ScTinvMail *mail, uint32_t mail_epoch) {
  const int first_half = (int)l / 2, second_half = (int)l - first_half;
  void *eqs;
  const size_t nleft = (size_t)1 << (first_half > 0 ? first_half : 1), nright = (size_t)2 << second_half;
  SP2_TRY(scratch(ctx, 13, (nleft + nright) * sizeof(fe), &eqs));
  fe *eq_left = (fe *)eqs, *eq_right = eq_left + nleft;
  k_cubic_init<<<2, 1024, 0, ctx->stream>>>(st, (int)l, eq_left, eq_right);
  SP2_LAUNCH_CHECK();
  const unsigned target = (unsigned)ctx->num_sms * 2;
  // scratch copies for the small (role-split, ping-pong) rounds: tables of <= SC_ROLE_LEN entries going in
  void *pp; SP2_TRY(scratch(ctx, 7, 3 * (SC_ROLE_LEN / 2) * sizeof(fe), &pp));
  fe *s2[3] = {(fe *)pp, (fe *)pp + SC_ROLE_LEN / 2, (fe *)pp + SC_ROLE_LEN};
  fe *src[3] = {A, B, C}, *dst[3] = {s2[0], s2[1], s2[2]};
  DevComm dc = dcp ? *dcp : comm_none();
  bool sharded = dc.n > 1;
  if (sharded) {
    if ((int)l < dc.k) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "sharded sum-check: fewer entries than ranks");
    k_shard_barrier<<<1, 32, 0, ctx->stream>>>(dc, 0);          // every rank has finished the previous sharded call
    SP2_LAUNCH_CHECK();
  }
  uint32_t round_start = 1;
  if (!sharded && use_persistent()) {
    // all multi-CTA rounds in one cooperative launch (one CTA per SM), then the single-CTA tail
    uint32_t round_end = 1;
    while (round_end <= l && ((round_end > 1 ? 4ull : 2ull) << (l - round_end)) > SC_TAIL_LEN) round_end++;
    uint32_t mid_first = 0, mid_k = 0; size_t mid_smem = 0;
    const bool mid = mid_plan(ctx, l, 2, round_end, 3, &mid_first, &mid_k, &mid_smem);
    const uint32_t persist_end = mid ? mid_first : round_end;
    if (persist_end > 1) {
      if (!mail) { mail = &st->mail; mail_epoch = 1; }               // standalone provers: inverses uploaded with the state (or none: derive_rounds = 0)
      PersistCubic pa{st, A, B, C, dst[0], dst[1], dst[2], (int)l, 1, (int)persist_end, eq_left, eq_right, mail, mail_epoch};
      void *args[] = {&pa};
      SP2_CUDA_OK(cudaEventRecord(ctx->ev_k0, ctx->stream));
      SP2_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_cubic_persist, dim3(ctx->num_sms * SC_PERSIST_MINB), dim3(SC_THREADS), args, 0, ctx->stream));
      SP2_CUDA_OK(cudaEventRecord(ctx->ev_k1, ctx->stream));
      ctx->ev_k_valid = true;
      ctx->launches++;
      for (uint32_t rr = 2; rr < persist_end; rr++)            // fused role rounds ping-pong src <-> dst
        if ((4ull << (l - rr)) <= SC_ROLE_LEN) for (int k = 0; k < 3; k++) std::swap(src[k], dst[k]);
      round_start = persist_end;
    }
    if (mid) {
      // rounds [mid_first, round_end): pipelined on 2^mid_k CTAs; the table bound by the last of them lands in dst (natural order)
      SP2_TRY(launch_mid_cubic(ctx, st, l, src, dst, mid_first, round_end, mid_k, mid_smem, eq_left, eq_right));
      for (int k = 0; k < 3; k++) std::swap(src[k], dst[k]);
      round_start = use_mid_finish() ? l + 1 : round_end;      // (finish mode: the kernel ran every remaining round and left the claims)
    }
  }
  for (uint32_t round1 = round_start; round1 <= l; round1++) {
    const bool fused = round1 > 1;
    const u64 Pg = (u64)1 << (l - round1);                      // pairs evaluated this round (global)
    const u64 len_in = fused ? 4 * Pg : 2 * Pg;                 // global table length going into this launch
    if (sharded && len_in <= SC_ROLE_LEN) {                      // hand-off: gather the shards, finish redundantly on every rank
      fe *g[3];
      SP2_TRY(shard_gather(ctx, dc, src, 3, len_in >> dc.k, g));
      for (int k = 0; k < 3; k++) src[k] = g[k];
      sharded = false; dc = comm_none();
    }
    const u64 P = sharded ? Pg >> dc.k : Pg;                    // local pairs
    if (!sharded && fused && len_in > SC_TAIL_LEN && len_in <= (1ull << MP_LOG_LEN_IN)) {
      // (the sharded provers arrive here after the all-gather: the remaining multi-CTA rounds go to the pipelined kernel, redundantly on every rank)
      uint32_t round_end = round1, mf = 0, mk = 0; size_t msm = 0;
      while (round_end <= l && (4ull << (l - round_end)) > SC_TAIL_LEN) round_end++;
      if (mid_plan(ctx, l, round1, round_end, 3, &mf, &mk, &msm) && mf == round1) {
        SP2_TRY(launch_mid_cubic(ctx, st, l, src, dst, mf, round_end, mk, msm, eq_left, eq_right));
        for (int k = 0; k < 3; k++) std::swap(src[k], dst[k]);
        if (use_mid_finish()) break;
        round1 = round_end - 1;
        continue;
      }
    }
    if (!sharded && len_in <= SC_TAIL_LEN) {
      if (use_tail_pipe()) k_cubic_tail_pipe<<<1, TP_THREADS, 0, ctx->stream>>>(st, src[0], src[1], src[2], dst[0], dst[1], dst[2], (int)round1, (int)l, eq_left, eq_right);
      else k_cubic_tail<<<1, SC_TAIL_THREADS + 32, 0, ctx->stream>>>(st, src[0], src[1], src[2], dst[0], dst[1], dst[2], (int)round1, (int)l, eq_left, eq_right);
      SP2_LAUNCH_CHECK();
      break;
    }
    const bool in_first = (int)round1 < first_half;
    const fe *el = nullptr, *er; u32 out_len = 1, sh = 0;
    if (in_first) {
      const int kl = first_half - (int)round1;
      el = eq_left + (((size_t)1 << kl) - 1); out_len = 1u << kl;
      er = eq_right + (((size_t)1 << second_half) - 1); sh = (u32)second_half;
    } else {
      er = eq_right + (((size_t)1 << (l - round1)) - 1);
    }
    if (!sharded && len_in <= SC_ROLE_LEN) {                      // small multi-CTA round: three items per pair, src -> dst
      const u64 per_cta = (SC_ROLE_THREADS / 96) * 32;             // pairs per CTA and pass
      u64 nb = (P + per_cta - 1) / per_cta; if (nb > (u64)ctx->num_sms) nb = ctx->num_sms;
      if (fused) {
        k_cubic_round_roles<true><<<(unsigned)nb, SC_ROLE_THREADS, 0, ctx->stream>>>(st, src[0], src[1], src[2], dst[0], dst[1], dst[2], P, (int)round1, (int)l, el, er, sh);
        for (int k = 0; k < 3; k++) std::swap(src[k], dst[k]);
      } else {
        k_cubic_round_roles<false><<<(unsigned)nb, SC_ROLE_THREADS, 0, ctx->stream>>>(st, src[0], src[1], src[2], dst[0], dst[1], dst[2], P, (int)round1, (int)l, el, er, sh);
      }
      SP2_LAUNCH_CHECK();
      continue;
    }
    // large round: in place on (A, B, C) — local shards when sharded
    const u64 in_len_local = ((u64)1 << sh) >> dc.k;             // local width of x_in (two-level mode)
    if (in_first && in_len_local >= SC_THREADS && in_len_local / SC_THREADS <= target) {
      dim3 grid((unsigned)(in_len_local / SC_THREADS), 1);
      unsigned gy = target / grid.x; if (gy < 1) gy = 1; if (gy > out_len) gy = out_len;
      grid.y = gy;
      const u32 sh_local = sh - (u32)dc.k;
      if (fused) k_cubic_round<true, 0><<<grid, SC_THREADS, 0, ctx->stream>>>(st, A, B, C, P, (int)round1, (int)l, el, er, out_len, sh_local, dc);
      else k_cubic_round<false, 0><<<grid, SC_THREADS, 0, ctx->stream>>>(st, A, B, C, P, (int)round1, (int)l, el, er, out_len, sh_local, dc);
    } else {
      u64 nb = (P + SC_THREADS - 1) / SC_THREADS; if (nb > target) nb = target; if (nb == 0) nb = 1;
      if (fused) k_cubic_round<true, 1><<<(unsigned)nb, SC_THREADS, 0, ctx->stream>>>(st, A, B, C, P, (int)round1, (int)l, el, er, out_len, sh, dc);
      else k_cubic_round<false, 1><<<(unsigned)nb, SC_THREADS, 0, ctx->stream>>>(st, A, B, C, P, (int)round1, (int)l, el, er, out_len, sh, dc);
    }
    SP2_LAUNCH_CHECK();
  }
  return SP2_OK;
}

int sumcheck_quad_enqueue(sp2_ctx *ctx, ScState *st, uint32_t rounds, fe *A, fe *B, uint64_t nvalid, cudaEvent_t after_first,
                          const DevComm *dcp) {
  bool recorded = false;
  const unsigned target = (unsigned)ctx->num_sms * 2;
  void *pp; SP2_TRY(scratch(ctx, 6, 2 * (SC_ROLE_LEN / 2) * sizeof(fe), &pp));
  fe *src[2] = {A, B}, *dst[2] = {(fe *)pp, (fe *)pp + SC_ROLE_LEN / 2};
  DevComm dc = dcp ? *dcp : comm_none();
  bool sharded = dc.n > 1;
  if (sharded) {
    if (nvalid != ~0ull) return set_error(ctx, SP2_ERR_UNSUPPORTED, "sharded quad sum-check needs dense tables");
    if ((int)rounds < dc.k) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "sharded sum-check: fewer entries than ranks");
    k_shard_barrier<<<1, 32, 0, ctx->stream>>>(dc, 0);
    SP2_LAUNCH_CHECK();
  }
  uint32_t round_start = 1;
  if (!sharded && use_persistent()) {
    uint32_t round_end = 1;
    while (round_end <= rounds && ((round_end > 1 ? 4ull : 2ull) << (rounds - round_end)) > SC_TAIL_LEN) round_end++;
    uint32_t mid_first = 0, mid_k = 0; size_t mid_smem = 0;
    const bool mid = mid_plan(ctx, rounds, 3, round_end, 2, &mid_first, &mid_k, &mid_smem);   // from round 3 on every entry is materialised
    const uint32_t persist_end = mid ? mid_first : round_end;
    if (persist_end > 1) {
      PersistQuad pa{st, A, B, dst[0], dst[1], (int)rounds, 1, (int)persist_end, nvalid};
      void *args[] = {&pa};
      SP2_CUDA_OK(cudaLaunchCooperativeKernel((const void *)k_quad_persist, dim3(ctx->num_sms * SC_PERSIST_MINB), dim3(SC_THREADS), args, 0, ctx->stream));
      ctx->launches++;
      if (after_first && !recorded) { SP2_CUDA_OK(cudaEventRecord(after_first, ctx->stream)); recorded = true; }
      for (uint32_t rr = 2; rr < persist_end; rr++)
        if ((4ull << (rounds - rr)) <= SC_ROLE_LEN) { std::swap(src[0], dst[0]); std::swap(src[1], dst[1]); }
      round_start = persist_end;
    }
    if (mid) {
      SP2_TRY(launch_mid_quad(ctx, st, rounds, src, dst, mid_first, round_end, mid_k, mid_smem));
      std::swap(src[0], dst[0]); std::swap(src[1], dst[1]);
      round_start = use_mid_finish() ? rounds + 1 : round_end;
    }
  }
  for (uint32_t round1 = round_start; round1 <= rounds; round1++) {
    const u64 Pg = (u64)1 << (rounds - round1);
    const u64 len_in = round1 > 1 ? 4 * Pg : 2 * Pg;
    if (sharded && len_in <= SC_ROLE_LEN) {
      fe *g[3];
      SP2_TRY(shard_gather(ctx, dc, src, 2, len_in >> dc.k, g));
      src[0] = g[0]; src[1] = g[1];
      sharded = false; dc = comm_none();
    }
    const u64 P = sharded ? Pg >> dc.k : Pg;
    if (!sharded && round1 >= 3 && len_in > SC_TAIL_LEN && len_in <= (1ull << MP_LOG_LEN_IN_QUAD)) {
      // (after the all-gather of a sharded prover: the remaining multi-CTA rounds go to the pipelined kernel)
      uint32_t round_end = round1, mf = 0, mk = 0; size_t msm = 0;
      while (round_end <= rounds && (4ull << (rounds - round_end)) > SC_TAIL_LEN) round_end++;
      if (mid_plan(ctx, rounds, round1, round_end, 2, &mf, &mk, &msm) && mf == round1) {
        SP2_TRY(launch_mid_quad(ctx, st, rounds, src, dst, mf, round_end, mk, msm));
        std::swap(src[0], dst[0]); std::swap(src[1], dst[1]);
        if (after_first && !recorded) { SP2_CUDA_OK(cudaEventRecord(after_first, ctx->stream)); recorded = true; }
        if (use_mid_finish()) break;
        round1 = round_end - 1;
        continue;
      }
    }
    if (!sharded && len_in <= SC_TAIL_LEN) {
      if (use_tail_pipe()) k_quad_tail_pipe<<<1, TP_THREADS, 0, ctx->stream>>>(st, src[0], src[1], dst[0], dst[1], (int)round1, (int)rounds, nvalid);
      else k_quad_tail<<<1, SC_TAIL_THREADS + 32, 0, ctx->stream>>>(st, src[0], src[1], dst[0], dst[1], (int)round1, (int)rounds, nvalid);
      SP2_LAUNCH_CHECK();
      if (after_first && !recorded) { SP2_CUDA_OK(cudaEventRecord(after_first, ctx->stream)); recorded = true; }
      break;
    }
    const u64 nv = round1 <= 2 ? nvalid : ~0ull;
    if (!sharded && len_in <= SC_ROLE_LEN) {
      const u64 per_cta = (SC_ROLE_THREADS / 96) * 32;
      u64 nb = (P + per_cta - 1) / per_cta; if (nb > (u64)ctx->num_sms) nb = ctx->num_sms;
      if (round1 > 1) {
        k_quad_round_roles<true><<<(unsigned)nb, SC_ROLE_THREADS, 0, ctx->stream>>>(st, src[0], src[1], dst[0], dst[1], P, (int)round1, (int)rounds, nv);
        std::swap(src[0], dst[0]); std::swap(src[1], dst[1]);
      } else {
        k_quad_round_roles<false><<<(unsigned)nb, SC_ROLE_THREADS, 0, ctx->stream>>>(st, src[0], src[1], dst[0], dst[1], P, (int)round1, (int)rounds, nv);
      }
    } else {
      u64 nb = (P + SC_THREADS - 1) / SC_THREADS; if (nb > target) nb = target; if (nb == 0) nb = 1;
      if (round1 > 1) k_quad_round<true><<<(unsigned)nb, SC_THREADS, 0, ctx->stream>>>(st, A, B, P, (int)round1, (int)rounds, nv, dc);
      else k_quad_round<false><<<(unsigned)nb, SC_THREADS, 0, ctx->stream>>>(st, A, B, P, (int)round1, (int)rounds, nv, dc);
    }
    SP2_LAUNCH_CHECK();
    if (after_first && !recorded) { SP2_CUDA_OK(cudaEventRecord(after_first, ctx->stream)); recorded = true; }
  }
  return SP2_OK;
}

}  // namespace sp2

extern "C" {

/* debug: clock64() stamps (SM cycles) of the last finalised sum-check round: [tail round start, finalize start,
 * squeeze start, message built, hashed, challenge ready, finalize end] */
/* debug: per-round %globaltimer stamps (ns) of the last persistent cubic kernel: rounds x [start, own compute done, all CTAs
 * arrived, finalised] */
/* tables of at most this many entries are finished by the single-CTA tail kernels; larger ones go through the persistent
 * multi-CTA kernels (bench.py derives from it which rounds k_cubic_persist covers) */
uint64_t sp2_sc_tail_len(void) { return SC_TAIL_LEN; }
/* tables of at most this many entries (going into a round >= 2) leave the persistent kernels for the pipelined multi-CTA kernels; 0 = disabled */
uint64_t sp2_sc_mid_len(void) { return use_mid_pipe() ? (1ull << MP_LOG_LEN_IN) : 0; }
#ifdef SP2_TAIL_TRACE
int32_t sp2_debug_tail_trace(long long *out) { return (int32_t)cudaMemcpyFromSymbol(out, g_tail_trace, sizeof(long long) * 2 * 40 * 10); }
#endif

int32_t sp2_debug_sc_round_profile(sp2_ctx *ctx, uint64_t *out, uint32_t rounds) {
  cudaSetDevice(ctx->device);
  if (!ctx->slot[14] || rounds > SC_MAX_ROUNDS) return set_error(ctx, SP2_ERR_INTERNAL, "no sum-check has run");
  ScState *st = (ScState *)ctx->slot[14];
  SP2_CUDA_OK(cudaMemcpyAsync(out, st->prof, (size_t)rounds * 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

int32_t sp2_debug_sc_clocks(sp2_ctx *ctx, uint64_t *out7 /* 11 values */) {
  cudaSetDevice(ctx->device);
  if (!ctx->slot[14]) return set_error(ctx, SP2_ERR_INTERNAL, "no sum-check has run");
  ScState *st = (ScState *)ctx->slot[14];
  SP2_CUDA_OK(cudaMemcpyAsync(out7, st->clk, 7 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(out7 + 7, st->gt, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(out7 + 11, &st->clk[7], sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(out7 + 12, &st->gt[4], sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

/* EqSumCheckInstance::evaluation_points_zero_check_round0 (src/sumcheck.rs:1163-1271) on device-resident tables of 2^l entries
 * (not modified): out3 = (eval_0, eval_2, eval_3) of the first round of a zero-check */
int32_t sp2_sc_zero_check_round0_dev(sp2_ctx *ctx, const uint64_t *taus, uint32_t l, const void *dA, const void *dB, uint64_t *out3) {
  cudaSetDevice(ctx->device);
  if (l < 1 || l > SC_MAX_ROUNDS) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "zero_check_round0: 1 <= num_rounds <= 40");
  ScState *st;
  sp2_transcript_state ts0; memset(&ts0, 0, sizeof(ts0));
  const uint64_t zero4[4] = {0, 0, 0, 0};
  SP2_TRY(sc_state_upload(ctx, &st, zero4, taus, l, &ts0));
  const int first_half = (int)l / 2, second_half = (int)l - first_half;
  void *eqs;
  const size_t nleft = (size_t)1 << (first_half > 0 ? first_half : 1), nright = (size_t)2 << second_half;
  SP2_TRY(scratch(ctx, 13, (nleft + nright) * sizeof(fe), &eqs));
  fe *eq_left = (fe *)eqs, *eq_right = eq_left + nleft;
  k_cubic_init<<<2, 1024, 0, ctx->stream>>>(st, (int)l, eq_left, eq_right);
  SP2_LAUNCH_CHECK();
  const u64 P = (u64)1 << (l - 1);
  const fe *el = nullptr, *er; u32 sh = 0;
  if (1 < first_half) { const int kl = first_half - 1; el = eq_left + (((size_t)1 << kl) - 1); er = eq_right + (((size_t)1 << second_half) - 1); sh = (u32)second_half; }
  else er = eq_right + (((size_t)1 << (l - 1)) - 1);
  u64 nb = (P + SC_THREADS - 1) / SC_THREADS; if (nb > (u64)ctx->num_sms * 4) nb = (u64)ctx->num_sms * 4;
  void *part; SP2_TRY(scratch(ctx, 7, (nb + 4) * sizeof(fe), &part));
  k_zero_check_round0<<<(unsigned)nb, SC_THREADS, 0, ctx->stream>>>((const fe *)dA, (const fe *)dB, P, el, er, sh, (fe *)part);
  SP2_LAUNCH_CHECK();
  k_zero_check_finish<<<1, 256, 0, ctx->stream>>>(st, (const fe *)part, (u32)nb, (fe *)part + nb);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(out3, (fe *)part + nb, 3 * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

int32_t sp2_sumcheck_cubic_prove_dev(sp2_ctx *ctx, const uint64_t *claim, const uint64_t *taus, uint32_t l,
                                     void *dA, void *dB, void *dC,
                                     sp2_transcript_state *ts, uint64_t *polys, uint64_t *r, uint64_t *claims) {
  cudaSetDevice(ctx->device);
  if (l < 1 || l > SC_MAX_ROUNDS) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "sumcheck_cubic: 1 <= num_rounds <= 40");
  (void)claim;   // the cubic prover sums t(0), t(1), t(inf) directly; the claim is implied by the tables
  ScState *st;
  SP2_TRY(sc_state_upload(ctx, &st, claim, taus, l, ts));
  SP2_TRY(sumcheck_cubic_enqueue(ctx, st, l, (fe *)dA, (fe *)dB, (fe *)dC));
  return sc_state_download(ctx, st, ts, polys, 4, r, claims, 3, l);
}

int32_t sp2_sumcheck_quad_prove_dev(sp2_ctx *ctx, const uint64_t *claim, uint32_t rounds, void *dA, void *dB,
                                    sp2_transcript_state *ts, uint64_t *polys, uint64_t *r, uint64_t *claims) {
  cudaSetDevice(ctx->device);
  if (rounds < 1 || rounds > SC_MAX_ROUNDS) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "sumcheck_quad: 1 <= num_rounds <= 40");
  ScState *st;
  SP2_TRY(sc_state_upload(ctx, &st, claim, nullptr, rounds, ts));
  SP2_TRY(sumcheck_quad_enqueue(ctx, st, rounds, (fe *)dA, (fe *)dB, ~0ull, nullptr));
  return sc_state_download(ctx, st, ts, polys, 3, r, claims, 2, rounds);
}

int32_t sp2_sumcheck_cubic_prove(sp2_ctx *ctx, const uint64_t *claim, const uint64_t *taus, uint32_t l,
                                 const uint64_t *A, const uint64_t *B, const uint64_t *C,
                                 sp2_transcript_state *ts, uint64_t *polys, uint64_t *r, uint64_t *claims) {
  cudaSetDevice(ctx->device);
  if (l < 1 || l > 32) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "sumcheck_cubic: 1 <= num_rounds <= 32");
  const size_t bytes = ((size_t)1 << l) * sizeof(fe);
  void *dA, *dB, *dC;
  SP2_TRY(scratch(ctx, 0, bytes, &dA)); SP2_TRY(scratch(ctx, 1, bytes, &dB)); SP2_TRY(scratch(ctx, 2, bytes, &dC));
  SP2_CUDA_OK(cudaMemcpyAsync(dA, A, bytes, cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(dB, B, bytes, cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(dC, C, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return sp2_sumcheck_cubic_prove_dev(ctx, claim, taus, l, dA, dB, dC, ts, polys, r, claims);
}

int32_t sp2_sumcheck_quad_prove(sp2_ctx *ctx, const uint64_t *claim, uint32_t rounds,
                                const uint64_t *A, const uint64_t *B,
                                sp2_transcript_state *ts, uint64_t *polys, uint64_t *r, uint64_t *claims) {
  cudaSetDevice(ctx->device);
  if (rounds < 1 || rounds > 32) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "sumcheck_quad: 1 <= num_rounds <= 32");
  const size_t bytes = ((size_t)1 << rounds) * sizeof(fe);
  void *dA, *dB;
  SP2_TRY(scratch(ctx, 0, bytes, &dA)); SP2_TRY(scratch(ctx, 1, bytes, &dB));
  SP2_CUDA_OK(cudaMemcpyAsync(dA, A, bytes, cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(dB, B, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return sp2_sumcheck_quad_prove_dev(ctx, claim, rounds, dA, dB, ts, polys, r, claims);
}

}  // extern "C"
