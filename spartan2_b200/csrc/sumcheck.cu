// Sum-check provers on the device — whole loops, no host round trip per round.
//
// Restates reference src/sumcheck.rs:
//   prove_cubic_with_three_inputs (:502-571) with EqSumCheckInstance (:934-1405): split-eq (Gruen)
//     weights + the two-sum trick (t(0), t(inf)); derive_from_claim (:1277-1324); bound (:1399-1405)
//   prove_quad (:190-247) with compute_eval_points_quad (:128-174)
// and the per-round transcript step (absorb b"p" / squeeze b"c", :536-548) of
// src/provider/keccak.rs:70-99.
//
// B200 design (not the reference's rayon fold/reduce):
//   * one launch per round; the launch for round i first binds the tables to challenge r_{i-1}
//     (in place: the thread that reads T[j], T[j+L/4], T[j+L/2], T[j+3L/4] writes T'[j], T'[j+L/4])
//     and evaluates round i on the freshly bound values while they are still in registers — every
//     table element is read once and written once per round (368*T bytes in total);
//   * per-thread 544-bit delayed-reduction accumulators, warp-shuffle + block reduction, one
//     partial pair per CTA, and the LAST CTA to finish (atomic ticket) sums the partials, does the
//     round's scalar algebra, runs the Keccak transcript on two warps and publishes the challenge
//     in device memory for the next launch: rounds are chained on the stream without host syncs;
//   * the claim is tracked in "t-space" (claim / eval_eq_left), so the only inversions are of the
//     tau_i themselves — done once, in parallel, by the init kernel — instead of one per round.
// Every emitted value is the canonical representative of the same field element the reference
// computes, hence bit-identical.
#include <string.h>
#include "ctx.cuh"
#include "keccak.cuh"
#include "polys.cuh"

using namespace sp2;

namespace sp2 {

constexpr int SC_MAX_ROUNDS = 40;
constexpr int SC_MAX_BLOCKS = 2048;
constexpr int SC_THREADS = 256;

struct ScState {
  DevTranscript ts;
  fe c;                       // cubic: claim / eval_eq_left ("t-space" claim);  quad: the claim
  fe p;                       // eval_eq_left (sumcheck.rs:951)
  fe L0, SL;                  // p*(1-tau_i), p*(2tau_i-1) for the round being evaluated
  u32 ticket, error, l, pad;
  fe taus[SC_MAX_ROUNDS];
  fe tau_inv[SC_MAX_ROUNDS];
  // ---- everything above is uploaded by the host; everything below is produced on the device ----
  fe r[SC_MAX_ROUNDS];
  fe polys[SC_MAX_ROUNDS * 4];
  fe claims[4];
  fe partial[2 * SC_MAX_BLOCKS];
};

struct FinSmem {
  unsigned char buf[2304];
  fe f[8];
  fe g[8];
  fe ch;
  fe red[2 * 32];
  int is_last;
};

__device__ __forceinline__ fe ld_state(const fe *p) {   // device-produced scalars: bypass L1
  fe r; u64 a, b, c, d;
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  r.v[0] = (u32)a; r.v[1] = (u32)(a >> 32); r.v[2] = (u32)b; r.v[3] = (u32)(b >> 32);
  r.v[4] = (u32)c; r.v[5] = (u32)(c >> 32); r.v[6] = (u32)d; r.v[7] = (u32)(d >> 32);
  return r;
}

// publish this CTA's partial sums and elect the last CTA of the grid
__device__ __forceinline__ bool publish_and_elect(ScState *st, fe (&x)[2], FinSmem &sm) {
  const int nblocks = gridDim.x * gridDim.y, bid = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) {
    stg_fe(&st->partial[2 * bid], x[0]);
    stg_fe(&st->partial[2 * bid + 1], x[1]);
    __threadfence();
    u32 t = atomicAdd(&st->ticket, 1u);
    sm.is_last = (t == (u32)nblocks - 1);
  }
  __syncthreads();
  if (!sm.is_last) return false;
  __threadfence();
  return true;
}

__device__ __forceinline__ void sum_partials(ScState *st, fe (&x)[2], FinSmem &sm) {
  const int nblocks = gridDim.x * gridDim.y;
  x[0] = Fq::zero(); x[1] = Fq::zero();
  for (int b = threadIdx.x; b < nblocks; b += blockDim.x) {
    x[0] = Fq::add(x[0], ld_state(&st->partial[2 * b]));
    x[1] = Fq::add(x[1], ld_state(&st->partial[2 * b + 1]));
  }
  __syncthreads();            // sm.red is reused
  block_sum_fq<2>(x, sm.red);
}

// ---------------------------------------------------------------------------------------------
// cubic: s(X) = l(X) * p * t(X),  l(X) = (1-tau) + (2tau-1) X,  t(X) = t0 + tb X + tinf X^2
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cubic_finalize(ScState *st, int round1, int l, const fe *A, const fe *B, const fe *C,
                                               FinSmem &sm) {
  const int tid = threadIdx.x, i = round1 - 1;
  fe x[2];
  sum_partials(st, x, sm);
  if (tid == 0) {
    const fe t0 = x[0], ti = x[1];
    const fe tau = ld_state(&st->taus[i]);
    const fe l0 = Fq::sub(Fq::one(), tau);
    // claim = s(0) + s(1)  =>  c = l(0) t(0) + l(1) t(1),  l(1) = tau      (derive_from_claim)
    const fe t1 = Fq::mul(Fq::sub(ld_state(&st->c), Fq::mul(l0, t0)), ld_state(&st->tau_inv[i]));
    sm.f[0] = t0; sm.f[1] = Fq::sub(Fq::sub(t1, t0), ti); sm.f[2] = ti;
  }
  __syncthreads();
  if (tid < 6) {              // L0*t0, L0*tb, SL*t0, L0*tinf, SL*tb, SL*tinf
    const bool use_sl = (tid == 2) | (tid == 4) | (tid == 5);
    const int which = (tid == 0 || tid == 2) ? 0 : ((tid == 1 || tid == 4) ? 1 : 2);
    sm.g[tid] = Fq::mul(ld_state(use_sl ? &st->SL : &st->L0), sm.f[which]);
  }
  __syncthreads();
  if (tid < 4) {              // coefficients low -> high (UniPoly, univariate.rs:102-118)
    fe co = tid == 0 ? sm.g[0] : tid == 1 ? Fq::add(sm.g[1], sm.g[2]) : tid == 2 ? Fq::add(sm.g[3], sm.g[4]) : sm.g[5];
    stg_fe(&st->polys[4 * i + tid], co);
    sm.f[4 + tid] = Fq::from_mont(co);
  }
  __syncthreads();
  if (tid == 0) {             // absorb(b"p", poly): all coefficients but the linear one, to_repr() LE
    const unsigned char lab = 'p';
    ts_push_bytes(&st->ts, &lab, 1);
    ts_push_fe_le(&st->ts, sm.f[4]); ts_push_fe_le(&st->ts, sm.f[6]); ts_push_fe_le(&st->ts, sm.f[7]);
  }
  __syncthreads();
  ts_squeeze_block(&st->ts, "c", 1, sm.buf, &sm.ch);
  const fe r = sm.ch;
  if (tid == 0) {             // next t-space claim: t(r)
    stg_fe(&st->r[i], r);
    stg_fe(&st->c, Fq::add(sm.f[0], Fq::mul(r, Fq::add(sm.f[1], Fq::mul(r, sm.f[2])))));
    st->ticket = 0;
  }
  if (tid == 32) {            // bound(): p <- p * (1 - tau - r + 2 r tau) = p * l(r)
    const fe tau = ld_state(&st->taus[i]);
    const fe l0 = Fq::sub(Fq::one(), tau), sl = Fq::sub(tau, l0);
    const fe pn = Fq::mul(ld_state(&st->p), Fq::add(l0, Fq::mul(sl, r)));
    stg_fe(&st->p, pn);
    if (round1 < l) {
      const fe tn = ld_state(&st->taus[i + 1]);
      const fe l0n = Fq::sub(Fq::one(), tn), sln = Fq::sub(tn, l0n);
      stg_fe(&st->L0, Fq::mul(pn, l0n));
      stg_fe(&st->SL, Fq::mul(pn, sln));
    }
  }
  if (round1 == l && (tid == 64 || tid == 96 || tid == 128)) {   // final claims: bind the length-2 tables
    const fe *T = tid == 64 ? A : tid == 96 ? B : C;
    stg_fe(&st->claims[(tid - 64) / 32], bind_pair(ld_state(T), ld_state(T + 1), r));
  }
}

// MODE 0: two-level split-eq.  Thread owns one x_in (coalesced across the warp) and walks x_out:
//         acc += el[x_out] * v(x_out, x_in), then one multiply by er[x_in]  (sumcheck.rs:1045-1100,
//         with the roles of the inner/outer sums swapped so the flush happens once per thread).
// MODE 1: generic — weight = el[id >> sh] * er[id & mask] (tiny tables) or er[id] (second half,
//         sumcheck.rs:1107-1142).
template <bool FUSED, int MODE>
__global__ void __launch_bounds__(SC_THREADS, 2)
k_cubic_round(ScState *st, fe *A, fe *B, fe *C, u64 P, int round1, int l, const fe *el, const fe *er, u32 out_len, u32 sh) {
  __shared__ FinSmem sm;
  fe r;
  if (FUSED) r = ld_state(&st->r[round1 - 2]);
  Fq::acc acc0 = Fq::acc_zero(), acci = Fq::acc_zero();

  auto pair = [&](u64 id, const fe &w) {
    fe a0, a1, b0, b1, c0;
    if (FUSED) {
      const fe a00 = ldg_fe(A + id), a01 = ldg_fe(A + id + P), a10 = ldg_fe(A + id + 2 * P), a11 = ldg_fe(A + id + 3 * P);
      const fe b00 = ldg_fe(B + id), b01 = ldg_fe(B + id + P), b10 = ldg_fe(B + id + 2 * P), b11 = ldg_fe(B + id + 3 * P);
      const fe c00 = ldg_fe(C + id), c01 = ldg_fe(C + id + P), c10 = ldg_fe(C + id + 2 * P), c11 = ldg_fe(C + id + 3 * P);
      a0 = bind_pair(a00, a10, r); a1 = bind_pair(a01, a11, r);
      stg_fe(A + id, a0); stg_fe(A + id + P, a1);
      b0 = bind_pair(b00, b10, r); b1 = bind_pair(b01, b11, r);
      stg_fe(B + id, b0); stg_fe(B + id + P, b1);
      c0 = bind_pair(c00, c10, r);
      stg_fe(C + id, c0); stg_fe(C + id + P, bind_pair(c01, c11, r));
    } else {
      a0 = ldg_fe(A + id); a1 = ldg_fe(A + id + P);
      b0 = ldg_fe(B + id); b1 = ldg_fe(B + id + P);
      c0 = ldg_fe(C + id);
    }
    const fe t0e = Fq::sub(Fq::mul(a0, b0), c0);
    const fe tie = Fq::mul(Fq::sub(a1, a0), Fq::sub(b1, b0));
    Fq::mul_acc(acc0, w, t0e);
    Fq::mul_acc(acci, w, tie);
  };

  fe x[2];
  if (MODE == 0) {
    const u64 xi = (u64)blockIdx.x * SC_THREADS + threadIdx.x;
    for (u32 xo = blockIdx.y; xo < out_len; xo += gridDim.y) pair(((u64)xo << sh) | xi, ldg_fe_ro(el + xo));
    const fe wr = ldg_fe_ro(er + xi);
    x[0] = Fq::mul(wr, Fq::acc_reduce(acc0));
    x[1] = Fq::mul(wr, Fq::acc_reduce(acci));
  } else {
    const u64 mask = ((u64)1 << sh) - 1;
    for (u64 id = (u64)blockIdx.x * SC_THREADS + threadIdx.x; id < P; id += (u64)gridDim.x * SC_THREADS) {
      fe w = ldg_fe_ro(er + (el ? (id & mask) : id));
      if (el) w = Fq::mul(ldg_fe_ro(el + (id >> sh)), w);
      pair(id, w);
    }
    x[0] = Fq::acc_reduce(acc0);
    x[1] = Fq::acc_reduce(acci);
  }
  block_sum_fq<2>(x, sm.red);
  if (!publish_and_elect(st, x, sm)) return;
  cubic_finalize(st, round1, l, A, B, C, sm);
}

// init: block 0 builds the split-eq prefix tables and the round-1 constants; block b >= 1 inverts tau_{b-1}
__global__ void __launch_bounds__(1024) k_cubic_init(ScState *st, int l, fe *eq_left, fe *eq_right) {
  const int first_half = l / 2, second_half = l - first_half;
  if (blockIdx.x == 0) {
    eq_prefix_block(st->taus + 1, first_half > 0 ? first_half - 1 : 0, eq_left);
    eq_prefix_block(st->taus + first_half, second_half, eq_right);
    if (threadIdx.x == 0) {
      const fe tau = ldg_fe(&st->taus[0]);
      const fe l0 = Fq::sub(Fq::one(), tau);
      stg_fe(&st->p, Fq::one());
      stg_fe(&st->L0, l0);
      stg_fe(&st->SL, Fq::sub(tau, l0));
    }
  } else if (threadIdx.x == 0) {
    const fe tau = ldg_fe(&st->taus[blockIdx.x - 1]);
    if (Fq::is_zero(tau)) atomicOr(&st->error, 1u);   // reference takes fallback_three_inputs (:1327-1396)
    stg_fe(&st->tau_inv[blockIdx.x - 1], Fq::inv(tau));
  }
}

// ---------------------------------------------------------------------------------------------
// quadratic: eval0 = sum a_lo b_lo, tinf = sum (a_hi-a_lo)(b_hi-b_lo)   (sumcheck.rs:128-174)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void quad_finalize(ScState *st, int round1, int rounds, const fe *A, const fe *B, FinSmem &sm) {
  const int tid = threadIdx.x, i = round1 - 1;
  fe x[2];
  sum_partials(st, x, sm);
  if (tid == 0) {
    // from_evals([e0, claim-e0, 2claim-3e0+2tinf]) = [e0, claim - 2 e0 - tinf, tinf]  (sumcheck.rs:205-216)
    const fe e0 = x[0], ti = x[1], claim = ld_state(&st->c);
    const fe b = Fq::sub(Fq::sub(claim, Fq::dbl(e0)), ti);
    sm.f[0] = e0; sm.f[1] = b; sm.f[2] = ti;
    stg_fe(&st->polys[4 * i], e0); stg_fe(&st->polys[4 * i + 1], b); stg_fe(&st->polys[4 * i + 2], ti);
    const unsigned char lab = 'p';
    ts_push_bytes(&st->ts, &lab, 1);
    ts_push_fe_le(&st->ts, Fq::from_mont(e0)); ts_push_fe_le(&st->ts, Fq::from_mont(ti));
  }
  __syncthreads();
  ts_squeeze_block(&st->ts, "c", 1, sm.buf, &sm.ch);
  const fe r = sm.ch;
  if (tid == 0) {
    stg_fe(&st->r[i], r);
    stg_fe(&st->c, Fq::add(sm.f[0], Fq::mul(r, Fq::add(sm.f[1], Fq::mul(r, sm.f[2])))));
    st->ticket = 0;
  }
  if (round1 == rounds && (tid == 64 || tid == 96)) {
    const fe *T = tid == 64 ? A : B;
    stg_fe(&st->claims[(tid - 64) / 32], bind_pair(ld_state(T), ld_state(T + 1), r));
  }
}

template <bool FUSED>
__global__ void __launch_bounds__(SC_THREADS, 2)
k_quad_round(ScState *st, fe *A, fe *B, u64 P, int round1, int rounds) {
  __shared__ FinSmem sm;
  fe r;
  if (FUSED) r = ld_state(&st->r[round1 - 2]);
  Fq::acc acc0 = Fq::acc_zero(), acci = Fq::acc_zero();
  for (u64 id = (u64)blockIdx.x * SC_THREADS + threadIdx.x; id < P; id += (u64)gridDim.x * SC_THREADS) {
    fe a0, a1, b0, b1;
    if (FUSED) {
      const fe a00 = ldg_fe(A + id), a01 = ldg_fe(A + id + P), a10 = ldg_fe(A + id + 2 * P), a11 = ldg_fe(A + id + 3 * P);
      const fe b00 = ldg_fe(B + id), b01 = ldg_fe(B + id + P), b10 = ldg_fe(B + id + 2 * P), b11 = ldg_fe(B + id + 3 * P);
      a0 = bind_pair(a00, a10, r); a1 = bind_pair(a01, a11, r);
      stg_fe(A + id, a0); stg_fe(A + id + P, a1);
      b0 = bind_pair(b00, b10, r); b1 = bind_pair(b01, b11, r);
      stg_fe(B + id, b0); stg_fe(B + id + P, b1);
    } else {
      a0 = ldg_fe(A + id); a1 = ldg_fe(A + id + P);
      b0 = ldg_fe(B + id); b1 = ldg_fe(B + id + P);
    }
    Fq::mul_acc(acc0, a0, b0);
    Fq::mul_acc(acci, Fq::sub(a1, a0), Fq::sub(b1, b0));
  }
  fe x[2] = {Fq::acc_reduce(acc0), Fq::acc_reduce(acci)};
  block_sum_fq<2>(x, sm.red);
  if (!publish_and_elect(st, x, sm)) return;
  quad_finalize(st, round1, rounds, A, B, sm);
}

static int sc_state_upload(sp2_ctx *ctx, ScState **d_st, const uint64_t *claim, const uint64_t *taus, uint32_t l,
                           const sp2_transcript_state *ts) {
  void *d, *h;
  SP2_TRY(scratch(ctx, 14, sizeof(ScState), &d));
  SP2_TRY(pinned(ctx, sizeof(ScState), &h));
  ScState *hs = (ScState *)h;
  const size_t head = offsetof(ScState, r);
  memset(hs, 0, head);
  hs->ts.round = ts->round; hs->ts.pending_len = 0; memcpy(hs->ts.state, ts->state, 64);
  memcpy(&hs->c, claim, sizeof(fe));
  hs->l = l;
  if (taus) memcpy(hs->taus, taus, (size_t)l * sizeof(fe));
  SP2_CUDA_OK(cudaMemcpyAsync(d, hs, head, cudaMemcpyHostToDevice, ctx->stream));
  *d_st = (ScState *)d;
  return SP2_OK;
}

static int sc_state_download(sp2_ctx *ctx, ScState *d_st, sp2_transcript_state *ts, uint64_t *polys, int ncoef, uint64_t *r,
                             uint64_t *claims, int nclaims, uint32_t l) {
  void *h;
  SP2_TRY(pinned(ctx, sizeof(ScState), &h));
  ScState *hs = (ScState *)h;
  const size_t upto = offsetof(ScState, partial);
  SP2_CUDA_OK(cudaMemcpyAsync(hs, d_st, upto, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (hs->error) return set_error(ctx, SP2_ERR_UNSUPPORTED, "sum-check: tau_i == 0 needs the reference's fallback_three_inputs path");
  ts->round = (uint16_t)hs->ts.round; memcpy(ts->state, hs->ts.state, 64);
  for (uint32_t i = 0; i < l; i++) memcpy(polys + (size_t)i * ncoef * 4, &hs->polys[4 * i], (size_t)ncoef * sizeof(fe));
  memcpy(r, hs->r, (size_t)l * sizeof(fe));
  memcpy(claims, hs->claims, (size_t)nclaims * sizeof(fe));
  return SP2_OK;
}

// enqueue the whole cubic sum-check on ctx->stream (state already uploaded)
int sumcheck_cubic_enqueue(sp2_ctx *ctx, ScState *st, uint32_t l, fe *A, fe *B, fe *C) {
  const int first_half = (int)l / 2, second_half = (int)l - first_half;
  void *eqs;
  const size_t nleft = (size_t)1 << (first_half > 0 ? first_half : 1), nright = (size_t)2 << second_half;
  SP2_TRY(scratch(ctx, 13, (nleft + nright) * sizeof(fe), &eqs));
  fe *eq_left = (fe *)eqs, *eq_right = eq_left + nleft;
  k_cubic_init<<<1 + l, 1024, 0, ctx->stream>>>(st, (int)l, eq_left, eq_right);
  SP2_LAUNCH_CHECK();
  const unsigned target = (unsigned)ctx->num_sms * 2;
  for (uint32_t round1 = 1; round1 <= l; round1++) {
    const bool fused = round1 > 1;
    const u64 P = (u64)1 << (l - round1);                       // pairs evaluated this round
    const bool in_first = (int)round1 < first_half;
    const fe *el = nullptr, *er; u32 out_len = 1, sh = 0;
    if (in_first) {
      const int kl = first_half - (int)round1;
      el = eq_left + (((size_t)1 << kl) - 1); out_len = 1u << kl;
      er = eq_right + (((size_t)1 << second_half) - 1); sh = (u32)second_half;
    } else {
      er = eq_right + (((size_t)1 << (l - round1)) - 1);
    }
    const u64 in_len = (u64)1 << sh;
    if (in_first && in_len >= SC_THREADS && in_len / SC_THREADS <= target) {
      dim3 grid((unsigned)(in_len / SC_THREADS), 1);
      unsigned gy = target / grid.x; if (gy < 1) gy = 1; if (gy > out_len) gy = out_len;
      grid.y = gy;
      if (fused) k_cubic_round<true, 0><<<grid, SC_THREADS, 0, ctx->stream>>>(st, A, B, C, P, (int)round1, (int)l, el, er, out_len, sh);
      else k_cubic_round<false, 0><<<grid, SC_THREADS, 0, ctx->stream>>>(st, A, B, C, P, (int)round1, (int)l, el, er, out_len, sh);
    } else {
      u64 nb = (P + SC_THREADS - 1) / SC_THREADS; if (nb > target) nb = target;
      if (fused) k_cubic_round<true, 1><<<(unsigned)nb, SC_THREADS, 0, ctx->stream>>>(st, A, B, C, P, (int)round1, (int)l, el, er, out_len, sh);
      else k_cubic_round<false, 1><<<(unsigned)nb, SC_THREADS, 0, ctx->stream>>>(st, A, B, C, P, (int)round1, (int)l, el, er, out_len, sh);
    }
    SP2_LAUNCH_CHECK();
  }
  return SP2_OK;
}

int sumcheck_quad_enqueue(sp2_ctx *ctx, ScState *st, uint32_t rounds, fe *A, fe *B) {
  const unsigned target = (unsigned)ctx->num_sms * 2;
  for (uint32_t round1 = 1; round1 <= rounds; round1++) {
    const u64 P = (u64)1 << (rounds - round1);
    u64 nb = (P + SC_THREADS - 1) / SC_THREADS; if (nb > target) nb = target;
    if (round1 > 1) k_quad_round<true><<<(unsigned)nb, SC_THREADS, 0, ctx->stream>>>(st, A, B, P, (int)round1, (int)rounds);
    else k_quad_round<false><<<(unsigned)nb, SC_THREADS, 0, ctx->stream>>>(st, A, B, P, (int)round1, (int)rounds);
    SP2_LAUNCH_CHECK();
  }
  return SP2_OK;
}

}  // namespace sp2

extern "C" {

int32_t sp2_sumcheck_cubic_prove_dev(sp2_ctx *ctx, const uint64_t *claim, const uint64_t *taus, uint32_t l,
                                     void *dA, void *dB, void *dC,
                                     sp2_transcript_state *ts, uint64_t *polys, uint64_t *r, uint64_t *claims) {
  cudaSetDevice(ctx->device);
  if (l < 1 || l > SC_MAX_ROUNDS) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "sumcheck_cubic: 1 <= num_rounds <= 40");
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
  SP2_TRY(sumcheck_quad_enqueue(ctx, st, rounds, (fe *)dA, (fe *)dB));
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
