// NeutronNova building blocks on the device (configs 3 and 5): the multi-folding NIFS rounds over instance layers,
// witness/vector folding, the pow-weighted cubic and the quadratic evaluation points with per-round binding (the ZK
// drivers take one challenge per round from the in-circuit verifier, so these are per-round {eval, bind} pairs, not
// whole loops — SURVEY.md §8b), and commitment folding.
//
// Restates reference
//   src/polys/power.rs:65-86                 PowPolynomial::split_evals
//   src/neutronnova_zk.rs:78-178, 738-776    suffix_weight_full, prove_helper, fold_abc_pair (standard field path;
//                                            the i64 "small value" layers of :255-432 are a CPU representation
//                                            optimisation that yields the same field values)
//   src/r1cs/mod.rs:153-166, 570-660         weights_from_r, R1CSWitness::fold_multiple
//   src/sumcheck.rs:262-342, 366-498         compute_eval_points_cubic_with_additive_term(_with_outer_pow)
//   src/sumcheck.rs:128-174                  compute_eval_points_quad
//   src/provider/pcs/hyrax_pc.rs:737-793     fold_commitments (msm_shared_weights, msm.rs:228-356, as group elements)
//
// Layers are stored layer-major; live layer q of round t sits at slot q * stride (stride = 2^t): folding pair
// (2p, 2p+1) in place into the even slot needs no compaction and no cross-thread hazards.
// every kernel of this file is latency-bound (<= 2^15-entry tables per instance, one launch per round): see field.cuh
#define SP2_FQ_OUTLINE 1
#include <string.h>
#include <vector>
#include "ctx.cuh"
#include "curve.cuh"
#include "devutil.cuh"
#include "host_transcript.h"
#include "msm.cuh"
#include "polys.cuh"

using namespace sp2;

namespace {

constexpr int NF_THREADS = 256;

__device__ __forceinline__ fe fq_pow_small(fe base, u32 e) {      // base^e, e < 2^32 (binary)
  fe r = Fq::one();
  while (e) { if (e & 1u) r = Fq::mul(r, base); base = Fq::sqr(base); e >>= 1; }
  return r;
}
__global__ void k_pow_split(const fe *t, u32 left, u32 right, fe *out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  const fe tt = ldg_fe(t);
  if (i < left) stg_fe(out + i, fq_pow_small(tt, i));
  else if (i < left + right) stg_fe(out + i, fq_pow_small(fq_pow_small(tt, left), i - left));
}

// partial sums -> out (one CTA)
template <int NV>
__global__ void __launch_bounds__(NF_THREADS) k_reduce_partials(const fe *partials, u32 nparts, fe *out) {
  __shared__ fe red[NV * 32];
  fe x[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) x[k] = Fq::zero();
  for (u32 b = threadIdx.x; b < nparts; b += blockDim.x)
#pragma unroll
    for (int k = 0; k < NV; k++) x[k] = Fq::add(x[k], ldg_fe(partials + (size_t)b * NV + k));
  block_sum_fq<NV>(x, red);
  if (threadIdx.x == 0)
#pragma unroll
    for (int k = 0; k < NV; k++) stg_fe(out + k, x[k]);
}

// (defined with the batched sum-check kernels below) accumulate-and-publish: every CTA adds its NV sums, the last one publishes them
template <int NV>
__device__ __forceinline__ void nn_publish_last(fe (&x)[NV], fe *, fe *red, int *is_last, u32 *ticket, fe *mail_out, u32 *mail_flag, u32 seq, u32 groups = 0, u32 group = 0);

// One NIFS round: grid (i-chunks, pairs); thread owns j (the contiguous index), walks its i-chunk:
//   acc += f[i] * v(i, j), then * e_left[j]   (prove_helper's nested sums with inner/outer swapped: one flush per thread)
__global__ void __launch_bounds__(NF_THREADS) k_nifs_round(u32 t, const fe *rhos, u32 ell_b, u32 left, u32 right, const fe *E, const fe *A,
                                                           const fe *B, const fe *C, u64 N, u64 stride, fe *partials, u32 pair_offset = 0,
                                                           u32 *ticket = nullptr, fe *mail_out = nullptr, u32 *mail_flag = nullptr, u32 seq = 0) {
  __shared__ fe red[2 * 32];
  __shared__ fe wsh;
  __shared__ int is_last;
  const u32 p = blockIdx.y;
  const fe *el = E, *f = E + left;
  const fe *A1 = A + (u64)(2 * p) * stride * N, *A2 = A + (u64)(2 * p + 1) * stride * N;
  const fe *B1 = B + (u64)(2 * p) * stride * N, *B2 = B + (u64)(2 * p + 1) * stride * N;
  const fe *C1 = C + (u64)(2 * p) * stride * N;
  if (threadIdx.x == 0) {                       // suffix_weight_full(t, ell_b, p, rhos)
    fe w = Fq::one(); u32 k = p + pair_offset;        // pair_offset: a multi-GPU shard's first GLOBAL pair index
    for (u32 s = t + 1; s < ell_b; s++) { const fe r = ldg_fe(rhos + s); w = Fq::mul(w, (k & 1u) ? r : Fq::sub(Fq::one(), r)); k >>= 1; }
    wsh = w;
  }
  fe x[2] = {Fq::zero(), Fq::zero()};
  const u32 per = (right + gridDim.x - 1) / gridDim.x, i0 = blockIdx.x * per, i1 = min(right, i0 + per);
  for (u32 j = threadIdx.x; j < left; j += blockDim.x) {
    Fq::acc a0 = Fq::acc_zero(), aq = Fq::acc_zero();
    for (u32 i = i0; i < i1; i++) {
      const u64 k = (u64)i * left + j;
      const fe fi = ldg_fe_ro(f + i);
      const fe a1 = ldg_fe(A1 + k), b1 = ldg_fe(B1 + k);
      if (t != 0) Fq::mul_acc(a0, fi, Fq::sub(Fq::mul(a1, b1), ldg_fe(C1 + k)));
      Fq::mul_acc(aq, fi, Fq::mul(Fq::sub(ldg_fe(A2 + k), a1), Fq::sub(ldg_fe(B2 + k), b1)));
    }
    const fe ej = ldg_fe_ro(el + j);
    if (t != 0) x[0] = Fq::add(x[0], Fq::mul(ej, Fq::acc_reduce(a0)));
    x[1] = Fq::add(x[1], Fq::mul(ej, Fq::acc_reduce(aq)));
  }
  block_sum_fq<2>(x, red);
  if (ticket) {
    // publish from this kernel (no separate k_publish launch): all CTAs of all pairs add into ONE group of accumulators
    if (threadIdx.x == 0) { x[0] = Fq::mul(x[0], wsh); x[1] = Fq::mul(x[1], wsh); }
    __syncthreads();
    nn_publish_last<2>(x, nullptr, red, &is_last, ticket, mail_out, mail_flag, seq, 1, 0);
    return;
  }
  if (threadIdx.x == 0) {
    const size_t b = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
    stg_fe(partials + 2 * b, Fq::mul(x[0], wsh));
    stg_fe(partials + 2 * b + 1, Fq::mul(x[1], wsh));
  }
}

// fold pair (2p, 2p+1) into slot 2p for up to three tables
__global__ void __launch_bounds__(NF_THREADS) k_nifs_fold(fe *A, fe *B, fe *C, u64 N, u64 pairs, u64 stride, const fe *r) {
  const fe rr = ldg_fe_ro(r);
  fe *T[3] = {A, B, C};
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < pairs * N; q += (u64)gridDim.x * blockDim.x) {
    const u64 p = q / N, k = q - p * N;
#pragma unroll
    for (int s = 0; s < 3; s++) {
      if (!T[s]) continue;
      fe *lo = T[s] + (2 * p) * stride * N + k;
      stg_fe(lo, bind_pair(ldg_fe(lo), ldg_fe(T[s] + (2 * p + 1) * stride * N + k), rr));
    }
  }
}

__global__ void k_weights_from_r(const fe *r_bs, u32 ell, u32 n, fe *out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe w = Fq::one(); u32 k = i;
  for (u32 t = 0; t < ell; t++) { const fe r = ldg_fe(r_bs + t); w = Fq::mul(w, (k & 1u) ? r : Fq::sub(Fq::one(), r)); k >>= 1; }
  stg_fe(out + i, w);
}

// out[j] = sum_i w_i * Ws[i*dim + j]  (delayed reduction, fold_multiple's general path r1cs/mod.rs:633-648)
__global__ void __launch_bounds__(NF_THREADS) k_fold_vectors(const fe *Ws, u64 n, u64 dim, const fe *w, fe *out) {
  for (u64 j = (u64)blockIdx.x * blockDim.x + threadIdx.x; j < dim; j += (u64)gridDim.x * blockDim.x) {
    Fq::acc a = Fq::acc_zero();
    for (u64 i = 0; i < n; i++) Fq::mul_acc(a, ldg_fe_ro(w + i), ldg_fe(Ws + i * dim + j));
    stg_fe(out + j, Fq::acc_reduce(a));
  }
}

// evaluation points at 0, 2, 3 of  sum_x pow(x) (A B - C)  with pow = left (x) right outer product, len >= left
__global__ void __launch_bounds__(NF_THREADS) k_pow_cubic_outer(const fe *pl, u32 left, const fe *pr, const fe *A, const fe *B, const fe *C,
                                                                u64 len, fe *partials) {
  __shared__ fe red[3 * 32];
  const u64 right = len / left;
  fe x[3] = {Fq::zero(), Fq::zero(), Fq::zero()};
  const u64 per = (right + gridDim.x - 1) / gridDim.x, j0 = blockIdx.x * per, j1 = min(right, j0 + per);
  for (u32 i = threadIdx.x; i < left; i += blockDim.x) {
    Fq::acc a0 = Fq::acc_zero(), a2 = Fq::acc_zero(), a3 = Fq::acc_zero();
    for (u64 j = j0; j < j1; j++) {
      const u64 low = i + j * left, high = low + len;
      const fe tl = ldg_fe_ro(pr + j), th = ldg_fe_ro(pr + j + right);
      const fe al = ldg_fe(A + low), ah = ldg_fe(A + high), bl = ldg_fe(B + low), bh = ldg_fe(B + high), cl = ldg_fe(C + low), ch = ldg_fe(C + high);
      Fq::mul_acc(a0, tl, Fq::sub(Fq::mul(al, bl), cl));
      const fe dt = Fq::sub(th, tl), da = Fq::sub(ah, al), db = Fq::sub(bh, bl), dc = Fq::sub(ch, cl);
      fe tb = Fq::add(th, dt), ab = Fq::add(ah, da), bb = Fq::add(bh, db), cb = Fq::add(ch, dc);      // 2*high - low
      Fq::mul_acc(a2, tb, Fq::sub(Fq::mul(ab, bb), cb));
      tb = Fq::add(tb, dt); ab = Fq::add(ab, da); bb = Fq::add(bb, db); cb = Fq::add(cb, dc);          // 3*high - 2*low
      Fq::mul_acc(a3, tb, Fq::sub(Fq::mul(ab, bb), cb));
    }
    const fe w = ldg_fe_ro(pl + i);
    x[0] = Fq::add(x[0], Fq::mul(w, Fq::acc_reduce(a0)));
    x[1] = Fq::add(x[1], Fq::mul(w, Fq::acc_reduce(a2)));
    x[2] = Fq::add(x[2], Fq::mul(w, Fq::acc_reduce(a3)));
  }
  block_sum_fq<3>(x, red);
  if (threadIdx.x == 0) for (int k = 0; k < 3; k++) stg_fe(partials + 3 * blockIdx.x + k, x[k]);
}
// len < left: the weight table itself is the first polynomial (sumcheck.rs:262-342)
__global__ void __launch_bounds__(NF_THREADS) k_pow_cubic_small(const fe *pl, const fe *A, const fe *B, const fe *C, u64 len, fe *partials) {
  __shared__ fe red[3 * 32];
  Fq::acc a0 = Fq::acc_zero(), a2 = Fq::acc_zero(), a3 = Fq::acc_zero();
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (u64)gridDim.x * blockDim.x) {
    const fe tl = ldg_fe_ro(pl + i), th = ldg_fe_ro(pl + i + len);
    const fe al = ldg_fe(A + i), ah = ldg_fe(A + i + len), bl = ldg_fe(B + i), bh = ldg_fe(B + i + len), cl = ldg_fe(C + i), ch = ldg_fe(C + i + len);
    Fq::mul_acc(a0, tl, Fq::sub(Fq::mul(al, bl), cl));
    const fe dt = Fq::sub(th, tl), da = Fq::sub(ah, al), db = Fq::sub(bh, bl), dc = Fq::sub(ch, cl);
    fe tb = Fq::add(th, dt), ab = Fq::add(ah, da), bb = Fq::add(bh, db), cb = Fq::add(ch, dc);
    Fq::mul_acc(a2, tb, Fq::sub(Fq::mul(ab, bb), cb));
    tb = Fq::add(tb, dt); ab = Fq::add(ab, da); bb = Fq::add(bb, db); cb = Fq::add(cb, dc);
    Fq::mul_acc(a3, tb, Fq::sub(Fq::mul(ab, bb), cb));
  }
  fe x[3] = {Fq::acc_reduce(a0), Fq::acc_reduce(a2), Fq::acc_reduce(a3)};
  block_sum_fq<3>(x, red);
  if (threadIdx.x == 0) for (int k = 0; k < 3; k++) stg_fe(partials + 3 * blockIdx.x + k, x[k]);
}

// compute_eval_points_quad: (sum a_lo b_lo, sum (a_hi - a_lo)(b_hi - b_lo))
__global__ void __launch_bounds__(NF_THREADS) k_quad_eval(const fe *A, const fe *B, u64 len, fe *partials) {
  __shared__ fe red[2 * 32];
  Fq::acc a0 = Fq::acc_zero(), ai = Fq::acc_zero();
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (u64)gridDim.x * blockDim.x) {
    const fe al = ldg_fe(A + i), bl = ldg_fe(B + i);
    Fq::mul_acc(a0, al, bl);
    Fq::mul_acc(ai, Fq::sub(ldg_fe(A + i + len), al), Fq::sub(ldg_fe(B + i + len), bl));
  }
  fe x[2] = {Fq::acc_reduce(a0), Fq::acc_reduce(ai)};
  block_sum_fq<2>(x, red);
  if (threadIdx.x == 0) { stg_fe(partials + 2 * blockIdx.x, x[0]); stg_fe(partials + 2 * blockIdx.x + 1, x[1]); }
}

struct TablePtrs { fe *t[8]; };
// bind_poly_var_top on several tables with one challenge, in place (thread reads i and i+n, writes i)
__global__ void __launch_bounds__(NF_THREADS) k_bind_tables(TablePtrs tp, u32 ntab, u64 n, const fe *r) {
  const fe rr = ldg_fe_ro(r);
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < (u64)ntab * n; q += (u64)gridDim.x * blockDim.x) {
    const u32 s = (u32)(q / n); const u64 i = q - (u64)s * n;
    fe *T = tp.t[s];
    stg_fe(T + i, bind_pair(ldg_fe(T + i), ldg_fe(T + i + n), rr));
  }
}

// w_i * P for arbitrary (non-key) points: variable-base double-and-add, one thread per term
__global__ void __launch_bounds__(64) k_scalar_mul_var(const aff *pts, const fe *w, u32 n, u32 rows, jac *out) {
  const u32 q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n * rows) return;
  const u32 i = q / rows;
  const aff p = {ldg_fe(&pts[q].x), ldg_fe(&pts[q].y)};
  const fe k = Fq::from_mont(ldg_fe(w + i));
  jac acc = jac_inf();
  for (int b = 255; b >= 0; b--) {
    acc = jac_dbl(acc);
    if ((k.v[b >> 5] >> (b & 31)) & 1u) acc = jac_add_mixed(acc, p);
  }
  out[q].x = acc.x; out[q].y = acc.y; out[q].z = acc.z;
}
// out[row] = sum_i terms[i*rows + row]
__global__ void __launch_bounds__(128) k_point_col_sum(const jac *terms, u32 n, u32 rows, jac *out) {
  __shared__ jac red[128];
  const u32 row = blockIdx.x, tid = threadIdx.x;
  jac acc = jac_inf();
  for (u32 i = tid; i < n; i += 128) { const jac t = terms[(size_t)i * rows + row]; acc = jac_add(acc, t); }
  red[tid] = acc;
  __syncthreads();
#pragma unroll 1
  for (u32 s = 64; s >= 1; s >>= 1) {
    if (tid < s) red[tid] = jac_add(red[tid], red[tid + s]);
    __syncthreads();
  }
  if (tid == 0) out[row] = red[0];
}

int upload_small(sp2_ctx *ctx, int slot, const uint64_t *h, size_t nfe, fe **d) {
  void *p; SP2_TRY(scratch(ctx, slot, nfe * sizeof(fe) + 64, &p));
  SP2_CUDA_OK(cudaMemcpyAsync(p, h, nfe * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  *d = (fe *)p;
  return SP2_OK;
}
template <int NV>
int finish_partials(sp2_ctx *ctx, fe *partials, u32 nparts, uint64_t *out) {
  fe *d_out = partials + (size_t)nparts * NV;
  k_reduce_partials<NV><<<1, NF_THREADS, 0, ctx->stream>>>(partials, nparts, d_out);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(out, d_out, NV * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

}  // namespace

extern "C" {

/* PowPolynomial::split_evals(t, ell, left, right) -> left + right scalars (host out) */
int32_t sp2_pow_split_evals(sp2_ctx *ctx, const uint64_t *t, uint32_t left, uint32_t right, uint64_t *out) {
  cudaSetDevice(ctx->device);
  if (!left || !right) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "split_evals: empty side");
  fe *dt; SP2_TRY(upload_small(ctx, 0, t, 1, &dt));
  void *d; SP2_TRY(scratch(ctx, 1, (size_t)(left + right) * sizeof(fe), &d));
  k_pow_split<<<(left + right + 127) / 128, 128, 0, ctx->stream>>>(dt, left, right, (fe *)d);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(out, d, (size_t)(left + right) * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

/* One NIFS round over m live layers (layer q at slot q*stride, N entries each): out2 = (e0, quad_coeff).
 * E: left + right split evals; rhos: ell_b scalars (host).  Round 0 returns e0 = 0 (neutronnova_zk.rs:116). */
int32_t sp2_nifs_round_dev(sp2_ctx *ctx, uint32_t t, const uint64_t *rhos, uint32_t ell_b, uint32_t left, uint32_t right, const void *dE,
                           const void *dA, const void *dB, const void *dC, uint64_t N, uint64_t m, uint64_t stride, uint64_t *out2) {
  cudaSetDevice(ctx->device);
  if ((uint64_t)left * right != N || m < 2 || (m & 1)) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "nifs_round: bad shape");
  fe *drho; SP2_TRY(upload_small(ctx, 0, rhos, ell_b ? ell_b : 1, &drho));
  const u32 pairs = (u32)(m / 2);
  u32 chunks = std::max<u32>(1, std::min<u32>(right, (u32)(ctx->num_sms * 4) / std::max<u32>(1, pairs)));
  void *part; SP2_TRY(scratch(ctx, 1, ((size_t)chunks * pairs + 1) * 2 * sizeof(fe), &part));
  const u32 threads = std::min<u32>(NF_THREADS, (left + 31) / 32 * 32);
  k_nifs_round<<<dim3(chunks, pairs), threads, 0, ctx->stream>>>(t, drho, ell_b, left, right, (const fe *)dE, (const fe *)dA, (const fe *)dB,
                                                                  (const fe *)dC, N, stride, (fe *)part);
  SP2_LAUNCH_CHECK();
  return finish_partials<2>(ctx, (fe *)part, chunks * pairs, out2);
}

/* fold_abc_pair for every pair: slot 2p*stride <- lo + r_b (hi - lo); any of dA/dB/dC may be NULL */
int32_t sp2_nifs_fold_dev(sp2_ctx *ctx, void *dA, void *dB, void *dC, uint64_t N, uint64_t m, uint64_t stride, const uint64_t *r_b) {
  cudaSetDevice(ctx->device);
  if (m < 2 || (m & 1)) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "nifs_fold: need an even number of layers");
  fe *dr; SP2_TRY(upload_small(ctx, 0, r_b, 1, &dr));
  const u64 work = (m / 2) * N;
  unsigned nb = (unsigned)std::min<u64>((work + NF_THREADS - 1) / NF_THREADS, (u64)ctx->num_sms * 8);
  k_nifs_fold<<<nb, NF_THREADS, 0, ctx->stream>>>((fe *)dA, (fe *)dB, (fe *)dC, N, m / 2, stride, dr);
  SP2_LAUNCH_CHECK();
  return SP2_OK;
}

/* weights_from_r(r_bs, n): n scalars, host out */
int32_t sp2_weights_from_r(sp2_ctx *ctx, const uint64_t *r_bs, uint32_t ell, uint32_t n, uint64_t *out) {
  cudaSetDevice(ctx->device);
  fe *dr; SP2_TRY(upload_small(ctx, 0, r_bs, ell ? ell : 1, &dr));
  void *d; SP2_TRY(scratch(ctx, 1, (size_t)n * sizeof(fe) + 64, &d));
  k_weights_from_r<<<(n + 127) / 128, 128, 0, ctx->stream>>>(dr, ell, n, (fe *)d);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(out, d, (size_t)n * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

/* R1CSWitness::fold_multiple (W part): d_out[j] = sum_i w[i] * dWs[i*dim + j] */
int32_t sp2_fold_vectors_dev(sp2_ctx *ctx, const void *dWs, uint64_t n, uint64_t dim, const uint64_t *w, void *d_out) {
  cudaSetDevice(ctx->device);
  if (!n) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "fold_multiple: empty witness list");
  fe *dw; SP2_TRY(upload_small(ctx, 0, w, n, &dw));
  unsigned nb = (unsigned)std::min<u64>((dim + NF_THREADS - 1) / NF_THREADS, (u64)ctx->num_sms * 8);
  k_fold_vectors<<<nb, NF_THREADS, 0, ctx->stream>>>((const fe *)dWs, n, dim, dw, (fe *)d_out);
  SP2_LAUNCH_CHECK();
  return SP2_OK;
}

/* evaluation points (0, 2, 3) of the pow-weighted cubic, unscaled by base_tau (sumcheck.rs:366-498 / :262-342) */
int32_t sp2_sc_pow_cubic_eval_dev(sp2_ctx *ctx, const void *d_pow_left, uint32_t left, const void *d_pow_right, const void *dA, const void *dB,
                                  const void *dC, uint64_t table_len, uint64_t *out3) {
  cudaSetDevice(ctx->device);
  if (table_len < 2 || (table_len & (table_len - 1))) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "pow_cubic_eval: bad table length");
  const u64 len = table_len / 2;
  void *part;
  if (len >= left) {
    const u64 right = len / left;
    const u32 chunks = (u32)std::max<u64>(1, std::min<u64>(right, (u64)ctx->num_sms * 2));
    SP2_TRY(scratch(ctx, 1, ((size_t)chunks + 1) * 3 * sizeof(fe), &part));
    const u32 threads = std::min<u32>(NF_THREADS, (left + 31) / 32 * 32);
    k_pow_cubic_outer<<<chunks, threads, 0, ctx->stream>>>((const fe *)d_pow_left, left, (const fe *)d_pow_right, (const fe *)dA, (const fe *)dB,
                                                           (const fe *)dC, len, (fe *)part);
    SP2_LAUNCH_CHECK();
    return finish_partials<3>(ctx, (fe *)part, chunks, out3);
  }
  const u32 nb = (u32)((len + NF_THREADS - 1) / NF_THREADS);
  SP2_TRY(scratch(ctx, 1, ((size_t)nb + 1) * 3 * sizeof(fe), &part));
  k_pow_cubic_small<<<nb, NF_THREADS, 0, ctx->stream>>>((const fe *)d_pow_left, (const fe *)dA, (const fe *)dB, (const fe *)dC, len, (fe *)part);
  SP2_LAUNCH_CHECK();
  return finish_partials<3>(ctx, (fe *)part, nb, out3);
}

/* compute_eval_points_quad: out2 = (eval_point_0, bound_coeff) */
int32_t sp2_sc_quad_eval_dev(sp2_ctx *ctx, const void *dA, const void *dB, uint64_t table_len, uint64_t *out2) {
  cudaSetDevice(ctx->device);
  if (table_len < 2 || (table_len & (table_len - 1))) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "quad_eval: bad table length");
  const u64 len = table_len / 2;
  const u32 nb = (u32)std::min<u64>((len + NF_THREADS - 1) / NF_THREADS, (u64)ctx->num_sms * 4);
  void *part; SP2_TRY(scratch(ctx, 1, ((size_t)nb + 1) * 2 * sizeof(fe), &part));
  k_quad_eval<<<nb, NF_THREADS, 0, ctx->stream>>>((const fe *)dA, (const fe *)dB, len, (fe *)part);
  SP2_LAUNCH_CHECK();
  return finish_partials<2>(ctx, (fe *)part, nb, out2);
}

/* bind_poly_var_top on up to 8 device tables of table_len entries with one challenge, in place */
int32_t sp2_bind_tables_dev(sp2_ctx *ctx, void *const *d_tables, uint32_t ntables, uint64_t table_len, const uint64_t *r) {
  cudaSetDevice(ctx->device);
  if (ntables == 0 || ntables > 8 || table_len < 2) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "bind_tables: 1..8 tables of length >= 2");
  fe *dr; SP2_TRY(upload_small(ctx, 0, r, 1, &dr));
  TablePtrs tp; for (u32 i = 0; i < 8; i++) tp.t[i] = i < ntables ? (fe *)d_tables[i] : nullptr;
  const u64 work = (u64)ntables * (table_len / 2);
  unsigned nb = (unsigned)std::min<u64>((work + NF_THREADS - 1) / NF_THREADS, (u64)ctx->num_sms * 8);
  k_bind_tables<<<nb, NF_THREADS, 0, ctx->stream>>>(tp, ntables, table_len / 2, dr);
  SP2_LAUNCH_CHECK();
  return SP2_OK;
}

/* HyraxPCS::fold_commitments as group elements (hyrax_pc.rs:737-793): out[row] = sum_i w[i] * comms[i*rows + row] (affine in/out,
 * host).  The reference batch-normalises and calls msm_shared_weights (msm.rs:228-356): here the same through the device's
 * shared-weight Pippenger (msm_var.cu), one (row, window) CTA each — the digits of the n weights are taken once per window. */
int32_t sp2_fold_commitments(sp2_ctx *ctx, const uint64_t *comms_xy, uint32_t n, uint32_t rows, const uint64_t *w, uint64_t *out_xy) {
  cudaSetDevice(ctx->device);
  if (!n || !rows) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "fold_commitments: empty input");
  std::vector<uint64_t> byrow((size_t)n * rows * 8);              // [instance][row] -> [row][instance]
  for (uint32_t i = 0; i < n; i++) for (uint32_t r = 0; r < rows; r++) memcpy(&byrow[((size_t)r * n + i) * 8], comms_xy + ((size_t)i * rows + r) * 8, 64);
  return sp2_msm_shared_weights(ctx, w, n, byrow.data(), rows, out_xy);
}

/* HyraxPCS::fold_blinds (hyrax_pc.rs:795-819): out[row] = sum_k w[k] * blinds[k*rows + row] */
int32_t sp2_fold_blinds(sp2_ctx *ctx, const uint64_t *blinds, uint32_t n, uint32_t rows, const uint64_t *w, uint64_t *out) {
  cudaSetDevice(ctx->device);
  if (!n || !rows) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "fold_blinds: blinds and weights must be non-empty and same length");
  void *db, *dout; fe *dw;
  SP2_TRY(scratch(ctx, 2, (size_t)n * rows * sizeof(fe) + 32, &db)); SP2_TRY(scratch(ctx, 3, (size_t)rows * sizeof(fe) + 32, &dout));
  SP2_TRY(upload_small(ctx, 0, w, n, &dw));
  SP2_CUDA_OK(cudaMemcpyAsync(db, blinds, (size_t)n * rows * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  k_fold_vectors<<<(rows + NF_THREADS - 1) / NF_THREADS, NF_THREADS, 0, ctx->stream>>>((const fe *)db, n, rows, dw, (fe *)dout);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(out, dout, (size_t)rows * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

/* HyraxPCS::fold_commitments_partial (hyrax_pc.rs:821-874): the first num_data_rows rows are folded as group elements, the rest
 * rows are folded_blind[row] * h (each instance's rest row is blind * h) — h from the key's fixed-base table */
int32_t sp2_fold_commitments_partial(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *comms_xy, uint32_t n, uint32_t rows, const uint64_t *w,
                                     uint32_t num_data_rows, const uint64_t *folded_blind, uint64_t *out_xy) {
  cudaSetDevice(ctx->device);
  if (!n || !rows) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "fold_commitments_partial: Commitments and weights must have the same length");
  if (num_data_rows > rows) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "fold_commitments_partial: num_data_rows exceeds total_rows");
  if (num_data_rows >= rows) return sp2_fold_commitments(ctx, comms_xy, n, rows, w, out_xy);
  if (num_data_rows) {
    std::vector<uint64_t> byrow((size_t)n * num_data_rows * 8);
    for (uint32_t i = 0; i < n; i++) for (uint32_t r = 0; r < num_data_rows; r++) memcpy(&byrow[((size_t)r * n + i) * 8], comms_xy + ((size_t)i * rows + r) * 8, 64);
    SP2_TRY(sp2_msm_shared_weights(ctx, w, n, byrow.data(), num_data_rows, out_xy));
  }
  const uint32_t rest = rows - num_data_rows;
  void *db, *dout;
  SP2_TRY(scratch(ctx, 2, (size_t)rest * sizeof(fe) + 32, &db)); SP2_TRY(scratch(ctx, 3, (size_t)rest * sizeof(jac) + 32, &dout));
  SP2_CUDA_OK(cudaMemcpyAsync(db, folded_blind + 4 * (size_t)num_data_rows, (size_t)rest * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<MsmJob> jobs(rest);
  for (uint32_t r = 0; r < rest; r++) { MsmJob &j = jobs[r]; memset(&j, 0, sizeof(j)); j.nextra = 1; j.extra_base[0] = ck->idx_h(); j.extra_scalar[0] = (const fe *)db + r; }
  SP2_TRY(msm_run(ctx, ck, jobs, (jac *)dout));
  std::vector<uint64_t> hj((size_t)rest * 12);
  SP2_CUDA_OK(cudaMemcpyAsync(hj.data(), dout, (size_t)rest * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  sp2h::batch_normalize(hj.data(), rest, out_xy + 8 * (size_t)num_data_rows);
  return SP2_OK;
}

}  // extern "C"

// =====================================================================================================================
// Fused NeutronNova hot path: HOT LOOPS A-C of NeutronNovaZkSNARK::prove (src/neutronnova_zk.rs:1609-2093) with the
// per-step products of prep_prove (:1538-1549) in front, driven from C++ inside the library.
//
// State is device-resident across the whole prove (layers, witnesses, tables); the host only does what the reference
// also does on its main thread between the parallel loops: the per-round scalar algebra (finish_round!, UniPoly
// interpolation, claim updates) and the Keccak transcript.  Per round the device publishes the <= 6 round sums into
// HOST-MAPPED pinned memory (st.global over PCIe + system fence + sequence flag) and the host spins on the flag: no
// cudaMemcpy, no stream synchronisation; the challenge travels back as a kernel ARGUMENT of the next launch (bind /
// fold), so a round costs one flag round trip instead of upload + download + two syncs.
//
// Challenges come from a plain Keccak transcript (absorb b"p" / squeeze b"c", as the non-ZK provers do,
// src/sumcheck.rs:536-548): the reference's ZK driver draws them through its in-circuit verifier (process_round,
// src/bellpepper/r1cs.rs:735-816), which is out of scope.  The data path, the host algebra per round and
// every intermediate value are the reference's; spartan2_b200/neutronnova.py is the same driver over the per-round C-ABI
// seams (and, in the tests, over the oracle), and tests/test_gpu_neutronnova.py compares the two value by value.
// =====================================================================================================================
#include <atomic>
#include <chrono>
#include <functional>
#include "hostfield.h"
#include "r1cs.cuh"
#include "sumcheck.cuh"

namespace sp2 {
int eq_table_dev(sp2_ctx *ctx, const fe *d_r, uint32_t k, fe *d_out, cudaStream_t stream = nullptr, int slot = 15);
// small_value.cu
int nifs_round0_small_enqueue(sp2_ctx *ctx, const fe *d_rhos, u32 ell_b, u32 left, u32 right, const fe *dE, const void *dA64, const void *dB64,
                              const fe *dA, const fe *dB, const void *d_positions, u64 n_large, u64 N, u64 m, fe *d_partials, fe *d_out,
                              u32 pair_offset = 0);
int small_layers_enqueue(sp2_ctx *ctx, const fe *const *d_tabs, void *const *d_i64, u32 ntab, u64 n_layers, u64 N, void *d_flags, void *d_positions,
                         u64 *d_count);
}

struct sp2_nn_prep {
  sp2_ctx *ctx = nullptr;
  const sp2_shape *S = nullptr;
  uint32_t n = 0, ell_b = 0, ell = 0, left = 0, right = 0;   // n: instances held by THIS rank; ell_b = log2 of the global count
  // multi-GPU: rank g of G owns the step instances [g*n, (g+1)*n) of n_total = G*n; the core instance is replicated
  int rank = 0, nranks = 1; uint32_t n_total = 0;
  fe *gath = nullptr;                        // nranks x 3 x N: the surviving (A, B, C) layers of all ranks
  fe *wpart = nullptr;                       // nranks x M: per-rank partial witness folds
  // gath and wpart live in ONE allocation (xbuf) that peers map through CUDA IPC, so the two bulk exchanges of a sharded
  // prove are plain stores into every peer's copy over NVLink (k_nn_scatter) instead of a host-driven collective
  fe *xbuf = nullptr; fe *peer_x[8] = {nullptr}; bool peer_opened[8] = {false}; bool peers_connected = false;
  uint64_t N = 0, M = 0, ncols = 0;
  fe *zs = nullptr, *zc = nullptr;           // n x ncols, ncols
  fe *Ws = nullptr;                          // n x M (contiguous copies of the witness sections)
  fe *L[3] = {nullptr, nullptr, nullptr};    // cached step layers Az_i, Bz_i, Cz_i: n x N, layer-major
  fe *Lc[3] = {nullptr, nullptr, nullptr};   // cached core layers
  fe *work[3] = {nullptr, nullptr, nullptr}, *workc[3] = {nullptr, nullptr, nullptr};   // folded / bound in place by prove
  fe *z_step = nullptr, *z_core = nullptr, *abc_s = nullptr, *abc_c = nullptr;          // 2M each
  fe *E = nullptr, *rx = nullptr, *small = nullptr, *partials = nullptr;
  // small-value layers (prep_prove, neutronnova_zk.rs:1551-1584): i64 copies of the cached step layers with the union of
  // the large positions zeroed, the ascending list of those positions
  void *L64[3] = {nullptr, nullptr, nullptr}, *large_pos = nullptr, *large_flags = nullptr;
  uint64_t n_large = 0; bool has_i64 = false;
  u32 *ticket = nullptr;
  // host-mapped mailbox: [0] sequence flag, results at +64 bytes
  unsigned char *h_mail = nullptr; unsigned char *d_mail = nullptr;
  unsigned char *h_stage = nullptr;          // pinned staging for small uploads / head read-backs
  u32 seq = 0;
  std::vector<void *> owned;
  // ---- commitment half (sp2_neutronnova_prep_commit / sp2_neutronnova_snark_prove) ----
  const sp2_ck *ck = nullptr;
  uint32_t rows = 0, pre_rows = 0, np = 0;   // commitment rows per instance, rows of the precommitted section, public inputs
  aff *U = nullptr;                          // (n + 1) x rows UNBLINDED row commitments (commit_without_blind), affine, identity = (0,0); index n = core
  aff *Ucore_tab = nullptr;                  // window tables [rows][33][128] of the core's unblinded rows (c_eval * U_core[row] by table lookups)
  fe *Wfold = nullptr;                       // M: copy of the folded witness (the inner sum-check binds z_step in place)
  fe *Wfin = nullptr;                        // M: W_fold + c_eval * W_core, the polynomial PCS::prove opens
  fe *blinds_dev = nullptr;                  // (n + 1) x rows blinds of this prove | rows folded blinds
  jac *pts = nullptr;                        // (n + 1) x rows + rows + 16 Jacobian scratch points
  fe *pcs = nullptr;                         // PCS scratch: LZ | L | R | d_vec | z_vec | small scalars
  std::vector<uint64_t> Xs, Xc;              // host copies of the public IO (n x np, np)
  cudaStream_t side = nullptr; cudaEvent_t ev_fold = nullptr, ev_side = nullptr;
  cudaEvent_t ev_r0a = nullptr, ev_r0b = nullptr; float round0_ms = -1.f;   // device time of the last NIFS round-0 kernel (bench.py: roofline)
};
namespace sp2 { struct NnHooks { std::function<int(const std::vector<fe> &)> after_fold; }; }

namespace {

constexpr size_t NN_MAIL_BYTES = 4096, NN_STAGE_BYTES = 1 << 16, NN_SMALL_FE = 1024, NN_MAX_PARTS = 4096;

__global__ void __launch_bounds__(NF_THREADS) k_nifs_fold_v(fe *A, fe *B, fe *C, u64 N, u64 pairs, u64 stride, fe rr) {
  fe *T[3] = {A, B, C};
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < pairs * N; q += (u64)gridDim.x * blockDim.x) {
    const u64 p = q / N, k = q - p * N;
#pragma unroll
    for (int s = 0; s < 3; s++) {
      fe *lo = T[s] + (2 * p) * stride * N + k;
      stg_fe(lo, bind_pair(ldg_fe(lo), ldg_fe(T[s] + (2 * p + 1) * stride * N + k), rr));
    }
  }
}
// group g (one CTA) sums its nparts partial NV-tuples and writes them to the host-mapped mailbox; the last group to
// finish publishes the sequence number (system-scope fence first: the host reads the sums after it sees the flag)
template <int NV>
__global__ void __launch_bounds__(NF_THREADS) k_publish(const fe *partials, u32 nparts, fe *mail_out, u32 *mail_flag, u32 seq, u32 *ticket) {
  __shared__ fe red[NV * 32];
  const fe *src = partials + (size_t)blockIdx.x * nparts * NV;
  fe x[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) x[k] = Fq::zero();
  for (u32 b = threadIdx.x; b < nparts; b += blockDim.x)
#pragma unroll
    for (int k = 0; k < NV; k++) x[k] = Fq::add(x[k], ldg_fe(src + (size_t)b * NV + k));
  block_sum_fq<NV>(x, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) stg_fe(mail_out + blockIdx.x * NV + k, x[k]);
    __threadfence_system();
    const u32 t = atomicAdd(ticket, 1u);
    if (t == gridDim.x - 1) { *ticket = 0; __threadfence_system(); *(volatile u32 *)mail_flag = seq; }
  }
}
// Multi-GPU variant of k_publish<2>: this rank's two round sums go straight into every peer's mailbox (CUDA-IPC mapped
// device memory, plain stores over NVLink + system fence + epoch flag — the exchange of sumcheck.cu: exchange_sums), the
// kernel waits for the peers' contributions, adds, and publishes the GLOBAL sums to the host: no host collective per
// NIFS round.  The wait is bounded (2 s of %globaltimer): a missing peer surfaces as a host-side timeout, not a hang.
__global__ void __launch_bounds__(NF_THREADS) k_publish_xchg(const fe *partials, u32 nparts, DevComm dc, int slot, fe *mail_out, u32 *mail_flag, u32 seq) {
  __shared__ fe red[2 * 32];
  __shared__ fe g[2];
  __shared__ int bad;
  fe x[2] = {Fq::zero(), Fq::zero()};
  for (u32 b = threadIdx.x; b < nparts; b += blockDim.x) { x[0] = Fq::add(x[0], ldg_fe(partials + (size_t)b * 2)); x[1] = Fq::add(x[1], ldg_fe(partials + (size_t)b * 2 + 1)); }
  block_sum_fq<2>(x, red);
  if (threadIdx.x == 0) { g[0] = x[0]; g[1] = x[1]; bad = 0; }
  __syncthreads();
  if (threadIdx.x < (u32)dc.n) {
    MailBox *mb = dc.peer[threadIdx.x];
    stg_fe(&mb->sums[slot][dc.rank][0], g[0]); stg_fe(&mb->sums[slot][dc.rank][1], g[1]);
    __threadfence_system();
    *(volatile u32 *)&mb->flag[slot][dc.rank] = dc.epoch;
    const MailBox *me = dc.peer[dc.rank];
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    while (*(volatile const u32 *)&me->flag[slot][threadIdx.x] != dc.epoch) {
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      if (t1 - t0 > 2000000000ull) { bad = 1; break; }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0 && !bad) {
    const MailBox *me = dc.peer[dc.rank];
    for (int k = 0; k < 2; k++) {
      fe acc = Fq::zero();
      for (int q = 0; q < dc.n; q++) {
        fe v; u64 a, b, c, d;
        asm volatile("ld.volatile.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(&me->sums[slot][q][k]));
        v.v[0] = (u32)a; v.v[1] = (u32)(a >> 32); v.v[2] = (u32)b; v.v[3] = (u32)(b >> 32);
        v.v[4] = (u32)c; v.v[5] = (u32)(c >> 32); v.v[6] = (u32)d; v.v[7] = (u32)(d >> 32);
        acc = Fq::add(acc, v);
      }
      stg_fe(mail_out + k, acc);
    }
    __threadfence_system();
    *(volatile u32 *)mail_flag = seq;
  }
}
// bulk exchange of a sharded prove: every rank stores its block straight into all ranks' exchange buffers (peer memory
// over NVLink; its own copy included) at the same offset, then a flag barrier over the comm mailboxes makes the data of
// all ranks visible everywhere (system-scope fences on both sides; bounded wait)
struct NnPeers { fe *x[8]; int n; };
__global__ void __launch_bounds__(256) k_nn_scatter(const fe *src, u64 count, NnPeers pe, u64 dst_offset) {
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (u64)gridDim.x * blockDim.x) {
    const fe v = ldg_fe(src + i);
    for (int q = 0; q < pe.n; q++) stg_fe(pe.x[q] + dst_offset + i, v);
  }
}
__global__ void k_nn_barrier(DevComm dc, int slot, u32 *err) {
  const int tid = threadIdx.x;
  if (tid < dc.n) {
    __threadfence_system();
    *(volatile u32 *)&dc.peer[tid]->flag[slot][dc.rank] = dc.epoch;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    while (*(volatile const u32 *)&dc.peer[dc.rank]->flag[slot][tid] != dc.epoch) {
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      if (t1 - t0 > 2000000000ull) { atomicExch(err, 1u); break; }   // a missing peer must not wedge the GPU: the prove returns an error
    }
    __threadfence_system();
  }
}
__global__ void k_set_one(fe *p) { if (threadIdx.x == 0) stg_fe(p, Fq::one()); }
__global__ void __launch_bounds__(256) k_copy_rows(const fe *src, u64 src_stride, fe *dst, u64 dst_stride, u64 len, u64 rows) {
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < rows * len; q += (u64)gridDim.x * blockDim.x) {
    const u64 rw = q / len, i = q - rw * len;
    stg_fe(dst + rw * dst_stride + i, ldg_fe(src + rw * src_stride + i));
  }
}

// ---- one launch per round of the batched sum-checks: [bind to the previous challenge] + evaluate both branches +
// last-CTA reduction + publication in the host-mapped mailbox ------------------------------------------------------
struct NnTables { fe *t[2][3]; };                     // [branch][A, B, C] (outer) or [branch][poly_ABC, z, -] (inner)
template <bool FUSED>
__device__ __forceinline__ void nn_ld_pair(fe *T, u64 low, u64 len, const fe &r, fe &lo, fe &hi) {
  if (FUSED) {                                         // the table still has 4*len entries: bind (low, low+2len), (low+len, low+3len) in place
    lo = bind_pair(ldg_fe(T + low), ldg_fe(T + low + 2 * len), r);
    hi = bind_pair(ldg_fe(T + low + len), ldg_fe(T + low + 3 * len), r);
    stg_fe(T + low, lo); stg_fe(T + low + len, hi);
  } else {
    lo = ldg_fe(T + low); hi = ldg_fe(T + low + len);
  }
}
// Every CTA adds the 16-bit limb halves of its NV partial sums (valid in thread 0) to u32 accumulators behind the ticket (L2 atomics:
// CTAs * 2^16 < 2^32); the last CTA of the grid (atomic ticket) rebuilds the 2*NV sums from 16 words each, publishes them in the host-mapped
// mailbox and clears the accumulators.  (The first version summed per-CTA partial tuples in the last CTA with two block reductions per branch:
// ~19 us of the ~25 us of a round — REDUX-heavy block sums of 9 values on 16 warps — measured with %globaltimer stamps, SP2_NN_STAMPS.)
#ifdef SP2_NN_STAMPS
__device__ unsigned long long g_nn_t0, g_nn_t1;     // debug: first CTA's kernel entry, last CTA's arrival
#endif
// groups = number of independent result tuples (default: one per blockIdx.y = the two branches of a batched sum-check), group = this CTA's
template <int NV>
__device__ __forceinline__ void nn_publish_last(fe (&x)[NV], fe *, fe *red, int *is_last, u32 *ticket, fe *mail_out, u32 *mail_flag, u32 seq, u32 groups, u32 group) {
  if (groups == 0) { groups = gridDim.y; group = blockIdx.y; }
  const u32 ncta = gridDim.x * gridDim.y;
  const int tid = threadIdx.x;
#ifdef SP2_NN_STAMPS
  if (tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); g_nn_t1 = t; }
#endif
  if (tid == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) red[k] = x[k];
  }
  __syncthreads();
  for (int t = tid; t < NV * 16; t += blockDim.x) {          // (a CTA may have as few as 32 threads)
    const u32 w = red[t >> 4].v[(t & 15) >> 1], half = (t & 1) ? (w >> 16) : (w & 0xffffu);
    if (half) atomicAdd(ticket + 16 + group * NV * 16 + t, half);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) *is_last = atomicAdd(ticket, 1u) == ncta - 1;
  __syncthreads();
  if (!*is_last) return;
  __threadfence();
  if (tid < (int)groups * NV) {
    u32 *a = ticket + 16 + tid * 16;
    u32 w[16];
#pragma unroll
    for (int q = 0; q < 4; q++) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4 * q]), "=r"(w[4 * q + 1]), "=r"(w[4 * q + 2]), "=r"(w[4 * q + 3]) : "l"(a + 4 * q));
    fe v; u64 carry = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { carry += (u64)w[2 * i] + ((u64)w[2 * i + 1] << 16); v.v[i] = (u32)carry; carry >>= 32; }
    u32 top = Fq::fold_top(v, (u32)carry);
    top = Fq::fold_top(v, top);
    cond_sub_p<FqParams>(v, top);
    cond_sub_p<FqParams>(v, 0);
    stg_fe(mail_out + tid, v);
#pragma unroll
    for (int q = 0; q < 16; q++) a[q] = 0;
  }
  __syncthreads();
#ifdef SP2_NN_STAMPS
  if (tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); ((unsigned long long *)(mail_out + 24))[0] = t; ((unsigned long long *)(mail_out + 24))[1] = g_nn_t0; ((unsigned long long *)(mail_out + 24))[2] = g_nn_t1; }
#endif
  if (tid == 0) { *ticket = 0; __threadfence_system(); *(volatile u32 *)mail_flag = seq; }
}
// outer: evaluation points (0, 2, 3) of sum_x pow(x) (A B - C) per branch (compute_eval_points_cubic_with_additive_term
// [_with_outer_pow], src/sumcheck.rs:262-342, 366-498); len = half the (bound) table length
template <bool FUSED>
__global__ void __launch_bounds__(NF_THREADS) k_nn_outer_round(NnTables tb, const fe *pl, u32 left, const fe *pr, u64 len, fe r, fe *partials,
                                                               u32 *ticket, fe *mail_out, u32 *mail_flag, u32 seq) {
  __shared__ fe red[3 * 32];
  __shared__ int is_last;
  fe *A = tb.t[blockIdx.y][0], *B = tb.t[blockIdx.y][1], *C = tb.t[blockIdx.y][2];
  fe x[3] = {Fq::zero(), Fq::zero(), Fq::zero()};
  if (len >= left) {
    const u64 right = len / left;
    const u64 per = (right + gridDim.x - 1) / gridDim.x, j0 = blockIdx.x * per, j1 = min(right, j0 + per);
    for (u32 i = threadIdx.x; i < left; i += blockDim.x) {
      Fq::acc a0 = Fq::acc_zero(), a2 = Fq::acc_zero(), a3 = Fq::acc_zero();
      for (u64 j = j0; j < j1; j++) {
        const u64 low = i + j * left;
        const fe tl = ldg_fe_ro(pr + j), th = ldg_fe_ro(pr + j + right);
        fe al, ah, bl, bh, cl, ch;
        nn_ld_pair<FUSED>(A, low, len, r, al, ah); nn_ld_pair<FUSED>(B, low, len, r, bl, bh); nn_ld_pair<FUSED>(C, low, len, r, cl, ch);
        Fq::mul_acc(a0, tl, Fq::sub(Fq::mul(al, bl), cl));
        const fe dt = Fq::sub(th, tl), da = Fq::sub(ah, al), db = Fq::sub(bh, bl), dc = Fq::sub(ch, cl);
        fe tb2 = Fq::add(th, dt), ab = Fq::add(ah, da), bb = Fq::add(bh, db), cb = Fq::add(ch, dc);      // 2*high - low
        Fq::mul_acc(a2, tb2, Fq::sub(Fq::mul(ab, bb), cb));
        tb2 = Fq::add(tb2, dt); ab = Fq::add(ab, da); bb = Fq::add(bb, db); cb = Fq::add(cb, dc);          // 3*high - 2*low
        Fq::mul_acc(a3, tb2, Fq::sub(Fq::mul(ab, bb), cb));
      }
      const fe w = ldg_fe_ro(pl + i);
      x[0] = Fq::add(x[0], Fq::mul(w, Fq::acc_reduce(a0)));
      x[1] = Fq::add(x[1], Fq::mul(w, Fq::acc_reduce(a2)));
      x[2] = Fq::add(x[2], Fq::mul(w, Fq::acc_reduce(a3)));
    }
  } else {                                            // len < left: the left table itself is the weight polynomial
    Fq::acc a0 = Fq::acc_zero(), a2 = Fq::acc_zero(), a3 = Fq::acc_zero();
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (u64)gridDim.x * blockDim.x) {
      const fe tl = ldg_fe_ro(pl + i), th = ldg_fe_ro(pl + i + len);
      fe al, ah, bl, bh, cl, ch;
      nn_ld_pair<FUSED>(A, i, len, r, al, ah); nn_ld_pair<FUSED>(B, i, len, r, bl, bh); nn_ld_pair<FUSED>(C, i, len, r, cl, ch);
      Fq::mul_acc(a0, tl, Fq::sub(Fq::mul(al, bl), cl));
      const fe dt = Fq::sub(th, tl), da = Fq::sub(ah, al), db = Fq::sub(bh, bl), dc = Fq::sub(ch, cl);
      fe tb2 = Fq::add(th, dt), ab = Fq::add(ah, da), bb = Fq::add(bh, db), cb = Fq::add(ch, dc);
      Fq::mul_acc(a2, tb2, Fq::sub(Fq::mul(ab, bb), cb));
      tb2 = Fq::add(tb2, dt); ab = Fq::add(ab, da); bb = Fq::add(bb, db); cb = Fq::add(cb, dc);
      Fq::mul_acc(a3, tb2, Fq::sub(Fq::mul(ab, bb), cb));
    }
    x[0] = Fq::acc_reduce(a0); x[1] = Fq::acc_reduce(a2); x[2] = Fq::acc_reduce(a3);
  }
  block_sum_fq<3>(x, red);
  __syncthreads();
  nn_publish_last<3>(x, partials, red, &is_last, ticket, mail_out, mail_flag, seq);
}
// inner: (sum a_lo b_lo, sum (a_hi - a_lo)(b_hi - b_lo)) per branch (compute_eval_points_quad, src/sumcheck.rs:128-174)
template <bool FUSED>
__global__ void __launch_bounds__(NF_THREADS) k_nn_inner_round(NnTables tb, u64 len, fe r, fe *partials, u32 *ticket, fe *mail_out, u32 *mail_flag, u32 seq) {
  __shared__ fe red[2 * 32];
  __shared__ int is_last;
  fe *A = tb.t[blockIdx.y][0], *B = tb.t[blockIdx.y][1];
  Fq::acc a0 = Fq::acc_zero(), ai = Fq::acc_zero();
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (u64)gridDim.x * blockDim.x) {
    fe al, ah, bl, bh;
    nn_ld_pair<FUSED>(A, i, len, r, al, ah); nn_ld_pair<FUSED>(B, i, len, r, bl, bh);
    Fq::mul_acc(a0, al, bl);
    Fq::mul_acc(ai, Fq::sub(ah, al), Fq::sub(bh, bl));
  }
  fe x[2] = {Fq::acc_reduce(a0), Fq::acc_reduce(ai)};
  block_sum_fq<2>(x, red);
  __syncthreads();
  nn_publish_last<2>(x, partials, red, &is_last, ticket, mail_out, mail_flag, seq);
}
// last bind of a sum-check: the tables have two entries left; out[s] = T_s[0] + r (T_s[1] - T_s[0]) -> mailbox
__global__ void k_bind_heads(TablePtrs tp, u32 ntab, fe r, fe *mail_out, u32 *mail_flag, u32 seq) {
  const u32 i = threadIdx.x;
  if (i < ntab) stg_fe(mail_out + i, bind_pair(ldg_fe(tp.t[i]), ldg_fe(tp.t[i] + 1), r));
  __syncthreads();
  if (i == 0) { __threadfence_system(); *(volatile u32 *)mail_flag = seq; }
}

// ---- pipelined rounds of the batched sum-checks -----------------------------------------------------------------------------
// A host-driven round costs launch + kernel + mailbox + host algebra + Keccak, all serial.  Binding is linear in the challenge, so
// the sums of round i+1 over the table bound to r_i are quadratics in r_i whose coefficients need only the table bound to r_(i-1)
// (sumcheck.cu: pipelined tails): kernel i binds to r_(i-1) and publishes the COEFFICIENT sums of round i+1 — Karatsuba triples
// (value at r = 0, value at r = 1, the r^2 coefficient) — while the host is already hashing round i.  The host evaluates a round with two
// multiplications per sum, is always one launch ahead of the device, and the kernels run back to back on the stream: a round costs
// one kernel, not launch + kernel + round trip + host.  Same sums, same field elements: every proof field stays bit-identical.
//
// entries of the table (after the fused bind to r) that pair index k of the NEXT round touches: k, k + h in the low half, k + len, k + len + h
// in the high half (len = pairs of this round, h = len / 2)
template <bool FUSED>
__device__ __forceinline__ void nn_ld_quad(fe *T, u64 k, u64 h, u64 len, const fe &r, fe (&e)[4]) {
  const u64 pos[4] = {k, k + h, k + len, k + len + h};
#pragma unroll
  for (int q = 0; q < 4; q++) {
    if (FUSED) { e[q] = bind_pair(ldg_fe(T + pos[q]), ldg_fe(T + pos[q] + 2 * len), r); stg_fe(T + pos[q], e[q]); }
    else e[q] = ldg_fe(T + pos[q]);
  }
}
// Both coefficient kernels split a pair's work over the threads of a CTA through shared memory, so that a thread runs 2-3 dependent
// multiplications instead of 32 (these kernels are latency-bound: ~5000 dependent instructions per thread at one instruction per ~8 cycles
// was the whole 20 us of a round): phase 1 — one task per (table, entry, pair): bind (FUSED) or load one of the 4 entries per table of NN_KC
// pairs, and the pow weights; phase 2 — one task per (coefficient, pair): two multiplications; warp c sums coefficient c.
constexpr int NN_KC = 32;                            // next-round pairs per CTA and pass (= one warp per coefficient in phase 2)
constexpr int NN_CT = 512;                           // threads of a coefficient CTA
__device__ __forceinline__ u64 nn_quad_pos(int e, u64 k, u64 h, u64 len) { return k + ((e & 1) ? h : 0) + ((e & 2) ? len : 0); }
// inner: per branch the triples of e0' = sum a_lo b_lo and t_inf' = sum (a_hi - a_lo)(b_hi - b_lo) of the next round
template <bool FUSED>
__global__ void __launch_bounds__(NN_CT) k_nn_inner_coef(NnTables tb, u64 len, fe r, fe *partials, u32 *ticket, fe *mail_out, u32 *mail_flag, u32 seq) {
  __shared__ fe E[2][4][NN_KC];                       // [table][entry][pair]
  __shared__ fe acc[6];
  __shared__ fe red[6 * 32];
  __shared__ int is_last;
  fe *T[2] = {tb.t[blockIdx.y][0], tb.t[blockIdx.y][1]};
  const u64 h = len / 2;
  const int tid = threadIdx.x;
  if (tid < 6) acc[tid] = Fq::zero();
  for (u64 k0 = (u64)blockIdx.x * NN_KC; k0 < h; k0 += (u64)gridDim.x * NN_KC) {
    if (tid < 2 * 4 * NN_KC) {
      const int tab = tid / (4 * NN_KC), e = (tid / NN_KC) % 4, kk = tid % NN_KC;
      fe v = Fq::zero();
      if (k0 + kk < h) {
        fe *p = T[tab] + nn_quad_pos(e, k0 + kk, h, len);
        if (FUSED) { v = bind_pair(ldg_fe(p), ldg_fe(p + 2 * len), r); stg_fe(p, v); } else v = ldg_fe(p);
      }
      E[tab][e][kk] = v;
    }
    __syncthreads();
    // next round: a_lo' = a0 + r' (a2 - a0), a_hi' = a1 + r' (a3 - a1); warp c takes coefficient c of the 32 pairs
    if (tid < 6 * NN_KC) {
      const int c = tid / NN_KC, kk = tid % NN_KC;
      const fe a0 = E[0][0][kk], a1 = E[0][1][kk], a2 = E[0][2][kk], a3 = E[0][3][kk], b0 = E[1][0][kk], b1 = E[1][1][kk], b2 = E[1][2][kk], b3 = E[1][3][kk];
      fe u, v;
      if (c == 0) { u = a0; v = b0; } else if (c == 1) { u = a2; v = b2; } else if (c == 2) { u = Fq::sub(a2, a0); v = Fq::sub(b2, b0); }
      else if (c == 3) { u = Fq::sub(a1, a0); v = Fq::sub(b1, b0); } else if (c == 4) { u = Fq::sub(a3, a2); v = Fq::sub(b3, b2); }
      else { u = Fq::sub(Fq::sub(a3, a2), Fq::sub(a1, a0)); v = Fq::sub(Fq::sub(b3, b2), Fq::sub(b1, b0)); }
      fe x[1] = {Fq::mul(u, v)};
      warp_sum_fq_cols<1>(x);
      if (kk == 0) acc[c] = Fq::add(acc[c], x[0]);
    }
    __syncthreads();
  }
  __syncthreads();
  fe tot[6];
#pragma unroll
  for (int q = 0; q < 6; q++) tot[q] = acc[q];
  nn_publish_last<6>(tot, partials, red, &is_last, ticket, mail_out, mail_flag, seq);
}
// outer: per branch, for each evaluation point t in {0, 2, 3} of the next round, the triple of sum_k w_t(k) (a_t b_t - c_t):
//   coefficient 3 s + t with s = 0: the entries at r' = 0 (low half), s = 1: at r' = 1 (high half), s = 2: the differences (no c term)
template <bool FUSED>
__global__ void __launch_bounds__(NN_CT) k_nn_outer_coef(NnTables tb, const fe *pl, u32 left, const fe *pr, u64 len, fe r, fe *partials,
                                                         u32 *ticket, fe *mail_out, u32 *mail_flag, u32 seq) {
  __shared__ fe E[3][4][NN_KC];                       // [table][entry][pair]
  __shared__ fe W[2][NN_KC];                          // pow weights (low, high) of the next round's pairs
  __shared__ fe acc[9];
  __shared__ fe red[9 * 32];
  __shared__ int is_last;
  fe *T[3] = {tb.t[blockIdx.y][0], tb.t[blockIdx.y][1], tb.t[blockIdx.y][2]};
  const u64 h = len / 2;
  const int tid = threadIdx.x;
#ifdef SP2_NN_STAMPS
  if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); g_nn_t0 = t; }
#endif
  if (tid < 9) acc[tid] = Fq::zero();
  for (u64 k0 = (u64)blockIdx.x * NN_KC; k0 < h; k0 += (u64)gridDim.x * NN_KC) {
    if (tid < 3 * 4 * NN_KC) {
      const int tab = tid / (4 * NN_KC), e = (tid / NN_KC) % 4, kk = tid % NN_KC;
      fe v = Fq::zero();
      if (k0 + kk < h) {
        fe *p = T[tab] + nn_quad_pos(e, k0 + kk, h, len);
        if (FUSED) { v = bind_pair(ldg_fe(p), ldg_fe(p + 2 * len), r); stg_fe(p, v); } else v = ldg_fe(p);
      }
      E[tab][e][kk] = v;
    } else if (tid < 3 * 4 * NN_KC + 2 * NN_KC) {
      // pow weights of the next round's pair k (low, high); len' = h (PowPolynomial::split_evals, power.rs:65-86)
      const int hi = (tid - 3 * 4 * NN_KC) / NN_KC, kk = tid % NN_KC;
      const u64 k = k0 + kk;
      fe w = Fq::zero();
      if (k < h) {
        if (h >= left) w = Fq::mul(ldg_fe_ro(pl + k % left), ldg_fe_ro(pr + k / left + (hi ? h / left : 0)));
        else w = ldg_fe_ro(pl + k + (hi ? h : 0));
      }
      W[hi][kk] = w;
    }
    __syncthreads();
    if (tid < 9 * NN_KC) {                            // warp c: coefficient c = 3 s + t of the 32 pairs
      const int c = tid / NN_KC, kk = tid % NN_KC, sidx = c / 3, t = c % 3;
      fe al, ah, bl, bh, cl, ch;
      if (sidx < 2) { al = E[0][2 * sidx][kk]; ah = E[0][2 * sidx + 1][kk]; bl = E[1][2 * sidx][kk]; bh = E[1][2 * sidx + 1][kk]; cl = E[2][2 * sidx][kk]; ch = E[2][2 * sidx + 1][kk]; }
      else {
        al = Fq::sub(E[0][2][kk], E[0][0][kk]); ah = Fq::sub(E[0][3][kk], E[0][1][kk]);
        bl = Fq::sub(E[1][2][kk], E[1][0][kk]); bh = Fq::sub(E[1][3][kk], E[1][1][kk]); cl = Fq::zero(); ch = Fq::zero();
      }
      // the entries and the weight at the evaluation point (0, 2, 3): low + t (high - low)
      const fe tl = W[0][kk], th = W[1][kk];
      fe xa = al, xb = bl, xc = cl, xw = tl;
      if (t > 0) {
        const fe da = Fq::sub(ah, al), db = Fq::sub(bh, bl), dc = Fq::sub(ch, cl), dw = Fq::sub(th, tl);
        xa = Fq::add(ah, da); xb = Fq::add(bh, db); xc = Fq::add(ch, dc); xw = Fq::add(th, dw);
        if (t > 1) { xa = Fq::add(xa, da); xb = Fq::add(xb, db); xc = Fq::add(xc, dc); xw = Fq::add(xw, dw); }
      }
      fe x[1] = {Fq::mul(xw, Fq::sub(Fq::mul(xa, xb), xc))};
      warp_sum_fq_cols<1>(x);
      if (kk == 0) acc[c] = Fq::add(acc[c], x[0]);
    }
    __syncthreads();
  }
  __syncthreads();
  fe tot[9];
#pragma unroll
  for (int q = 0; q < 9; q++) tot[q] = acc[q];
  nn_publish_last<9>(tot, partials, red, &is_last, ticket, mail_out, mail_flag, seq);
}
// bind several tables to a challenge passed by value (the last bind of a pipelined sum-check)
__global__ void __launch_bounds__(NF_THREADS) k_bind_tables_v(TablePtrs tp, u32 ntab, u64 n, fe r) {
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < (u64)ntab * n; q += (u64)gridDim.x * blockDim.x) {
    const u32 s = (u32)(q / n); const u64 i = q - (u64)s * n;
    fe *T = tp.t[s];
    stg_fe(T + i, bind_pair(ldg_fe(T + i), ldg_fe(T + i + n), r));
  }
}
// mailbox slots of the pipelined loops (the host runs one launch ahead: consecutive kernels must not share a slot)
constexpr int NN_SLOTS = 3;

// wait for the device to publish sequence number `seq` in a mailbox flag (bounded: a wedged stream surfaces as an error)
int nn_wait_flag(sp2_nn_prep *P, volatile u32 *flag, u32 seq) {
  sp2_ctx *ctx = P->ctx;
  const auto t0 = std::chrono::steady_clock::now();
  uint64_t spins = 0;
  while (*flag != seq) {
    if ((++spins & 0xfffff) == 0) {
      const cudaError_t e = cudaStreamQuery(ctx->stream);
      if (e != cudaSuccess && e != cudaErrorNotReady) return set_cuda_error(ctx, e, "neutronnova round", __LINE__);
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 20.0)
        return set_error(ctx, SP2_ERR_INTERNAL, "neutronnova: device did not publish a round result within 20 s");
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return SP2_OK;
}
int nn_wait(sp2_nn_prep *P, u32 seq) { return nn_wait_flag(P, (volatile u32 *)P->h_mail, seq); }
// slot s of the pipelined loops: flag at 16 (s + 1), <= 30 scalars at 1024 (s + 1) of the 4 KB mailbox
inline int nn_wait_slot(sp2_nn_prep *P, int s, u32 seq) { return nn_wait_flag(P, (volatile u32 *)(P->h_mail + 16 * (s + 1)), seq); }
inline const fe *nn_slot(const sp2_nn_prep *P, int s) { return (const fe *)(P->h_mail + 1024 * (s + 1)); }
inline fe *nn_slot_dev(const sp2_nn_prep *P, int s) { return (fe *)(P->d_mail + 1024 * (s + 1)); }
inline u32 *nn_slot_flag_dev(const sp2_nn_prep *P, int s) { return (u32 *)(P->d_mail + 16 * (s + 1)); }
// SP2_NN_PIPE=0 (read per prove) restores the one-launch-per-round loops (bind + direct sums, host waits for every kernel).  Measured on
// B200, config 3: 2.18 vs 2.27 ms per prove (outer 0.37 vs 0.51, inner 0.40 vs 0.45 ms).  What a pipelined round costs now: ~6 us between
// consecutive kernels, ~4.3 us of kernel once the tables are small (55 / 30 / 17 / 11 us for the first four: 32 instead of 12 multiplications per
// pair), 1.6 us publish, against ~9 us of host algebra + Keccak + launch per round — host and device are balanced; the device-resident
// transcript of sumcheck.cu is what would remove both.
static bool nn_pipe() { const char *e = getenv("SP2_NN_PIPE"); return !(e && e[0] == '0'); }
inline const fe *nn_mail(const sp2_nn_prep *P) { return (const fe *)(P->h_mail + 64); }
inline fe *nn_mail_dev(const sp2_nn_prep *P) { return (fe *)(P->d_mail + 64); }

int nn_alloc(sp2_nn_prep *P, size_t nfe, fe **out) {
  sp2_ctx *ctx = P->ctx; void *p;
  SP2_CUDA_OK(cudaMalloc(&p, std::max<size_t>(nfe, 1) * sizeof(fe)));
  P->owned.push_back(p); *out = (fe *)p;
  return SP2_OK;
}

}  // namespace

extern "C" {

void sp2_neutronnova_prep_free(sp2_nn_prep *P) {
  if (!P) return;
  cudaSetDevice(P->ctx->device);
  cudaStreamSynchronize(P->ctx->stream);
  for (int q = 0; q < 8; q++) if (P->peer_opened[q]) cudaIpcCloseMemHandle(P->peer_x[q]);
  for (void *p : P->owned) cudaFree(p);
  if (P->Ucore_tab) cudaFree(P->Ucore_tab);
  if (P->side) cudaStreamDestroy(P->side);
  if (P->ev_fold) cudaEventDestroy(P->ev_fold);
  if (P->ev_side) cudaEventDestroy(P->ev_side);
  if (P->ev_r0a) cudaEventDestroy(P->ev_r0a);
  if (P->ev_r0b) cudaEventDestroy(P->ev_r0b);
  if (P->h_mail) cudaFreeHost(P->h_mail);
  if (P->h_stage) cudaFreeHost(P->h_stage);
  delete P;
}

/* NeutronNovaZkSNARK::prep_prove, data path (src/neutronnova_zk.rs:1477-1603): the n step instances z_i = [W_i | 1 | X_i]
 * (num_cols scalars each, host) and the core instance are uploaded once; Az_i, Bz_i, Cz_i for every step and for the
 * core circuit are computed on the device and cached, layer-major.  n must be a power of two >= 2. */
static int32_t nn_prep_impl(sp2_ctx *ctx, const sp2_shape *S, int rank, int nranks, uint32_t n_steps, const uint64_t *step_zs, const uint64_t *core_z,
                            sp2_nn_prep **out) {
  cudaSetDevice(ctx->device);
  if (!out) return SP2_ERR_INTERNAL;
  *out = nullptr;
  if (nranks < 1 || (nranks & (nranks - 1)) || rank < 0 || rank >= nranks) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "neutronnova: ranks must be a power of two");
  const uint64_t n_total = (uint64_t)n_steps * nranks;
  if (n_steps < 1 || (n_steps & (n_steps - 1)) || n_total < 2) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "neutronnova: the number of step instances must be a power of two >= 2");
  if (n_total > 256) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "neutronnova: at most 256 step instances");
  if (S->nranks != 1) return set_error(ctx, SP2_ERR_UNSUPPORTED, "neutronnova: whole shape expected");
  sp2_nn_prep *P = new sp2_nn_prep();
  P->ctx = ctx; P->S = S; P->n = n_steps; P->N = S->num_cons; P->M = S->num_vars; P->ncols = S->num_cols;
  P->rank = rank; P->nranks = nranks; P->n_total = (uint32_t)n_total;
  P->np = (uint32_t)S->num_public;
  if (P->np) {                                      // public IO of the instances (folded on the host: X_acc = sum_i w_i X_i)
    if (nranks > 1) { delete P; return set_error(ctx, SP2_ERR_UNSUPPORTED, "neutronnova: instance sharding needs step circuits without public IO"); }
    P->Xs.resize((size_t)n_steps * P->np * 4); P->Xc.resize((size_t)P->np * 4);
    for (uint32_t i = 0; i < n_steps; i++) memcpy(&P->Xs[(size_t)i * P->np * 4], step_zs + ((size_t)i * S->num_cols + S->num_vars + 1) * 4, (size_t)P->np * 32);
    memcpy(P->Xc.data(), core_z + (S->num_vars + 1) * 4, (size_t)P->np * 32);
  }
  while ((1u << P->ell_b) < n_total) P->ell_b++;
  while ((1ull << P->ell) < P->N) P->ell++;
  P->left = 1u << ((P->ell + 1) / 2); P->right = 1u << (P->ell / 2);          // compute_tensor_decomp (:58-67)
  int rc = SP2_OK;
  auto fail = [&](int code) { sp2_neutronnova_prep_free(P); return code; };
  const size_t n = n_steps, N = P->N, M = P->M, nc = P->ncols;
  if ((rc = nn_alloc(P, n * nc, &P->zs)) || (rc = nn_alloc(P, nc, &P->zc)) || (rc = nn_alloc(P, n * M, &P->Ws))) return fail(rc);
  for (int k = 0; k < 3; k++)
    if ((rc = nn_alloc(P, n * N, &P->L[k])) || (rc = nn_alloc(P, N, &P->Lc[k])) || (rc = nn_alloc(P, n * N, &P->work[k])) || (rc = nn_alloc(P, N, &P->workc[k])))
      return fail(rc);
  if ((rc = nn_alloc(P, 2 * M, &P->z_step)) || (rc = nn_alloc(P, 2 * M, &P->z_core)) || (rc = nn_alloc(P, 2 * M, &P->abc_s)) || (rc = nn_alloc(P, 2 * M, &P->abc_c)) ||
      (rc = nn_alloc(P, (size_t)P->left + P->right, &P->E)) || (rc = nn_alloc(P, N, &P->rx)) || (rc = nn_alloc(P, NN_SMALL_FE, &P->small)) ||
      (rc = nn_alloc(P, NN_MAX_PARTS * 6, &P->partials)))
    return fail(rc);
  if (nranks > 1) {
    if ((rc = nn_alloc(P, (size_t)nranks * (3 * N + M), &P->xbuf))) return fail(rc);
    P->gath = P->xbuf; P->wpart = P->xbuf + (size_t)nranks * 3 * N; P->peer_x[rank] = P->xbuf;
  }
  { void *p; if (cudaMalloc(&p, 2048) != cudaSuccess) return fail(set_error(ctx, SP2_ERR_CUDA, "cudaMalloc")); P->owned.push_back(p); P->ticket = (u32 *)p;
    cudaMemsetAsync(p, 0, 2048, ctx->stream); }   // [0] ticket, [1] peer-barrier error, [16 ...] column accumulators of nn_publish_last
  if (cudaHostAlloc((void **)&P->h_mail, NN_MAIL_BYTES, cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void **)&P->d_mail, P->h_mail, 0) != cudaSuccess ||
      cudaMallocHost((void **)&P->h_stage, NN_STAGE_BYTES) != cudaSuccess)
    return fail(set_error(ctx, SP2_ERR_CUDA, "neutronnova: pinned allocation failed"));
  memset(P->h_mail, 0, NN_MAIL_BYTES);
  // upload the instances (one copy each), split off the witness sections, per-step products
  if (cudaMemcpyAsync(P->zs, step_zs, n * nc * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
      cudaMemcpyAsync(P->zc, core_z, nc * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
    return fail(set_error(ctx, SP2_ERR_CUDA, "neutronnova: upload failed"));
  k_copy_rows<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(P->zs, nc, P->Ws, M, M, n);
  ctx->launches++;
  for (size_t i = 0; i < n; i++) {
    fe *o[3] = {P->L[0] + i * N, P->L[1] + i * N, P->L[2] + i * N};
    if ((rc = spmv3_dev(ctx, S, S->M, P->zs + i * nc, nullptr, o))) return fail(rc);
  }
  { fe *o[3] = {P->Lc[0], P->Lc[1], P->Lc[2]}; if ((rc = spmv3_dev(ctx, S, S->M, P->zc, nullptr, o))) return fail(rc); }
  { // i64 layers for the small-value NIFS round 0 (SP2_NN_NO_SMALL=1: field path only)
    const char *e = getenv("SP2_NN_NO_SMALL");
    if (!(e && e[0] == '1') && n >= 2) {
      for (int k = 0; k < 3; k++) { void *p; if (cudaMalloc(&p, n * N * 8) != cudaSuccess) return fail(set_error(ctx, SP2_ERR_CUDA, "cudaMalloc")); P->owned.push_back(p); P->L64[k] = p; }
      { void *p; if (cudaMalloc(&p, N * 8 + 64) != cudaSuccess) return fail(set_error(ctx, SP2_ERR_CUDA, "cudaMalloc")); P->owned.push_back(p); P->large_pos = p; }
      { void *p; if (cudaMalloc(&p, N + 64) != cudaSuccess) return fail(set_error(ctx, SP2_ERR_CUDA, "cudaMalloc")); P->owned.push_back(p); P->large_flags = p; }
      const fe *tabs[3] = {P->L[0], P->L[1], P->L[2]};
      u64 *d_count = (u64 *)((unsigned char *)P->large_pos + N * 8);
      if ((rc = small_layers_enqueue(ctx, tabs, P->L64, 3, n, N, P->large_flags, P->large_pos, d_count))) return fail(rc);
      if (cudaMemcpyAsync(&P->n_large, d_count, 8, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return fail(set_error(ctx, SP2_ERR_CUDA, "neutronnova: prep_prove"));
      P->has_i64 = true;
    }
  }
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(set_error(ctx, SP2_ERR_CUDA, "neutronnova: prep_prove failed on the device"));
  *out = P;
  return SP2_OK;
}
int32_t sp2_neutronnova_prep_prove(sp2_ctx *ctx, const sp2_shape *S, uint32_t n_steps, const uint64_t *step_zs, const uint64_t *core_z,
                                   sp2_nn_prep **out) {
  return nn_prep_impl(ctx, S, 0, 1, n_steps, step_zs, core_z, out);
}
/* Multi-GPU (one process per GPU): rank g of G = nranks holds the n_local step instances [g * n_local, (g+1) * n_local)
 * of n_local * G in total (instances are the independent units); the core instance is replicated. */
int32_t sp2_neutronnova_prep_prove_sharded(sp2_ctx *ctx, const sp2_shape *S, int32_t rank, int32_t nranks, uint32_t n_local,
                                           const uint64_t *local_step_zs, const uint64_t *core_z, sp2_nn_prep **out) {
  return nn_prep_impl(ctx, S, rank, nranks, n_local, local_step_zs, core_z, out);
}

/* HOT LOOPS A-C (see the section header).  `ts` is advanced exactly as the Python driver advances its transcript.
 * phase_ms (optional, 6 floats): nifs, fold_witness, outer_sumcheck_batched, compute_eval_table_sparse,
 * inner_sumcheck_batched, total — host wall clock (every phase ends in a host wait). */
static int32_t nn_prove_impl(sp2_ctx *ctx, sp2_nn_prep *P, sp2_transcript *tsh, sp2_comm *comm, sp2_allgather_fn allgather, void *user,
                             sp2_nn_proof *pf, float *phase_ms, const sp2::NnHooks *hooks = nullptr) {
  cudaSetDevice(ctx->device);
  if (!P || !tsh || !pf) return set_error(ctx, SP2_ERR_INTERNAL, "neutronnova_prove: null argument");
  if (comm && (!comm->connected || comm->dc.n != P->nranks || comm->dc.rank != P->rank)) return set_error(ctx, SP2_ERR_INTERNAL, "neutronnova_prove: comm does not match the prep state");
  const bool xchg = comm && P->nranks > 1;          // per-round sums exchanged inside the publish kernel over NVLink
  const bool peer_x = xchg && P->peers_connected;   // bulk exchanges as peer stores too (else through the allgather callback)
  if (P->nranks > 1 && !peer_x && !allgather) return set_error(ctx, SP2_ERR_INTERNAL, "neutronnova_prove: a sharded prep state needs connected peers or the all-gather callback");
  if (xchg) comm->dc.epoch++;
  sp2h::Transcript &ts = tsh->t;
  const sp2_shape *S = P->S;
  const u32 n = P->n, ell_b = P->ell_b, ell = P->ell, left = P->left, right = P->right;
  const u32 G = (u32)P->nranks, rank = (u32)P->rank, n_total = P->n_total;
  u32 ell_local = 0; while ((1u << ell_local) < n) ell_local++;          // NIFS rounds that are local to a rank
  const u64 N = P->N, M = P->M;
  u32 my = 0; while ((1ull << my) < 2 * M) my++;                       // inner rounds: log2(2M)
  if (N > (1ull << 31) || n_total > 256 || ell_b > 8 || my > 40 || ell > 40) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "neutronnova_prove: size out of range");
  pf->n_steps = n_total; pf->ell_b = ell_b; pf->ell = ell; pf->rounds_y = my; pf->outer_ok = 0; pf->inner_ok = 0;
  typedef NnHost H;
  const fe one = HF::one(), zero = HF::zero();
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [](std::chrono::steady_clock::time_point a) { return (float)(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count()); };
  const auto t_begin = now(); auto t_phase = t_begin;
  float ph[6] = {0, 0, 0, 0, 0, 0};
  auto squeeze = [&](const char *label) { uint8_t dg[64]; uint64_t o[4]; ts.squeeze(label, dg); sp2h::fq_from_uniform(dg, o); return H::load(o); };
  auto absorb = [&](const char *label, const fe *v, size_t k) { ts.absorb_scalars(label, (const uint64_t *)v, k); };
  size_t stage_off = 0;
  auto stage = [&](const void *src, size_t bytes, fe *dst) -> int {   // small async upload through pinned staging
    if (stage_off + bytes > NN_STAGE_BYTES) return set_error(ctx, SP2_ERR_INTERNAL, "neutronnova: staging overflow");
    memcpy(P->h_stage + stage_off, src, bytes);
    SP2_CUDA_OK(cudaMemcpyAsync(dst, P->h_stage + stage_off, bytes, cudaMemcpyHostToDevice, ctx->stream));
    stage_off += (bytes + 63) & ~(size_t)63;
    return SP2_OK;
  };
  u32 *mail_flag = (u32 *)P->d_mail;

  // working copies of the cached layers (the NIFS folds and the sum-check binds are in place)
  for (int k = 0; k < 3; k++) {
    SP2_CUDA_OK(cudaMemcpyAsync(P->work[k], P->L[k], (size_t)n * N * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
    SP2_CUDA_OK(cudaMemcpyAsync(P->workc[k], P->Lc[k], (size_t)N * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  fe *As = P->work[0], *Bs = P->work[1], *Cs = P->work[2], *Ac = P->workc[0], *Bc = P->workc[1], *Cc = P->workc[2];

  // ---- HOT LOOP A: NIFS (neutronnova_zk.rs:511-1273) ----------------------------------------------------------
  absorb("T", &zero, 1);
  const fe tau = squeeze("tau");
  SP2_TRY(stage(&tau, sizeof(fe), P->small));
  k_pow_split<<<(left + right + 127) / 128, 128, 0, ctx->stream>>>(P->small, left, right, P->E);
  SP2_LAUNCH_CHECK();
  std::vector<fe> rhos(ell_b), rho_inv(ell_b);
  for (u32 t = 0; t < ell_b; t++) rhos[t] = squeeze("rho");
  SP2_TRY(stage(rhos.data(), ell_b * sizeof(fe), P->small + 8));
  const fe *d_rhos = P->small + 8;
  { // batch inversion of the rhos (finish_round! divides by rho, :703-735): one inversion, on the host while the device copies
    std::vector<fe> pre(ell_b + 1); pre[0] = one;
    for (u32 t = 0; t < ell_b; t++) { if (HF::is_zero(rhos[t])) return set_error(ctx, SP2_ERR_DIVISION_BY_ZERO, "neutronnova: rho = 0"); pre[t + 1] = HF::mul(pre[t], rhos[t]); }
    fe inv = HF::inv(pre[ell_b]);
    for (u32 t = ell_b; t-- > 0;) { rho_inv[t] = HF::mul(inv, pre[t]); inv = HF::mul(inv, rhos[t]); }
  }
  fe T_cur = zero, acc_eq = one;
  std::vector<fe> r_bs(ell_b);
  u64 m = n, stride = 1;
  bool gathered = G == 1, timed_r0 = false;
  for (u32 t = 0; t < ell_b; t++) {
    if (!gathered && t >= ell_local) {
      // every rank is down to ONE layer triple: all-gather them (rank order = instance order) and finish the last
      // log2(G) rounds replicated.  recv layout [rank][A|B|C][N]: table k of layer q sits at gath + (3q + k) N, i.e.
      // layer-major with stride 3 from base gath + k N — exactly what the round / fold kernels take.
      fe *mine = P->gath + (size_t)rank * 3 * N;
      if (peer_x) {
        NnPeers pe; pe.n = (int)G; for (u32 q = 0; q < 8; q++) pe.x[q] = q < G ? P->peer_x[q] : nullptr;
        k_nn_barrier<<<1, 32, 0, ctx->stream>>>(comm->dc, 30, P->ticket + 1);      // every rank has entered this prove: its exchange buffer is free
        SP2_LAUNCH_CHECK();
        for (int k = 0; k < 3; k++) {
          k_nn_scatter<<<ctx->num_sms, 256, 0, ctx->stream>>>(P->work[k], N, pe, (u64)rank * 3 * N + (u64)k * N);
          SP2_LAUNCH_CHECK();
        }
        k_nn_barrier<<<1, 32, 0, ctx->stream>>>(comm->dc, 31, P->ticket + 1);
        SP2_LAUNCH_CHECK();
      } else {
        for (int k = 0; k < 3; k++) SP2_CUDA_OK(cudaMemcpyAsync(mine + (size_t)k * N, P->work[k], (size_t)N * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
        SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        if (allgather(user, mine, (uint64_t)3 * N * sizeof(fe), P->gath, 1) != 0) return set_error(ctx, SP2_ERR_INTERNAL, "neutronnova: all-gather of the surviving layers failed");
      }
      As = P->gath; Bs = P->gath + N; Cs = P->gath + 2 * N;
      m = G; stride = 3; gathered = true;
    }
    const bool local = t < ell_local && G > 1;
    const u32 pairs = (u32)(m / 2);
    const u32 pair_offset = local ? rank * pairs : 0;
    u32 chunks = std::max<u32>(1, std::min<u32>(right, (u32)(ctx->num_sms * 4) / std::max<u32>(1, pairs)));
    if ((size_t)chunks * pairs > NN_MAX_PARTS) chunks = std::max<u32>(1, (u32)(NN_MAX_PARTS / pairs));
    const u32 threads = std::min<u32>(NF_THREADS, (left + 31) / 32 * 32);
    const u32 seq = ++P->seq;
    if (t == 0 && P->has_i64 && ell_local > 0) {
      // round 0 on the i64 layers (prove_helper_small, :255-325): e0 = 0, quad from i64 differences / i128 products
      SP2_CUDA_OK(cudaMemsetAsync(P->partials, 0, sizeof(fe), ctx->stream));
      if (!P->ev_r0a) { SP2_CUDA_OK(cudaEventCreate(&P->ev_r0a)); SP2_CUDA_OK(cudaEventCreate(&P->ev_r0b)); }
      SP2_CUDA_OK(cudaEventRecord(P->ev_r0a, ctx->stream));
      SP2_TRY(nifs_round0_small_enqueue(ctx, d_rhos, ell_b, left, right, P->E, P->L64[0], P->L64[1], As, Bs, P->large_pos, P->n_large, N, m,
                                        P->partials + 64, P->partials + 1, pair_offset));
      SP2_CUDA_OK(cudaEventRecord(P->ev_r0b, ctx->stream));
      timed_r0 = true;
      if (local && xchg) k_publish_xchg<<<1, NF_THREADS, 0, ctx->stream>>>(P->partials, 1, comm->dc, (int)t + 1, nn_mail_dev(P), mail_flag, seq);
      else k_publish<2><<<1, NF_THREADS, 0, ctx->stream>>>(P->partials, 1, nn_mail_dev(P), mail_flag, seq, P->ticket);
    } else {
      if (local && xchg) {
        k_nifs_round<<<dim3(chunks, pairs), threads, 0, ctx->stream>>>(t, d_rhos, ell_b, left, right, P->E, As, Bs, Cs, N, stride, P->partials, pair_offset);
        SP2_LAUNCH_CHECK();
        k_publish_xchg<<<1, NF_THREADS, 0, ctx->stream>>>(P->partials, chunks * pairs, comm->dc, (int)t + 1, nn_mail_dev(P), mail_flag, seq);
      } else {
        // the round kernel publishes its two sums itself (column accumulators + last-CTA ticket): no k_publish launch
        k_nifs_round<<<dim3(chunks, pairs), threads, 0, ctx->stream>>>(t, d_rhos, ell_b, left, right, P->E, As, Bs, Cs, N, stride, P->partials, pair_offset,
                                                                        P->ticket, nn_mail_dev(P), mail_flag, seq);
      }
    }
    SP2_LAUNCH_CHECK();
    SP2_TRY(nn_wait(P, seq));
    if (timed_r0 && t == 0 && cudaEventSynchronize(P->ev_r0b) == cudaSuccess) cudaEventElapsedTime(&P->round0_ms, P->ev_r0a, P->ev_r0b);
    fe e0 = nn_mail(P)[0], quad = nn_mail(P)[1];
    if (local && !xchg) {
      // the round's two sums over ALL ranks' pairs: one 64-byte all-gather + modular adds on the host — every rank then
      // derives the identical polynomial and challenge
      fe mine2[2] = {e0, quad};
      std::vector<fe> all(2 * G);
      if (allgather(user, mine2, sizeof(mine2), all.data(), 0) != 0) return set_error(ctx, SP2_ERR_INTERNAL, "neutronnova: all-gather of the round sums failed");
      e0 = zero; quad = zero;
      for (u32 g = 0; g < G; g++) { e0 = HF::add(e0, all[2 * g]); quad = HF::add(quad, all[2 * g + 1]); }
    }
    H::store(pf->nifs_evals + 8 * t, e0); H::store(pf->nifs_evals + 8 * t + 4, quad);
    // finish_round! (neutronnova_zk.rs:703-735)
    const fe rho = rhos[t], omr = HF::sub(one, rho), trm = HF::sub(rho, omr);
    const fe c = HF::mul(e0, acc_eq), a = HF::mul(quad, acc_eq);
    const fe a_b_c = HF::mul(HF::sub(T_cur, HF::mul(c, omr)), rho_inv[t]);
    const fe b = HF::sub(HF::sub(a_b_c, a), c);
    fe co[4] = {HF::mul(c, omr), HF::add(HF::mul(c, trm), HF::mul(b, omr)), HF::add(HF::mul(b, trm), HF::mul(a, omr)), HF::mul(a, trm)};
    for (int k = 0; k < 4; k++) H::store(pf->nifs_polys + 16 * t + 4 * k, co[k]);
    absorb("p", co, 4);
    const fe r_b = squeeze("c");
    r_bs[t] = r_b; H::store(pf->r_b + 4 * t, r_b);
    acc_eq = HF::mul(acc_eq, HF::add(HF::mul(HF::sub(one, r_b), omr), HF::mul(r_b, rho)));
    T_cur = H::eval(co, 4, r_b);
    const u64 work = (m / 2) * N;
    const unsigned nb = (unsigned)std::min<u64>((work + NF_THREADS - 1) / NF_THREADS, (u64)ctx->num_sms * 8);
    k_nifs_fold_v<<<nb, NF_THREADS, 0, ctx->stream>>>(As, Bs, Cs, N, m / 2, stride, r_b);
    SP2_LAUNCH_CHECK();
    m /= 2; stride *= 2;
  }
  if (HF::is_zero(acc_eq)) return set_error(ctx, SP2_ERR_DIVISION_BY_ZERO, "neutronnova: acc_eq = 0");
  const fe T_out = HF::mul(T_cur, HF::inv(acc_eq));                      // :1206-1208
  H::store(pf->T_out, T_out);
  unsigned char *heads_stage = P->h_stage + NN_STAGE_BYTES - 28 * sizeof(fe);   // parity read-backs, copied out at the end
  if (pf->heads) { const fe *fold3[3] = {As, Bs, Cs};
    for (int k = 0; k < 3; k++) SP2_CUDA_OK(cudaMemcpyAsync(heads_stage + k * 4 * sizeof(fe), fold3[k], 4 * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream)); }
  ph[0] = ms_since(t_phase); t_phase = now();

  // ---- R1CSWitness::fold_multiple (src/r1cs/mod.rs:570-660) + the z tables of the inner sum-check -----------------
  SP2_TRY(stage(r_bs.data(), ell_b * sizeof(fe), P->small + 64));
  k_weights_from_r<<<(n_total + 127) / 128, 128, 0, ctx->stream>>>(P->small + 64, ell_b, n_total, P->small + 128);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemsetAsync(P->z_step + M, 0, (size_t)M * sizeof(fe), ctx->stream));
  SP2_CUDA_OK(cudaMemsetAsync(P->z_core + M, 0, (size_t)M * sizeof(fe), ctx->stream));
  { const unsigned nb = (unsigned)std::min<u64>((M + NF_THREADS - 1) / NF_THREADS, (u64)ctx->num_sms * 8);
    if (G == 1) {
      k_fold_vectors<<<nb, NF_THREADS, 0, ctx->stream>>>(P->Ws, n, M, P->small + 128, P->z_step);
      SP2_LAUNCH_CHECK();
    } else {
      // this rank's part of the weighted sum (its instances' global weights), all-gather of the G partial vectors, sum
      fe *mine = P->wpart + (size_t)rank * M;
      if (peer_x) {
        // the partial goes to a private scratch first (the exchange buffer may still be read by a slower peer's previous
        // phase only before barrier 30 of this prove, which every rank has passed by now), then to every rank's wpart
        fe *tmp = P->z_core;                                       // not yet initialised for this prove: free scratch of >= M entries
        k_fold_vectors<<<nb, NF_THREADS, 0, ctx->stream>>>(P->Ws, n, M, P->small + 128 + (size_t)rank * n, tmp);
        SP2_LAUNCH_CHECK();
        NnPeers pe; pe.n = (int)G; for (u32 q = 0; q < 8; q++) pe.x[q] = q < G ? P->peer_x[q] : nullptr;
        k_nn_scatter<<<ctx->num_sms, 256, 0, ctx->stream>>>(tmp, M, pe, (u64)G * 3 * N + (u64)rank * M);
        SP2_LAUNCH_CHECK();
        k_nn_barrier<<<1, 32, 0, ctx->stream>>>(comm->dc, 32, P->ticket + 1);
        SP2_LAUNCH_CHECK();
      } else {
        k_fold_vectors<<<nb, NF_THREADS, 0, ctx->stream>>>(P->Ws, n, M, P->small + 128 + (size_t)rank * n, mine);
        SP2_LAUNCH_CHECK();
        SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        if (allgather(user, mine, (uint64_t)M * sizeof(fe), P->wpart, 1) != 0) return set_error(ctx, SP2_ERR_INTERNAL, "neutronnova: all-gather of the witness partials failed");
      }
      std::vector<fe> ones(G, one);
      SP2_TRY(stage(ones.data(), G * sizeof(fe), P->small + 700));
      k_fold_vectors<<<nb, NF_THREADS, 0, ctx->stream>>>(P->wpart, G, M, P->small + 700, P->z_step);
      SP2_LAUNCH_CHECK();
    }
  }
  const u32 np = P->np;
  SP2_CUDA_OK(cudaMemcpyAsync(P->z_core, P->zc, (size_t)(M + 1 + np) * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));   // [W | 1 | X_core]
  k_set_one<<<1, 32, 0, ctx->stream>>>(P->z_step + M); SP2_LAUNCH_CHECK();
  std::vector<fe> Xacc(np + 1, zero), Xcore(np + 1, zero);                // [1 | X]: R1CSInstance::fold_multiple's X part (neutronnova_zk.rs:1231-1239)
  Xacc[0] = one; Xcore[0] = one;
  if (np) {
    std::vector<fe> w(n_total);
    for (u32 i = 0; i < n_total; i++) { fe wi = one; u32 k = i; for (u32 t = 0; t < ell_b; t++) { wi = HF::mul(wi, (k & 1u) ? r_bs[t] : HF::sub(one, r_bs[t])); k >>= 1; } w[i] = wi; }
    for (u32 i = 0; i < n_total; i++) for (u32 j = 0; j < np; j++) Xacc[1 + j] = HF::add(Xacc[1 + j], HF::mul(w[i], H::load(&P->Xs[((size_t)i * np + j) * 4])));
    for (u32 j = 0; j < np; j++) Xcore[1 + j] = H::load(&P->Xc[(size_t)j * 4]);
    SP2_TRY(stage(&Xacc[1], np * sizeof(fe), P->z_step + M + 1));
  }
  if (hooks && hooks->after_fold) SP2_TRY(hooks->after_fold(r_bs));
  if (pf->heads) SP2_CUDA_OK(cudaMemcpyAsync(heads_stage + 12 * sizeof(fe), P->z_step, 8 * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  ph[1] = ms_since(t_phase); t_phase = now();      // (asynchronous: the device time of this phase lands in the next one)

  // ---- HOT LOOP B: batched outer sum-check, 2 branches (src/sumcheck.rs:786-917) ----------------------------------
  std::vector<fe> tau_pow(ell + 1);                 // tau^(2^k)
  tau_pow[0] = tau; for (u32 k = 1; k <= ell; k++) tau_pow[k] = HF::sqr(tau_pow[k - 1]);
  fe base_tau = one, claim_s = T_out, claim_c = zero;
  std::vector<fe> r_x(ell);
  fe *step_t[3] = {As, Bs, Cs}, *core_t[3] = {Ac, Bc, Cc};
  u64 tl = N;
  NnTables otb; for (int k = 0; k < 3; k++) { otb.t[0][k] = step_t[k]; otb.t[1][k] = core_t[k]; }
  fe r_prev = zero;
  // evaluate a Karatsuba triple (value at 0, value at 1, r^2 coefficient) at r
  auto eval3 = [&](const fe &c00, const fe &c22, const fe &cdd, const fe &r) {
    return HF::add(c00, HF::mul(r, HF::add(HF::sub(HF::sub(c22, c00), cdd), HF::mul(r, cdd))));
  };
  const bool piped_outer = nn_pipe() && ell >= 2;
  u32 slot_seq[NN_SLOTS] = {0, 0, 0};
  auto coef_grid = [&](u64 h) { return (u32)std::max<u64>(1, std::min<u64>((h + NN_KC - 1) / NN_KC, (u64)ctx->num_sms * 2)); };
  auto outer_grid = [&](u64 work) { return (u32)std::max<u64>(1, std::min<u64>((work + NF_THREADS - 1) / NF_THREADS, (u64)ctx->num_sms)); };
  if (piped_outer) {
    // round 0 directly (slot 0) and the coefficient sums of round 1 from the unbound tables (slot 1): both before any challenge exists
    const u64 len = tl / 2;
    u32 nb;
    if (len >= left) nb = (u32)std::max<u64>(1, std::min<u64>(len / left, (u64)ctx->num_sms));
    else nb = (u32)((len + NF_THREADS - 1) / NF_THREADS);
    const u32 threads = len >= left ? std::min<u32>(NF_THREADS, (left + 31) / 32 * 32) : NF_THREADS;
    slot_seq[0] = ++P->seq;
    k_nn_outer_round<false><<<dim3(nb, 2), threads, 0, ctx->stream>>>(otb, P->E, left, P->E + left, len, zero, P->partials, P->ticket, nn_slot_dev(P, 0), nn_slot_flag_dev(P, 0), slot_seq[0]);
    SP2_LAUNCH_CHECK();
    slot_seq[1] = ++P->seq;
    k_nn_outer_coef<false><<<dim3(coef_grid(len / 2), 2), NN_CT, 0, ctx->stream>>>(otb, P->E, left, P->E + left, len, zero, P->partials, P->ticket, nn_slot_dev(P, 1), nn_slot_flag_dev(P, 1), slot_seq[1]);
    SP2_LAUNCH_CHECK();
    SP2_TRY(nn_wait_slot(P, 0, slot_seq[0]));
  }
  fe raw[6];
  if (piped_outer) for (int k = 0; k < 6; k++) raw[k] = nn_slot(P, 0)[k];
  for (u32 i = 0; i < ell; i++) {
    const u64 len = tl / 2;
    if (!piped_outer) {
      u32 nb;                                            // CTAs per branch
      if (len >= left) nb = (u32)std::max<u64>(1, std::min<u64>(len / left, (u64)ctx->num_sms));
      else nb = (u32)((len + NF_THREADS - 1) / NF_THREADS);
      const u32 threads = len >= left ? std::min<u32>(NF_THREADS, (left + 31) / 32 * 32) : NF_THREADS;
      const u32 seq = ++P->seq;
      if (i == 0) k_nn_outer_round<false><<<dim3(nb, 2), threads, 0, ctx->stream>>>(otb, P->E, left, P->E + left, len, r_prev, P->partials, P->ticket, nn_mail_dev(P), mail_flag, seq);
      else k_nn_outer_round<true><<<dim3(nb, 2), threads, 0, ctx->stream>>>(otb, P->E, left, P->E + left, len, r_prev, P->partials, P->ticket, nn_mail_dev(P), mail_flag, seq);
      SP2_LAUNCH_CHECK();
      SP2_TRY(nn_wait(P, seq));
      for (int k = 0; k < 6; k++) raw[k] = nn_mail(P)[k];
    }
    fe ev[6];
    for (int k = 0; k < 6; k++) { ev[k] = raw[k]; H::store(pf->outer_evals + 24 * i + 4 * k, ev[k]); ev[k] = HF::mul(ev[k], base_tau); }
    fe co[8];
    H::from_evals4(ev[0], HF::sub(claim_s, ev[0]), ev[1], ev[2], co);
    H::from_evals4(ev[3], HF::sub(claim_c, ev[3]), ev[4], ev[5], co + 4);
    for (int k = 0; k < 8; k++) H::store(pf->outer_polys + 32 * i + 4 * k, co[k]);
    absorb("p", co, 8);
    const fe r_i = squeeze("c");
    r_x[i] = r_i; H::store(pf->r_x + 4 * i, r_i);
    claim_s = H::eval(co, 4, r_i); claim_c = H::eval(co + 4, 4, r_i);
    r_prev = r_i;                                      // bound by the next round's launch (or by k_bind_heads after the last)
    tl /= 2;
    const fe pw = tau_pow[ell - 1 - i];             // tau^(N >> (i+1)) = E[len_pow % left] * E[left + len_pow / left]
    base_tau = HF::mul(base_tau, HF::add(HF::mul(HF::sub(pw, one), r_i), one));
    if (piped_outer && i + 1 < ell) {
      // kernel i+1: bind to r_i (+ the coefficient sums of round i+2); then this thread evaluates round i+1 from the coefficients kernel i published
      const u64 len_next = len / 2;
      if (i + 2 < ell) {
        const int sl = (int)((i + 2) % NN_SLOTS);
        slot_seq[sl] = ++P->seq;
        k_nn_outer_coef<true><<<dim3(coef_grid(len_next / 2), 2), NN_CT, 0, ctx->stream>>>(otb, P->E, left, P->E + left, len_next, r_i, P->partials, P->ticket, nn_slot_dev(P, sl), nn_slot_flag_dev(P, sl), slot_seq[sl]);
      } else {
        TablePtrs tp; tp.t[0] = As; tp.t[1] = Bs; tp.t[2] = Cs; tp.t[3] = Ac; tp.t[4] = Bc; tp.t[5] = Cc; tp.t[6] = nullptr; tp.t[7] = nullptr;
        k_bind_tables_v<<<1, NF_THREADS, 0, ctx->stream>>>(tp, 6, 2 * len_next, r_i);
      }
      SP2_LAUNCH_CHECK();
      const int sk = (int)((i + 1) % NN_SLOTS);
#ifdef SP2_NN_STAMPS
      static std::chrono::steady_clock::time_point t_prev; const auto t_l = now();
#endif
      SP2_TRY(nn_wait_slot(P, sk, slot_seq[sk]));
#ifdef SP2_NN_STAMPS
      fprintf(stderr, "host round %u: algebra+hash+launch since the previous wait ended %.1f us, waited %.1f us\n", i, std::chrono::duration<double, std::micro>(t_l - t_prev).count(), ms_since(t_l) * 1e3); t_prev = now();
#endif
      const fe *K = nn_slot(P, sk);
#ifdef SP2_NN_STAMPS
      { const unsigned long long *st = (const unsigned long long *)(K + 24); static unsigned long long prev_end = 0;
        fprintf(stderr, "outer coef kernel for round %u: gap since the previous kernel's end %.1f us, entry -> last CTA arrives %.1f us, -> sums published %.1f us\n", i + 1,
                prev_end ? (double)(st[1] - prev_end) / 1e3 : 0.0, (double)(st[2] - st[1]) / 1e3, (double)(st[0] - st[2]) / 1e3); prev_end = st[0]; }
#endif
      for (int b = 0; b < 2; b++) for (int t = 0; t < 3; t++) raw[3 * b + t] = eval3(K[9 * b + t], K[9 * b + 3 + t], K[9 * b + 6 + t], r_i);
    }
  }
  fe cl[6];
  { TablePtrs tp; tp.t[0] = As; tp.t[1] = Bs; tp.t[2] = Cs; tp.t[3] = Ac; tp.t[4] = Bc; tp.t[5] = Cc; tp.t[6] = nullptr; tp.t[7] = nullptr;
    const u32 seq = ++P->seq;
    k_bind_heads<<<1, 32, 0, ctx->stream>>>(tp, 6, r_prev, nn_mail_dev(P), mail_flag, seq);
    SP2_LAUNCH_CHECK();
    SP2_TRY(nn_wait(P, seq));
    for (int k = 0; k < 6; k++) { cl[k] = nn_mail(P)[k]; H::store(pf->claims_outer + 4 * k, cl[k]); } }
  H::store(pf->tau_at_rx, base_tau);
  pf->outer_ok = HF::eq(claim_s, HF::mul(base_tau, HF::sub(HF::mul(cl[0], cl[1]), cl[2]))) &&
                 HF::eq(claim_c, HF::mul(base_tau, HF::sub(HF::mul(cl[3], cl[4]), cl[5])));
  ph[2] = ms_since(t_phase); t_phase = now();

  // ---- batching challenge, eq(r_x), poly_ABC for both branches (neutronnova_zk.rs:1853-1875) ------------------------
  absorb("claims_outer", cl, 6);
  const fe r = squeeze("r");
  H::store(pf->r, r);
  const fe r2 = HF::sqr(r);
  fe claim_js = HF::add(HF::add(cl[0], HF::mul(r, cl[1])), HF::mul(r2, cl[2]));
  fe claim_jc = HF::add(HF::add(cl[3], HF::mul(r, cl[4])), HF::mul(r2, cl[5]));
  SP2_TRY(stage(r_x.data(), ell * sizeof(fe), P->small + 512));
  SP2_TRY(stage(&r, sizeof(fe), P->small + 600));
  SP2_TRY(eq_table_dev(ctx, P->small + 512, ell, P->rx));
  SP2_TRY(abc_dev(ctx, S, P->rx, P->small + 600, P->abc_s, 2 * M));
  SP2_TRY(abc_dev(ctx, S, P->rx, P->small + 600, P->abc_c, 2 * M));   // S_core == S_step for the SHA-256 chain; two calls as in the reference
  if (pf->heads) SP2_CUDA_OK(cudaMemcpyAsync(heads_stage + 20 * sizeof(fe), P->abc_s, 8 * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  ph[3] = ms_since(t_phase); t_phase = now();

  // ---- HOT LOOP C: batched inner sum-check (src/sumcheck.rs:702-782) -------------------------------------------------
  std::vector<fe> r_y(my);
  tl = 2 * M;
  NnTables itb; itb.t[0][0] = P->abc_s; itb.t[0][1] = P->z_step; itb.t[0][2] = nullptr; itb.t[1][0] = P->abc_c; itb.t[1][1] = P->z_core; itb.t[1][2] = nullptr;
  r_prev = zero;
  const bool piped_inner = nn_pipe() && my >= 2;
  fe rawi[4];
  if (piped_inner) {
    const u64 len = tl / 2;
    slot_seq[0] = ++P->seq;
    k_nn_inner_round<false><<<dim3(outer_grid(len), 2), NF_THREADS, 0, ctx->stream>>>(itb, len, zero, P->partials, P->ticket, nn_slot_dev(P, 0), nn_slot_flag_dev(P, 0), slot_seq[0]);
    SP2_LAUNCH_CHECK();
    slot_seq[1] = ++P->seq;
    k_nn_inner_coef<false><<<dim3(coef_grid(len / 2), 2), NN_CT, 0, ctx->stream>>>(itb, len, zero, P->partials, P->ticket, nn_slot_dev(P, 1), nn_slot_flag_dev(P, 1), slot_seq[1]);
    SP2_LAUNCH_CHECK();
    SP2_TRY(nn_wait_slot(P, 0, slot_seq[0]));
    for (int k = 0; k < 4; k++) rawi[k] = nn_slot(P, 0)[k];
  }
  for (u32 j = 0; j < my; j++) {
    const u64 len = tl / 2;
    if (!piped_inner) {
      const u32 nb = (u32)std::max<u64>(1, std::min<u64>((len + NF_THREADS - 1) / NF_THREADS, (u64)ctx->num_sms));
      const u32 seq = ++P->seq;
      if (j == 0) k_nn_inner_round<false><<<dim3(nb, 2), NF_THREADS, 0, ctx->stream>>>(itb, len, r_prev, P->partials, P->ticket, nn_mail_dev(P), mail_flag, seq);
      else k_nn_inner_round<true><<<dim3(nb, 2), NF_THREADS, 0, ctx->stream>>>(itb, len, r_prev, P->partials, P->ticket, nn_mail_dev(P), mail_flag, seq);
      SP2_LAUNCH_CHECK();
      SP2_TRY(nn_wait(P, seq));
      for (int k = 0; k < 4; k++) rawi[k] = nn_mail(P)[k];
    }
    fe co[6];
    for (int b = 0; b < 2; b++) {
      const fe e0 = rawi[2 * b], tinf = rawi[2 * b + 1], claim = b ? claim_jc : claim_js;
      H::store(pf->inner_evals + 16 * j + 8 * b, e0); H::store(pf->inner_evals + 16 * j + 8 * b + 4, tinf);
      const fe e2 = HF::add(HF::sub(HF::dbl(claim), HF::add(HF::dbl(e0), e0)), HF::dbl(tinf));        // BDDT (sumcheck.rs:731-733)
      H::from_evals3(e0, HF::sub(claim, e0), e2, co + 3 * b);
    }
    for (int k = 0; k < 6; k++) H::store(pf->inner_polys + 24 * j + 4 * k, co[k]);
    absorb("p", co, 6);
    const fe r_j = squeeze("c");
    r_y[j] = r_j; H::store(pf->r_y + 4 * j, r_j);
    r_prev = r_j;
    tl /= 2;
    claim_js = H::eval(co, 3, r_j); claim_jc = H::eval(co + 3, 3, r_j);
    if (piped_inner && j + 1 < my) {
      const u64 len_next = len / 2;
      if (j + 2 < my) {
        const int sl = (int)((j + 2) % NN_SLOTS);
        slot_seq[sl] = ++P->seq;
        k_nn_inner_coef<true><<<dim3(coef_grid(len_next / 2), 2), NN_CT, 0, ctx->stream>>>(itb, len_next, r_j, P->partials, P->ticket, nn_slot_dev(P, sl), nn_slot_flag_dev(P, sl), slot_seq[sl]);
      } else {
        TablePtrs tp; tp.t[0] = P->abc_s; tp.t[1] = P->abc_c; tp.t[2] = P->z_step; tp.t[3] = P->z_core; for (int k = 4; k < 8; k++) tp.t[k] = nullptr;
        k_bind_tables_v<<<1, NF_THREADS, 0, ctx->stream>>>(tp, 4, 2 * len_next, r_j);
      }
      SP2_LAUNCH_CHECK();
      const int sk = (int)((j + 1) % NN_SLOTS);
      SP2_TRY(nn_wait_slot(P, sk, slot_seq[sk]));
      const fe *K = nn_slot(P, sk);
      for (int b = 0; b < 2; b++) for (int v = 0; v < 2; v++) rawi[2 * b + v] = eval3(K[6 * b + 3 * v], K[6 * b + 3 * v + 1], K[6 * b + 3 * v + 2], r_j);
    }
  }
  fe fi[4];
  { TablePtrs tp; tp.t[0] = P->abc_s; tp.t[1] = P->abc_c; tp.t[2] = P->z_step; tp.t[3] = P->z_core; for (int k = 4; k < 8; k++) tp.t[k] = nullptr;
    const u32 seq = ++P->seq;
    k_bind_heads<<<1, 32, 0, ctx->stream>>>(tp, 4, r_prev, nn_mail_dev(P), mail_flag, seq);
    SP2_LAUNCH_CHECK();
    SP2_TRY(nn_wait(P, seq));
    for (int k = 0; k < 4; k++) { fi[k] = nn_mail(P)[k]; H::store(pf->inner_final + 4 * k, fi[k]); } }
  // eval_W = (eval_Z - r_y[0] * eval_X) / (1 - r_y[0]); X = [1]: eval_X = prod (1 - r_y[1..])
  // SparsePolynomial::evaluate of [1 | X] at r_y[1..] (polys/multilinear.rs:190-207)
  auto sparse_eval = [&](const std::vector<fe> &Xv) {
    const size_t zlen = Xv.size(), nvars = my - 1; size_t p2 = 1, nvz = 0; while (p2 < zlen) { p2 <<= 1; nvz++; }
    const size_t skip = nvars - 1 - nvz, k = nvars - skip;
    std::vector<fe> chis((size_t)1 << k); chis[0] = one; size_t size = 1;
    for (size_t t = k; t-- > 0;) {                      // EqPolynomial::evals_from_points (eq.rs:59-92), MSB-first
      const fe rt = r_y[1 + skip + t];
      for (size_t i = 0; i < size; i++) { const fe hi = HF::mul(chis[i], rt); chis[size + i] = hi; chis[i] = HF::sub(chis[i], hi); }
      size *= 2;
    }
    fe acc = zero;
    for (size_t i = 0; i < zlen; i++) acc = HF::add(acc, HF::mul(Xv[i], chis[i]));
    fe common = one;
    for (size_t i = 0; i < skip; i++) common = HF::mul(common, HF::sub(one, r_y[1 + i]));
    return HF::mul(common, acc);
  };
  const fe eval_X = sparse_eval(Xacc), eval_Xc = sparse_eval(Xcore);
  const fe den = HF::sub(one, r_y[0]);
  if (HF::is_zero(den)) return set_error(ctx, SP2_ERR_DIVISION_BY_ZERO, "neutronnova: r_y[0] = 1");
  const fe inv = HF::inv(den);
  H::store(pf->eval_W, HF::mul(HF::sub(fi[2], HF::mul(r_y[0], eval_X)), inv));
  H::store(pf->eval_W + 4, HF::mul(HF::sub(fi[3], HF::mul(r_y[0], eval_Xc)), inv));
  pf->inner_ok = HF::eq(claim_js, HF::mul(fi[0], fi[2])) && HF::eq(claim_jc, HF::mul(fi[1], fi[3]));
  ph[4] = ms_since(t_phase);
  if (pf->heads) { SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream)); memcpy(pf->heads, heads_stage, 28 * sizeof(fe)); }
  if (peer_x) {                                     // did every peer-store barrier complete?
    u32 *h_err = (u32 *)(P->h_stage + NN_STAGE_BYTES - 28 * sizeof(fe) - 64);
    SP2_CUDA_OK(cudaMemcpyAsync(h_err, P->ticket + 1, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
    SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (*h_err) { cudaMemsetAsync(P->ticket + 1, 0, sizeof(u32), ctx->stream); return set_error(ctx, SP2_ERR_INTERNAL, "neutronnova: a peer did not reach an exchange barrier within 2 s"); }
  }
  ph[5] = ms_since(t_begin);
  if (phase_ms) memcpy(phase_ms, ph, sizeof(ph));
  return SP2_OK;
}
/* CUDA-IPC handle (64 bytes) of this rank's exchange buffer, to be all-gathered by the host and passed to
 * sp2_neutronnova_prep_connect on every rank: afterwards the two bulk exchanges of a sharded prove are peer stores. */
int32_t sp2_neutronnova_prep_ipc_handle(sp2_nn_prep *P, uint8_t *out64) {
  sp2_ctx *ctx = P->ctx;
  cudaSetDevice(ctx->device);
  if (!P->xbuf) return set_error(ctx, SP2_ERR_INTERNAL, "neutronnova: not a sharded prep state");
  cudaIpcMemHandle_t h;
  SP2_CUDA_OK(cudaIpcGetMemHandle(&h, P->xbuf));
  memcpy(out64, &h, 64);
  return SP2_OK;
}
int32_t sp2_neutronnova_prep_connect(sp2_nn_prep *P, const uint8_t *all_handles) {
  sp2_ctx *ctx = P->ctx;
  cudaSetDevice(ctx->device);
  if (!P->xbuf) return set_error(ctx, SP2_ERR_INTERNAL, "neutronnova: not a sharded prep state");
  for (int q = 0; q < P->nranks; q++) {
    if (q == P->rank || P->peer_opened[q]) continue;
    cudaIpcMemHandle_t h; memcpy(&h, all_handles + 64 * q, 64);
    void *p = nullptr;
    SP2_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    P->peer_x[q] = (fe *)p; P->peer_opened[q] = true;
  }
  P->peers_connected = true;
  return SP2_OK;
}
/* in-process variant (ranks in one process: no CUDA IPC): xbufs[q] = rank q's sp2_neutronnova_prep_xbuf() pointer */
int32_t sp2_neutronnova_prep_xbuf(sp2_nn_prep *P, void **out) {
  if (!P->xbuf) return set_error(P->ctx, SP2_ERR_INTERNAL, "neutronnova: not a sharded prep state");
  *out = P->xbuf;
  return SP2_OK;
}
int32_t sp2_neutronnova_prep_connect_ptrs(sp2_nn_prep *P, void *const *xbufs) {
  if (!P->xbuf) return set_error(P->ctx, SP2_ERR_INTERNAL, "neutronnova: not a sharded prep state");
  for (int q = 0; q < P->nranks; q++) if (q != P->rank) P->peer_x[q] = (fe *)xbufs[q];
  P->peers_connected = true;
  return SP2_OK;
}
/* measurement hook: device time (CUDA events on the library's stream) of the last NIFS round-0 kernel (the i64 layers pass) */
int32_t sp2_neutronnova_last_round0_ms(sp2_nn_prep *P, float *ms) { if (!P || P->round0_ms < 0) return SP2_ERR_INTERNAL; *ms = P->round0_ms; return SP2_OK; }
int32_t sp2_neutronnova_prove(sp2_ctx *ctx, sp2_nn_prep *P, sp2_transcript *tsh, sp2_nn_proof *pf, float *phase_ms) {
  return nn_prove_impl(ctx, P, tsh, nullptr, nullptr, nullptr, pf, phase_ms);
}
/* Sharded prove: every rank calls with its own prep state and an identical transcript and receives the identical
 * proof.  `allgather(user, send, bytes, recv, on_device)` is the host's collective (NCCL through torch.distributed in
 * this repo's harness): recv = the ranks' `bytes`-long contributions in rank order; on_device = 1 when both pointers are
 * device memory (the library has synchronised its stream before the call and expects the data to be complete on
 * return).  Per prove: one device gather of the surviving layer triples (3 N scalars per rank) and one of the witness
 * partials (M scalars per rank); the two sums of every local NIFS round cross ranks INSIDE the publish kernel through
 * `comm`'s peer mailboxes (NVLink stores), or, when comm is NULL, through a 64-byte host gather per round. */
int32_t sp2_neutronnova_prove_sharded(sp2_ctx *ctx, sp2_nn_prep *P, sp2_transcript *tsh, sp2_comm *comm, sp2_allgather_fn allgather, void *user,
                                      sp2_nn_proof *pf, float *phase_ms) {
  return nn_prove_impl(ctx, P, tsh, comm, allgather, user, pf, phase_ms);
}

}  // extern "C"

// =====================================================================================================================
// The commitment half of the NeutronNova prove (non-ZK variant; the checker is oracle/oracle.c: orc_neutronnova_prove /
// orc_neutronnova_verify, which follow NeutronNovaZkSNARK::{prove, verify}, src/neutronnova_zk.rs:1609-2093, 2096-2330, with
// the in-circuit verifier replaced by direct absorbs):
//   prep_commit : per-instance commit of the precommitted section (bellpepper/r1cs.rs:359-408 -> HyraxPCS::commit)
//   snark_prove : rerandomize_commitment per step (hyrax_pc.rs:321-344) + commit_zeros of the rest rows (:305-319), the
//                 transcript over all instances, HOT LOOPS A-C (nn_prove_impl), fold_blinds / fold_commitments_partial
//                 (:795-874), the c_eval fold of the step and core claims (neutronnova_zk.rs:2019-2051) and PCS::prove
//                 on W_fold + c_eval W_core (:2053-2064 -> hyrax_pc.rs:387-478, ipa.rs:125-170).
// B200 design: every commitment operation is FIXED-BASE.  A Hyrax row commitment is linear in (row, blind), so
//   rerandomised row           = U_row + r_new h                      (U_row = the unblinded row, cached by prep_commit)
//   fold of n commitments      = commit(folded witness row, folded blind)           instead of n variable-base scalar
//   fold with c_eval           = comm(W_fold row) + c_eval U_core_row + (b_fold + c_eval b_core) h       multiplications,
// i.e. table gathers only (the key's window tables, plus 33 x 128 tables of the core's unblinded rows built at prep time):
// no doubling chain anywhere — a 256-bit variable-base scalar multiplication is ~256 serial doublings of ~5 us each on a GPU
// thread.  The one full-width MSM (13 rows x 2048 of the folded witness) runs on a side stream under the two sum-checks.
// The group elements are the ones the reference computes, so the affine outputs are bit-identical (tests).
// =====================================================================================================================
namespace {

__global__ void __launch_bounds__(256) k_nn_axpy(fe *out, const fe *a, const fe *b, fe c, u64 n) {      // out = a + c * b
  for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) stg_fe(out + i, Fq::add(ldg_fe(a + i), Fq::mul(c, ldg_fe(b + i))));
}
__global__ void __launch_bounds__(256) k_nn_dot(const fe *a, const fe *b, u64 n, fe *out) {
  __shared__ fe red[32];
  Fq::acc acc = Fq::acc_zero();
  for (u64 i = threadIdx.x; i < n; i += blockDim.x) Fq::mul_acc(acc, ldg_fe(a + i), ldg_fe(b + i));
  fe x[1] = {Fq::acc_reduce(acc)};
  block_sum_fq<1>(x, red);
  if (threadIdx.x == 0) stg_fe(out, x[0]);
}
enum NnSlot { NS_EVAL = 0 /* 2 */, NS_BEVAL = 2 /* 2 */, NS_EVALF = 4, NS_BEVALF = 5, NS_CEVAL = 6, NS_RLZ = 7, NS_IP = 8, NS_RDELTA = 9, NS_RBETA = 10, NS_RLZC = 11, NS_RY = 16 /* <= 40 */, NS_COUNT = 64 };

// gather CTAs the background commitment of the folded witness may occupy while the sum-checks run (SP2_NN_SIDE_CTAS, default 96; swept on B200 at 32 steps: 0 = uncapped: outer 1.04 / pcs 0.33 ms, 48: 0.57 / 0.83, 16: 0.56 / 3.9)
unsigned nn_side_ctas() { static int v = -1; if (v < 0) { const char *e = getenv("SP2_NN_SIDE_CTAS"); v = e ? atoi(e) : 96; } return (unsigned)v; }

int nn_palloc(sp2_nn_prep *P, size_t bytes, void **out) {
  sp2_ctx *ctx = P->ctx; void *p;
  SP2_CUDA_OK(cudaMalloc(&p, std::max<size_t>(bytes, 64)));
  P->owned.push_back(p); *out = p;
  return SP2_OK;
}
}  // namespace

extern "C" {

/* prep_prove's commitment half: blinds_pre_* are the blinds of the precommitted rows (n_steps x pre_rows, pre_rows); comm_pre_*_out
 * receive commit(W_i[precommitted], blinds) — the PrecommittedState a Rust caller keeps (bellpepper/r1cs.rs:359-408). */
int32_t sp2_neutronnova_prep_commit(sp2_ctx *ctx, sp2_nn_prep *P, const sp2_ck *ck, const uint64_t *blinds_pre_steps, const uint64_t *blinds_pre_core,
                                    uint64_t *comm_pre_steps_out, uint64_t *comm_pre_core_out) {
  cudaSetDevice(ctx->device);
  if (!P || !ck) return set_error(ctx, SP2_ERR_INTERNAL, "prep_commit: null argument");
  const sp2_shape *S = P->S;
  const uint64_t width = ck->n, M = P->M;
  if (S->num_shared) return set_error(ctx, SP2_ERR_UNSUPPORTED, "prep_commit: circuits with a shared witness section are not offloaded");
  if (M % width || S->num_precommitted % width) return set_error(ctx, SP2_ERR_INVALID_WITNESS_LENGTH, "prep_commit: witness sections must be multiples of the commitment width");
  if (P->U) return set_error(ctx, SP2_ERR_INTERNAL, "prep_commit: already committed");
  const uint32_t n = P->n, rows = (uint32_t)(M / width), pre_rows = (uint32_t)(S->num_precommitted / width);
  P->ck = ck; P->rows = rows; P->pre_rows = pre_rows;
  const size_t nrow = (size_t)(n + 1) * rows;
  void *p;
  SP2_TRY(nn_palloc(P, nrow * sizeof(aff), &p)); P->U = (aff *)p;
  SP2_TRY(nn_palloc(P, (nrow + 2 * rows + 16) * sizeof(jac) + nrow * sizeof(aff), &p)); P->pts = (jac *)p;   // [instance rows | folded rows | 2 eval | final rows + 4] | affine rows
  SP2_TRY(nn_palloc(P, (nrow + 2 * rows + 16 + (size_t)P->n_total * rows) * sizeof(fe), &p)); P->blinds_dev = (fe *)p;   // local rows | folded | final | all instances' blinds
  SP2_TRY(nn_palloc(P, M * sizeof(fe), &p)); P->Wfold = (fe *)p;
  SP2_TRY(nn_palloc(P, M * sizeof(fe), &p)); P->Wfin = (fe *)p;
  SP2_TRY(nn_palloc(P, (5 * width + 2 * rows + NS_COUNT + 64) * sizeof(fe), &p)); P->pcs = (fe *)p;
  SP2_CUDA_OK(cudaStreamCreateWithFlags(&P->side, cudaStreamNonBlocking));
  SP2_CUDA_OK(cudaEventCreateWithFlags(&P->ev_fold, cudaEventDisableTiming));
  SP2_CUDA_OK(cudaEventCreateWithFlags(&P->ev_side, cudaEventDisableTiming));
  // commit_without_blind of every row of every instance (hyrax_pc.rs:533-567); the zero rest rows come out as the identity
  std::vector<MsmJob> jobs(nrow);
  for (uint32_t i = 0; i <= n; i++)
    for (uint32_t r = 0; r < rows; r++) {
      MsmJob &j = jobs[(size_t)i * rows + r]; memset(&j, 0, sizeof(j));
      j.scalars = (i < n ? P->Ws + (size_t)i * M : P->zc) + (size_t)r * width; j.len = (u32)width;
    }
  SP2_TRY(msm_run(ctx, ck, jobs, P->pts));
  SP2_TRY(batch_normalize_dev(ctx, P->pts, nrow, P->U));
  // the caller's PrecommittedState: commit(W[precommitted], blinds) = U + blind * h
  const size_t npre = (size_t)(n + 1) * pre_rows;
  if (npre) {
    std::vector<uint64_t> hb(npre * 4);
    memcpy(hb.data(), blinds_pre_steps, (size_t)n * pre_rows * 32); memcpy(hb.data() + (size_t)n * pre_rows * 4, blinds_pre_core, (size_t)pre_rows * 32);
    SP2_CUDA_OK(cudaMemcpyAsync(P->blinds_dev, hb.data(), npre * 32, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<MsmJob> jb(npre);
    for (uint32_t i = 0; i <= n; i++)
      for (uint32_t r = 0; r < pre_rows; r++) {
        MsmJob &j = jb[(size_t)i * pre_rows + r]; memset(&j, 0, sizeof(j));
        j.nextra = 1; j.extra_base[0] = ck->idx_h(); j.extra_scalar[0] = P->blinds_dev + (size_t)i * pre_rows + r;
        j.add_aff = P->U + (size_t)i * rows + r;
      }
    SP2_TRY(msm_run(ctx, ck, jb, P->pts));
    std::vector<uint64_t> hj(npre * 12), ha(npre * 8);
    SP2_CUDA_OK(cudaMemcpyAsync(hj.data(), P->pts, npre * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
    SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    sp2h::batch_normalize(hj.data(), npre, ha.data());
    if (comm_pre_steps_out) memcpy(comm_pre_steps_out, ha.data(), (size_t)n * pre_rows * 64);
    if (comm_pre_core_out) memcpy(comm_pre_core_out, ha.data() + (size_t)n * pre_rows * 8, (size_t)pre_rows * 64);
  }
  // window tables of the core's unblinded rows: c_eval * U_core[row] becomes 33 table lookups
  SP2_TRY(msm_build_tables(ctx, P->U + (size_t)n * rows, rows, &P->Ucore_tab));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

/* phase_ms (optional, 10 floats, host wall clock): rerandomize + commit_zeros, instance transcript, nifs, fold_witness,
 * outer_sumcheck_batched, compute_eval_table_sparse, inner_sumcheck_batched, eval commitments + c_eval, pcs_prove, total */
static int32_t nn_snark_impl(sp2_ctx *ctx, sp2_nn_prep *P, sp2_comm *xcomm, sp2_allgather_fn allgather, void *user, const uint8_t *vk_digest,
                             const sp2_nn_rand *rnd, sp2_nn_snark *sn, float *phase_ms) {
  cudaSetDevice(ctx->device);
  if (!P || !rnd || !sn || !vk_digest) return set_error(ctx, SP2_ERR_INTERNAL, "snark_prove: null argument");
  if (P->nranks > 1 && !allgather) return set_error(ctx, SP2_ERR_INTERNAL, "snark_prove: a sharded prep state needs the all-gather callback (instance commitments)");
  if (!P->U) return set_error(ctx, SP2_ERR_INTERNAL, "snark_prove: sp2_neutronnova_prep_commit has not run on this prep state");
  typedef NnHost H;
  const sp2_ck *ck = P->ck;
  // n: the step instances THIS rank holds (all of them on one GPU); their blinds are rows [rank * n, (rank + 1) * n) of rnd->blinds_steps
  const uint32_t n = P->n, n_total = P->n_total, rows = P->rows, pre_rows = P->pre_rows, np = P->np;
  const uint64_t width = ck->n, M = P->M;
  const size_t nrow = (size_t)(n + 1) * rows;
  const uint64_t *my_blinds = rnd->blinds_steps + (size_t)P->rank * n * rows * 4;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms_since = [](std::chrono::steady_clock::time_point a) { return (float)(std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - a).count()); };
  const auto t_begin = now(); auto t_phase = t_begin;
  float ph[10] = {0};
  sn->rows = rows;
  fe *small = P->pcs, *LZ = small + NS_COUNT, *Ltab = LZ + width, *Rtab = Ltab + rows, *dvec = Rtab + width, *zvec = dvec + width;
  fe *blind_fold = P->blinds_dev + nrow, *blind_fin = blind_fold + rows, *blinds_all = blind_fin + rows + 16;
  // pinned staging: [blinds | d_vec | all instances' blinds | Jacobian read-backs]
  const size_t in_fe = nrow + width + NS_COUNT + 16 + (P->nranks > 1 ? (size_t)n_total * rows : 0);      // (also receives the NS_COUNT small scalars read back at the end)
  const size_t stage_need = in_fe * sizeof(fe) + (nrow + rows + 8) * sizeof(jac);
  void *hp; SP2_TRY(pinned(ctx, stage_need, &hp));
  uint8_t *h_in = (uint8_t *)hp; uint64_t *h_jac = (uint64_t *)(h_in + in_fe * sizeof(fe));
  // ---- rerandomize_commitment (precommitted rows) + commit_zeros (rest rows): row = U_row + blind * h ----------------
  memcpy(h_in, my_blinds, (size_t)n * rows * 32); memcpy(h_in + (size_t)n * rows * 32, rnd->blinds_core, (size_t)rows * 32);
  memcpy(h_in + nrow * 32, rnd->d_vec, width * 32);
  SP2_CUDA_OK(cudaMemcpyAsync(P->blinds_dev, h_in, nrow * 32, cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(dvec, h_in + nrow * 32, width * 32, cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(small + NS_RDELTA, rnd->r_delta, 32, cudaMemcpyHostToDevice, ctx->stream));   // (for the early delta MSM below)
  const fe *fold_src = P->blinds_dev;                        // [instance][row] blinds the fold runs over
  if (P->nranks > 1) {                                       // fold_blinds needs every instance's blinds (identical inputs on all ranks)
    uint8_t *h_all = h_in + (nrow + width + NS_COUNT + 16) * 32;
    memcpy(h_all, rnd->blinds_steps, (size_t)n_total * rows * 32);
    SP2_CUDA_OK(cudaMemcpyAsync(blinds_all, h_all, (size_t)n_total * rows * 32, cudaMemcpyHostToDevice, ctx->stream));
    fold_src = blinds_all;
  }
  { std::vector<MsmJob> jobs(nrow);
    for (size_t k = 0; k < nrow; k++) {
      MsmJob &j = jobs[k]; memset(&j, 0, sizeof(j));
      j.nextra = 1; j.extra_base[0] = ck->idx_h(); j.extra_scalar[0] = P->blinds_dev + k; j.add_aff = P->U + k;
    }
    SP2_TRY(msm_run(ctx, ck, jobs, P->pts)); }
  if (nrow > 1024) {
    // many instances: normalise on the device (one inversion per 16-point chunk, chunks in parallel) and read back affine rows;
    // a serial host pass over thousands of points costs ~1 ms
    aff *d_aff = (aff *)(P->pts + nrow + 2 * rows + 16);
    SP2_TRY(batch_normalize_dev(ctx, P->pts, nrow, d_aff));
    uint64_t *dst = sn->comm_W_steps + (size_t)P->rank * n * rows * 8;
    SP2_CUDA_OK(cudaMemcpyAsync(dst, d_aff, (size_t)n * rows * sizeof(aff), cudaMemcpyDeviceToHost, ctx->stream));
    SP2_CUDA_OK(cudaMemcpyAsync(sn->comm_W_core, d_aff + (size_t)n * rows, (size_t)rows * sizeof(aff), cudaMemcpyDeviceToHost, ctx->stream));
    SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  } else {
    SP2_CUDA_OK(cudaMemcpyAsync(h_jac, P->pts, nrow * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
    SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    std::vector<uint64_t> aff_all(nrow * 8);
    sp2h::batch_normalize(h_jac, nrow, aff_all.data());
    memcpy(sn->comm_W_steps + (size_t)P->rank * n * rows * 8, aff_all.data(), (size_t)n * rows * 64); memcpy(sn->comm_W_core, aff_all.data() + (size_t)n * rows * 8, (size_t)rows * 64);
  }
  if (P->nranks > 1) {
    // every rank rerandomised its own instances: all-gather the rows (rank order = instance order; 64 B per row) so that every
    // rank absorbs the same transcript.  A host collective: the rows have to reach the host for the Keccak transcript anyway.
    std::vector<uint64_t> mine(sn->comm_W_steps + (size_t)P->rank * n * rows * 8, sn->comm_W_steps + (size_t)(P->rank + 1) * n * rows * 8);
    if (allgather(user, mine.data(), (uint64_t)n * rows * 64, sn->comm_W_steps, 0) != 0) return set_error(ctx, SP2_ERR_INTERNAL, "snark_prove: all-gather of the instance commitments failed");
  }
  ph[0] = ms_since(t_phase); t_phase = now();
  // ---- transcript over the instances (neutronnova_zk.rs:1727-1733, 552-556; R1CSInstance bytes r1cs/mod.rs:728-736) -----
  sp2_transcript tsobj("neutronnova_prove");
  sp2h::Transcript &ts = tsobj.t;
  ts.absorb_bytes("vk", vk_digest, 32);
  auto absorb_instance = [&](const char *label, const uint64_t *comm, const uint64_t *X) {
    ts.push(label, strlen(label)); ts.push("poly_commitment_begin", 21);
    for (uint32_t r = 0; r < rows; r++) ts.push_point(comm + 8 * r);
    ts.push("poly_commitment_end", 19);
    for (uint32_t j = 0; j < np; j++) { uint64_t c[4]; uint8_t b[32]; sp2h::from_mont(X + 4 * j, sp2h::FQ_MOD, sp2h::FQ_INV, c); sp2h::limbs_to_be(c, b); ts.push(b, 32); }
  };
  absorb_instance("core_instance", sn->comm_W_core, np ? P->Xc.data() : nullptr);
  for (uint32_t i = 0; i < n_total; i++) absorb_instance("U", sn->comm_W_steps + (size_t)i * rows * 8, np ? &P->Xs[(size_t)i * np * 4] : nullptr);
  ph[1] = ms_since(t_phase); t_phase = now();
  // ---- HOT LOOPS A-C; right after the witness fold: copy W_fold, fold the blinds, and start the commitment of the folded
  // witness rows on the side stream (it overlaps the outer and inner sum-checks) -------------------------------------------
  jac *pts_fold = P->pts + nrow;
  sp2::NnHooks hooks;
  hooks.after_fold = [&](const std::vector<fe> &) -> int {
    SP2_CUDA_OK(cudaMemcpyAsync(P->Wfold, P->z_step, M * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
    // fold_blinds (hyrax_pc.rs:795-819): blinds_dev is [instance][row] = n vectors of `rows` entries; weights at small + 128
    k_fold_vectors<<<1, NF_THREADS, 0, ctx->stream>>>(fold_src, n_total, rows, P->small + 128, blind_fold);
    SP2_LAUNCH_CHECK();
    SP2_CUDA_OK(cudaEventRecord(P->ev_fold, ctx->stream));
    SP2_CUDA_OK(cudaStreamWaitEvent(P->side, P->ev_fold, 0));
    if (pre_rows) {
      std::vector<MsmJob> jobs(pre_rows);
      for (uint32_t r = 0; r < pre_rows; r++) { MsmJob &j = jobs[r]; memset(&j, 0, sizeof(j)); j.scalars = P->Wfold + (size_t)r * width; j.len = (u32)width; }
      SP2_TRY(msm_run(ctx, ck, jobs, pts_fold, P->side, 16, 17, nn_side_ctas()));
    }
    // delta = <d, ck> + r_delta h (ipa.rs:140-146) depends on the prover's randomness only: same side stream, under the sum-checks
    { MsmJob j; memset(&j, 0, sizeof(j)); j.scalars = dvec; j.len = (u32)width; j.nextra = 1; j.extra_base[0] = ck->idx_h(); j.extra_scalar[0] = small + NS_RDELTA;
      SP2_TRY(msm_run(ctx, ck, std::vector<MsmJob>{j}, pts_fold + rows + 2 + rows + 3, P->side, 21, 22, nn_side_ctas())); }
    SP2_CUDA_OK(cudaEventRecord(P->ev_side, P->side));
    return SP2_OK;
  };
  float ph_core[6];
  SP2_TRY(nn_prove_impl(ctx, P, &tsobj, xcomm, allgather, user, &sn->base, ph_core, &hooks));
  for (int k = 0; k < 5; k++) ph[2 + k] = ph_core[k];
  t_phase = now();
  // ---- commitments to eval_W_step / eval_W_core, c_eval (neutronnova_zk.rs:1953-2017 in its non-ZK form) ------------------
  const fe eval_s = H::load(sn->base.eval_W), eval_c = H::load(sn->base.eval_W + 4);
  const fe be_s = H::load(rnd->blind_eval_W), be_c = H::load(rnd->blind_eval_W + 4);
  memcpy(sn->blind_eval_W, rnd->blind_eval_W, 64);
  { fe up[4] = {eval_s, eval_c, be_s, be_c};
    memcpy(h_in, up, sizeof(up));
    SP2_CUDA_OK(cudaMemcpyAsync(small + NS_EVAL, h_in, sizeof(up), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<MsmJob> jobs(2);
    for (int b = 0; b < 2; b++) { MsmJob &j = jobs[b]; memset(&j, 0, sizeof(j)); j.nextra = 2; j.extra_base[0] = ck->idx_ck_s(); j.extra_scalar[0] = small + NS_EVAL + b;
      j.extra_base[1] = ck->idx_h_s(); j.extra_scalar[1] = small + NS_BEVAL + b; }
    SP2_TRY(msm_run(ctx, ck, jobs, pts_fold + rows));
    SP2_CUDA_OK(cudaMemcpyAsync(h_jac, pts_fold + rows, 2 * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
    SP2_CUDA_OK(cudaEventRecord(P->ev_fold, ctx->stream)); }
  // ---- everything of PCS::prove that does not depend on c_eval runs while the host normalises / hashes the two evaluation
  // commitments (a host round trip of ~65 us): the eq tables of r_y, <R, d>, and — by linearity of the bind — L^T W_fold, L^T W_core,
  // <L, b_fold>, <L, b_core>; after c_eval only three small axpys remain (LZ, r_LZ, the folded blinds): W_fold + c W_core itself is
  // never materialised
  const uint32_t my = sn->base.rounds_y, m = my - 1;
  int nvr = 0; while ((1u << nvr) < rows) nvr++;
  if ((1u << nvr) != rows || (width & (width - 1))) return set_error(ctx, SP2_ERR_INVALID_CK_LENGTH, "snark_prove: rows and width must be powers of two");
  { fe up[SC_MAX_ROUNDS + 8];
    for (uint32_t j = 0; j < m; j++) up[j] = H::load(sn->base.r_y + 4 * (j + 1));
    up[m] = H::load(rnd->r_beta);
    memcpy(h_in + 64 * sizeof(fe), up, (m + 1) * sizeof(fe));
    SP2_CUDA_OK(cudaMemcpyAsync(small + NS_RY, h_in + 64 * sizeof(fe), m * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
    SP2_CUDA_OK(cudaMemcpyAsync(small + NS_RBETA, h_in + (64 + m) * sizeof(fe), sizeof(fe), cudaMemcpyHostToDevice, ctx->stream)); }
  SP2_TRY(eq_table_dev(ctx, small + NS_RY + nvr, (uint32_t)(m - nvr), Rtab));
  k_nn_dot<<<1, 256, 0, ctx->stream>>>(Rtab, dvec, width, small + NS_IP);
  SP2_LAUNCH_CHECK();
  if (nvr > 0) {
    SP2_TRY(eq_table_dev(ctx, small + NS_RY, (uint32_t)nvr, Ltab));
    SP2_TRY(hyrax_bind_dev(ctx, P->Wfold, Ltab, rows, width, LZ));              // L^T W_fold
    SP2_TRY(hyrax_bind_dev(ctx, P->zc, Ltab, rows, width, zvec));               // L^T W_core (zvec is free until the IPA response)
    k_nn_dot<<<1, 256, 0, ctx->stream>>>(Ltab, blind_fold, rows, small + NS_RLZ);
    SP2_LAUNCH_CHECK();
    k_nn_dot<<<1, 256, 0, ctx->stream>>>(Ltab, P->blinds_dev + (size_t)n * rows, rows, small + NS_RLZC);
    SP2_LAUNCH_CHECK();
  }
  SP2_CUDA_OK(cudaEventSynchronize(P->ev_fold));                                // the two evaluation commitments are on the host
  uint64_t ce[16];
  sp2h::batch_normalize(h_jac, 2, ce);
  if (sn->comm_eval_W) memcpy(sn->comm_eval_W, ce, 128);
  ts.absorb_point("comm_eval_W_step", ce); ts.absorb_point("comm_eval_W_core", ce + 8);
  fe c_eval; { uint8_t dg[64]; uint64_t o[4]; ts.squeeze("c_eval", dg); sp2h::fq_from_uniform(dg, o); c_eval = H::load(o); }
  if (sn->c_eval) H::store(sn->c_eval, c_eval);
  ph[7] = ms_since(t_phase); t_phase = now();
  // ---- fold the step and core claims with c_eval and open (neutronnova_zk.rs:2019-2064) -----------------------------------
  const fe eval_f = HF::add(eval_s, HF::mul(c_eval, eval_c)), be_f = HF::add(be_s, HF::mul(c_eval, be_c));
  { fe up[4] = {eval_f, be_f, c_eval, HF::zero()};
    memcpy(h_in, up, sizeof(up));
    SP2_CUDA_OK(cudaMemcpyAsync(small + NS_EVALF, h_in, 3 * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream)); }
  k_nn_axpy<<<1, 256, 0, ctx->stream>>>(blind_fin, blind_fold, P->blinds_dev + (size_t)n * rows, c_eval, rows);   // fold_blinds with [1, c_eval]
  SP2_LAUNCH_CHECK();
  const fe *LZp = LZ, *rLZ = small + NS_RLZ;
  if (nvr > 0) {
    k_nn_axpy<<<(unsigned)((width + 255) / 256), 256, 0, ctx->stream>>>(LZ, LZ, zvec, c_eval, width);          // LZ = L^T W_fold + c_eval L^T W_core
    SP2_LAUNCH_CHECK();
    k_nn_axpy<<<1, 32, 0, ctx->stream>>>(small + NS_RLZ, small + NS_RLZ, small + NS_RLZC, c_eval, 1);           // r_LZ likewise
    SP2_LAUNCH_CHECK();
  } else {
    // one commitment row: LZ is the folded witness itself (hyrax_pc.rs:420-424)
    const unsigned nbM = (unsigned)std::min<u64>((M + 255) / 256, (u64)ctx->num_sms * 8);
    k_nn_axpy<<<nbM, 256, 0, ctx->stream>>>(P->Wfin, P->Wfold, P->zc, c_eval, M);                     // W = W_fold + c_eval W_core
    SP2_LAUNCH_CHECK();
    LZp = P->Wfin; rLZ = blind_fin;
  }
  SP2_CUDA_OK(cudaStreamWaitEvent(ctx->stream, P->ev_side, 0));                                        // the folded-witness rows are committed
  std::vector<MsmJob> jobs(rows + 3);                              // (delta, point rows + 3 of the output, was committed on the side stream)
  for (uint32_t r = 0; r < rows; r++) {        // comm[r] = comm(W_fold row) + c_eval U_core[r] + (b_fold[r] + c_eval b_core[r]) h
    MsmJob &j = jobs[r]; memset(&j, 0, sizeof(j));
    j.nextra = 2; j.extra_base[0] = r; j.extra_scalar[0] = small + NS_CEVAL; j.extra_tab[0] = P->Ucore_tab;
    j.extra_base[1] = ck->idx_h(); j.extra_scalar[1] = blind_fin + r;
    if (r < pre_rows) j.add_jac = pts_fold + r;
  }
  { MsmJob &j = jobs[rows]; memset(&j, 0, sizeof(j)); j.nextra = 2; j.extra_base[0] = ck->idx_ck_s(); j.extra_scalar[0] = small + NS_EVALF;
    j.extra_base[1] = ck->idx_h_s(); j.extra_scalar[1] = small + NS_BEVALF; }                           // comm_eval
  { MsmJob &j = jobs[rows + 1]; memset(&j, 0, sizeof(j)); j.scalars = LZp; j.len = (u32)width; j.nextra = 1; j.extra_base[0] = ck->idx_h(); j.extra_scalar[0] = rLZ; }   // comm_LZ
  { MsmJob &j = jobs[rows + 2]; memset(&j, 0, sizeof(j)); j.nextra = 2; j.extra_base[0] = ck->idx_ck_s(); j.extra_scalar[0] = small + NS_IP;
    j.extra_base[1] = ck->idx_h_s(); j.extra_scalar[1] = small + NS_RBETA; }                            // beta
  jac *pts_out = pts_fold + rows + 2;            // (the row jobs read pts_fold[0..pre_rows) as addends and write pts_out[0..rows): disjoint)
  SP2_TRY(msm_run(ctx, ck, jobs, pts_out));
  SP2_CUDA_OK(cudaMemcpyAsync(h_jac, pts_out, (rows + 4) * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
  fe *h_small = (fe *)h_in;
  SP2_CUDA_OK(cudaMemcpyAsync(h_small, small, NS_COUNT * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  if (nvr == 0) SP2_CUDA_OK(cudaMemcpyAsync(h_small + NS_RLZ, blind_fin, sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  std::vector<uint64_t> outp((size_t)(rows + 4) * 8);
  sp2h::batch_normalize(h_jac, rows + 4, outp.data());
  const uint64_t *comm = outp.data(), *comm_eval = comm + 8 * (size_t)rows, *comm_LZ = nvr > 0 ? comm_eval + 8 : comm, *p_beta = comm_eval + 16, *p_delta = comm_eval + 24;
  if (sn->comm_fold) memcpy(sn->comm_fold, comm, (size_t)rows * 64);
  memcpy(sn->delta, p_delta, 64); memcpy(sn->beta, p_beta, 64);
  ts.absorb_commitment("poly_com", comm, rows);                                                         // hyrax_pc.rs:410
  ts.dom_sep("inner product argument (linear)");                                                        // ipa.rs:134-153
  ts.push("U", 1); ts.push_point(comm_LZ); ts.push_point(comm_eval);
  ts.absorb_point("delta", p_delta); ts.absorb_point("beta", p_beta);
  fe r_ipa; { uint8_t dg[64]; uint64_t o[4]; ts.squeeze("r", dg); sp2h::fq_from_uniform(dg, o); r_ipa = H::load(o); }
  k_nn_axpy<<<(unsigned)((width + 255) / 256), 256, 0, ctx->stream>>>(zvec, dvec, LZp, r_ipa, width);   // z_vec = d + r LZ   (ipa.rs:155-167)
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(sn->z_vec, zvec, width * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  H::store(sn->z_delta, HF::add(HF::mul(r_ipa, h_small[NS_RLZ]), H::load(rnd->r_delta)));
  H::store(sn->z_beta, HF::add(HF::mul(r_ipa, be_f), H::load(rnd->r_beta)));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ph[8] = ms_since(t_phase);
  ph[9] = ms_since(t_begin);
  if (phase_ms) memcpy(phase_ms, ph, sizeof(ph));
  return SP2_OK;
}

int32_t sp2_neutronnova_snark_prove(sp2_ctx *ctx, sp2_nn_prep *P, const uint8_t *vk_digest, const sp2_nn_rand *rnd, sp2_nn_snark *sn, float *phase_ms) {
  if (P && P->nranks > 1) return set_error(ctx, SP2_ERR_INTERNAL, "snark_prove: sharded prep state: use sp2_neutronnova_snark_prove_sharded");
  return nn_snark_impl(ctx, P, nullptr, nullptr, nullptr, vk_digest, rnd, sn, phase_ms);
}
/* Instance-sharded (one process per GPU; see sp2_neutronnova_prove_sharded): every rank passes the SAME rand (blinds of all
 * n_local * nranks instances) and receives the identical proof; a rank rerandomises / commits only its own instances and the rows
 * are all-gathered through `allgather` (host buffers, on_device = 0). */
int32_t sp2_neutronnova_snark_prove_sharded(sp2_ctx *ctx, sp2_nn_prep *P, sp2_comm *comm, sp2_allgather_fn allgather, void *user, const uint8_t *vk_digest,
                                            const sp2_nn_rand *rnd, sp2_nn_snark *sn, float *phase_ms) {
  return nn_snark_impl(ctx, P, comm, allgather, user, vk_digest, rnd, sn, phase_ms);
}

}  // extern "C"
