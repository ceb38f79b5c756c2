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
#include <string.h>
#include <vector>
#include "ctx.cuh"
#include "curve.cuh"
#include "devutil.cuh"
#include "host_transcript.h"
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

// One NIFS round: grid (i-chunks, pairs); thread owns j (the contiguous index), walks its i-chunk:
//   acc += f[i] * v(i, j), then * e_left[j]   (prove_helper's nested sums with inner/outer swapped: one flush per thread)
__global__ void __launch_bounds__(NF_THREADS) k_nifs_round(u32 t, const fe *rhos, u32 ell_b, u32 left, u32 right, const fe *E, const fe *A,
                                                           const fe *B, const fe *C, u64 N, u64 stride, fe *partials) {
  __shared__ fe red[2 * 32];
  __shared__ fe wsh;
  const u32 p = blockIdx.y;
  const fe *el = E, *f = E + left;
  const fe *A1 = A + (u64)(2 * p) * stride * N, *A2 = A + (u64)(2 * p + 1) * stride * N;
  const fe *B1 = B + (u64)(2 * p) * stride * N, *B2 = B + (u64)(2 * p + 1) * stride * N;
  const fe *C1 = C + (u64)(2 * p) * stride * N;
  if (threadIdx.x == 0) {                       // suffix_weight_full(t, ell_b, p, rhos)
    fe w = Fq::one(); u32 k = p;
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

/* HyraxPCS::fold_commitments as group elements: out[row] = sum_i w[i] * comms[i*rows + row] (affine in/out, host) */
int32_t sp2_fold_commitments(sp2_ctx *ctx, const uint64_t *comms_xy, uint32_t n, uint32_t rows, const uint64_t *w, uint64_t *out_xy) {
  cudaSetDevice(ctx->device);
  if (!n || !rows) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "fold_commitments: empty input");
  void *dp, *dt, *dout; fe *dw;
  const size_t tot = (size_t)n * rows;
  SP2_TRY(scratch(ctx, 2, tot * sizeof(aff), &dp)); SP2_TRY(scratch(ctx, 3, tot * sizeof(jac), &dt)); SP2_TRY(scratch(ctx, 4, (size_t)rows * sizeof(jac), &dout));
  SP2_TRY(upload_small(ctx, 0, w, n, &dw));
  SP2_CUDA_OK(cudaMemcpyAsync(dp, comms_xy, tot * sizeof(aff), cudaMemcpyHostToDevice, ctx->stream));
  k_scalar_mul_var<<<(unsigned)((tot + 63) / 64), 64, 0, ctx->stream>>>((const aff *)dp, dw, n, rows, (jac *)dt);
  SP2_LAUNCH_CHECK();
  k_point_col_sum<<<rows, 128, 0, ctx->stream>>>((const jac *)dt, n, rows, (jac *)dout);
  SP2_LAUNCH_CHECK();
  std::vector<uint64_t> hj((size_t)rows * 12);
  SP2_CUDA_OK(cudaMemcpyAsync(hj.data(), dout, (size_t)rows * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  sp2h::batch_normalize(hj.data(), rows, out_xy);
  return SP2_OK;
}

}  // extern "C"
