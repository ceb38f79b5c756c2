// Small-value path on the device: i64 layers for values that fit 62 bits, signed 448-bit
// accumulation of field x i128 products, and the NIFS round-0 / c_vals kernels that consume them.
//
// Restates reference
//   src/big_num/small_value.rs:32-85     SMALL_VALUE_MAX, to_small_vec_or_zero (value -> i64 or 0 + recorded position)
//   src/big_num/small_value.rs:96-196    SmallAccumulator (pos / neg buckets of field_mont * |i128|, 7 limbs each)
//   src/big_num/small_value.rs:204-222   reduce_7_to_field (acc mod p read as Montgomery limbs)
//   src/neutronnova_zk.rs:255-325        prove_helper_small (round 0 of the NIFS on i64 layers + field correction)
//   src/neutronnova_zk.rs:649-693        c_vals
//   src/neutronnova_zk.rs:1551-1584      prep_prove: conversion of every layer, union of the large positions, zeroing
//
// B200 design: a layer entry is 8 bytes instead of 32 and a product-accumulate is a 256 x 128-bit multiply with no
// modular reduction, so the round-0 kernel moves 1/4 of the bytes and issues ~1/4 of the IMADs of the field path.
// The thread that owns the contiguous index j walks i (as k_nifs_round does) with ONE signed accumulator pair,
// weights f[i] inside and e_left[j] outside — the reference's nested sums with the roles swapped; the field value is
// the same, hence bit-identical.  reduce_7 needs no long division here: acc = lo + hi * 2^256 with hi < 2^192, and
// hi * 2^256 mod p is one Montgomery multiplication of the raw integer hi by R^2.
#include <algorithm>
#include "ctx.cuh"
#include "devutil.cuh"

using namespace sp2;

namespace {

constexpr int SV_THREADS = 256;
typedef long long i64;
typedef unsigned __int128 u128d;

struct SmallAcc { u64 pos[7], neg[7]; };
__device__ __forceinline__ void small_acc_zero(SmallAcc &a) {
#pragma unroll
  for (int i = 0; i < 7; i++) { a.pos[i] = 0; a.neg[i] = 0; }
}
// t += f * v  (f: 4 x u64 Montgomery limbs, v < 2^127); small_value.rs:120-164
__device__ __forceinline__ void small_mad(u64 (&t)[7], const u64 (&f)[4], u128d v) {
  const u64 lo = (u64)v, hi = (u64)(v >> 64);
  u128d c = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) { const u128d p = (u128d)f[j] * lo + t[j] + c; t[j] = (u64)p; c = p >> 64; }
#pragma unroll
  for (int j = 4; j < 7; j++) { const u128d s = (u128d)t[j] + c; t[j] = (u64)s; c = s >> 64; }
  if (hi) {
    c = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { const u128d p = (u128d)f[j] * hi + t[j + 1] + c; t[j + 1] = (u64)p; c = p >> 64; }
#pragma unroll
    for (int j = 5; j < 7; j++) { const u128d s = (u128d)t[j] + c; t[j] = (u64)s; c = s >> 64; }
  }
}
__device__ __forceinline__ void small_accumulate(SmallAcc &a, const fe &f, __int128 val) {
  if (val == 0) return;
  u64 fl[4];
#pragma unroll
  for (int i = 0; i < 4; i++) fl[i] = (u64)f.v[2 * i] | ((u64)f.v[2 * i + 1] << 32);
  if (val > 0) small_mad(a.pos, fl, (u128d)val); else small_mad(a.neg, fl, (u128d)(-val));
}
// reduce_7_to_field: acc mod p as a (Montgomery-form) field element
__device__ __forceinline__ fe small_reduce7(const u64 (&acc)[7]) {
  fe lo, hi;
#pragma unroll
  for (int i = 0; i < 4; i++) { lo.v[2 * i] = (u32)acc[i]; lo.v[2 * i + 1] = (u32)(acc[i] >> 32); }
#pragma unroll
  for (int i = 0; i < 3; i++) { hi.v[2 * i] = (u32)acc[4 + i]; hi.v[2 * i + 1] = (u32)(acc[4 + i] >> 32); }
  hi.v[6] = 0; hi.v[7] = 0;
  cond_sub_p<FqParams>(lo, 0);                                  // lo < 2^256 < 2p
  if (Fq::is_zero(hi)) return lo;
  return Fq::add(lo, Fq::mul(hi, Fq::cst_r2()));               // hi * R^2 / R = hi * 2^256 (mod p)
}
__device__ __forceinline__ fe small_acc_reduce(const SmallAcc &a) { return Fq::sub(small_reduce7(a.pos), small_reduce7(a.neg)); }

// to_small_vec_or_zero over n_layers layers of N entries; flags[k] |= 1 where ANY layer is large at k
__global__ void __launch_bounds__(SV_THREADS) k_to_small(const fe *layers, u64 total, u64 N, i64 *out, unsigned char *flags) {
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (u64)gridDim.x * blockDim.x) {
    const fe c = Fq::from_mont(ldg_fe(layers + q));            // to_repr(): canonical
    const u64 l0 = (u64)c.v[0] | ((u64)c.v[1] << 32);
    const u32 up = c.v[2] | c.v[3] | c.v[4] | c.v[5] | c.v[6] | c.v[7];
    const u64 SMAX = (1ull << 62) - 1;
    i64 r = 0; bool large = true;
    if (up == 0 && l0 <= SMAX) { r = (i64)l0; large = false; }
    else {
      const fe d = Fq::sub(Fq::zero(), c);                     // p - value (value != 0 here)
      const u64 d0 = (u64)d.v[0] | ((u64)d.v[1] << 32);
      const u32 dup = d.v[2] | d.v[3] | d.v[4] | d.v[5] | d.v[6] | d.v[7];
      if (dup == 0 && d0 > 0 && d0 <= SMAX) { r = -(i64)d0; large = false; }
    }
    out[q] = r;
    if (large) flags[q % N] = 1;
  }
}
// zero every table at the union of the large positions (neutronnova_zk.rs:1575-1584)
struct SvTables { i64 *t[4]; };
__global__ void __launch_bounds__(SV_THREADS) k_zero_large(SvTables tb, u32 ntab, u64 total, u64 N, const unsigned char *flags) {
  for (u64 q = (u64)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (u64)gridDim.x * blockDim.x)
    if (flags[q % N]) for (u32 s = 0; s < ntab; s++) tb.t[s][q] = 0;
}
// ascending list of the flagged positions (one CTA; thread t owns a contiguous segment)
__global__ void __launch_bounds__(1024) k_compact_flags(const unsigned char *flags, u64 N, u64 *positions, u64 *count) {
  __shared__ u32 cnt[1024];
  const u64 seg = (N + 1023) / 1024, s = threadIdx.x * seg, e = min(N, s + seg);
  u32 c = 0;
  for (u64 k = s; k < e; k++) c += flags[k] != 0;
  cnt[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) { u32 run = 0; for (int t = 0; t < 1024; t++) { const u32 v = cnt[t]; cnt[t] = run; run += v; } *count = run; }
  __syncthreads();
  u64 o = cnt[threadIdx.x];
  for (u64 k = s; k < e; k++) if (flags[k]) positions[o++] = k;
}

__device__ __forceinline__ fe suffix_weight0(u32 ell_b, u32 p, const fe *rhos) {      // suffix_weight_full(0, ell_b, p, rhos); p: GLOBAL pair index
  fe w = Fq::one(); u32 k = p;
  for (u32 s = 1; s < ell_b; s++) { const fe r = ldg_fe(rhos + s); w = Fq::mul(w, (k & 1u) ? r : Fq::sub(Fq::one(), r)); k >>= 1; }
  return w;
}
// round 0 of the NIFS on i64 layers: grid (i-chunks, pairs); partial[pair * chunks + chunk] = w_pair * sum
__global__ void __launch_bounds__(SV_THREADS) k_nifs_round0_small(u32 ell_b, const fe *rhos, u32 left, u32 right, const fe *E, const i64 *A64,
                                                                  const i64 *B64, u64 N, fe *partials, u32 pair_offset) {
  __shared__ fe red[32];
  __shared__ fe wsh;
  const u32 p = blockIdx.y;
  const fe *el = E, *f = E + left;
  const i64 *a1 = A64 + (u64)(2 * p) * N, *a2 = a1 + N, *b1 = B64 + (u64)(2 * p) * N, *b2 = b1 + N;
  if (threadIdx.x == 0) wsh = suffix_weight0(ell_b, p + pair_offset, rhos);
  fe x[1] = {Fq::zero()};
  const u32 per = (right + gridDim.x - 1) / gridDim.x, i0 = blockIdx.x * per, i1 = min(right, i0 + per);
  for (u32 j = threadIdx.x; j < left; j += blockDim.x) {
    SmallAcc acc; small_acc_zero(acc);
    for (u32 i = i0; i < i1; i++) {
      const u64 k = (u64)i * left + j;
      const __int128 da = (__int128)a2[k] - (__int128)a1[k], db = (__int128)b2[k] - (__int128)b1[k];
      small_accumulate(acc, ldg_fe_ro(f + i), da * db);
    }
    x[0] = Fq::add(x[0], Fq::mul(ldg_fe_ro(el + j), small_acc_reduce(acc)));
  }
  block_sum_fq<1>(x, red);
  __syncthreads();
  if (threadIdx.x == 0) stg_fe(partials + (size_t)blockIdx.y * gridDim.x + blockIdx.x, Fq::mul(x[0], wsh));
}
// field correction at the zeroed positions: one CTA per pair
__global__ void __launch_bounds__(SV_THREADS) k_nifs_round0_fix(u32 ell_b, const fe *rhos, u32 left, u32 right, const fe *E, const fe *A, const fe *B,
                                                                const u64 *positions, u64 n_large, u64 N, fe *partials, u32 pair_offset) {
  __shared__ fe red[32];
  const u32 p = blockIdx.x;
  const fe *el = E, *f = E + left;
  const fe *A1 = A + (u64)(2 * p) * N, *A2 = A1 + N, *B1 = B + (u64)(2 * p) * N, *B2 = B1 + N;
  fe x[1] = {Fq::zero()};
  for (u64 q = threadIdx.x; q < n_large; q += blockDim.x) {
    const u64 k = positions[q];
    if (k >= (u64)left * right) continue;
    const fe da = Fq::sub(ldg_fe(A2 + k), ldg_fe(A1 + k)), db = Fq::sub(ldg_fe(B2 + k), ldg_fe(B1 + k));
    x[0] = Fq::add(x[0], Fq::mul(Fq::mul(Fq::mul(ldg_fe_ro(f + k / left), ldg_fe_ro(el + k % left)), da), db));
  }
  block_sum_fq<1>(x, red);
  __syncthreads();
  if (threadIdx.x == 0) stg_fe(partials + p, Fq::mul(x[0], suffix_weight0(ell_b, p + pair_offset, rhos)));
}
// c_vals[b] = sum_k E[k] * Cz_b[k]: one CTA per instance
__global__ void __launch_bounds__(SV_THREADS) k_nifs_cvals_small(u32 left, u32 right, const fe *E, const fe *Cl, const i64 *C64, const u64 *positions,
                                                                 u64 n_large, u64 N, fe *vals) {
  __shared__ fe red[32];
  const u32 b = blockIdx.x;
  const fe *el = E, *f = E + left;
  const i64 *c = C64 + (u64)b * N;
  fe x[1] = {Fq::zero()};
  for (u32 j = threadIdx.x; j < left; j += blockDim.x) {
    SmallAcc acc; small_acc_zero(acc);
    for (u32 i = 0; i < right; i++) small_accumulate(acc, ldg_fe_ro(f + i), (__int128)c[(u64)i * left + j]);
    x[0] = Fq::add(x[0], Fq::mul(ldg_fe_ro(el + j), small_acc_reduce(acc)));
  }
  for (u64 q = threadIdx.x; q < n_large; q += blockDim.x) {
    const u64 k = positions[q];
    if (k >= (u64)left * right) continue;
    x[0] = Fq::add(x[0], Fq::mul(Fq::mul(ldg_fe_ro(el + k % left), ldg_fe_ro(f + k / left)), ldg_fe(Cl + (u64)b * N + k)));
  }
  block_sum_fq<1>(x, red);
  if (threadIdx.x == 0) stg_fe(vals + b, x[0]);
}
__global__ void __launch_bounds__(SV_THREADS) k_sum_partials1(const fe *partials, u32 nparts, fe *out) {
  __shared__ fe red[32];
  fe x[1] = {Fq::zero()};
  for (u32 b = threadIdx.x; b < nparts; b += blockDim.x) x[0] = Fq::add(x[0], ldg_fe(partials + b));
  block_sum_fq<1>(x, red);
  if (threadIdx.x == 0) stg_fe(out, x[0]);
}

}  // namespace

namespace sp2 {
// quad coefficient of NIFS round 0 from the i64 layers -> *d_out (device); positions may be null when n_large == 0
int nifs_round0_small_enqueue(sp2_ctx *ctx, const fe *d_rhos, u32 ell_b, u32 left, u32 right, const fe *dE, const void *dA64, const void *dB64,
                              const fe *dA, const fe *dB, const void *d_positions, u64 n_large, u64 N, u64 m, fe *d_partials, fe *d_out, u32 pair_offset) {
  const u32 pairs = (u32)(m / 2);
  u32 chunks = std::max<u32>(1, std::min<u32>(right, (u32)(ctx->num_sms * 4) / std::max<u32>(1, pairs)));
  const u32 threads = std::min<u32>(SV_THREADS, (left + 31) / 32 * 32);
  k_nifs_round0_small<<<dim3(chunks, pairs), threads, 0, ctx->stream>>>(ell_b, d_rhos, left, right, dE, (const i64 *)dA64, (const i64 *)dB64, N, d_partials, pair_offset);
  SP2_LAUNCH_CHECK();
  u32 nparts = chunks * pairs;
  if (n_large) {
    k_nifs_round0_fix<<<pairs, SV_THREADS, 0, ctx->stream>>>(ell_b, d_rhos, left, right, dE, dA, dB, (const u64 *)d_positions, n_large, N, d_partials + nparts, pair_offset);
    SP2_LAUNCH_CHECK();
    nparts += pairs;
  }
  k_sum_partials1<<<1, SV_THREADS, 0, ctx->stream>>>(d_partials, nparts, d_out);
  SP2_LAUNCH_CHECK();
  return SP2_OK;
}
// convert ntab tables of n_layers x N field elements; d_flags (N bytes) and d_positions (N u64) are outputs
int small_layers_enqueue(sp2_ctx *ctx, const fe *const *d_tabs, void *const *d_i64, u32 ntab, u64 n_layers, u64 N, void *d_flags, void *d_positions,
                         u64 *d_count) {
  const u64 total = n_layers * N;
  SP2_CUDA_OK(cudaMemsetAsync(d_flags, 0, N, ctx->stream));
  const unsigned nb = (unsigned)std::min<u64>((total + SV_THREADS - 1) / SV_THREADS, (u64)ctx->num_sms * 8);
  SvTables tb; for (int s = 0; s < 4; s++) tb.t[s] = s < (int)ntab ? (i64 *)d_i64[s] : nullptr;
  for (u32 s = 0; s < ntab; s++) {
    k_to_small<<<nb, SV_THREADS, 0, ctx->stream>>>(d_tabs[s], total, N, (i64 *)d_i64[s], (unsigned char *)d_flags);
    SP2_LAUNCH_CHECK();
  }
  k_zero_large<<<nb, SV_THREADS, 0, ctx->stream>>>(tb, ntab, total, N, (const unsigned char *)d_flags);
  SP2_LAUNCH_CHECK();
  k_compact_flags<<<1, 1024, 0, ctx->stream>>>((const unsigned char *)d_flags, N, (u64 *)d_positions, d_count);
  SP2_LAUNCH_CHECK();
  return SP2_OK;
}
}  // namespace sp2

extern "C" {

/* to_small_vec_or_zero for up to 4 device tables of n_layers x N scalars each (e.g. the Az, Bz, Cz layers), with the
 * union of the large positions over ALL layers of ALL tables zeroed in every i64 table and returned ascending
 * (prep_prove, neutronnova_zk.rs:1551-1584).  d_i64[s]: n_layers * N int64; d_positions: N uint64 (device). */
int32_t sp2_to_small_layers_dev(sp2_ctx *ctx, const void *const *d_tables, void *const *d_i64, uint32_t ntables, uint64_t n_layers, uint64_t N,
                                void *d_positions, uint64_t *n_large) {
  cudaSetDevice(ctx->device);
  if (ntables < 1 || ntables > 4 || !N || !n_layers) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "to_small_layers: 1..4 non-empty tables");
  void *flags; SP2_TRY(scratch(ctx, 0, ((N + 15) & ~(uint64_t)15) + 64, &flags));
  u64 *d_count = (u64 *)((unsigned char *)flags + ((N + 15) & ~(uint64_t)15));
  const fe *tabs[4]; for (uint32_t s = 0; s < 4; s++) tabs[s] = s < ntables ? (const fe *)d_tables[s] : nullptr;
  SP2_TRY(small_layers_enqueue(ctx, tabs, d_i64, ntables, n_layers, N, flags, d_positions, d_count));
  SP2_CUDA_OK(cudaMemcpyAsync(n_large, d_count, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

/* NIFS round 0 on the i64 layers (prove_helper_small per pair, suffix weights, summed over pairs): out2 = (0, quad).
 * m live layers of N = left*right entries, layer-major; dA/dB: the field layers (used only at the large positions). */
int32_t sp2_nifs_round0_small_dev(sp2_ctx *ctx, const uint64_t *rhos, uint32_t ell_b, uint32_t left, uint32_t right, const void *dE, const void *dA64,
                                  const void *dB64, const void *dA, const void *dB, const void *d_positions, uint64_t n_large, uint64_t N, uint64_t m,
                                  uint64_t *out2) {
  cudaSetDevice(ctx->device);
  if ((uint64_t)left * right != N || m < 2 || (m & 1)) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "nifs_round0_small: bad shape");
  void *dr, *part;
  SP2_TRY(scratch(ctx, 0, (size_t)(ell_b + 1) * sizeof(fe), &dr));
  SP2_CUDA_OK(cudaMemcpyAsync(dr, rhos, (size_t)std::max<uint32_t>(ell_b, 1) * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_TRY(scratch(ctx, 1, ((size_t)ctx->num_sms * 4 + m + 8) * sizeof(fe), &part));
  fe *d_out = (fe *)part + (size_t)ctx->num_sms * 4 + m;
  SP2_TRY(nifs_round0_small_enqueue(ctx, (const fe *)dr, ell_b, left, right, (const fe *)dE, dA64, dB64, (const fe *)dA, (const fe *)dB, d_positions, n_large,
                                    N, m, (fe *)part, d_out, 0));
  memset(out2, 0, sizeof(fe));
  SP2_CUDA_OK(cudaMemcpyAsync(out2 + 4, d_out, sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

/* c_vals[b] = sum_k E[k] * Cz_b[k] for n instances (neutronnova_zk.rs:649-693), host out */
int32_t sp2_nifs_cvals_small_dev(sp2_ctx *ctx, uint32_t left, uint32_t right, const void *dE, const void *dC, const void *dC64, const void *d_positions,
                                 uint64_t n_large, uint64_t N, uint64_t n, uint64_t *out_vals) {
  cudaSetDevice(ctx->device);
  if ((uint64_t)left * right != N || !n) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "nifs_cvals_small: bad shape");
  void *dv; SP2_TRY(scratch(ctx, 1, (size_t)n * sizeof(fe) + 64, &dv));
  const u32 threads = std::min<u32>(SV_THREADS, (left + 31) / 32 * 32);
  k_nifs_cvals_small<<<(unsigned)n, threads, 0, ctx->stream>>>(left, right, (const fe *)dE, (const fe *)dC, (const i64 *)dC64, (const u64 *)d_positions, n_large, N,
                                                               (fe *)dv);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(out_vals, dv, (size_t)n * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

}  // extern "C"
