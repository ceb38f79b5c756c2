// Multi-scalar multiplication and Hyrax row commitments on the device.
//
// Restates (as group elements — the affine results are what the reference emits and absorbs):
//   src/provider/msm.rs:59-222   cpu_msm_serial / msm  (signed-digit Pippenger, c = 8 at n = 2048)
//   src/provider/msm.rs:367-620  msm_small family (binary / <=10 bit / windowed) — same sums
//   src/provider/msm.rs:637-774  FixedBaseMul (8-bit window tables) — blind * h terms
//   src/provider/pcs/hyrax_pc.rs:207-319  HyraxPCS::commit / commit_zeros (one Pedersen row per 2048)
//   src/provider/pcs/hyrax_pc.rs:38-54    bind_with_delayed (LZ = L^T W)
//
// B200 design.  The commitment key is fixed at setup, so (like the reference's FixedBaseMul, but for
// every base) the key upload precomputes table[w][i] = 2^(8w) * base_i in affine form: all 33 signed
// byte-digit windows of all scalars then fall into ONE set of 128 buckets and the window-combining
// Horner chain (256 serial doublings in the reference's loop, msm.rs:151-175) disappears.  Per MSM:
//   accumulate: CTAs of 128 threads own 64 terms each; digits are counting-sorted by bucket in shared
//               memory so that thread b walks only bucket b's entries (mixed adds, no atomics);
//   reduce:     one CTA sums the per-CTA partial buckets, then sum_b b*B_b via 8 bit-plane tree sums
//               and a 7-step Horner (instead of the 2*128 serial running-sum adds), then one inversion
//               to affine.
// Many MSMs (Hyrax rows, the prover's blinded terms) are batched into one pair of launches.
#include <string.h>
#include "msm.cuh"
#include "devutil.cuh"

using namespace sp2;

namespace {

__device__ __forceinline__ jac ld_jac(const jac *p) { jac r; r.x = ldg_fe(&p->x); r.y = ldg_fe(&p->y); r.z = ldg_fe(&p->z); return r; }
__device__ __forceinline__ void st_jac(jac *p, const jac &v) { stg_fe(&p->x, v.x); stg_fe(&p->y, v.y); stg_fe(&p->z, v.z); }
__device__ __forceinline__ aff ld_aff_ro(const aff *p) { aff r; r.x = ldg_fe_ro(&p->x); r.y = ldg_fe_ro(&p->y); return r; }

// ---- key precompute ----------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_ck_windows(const aff *bases, u32 nbase, jac *tmp) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbase) return;
  jac p = jac_from_aff(ld_aff_ro(bases + i));
  for (int w = 0; w < MSM_NW; w++) {
    st_jac(tmp + (size_t)w * nbase + i, p);
#pragma unroll 1
    for (int k = 0; k < MSM_C; k++) p = jac_dbl(p);
  }
}
// batch normalisation (Montgomery's trick over the 33 windows of one base; cf. batch_normalize, msm.rs:669-676)
__global__ void __launch_bounds__(128) k_ck_normalize(const jac *tmp, u32 nbase, aff *table) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nbase) return;
  fe pre[MSM_NW];
  fe acc = Fp::one();
  for (int w = 0; w < MSM_NW; w++) {
    pre[w] = acc;
    const fe z = ldg_fe(&tmp[(size_t)w * nbase + i].z);
    if (!Fp::is_zero(z)) acc = Fp::mul(acc, z);
  }
  fe inv = Fp::inv(acc);
  for (int w = MSM_NW - 1; w >= 0; w--) {
    const jac p = ld_jac(tmp + (size_t)w * nbase + i);
    aff a;
    if (Fp::is_zero(p.z)) { a.x = Fp::zero(); a.y = Fp::zero(); }
    else { a = jac_to_aff_with_inv(p, Fp::mul(inv, pre[w])); inv = Fp::mul(inv, p.z); }
    stg_fe(&table[(size_t)w * nbase + i].x, a.x);
    stg_fe(&table[(size_t)w * nbase + i].y, a.y);
  }
}

// ---- accumulate --------------------------------------------------------------------------------
struct AccSmem {
  short dig[MSM_SLICE][MSM_NW + 1];
  u32 bidx[MSM_SLICE];
  u32 hist[MSM_NBUCKET + 2];
  u32 cur[MSM_NBUCKET + 2];
  unsigned short sorted[MSM_SLICE * MSM_NW];
};

__global__ void __launch_bounds__(MSM_THREADS) k_msm_accumulate(const MsmJob *jobs, const aff *table, u32 nbase, jac *partial, u32 maxblk) {
  __shared__ AccSmem sm;
  const MsmJob job = jobs[blockIdx.y];
  if (blockIdx.x >= job.nblk) return;
  const u32 tid = threadIdx.x;
  const u32 total = job.len + job.nextra, t0 = blockIdx.x * MSM_SLICE;
  const u32 nterm = min((u32)MSM_SLICE, total - t0);
  for (u32 i = tid; i < MSM_NBUCKET + 2; i += blockDim.x) sm.hist[i] = 0;
  // signed byte digits of each scalar (to_repr() little-endian bytes, msm.rs:97-100; digit recoding :122-148)
  if (tid < nterm) {
    const u32 g = t0 + tid;
    fe s; u32 b;
    if (g < job.len) { s = ldg_fe(job.scalars + g); b = job.base0 + g; }
    else { s = ldg_fe(job.extra_scalar[g - job.len]); b = job.extra_base[g - job.len]; }
    s = Fq::from_mont(s);
    sm.bidx[tid] = b;
    int carry = 0;
#pragma unroll
    for (int w = 0; w < 32; w++) {
      int d = (int)((s.v[w >> 2] >> (8 * (w & 3))) & 0xffu) + carry;
      carry = 0;
      if (d > 128) { d -= 256; carry = 1; }
      sm.dig[tid][w] = (short)d;
    }
    sm.dig[tid][32] = (short)carry;
  }
  __syncthreads();
  if (tid < nterm) {
    for (int w = 0; w < MSM_NW; w++) { const int d = sm.dig[tid][w]; if (d) atomicAdd(&sm.hist[d < 0 ? -d : d], 1u); }
  }
  __syncthreads();
  if (tid == 0) { u32 run = 0; for (int b = 1; b <= MSM_NBUCKET + 1; b++) { const u32 c = sm.hist[b]; sm.hist[b] = run; sm.cur[b] = run; run += c; } }
  __syncthreads();
  if (tid < nterm) {
    for (int w = 0; w < MSM_NW; w++) {
      const int d = sm.dig[tid][w];
      if (d) { const u32 pos = atomicAdd(&sm.cur[d < 0 ? -d : d], 1u); sm.sorted[pos] = (unsigned short)(tid | (w << 6) | (d < 0 ? 0x8000 : 0)); }
    }
  }
  __syncthreads();
  // thread `tid` owns bucket tid + 1
  jac acc = jac_inf();
  const u32 s = sm.hist[tid + 1], e = sm.hist[tid + 2];
  for (u32 k = s; k < e; k++) {
    const u32 code = sm.sorted[k];
    const u32 t = code & 63u, w = (code >> 6) & 63u;
    aff p = ld_aff_ro(table + (size_t)w * nbase + sm.bidx[t]);
    if (code & 0x8000u) p.y = Fp::neg(p.y);
    acc = jac_add_mixed(acc, p);
  }
  st_jac(partial + ((size_t)blockIdx.y * maxblk + blockIdx.x) * MSM_NBUCKET + tid, acc);
}

// ---- reduce ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MSM_RED_THREADS) k_msm_reduce(const MsmJob *jobs, const jac *partial, u32 maxblk, aff *out) {
  extern __shared__ __align__(32) unsigned char smem_raw[];
  jac *sm = (jac *)smem_raw;                       // MSM_RED_THREADS entries
  __shared__ jac G[8];
  const MsmJob job = jobs[blockIdx.x];
  const u32 tid = threadIdx.x, b = tid & (MSM_NBUCKET - 1), part = tid >> 7;   // 4 parts
  const jac *base = partial + (size_t)blockIdx.x * maxblk * MSM_NBUCKET;
  jac acc = jac_inf();
  for (u32 blk = part; blk < job.nblk; blk += MSM_RED_THREADS / MSM_NBUCKET) acc = jac_add(acc, ld_jac(base + (size_t)blk * MSM_NBUCKET + b));
  sm[tid] = acc;
  __syncthreads();
  for (u32 s = 2; s >= 1; s >>= 1) {
    if (part < s) sm[tid] = jac_add(sm[tid], sm[tid + s * MSM_NBUCKET]);
    __syncthreads();
  }
  const jac Bb = sm[b];                            // bucket b + 1
  __syncthreads();
  // sum_b (b+1) * B_b = sum_j 2^j * (sum over buckets whose index has bit j set)
  for (int pass = 0; pass < 2; pass++) {
    const int j = pass * 4 + (int)part;
    sm[tid] = (((b + 1) >> j) & 1u) ? Bb : jac_inf();
    __syncthreads();
    for (u32 s = MSM_NBUCKET / 2; s >= 1; s >>= 1) {
      if (b < s) sm[tid] = jac_add(sm[tid], sm[tid + s]);
      __syncthreads();
    }
    if (b == 0) G[j] = sm[tid];
    __syncthreads();
  }
  if (tid == 0) {
    jac r = G[7];
    for (int j = 6; j >= 0; j--) r = jac_add(jac_dbl(r), G[j]);
    const aff a = jac_to_aff(r);
    stg_fe(&out[blockIdx.x].x, a.x);
    stg_fe(&out[blockIdx.x].y, a.y);
  }
}

// ---- Hyrax bind: LZ[i] = sum_j L[j] * W[j * r_len + i], delayed reduction (hyrax_pc.rs:38-54) ----
constexpr int BIND_RG = 32;
__global__ void __launch_bounds__(128) k_hyrax_bind(const fe *poly, const fe *L, u64 rows, u64 r_len, fe *partial) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r_len) return;
  Fq::acc acc = Fq::acc_zero();
  for (u64 j = blockIdx.y; j < rows; j += gridDim.y) Fq::mul_acc(acc, ldg_fe_ro(L + j), ldg_fe(poly + j * r_len + i));
  stg_fe(partial + (u64)blockIdx.y * r_len + i, Fq::acc_reduce(acc));
}
__global__ void __launch_bounds__(128) k_sum_rows(const fe *partial, u64 nparts, u64 r_len, fe *out) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r_len) return;
  fe s = ldg_fe(partial + i);
  for (u64 p = 1; p < nparts; p++) s = Fq::add(s, ldg_fe(partial + p * r_len + i));
  stg_fe(out + i, s);
}


// ---- harness support: n pseudo-random curve points s_i * G (no hash-to-curve on this side of the ABI; the
// reference derives its generators in halo2curves, src/provider/traits.rs:205-249, and ships them as bases) ----
__device__ __forceinline__ u64 splitmix64(u64 &x) { x += 0x9E3779B97F4A7C15ull; u64 z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
__global__ void __launch_bounds__(64) k_test_points(u64 seed, u32 n, aff *out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // generator (3, 0x5a6dd32d...f02d) of T256 (verified there)
  fe gx, gy;
  const u32 GY[8] = {0x25b1f02du, 0x30b49549u, 0xa351bb3cu, 0xecd9d538u, 0xbe66600du, 0x4e97345cu, 0xf58708e6u, 0x5a6dd32du};
#pragma unroll
  for (int k = 0; k < 8; k++) { gx.v[k] = k == 0 ? 3u : 0u; gy.v[k] = GY[k]; }
  aff g; g.x = Fp::to_mont(gx); g.y = Fp::to_mont(gy);
  u64 st = seed ^ (0xD1B54A32D192ED03ull * (u64)(i + 1));
  u32 k[8];
#pragma unroll
  for (int j = 0; j < 4; j++) { const u64 w = splitmix64(st); k[2 * j] = (u32)w; k[2 * j + 1] = (u32)(w >> 32); }
  k[7] &= 0x7fffffffu; k[0] |= 1u;                   // 0 < k < group order
  jac acc = jac_inf();
  for (int b = 254; b >= 0; b--) {
    acc = jac_dbl(acc);
    if ((k[b >> 5] >> (b & 31)) & 1u) acc = jac_add_mixed(acc, g);
  }
  const aff a = jac_to_aff(acc);
  stg_fe(&out[i].x, a.x); stg_fe(&out[i].y, a.y);
}

bool g_reduce_attr_set = false;

}  // namespace

namespace sp2 {

int msm_run(sp2_ctx *ctx, const sp2_ck *ck, const std::vector<MsmJob> &jobs_in, aff *d_out) {
  if (jobs_in.empty()) return SP2_OK;
  const size_t CHUNK = 96;                         // bounds the partial-bucket scratch
  if (!g_reduce_attr_set) {
    SP2_CUDA_OK(cudaFuncSetAttribute(k_msm_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MSM_RED_THREADS * sizeof(jac))));
    g_reduce_attr_set = true;
  }
  for (size_t c0 = 0; c0 < jobs_in.size(); c0 += CHUNK) {
    const size_t nj = std::min(CHUNK, jobs_in.size() - c0);
    std::vector<MsmJob> jobs(jobs_in.begin() + c0, jobs_in.begin() + c0 + nj);
    u32 maxblk = 1;
    for (auto &j : jobs) {
      if (j.base0 + j.len > ck->nbase) return set_error(ctx, SP2_ERR_INVALID_CK_LENGTH, "msm: more scalars than commitment-key bases");
      j.nblk = (j.len + j.nextra + MSM_SLICE - 1) / MSM_SLICE;
      if (j.nblk == 0) j.nblk = 1;
      maxblk = std::max(maxblk, j.nblk);
    }
    void *d_jobs, *d_partial;
    SP2_TRY(scratch(ctx, 10, nj * sizeof(MsmJob), &d_jobs));
    SP2_TRY(scratch(ctx, 11, nj * maxblk * MSM_NBUCKET * sizeof(jac), &d_partial));
    // the job list is consumed asynchronously: stage it through a per-call copy (pageable -> the runtime
    // copies it out before cudaMemcpyAsync returns)
    SP2_CUDA_OK(cudaMemcpyAsync(d_jobs, jobs.data(), nj * sizeof(MsmJob), cudaMemcpyHostToDevice, ctx->stream));
    k_msm_accumulate<<<dim3(maxblk, (unsigned)nj), MSM_THREADS, 0, ctx->stream>>>((const MsmJob *)d_jobs, ck->table, ck->nbase, (jac *)d_partial, maxblk);
    SP2_LAUNCH_CHECK();
    k_msm_reduce<<<(unsigned)nj, MSM_RED_THREADS, MSM_RED_THREADS * sizeof(jac), ctx->stream>>>((const MsmJob *)d_jobs, (const jac *)d_partial, maxblk, d_out + c0);
    SP2_LAUNCH_CHECK();
  }
  return SP2_OK;
}

int hyrax_bind_dev(sp2_ctx *ctx, const fe *d_poly, const fe *d_L, uint64_t rows, uint64_t r_len, fe *d_out) {
  const unsigned rg = (unsigned)std::min<uint64_t>(rows, BIND_RG);
  void *part;
  SP2_TRY(scratch(ctx, 12, (size_t)rg * r_len * sizeof(fe), &part));
  k_hyrax_bind<<<dim3((unsigned)((r_len + 127) / 128), rg), 128, 0, ctx->stream>>>(d_poly, d_L, rows, r_len, (fe *)part);
  SP2_LAUNCH_CHECK();
  k_sum_rows<<<(unsigned)((r_len + 127) / 128), 128, 0, ctx->stream>>>((const fe *)part, rg, r_len, d_out);
  SP2_LAUNCH_CHECK();
  return SP2_OK;
}

}  // namespace sp2

extern "C" {

int32_t sp2_ck_upload(sp2_ctx *ctx, const uint64_t *bases_xy, uint32_t n, const uint64_t *h_xy, const uint64_t *ck_s_xy,
                      const uint64_t *h_s_xy, sp2_ck **out) {
  cudaSetDevice(ctx->device);
  if (!out) return SP2_ERR_INTERNAL;
  *out = nullptr;
  if (n == 0 || n > (1u << 24)) return set_error(ctx, SP2_ERR_INVALID_CK_LENGTH, "ck: bad number of bases");
  sp2_ck *ck = new sp2_ck();
  ck->ctx = ctx; ck->n = n; ck->nbase = n + 3;
  aff *d_bases = nullptr; jac *tmp = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_bases, (size_t)ck->nbase * sizeof(aff));
  if (e == cudaSuccess) e = cudaMalloc((void **)&tmp, (size_t)MSM_NW * ck->nbase * sizeof(jac));
  if (e == cudaSuccess) e = cudaMalloc((void **)&ck->table, (size_t)MSM_NW * ck->nbase * sizeof(aff));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_bases, bases_xy, (size_t)n * sizeof(aff), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_bases + n, h_xy, sizeof(aff), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_bases + n + 1, ck_s_xy, sizeof(aff), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_bases + n + 2, h_s_xy, sizeof(aff), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) {
    const unsigned blocks = (ck->nbase + 127) / 128;
    k_ck_windows<<<blocks, 128, 0, ctx->stream>>>(d_bases, ck->nbase, tmp);
    k_ck_normalize<<<blocks, 128, 0, ctx->stream>>>(tmp, ck->nbase, ck->table);
    ctx->launches += 2;
    e = cudaStreamSynchronize(ctx->stream);
  }
  if (d_bases) cudaFree(d_bases);
  if (tmp) cudaFree(tmp);
  if (e != cudaSuccess) { if (ck->table) cudaFree(ck->table); delete ck; return set_cuda_error(ctx, e, "ck upload", __LINE__); }
  *out = ck;
  return SP2_OK;
}

void sp2_ck_free(sp2_ck *ck) {
  if (!ck) return;
  cudaSetDevice(ck->ctx->device);
  cudaStreamSynchronize(ck->ctx->stream);
  if (ck->table) cudaFree(ck->table);
  delete ck;
}

/* DlogGroupExt::vartime_multiscalar_mul(scalars, ck[..n]) (provider/traits.rs:118-134 -> msm.rs:187) */
int32_t sp2_msm(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *scalars, uint32_t n, uint64_t *out_xy) {
  cudaSetDevice(ctx->device);
  if (n > ck->n) return set_error(ctx, SP2_ERR_INVALID_CK_LENGTH, "msm: more scalars than commitment-key bases");
  void *d_s, *d_o;
  SP2_TRY(scratch(ctx, 0, (size_t)n * sizeof(fe) + 32, &d_s)); SP2_TRY(scratch(ctx, 1, sizeof(aff), &d_o));
  if (n) SP2_CUDA_OK(cudaMemcpyAsync(d_s, scalars, (size_t)n * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  MsmJob j; memset(&j, 0, sizeof(j));
  j.scalars = (const fe *)d_s; j.len = n; j.base0 = 0; j.nextra = 0;
  SP2_TRY(msm_run(ctx, ck, std::vector<MsmJob>{j}, (aff *)d_o));
  SP2_CUDA_OK(cudaMemcpyAsync(out_xy, d_o, sizeof(aff), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

/* HyraxPCS::commit (hyrax_pc.rs:207-303; commit_zeros :305-319 is the all-zero v): rows of ck->n scalars,
 * out_rows[i] = sum_j v[i*n + j] * ck_j + blinds[i] * h.  d_v may be shorter than rows*n (last row ragged). */
int32_t sp2_hyrax_commit_dev(sp2_ctx *ctx, const sp2_ck *ck, const void *d_v, uint64_t len, const void *d_blinds, uint64_t rows,
                             void *d_out_rows) {
  cudaSetDevice(ctx->device);
  if (rows < (len + ck->n - 1) / ck->n) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "hyrax commit: too few rows for the vector");
  std::vector<MsmJob> jobs(rows);
  for (uint64_t i = 0; i < rows; i++) {
    MsmJob &j = jobs[i]; memset(&j, 0, sizeof(j));
    const uint64_t lo = i * ck->n, hi = std::min<uint64_t>(len, lo + ck->n);
    j.scalars = (const fe *)d_v + lo; j.len = hi > lo ? (u32)(hi - lo) : 0; j.base0 = 0;
    j.nextra = 1; j.extra_base[0] = ck->idx_h(); j.extra_scalar[0] = (const fe *)d_blinds + i;
  }
  return msm_run(ctx, ck, jobs, (aff *)d_out_rows);
}

int32_t sp2_hyrax_commit(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *v, uint64_t len, const uint64_t *blinds, uint64_t rows,
                         int32_t is_small, uint64_t *out_rows) {
  (void)is_small;   // the reference's small-scalar hint only selects a faster CPU path (hyrax_pc.rs:266-290); same group element
  cudaSetDevice(ctx->device);
  void *d_v, *d_b, *d_o;
  SP2_TRY(scratch(ctx, 0, len * sizeof(fe) + 32, &d_v)); SP2_TRY(scratch(ctx, 1, rows * sizeof(fe) + 32, &d_b));
  SP2_TRY(scratch(ctx, 2, rows * sizeof(aff) + 32, &d_o));
  if (len) SP2_CUDA_OK(cudaMemcpyAsync(d_v, v, len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(d_b, blinds, rows * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_TRY(sp2_hyrax_commit_dev(ctx, ck, d_v, len, d_b, rows, d_o));
  SP2_CUDA_OK(cudaMemcpyAsync(out_rows, d_o, rows * sizeof(aff), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

/* bind_with_delayed (hyrax_pc.rs:38-54): out[i] = sum_j L[j] * poly[j * r_len + i] */
int32_t sp2_hyrax_bind(sp2_ctx *ctx, const uint64_t *poly, const uint64_t *L, uint64_t rows, uint64_t r_len, uint64_t *out) {
  cudaSetDevice(ctx->device);
  void *d_p, *d_L, *d_o;
  SP2_TRY(scratch(ctx, 0, rows * r_len * sizeof(fe), &d_p)); SP2_TRY(scratch(ctx, 1, rows * sizeof(fe), &d_L)); SP2_TRY(scratch(ctx, 2, r_len * sizeof(fe), &d_o));
  SP2_CUDA_OK(cudaMemcpyAsync(d_p, poly, rows * r_len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(d_L, L, rows * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_TRY(hyrax_bind_dev(ctx, (const fe *)d_p, (const fe *)d_L, rows, r_len, (fe *)d_o));
  SP2_CUDA_OK(cudaMemcpyAsync(out, d_o, r_len * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

/* harness support: n pseudo-random T256 points (seeded multiples of the generator), affine, to the host */
int32_t sp2_test_points(sp2_ctx *ctx, uint64_t seed, uint32_t n, uint64_t *out_xy) {
  cudaSetDevice(ctx->device);
  void *d; SP2_TRY(scratch(ctx, 0, (size_t)n * sizeof(aff) + 64, &d));
  k_test_points<<<(n + 63) / 64, 64, 0, ctx->stream>>>(seed, n, (aff *)d);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(out_xy, d, (size_t)n * sizeof(aff), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

}  // extern "C"
