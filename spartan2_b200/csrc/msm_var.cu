// Variable-base multi-scalar multiplication on the device: the DlogGroupExt surface for ARBITRARY bases
// (reference src/provider/traits.rs:118-162):
//   vartime_multiscalar_mul            -> sp2_msm_var            (msm.rs:59-222, signed-digit Pippenger, c = 8)
//   batch_vartime_multiscalar_mul      -> sp2_msm_batch_var      (traits.rs:127-135: many scalar vectors, one base slice)
//   vartime_multiscalar_mul_small      -> sp2_msm_small_var      (msm.rs:367-620: u64 scalars — only 9 windows are non-zero)
//   vartime_multiscalar_mul_shared_weights -> sp2_msm_shared_weights (msm.rs:228-356: one scalar vector, many base rows)
// One engine serves all four: a batch of jobs (scalars_j, bases_j, n_j), signed byte digits (msm.rs:122-148) taken straight
// from the canonical scalar by a closed form (carry into window w = [low w bytes > 0x80..80]), and per (job, window) ONE CTA:
//   counting sort of the terms by |digit| in shared memory -> thread b accumulates bucket b with mixed additions over its
//   contiguous list (no warp divergence from scattered matches) -> (b) * bucket by an 8-bit double-and-add -> CTA tree sum,
// then per job one warp combines the windows: lane w doubles its window sum 8w times, shuffle tree.  The group elements
// are the reference's (any summation order gives the same point), so the affine results are bit-identical to the oracle.
// Where the bases are the commitment key's, the table-driven sp2_msm / sp2_hyrax_commit (msm.cu) is ~10x faster: a
// variable base costs a chain of 8 doublings per window that no table can remove.
#include <string.h>
#include <vector>
#include "msm.cuh"
#include "devutil.cuh"
#include "host_transcript.h"

using namespace sp2;

namespace {

constexpr int VB = 128;                   // buckets per window = CTA threads
constexpr int VCHUNK = 2048;              // terms sorted per pass

struct VarJob { const fe *scalars; const u64 *small; const aff *bases; u32 n; u32 nwin; };

__device__ __forceinline__ jac ldj(const jac *p) { jac r; r.x = ldg_fe(&p->x); r.y = ldg_fe(&p->y); r.z = ldg_fe(&p->z); return r; }
__device__ __forceinline__ void stj(jac *p, const jac &v) { stg_fe(&p->x, v.x); stg_fe(&p->y, v.y); stg_fe(&p->z, v.z); }

// signed byte digit w (0..32) of the canonical 256-bit integer s: d = byte_w + carry_w, minus 256 if > 128 (msm.rs:122-148);
// carry_w = 1 iff the low w bytes exceed 0x80 80 ... 80 (the recurrence c_w = [byte_{w-1} + c_{w-1} > 128] in closed form)
__device__ __forceinline__ int signed_digit(const fe &s, int w) {
  int carry = 0;
  if (w > 0) {
    const int full = w >> 2, rem = w & 3;               // w bytes = `full` limbs + `rem` bytes of limb `full`
    int cmp = 0;                                         // sign of (low part - pattern), decided from the top down
    if (rem) { const u32 m = (1u << (8 * rem)) - 1; const u32 a = s.v[full] & m, b = 0x80808080u & m; cmp = a > b ? 1 : (a < b ? -1 : 0); }
#pragma unroll
    for (int k = 7; k >= 0; k--) if (k < full && cmp == 0) { const u32 a = s.v[k]; cmp = a > 0x80808080u ? 1 : (a < 0x80808080u ? -1 : 0); }
    carry = cmp > 0;
  }
  int d = (w < 32 ? (int)((s.v[w >> 2] >> (8 * (w & 3))) & 0xffu) : 0) + carry;
  if (d > 128) d -= 256;
  return d;
}

struct VarSmem { short dig[VCHUNK]; unsigned short order[VCHUNK]; int hist[VB + 1]; int start[VB + 1]; jac red[VB / 32]; };

__global__ void __launch_bounds__(VB) k_var_window(const VarJob *jobs, jac *wsum, u32 maxwin) {
  __shared__ VarSmem sm;
  const VarJob job = jobs[blockIdx.y];
  const u32 w = blockIdx.x, tid = threadIdx.x;
  jac *out = wsum + (size_t)blockIdx.y * maxwin + w;
  if (w >= job.nwin) { if (tid == 0) stj(out, jac_inf()); return; }
  jac acc = jac_inf();                                   // bucket tid + 1
  for (u32 c0 = 0; c0 < job.n; c0 += VCHUNK) {
    const u32 cn = min((u32)VCHUNK, job.n - c0);
    for (u32 i = tid; i <= VB; i += VB) sm.hist[i] = 0;
    __syncthreads();
    for (u32 i = tid; i < cn; i += VB) {
      fe s;
      if (job.small) { const u64 v = job.small[c0 + i]; s = Fq::zero(); s.v[0] = (u32)v; s.v[1] = (u32)(v >> 32); }
      else s = Fq::from_mont(ldg_fe(job.scalars + c0 + i));
      const int d = signed_digit(s, (int)w);
      sm.dig[i] = (short)d;
      if (d) atomicAdd(&sm.hist[d < 0 ? -d : d], 1);
    }
    __syncthreads();
    if (tid == 0) { int run = 0; for (int b = 1; b <= VB; b++) { sm.start[b] = run; run += sm.hist[b]; } }
    __syncthreads();
    for (u32 i = tid; i <= VB; i += VB) sm.hist[i] = 0;   // reuse as fill counters
    __syncthreads();
    for (u32 i = tid; i < cn; i += VB) {
      const int d = sm.dig[i];
      if (d) { const int b = d < 0 ? -d : d; const int pos = sm.start[b] + atomicAdd(&sm.hist[b], 1); sm.order[pos] = (unsigned short)i; }
    }
    __syncthreads();
    const int b = (int)tid + 1, s0 = sm.start[b], cnt = sm.hist[b];
    for (int k = 0; k < cnt; k++) {
      const u32 i = sm.order[s0 + k];
      aff pt; pt.x = ldg_fe_ro(&job.bases[c0 + i].x); pt.y = ldg_fe_ro(&job.bases[c0 + i].y);
      if (sm.dig[i] < 0) pt.y = Fp::neg(pt.y);
      acc = jac_add_mixed(acc, pt);
    }
    __syncthreads();
  }
  // (tid + 1) * bucket, 8-bit double-and-add from the top bit
  jac m = jac_inf();
  const u32 k = tid + 1;
#pragma unroll 1
  for (int bit = 7; bit >= 0; bit--) { m = jac_dbl(m); if ((k >> bit) & 1u) m = jac_add(m, acc); }
  // CTA sum: shuffle tree per warp, then the 4 warp results
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    jac o;
#pragma unroll
    for (int q = 0; q < 8; q++) { o.x.v[q] = __shfl_down_sync(0xffffffffu, m.x.v[q], d); o.y.v[q] = __shfl_down_sync(0xffffffffu, m.y.v[q], d); o.z.v[q] = __shfl_down_sync(0xffffffffu, m.z.v[q], d); }
    if ((tid & 31) < (u32)d) m = jac_add(m, o);
  }
  if ((tid & 31) == 0) sm.red[tid >> 5] = m;
  __syncthreads();
  if (tid == 0) stj(out, jac_add(jac_add(sm.red[0], sm.red[1]), jac_add(sm.red[2], sm.red[3])));
}

// sum_w 2^(8w) S_w per job: lane w doubles S_w 8w times (window 32, if present, rides on lane 31 for 8 more), shuffle tree
__global__ void __launch_bounds__(32) k_var_combine(const VarJob *jobs, const jac *wsum, u32 maxwin, jac *out) {
  const VarJob job = jobs[blockIdx.x];
  const u32 lane = threadIdx.x;
  jac acc = lane < job.nwin ? ldj(wsum + (size_t)blockIdx.x * maxwin + lane) : jac_inf();
  if (lane == 31 && job.nwin > 32) {                     // 2^256 S_32 + 2^248 S_31 = 2^248 (2^8 S_32 + S_31)
    jac top = ldj(wsum + (size_t)blockIdx.x * maxwin + 32);
#pragma unroll 1
    for (int k = 0; k < 8; k++) top = jac_dbl(top);
    acc = jac_add(acc, top);
  }
#pragma unroll 1
  for (u32 k = 0; k < 8 * lane; k++) acc = jac_dbl(acc);
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    jac o;
#pragma unroll
    for (int q = 0; q < 8; q++) { o.x.v[q] = __shfl_down_sync(0xffffffffu, acc.x.v[q], d); o.y.v[q] = __shfl_down_sync(0xffffffffu, acc.y.v[q], d); o.z.v[q] = __shfl_down_sync(0xffffffffu, acc.z.v[q], d); }
    if (lane < (u32)d) acc = jac_add(acc, o);
  }
  if (lane == 0) stj(out + blockIdx.x, acc);
}

// run `jobs` (device pointers inside) and return the affine results on the host
int var_run(sp2_ctx *ctx, std::vector<VarJob> &jobs, uint64_t *out_xy) {
  const size_t nj = jobs.size();
  if (!nj) return SP2_OK;
  u32 maxwin = 1;
  for (auto &j : jobs) maxwin = std::max(maxwin, j.nwin);
  void *d_jobs, *d_w, *d_o;
  SP2_TRY(scratch(ctx, 10, nj * sizeof(VarJob), &d_jobs));
  SP2_TRY(scratch(ctx, 11, nj * maxwin * sizeof(jac), &d_w));
  SP2_TRY(scratch(ctx, 5, nj * sizeof(jac), &d_o));
  SP2_CUDA_OK(cudaMemcpyAsync(d_jobs, jobs.data(), nj * sizeof(VarJob), cudaMemcpyHostToDevice, ctx->stream));
  k_var_window<<<dim3(maxwin, (unsigned)nj), VB, 0, ctx->stream>>>((const VarJob *)d_jobs, (jac *)d_w, maxwin);
  SP2_LAUNCH_CHECK();
  k_var_combine<<<(unsigned)nj, 32, 0, ctx->stream>>>((const VarJob *)d_jobs, (const jac *)d_w, maxwin, (jac *)d_o);
  SP2_LAUNCH_CHECK();
  std::vector<uint64_t> hj(nj * 12);
  SP2_CUDA_OK(cudaMemcpyAsync(hj.data(), d_o, nj * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  sp2h::batch_normalize(hj.data(), nj, out_xy);
  return SP2_OK;
}

int upload(sp2_ctx *ctx, int slot, const void *h, size_t bytes, void **d) {
  SP2_TRY(scratch(ctx, slot, bytes + 64, d));
  if (bytes) SP2_CUDA_OK(cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return SP2_OK;
}

}  // namespace

extern "C" {

/* DlogGroupExt::vartime_multiscalar_mul (provider/traits.rs:118-125 -> msm.rs:187-222): out = sum_i scalars[i] * bases[i] */
int32_t sp2_msm_var(sp2_ctx *ctx, const uint64_t *scalars, const uint64_t *bases_xy, uint32_t n, uint64_t *out_xy) {
  cudaSetDevice(ctx->device);
  if (n == 0) { memset(out_xy, 0, 64); return SP2_OK; }
  void *ds, *db;
  SP2_TRY(upload(ctx, 0, scalars, (size_t)n * sizeof(fe), &ds)); SP2_TRY(upload(ctx, 1, bases_xy, (size_t)n * sizeof(aff), &db));
  std::vector<VarJob> jobs(1); jobs[0] = VarJob{(const fe *)ds, nullptr, (const aff *)db, n, 33};
  return var_run(ctx, jobs, out_xy);
}

/* vartime_multiscalar_mul_small (traits.rs:137-142 -> msm.rs:367-620): u64 scalars */
int32_t sp2_msm_small_var(sp2_ctx *ctx, const uint64_t *scalars_u64, const uint64_t *bases_xy, uint32_t n, uint64_t *out_xy) {
  cudaSetDevice(ctx->device);
  if (n == 0) { memset(out_xy, 0, 64); return SP2_OK; }
  uint64_t mx = 0; for (uint32_t i = 0; i < n; i++) mx |= scalars_u64[i];
  u32 bytes = 0; while (bytes < 8 && (mx >> (8 * bytes))) bytes++;
  void *ds, *db;
  SP2_TRY(upload(ctx, 0, scalars_u64, (size_t)n * 8, &ds)); SP2_TRY(upload(ctx, 1, bases_xy, (size_t)n * sizeof(aff), &db));
  std::vector<VarJob> jobs(1); jobs[0] = VarJob{nullptr, (const u64 *)ds, (const aff *)db, n, bytes + 1};   // + the carry window
  return var_run(ctx, jobs, out_xy);
}

/* batch_vartime_multiscalar_mul (traits.rs:127-135): k scalar vectors (concatenated, lens[j] each) against the SAME base slice
 * bases[..lens[j]]; out_xy: k points */
int32_t sp2_msm_batch_var(sp2_ctx *ctx, const uint64_t *scalars, const uint32_t *lens, uint32_t k, const uint64_t *bases_xy, uint64_t *out_xy) {
  cudaSetDevice(ctx->device);
  if (k == 0) return SP2_OK;
  size_t tot = 0; u32 mxl = 0;
  for (u32 j = 0; j < k; j++) { tot += lens[j]; mxl = std::max(mxl, lens[j]); }
  void *ds, *db;
  SP2_TRY(upload(ctx, 0, scalars, tot * sizeof(fe), &ds)); SP2_TRY(upload(ctx, 1, bases_xy, (size_t)mxl * sizeof(aff), &db));
  std::vector<VarJob> jobs(k); size_t off = 0;
  for (u32 j = 0; j < k; j++) { jobs[j] = VarJob{(const fe *)ds + off, nullptr, (const aff *)db, lens[j], lens[j] ? 33u : 0u}; off += lens[j]; }
  return var_run(ctx, jobs, out_xy);
}

/* vartime_multiscalar_mul_shared_weights (traits.rs:155-161 -> msm.rs:228-356): out[r] = sum_i weights[i] * bases_rows[r][i];
 * bases_rows: rows x n points, row-major */
int32_t sp2_msm_shared_weights(sp2_ctx *ctx, const uint64_t *weights, uint32_t n, const uint64_t *bases_rows_xy, uint32_t rows, uint64_t *out_xy) {
  cudaSetDevice(ctx->device);
  if (rows == 0) return SP2_OK;
  if (n == 0) { memset(out_xy, 0, (size_t)rows * 64); return SP2_OK; }
  void *ds, *db;
  SP2_TRY(upload(ctx, 0, weights, (size_t)n * sizeof(fe), &ds)); SP2_TRY(upload(ctx, 1, bases_rows_xy, (size_t)rows * n * sizeof(aff), &db));
  std::vector<VarJob> jobs(rows);
  for (u32 r = 0; r < rows; r++) jobs[r] = VarJob{(const fe *)ds, nullptr, (const aff *)db + (size_t)r * n, n, 33};
  return var_run(ctx, jobs, out_xy);
}

}  // extern "C"
