// eq~ table and top-variable bind entry points (reference src/polys/eq.rs, src/polys/multilinear.rs).
#include "ctx.cuh"
#include "polys.cuh"

using namespace sp2;

// Both sqrt-size factor tables in one launch: block 0 -> hi (first k-s points), block 1 -> lo (last s).
__global__ void __launch_bounds__(1024) k_eq_factors(const fe *r, int k, int s, fe *hi_pref, fe *lo_pref) {
  if (blockIdx.x == 0) eq_prefix_block(r, k - s, hi_pref);
  else eq_prefix_block(r + (k - s), s, lo_pref);
}

// eq~(r, i) = eq~(r[0..k-s), i >> s) * eq~(r[k-s..k), i & (2^s - 1)): one modmul and one 32-byte
// store per entry (write-only HBM traffic), instead of the reference's k doubling passes.
__global__ void __launch_bounds__(256) k_eq_product(const fe *hi, const fe *lo, int s, size_t n, fe *out) {
  const size_t mask = ((size_t)1 << s) - 1;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    stg_fe(out + i, Fq::mul(ldg_fe_ro(hi + (i >> s)), ldg_fe_ro(lo + (i & mask))));
}

__global__ void __launch_bounds__(256) k_bind_top(const fe *Z, size_t n, const fe *r, fe *out) {
  const fe rr = ldg_fe_ro(r);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    stg_fe(out + i, bind_pair(ldg_fe(Z + i), ldg_fe(Z + n + i), rr));
}

namespace sp2 {
// device-resident eq table: d_r (k points) -> d_out (2^k); uses scratch slot 15 for the factor tables
int eq_table_dev(sp2_ctx *ctx, const fe *d_r, uint32_t k, fe *d_out, cudaStream_t stream = nullptr, int slot = 15) {
  if (!stream) stream = ctx->stream;
  const int s = (int)k / 2;
  const size_t n = (size_t)1 << k;
  void *fac;
  SP2_TRY(scratch(ctx, slot, (((size_t)2 << (k - s)) + ((size_t)2 << s)) * sizeof(fe), &fac));
  fe *hi_pref = (fe *)fac, *lo_pref = hi_pref + ((size_t)2 << (k - s));
  k_eq_factors<<<2, 1024, 0, stream>>>(d_r, (int)k, s, hi_pref, lo_pref);
  SP2_LAUNCH_CHECK();
  const fe *hi = hi_pref + (((size_t)1 << (k - s)) - 1), *lo = lo_pref + (((size_t)1 << s) - 1);
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > (unsigned)ctx->num_sms * 8) blocks = ctx->num_sms * 8;
  k_eq_product<<<blocks, 256, 0, stream>>>(hi, lo, s, n, d_out);
  SP2_LAUNCH_CHECK();
  return SP2_OK;
}
// (see sumcheck_cubic_preload / _reserve: the same for an eq table built behind the prover's gate)
int eq_table_reserve(sp2_ctx *ctx, uint32_t k, int slot) {
  static unsigned loaded_mask = 0;
  if (!(loaded_mask >> ctx->device & 1u)) { cudaFuncAttributes a; cudaFuncGetAttributes(&a, (const void *)k_eq_factors); cudaFuncGetAttributes(&a, (const void *)k_eq_product); loaded_mask |= 1u << ctx->device; }
  const int s = (int)k / 2;
  void *fac;
  return scratch(ctx, slot, (((size_t)2 << (k - s)) + ((size_t)2 << s)) * sizeof(fe), &fac);
}
int bind_top_dev(sp2_ctx *ctx, const fe *d_Z, size_t len, const fe *d_r, fe *d_out) {
  const size_t n = len / 2;
  if (n == 0) return SP2_OK;
  unsigned blocks = (unsigned)((n + 255) / 256);
  if (blocks > (unsigned)ctx->num_sms * 8) blocks = ctx->num_sms * 8;
  k_bind_top<<<blocks, 256, 0, ctx->stream>>>(d_Z, n, d_r, d_out);
  SP2_LAUNCH_CHECK();
  return SP2_OK;
}
}  // namespace sp2

extern "C" {

int32_t sp2_eq_table_dev(sp2_ctx *ctx, const void *d_r, uint32_t k, void *d_out) {
  cudaSetDevice(ctx->device);
  if (k > 34) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "eq_table: k too large");
  return eq_table_dev(ctx, (const fe *)d_r, k, (fe *)d_out);
}

int32_t sp2_eq_table(sp2_ctx *ctx, const uint64_t *r, uint32_t k, uint64_t *out) {
  cudaSetDevice(ctx->device);
  if (k > 30) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "eq_table: k too large");
  const size_t n = (size_t)1 << k;
  void *d_r, *d_out;
  SP2_TRY(scratch(ctx, 0, (k + 1) * sizeof(fe), &d_r));
  SP2_TRY(scratch(ctx, 1, n * sizeof(fe), &d_out));
  if (k) SP2_CUDA_OK(cudaMemcpyAsync(d_r, r, k * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_TRY(eq_table_dev(ctx, (const fe *)d_r, k, (fe *)d_out));
  SP2_CUDA_OK(cudaMemcpyAsync(out, d_out, n * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

int32_t sp2_bind_top_dev(sp2_ctx *ctx, const void *d_Z, uint64_t len, const void *d_r, void *d_out) {
  cudaSetDevice(ctx->device);
  if (len & (len - 1)) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "bind_top: length must be a power of two");
  return bind_top_dev(ctx, (const fe *)d_Z, len, (const fe *)d_r, (fe *)d_out);
}

int32_t sp2_bind_top(sp2_ctx *ctx, const uint64_t *Z, uint64_t len, const uint64_t *r, uint64_t *out) {
  cudaSetDevice(ctx->device);
  if (len < 2 || (len & (len - 1))) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "bind_top: length must be a power of two >= 2");
  void *d_Z, *d_r, *d_out;
  SP2_TRY(scratch(ctx, 0, len * sizeof(fe), &d_Z));
  SP2_TRY(scratch(ctx, 1, sizeof(fe), &d_r));
  SP2_TRY(scratch(ctx, 2, len / 2 * sizeof(fe), &d_out));
  SP2_CUDA_OK(cudaMemcpyAsync(d_Z, Z, len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(d_r, r, sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_TRY(bind_top_dev(ctx, (const fe *)d_Z, len, (const fe *)d_r, (fe *)d_out));
  SP2_CUDA_OK(cudaMemcpyAsync(out, d_out, len / 2 * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

}  // extern "C"
