// Context, device-memory helpers and element-wise test hooks of the C ABI (include/spartan2_b200.h).
#include <stdlib.h>
#include <string.h>
#include "ctx.cuh"
#include "host_transcript.h"
#include "devutil.cuh"
#include "keccak.cuh"

using namespace sp2;

extern "C" {

int32_t sp2_ctx_create(int32_t device, sp2_ctx **out) {
  if (!out) return SP2_ERR_INTERNAL;
  *out = nullptr;
  // A context uses up to three streams whose kernels wait on each other on the device (bounded spins); with the default 8 hardware
  // queues, streams of several contexts in one process would share a queue and a spinning kernel could sit in front of the kernel
  // it waits for.  Effective only if CUDA is not initialised yet in this process; never overrides the user's setting.
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) return SP2_ERR_CUDA;      // no silent CPU fallback: fail loudly
  if (device < 0 || device >= ndev) return SP2_ERR_CUDA;
  sp2_ctx *ctx = new sp2_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return SP2_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return SP2_ERR_CUDA; }
  if (prop.major < 10) { delete ctx; return SP2_ERR_CUDA; }    // built for sm_100a only
  ctx->num_sms = prop.multiProcessorCount;
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return SP2_ERR_CUDA; }
  ctx->own_stream = true;
  cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking);
  cudaEventCreate(&ctx->ev_a); cudaEventCreate(&ctx->ev_b); cudaEventCreate(&ctx->ev_k0); cudaEventCreate(&ctx->ev_k1);
  cudaEventCreateWithFlags(&ctx->ev_side, cudaEventDisableTiming);
  *out = ctx;
  return SP2_OK;
}

void sp2_ctx_destroy(sp2_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (int i = 0; i < sp2_ctx::NSLOT; i++) if (ctx->slot[i]) cudaFree(ctx->slot[i]);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (void *p : ctx->retired) cudaFree(p);
  for (void *p : ctx->retired_pinned) cudaFreeHost(p);
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
  if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
  if (ctx->ev_k0) cudaEventDestroy(ctx->ev_k0);
  if (ctx->ev_k1) cudaEventDestroy(ctx->ev_k1);
  if (ctx->ev_side) cudaEventDestroy(ctx->ev_side);
  delete ctx;
}

int32_t sp2_ctx_set_stream(sp2_ctx *ctx, void *cuda_stream) {
  if (!ctx) return SP2_ERR_INTERNAL;
  cudaSetDevice(ctx->device);
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->own_stream = false;
  return SP2_OK;
}

const char *sp2_last_error(const sp2_ctx *ctx) { return ctx ? ctx->err.c_str() : "no context"; }
uint64_t sp2_launch_count(const sp2_ctx *ctx) { return ctx ? ctx->launches : 0; }
int32_t sp2_num_sms(const sp2_ctx *ctx) { return ctx ? ctx->num_sms : 0; }

int32_t sp2_synchronize(sp2_ctx *ctx) {
  cudaSetDevice(ctx->device);
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

int32_t sp2_timer_start(sp2_ctx *ctx) { cudaSetDevice(ctx->device); SP2_CUDA_OK(cudaEventRecord(ctx->ev_a, ctx->stream)); return SP2_OK; }
int32_t sp2_timer_stop(sp2_ctx *ctx, float *ms) {
  cudaSetDevice(ctx->device);
  SP2_CUDA_OK(cudaEventRecord(ctx->ev_b, ctx->stream));
  SP2_CUDA_OK(cudaEventSynchronize(ctx->ev_b));
  SP2_CUDA_OK(cudaEventElapsedTime(ms, ctx->ev_a, ctx->ev_b));
  return SP2_OK;
}

/* duration (CUDA events on the library's stream) of the most recent persistent cubic sum-check kernel k_cubic_persist —
 * the largest single kernel of a prove; the caller must have synchronised the stream */
int32_t sp2_last_cubic_persist_ms(sp2_ctx *ctx, float *ms) {
  cudaSetDevice(ctx->device);
  if (!ctx->ev_k_valid) return set_error(ctx, SP2_ERR_INTERNAL, "no persistent cubic kernel has run");
  SP2_CUDA_OK(cudaEventSynchronize(ctx->ev_k1));
  SP2_CUDA_OK(cudaEventElapsedTime(ms, ctx->ev_k0, ctx->ev_k1));
  return SP2_OK;
}
int32_t sp2_dev_alloc(sp2_ctx *ctx, uint64_t bytes, void **out) {
  cudaSetDevice(ctx->device);
  SP2_CUDA_OK(cudaMalloc(out, bytes ? bytes : 32));
  return SP2_OK;
}
int32_t sp2_dev_free(sp2_ctx *ctx, void *p) {
  cudaSetDevice(ctx->device);
  SP2_CUDA_OK(cudaFree(p));
  return SP2_OK;
}
int32_t sp2_dev_upload(sp2_ctx *ctx, void *dst, const void *src, uint64_t bytes) {
  cudaSetDevice(ctx->device);
  SP2_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return SP2_OK;
}
int32_t sp2_dev_download(sp2_ctx *ctx, void *dst, const void *src, uint64_t bytes) {
  cudaSetDevice(ctx->device);
  SP2_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}
int32_t sp2_dev_copy(sp2_ctx *ctx, void *dst, const void *src, uint64_t bytes) {
  cudaSetDevice(ctx->device);
  SP2_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return SP2_OK;
}
int32_t sp2_dev_memset(sp2_ctx *ctx, void *dst, int32_t value, uint64_t bytes) {
  cudaSetDevice(ctx->device);
  SP2_CUDA_OK(cudaMemsetAsync(dst, value, bytes, ctx->stream));
  return SP2_OK;
}
int32_t sp2_host_alloc(sp2_ctx *ctx, uint64_t bytes, void **out) {
  cudaSetDevice(ctx->device);
  SP2_CUDA_OK(cudaMallocHost(out, bytes ? bytes : 32));
  return SP2_OK;
}
int32_t sp2_host_free(sp2_ctx *ctx, void *p) {
  SP2_CUDA_OK(cudaFreeHost(p));
  return SP2_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// test hooks
// ---------------------------------------------------------------------------------------------
template <class F>
__global__ void k_field_op(int op, const fe *a, const fe *b, fe *out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe x = ldg_fe(a + i), y = b ? ldg_fe(b + i) : F::zero(), r;
  switch (op) {
    case 0: r = F::mul(x, y); break;
    case 1: r = F::add(x, y); break;
    case 2: r = F::sub(x, y); break;
    case 3: r = F::inv(x); break;
    case 4: r = F::from_mont(x); break;
    case 5: r = F::to_mont(x); break;
    default: r = F::half(x); break;
  }
  stg_fe(out + i, r);
}

// one block: sum_i a_i*b_i through the delayed-reduction accumulator, then a tree of wide adds
__global__ void k_dot_delayed(const fe *a, const fe *b, size_t n, fe *out) {
  __shared__ Fq::acc sh[256];
  Fq::acc acc = Fq::acc_zero();
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) Fq::mul_acc(acc, ldg_fe(a + i), ldg_fe(b + i));
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) Fq::acc_add(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) stg_fe(out, Fq::acc_reduce(sh[0]));
}

__global__ void k_transcript(DevTranscript *ts, const char *label, int label_len, fe *out) {
  __shared__ unsigned char buf[2304];
  __shared__ fe ch;
  ts_squeeze_block(ts, label, label_len, buf, &ch);
  if (threadIdx.x == 0) stg_fe(out, ch);
}

extern "C" {

int32_t sp2_test_field_op(sp2_ctx *ctx, int32_t field, int32_t op, const uint64_t *a, const uint64_t *b,
                          uint64_t *out, uint64_t n) {
  cudaSetDevice(ctx->device);
  if (n == 0) return SP2_OK;
  void *da, *db, *dout;
  SP2_TRY(scratch(ctx, 0, n * 32, &da)); SP2_TRY(scratch(ctx, 1, n * 32, &db)); SP2_TRY(scratch(ctx, 2, n * 32, &dout));
  SP2_CUDA_OK(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  if (b) SP2_CUDA_OK(cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  unsigned blocks = (unsigned)((n + 127) / 128);
  if (field == 0) k_field_op<Fq><<<blocks, 128, 0, ctx->stream>>>(op, (fe *)da, b ? (fe *)db : nullptr, (fe *)dout, n);
  else k_field_op<Fp><<<blocks, 128, 0, ctx->stream>>>(op, (fe *)da, b ? (fe *)db : nullptr, (fe *)dout, n);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(out, dout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

int32_t sp2_test_dot_delayed(sp2_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t n, uint64_t *out) {
  cudaSetDevice(ctx->device);
  void *da, *db, *dout;
  SP2_TRY(scratch(ctx, 0, n * 32 + 32, &da)); SP2_TRY(scratch(ctx, 1, n * 32 + 32, &db)); SP2_TRY(scratch(ctx, 2, 32, &dout));
  SP2_CUDA_OK(cudaMemcpyAsync(da, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(db, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
  k_dot_delayed<<<1, 256, 0, ctx->stream>>>((fe *)da, (fe *)db, n, (fe *)dout);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(out, dout, 32, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

int32_t sp2_test_transcript(sp2_ctx *ctx, sp2_transcript_state *ts, const uint8_t *pending, uint32_t pending_len,
                            const char *sq_label, uint64_t *challenge_out) {
  cudaSetDevice(ctx->device);
  if (pending_len > sizeof(((DevTranscript *)0)->pending)) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "pending too long");
  size_t ll = strlen(sq_label);
  if (ll > 32) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "label too long");
  DevTranscript h; memset(&h, 0, sizeof(h));
  h.round = ts->round; h.pending_len = pending_len; memcpy(h.state, ts->state, 64);
  if (pending_len) memcpy(h.pending, pending, pending_len);
  void *d; SP2_TRY(scratch(ctx, 0, sizeof(DevTranscript) + 64 + 32, &d));
  DevTranscript *dts = (DevTranscript *)d; char *dlabel = (char *)d + sizeof(DevTranscript); fe *dout = (fe *)(dlabel + 64);
  SP2_CUDA_OK(cudaMemcpyAsync(dts, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(dlabel, sq_label, ll, cudaMemcpyHostToDevice, ctx->stream));
  k_transcript<<<1, 64, 0, ctx->stream>>>(dts, dlabel, (int)ll, dout);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(&h, dts, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(challenge_out, dout, 32, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ts->round = (uint16_t)h.round; memcpy(ts->state, h.state, 64);
  return SP2_OK;
}

/* ---- host Keccak256Transcript (src/provider/keccak.rs:18-105; TranscriptEngineTrait, src/traits/transcript.rs) ----
 * For host-side drivers that run the Fiat-Shamir steps between per-round device calls (the NeutronNova seams). */

int32_t sp2_transcript_new(const char *label, sp2_transcript **out) {
  if (!out || !label) return SP2_ERR_INTERNAL;
  *out = new sp2_transcript(label);
  return SP2_OK;
}
void sp2_transcript_free(sp2_transcript *t) { delete t; }
int32_t sp2_transcript_absorb_bytes(sp2_transcript *t, const char *label, const uint8_t *data, uint64_t n) {
  t->t.absorb_bytes(label, data, (size_t)n);
  return SP2_OK;
}
/* scalars in Montgomery form; absorbed as 32 big-endian canonical bytes each (src/provider/traits.rs:282-286) */
int32_t sp2_transcript_absorb_scalars(sp2_transcript *t, const char *label, const uint64_t *scalars, uint64_t n) {
  t->t.absorb_scalars(label, scalars, (size_t)n);
  return SP2_OK;
}
/* commitment rows (affine points), framed as HyraxCommitment::to_transcript_bytes (hyrax_pc.rs:714-729) */
int32_t sp2_transcript_absorb_commitment(sp2_transcript *t, const char *label, const uint64_t *rows_xy, uint64_t rows) {
  t->t.absorb_commitment(label, rows_xy, (size_t)rows);
  return SP2_OK;
}
int32_t sp2_transcript_dom_sep(sp2_transcript *t, const char *label) { t->t.dom_sep(label); return SP2_OK; }
/* squeeze (keccak.rs:70-94) -> challenge as a Montgomery scalar */
int32_t sp2_transcript_squeeze(sp2_transcript *t, const char *label, uint64_t *out_scalar) {
  uint8_t dg[64];
  t->t.squeeze(label, dg);
  sp2h::fq_from_uniform(dg, out_scalar);
  return SP2_OK;
}
int32_t sp2_transcript_get_state(const sp2_transcript *t, sp2_transcript_state *out) {
  out->round = t->t.round; memcpy(out->state, t->t.state, 64);
  return SP2_OK;
}

}  // extern "C"
