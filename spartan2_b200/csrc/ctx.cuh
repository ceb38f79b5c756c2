// Library context: one per (process, GPU).  Owns the stream, cached scratch allocations and the
// last error string (errors are values, never C++ exceptions across the ABI; mirrors the
// reference's SpartanError, src/errors.rs:12-110).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/spartan2_b200.h"

struct sp2_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;        // all kernels of this context are launched here
  bool own_stream = false;
  cudaStream_t side = nullptr;          // side stream for independent prologue work
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_side = nullptr;
  cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr;   // bracket the last persistent cubic sum-check kernel (bench.py: roofline of the dominant kernel)
  bool ev_k_valid = false;
  int num_sms = 148;
  std::string err;
  uint64_t launches = 0;                // kernels launched through this context (bench's gpu_launches)
  // scratch slots: grown on demand, reused across calls (cudaMalloc/cudaFree are synchronising)
  static const int NSLOT = 24;
  void *slot[NSLOT] = {nullptr};
  size_t slot_bytes[NSLOT] = {0};
  void *pinned = nullptr;               // small pinned staging buffer for result read-back
  size_t pinned_bytes = 0;
  // buffers replaced by a larger one: kept until the context goes (cudaFree / cudaFreeHost synchronise the device, and the
  // library keeps kernels in flight that wait for the HOST — the prover's gate — or for a peer)
  std::vector<void *> retired, retired_pinned;
};

namespace sp2 {
inline int set_error(sp2_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg;
  return code;
}
inline int set_cuda_error(sp2_ctx *ctx, cudaError_t e, const char *what, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s [line %d]", (int)e, cudaGetErrorString(e), what, line);
  return set_error(ctx, SP2_ERR_CUDA, buf);
}
#define SP2_TRY(expr) do { int rc_ = (expr); if (rc_ != SP2_OK) return rc_; } while (0)
#define SP2_CUDA_OK(call)                                                     \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess) return sp2::set_cuda_error(ctx, e_, #call, __LINE__); \
  } while (0)
#define SP2_LAUNCH_CHECK() do { ctx->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return sp2::set_cuda_error(ctx, e_, "kernel launch", __LINE__); } while (0)

// scratch slot `i` of at least `bytes` bytes (contents undefined)
inline int scratch(sp2_ctx *ctx, int i, size_t bytes, void **out) {
  if (ctx->slot_bytes[i] < bytes) {
    if (ctx->slot[i]) { ctx->retired.push_back(ctx->slot[i]); ctx->slot[i] = nullptr; ctx->slot_bytes[i] = 0; }   // (work in flight may still use it)
    size_t want = bytes + bytes / 4 + 256;
    SP2_CUDA_OK(cudaMalloc(&ctx->slot[i], want));
    ctx->slot_bytes[i] = want;
  }
  *out = ctx->slot[i];
  return SP2_OK;
}
inline int pinned(sp2_ctx *ctx, size_t bytes, void **out) {
  if (ctx->pinned_bytes < bytes) {
    if (ctx->pinned) { ctx->retired_pinned.push_back(ctx->pinned); ctx->pinned = nullptr; ctx->pinned_bytes = 0; }
    size_t want = bytes * 2 + 4096;
    SP2_CUDA_OK(cudaMallocHost(&ctx->pinned, want));
    ctx->pinned_bytes = want;
  }
  *out = ctx->pinned;
  return SP2_OK;
}
}  // namespace sp2
