// Library context: one per (process, GPU).  Owns the stream, a bump-free set of scratch
// allocations and the last error string (errors are values, never C++ exceptions across the ABI;
// mirrors the reference's SpartanError, src/errors.rs:12-110).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "../../include/spartan2_b200.h"

struct sp2_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side = nullptr;          // side stream for independent prologue work
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  int num_sms = 148;
  std::string err;
  uint64_t launches = 0;                // kernels launched through this context (bench's gpu_launches)
  std::vector<void *> owned;            // device allocations freed at destroy
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
template <class T>
inline int dev_alloc(sp2_ctx *ctx, T **p, size_t n) {
  cudaError_t e = cudaMalloc((void **)p, n * sizeof(T) + 32);
  if (e != cudaSuccess) return set_cuda_error(ctx, e, "cudaMalloc", __LINE__);
  return SP2_OK;
}
#define SP2_TRY(expr) do { int rc_ = (expr); if (rc_ != SP2_OK) return rc_; } while (0)
#define SP2_LAUNCH_CHECK() do { ctx->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return sp2::set_cuda_error(ctx, e_, "kernel launch", __LINE__); } while (0)
}  // namespace sp2
