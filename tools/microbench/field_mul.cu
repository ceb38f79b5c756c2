// Microbenchmark of the 256-bit field multiplications of spartan2_b200/csrc/field.cuh on sm_100a:
//   latency   : one warp, a chain of N dependent multiplications x <- x * y (clock64 cycles per multiplication)
//   throughput: a full grid (CTAs x 256 threads, U independent chains per thread), multiplications / s
// Variants: the current Fq::mul (wide product + multiplication-free P-256 REDC), the current Fp::mul (wide product +
// word-by-word REDC shaped on the T256 base modulus), the lockstep pair Fp::mul2, and mont_mul_interleaved<> for both.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/microbench/field_mul tools/microbench/field_mul.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../spartan2_b200/csrc/field.cuh"
using namespace sp2;

struct VFq { static __device__ __forceinline__ fe mul(const fe &a, const fe &b) { return Fq::mul(a, b); } static const char *name() { return "Fq::mul (wide + mult-free REDC)"; } };
struct VFqIl { static __device__ __forceinline__ fe mul(const fe &a, const fe &b) { return mont_mul_interleaved<FqParams>(a, b); } static const char *name() { return "Fq interleaved"; } };
struct VFp { static __device__ __forceinline__ fe mul(const fe &a, const fe &b) { return Fp::mul_inl(a, b); } static const char *name() { return "Fp::mul_inl (wide + shaped CIOS)"; } };
struct VFpNi { static __device__ __forceinline__ fe mul(const fe &a, const fe &b) { return Fp::mul(a, b); } static const char *name() { return "Fp::mul (out of line)"; } };
struct VFpIl { static __device__ __forceinline__ fe mul(const fe &a, const fe &b) { return mont_mul_interleaved<FpParams>(a, b); } static const char *name() { return "Fp interleaved"; } };

template <class V>
__global__ void k_latency(fe x0, fe y0, int n, unsigned long long *cycles, fe *sink) {
  fe x = x0, y = y0;
  x.v[0] ^= threadIdx.x;
  const unsigned long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) x = V::mul(x, y);
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// two independent chains per thread through Fp::mul2 (the lockstep pair)
__global__ void k_latency_mul2(fe x0, fe y0, int n, unsigned long long *cycles, fe *sink) {
  fe x = x0, y = y0, z = y0;
  x.v[0] ^= threadIdx.x; z.v[1] ^= threadIdx.x;
  const unsigned long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) Fp::mul2(x, y, z, y, x, z);
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = Fp::add(x, z);
}
template <class V, int U>
__global__ void __launch_bounds__(256) k_throughput(fe x0, fe y0, int n, fe *sink) {
  fe x[U], y = y0;
#pragma unroll
  for (int u = 0; u < U; u++) { x[u] = x0; x[u].v[0] ^= (threadIdx.x + 977 * u + blockIdx.x); }
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < U; u++) x[u] = V::mul(x[u], y);
  }
  fe s = x[0];
#pragma unroll
  for (int u = 1; u < U; u++) s = Field<FqParams>::add(s, x[u]);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// raw pipes: dependent and independent IMAD.WIDE.X / IADD3.X chains
__global__ void k_imad_chain(unsigned a, unsigned b, int n, unsigned long long *cycles, unsigned *sink, int ilp) {
  u32 acc[8][2];
  for (int k = 0; k < 8; k++) { acc[k][0] = a + k + threadIdx.x; acc[k][1] = b; }
  const unsigned long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) if (k < ilp) { acc[k][0] = mad_lo_cc(acc[k][1], b, acc[k][0]); acc[k][1] = madc_hi_cc(acc[k][1], b, acc[k][1]); }
  }
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  unsigned s = 0; for (int k = 0; k < 8; k++) s ^= acc[k][0] ^ acc[k][1];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_iadd_chain(unsigned a, unsigned b, int n, unsigned long long *cycles, unsigned *sink, int ilp) {
  u32 acc[8];
  for (int k = 0; k < 8; k++) acc[k] = a + k + threadIdx.x;
  const unsigned long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) if (k < ilp) { acc[k] = add_cc(acc[k], b); acc[k] = addc_cc(acc[k], b); acc[k] = addc_cc(acc[k], a); acc[k] = addc(acc[k], b); }
  }
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) *cycles = t1 - t0;
  unsigned s = 0; for (int k = 0; k < 8; k++) s ^= acc[k];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class V>
void run(fe x, fe y, fe *sink, unsigned long long *d_cyc, int sms) {
  const int n = 512;
  unsigned long long h;
  for (int warps = 1; warps <= 8; warps *= 2) {
    k_latency<V><<<1, 32 * warps>>>(x, y, n, d_cyc, sink);
    k_latency<V><<<1, 32 * warps>>>(x, y, n, d_cyc, sink);
    cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-36s latency  %d warp(s)/SM: %7.1f cycles per multiplication\n", V::name(), warps, (double)h / n);
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const int iters = 256;
#define TP(U, CTAS)                                                                                                          \
  k_throughput<V, U><<<sms * CTAS, 256>>>(x, y, 16, sink); cudaEventRecord(e0); k_throughput<V, U><<<sms * CTAS, 256>>>(x, y, iters, sink); \
  cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);                                         \
  printf("%-36s throughput %d CTAs/SM x 256 thr, %d chains/thread: %8.2f G mul/s  (%.1f cycles per warp-mul per SMSP at 1.965 GHz)\n", V::name(), CTAS, U, \
         (double)sms * CTAS * 256 * U * iters / (ms * 1e-3) / 1e9, 1.965e9 * (ms * 1e-3) / ((double)CTAS * 2 * U * iters));
  TP(1, 1) TP(1, 2) TP(2, 1) TP(2, 2) TP(4, 1)
#undef TP
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  printf("%s, %d SMs\n", pr.name, sms);
  fe x, y;
  for (int i = 0; i < 8; i++) { x.v[i] = 0x12345678u * (i + 1); y.v[i] = 0x9e3779b9u * (i + 3); }
  x.v[7] &= 0x7fffffffu; y.v[7] &= 0x7fffffffu;
  fe *sink; cudaMalloc(&sink, (size_t)sms * 4 * 256 * sizeof(fe));
  unsigned long long *d_cyc; cudaMalloc(&d_cyc, 8);
  unsigned long long h;
  for (int ilp = 1; ilp <= 8; ilp *= 2) {
    k_imad_chain<<<1, 32>>>(3, 5, 4096, d_cyc, (unsigned *)sink, ilp); cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("IMAD.WIDE carry chain, 1 warp, %d independent chains: %.2f cycles per IMAD.WIDE\n", ilp, (double)h / (4096.0 * ilp));
    k_imad_chain<<<1, 256>>>(3, 5, 4096, d_cyc, (unsigned *)sink, ilp); cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("IMAD.WIDE carry chain, 8 warps, %d independent chains: %.2f cycles per IMAD.WIDE per SMSP\n", ilp, (double)h / (4096.0 * ilp * 2));
    k_iadd_chain<<<1, 32>>>(3, 5, 4096, d_cyc, (unsigned *)sink, ilp); cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("IADD3.X carry chain,   1 warp, %d independent chains: %.2f cycles per add\n", ilp, (double)h / (4096.0 * ilp * 4));
    k_iadd_chain<<<1, 256>>>(3, 5, 4096, d_cyc, (unsigned *)sink, ilp); cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
    printf("IADD3.X carry chain,   8 warps, %d independent chains: %.2f cycles per add per SMSP\n", ilp, (double)h / (4096.0 * ilp * 4 * 2));
  }
  run<VFq>(x, y, sink, d_cyc, sms);
  run<VFqIl>(x, y, sink, d_cyc, sms);
  run<VFp>(x, y, sink, d_cyc, sms);
  run<VFpNi>(x, y, sink, d_cyc, sms);
  run<VFpIl>(x, y, sink, d_cyc, sms);
  k_latency_mul2<<<1, 32>>>(x, y, 512, d_cyc, sink); cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
  printf("Fp::mul2 (lockstep pair, out of line) latency 1 warp: %.1f cycles per PAIR\n", (double)h / 512);
  k_latency_mul2<<<1, 256>>>(x, y, 512, d_cyc, sink); cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost);
  printf("Fp::mul2 (lockstep pair, out of line) latency 8 warps: %.1f cycles per PAIR\n", (double)h / 512);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
