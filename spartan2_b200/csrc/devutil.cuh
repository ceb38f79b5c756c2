// Device utilities: 256-bit global loads/stores, warp/block reductions of field elements.
#pragma once
#include <cuda_runtime.h>
#include "field.cuh"

namespace sp2 {

// One field element per 256-bit transaction (LDG.E.ENL2.256 / STG.E.ENL2.256 on sm_100a):
// a warp moves 1 KiB contiguous per request.  Table loads are .cg (L2 only): streaming data has no L1 reuse, and the
// persistent multi-round kernels re-read entries other CTAs wrote in the previous round.
__device__ __forceinline__ fe ldg_fe(const fe *p) {
  fe r; u64 a, b, c, d;
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  r.v[0] = (u32)a; r.v[1] = (u32)(a >> 32); r.v[2] = (u32)b; r.v[3] = (u32)(b >> 32);
  r.v[4] = (u32)c; r.v[5] = (u32)(c >> 32); r.v[6] = (u32)d; r.v[7] = (u32)(d >> 32);
  return r;
}
// read-only data reused across threads/CTAs (eq tables, z vector): keep it cacheable
__device__ __forceinline__ fe ldg_fe_ro(const fe *p) {
  fe r; u64 a, b, c, d;
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
  r.v[0] = (u32)a; r.v[1] = (u32)(a >> 32); r.v[2] = (u32)b; r.v[3] = (u32)(b >> 32);
  r.v[4] = (u32)c; r.v[5] = (u32)(c >> 32); r.v[6] = (u32)d; r.v[7] = (u32)(d >> 32);
  return r;
}
__device__ __forceinline__ void stg_fe(fe *p, const fe &x) {
  u64 a = (u64)x.v[0] | ((u64)x.v[1] << 32), b = (u64)x.v[2] | ((u64)x.v[3] << 32);
  u64 c = (u64)x.v[4] | ((u64)x.v[5] << 32), d = (u64)x.v[6] | ((u64)x.v[7] << 32);
  asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// c ? a : b, limb by limb (a reference-valued ?: makes the compiler spill both operands and index them through local memory)
__device__ __forceinline__ fe fe_sel(bool c, const fe &a, const fe &b) {
  fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
  return r;
}
__device__ __forceinline__ fe shfl_xor_fe(const fe &x, int m) {
  fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_xor_sync(0xffffffffu, x.v[i], m);
  return r;
}
// sum over the warp (modular adds); result valid in every lane
__device__ __forceinline__ fe warp_sum_fq(fe x) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) x = Fq::add(x, shfl_xor_fe(x, m));
  return x;
}
// ---- limb-column reductions -------------------------------------------------------------------------------
// Summing field elements across a warp with modular adds costs 5 dependent (shuffle x8, 8-limb carry chain,
// conditional subtract) stages per value.  Instead every 16-bit half of every limb is summed over the warp as an
// independent integer by the hardware warp reduction (redux.sync.add.u32: 32 lanes * 2^16 < 2^21, no overflow) —
// 16 REDUX per value instead of 80 SHFL + 40 64-bit adds — and carries are propagated / reduced mod p once at the
// end.  Every lane of the (fully converged) warp gets the sum.
struct col16 { u32 lo[8], hi[8]; };     // per-lane partial column sums of one value: < 2^32 / 32 each
__device__ __forceinline__ void col16_zero(col16 &c) {
#pragma unroll
  for (int i = 0; i < 8; i++) { c.lo[i] = 0; c.hi[i] = 0; }
}
__device__ __forceinline__ void col16_add(col16 &c, const fe &v) {          // up to 2^11 values per lane
#pragma unroll
  for (int i = 0; i < 8; i++) { c.lo[i] += v.v[i] & 0xffffu; c.hi[i] += v.v[i] >> 16; }
}
__device__ __forceinline__ fe col16_warp_sum(const col16 &c) {
  fe x;
  u64 carry = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const u32 lo = __reduce_add_sync(0xffffffffu, c.lo[i]), hi = __reduce_add_sync(0xffffffffu, c.hi[i]);
    carry += (u64)lo + ((u64)hi << 16);
    x.v[i] = (u32)carry; carry >>= 32;
  }
  // 8 limbs + a small top word, folded mod p
  u32 top = Fq::fold_top(x, (u32)carry);
  top = Fq::fold_top(x, top);
  cond_sub_p<FqParams>(x, top);
  cond_sub_p<FqParams>(x, 0);
  return x;
}
template <int NV>
__device__ __forceinline__ void warp_sum_fq_cols(fe (&x)[NV]) {
#pragma unroll
  for (int k = 0; k < NV; k++) {
    col16 c;
#pragma unroll
    for (int i = 0; i < 8; i++) { c.lo[i] = x[k].v[i] & 0xffffu; c.hi[i] = x[k].v[i] >> 16; }
    x[k] = col16_warp_sum(c);
  }
}
// sum of NV field elements per thread over the block; result in every lane of warp 0.  smem: NV * 32 fe.
template <int NV>
__device__ __forceinline__ void block_sum_fq(fe (&x)[NV], fe *smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  warp_sum_fq_cols<NV>(x);
  if (nw == 1) return;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) smem[k * 32 + warp] = x[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) x[k] = lane < nw ? smem[k * 32 + lane] : Fq::zero();
    warp_sum_fq_cols<NV>(x);
  }
}

}  // namespace sp2
