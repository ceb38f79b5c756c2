// Carry-chain primitives for 256-bit prime-field arithmetic on sm_100a.
//
// On the device every primitive is exactly one PTX instruction (add.cc / addc.cc / mad.lo.cc /
// madc.hi.cc ...); ptxas fuses a mad.lo.cc + madc.hi.cc pair on the same operands into one
// IMAD.WIDE.U32(.X).  When the same headers are compiled by a host compiler (unit tests of the
// kernel logic, tests/test_host_field.py) the primitives are emulated with an explicit carry
// flag, so the arithmetic above this layer is a single source for host and device.
//
// This is the B200 counterpart of the reference's only native code, the x86-64 BMI2/ADX
// multiply-accumulate (reference src/big_num/limbs.rs:200-331).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SP2_HD __host__ __device__ __forceinline__
#define SP2_D __device__ __forceinline__
#else
#define SP2_HD inline
#define SP2_D inline
#endif

namespace sp2 {
typedef uint32_t u32;
typedef uint64_t u64;

#if defined(__CUDA_ARCH__)
SP2_D u32 add_cc(u32 a, u32 b) { u32 r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SP2_D u32 addc_cc(u32 a, u32 b) { u32 r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SP2_D u32 addc(u32 a, u32 b) { u32 r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SP2_D u32 sub_cc(u32 a, u32 b) { u32 r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SP2_D u32 subc_cc(u32 a, u32 b) { u32 r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SP2_D u32 subc(u32 a, u32 b) { u32 r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SP2_D u32 mul_lo(u32 a, u32 b) { u32 r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SP2_D u32 mul_hi(u32 a, u32 b) { u32 r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SP2_D u32 mad_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SP2_D u32 madc_lo_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SP2_D u32 mad_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SP2_D u32 madc_hi_cc(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SP2_D u32 madc_hi(u32 a, u32 b, u32 c) { u32 r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// host emulation: one carry flag per thread, same semantics as the PTX CC.CF bit
static thread_local u32 g_cf = 0;
inline u32 add_cc(u32 a, u32 b) { u64 s = (u64)a + b; g_cf = (u32)(s >> 32); return (u32)s; }
inline u32 addc_cc(u32 a, u32 b) { u64 s = (u64)a + b + g_cf; g_cf = (u32)(s >> 32); return (u32)s; }
inline u32 addc(u32 a, u32 b) { return (u32)((u64)a + b + g_cf); }
inline u32 sub_cc(u32 a, u32 b) { u64 s = (u64)a - b; g_cf = (u32)((s >> 32) & 1); return (u32)s; }
inline u32 subc_cc(u32 a, u32 b) { u64 s = (u64)a - b - g_cf; g_cf = (u32)((s >> 32) & 1); return (u32)s; }
inline u32 subc(u32 a, u32 b) { return (u32)((u64)a - b - g_cf); }
inline u32 mul_lo(u32 a, u32 b) { return (u32)((u64)a * b); }
inline u32 mul_hi(u32 a, u32 b) { return (u32)(((u64)a * b) >> 32); }
inline u32 mad_lo_cc(u32 a, u32 b, u32 c) { u64 s = (u64)(u32)((u64)a * b) + c; g_cf = (u32)(s >> 32); return (u32)s; }
inline u32 madc_lo_cc(u32 a, u32 b, u32 c) { u64 s = (u64)(u32)((u64)a * b) + c + g_cf; g_cf = (u32)(s >> 32); return (u32)s; }
inline u32 mad_hi_cc(u32 a, u32 b, u32 c) { u64 s = (((u64)a * b) >> 32) + c; g_cf = (u32)(s >> 32); return (u32)s; }
inline u32 madc_hi_cc(u32 a, u32 b, u32 c) { u64 s = (((u64)a * b) >> 32) + c + g_cf; g_cf = (u32)(s >> 32); return (u32)s; }
inline u32 madc_hi(u32 a, u32 b, u32 c) { return (u32)((((u64)a * b) >> 32) + c + g_cf); }
#endif
// NOTE: on the device subtraction borrow uses the same CC.CF bit with inverted sense handled by
// the sub/subc instructions themselves; the host emulation keeps "1 = borrow" for sub chains and
// "1 = carry" for add chains, which is equivalent as long as add and sub chains are not mixed.
}  // namespace sp2
