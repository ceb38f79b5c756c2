// 256-bit prime-field arithmetic for sm_100a: 8 x u32 limbs in registers, Montgomery form with
// R = 2^256, canonical [0,p) results — bit-compatible with the reference's field elements
// (4 x u64 little-endian Montgomery limbs, reference src/big_num/montgomery.rs:14-22,
// src/big_num/macros.rs:59-73), so tables cross the C ABI without conversion.
//
//   Fq = T256 scalar field = NIST P-256 base prime p = 2^256 - 2^224 + 2^192 + 2^96 - 1.
//        -p^-1 mod 2^32 = 1, so the Montgomery quotient digit is the limb itself and m*p is
//        shifts/adds only: REDC costs no multiplications.  (SURVEY.md Appendix A.)
//   Fp = T256 base field (curve coordinates for the MSM), generic modulus: CIOS.
//
// Restates the arithmetic of reference src/big_num/limbs.rs:178-349 (wide multiply-accumulate),
// src/big_num/montgomery.rs:39-177 (REDC) and halo2curves' field ops on 32-bit integer pipes.
// Single source for host (unit tests) and device (see prim.cuh).
#pragma once
#include "prim.cuh"
#include "consts.cuh"

namespace sp2 {

struct alignas(32) fe { u32 v[8]; };     // one field element: 32 bytes, 256-bit loads/stores

// ------------------------------------------------------------------------------------------
// 256 x 256 -> 512 bit product.  Even/odd column accumulators so every a_j*b_i is one
// IMAD.WIDE with carry (mad.lo.cc + madc.hi.cc pair).
// ------------------------------------------------------------------------------------------
template <int LEN>
SP2_HD void mad_row4(u32 (&acc)[LEN], const int k, u32 x0, u32 x1, u32 x2, u32 x3, u32 b) {
  acc[k + 0] = mad_lo_cc(x0, b, acc[k + 0]);
  acc[k + 1] = madc_hi_cc(x0, b, acc[k + 1]);
  acc[k + 2] = madc_lo_cc(x1, b, acc[k + 2]);
  acc[k + 3] = madc_hi_cc(x1, b, acc[k + 3]);
  acc[k + 4] = madc_lo_cc(x2, b, acc[k + 4]);
  acc[k + 5] = madc_hi_cc(x2, b, acc[k + 5]);
  acc[k + 6] = madc_lo_cc(x3, b, acc[k + 6]);
  acc[k + 7] = madc_hi_cc(x3, b, acc[k + 7]);
  if (k + 8 < LEN) acc[k + 8] = addc(acc[k + 8], 0);
}

SP2_HD void mul_wide(u32 (&r)[16], const fe &a, const fe &b) {
  u32 E[16], O[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    // even i: a_even*b_i -> even columns (E at i), a_odd*b_i -> odd columns (O index i)
    mad_row4<16>(E, i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
    mad_row4<16>(O, i, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
    // odd i+1: a_even*b -> odd columns (O index i), a_odd*b -> even columns (E at i+2)
    mad_row4<16>(O, i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i + 1]);
    mad_row4<16>(E, i + 2, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i + 1]);
  }
  r[0] = E[0];
  r[1] = add_cc(E[1], O[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) r[k] = addc_cc(E[k], O[k - 1]);
  r[15] = addc(E[15], O[14]);
}

// Two independent 256 x 256 products in LOCKSTEP: the carry chains of the two products alternate in program order.
// ptxas interleaves independent chains only within a short window (it does so for the even / odd chains of one product,
// not across two ~360-instruction multiplications placed one after the other — checked in SASS), so the instruction-
// level parallelism of a pair of independent products has to be laid out in the source.
SP2_HD void mul_wide2(u32 (&r0)[16], u32 (&r1)[16], const fe &a0, const fe &b0, const fe &a1, const fe &b1) {
  u32 E0[16], O0[16], E1[16], O1[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { E0[i] = 0; O0[i] = 0; E1[i] = 0; O1[i] = 0; }
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    mad_row4<16>(E0, i, a0.v[0], a0.v[2], a0.v[4], a0.v[6], b0.v[i]);
    mad_row4<16>(E1, i, a1.v[0], a1.v[2], a1.v[4], a1.v[6], b1.v[i]);
    mad_row4<16>(O0, i, a0.v[1], a0.v[3], a0.v[5], a0.v[7], b0.v[i]);
    mad_row4<16>(O1, i, a1.v[1], a1.v[3], a1.v[5], a1.v[7], b1.v[i]);
    mad_row4<16>(O0, i, a0.v[0], a0.v[2], a0.v[4], a0.v[6], b0.v[i + 1]);
    mad_row4<16>(O1, i, a1.v[0], a1.v[2], a1.v[4], a1.v[6], b1.v[i + 1]);
    mad_row4<16>(E0, i + 2, a0.v[1], a0.v[3], a0.v[5], a0.v[7], b0.v[i + 1]);
    mad_row4<16>(E1, i + 2, a1.v[1], a1.v[3], a1.v[5], a1.v[7], b1.v[i + 1]);
  }
  r0[0] = E0[0];
  r0[1] = add_cc(E0[1], O0[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) r0[k] = addc_cc(E0[k], O0[k - 1]);
  r0[15] = addc(E0[15], O0[14]);
  r1[0] = E1[0];
  r1[1] = add_cc(E1[1], O1[0]);
#pragma unroll
  for (int k = 2; k < 15; k++) r1[k] = addc_cc(E1[k], O1[k - 1]);
  r1[15] = addc(E1[15], O1[14]);
}

// ------------------------------------------------------------------------------------------
// Generic helpers parameterised by the modulus
// ------------------------------------------------------------------------------------------
template <class PR>
SP2_HD void cond_sub_p(fe &x, u32 top) {     // x (+ top*2^256) in [0, 2p)  ->  [0, p)
  u32 d[8];
  d[0] = sub_cc(x.v[0], PR::P(0));
#pragma unroll
  for (int i = 1; i < 8; i++) d[i] = subc_cc(x.v[i], PR::P(i));
  u32 b = subc(top, 0);                      // 0xffffffff iff x + top*2^256 < p
  if (b != 0xffffffffu) {
#pragma unroll
    for (int i = 0; i < 8; i++) x.v[i] = d[i];
  }
}

template <class PR>
struct Field {
  SP2_HD static fe zero() { fe r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
  SP2_HD static fe one() { fe r; for (int i = 0; i < 8; i++) r.v[i] = PR::ONE(i); return r; }
  SP2_HD static fe cst_r2() { fe r; for (int i = 0; i < 8; i++) r.v[i] = PR::R2(i); return r; }
  SP2_HD static fe cst_r3() { fe r; for (int i = 0; i < 8; i++) r.v[i] = PR::R3(i); return r; }
  SP2_HD static fe two_inv() { fe r; for (int i = 0; i < 8; i++) r.v[i] = PR::TWO_INV(i); return r; }
  SP2_HD static fe six_inv() { fe r; for (int i = 0; i < 8; i++) r.v[i] = PR::SIX_INV(i); return r; }
  SP2_HD static bool is_zero(const fe &a) { u32 o = 0; for (int i = 0; i < 8; i++) o |= a.v[i]; return o == 0; }
  SP2_HD static bool eq(const fe &a, const fe &b) { u32 o = 0; for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i]; return o == 0; }

  SP2_HD static fe add(const fe &a, const fe &b) {
    fe r;
    r.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = addc_cc(a.v[i], b.v[i]);
    u32 top = addc(0, 0);
    cond_sub_p<PR>(r, top);
    return r;
  }
  SP2_HD static fe sub(const fe &a, const fe &b) {
    fe r;
    r.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) r.v[i] = subc_cc(a.v[i], b.v[i]);
    u32 m = subc(0, 0);                      // 0xffffffff iff borrow
    r.v[0] = add_cc(r.v[0], PR::P(0) & m);
#pragma unroll
    for (int i = 1; i < 7; i++) r.v[i] = addc_cc(r.v[i], PR::P(i) & m);
    r.v[7] = addc(r.v[7], PR::P(7) & m);
    return r;
  }
  SP2_HD static fe dbl(const fe &a) { return add(a, a); }
  SP2_HD static fe neg(const fe &a) { return sub(zero(), a); }
  // a/2 : (a + (a odd ? p : 0)) >> 1
  SP2_HD static fe half(const fe &a) {
    u32 m = 0u - (a.v[0] & 1u);
    u32 t[9];
    t[0] = add_cc(a.v[0], PR::P(0) & m);
#pragma unroll
    for (int i = 1; i < 8; i++) t[i] = addc_cc(a.v[i], PR::P(i) & m);
    t[8] = addc(0, 0);
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = (t[i] >> 1) | (t[i + 1] << 31);
    return r;
  }
};

// ------------------------------------------------------------------------------------------
// Interleaved Montgomery multiplication (low latency): the eight reduction rows m_i * p are accumulated as further rows of
// the product, in the same even / odd carry-save form as mul_wide (each row = two independent 4-IMAD.WIDE carry chains
// into a big-integer accumulator; E takes the chains that start at an even limb position, O the odd ones), so no carry
// ever ripples further than its own row.  The only serial coupling between rounds is the quotient digit:
//   t_i = limb i of (E + O) incl. the carry c out of the retired low limbs;  m_i = t_i * (-p^-1) mod 2^32
// (two adds and one multiply per round).  Retiring limb i: E[i] + O[i] + c = 0 (mod 2^32) and < 2^33, hence the new
// c = 1 iff any of the three is non-zero.  ~240 instructions with 4-long dependency chains instead of ~360 with
// 17-long carry ripples per round: the latency-bound group arithmetic (MSM trees, scalar tails) runs on this.
// a, b < p  ->  a b / 2^256 mod p, canonical.
// ------------------------------------------------------------------------------------------
template <class PR>
SP2_HD fe mont_mul_interleaved(const fe &a, const fe &b) {
  u32 E[18], O[18];
#pragma unroll
  for (int i = 0; i < 18; i++) { E[i] = 0; O[i] = 0; }
  u32 c = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    if ((i & 1) == 0) {
      mad_row4<18>(E, i, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);
      mad_row4<18>(O, i, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);           // O[k] sits at limb position k + 1
      const u32 m = mul_lo(E[i] + (i ? O[i - 1] : 0u) + c, PR::INV32);
      mad_row4<18>(E, i, PR::P(0), PR::P(2), PR::P(4), PR::P(6), m);
      mad_row4<18>(O, i, PR::P(1), PR::P(3), PR::P(5), PR::P(7), m);
      c = ((E[i] | (i ? O[i - 1] : 0u) | c) != 0u) ? 1u : 0u;
    } else {
      mad_row4<18>(O, i - 1, a.v[0], a.v[2], a.v[4], a.v[6], b.v[i]);       // starts at limb i = O index i - 1
      mad_row4<18>(E, i + 1, a.v[1], a.v[3], a.v[5], a.v[7], b.v[i]);
      const u32 m = mul_lo(E[i] + O[i - 1] + c, PR::INV32);
      mad_row4<18>(O, i - 1, PR::P(0), PR::P(2), PR::P(4), PR::P(6), m);
      mad_row4<18>(E, i + 1, PR::P(1), PR::P(3), PR::P(5), PR::P(7), m);
      c = ((E[i] | O[i - 1] | c) != 0u) ? 1u : 0u;
    }
  }
  // (E + O + c 2^256) >> 256 : limbs 8..16  (O[k] is limb k + 1)
  fe r;
  r.v[0] = add_cc(E[8], O[7]);
#pragma unroll
  for (int k = 1; k < 8; k++) r.v[k] = addc_cc(E[8 + k], O[7 + k]);
  u32 top = addc(E[16], O[15]);
  r.v[0] = add_cc(r.v[0], c);
#pragma unroll
  for (int k = 1; k < 8; k++) r.v[k] = addc_cc(r.v[k], 0);
  top = addc(top, 0);
  cond_sub_p<PR>(r, top);
  return r;
}

// ------------------------------------------------------------------------------------------
// Fq: P-256 prime, multiplication-free REDC
// ------------------------------------------------------------------------------------------
struct Fq : Field<FqParams> {
  // Four 64-bit Montgomery rounds on t[0..NT-1] (NT >= 17; limbs above 15 absorb carries).
  // After the call t[0..7] are (logically) zero and the value/2^256 sits in t[8..NT-1].
  template <int NT>
  SP2_HD static void redc_rounds(u32 (&t)[NT]) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int i = 2 * k;
      const u32 m0 = t[i], m1 = t[i + 1];
      // + m*2^96 + m*2^192 + m*2^256
      t[i + 3] = add_cc(t[i + 3], m0);
      t[i + 4] = addc_cc(t[i + 4], m1);
      t[i + 5] = addc_cc(t[i + 5], 0);
      t[i + 6] = addc_cc(t[i + 6], m0);
      t[i + 7] = addc_cc(t[i + 7], m1);
      t[i + 8] = addc_cc(t[i + 8], m0);
      t[i + 9] = addc_cc(t[i + 9], m1);
#pragma unroll
      for (int j = i + 10; j < NT - 1; j++) t[j] = addc_cc(t[j], 0);
      t[NT - 1] = addc(t[NT - 1], 0);
      // - m*2^224   (the "- m" term cancels t[i], t[i+1])
      t[i + 7] = sub_cc(t[i + 7], m0);
      t[i + 8] = subc_cc(t[i + 8], m1);
#pragma unroll
      for (int j = i + 9; j < NT - 1; j++) t[j] = subc_cc(t[j], 0);
      t[NT - 1] = subc(t[NT - 1], 0);
    }
  }
  // t (16 limbs, < p*2^256)  ->  t / 2^256 mod p, canonical
  SP2_HD static fe redc16(const u32 (&w)[16]) {
    u32 t[17];
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = w[i];
    t[16] = 0;
    redc_rounds<17>(t);
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[8 + i];
    cond_sub_p<FqParams>(r, t[16]);
    return r;
  }
  SP2_HD static fe mul_inl(const fe &a, const fe &b) { u32 w[16]; mul_wide(w, a, b); return redc16(w); }
#if defined(__CUDA_ARCH__) && defined(SP2_FQ_OUTLINE)
  // Translation units of LATENCY-bound kernels (a few thousand threads, each running its straight-line body once) define
  // SP2_FQ_OUTLINE: the multiplication and the wide multiply-accumulate become out-of-line calls (operands and results in
  // registers), so a kernel is a few KB of code that stays in the instruction cache instead of >100 KB of inlined carry
  // chains fetched once per warp from L2 (ncu: 2-3 `no_instruction` stall cycles per issued instruction in k_nn_outer_round).
  struct w16 { u32 v[16]; };
  static __device__ __noinline__ fe mul_ni(fe a, fe b) { return mul_inl(a, b); }
  static __device__ __noinline__ w16 mul_wide_ni(fe a, fe b) { w16 r; mul_wide(r.v, a, b); return r; }
  static __device__ __forceinline__ fe mul(const fe &a, const fe &b) { return mul_ni(a, b); }
#else
  SP2_HD static fe mul(const fe &a, const fe &b) { return mul_inl(a, b); }
#endif
  SP2_HD static fe sqr(const fe &a) { return mul(a, a); }

  // x + hi*2^256 (hi < 2^32)  ==  x + hi*(2^224 - 2^192 - 2^96 + 1)  (mod p); returns the new top limb
  SP2_HD static u32 fold_top(fe &x, u32 hi) {
    u32 t[9];
    t[0] = add_cc(x.v[0], hi);
#pragma unroll
    for (int i = 1; i < 7; i++) t[i] = addc_cc(x.v[i], 0);
    t[7] = addc_cc(x.v[7], hi);
    t[8] = addc(0, 0);
    t[3] = sub_cc(t[3], hi);
    t[4] = subc_cc(t[4], 0);
    t[5] = subc_cc(t[5], 0);
    t[6] = subc_cc(t[6], hi);
    t[7] = subc_cc(t[7], 0);
    t[8] = subc(t[8], 0);
#pragma unroll
    for (int i = 0; i < 8; i++) x.v[i] = t[i];
    return t[8];
  }

  // ---- delayed reduction: 544-bit accumulator (WideLimbs<9> analogue, limbs.rs:69-89) ----
  struct acc { u32 v[17]; };
  SP2_HD static acc acc_zero() { acc a; for (int i = 0; i < 17; i++) a.v[i] = 0; return a; }
  // a += x*y  (unreduced_multiply_accumulate, delayed_reduction.rs:52-58); up to 2^32 products
  SP2_HD static void mul_acc(acc &a, const fe &x, const fe &y) {
#if defined(__CUDA_ARCH__) && defined(SP2_FQ_OUTLINE)
    const w16 ww = mul_wide_ni(x, y);
    const u32 (&w)[16] = ww.v;
#else
    u32 w[16];
    mul_wide(w, x, y);
#endif
    a.v[0] = add_cc(a.v[0], w[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) a.v[i] = addc_cc(a.v[i], w[i]);
    a.v[16] = addc(a.v[16], 0);
  }
  // a += x*y and b += x*y with one product
  SP2_HD static void mul_acc2(acc &a, acc &b, const fe &x, const fe &y) {
#if defined(__CUDA_ARCH__) && defined(SP2_FQ_OUTLINE)
    const w16 ww = mul_wide_ni(x, y);
    const u32 (&w)[16] = ww.v;
#else
    u32 w[16];
    mul_wide(w, x, y);
#endif
    a.v[0] = add_cc(a.v[0], w[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) a.v[i] = addc_cc(a.v[i], w[i]);
    a.v[16] = addc(a.v[16], 0);
    b.v[0] = add_cc(b.v[0], w[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) b.v[i] = addc_cc(b.v[i], w[i]);
    b.v[16] = addc(b.v[16], 0);
  }
  SP2_HD static void acc_add(acc &a, const acc &b) {
    a.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 16; i++) a.v[i] = addc_cc(a.v[i], b.v[i]);
    a.v[16] = addc(a.v[16], b.v[16]);
  }
  // reduce (montgomery_reduce_9 analogue, montgomery.rs:39-109): acc / 2^256 mod p, canonical
  SP2_HD static fe acc_reduce(const acc &a) {
    u32 t[18];
#pragma unroll
    for (int i = 0; i < 17; i++) t[i] = a.v[i];
    t[17] = 0;
    redc_rounds<18>(t);                       // value/2^256 < 2^288 + p : limbs t[8..17]
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[8 + i];
    // t[17] can only be 0 here (value < 2^289); fold the 9th limb (and the rare re-carries)
    u32 top = fold_top(r, t[16]);
    top = fold_top(r, top);
    top = fold_top(r, top);
    cond_sub_p<FqParams>(r, top);
    cond_sub_p<FqParams>(r, 0);
    return r;
  }
  // canonical integer <-> Montgomery
  SP2_HD static fe to_mont(const fe &raw) { return mul(raw, cst_r2()); }
  SP2_HD static fe from_mont(const fe &a) {
    u32 w[16];
#pragma unroll
    for (int i = 0; i < 8; i++) { w[i] = a.v[i]; w[8 + i] = 0; }
    return redc16(w);
  }
  // halo2curves from_uniform_bytes: 512-bit little-endian integer mod p (lo, hi: raw 256-bit halves)
  SP2_HD static fe from_uniform(const fe &lo, const fe &hi) { return add(mul(lo, cst_r2()), mul(hi, cst_r3())); }
  SP2_HD static fe inv(const fe &a) {         // Fermat, a^(p-2); inv(0) = 0
    fe r = one();
    for (int i = 255; i >= 0; i--) {
      r = sqr(r);
      u32 e = FqParams::P(i >> 5);
      if ((i >> 5) == 0) e -= 2;              // p - 2 (p's low limb is 0xffffffff: no borrow)
      if ((e >> (i & 31)) & 1) r = mul(r, a);
    }
    return r;
  }
};

// ------------------------------------------------------------------------------------------
// Fp: generic modulus (T256 base field), CIOS Montgomery multiplication
// ------------------------------------------------------------------------------------------
struct Fp : Field<FpParams> {
  // Montgomery multiplication = 256x256 wide product (IMAD.WIDE chains, shared with Fq) + word-by-word REDC that
  // exploits the shape of the T256 base modulus: p = 2^256 - 2^224 + 2^192 + 2^128 + c (c < 2^128), i.e. limbs
  // [p0,p1,p2,p3, 1, 0, 1, 2^32-1].  Per 32-bit round only m*p0..p3 are real multiplications (4 IMAD.WIDE, split
  // into an even and an odd carry chain); the upper limbs are m<<128, m<<192 and (m<<256) - (m<<224): adds/subs.
  // (The previous 32-bit CIOS used mad.hi, which SASS implements as the slow IMAD.HI.)
  // On the device the curve code calls the multiplication OUT OF LINE (operands and result in registers, no stack):
  // a Jacobian addition is 16 multiplications of ~360 instructions each, and with everything inlined the MSM kernels
  // were 280-300 KB of straight-line code executed once per tree level by a handful of warps — bound by instruction
  // fetch from L2, not by arithmetic (ncu: 115 us for a 7-level tree of 128 points).  One shared 5.8 KB body stays
  // resident in the instruction cache.
  // mul2: TWO INDEPENDENT products per out-of-line call, computed in lockstep (mul_inl2).  A point addition is a
  // dependency chain of 11-16 multiplications executed by few warps (the MSM kernels run at IPC ~0.3), and one
  // multiplication alone is a ~1100-cycle chain of dependent carry instructions; the independent products of a formula
  // level (e.g. Z1^2 and Z2^2) share one call whose interleaved chains fill the idle issue slots.  Body: ~12 KB, still
  // instruction-cache resident, unlike the fully inlined additions.
  struct fe2 { fe a, b; };
#if defined(__CUDA_ARCH__)
  static __device__ __noinline__ fe mul_ni(fe a, fe b) { return mul_inl(a, b); }
  static __device__ __forceinline__ fe mul(const fe &a, const fe &b) { return mul_ni(a, b); }
  static __device__ __noinline__ fe2 mul2_ni(fe a0, fe b0, fe a1, fe b1) { fe2 r; mul_inl2(a0, b0, a1, b1, r.a, r.b); return r; }
  static __device__ __forceinline__ void mul2(const fe &a0, const fe &b0, const fe &a1, const fe &b1, fe &o0, fe &o1) {
    const fe2 r = mul2_ni(a0, b0, a1, b1); o0 = r.a; o1 = r.b;
  }
#else
  SP2_HD static fe mul(const fe &a, const fe &b) { return mul_inl(a, b); }
  SP2_HD static void mul2(const fe &a0, const fe &b0, const fe &a1, const fe &b1, fe &o0, fe &o1) { mul_inl2(a0, b0, a1, b1, o0, o1); }
#endif
  // one word-by-word REDC round (see mul_inl) on t[i..17]
  SP2_HD static void redc_round(u32 (&t)[18], const int i) {
    const u32 m = mul_lo(t[i], FpParams::INV32);
    t[i + 0] = mad_lo_cc(m, FpParams::P(0), t[i + 0]);
    t[i + 1] = madc_hi_cc(m, FpParams::P(0), t[i + 1]);
    t[i + 2] = madc_lo_cc(m, FpParams::P(2), t[i + 2]);
    t[i + 3] = madc_hi_cc(m, FpParams::P(2), t[i + 3]);
    t[i + 4] = addc_cc(t[i + 4], m);
    t[i + 5] = addc_cc(t[i + 5], 0);
    t[i + 6] = addc_cc(t[i + 6], m);
    t[i + 7] = addc_cc(t[i + 7], 0);
    t[i + 8] = addc_cc(t[i + 8], m);
#pragma unroll
    for (int j = i + 9; j < 17; j++) t[j] = addc_cc(t[j], 0);
    t[17] = addc(t[17], 0);
    t[i + 1] = mad_lo_cc(m, FpParams::P(1), t[i + 1]);
    t[i + 2] = madc_hi_cc(m, FpParams::P(1), t[i + 2]);
    t[i + 3] = madc_lo_cc(m, FpParams::P(3), t[i + 3]);
    t[i + 4] = madc_hi_cc(m, FpParams::P(3), t[i + 4]);
#pragma unroll
    for (int j = i + 5; j < 17; j++) t[j] = addc_cc(t[j], 0);
    t[17] = addc(t[17], 0);
    t[i + 7] = sub_cc(t[i + 7], m);
#pragma unroll
    for (int j = i + 8; j < 17; j++) t[j] = subc_cc(t[j], 0);
    t[17] = subc(t[17], 0);
  }
  // two independent Montgomery products in lockstep (mul_wide2 + alternating REDC rounds)
  SP2_HD static void mul_inl2(const fe &a0, const fe &b0, const fe &a1, const fe &b1, fe &o0, fe &o1) {
    o0 = mont_mul_interleaved<FpParams>(a0, b0); o1 = mont_mul_interleaved<FpParams>(a1, b1);
  }
  SP2_HD static void mul_cios2(const fe &a0, const fe &b0, const fe &a1, const fe &b1, fe &o0, fe &o1) {
    u32 w0[16], w1[16];
    mul_wide2(w0, w1, a0, b0, a1, b1);
    u32 t0[18], t1[18];
#pragma unroll
    for (int i = 0; i < 16; i++) { t0[i] = w0[i]; t1[i] = w1[i]; }
    t0[16] = 0; t0[17] = 0; t1[16] = 0; t1[17] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { redc_round(t0, i); redc_round(t1, i); }
#pragma unroll
    for (int i = 0; i < 8; i++) { o0.v[i] = t0[8 + i]; o1.v[i] = t1[8 + i]; }
    cond_sub_p<FpParams>(o0, t0[16]);
    cond_sub_p<FpParams>(o1, t1[16]);
  }
  // Measured on B200 (tools/microbench/field_mul.cu, profiles/r2_b_microbench_field_mul.txt): the interleaved form runs at
  // 598 cycles per warp-multiplication per SMSP (62 G mul/s) against 711 (52 G mul/s) for the wide-product + shaped-CIOS form
  // below, 879 against 938 cycles of single-warp latency, in 280 instead of 408 instructions; the lockstep pair bought no
  // latency (1913 cycles per pair).  Both remain: mul_cios is the cross-check of the host tests.
  SP2_HD static fe mul_inl(const fe &a, const fe &b) { return mont_mul_interleaved<FpParams>(a, b); }
  SP2_HD static fe mul_cios(const fe &a, const fe &b) {
    u32 w[16];
    mul_wide(w, a, b);
    u32 t[18];
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = w[i];
    t[16] = 0; t[17] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const u32 m = mul_lo(t[i], FpParams::INV32);
      // even chain: m*p0 -> limbs i,i+1; m*p2 -> limbs i+2,i+3; then + m at limb i+4, + m at limb i+6, + m at limb i+8
      t[i + 0] = mad_lo_cc(m, FpParams::P(0), t[i + 0]);
      t[i + 1] = madc_hi_cc(m, FpParams::P(0), t[i + 1]);
      t[i + 2] = madc_lo_cc(m, FpParams::P(2), t[i + 2]);
      t[i + 3] = madc_hi_cc(m, FpParams::P(2), t[i + 3]);
      t[i + 4] = addc_cc(t[i + 4], m);
      t[i + 5] = addc_cc(t[i + 5], 0);
      t[i + 6] = addc_cc(t[i + 6], m);
      t[i + 7] = addc_cc(t[i + 7], 0);
      t[i + 8] = addc_cc(t[i + 8], m);
#pragma unroll
      for (int j = i + 9; j < 17; j++) t[j] = addc_cc(t[j], 0);
      t[17] = addc(t[17], 0);
      // odd chain: m*p1 -> limbs i+1,i+2; m*p3 -> limbs i+3,i+4
      t[i + 1] = mad_lo_cc(m, FpParams::P(1), t[i + 1]);
      t[i + 2] = madc_hi_cc(m, FpParams::P(1), t[i + 2]);
      t[i + 3] = madc_lo_cc(m, FpParams::P(3), t[i + 3]);
      t[i + 4] = madc_hi_cc(m, FpParams::P(3), t[i + 4]);
#pragma unroll
      for (int j = i + 5; j < 17; j++) t[j] = addc_cc(t[j], 0);
      t[17] = addc(t[17], 0);
      // - m at limb i+7   (the 2^32-1 limb: m*(2^32-1)<<224 = (m<<256) - (m<<224); the + part went in above)
      t[i + 7] = sub_cc(t[i + 7], m);
#pragma unroll
      for (int j = i + 8; j < 17; j++) t[j] = subc_cc(t[j], 0);
      t[17] = subc(t[17], 0);
    }
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[8 + i];
    cond_sub_p<FpParams>(r, t[16]);
    return r;
  }
  SP2_HD static fe sqr(const fe &a) { return mul(a, a); }
  SP2_HD static fe to_mont(const fe &raw) { return mul(raw, cst_r2()); }
  SP2_HD static fe from_mont(const fe &a) { fe o; for (int i = 0; i < 8; i++) o.v[i] = i == 0 ? 1u : 0u; return mul(a, o); }
  SP2_HD static fe inv(const fe &a) {
    fe r = one();
    // p - 2: low limb 0xb1c4b117 - 2, no borrow
    for (int i = 255; i >= 0; i--) {
      r = sqr(r);
      u32 e = FpParams::P(i >> 5);
      if ((i >> 5) == 0) e -= 2;
      if ((e >> (i & 31)) & 1) r = mul(r, a);
    }
    return r;
  }
};

}  // namespace sp2
