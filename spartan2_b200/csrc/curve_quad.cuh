// Lane-parallel point addition for the latency-bound MSM trees (device only).
//
// One Fp multiplication is a ~900-cycle dependent carry chain for a single warp, and a Jacobian addition (add-2007-bl) is
// 16 of them in a row when one lane does it alone: a 5-level warp tree costs ~76k cycles although at level k only 32 / 2^k
// lanes have work.  Here an aligned QUAD of lanes shares one addition: the formula's multiplications are spread over the
// four lanes level by level (5 dependent multiplication levels instead of 16), operands travel by width-4 shuffles, and all
// four lanes end up with the same result.  Same group law as curve.cuh (any correct law gives the same AFFINE point, which is
// what leaves the device); the exceptional cases (identity operands, P + P, P + (-P)) are resolved by selects after the
// common path, P + P by falling back to jac_dbl — no divergence in the common path.
#pragma once
#include "curve.cuh"
#include "devutil.cuh"

namespace sp2 {

__device__ __forceinline__ fe shfl_quad_fe(const fe &v, int src) {      // value of lane `src` (0..3) of the caller's quad
  fe r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, v.v[i], src, 4);
  return r;
}
__device__ __forceinline__ jac jac_sel(bool c, const jac &a, const jac &b) {
  jac r; r.x = fe_sel(c, a.x, b.x); r.y = fe_sel(c, a.y, b.y); r.z = fe_sel(c, a.z, b.z); return r;
}
__device__ __forceinline__ jac shfl_idx_jac(const jac &p, int src) {
  jac r;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r.x.v[i] = __shfl_sync(0xffffffffu, p.x.v[i], src);
    r.y.v[i] = __shfl_sync(0xffffffffu, p.y.v[i], src);
    r.z.v[i] = __shfl_sync(0xffffffffu, p.z.v[i], src);
  }
  return r;
}

// p + q.  The whole (converged) warp calls; within every aligned quad the four lanes pass identical p and q and receive the
// identical sum.  Multiplication levels (lane g of the quad):
//   L1  g0: Z1^2        g1: Z2^2        g2: Y1 Z2        g3: Y2 Z1
//   L2  g0: U1=X1 Z2Z2  g1: U2=X2 Z1Z1  g2: S1=t1 Z2Z2   g3: S2=t2 Z1Z1      H = U2-U1, r = 2(S2-S1)
//   L3  g0: I=(2H)^2    g1: Z1 Z2       g2,g3: r^2
//   L4  g0: J=H I       g1: V=U1 I      g2,g3: (Z1 Z2) H                     X3 = r^2 - J - 2V, Z3 = 2 Z1 Z2 H
//   L5  even: r (V-X3)  odd: S1 J                                            Y3 = r(V-X3) - 2 S1 J
__device__ __forceinline__ jac quad_jac_add(const jac &p, const jac &q) {
  const int g = threadIdx.x & 3;
  const bool g0 = g == 0, g1 = g == 1, odd = (g & 1) != 0;
  // L1
  fe a = fe_sel(g0, p.z, fe_sel(g1, q.z, fe_sel(g == 2, p.y, q.y)));
  fe b = fe_sel(g0 || g == 3, p.z, q.z);
  const fe m1 = Fp::mul(a, b);
  // L2: the other operand's Z^2 sits in lane 1 (Z2Z2) for even lanes, lane 0 (Z1Z1) for odd lanes
  const fe zz = shfl_quad_fe(m1, odd ? 0 : 1);
  a = fe_sel(g0, p.x, fe_sel(g1, q.x, m1));
  const fe m2 = Fp::mul(a, zz);
  const fe u1 = shfl_quad_fe(m2, 0), s1 = shfl_quad_fe(m2, 2);
  const fe h = Fp::sub(shfl_quad_fe(m2, 1), u1);
  fe rr = Fp::sub(shfl_quad_fe(m2, 3), s1);
  const bool h_zero = Fp::is_zero(h), r_zero = Fp::is_zero(rr);
  rr = Fp::dbl(rr);
  // L3
  const fe h2 = Fp::dbl(h);
  a = fe_sel(g0, h2, fe_sel(g1, p.z, rr));
  b = fe_sel(g0, h2, fe_sel(g1, q.z, rr));
  const fe m3 = Fp::mul(a, b);
  const fe i_ = shfl_quad_fe(m3, 0), z12 = shfl_quad_fe(m3, 1), r2 = shfl_quad_fe(m3, 2);
  // L4
  a = fe_sel(g0, h, fe_sel(g1, u1, z12));
  b = fe_sel(g < 2, i_, h);
  const fe m4 = Fp::mul(a, b);
  const fe j = shfl_quad_fe(m4, 0), v = shfl_quad_fe(m4, 1), zh = shfl_quad_fe(m4, 2);
  jac r;
  r.x = Fp::sub(Fp::sub(Fp::sub(r2, j), v), v);
  // L5
  a = fe_sel(odd, s1, rr);
  b = fe_sel(odd, j, Fp::sub(v, r.x));
  const fe m5 = Fp::mul(a, b);
  r.y = Fp::sub(shfl_quad_fe(m5, 0), Fp::dbl(shfl_quad_fe(m5, 1)));
  r.z = Fp::dbl(zh);
  // exceptional cases (quad-uniform)
  const bool pinf = jac_is_inf(p), qinf = jac_is_inf(q);
  r = jac_sel(pinf, q, jac_sel(qinf, p, r));
  const bool special = !pinf && !qinf && h_zero;
  if (__any_sync(0xffffffffu, special)) {
    if (special) r = r_zero ? jac_dbl(p) : jac_inf();
  }
  return r;
}

// affine + affine -> Jacobian (mmadd-2007-bl, 4M + 2S): the first addition of a gather lane, whose operands are both table
// entries.  Exceptional inputs fall back to the general mixed addition.
__device__ __forceinline__ jac aff_add_to_jac(const aff &p, const aff &q) {
  const fe h = Fp::sub(q.x, p.x);
  if (aff_is_inf(p) || aff_is_inf(q) || Fp::is_zero(h)) return jac_add_mixed(jac_from_aff(p), q);
  const fe rr = Fp::dbl(Fp::sub(q.y, p.y));
  fe hh, r2;
  Fp::mul2(h, h, rr, rr, hh, r2);
  const fe i_ = Fp::dbl(Fp::dbl(hh));
  fe j, v;
  Fp::mul2(h, i_, p.x, i_, j, v);
  jac r;
  r.x = Fp::sub(Fp::sub(Fp::sub(r2, j), v), v);
  fe a, b;
  Fp::mul2(Fp::sub(v, r.x), rr, p.y, j, a, b);
  r.y = Fp::sub(a, Fp::dbl(b));
  r.z = Fp::dbl(h);
  return r;
}

// Sum of the 32 lanes' points: 6 quad additions (levels 1+2 pair lanes (k, k+16) and (k+8, k+24) inside quad k, then three
// shuffle levels across quads) instead of 5 single-lane additions — 30 multiplication latencies instead of 80.
// Result valid in lanes 0..3.
__device__ __forceinline__ jac warp_sum_jac_quad(const jac &acc) {
  const int lane = threadIdx.x & 31, k = lane >> 2;
  jac s = jac_inf(), r0 = jac_inf();
#pragma unroll 1
  for (int it = 0; it < 6; it++) {
    jac a, b;
    if (it < 2) { a = shfl_idx_jac(acc, k + 8 * it); b = shfl_idx_jac(acc, k + 8 * it + 16); }
    else if (it == 2) { a = r0; b = s; }
    else {
      const int d = 16 >> (it - 3);
      a = s; b = shfl_idx_jac(s, (lane + d) & 31);
      if (lane + d >= 32) b = jac_inf();          // quads without a partner keep their (unused) value: no P + P fallback
    }
    const jac t = quad_jac_add(a, b);
    if (it == 0) r0 = t; else s = t;
  }
  return s;
}

}  // namespace sp2
