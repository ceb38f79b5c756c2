// T256 group arithmetic (y^2 = x^3 - 3x + b over the T256 base field Fp) for the commitment
// provider: Jacobian coordinates, a = -3 doubling, mixed and full addition with explicit identity /
// P+P / P+(-P) handling ("vartime" adds of the reference: src/provider/msm.rs:25-57 Bucket,
// add_mixed_vartime).  The curve arithmetic of the reference lives in halo2curves (un-vendored,
// Cargo.toml:40-45); any correct group law gives the same AFFINE result, which is what crosses the
// ABI and enters the transcript (src/provider/traits.rs:288-305).  Formulas: EFD dbl-2001-b,
// madd-2007-bl, add-2007-bl.  Single source for host unit tests and device (see prim.cuh).
#pragma once
#include "field.cuh"

namespace sp2 {

struct alignas(32) aff { fe x, y; };        // identity encoded as (0, 0) — (0,0) is not on the curve (b != 0)
struct alignas(32) jac { fe x, y, z; };     // identity: z == 0

SP2_HD bool aff_is_inf(const aff &p) { return Fp::is_zero(p.x) && Fp::is_zero(p.y); }
SP2_HD bool jac_is_inf(const jac &p) { return Fp::is_zero(p.z); }
SP2_HD jac jac_inf() { jac r; r.x = Fp::one(); r.y = Fp::one(); r.z = Fp::zero(); return r; }
SP2_HD jac jac_from_aff(const aff &a) {
  if (aff_is_inf(a)) return jac_inf();
  jac r; r.x = a.x; r.y = a.y; r.z = Fp::one(); return r;
}
SP2_HD aff aff_neg(const aff &p) { aff r; r.x = p.x; r.y = Fp::neg(p.y); return r; }   // neg(0) = 0 keeps the identity

// The formulas below are written in PAIRS of independent products, each pair one Fp::mul2 call (two multiplications in
// lockstep, field.cuh): a full addition is 8 call latencies instead of 16, a mixed addition 6 instead of 11, a doubling
// 4 instead of 8.

// 2P, a = -3: dbl-2001-b (3M + 5S)
SP2_HD jac jac_dbl(const jac &p) {
  if (jac_is_inf(p) || Fp::is_zero(p.y)) return jac_inf();
  const fe ypz = Fp::add(p.y, p.z);
  fe delta, gamma;
  Fp::mul2(p.z, p.z, p.y, p.y, delta, gamma);
  fe yz2, alpha;
  Fp::mul2(ypz, ypz, Fp::sub(p.x, delta), Fp::add(p.x, delta), yz2, alpha);
  fe beta, g2;
  Fp::mul2(p.x, gamma, gamma, gamma, beta, g2);
  alpha = Fp::add(Fp::dbl(alpha), alpha);
  const fe beta4 = Fp::dbl(Fp::dbl(beta));
  jac r;
  r.x = Fp::sub(Fp::sqr(alpha), Fp::dbl(beta4));
  r.z = Fp::sub(Fp::sub(yz2, gamma), delta);
  r.y = Fp::sub(Fp::mul(alpha, Fp::sub(beta4, r.x)), Fp::dbl(Fp::dbl(Fp::dbl(g2))));
  return r;
}

// P + Q, Q affine: madd-2007-bl (7M + 4S)
SP2_HD jac jac_add_mixed(const jac &p, const aff &q) {
  if (aff_is_inf(q)) return p;
  if (jac_is_inf(p)) return jac_from_aff(q);
  fe z1z1, t;
  Fp::mul2(p.z, p.z, q.y, p.z, z1z1, t);
  fe u2, s2;
  Fp::mul2(q.x, z1z1, t, z1z1, u2, s2);
  const fe h = Fp::sub(u2, p.x);
  fe rr = Fp::sub(s2, p.y);
  if (Fp::is_zero(h)) return Fp::is_zero(rr) ? jac_dbl(p) : jac_inf();
  rr = Fp::dbl(rr);
  const fe zh = Fp::add(p.z, h);
  fe hh, zh2;
  Fp::mul2(h, h, zh, zh, hh, zh2);
  const fe i = Fp::dbl(Fp::dbl(hh));
  fe j, v;
  Fp::mul2(h, i, p.x, i, j, v);
  fe r2, yj;
  Fp::mul2(rr, rr, p.y, j, r2, yj);
  jac r;
  r.x = Fp::sub(Fp::sub(Fp::sub(r2, j), v), v);
  r.y = Fp::sub(Fp::mul(Fp::sub(v, r.x), rr), Fp::dbl(yj));
  r.z = Fp::sub(Fp::sub(zh2, z1z1), hh);
  return r;
}

// P + Q: add-2007-bl (11M + 5S)
SP2_HD jac jac_add(const jac &p, const jac &q) {
  if (jac_is_inf(p)) return q;
  if (jac_is_inf(q)) return p;
  fe z1z1, z2z2, t1, t2;
  Fp::mul2(p.z, p.z, q.z, q.z, z1z1, z2z2);
  Fp::mul2(p.y, q.z, q.y, p.z, t1, t2);
  fe u1, u2, s1, s2;
  Fp::mul2(p.x, z2z2, q.x, z1z1, u1, u2);
  Fp::mul2(t1, z2z2, t2, z1z1, s1, s2);
  const fe h = Fp::sub(u2, u1);
  fe rr = Fp::sub(s2, s1);
  if (Fp::is_zero(h)) return Fp::is_zero(rr) ? jac_dbl(p) : jac_inf();
  rr = Fp::dbl(rr);
  const fe h2 = Fp::dbl(h), zs = Fp::add(p.z, q.z);
  fe i, zz;
  Fp::mul2(h2, h2, zs, zs, i, zz);
  const fe zc = Fp::sub(Fp::sub(zz, z1z1), z2z2);
  fe j, v, z3, r2;
  Fp::mul2(h, i, u1, i, j, v);
  Fp::mul2(zc, h, rr, rr, z3, r2);
  jac r;
  r.x = Fp::sub(Fp::sub(Fp::sub(r2, j), v), v);
  fe a, b;
  Fp::mul2(Fp::sub(v, r.x), rr, s1, j, a, b);
  r.y = Fp::sub(a, Fp::dbl(b));
  r.z = z3;
  return r;
}

// affine (x, y) = (X/Z^2, Y/Z^3) given zinv = 1/Z
SP2_HD aff jac_to_aff_with_inv(const jac &p, const fe &zinv) {
  aff r;
  if (jac_is_inf(p)) { r.x = Fp::zero(); r.y = Fp::zero(); return r; }
  const fe zi2 = Fp::sqr(zinv);
  r.x = Fp::mul(p.x, zi2);
  r.y = Fp::mul(p.y, Fp::mul(zi2, zinv));
  return r;
}
SP2_HD aff jac_to_aff(const jac &p) { return jac_to_aff_with_inv(p, Fp::inv(p.z)); }

// y^2 == x^3 - 3x + b
SP2_HD bool aff_on_curve(const aff &p) {
  fe b, a;
  for (int i = 0; i < 8; i++) { b.v[i] = CurveT256::B(i); a.v[i] = CurveT256::A(i); }
  const fe lhs = Fp::sqr(p.y);
  const fe rhs = Fp::add(Fp::add(Fp::mul(Fp::sqr(p.x), p.x), Fp::mul(a, p.x)), b);
  return Fp::eq(lhs, rhs);
}

}  // namespace sp2
