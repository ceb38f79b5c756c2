// Host-side scalar field (T256 scalar field, Montgomery limbs identical to the device's `fe`) for the per-round algebra of the
// drivers that keep the transcript on the host (nifs.cu, verifier.cu).
#pragma once
#include <string.h>
#include "field.cuh"
#include "host_transcript.h"

namespace sp2 {
// Host-side scalar field for the per-round algebra: the same Montgomery representation as `fe` (8 x u32 == 4 x u64
// little-endian), multiplied with 64-bit limbs / __int128 (host_transcript.h: mont_mul) — the 32-bit carry-chain
// emulation that field.cuh falls back to on the host costs ~1 us per multiplication, this ~40 ns.
struct HF {
  static void ld(const fe &a, uint64_t o[4]) { memcpy(o, a.v, 32); }
  static fe st(const uint64_t a[4]) { fe r; memcpy(r.v, a, 32); return r; }
  static fe zero() { return Fq::zero(); }
  static fe one() { return Fq::one(); }
  static bool is_zero(const fe &a) { return Fq::is_zero(a); }
  static bool eq(const fe &a, const fe &b) { return Fq::eq(a, b); }
  static fe add(const fe &x, const fe &y) {
    uint64_t a[4], b[4], s[4], d[4]; ld(x, a); ld(y, b);
    unsigned carry = 0, borrow = 0;
    for (int j = 0; j < 4; j++) { const sp2h::u128 t = (sp2h::u128)a[j] + b[j] + carry; s[j] = (uint64_t)t; carry = (unsigned)(t >> 64); }
    for (int j = 0; j < 4; j++) { const sp2h::u128 t = (sp2h::u128)s[j] - sp2h::FQ_MOD[j] - borrow; d[j] = (uint64_t)t; borrow = (unsigned)((t >> 64) & 1); }
    return st((carry || !borrow) ? d : s);
  }
  static fe sub(const fe &x, const fe &y) {
    uint64_t a[4], b[4], d[4]; ld(x, a); ld(y, b);
    unsigned borrow = 0;
    for (int j = 0; j < 4; j++) { const sp2h::u128 t = (sp2h::u128)a[j] - b[j] - borrow; d[j] = (uint64_t)t; borrow = (unsigned)((t >> 64) & 1); }
    if (borrow) { unsigned carry = 0; for (int j = 0; j < 4; j++) { const sp2h::u128 t = (sp2h::u128)d[j] + sp2h::FQ_MOD[j] + carry; d[j] = (uint64_t)t; carry = (unsigned)(t >> 64); } }
    return st(d);
  }
  static fe dbl(const fe &x) { return add(x, x); }
  static fe mul(const fe &x, const fe &y) { uint64_t a[4], b[4], o[4]; ld(x, a); ld(y, b); sp2h::mont_mul(a, b, sp2h::FQ_MOD, sp2h::FQ_INV, o); return st(o); }
  static fe sqr(const fe &x) { return mul(x, x); }
  static fe inv(const fe &x) {              // Fermat, x^(q-2); inv(0) = 0
    const uint64_t e[4] = {sp2h::FQ_MOD[0] - 2, sp2h::FQ_MOD[1], sp2h::FQ_MOD[2], sp2h::FQ_MOD[3]};
    fe r = one();
    for (int i = 255; i >= 0; i--) { r = sqr(r); if ((e[i >> 6] >> (i & 63)) & 1) r = mul(r, x); }
    return r;
  }
  static fe two_inv() { return Fq::two_inv(); }
  static fe six_inv() { return Fq::six_inv(); }
};
struct NnHost {
  static fe load(const uint64_t *p) { fe r; memcpy(r.v, p, 32); return r; }
  static void store(uint64_t *p, const fe &x) { memcpy(p, x.v, 32); }
  static fe eval(const fe *c, int n, const fe &r) { fe acc = c[n - 1]; for (int i = n - 2; i >= 0; i--) acc = HF::add(HF::mul(acc, r), c[i]); return acc; }
  // UniPoly::from_evals (src/polys/univariate.rs:84-120): evaluations at 0, 1, 2[, 3] -> coefficients low to high
  static void from_evals3(const fe &e0, const fe &e1, const fe &e2, fe *c) {
    c[2] = HF::mul(HF::add(HF::sub(e2, HF::dbl(e1)), e0), HF::two_inv());
    c[1] = HF::sub(HF::sub(e1, e0), c[2]);
    c[0] = e0;
  }
  static void from_evals4(const fe &e0, const fe &e1, const fe &e2, const fe &e3, fe *c) {
    const fe t1 = HF::add(HF::dbl(e1), e1), t2 = HF::add(HF::dbl(e2), e2);
    c[3] = HF::mul(HF::sub(HF::add(HF::sub(e3, t2), t1), e0), HF::six_inv());
    c[2] = HF::sub(HF::mul(HF::add(HF::sub(e2, HF::dbl(e1)), e0), HF::two_inv()), HF::add(HF::dbl(c[3]), c[3]));
    c[1] = HF::sub(HF::sub(HF::sub(e1, e0), c[2]), c[3]);
    c[0] = e0;
  }
};

}  // namespace sp2
