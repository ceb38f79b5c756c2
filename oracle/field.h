/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement of the reference's 256-bit Montgomery field layer:
 *   - 4x4 limb schoolbook multiply           (reference src/big_num/limbs.rs:178-194)
 *   - fused multiply-accumulate into 9 limbs (reference src/big_num/limbs.rs:335-349)
 *   - 9->8 limb fold + 4-round REDC          (reference src/big_num/montgomery.rs:39-177)
 *   - delayed-reduction accumulator          (reference src/big_num/delayed_reduction.rs:52-64)
 * The per-element field operations (add/sub/mul/invert) live in the un-vendored
 * dependency halo2curves 0.10.0 (reference Cargo.toml:40-45); their published
 * algorithm (canonical Montgomery form, R = 2^256, values in [0,p)) is restated here.
 *
 * The modulus is a run-time parameter (fctx) so the same code instantiates the
 * T256 scalar field (benchmark engine), the T256 base field (curve coordinates)
 * and the Pallas scalar field (needed only for the reference's transcript KAT,
 * src/provider/keccak.rs:147-152).
 */
#ifndef ORACLE_FIELD_H
#define ORACLE_FIELD_H
#include <stdint.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;          /* Montgomery form, canonical [0,p) */
typedef struct { uint64_t l[9]; } acc9;        /* WideLimbs<9> */

typedef struct {
  uint64_t mod[4];   /* MODULUS */
  uint64_t r1[4];    /* R_MOD   = 2^256 mod p (Montgomery ONE) */
  uint64_t r2[4];    /* R512_MOD = 2^512 mod p */
  uint64_t r3[4];    /* 2^768 mod p (for from_uniform) */
  uint64_t inv;      /* MONT_INV = -p^-1 mod 2^64 */
  int max_sub;       /* MAX_REDC_SUB_CORRECTIONS = floor(R/p) */
  fe two_inv;        /* TWO_INV (Montgomery form) */
} fctx;

static inline int f_gte4(const uint64_t *a, const uint64_t *b) {
  for (int i = 3; i >= 0; i--) { if (a[i] > b[i]) return 1; if (a[i] < b[i]) return 0; }
  return 1;
}
static inline uint64_t f_add4(uint64_t *r, const uint64_t *a, const uint64_t *b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
static inline uint64_t f_sub4(uint64_t *r, const uint64_t *a, const uint64_t *b) {
  uint64_t borrow = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t d = a[i] - b[i]; uint64_t b1 = a[i] < b[i];
    uint64_t d2 = d - borrow; uint64_t b2 = d < borrow;
    r[i] = d2; borrow = b1 | b2;
  }
  return borrow;
}

/* limbs.rs:178-194 */
static inline void f_mul_4_by_4(uint64_t *res, const uint64_t *a, const uint64_t *b) {
  memset(res, 0, 8 * sizeof(uint64_t));
  for (int i = 0; i < 4; i++) {
    u128 carry = 0;
    for (int j = 0; j < 4; j++) {
      u128 p = (u128)a[i] * b[j] + res[i + j] + carry;
      res[i + j] = (uint64_t)p; carry = p >> 64;
    }
    res[i + 4] = (uint64_t)carry;
  }
}

/* limbs.rs:335-349 (portable path) */
static inline void f_mul_acc(acc9 *acc, const fe *a, const fe *b) {
  uint64_t prod[8];
  f_mul_4_by_4(prod, a->l, b->l);
  u128 carry = 0;
  for (int i = 0; i < 8; i++) {
    u128 s = (u128)acc->l[i] + prod[i] + carry;
    acc->l[i] = (uint64_t)s; carry = s >> 64;
  }
  acc->l[8] += (uint64_t)carry;
}
static inline void f_acc_add(acc9 *a, const acc9 *b) {   /* limbs.rs:77-89 */
  u128 c = 0;
  for (int i = 0; i < 9; i++) { c += (u128)a->l[i] + b->l[i]; a->l[i] = (uint64_t)c; c >>= 64; }
}

/* montgomery.rs:122-177 */
static inline void f_reduce8(const fctx *F, uint64_t *out, const uint64_t *t) {
  uint64_t r[9];
  memcpy(r, t, 8 * sizeof(uint64_t)); r[8] = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t q = r[i] * F->inv;
    u128 carry = 0;
    for (int j = 0; j < 4; j++) {
      u128 p = (u128)q * F->mod[j] + r[i + j] + carry;
      r[i + j] = (uint64_t)p; carry = p >> 64;
    }
    for (int k = i + 4; k < 9 && carry; k++) {
      u128 s = (u128)r[k] + carry; r[k] = (uint64_t)s; carry = s >> 64;
    }
  }
  uint64_t x5[5] = { r[4], r[5], r[6], r[7], r[8] };
  if (x5[4] == 1) {                       /* sub_5_4 */
    uint64_t b = f_sub4(x5, x5, F->mod);
    x5[4] -= b;
  }
  for (int k = 0; k < F->max_sub; k++)
    if (f_gte4(x5, F->mod)) f_sub4(x5, x5, F->mod);
  memcpy(out, x5, 4 * sizeof(uint64_t));
}

/* montgomery.rs:39-109 */
static inline void f_reduce9(const fctx *F, fe *out, const acc9 *c) {
  uint64_t low8[8]; memcpy(low8, c->l, sizeof(low8));
  uint64_t h = c->l[8], fold_carry = 0;
  if (h) {
    u128 carry = 0;
    for (int i = 0; i < 4; i++) {
      u128 p = (u128)h * F->r2[i] + low8[i] + carry;
      low8[i] = (uint64_t)p; carry = p >> 64;
    }
    for (int i = 4; i < 8; i++) { u128 s = (u128)low8[i] + carry; low8[i] = (uint64_t)s; carry = s >> 64; }
    fold_carry = (uint64_t)carry;
  }
  f_reduce8(F, out->l, low8);
  if (fold_carry) {
    uint64_t cy = f_add4(out->l, out->l, F->r1);
    if (cy || f_gte4(out->l, F->mod)) f_sub4(out->l, out->l, F->mod);
  }
}

static inline void f_mul(const fctx *F, fe *r, const fe *a, const fe *b) {
  uint64_t t[8]; f_mul_4_by_4(t, a->l, b->l); f_reduce8(F, r->l, t);
}
static inline void f_sqr(const fctx *F, fe *r, const fe *a) { f_mul(F, r, a, a); }
static inline void f_add(const fctx *F, fe *r, const fe *a, const fe *b) {
  uint64_t c = f_add4(r->l, a->l, b->l);
  if (c || f_gte4(r->l, F->mod)) f_sub4(r->l, r->l, F->mod);
}
static inline void f_sub(const fctx *F, fe *r, const fe *a, const fe *b) {
  if (f_sub4(r->l, a->l, b->l)) f_add4(r->l, r->l, F->mod);
}
static inline void f_dbl(const fctx *F, fe *r, const fe *a) { f_add(F, r, a, a); }
static inline int f_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int f_eq(const fe *a, const fe *b) { return memcmp(a, b, sizeof(fe)) == 0; }
static inline void f_neg(const fctx *F, fe *r, const fe *a) {
  if (f_is_zero(a)) { *r = *a; return; }
  f_sub4(r->l, F->mod, a->l);
}
static inline void f_zero(fe *r) { memset(r, 0, sizeof(fe)); }
static inline void f_one(const fctx *F, fe *r) { memcpy(r->l, F->r1, 32); }

/* canonical integer (little-endian limbs) <-> Montgomery */
static inline void f_from_raw(const fctx *F, fe *r, const uint64_t *raw) {  /* raw < 2^256 */
  fe a, b; memcpy(a.l, raw, 32); memcpy(b.l, F->r2, 32); f_mul(F, r, &a, &b);
}
static inline void f_to_raw(const fctx *F, uint64_t *raw, const fe *a) {
  uint64_t t[8] = { a->l[0], a->l[1], a->l[2], a->l[3], 0, 0, 0, 0 };
  f_reduce8(F, raw, t);
}
static inline void f_from_u64(const fctx *F, fe *r, uint64_t v) {
  uint64_t raw[4] = { v, 0, 0, 0 }; f_from_raw(F, r, raw);
}
/* halo2curves from_uniform_bytes: 64 LE bytes as a 512-bit integer, mod p */
static inline void f_from_uniform(const fctx *F, fe *r, const uint8_t *b64) {
  fe lo, hi, t, c2, c3;
  memcpy(lo.l, b64, 32); memcpy(hi.l, b64 + 32, 32);
  memcpy(c2.l, F->r2, 32); memcpy(c3.l, F->r3, 32);
  f_mul(F, &t, &lo, &c2);           /* lo * R   */
  f_mul(F, &hi, &hi, &c3);          /* hi * R^2 = (hi * 2^256) * R */
  f_add(F, r, &t, &hi);
}
/* Fermat inversion; returns 0 if a == 0 */
static inline int f_inv(const fctx *F, fe *r, const fe *a) {
  if (f_is_zero(a)) { f_zero(r); return 0; }
  uint64_t e[4], two[4] = { 2, 0, 0, 0 };
  f_sub4(e, F->mod, two);
  fe acc; f_one(F, &acc);
  for (int i = 255; i >= 0; i--) {
    f_sqr(F, &acc, &acc);
    if ((e[i >> 6] >> (i & 63)) & 1) f_mul(F, &acc, &acc, a);
  }
  *r = acc; return 1;
}
static inline void f_pow(const fctx *F, fe *r, const fe *a, const uint64_t *e) {
  fe acc; f_one(F, &acc);
  for (int i = 255; i >= 0; i--) {
    f_sqr(F, &acc, &acc);
    if ((e[i >> 6] >> (i & 63)) & 1) f_mul(F, &acc, &acc, a);
  }
  *r = acc;
}

/* Derive all constants from the modulus (mirrors big_num/macros.rs compile-time derivation). */
static inline void f_ctx_init(fctx *F, const uint64_t mod[4]) {
  memcpy(F->mod, mod, 32);
  uint64_t inv = 1;                                  /* Newton: inv = p0^-1 mod 2^64 */
  for (int i = 0; i < 6; i++) inv *= 2 - mod[0] * inv;
  F->inv = (uint64_t)(0 - inv);
  /* 2^k mod p by repeated doubling of 1 */
  uint64_t x[4] = { 1, 0, 0, 0 };
  for (int k = 1; k <= 768; k++) {
    uint64_t top = x[3] >> 63;
    for (int i = 3; i > 0; i--) x[i] = (x[i] << 1) | (x[i - 1] >> 63);
    x[0] <<= 1;
    if (top || f_gte4(x, mod)) f_sub4(x, x, mod);
    if (k == 256) memcpy(F->r1, x, 32);
    if (k == 512) memcpy(F->r2, x, 32);
    if (k == 768) memcpy(F->r3, x, 32);
  }
  /* floor(2^256 / p): count how many times p fits below 2^256 */
  int q = 0; uint64_t s[4] = { 0, 0, 0, 0 };
  for (;;) { uint64_t c = f_add4(s, s, mod); if (c) break; q++; }
  F->max_sub = q;
  fe two; f_from_u64(F, &two, 2); f_inv(F, &F->two_inv, &two);
}
#endif
