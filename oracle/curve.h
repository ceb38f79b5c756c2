/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Short-Weierstrass group arithmetic y^2 = x^3 + a x + b over the T256 base field.
 * The curve arithmetic of the reference lives in the un-vendored dependency
 * halo2curves 0.10.0 (reference Cargo.toml:40-45; call sites src/provider/msm.rs:37-46,
 * 153,173 and src/provider/traits.rs:191-272).  Any correct group law yields the same
 * affine result, so the published Jacobian formulas (dbl-2007-bl, add-2007-bl,
 * madd-2007-bl from the EFD) are restated here with explicit handling of the identity,
 * P+P and P+(-P) cases (the reference's "vartime" adds branch on those).
 * Curve constants for T256 (a = -3, b, generator) are those verified in SURVEY.md §8c.
 */
#ifndef ORACLE_CURVE_H
#define ORACLE_CURVE_H
#include "field.h"

typedef struct { fe x, y, z; } pt;      /* Jacobian, identity: z == 0 */
typedef struct { fe x, y; } apt;        /* affine, identity encoded as (0,0) */

typedef struct { fctx F; fe a, b; } cctx;

static inline int apt_is_inf(const apt *p) { return f_is_zero(&p->x) && f_is_zero(&p->y); }
static inline int pt_is_inf(const pt *p) { return f_is_zero(&p->z); }
static inline void pt_set_inf(const cctx *C, pt *p) { f_one(&C->F, &p->x); f_one(&C->F, &p->y); f_zero(&p->z); }
static inline void pt_from_affine(const cctx *C, pt *r, const apt *a) {
  if (apt_is_inf(a)) { pt_set_inf(C, r); return; }
  r->x = a->x; r->y = a->y; f_one(&C->F, &r->z);
}

static inline void pt_dbl(const cctx *C, pt *r, const pt *p) {
  const fctx *F = &C->F;
  if (pt_is_inf(p) || f_is_zero(&p->y)) { pt_set_inf(C, r); return; }
  fe xx, yy, yyyy, zz, s, m, t, tmp;
  f_sqr(F, &xx, &p->x); f_sqr(F, &yy, &p->y); f_sqr(F, &yyyy, &yy); f_sqr(F, &zz, &p->z);
  f_add(F, &s, &p->x, &yy); f_sqr(F, &s, &s); f_sub(F, &s, &s, &xx); f_sub(F, &s, &s, &yyyy); f_dbl(F, &s, &s);
  f_dbl(F, &m, &xx); f_add(F, &m, &m, &xx);
  f_sqr(F, &tmp, &zz); f_mul(F, &tmp, &tmp, &C->a); f_add(F, &m, &m, &tmp);
  f_sqr(F, &t, &m); f_sub(F, &t, &t, &s); f_sub(F, &t, &t, &s);
  fe z3; f_add(F, &z3, &p->y, &p->z); f_sqr(F, &z3, &z3); f_sub(F, &z3, &z3, &yy); f_sub(F, &z3, &z3, &zz);
  fe y3; f_sub(F, &y3, &s, &t); f_mul(F, &y3, &y3, &m);
  f_dbl(F, &yyyy, &yyyy); f_dbl(F, &yyyy, &yyyy); f_dbl(F, &yyyy, &yyyy);
  f_sub(F, &y3, &y3, &yyyy);
  r->x = t; r->y = y3; r->z = z3;
}

static inline void pt_add(const cctx *C, pt *r, const pt *p, const pt *q) {
  const fctx *F = &C->F;
  if (pt_is_inf(p)) { *r = *q; return; }
  if (pt_is_inf(q)) { *r = *p; return; }
  fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, rr, v, t;
  f_sqr(F, &z1z1, &p->z); f_sqr(F, &z2z2, &q->z);
  f_mul(F, &u1, &p->x, &z2z2); f_mul(F, &u2, &q->x, &z1z1);
  f_mul(F, &s1, &p->y, &q->z); f_mul(F, &s1, &s1, &z2z2);
  f_mul(F, &s2, &q->y, &p->z); f_mul(F, &s2, &s2, &z1z1);
  f_sub(F, &h, &u2, &u1); f_sub(F, &rr, &s2, &s1);
  if (f_is_zero(&h)) { if (f_is_zero(&rr)) pt_dbl(C, r, p); else pt_set_inf(C, r); return; }
  f_dbl(F, &rr, &rr);
  f_dbl(F, &i, &h); f_sqr(F, &i, &i); f_mul(F, &j, &h, &i);
  f_mul(F, &v, &u1, &i);
  fe x3, y3, z3;
  f_sqr(F, &x3, &rr); f_sub(F, &x3, &x3, &j); f_sub(F, &x3, &x3, &v); f_sub(F, &x3, &x3, &v);
  f_sub(F, &y3, &v, &x3); f_mul(F, &y3, &y3, &rr);
  f_mul(F, &t, &s1, &j); f_dbl(F, &t, &t); f_sub(F, &y3, &y3, &t);
  f_add(F, &z3, &p->z, &q->z); f_sqr(F, &z3, &z3); f_sub(F, &z3, &z3, &z1z1); f_sub(F, &z3, &z3, &z2z2);
  f_mul(F, &z3, &z3, &h);
  r->x = x3; r->y = y3; r->z = z3;
}

static inline void pt_add_mixed(const cctx *C, pt *r, const pt *p, const apt *q) {
  const fctx *F = &C->F;
  if (apt_is_inf(q)) { *r = *p; return; }
  if (pt_is_inf(p)) { pt_from_affine(C, r, q); return; }
  fe z1z1, u2, s2, h, hh, i, j, rr, v, t;
  f_sqr(F, &z1z1, &p->z);
  f_mul(F, &u2, &q->x, &z1z1);
  f_mul(F, &s2, &q->y, &p->z); f_mul(F, &s2, &s2, &z1z1);
  f_sub(F, &h, &u2, &p->x); f_sub(F, &rr, &s2, &p->y);
  if (f_is_zero(&h)) { if (f_is_zero(&rr)) pt_dbl(C, r, p); else pt_set_inf(C, r); return; }
  f_dbl(F, &rr, &rr);
  f_sqr(F, &hh, &h); f_dbl(F, &i, &hh); f_dbl(F, &i, &i);
  f_mul(F, &j, &h, &i); f_mul(F, &v, &p->x, &i);
  fe x3, y3, z3;
  f_sqr(F, &x3, &rr); f_sub(F, &x3, &x3, &j); f_sub(F, &x3, &x3, &v); f_sub(F, &x3, &x3, &v);
  f_sub(F, &y3, &v, &x3); f_mul(F, &y3, &y3, &rr);
  f_mul(F, &t, &p->y, &j); f_dbl(F, &t, &t); f_sub(F, &y3, &y3, &t);
  f_add(F, &z3, &p->z, &h); f_sqr(F, &z3, &z3); f_sub(F, &z3, &z3, &z1z1); f_sub(F, &z3, &z3, &hh);
  r->x = x3; r->y = y3; r->z = z3;
}
static inline void apt_neg(const cctx *C, apt *r, const apt *p) { r->x = p->x; f_neg(&C->F, &r->y, &p->y); }
static inline void pt_neg(const cctx *C, pt *r, const pt *p) { r->x = p->x; r->z = p->z; f_neg(&C->F, &r->y, &p->y); }

static inline void pt_to_affine(const cctx *C, apt *r, const pt *p) {
  const fctx *F = &C->F;
  if (pt_is_inf(p)) { f_zero(&r->x); f_zero(&r->y); return; }
  fe zi, zi2, zi3; f_inv(F, &zi, &p->z); f_sqr(F, &zi2, &zi); f_mul(F, &zi3, &zi2, &zi);
  f_mul(F, &r->x, &p->x, &zi2); f_mul(F, &r->y, &p->y, &zi3);
}
static inline int apt_on_curve(const cctx *C, const apt *p) {
  const fctx *F = &C->F; fe l, r, t;
  f_sqr(F, &l, &p->y);
  f_sqr(F, &r, &p->x); f_mul(F, &r, &r, &p->x);
  f_mul(F, &t, &C->a, &p->x); f_add(F, &r, &r, &t); f_add(F, &r, &r, &C->b);
  return f_eq(&l, &r);
}
static inline int pt_eq(const cctx *C, const pt *p, const pt *q) {
  apt a, b; pt_to_affine(C, &a, p); pt_to_affine(C, &b, q);
  return f_eq(&a.x, &b.x) && f_eq(&a.y, &b.y);
}
/* variable-time double-and-add by a canonical 256-bit little-endian scalar */
static inline void pt_mul_raw(const cctx *C, pt *r, const pt *p, const uint64_t k[4]) {
  pt acc; pt_set_inf(C, &acc);
  for (int i = 255; i >= 0; i--) {
    pt_dbl(C, &acc, &acc);
    if ((k[i >> 6] >> (i & 63)) & 1) pt_add(C, &acc, &acc, p);
  }
  *r = acc;
}
#endif
