/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement (plain C + OpenMP where the reference uses rayon) of the Spartan2
 * prover hot path and of the verifier that checks it.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * The reference is Rust and there is no Rust toolchain in this image,
 * so the oracle cannot be validated against a run of the reference.  It is pinned by the
 * known-answer vectors the reference's own tests hold (Keccak-256 digest and transcript
 * challenges src/provider/keccak.rs:146-163; UniPoly interpolation
 * src/polys/univariate.rs:298-395; SpMV [25,9,4] src/r1cs/sparse.rs:637-653; eq/MLE small
 * cases src/polys/eq.rs:132-148, multilinear.rs:247-320; big_num property tests
 * src/big_num/) and by prove->verify self-consistency.  PARITY UNPINNED below that level:
 * no golden exists in the reference for any sum-check coefficient, commitment or proof
 * byte; generator derivation (halo2curves hash_to_curve) and T256 from_uniform are
 * restated from the published algorithms only.
 *
 * Every function cites the reference file:line it follows.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "field.h"
#include "keccak.h"
#include "curve.h"

#define EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------
 * Contexts: 0 = T256 scalar field (= P-256 base prime), 1 = T256 base field,
 *           2 = Pallas scalar field (transcript KAT only).   reference src/provider/pt256.rs:51-57
 * ---------------------------------------------------------------------------------- */
static fctx FQ, FPB, FPAL;
static cctx CV;
static int g_init = 0;
static int g_threads = 1;

static const uint64_t MOD_T256_SCALAR[4] = { 0xffffffffffffffffULL, 0x00000000ffffffffULL, 0x0ULL, 0xffffffff00000001ULL };
static const uint64_t MOD_T256_BASE[4]   = { 0x93135661b1c4b117ULL, 0x7e72b42b30e73177ULL, 0x1ULL, 0xffffffff00000001ULL };
static const uint64_t MOD_PALLAS_SCALAR[4] = { 0x8c46eb2100000001ULL, 0x224698fc0994a8ddULL, 0x0ULL, 0x4000000000000000ULL };
/* T256 curve: y^2 = x^3 - 3x + b  (verified there) */
static const uint64_t T256_B_RAW[4] = { 0x863e60f20219fc56ULL, 0x36b06aceeb354224ULL, 0x6fb552f8e21ed4acULL, 0xb441071b12f4a036ULL };

EXPORT void orc_init(void) {
  if (g_init) return;
  f_ctx_init(&FQ, MOD_T256_SCALAR);
  f_ctx_init(&FPB, MOD_T256_BASE);
  f_ctx_init(&FPAL, MOD_PALLAS_SCALAR);
  CV.F = FPB;
  fe three; f_from_u64(&FPB, &three, 3); f_neg(&FPB, &CV.a, &three);
  f_from_raw(&FPB, &CV.b, T256_B_RAW);
  g_init = 1;
}
EXPORT void orc_set_threads(int n) {
  g_threads = n < 1 ? 1 : n;
#ifdef _OPENMP
  omp_set_num_threads(g_threads);
#endif
}
EXPORT int orc_get_max_threads(void) {
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}
static const fctx *ctx_of(int id) { return id == 0 ? &FQ : id == 1 ? &FPB : &FPAL; }

/* ---- field test hooks (used by tests to pin the field layer against Python big ints) ---- */
EXPORT void orc_f_constants(int id, uint64_t *mod, uint64_t *r1, uint64_t *r2, uint64_t *inv, int *max_sub) {
  const fctx *F = ctx_of(id);
  memcpy(mod, F->mod, 32); memcpy(r1, F->r1, 32); memcpy(r2, F->r2, 32); *inv = F->inv; *max_sub = F->max_sub;
}
EXPORT void orc_f_mul(int id, const fe *a, const fe *b, fe *o, size_t n) { for (size_t i = 0; i < n; i++) f_mul(ctx_of(id), &o[i], &a[i], &b[i]); }
EXPORT void orc_f_add(int id, const fe *a, const fe *b, fe *o, size_t n) { for (size_t i = 0; i < n; i++) f_add(ctx_of(id), &o[i], &a[i], &b[i]); }
EXPORT void orc_f_sub(int id, const fe *a, const fe *b, fe *o, size_t n) { for (size_t i = 0; i < n; i++) f_sub(ctx_of(id), &o[i], &a[i], &b[i]); }
EXPORT void orc_f_inv(int id, const fe *a, fe *o, size_t n) { for (size_t i = 0; i < n; i++) f_inv(ctx_of(id), &o[i], &a[i]); }
EXPORT void orc_f_from_raw(int id, const uint64_t *raw, fe *o, size_t n) { for (size_t i = 0; i < n; i++) f_from_raw(ctx_of(id), &o[i], raw + 4 * i); }
EXPORT void orc_f_to_raw(int id, const fe *a, uint64_t *raw, size_t n) { for (size_t i = 0; i < n; i++) f_to_raw(ctx_of(id), raw + 4 * i, &a[i]); }
EXPORT void orc_f_from_uniform(int id, const uint8_t *b64, fe *o, size_t n) { for (size_t i = 0; i < n; i++) f_from_uniform(ctx_of(id), &o[i], b64 + 64 * i); }
/* sum_i a_i*b_i by delayed reduction (big_num/delayed_reduction.rs:70-94 test shape) */
EXPORT void orc_f_dot_delayed(int id, const fe *a, const fe *b, size_t n, fe *o) {
  acc9 acc; memset(&acc, 0, sizeof(acc));
  for (size_t i = 0; i < n; i++) f_mul_acc(&acc, &a[i], &b[i]);
  f_reduce9(ctx_of(id), o, &acc);
}
/* reduce an arbitrary 9-limb value (montgomery.rs R512 fold identity tests) */
EXPORT void orc_f_reduce9(int id, const uint64_t *limbs9, fe *o) {
  acc9 a; memcpy(a.l, limbs9, 72); f_reduce9(ctx_of(id), o, &a);
}
EXPORT void orc_keccak256(const uint8_t *in, size_t n, uint8_t *out) { keccak256(in, n, out); }

/* ---- transcript handles ---- */
EXPORT transcript *orc_ts_new(const char *label) { transcript *t = (transcript *)malloc(sizeof(transcript)); ts_new(t, label); return t; }
EXPORT void orc_ts_free(transcript *t) { ts_free(t); free(t); }
EXPORT void orc_ts_absorb_bytes(transcript *t, const char *label, const uint8_t *p, size_t n) { ts_absorb_bytes(t, label, p, n); }
EXPORT void orc_ts_absorb_scalars(transcript *t, int id, const char *label, const fe *v, size_t n) { ts_absorb_scalars(t, ctx_of(id), label, v, n); }
EXPORT void orc_ts_dom_sep(transcript *t, const char *b) { ts_dom_sep(t, b); }
EXPORT void orc_ts_squeeze(transcript *t, int id, const char *label, fe *out) { ts_squeeze(t, ctx_of(id), label, out); }
EXPORT void orc_ts_state(const transcript *t, uint8_t *state64, uint16_t *round) { memcpy(state64, t->state, 64); *round = t->round; }

/* point -> x_BE || y_BE (provider/traits.rs:288-305); the reference panics on identity */
static void apt_to_bytes(const apt *p, uint8_t out[64]) {
  fe_to_be_bytes(&FPB, &p->x, out); fe_to_be_bytes(&FPB, &p->y, out + 32);
}
static void ts_absorb_point(transcript *t, const char *label, const apt *p) {
  uint8_t b[64]; apt_to_bytes(p, b); ts_absorb_bytes(t, label, b, 64);
}
/* HyraxCommitment transcript bytes (hyrax_pc.rs:714-729) */
static void ts_absorb_commitment(transcript *t, const char *label, const apt *rows, size_t n) {
  ts_push(t, label, strlen(label));
  ts_push(t, "poly_commitment_begin", 21);
  for (size_t i = 0; i < n; i++) { uint8_t b[64]; apt_to_bytes(&rows[i], b); ts_push(t, b, 64); }
  ts_push(t, "poly_commitment_end", 19);
}
EXPORT void orc_ts_absorb_commitment(transcript *t, const char *label, const apt *rows, size_t n) { ts_absorb_commitment(t, label, rows, n); }

/* ------------------------------------------------------------------------------------
 * polys
 * ---------------------------------------------------------------------------------- */
/* EqPolynomial::evals_from_points (polys/eq.rs:59-117): MSB-first */
static void eq_evals(const fe *r, size_t k, fe *out) {
  f_one(&FQ, &out[0]);
  size_t size = 1;
  for (size_t t = k; t-- > 0;) {
    const fe *rv = &r[t];
    if (size >= 4096 && g_threads > 1) {
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < size; i++) { f_mul(&FQ, &out[size + i], &out[i], rv); f_sub(&FQ, &out[i], &out[i], &out[size + i]); }
    } else {
      for (size_t i = 0; i < size; i++) { f_mul(&FQ, &out[size + i], &out[i], rv); f_sub(&FQ, &out[i], &out[i], &out[size + i]); }
    }
    size *= 2;
  }
}
EXPORT void orc_eq_evals(const fe *r, size_t k, fe *out) { eq_evals(r, k, out); }

/* serial variant used by spartan.rs:320 (evals_from_points_into is serial in the reference) */
static void eq_evals_serial(const fe *r, size_t k, fe *out) {
  f_one(&FQ, &out[0]);
  size_t size = 1;
  for (size_t t = k; t-- > 0;) {
    for (size_t i = 0; i < size; i++) { f_mul(&FQ, &out[size + i], &out[i], &r[t]); f_sub(&FQ, &out[i], &out[i], &out[size + i]); }
    size *= 2;
  }
}

/* MultilinearPolynomial::bind_poly_var_top (polys/multilinear.rs:95-164); zero-prefix
 * bookkeeping (lo_eff/hi_eff) is an optimisation that does not change any value, the
 * oracle binds every pair. */
static void bind_top(fe *Z, size_t len, const fe *r) {
  size_t n = len / 2;
  if (n >= 4096 && g_threads > 1) {
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) { fe d; f_sub(&FQ, &d, &Z[n + i], &Z[i]); f_mul(&FQ, &d, &d, r); f_add(&FQ, &Z[i], &Z[i], &d); }
  } else {
    for (size_t i = 0; i < n; i++) { fe d; f_sub(&FQ, &d, &Z[n + i], &Z[i]); f_mul(&FQ, &d, &d, r); f_add(&FQ, &Z[i], &Z[i], &d); }
  }
}
EXPORT void orc_bind_top(fe *Z, size_t len, const fe *r) { bind_top(Z, len, r); }

/* UniPoly::from_evals_deg2 / deg3 (polys/univariate.rs:84-118); coefficients low->high */
static void unipoly_from_evals_deg2(const fe *e, fe *c) {
  fe a, t;
  f_dbl(&FQ, &t, &e[1]); f_sub(&FQ, &a, &e[0], &t); f_add(&FQ, &a, &a, &e[2]); f_mul(&FQ, &a, &a, &FQ.two_inv);
  c[0] = e[0];
  f_sub(&FQ, &c[1], &e[1], &c[0]); f_sub(&FQ, &c[1], &c[1], &a);
  c[2] = a;
}
static fe g_six_inv; static int g_six_init = 0;
static void unipoly_from_evals_deg3(const fe *e, fe *c) {
  if (!g_six_init) { fe six; f_from_u64(&FQ, &six, 6); f_inv(&FQ, &g_six_inv, &six); g_six_init = 1; }
  fe e1_3, e2_3, d3, a, d2, b, t;
  f_dbl(&FQ, &e1_3, &e[1]); f_add(&FQ, &e1_3, &e1_3, &e[1]);
  f_dbl(&FQ, &e2_3, &e[2]); f_add(&FQ, &e2_3, &e2_3, &e[2]);
  f_sub(&FQ, &d3, &e[3], &e2_3); f_add(&FQ, &d3, &d3, &e1_3); f_sub(&FQ, &d3, &d3, &e[0]);
  f_mul(&FQ, &a, &d3, &g_six_inv);
  f_dbl(&FQ, &t, &e[1]); f_sub(&FQ, &d2, &e[2], &t); f_add(&FQ, &d2, &d2, &e[0]);
  f_mul(&FQ, &b, &d2, &FQ.two_inv);
  f_dbl(&FQ, &t, &a); f_add(&FQ, &t, &t, &a); f_sub(&FQ, &b, &b, &t);
  c[0] = e[0];
  f_sub(&FQ, &c[1], &e[1], &c[0]); f_sub(&FQ, &c[1], &c[1], &b); f_sub(&FQ, &c[1], &c[1], &a);
  c[2] = b; c[3] = a;
}
/* UniPoly::evaluate (univariate.rs:136-144) */
static void unipoly_eval(const fe *c, int ncoef, const fe *r, fe *out) {
  fe ev = c[0], pw = *r, t;
  for (int i = 1; i < ncoef; i++) { f_mul(&FQ, &t, &pw, &c[i]); f_add(&FQ, &ev, &ev, &t); f_mul(&FQ, &pw, &pw, r); }
  *out = ev;
}
EXPORT void orc_unipoly_from_evals(const fe *e, int n, fe *c) { if (n == 3) unipoly_from_evals_deg2(e, c); else unipoly_from_evals_deg3(e, c); }
EXPORT void orc_unipoly_eval(const fe *c, int n, const fe *r, fe *o) { unipoly_eval(c, n, r, o); }

/* absorb b"p" + UniPoly transcript bytes: coefficients except the linear one, each to_repr()
 * little-endian (univariate.rs:182-190) */
static void ts_absorb_unipoly(transcript *t, const fe *c, int ncoef) {
  ts_push(t, "p", 1);
  for (int i = 0; i < ncoef; i++) { if (i == 1) continue; uint8_t b[32]; fe_to_le_bytes(&FQ, &c[i], b); ts_push(t, b, 32); }
}

/* ------------------------------------------------------------------------------------
 * sum-check  (sumcheck.rs)
 * ---------------------------------------------------------------------------------- */
typedef struct {                /* eq_sumcheck::EqSumCheckInstance (sumcheck.rs:934-1016) */
  size_t l, first_half, second_half, round;
  fe *taus; fe eval_eq_left;
  fe **eq_left; fe **eq_right;  /* all prefixes */
  fe *c0, *slope, *m1;          /* per-tau (1-tau, 2tau-1, 2-3tau) */
} eqinst;

static fe **eq_prefixes(const fe **taus, size_t len) {   /* compute_eq_polynomials, sumcheck.rs:960-979 */
  fe **res = (fe **)malloc((len + 1) * sizeof(fe *));
  res[0] = (fe *)malloc(sizeof(fe)); f_one(&FQ, &res[0][0]);
  for (size_t i = 0; i < len; i++) {
    size_t n = (size_t)1 << i;
    res[i + 1] = (fe *)malloc(2 * n * sizeof(fe));
    for (size_t j = 0; j < n; j++) {
      f_mul(&FQ, &res[i + 1][n + j], &res[i][j], taus[i]);
      f_sub(&FQ, &res[i + 1][j], &res[i][j], &res[i + 1][n + j]);
    }
  }
  return res;
}
static void eqinst_new(eqinst *E, const fe *taus, size_t l) {
  memset(E, 0, sizeof(*E));
  E->l = l; E->first_half = l / 2; E->second_half = l - E->first_half; E->round = 1;
  E->taus = (fe *)malloc(l * sizeof(fe)); memcpy(E->taus, taus, l * sizeof(fe));
  f_one(&FQ, &E->eval_eq_left);
  /* left: skip tau_0, reversed; right: reversed (sumcheck.rs:981-987) */
  size_t nl = E->first_half > 0 ? E->first_half - 1 : 0, nr = E->second_half;
  const fe **lt = (const fe **)malloc((nl + 1) * sizeof(fe *)), **rt = (const fe **)malloc((nr + 1) * sizeof(fe *));
  for (size_t i = 0; i < nl; i++) lt[i] = &E->taus[E->first_half - 1 - i];
  for (size_t i = 0; i < nr; i++) rt[i] = &E->taus[l - 1 - i];
  E->eq_left = eq_prefixes(lt, nl); E->eq_right = eq_prefixes(rt, nr);
  free(lt); free(rt);
  E->c0 = (fe *)malloc(l * sizeof(fe)); E->slope = (fe *)malloc(l * sizeof(fe)); E->m1 = (fe *)malloc(l * sizeof(fe));
  fe one; f_one(&FQ, &one);
  for (size_t i = 0; i < l; i++) {
    f_sub(&FQ, &E->c0[i], &one, &taus[i]);
    f_sub(&FQ, &E->slope[i], &taus[i], &E->c0[i]);
    f_sub(&FQ, &E->m1[i], &E->c0[i], &E->slope[i]);
  }
}
static void eqinst_free(eqinst *E) {
  size_t nl = E->first_half > 0 ? E->first_half - 1 : 0, nr = E->second_half;
  for (size_t i = 0; i <= nl; i++) free(E->eq_left[i]);
  for (size_t i = 0; i <= nr; i++) free(E->eq_right[i]);
  free(E->eq_left); free(E->eq_right); free(E->taus); free(E->c0); free(E->slope); free(E->m1);
}
/* Horner conversion shared by derive_from_claim / fallback (sumcheck.rs:1306-1323) */
static void cubic_points_from_s(const fe *s0, const fe *s1, const fe *slead, const fe *sm1, fe out[3]) {
  fe c1, c2, t, in2, c33, in3, mid3;
  f_sub(&FQ, &t, s1, sm1); f_mul(&FQ, &t, &t, &FQ.two_inv); f_sub(&FQ, &c1, &t, slead);
  f_add(&FQ, &t, s1, sm1); f_mul(&FQ, &t, &t, &FQ.two_inv); f_sub(&FQ, &c2, &t, s0);
  f_dbl(&FQ, &t, slead); f_add(&FQ, &in2, &c2, &t);
  f_dbl(&FQ, &t, &in2); f_add(&FQ, &t, &c1, &t); f_dbl(&FQ, &t, &t); f_add(&FQ, &out[1], s0, &t);
  f_dbl(&FQ, &c33, slead); f_add(&FQ, &c33, &c33, slead);
  f_add(&FQ, &in3, &c2, &c33);
  f_dbl(&FQ, &t, &in3); f_add(&FQ, &mid3, &c1, &t); f_add(&FQ, &mid3, &mid3, &in3);
  f_dbl(&FQ, &t, &mid3); f_add(&FQ, &t, s0, &t); f_add(&FQ, &out[2], &t, &mid3);
  out[0] = *s0;
}

/* evaluation_points_cubic_with_three_inputs (sumcheck.rs:1025-1156) + derive_from_claim
 * (:1277-1324) + fallback_three_inputs (:1327-1396).  Also returns the raw (t0, tinf). */
static void eq_cubic_points(const eqinst *E, const fe *A, const fe *B, const fe *C, size_t len,
                            const fe *claim, fe out[3], fe *t0_out, fe *tinf_out) {
  size_t half_p = len / 2;
  int in_first = E->round < E->first_half;
  acc9 tot0, toti; memset(&tot0, 0, sizeof(tot0)); memset(&toti, 0, sizeof(toti));
  if (in_first) {
    const fe *el = E->eq_left[E->first_half - E->round];
    const fe *er = E->eq_right[E->second_half];
    size_t out_len = (size_t)1 << (E->first_half - E->round), in_len = (size_t)1 << E->second_half;
#pragma omp parallel if (g_threads > 1)
    {
      acc9 o0, oi; memset(&o0, 0, sizeof(o0)); memset(&oi, 0, sizeof(oi));
#pragma omp for schedule(static) nowait
      for (size_t xo = 0; xo < out_len; xo++) {
        acc9 i0, ii; memset(&i0, 0, sizeof(i0)); memset(&ii, 0, sizeof(ii));
        for (size_t xi = 0; xi < in_len; xi++) {
          size_t id = (xo << E->second_half) | xi;
          fe t0e, tie, da, db;
          f_mul(&FQ, &t0e, &A[id], &B[id]); f_sub(&FQ, &t0e, &t0e, &C[id]);
          f_sub(&FQ, &da, &A[id + half_p], &A[id]); f_sub(&FQ, &db, &B[id + half_p], &B[id]); f_mul(&FQ, &tie, &da, &db);
          f_mul_acc(&i0, &er[xi], &t0e); f_mul_acc(&ii, &er[xi], &tie);
        }
        fe r0, ri; f_reduce9(&FQ, &r0, &i0); f_reduce9(&FQ, &ri, &ii);
        f_mul_acc(&o0, &el[xo], &r0); f_mul_acc(&oi, &el[xo], &ri);
      }
#pragma omp critical
      { f_acc_add(&tot0, &o0); f_acc_add(&toti, &oi); }
    }
  } else {
    const fe *er = E->eq_right[E->l - E->round];
#pragma omp parallel if (g_threads > 1 && half_p >= 1024)
    {
      acc9 o0, oi; memset(&o0, 0, sizeof(o0)); memset(&oi, 0, sizeof(oi));
#pragma omp for schedule(static) nowait
      for (size_t id = 0; id < half_p; id++) {
        fe t0e, tie, da, db;
        f_mul(&FQ, &t0e, &A[id], &B[id]); f_sub(&FQ, &t0e, &t0e, &C[id]);
        f_sub(&FQ, &da, &A[id + half_p], &A[id]); f_sub(&FQ, &db, &B[id + half_p], &B[id]); f_mul(&FQ, &tie, &da, &db);
        f_mul_acc(&o0, &er[id], &t0e); f_mul_acc(&oi, &er[id], &tie);
      }
#pragma omp critical
      { f_acc_add(&tot0, &o0); f_acc_add(&toti, &oi); }
    }
  }
  fe t0, tinf; f_reduce9(&FQ, &t0, &tot0); f_reduce9(&FQ, &tinf, &toti);
  if (t0_out) *t0_out = t0;
  if (tinf_out) *tinf_out = tinf;

  const fe *p = &E->eval_eq_left;
  const fe *eq0 = &E->c0[E->round - 1], *sl = &E->slope[E->round - 1], *em1 = &E->m1[E->round - 1];
  fe l0p, l1p, l1pinv, s0, s1, slead, tm1, sm1, t;
  f_mul(&FQ, &l0p, eq0, p);
  f_add(&FQ, &t, eq0, sl); f_mul(&FQ, &l1p, &t, p);
  f_mul(&FQ, &slead, sl, p); f_mul(&FQ, &slead, &slead, &tinf);
  if (f_inv(&FQ, &l1pinv, &l1p)) {                 /* derive_from_claim */
    fe t1;
    f_mul(&FQ, &s0, &l0p, &t0);
    f_sub(&FQ, &s1, claim, &s0);
    f_mul(&FQ, &t1, &s1, &l1pinv);
    f_dbl(&FQ, &tm1, &tinf); f_dbl(&FQ, &t, &t0); f_add(&FQ, &tm1, &tm1, &t); f_sub(&FQ, &tm1, &tm1, &t1);
    f_mul(&FQ, &sm1, em1, p); f_mul(&FQ, &sm1, &sm1, &tm1);
  } else {                                         /* fallback_three_inputs */
    fe acc; f_zero(&acc);
    for (size_t id = 0; id < half_p; id++) {
      fe e;
      if (in_first) {
        size_t xo = id >> E->second_half, xi = id & (((size_t)1 << E->second_half) - 1);
        f_mul(&FQ, &e, &E->eq_left[E->first_half - E->round][xo], &E->eq_right[E->second_half][xi]);
      } else e = E->eq_right[E->l - E->round][id];
      fe ma, mb, mc, v;
      f_dbl(&FQ, &ma, &A[id]); f_sub(&FQ, &ma, &ma, &A[id + half_p]);
      f_dbl(&FQ, &mb, &B[id]); f_sub(&FQ, &mb, &mb, &B[id + half_p]);
      f_dbl(&FQ, &mc, &C[id]); f_sub(&FQ, &mc, &mc, &C[id + half_p]);
      f_mul(&FQ, &v, &ma, &mb); f_sub(&FQ, &v, &v, &mc); f_mul(&FQ, &v, &v, &e); f_add(&FQ, &acc, &acc, &v);
    }
    tm1 = acc;
    f_mul(&FQ, &s0, eq0, p); f_mul(&FQ, &s0, &s0, &t0);
    f_sub(&FQ, &s1, claim, &s0);
    f_mul(&FQ, &sm1, em1, p); f_mul(&FQ, &sm1, &sm1, &tm1);
  }
  cubic_points_from_s(&s0, &s1, &slead, &sm1, out);
}
/* evaluation_points_zero_check_round0 (sumcheck.rs:1163-1271): first round of a zero-check (claim = 0, t(0) = 0 on a satisfied
 * instance): only t(inf) is summed, no Cz reads; returns (eval_0, eval_2, eval_3) through derive_from_claim / its fallback */
EXPORT void orc_zero_check_round0(const fe *taus, size_t l, const fe *A, const fe *B, fe out[3]) {
  eqinst E; eqinst_new(&E, taus, l);
  size_t len = (size_t)1 << l, half_p = len / 2;
  int in_first = E.round < E.first_half;
  acc9 tot; memset(&tot, 0, sizeof(tot));
  if (in_first) {
    const fe *el = E.eq_left[E.first_half - E.round], *er = E.eq_right[E.second_half];
    size_t out_len = (size_t)1 << (E.first_half - E.round), in_len = (size_t)1 << E.second_half;
    for (size_t xo = 0; xo < out_len; xo++) {
      acc9 in; memset(&in, 0, sizeof(in));
      for (size_t xi = 0; xi < in_len; xi++) {
        size_t id = (xo << E.second_half) | xi; fe da, db, t;
        f_sub(&FQ, &da, &A[id + half_p], &A[id]); f_sub(&FQ, &db, &B[id + half_p], &B[id]); f_mul(&FQ, &t, &da, &db);
        f_mul_acc(&in, &er[xi], &t);
      }
      fe red; f_reduce9(&FQ, &red, &in); f_mul_acc(&tot, &el[xo], &red);
    }
  } else {
    const fe *er = E.eq_right[E.l - E.round];
    for (size_t id = 0; id < half_p; id++) { fe da, db, t; f_sub(&FQ, &da, &A[id + half_p], &A[id]); f_sub(&FQ, &db, &B[id + half_p], &B[id]); f_mul(&FQ, &t, &da, &db); f_mul_acc(&tot, &er[id], &t); }
  }
  fe tinf, zero; f_reduce9(&FQ, &tinf, &tot); f_zero(&zero);
  const fe *p = &E.eval_eq_left, *eq0 = &E.c0[0], *sl = &E.slope[0], *em1 = &E.m1[0];
  fe l1p, l1pinv, s0 = zero, s1 = zero, slead, tm1, sm1, t;
  f_add(&FQ, &t, eq0, sl); f_mul(&FQ, &l1p, &t, p);
  f_mul(&FQ, &slead, sl, p); f_mul(&FQ, &slead, &slead, &tinf);
  if (f_inv(&FQ, &l1pinv, &l1p)) { fe t1; f_mul(&FQ, &t1, &s1, &l1pinv); f_dbl(&FQ, &tm1, &tinf); f_sub(&FQ, &tm1, &tm1, &t1); }
  else f_dbl(&FQ, &tm1, &tinf);                                   /* fallback: t(-1) = 2 t_inf */
  f_mul(&FQ, &sm1, em1, p); f_mul(&FQ, &sm1, &sm1, &tm1);
  cubic_points_from_s(&s0, &s1, &slead, &sm1, out);
  eqinst_free(&E);
}

static void eqinst_bound(eqinst *E, const fe *r) {          /* sumcheck.rs:1399-1405 */
  const fe *tau = &E->taus[E->round - 1];
  fe one, t, rt; f_one(&FQ, &one);
  f_sub(&FQ, &t, &one, tau); f_sub(&FQ, &t, &t, r);
  f_mul(&FQ, &rt, r, tau); f_dbl(&FQ, &rt, &rt); f_add(&FQ, &t, &t, &rt);
  f_mul(&FQ, &E->eval_eq_left, &E->eval_eq_left, &t);
  E->round++;
}

/* SumcheckProof::prove_cubic_with_three_inputs (sumcheck.rs:502-571).
 * polys_out: l x 4 full coefficients (low->high); r_out: l; claims_out: 3; t_out (optional): l x 2 raw sums */
EXPORT void orc_sumcheck_cubic_prove(const fe *claim, const fe *taus, size_t l, fe *A, fe *B, fe *C,
                                     transcript *ts, fe *polys_out, fe *r_out, fe *claims_out, fe *t_out) {
  eqinst E; eqinst_new(&E, taus, l);
  fe cl = *claim; size_t len = (size_t)1 << l;
  for (size_t round = 0; round < l; round++) {
    fe pts[3], ev[4], co[4], t0, ti;
    eq_cubic_points(&E, A, B, C, len, &cl, pts, &t0, &ti);
    if (t_out) { t_out[2 * round] = t0; t_out[2 * round + 1] = ti; }
    ev[0] = pts[0]; f_sub(&FQ, &ev[1], &cl, &pts[0]); ev[2] = pts[1]; ev[3] = pts[2];
    unipoly_from_evals_deg3(ev, co);
    ts_absorb_unipoly(ts, co, 4);
    fe r; ts_squeeze(ts, &FQ, "c", &r);
    r_out[round] = r; memcpy(&polys_out[4 * round], co, 4 * sizeof(fe));
    unipoly_eval(co, 4, &r, &cl);
    bind_top(A, len, &r); bind_top(B, len, &r); bind_top(C, len, &r);
    eqinst_bound(&E, &r);
    len /= 2;
  }
  claims_out[0] = A[0]; claims_out[1] = B[0]; claims_out[2] = C[0];
  eqinst_free(&E);
}

/* compute_eval_points_quad (sumcheck.rs:128-174) */
static void quad_points(const fe *A, const fe *B, size_t len, fe *e0, fe *tinf) {
  size_t n = len / 2;
  acc9 tot0, toti; memset(&tot0, 0, sizeof(tot0)); memset(&toti, 0, sizeof(toti));
#pragma omp parallel if (g_threads > 1 && n >= 1024)
  {
    acc9 a0, ai; memset(&a0, 0, sizeof(a0)); memset(&ai, 0, sizeof(ai));
#pragma omp for schedule(static) nowait
    for (size_t i = 0; i < n; i++) {
      fe da, db;
      f_mul_acc(&a0, &A[i], &B[i]);
      f_sub(&FQ, &da, &A[n + i], &A[i]); f_sub(&FQ, &db, &B[n + i], &B[i]);
      f_mul_acc(&ai, &da, &db);
    }
#pragma omp critical
    { f_acc_add(&tot0, &a0); f_acc_add(&toti, &ai); }
  }
  f_reduce9(&FQ, e0, &tot0); f_reduce9(&FQ, tinf, &toti);
}
/* one quadratic round message from (eval0, tinf, claim) (sumcheck.rs:205-216) */
static void quad_round_poly(const fe *e0, const fe *tinf, const fe *claim, fe *co) {
  fe ev[3], t3, t;
  f_dbl(&FQ, &t3, e0); f_add(&FQ, &t3, &t3, e0);
  f_dbl(&FQ, &t, claim); f_sub(&FQ, &t, &t, &t3); f_add(&FQ, &t, &t, tinf); f_add(&FQ, &ev[2], &t, tinf);
  ev[0] = *e0; f_sub(&FQ, &ev[1], claim, e0);
  unipoly_from_evals_deg2(ev, co);
}
/* SumcheckProof::prove_quad (sumcheck.rs:190-247). polys_out: rounds x 3 */
EXPORT void orc_sumcheck_quad_prove(const fe *claim, size_t rounds, fe *A, fe *B, transcript *ts,
                                    fe *polys_out, fe *r_out, fe *claims_out) {
  fe cl = *claim; size_t len = (size_t)1 << rounds;
  for (size_t round = 0; round < rounds; round++) {
    fe e0, ti, co[3];
    quad_points(A, B, len, &e0, &ti);
    quad_round_poly(&e0, &ti, &cl, co);
    ts_absorb_unipoly(ts, co, 3);
    fe r; ts_squeeze(ts, &FQ, "c", &r);
    r_out[round] = r; memcpy(&polys_out[3 * round], co, 3 * sizeof(fe));
    unipoly_eval(co, 3, &r, &cl);
    bind_top(A, len, &r); bind_top(B, len, &r);
    len /= 2;
  }
  claims_out[0] = A[0]; claims_out[1] = B[0];
}

/* SumcheckProof::verify (sumcheck.rs:67-114).  polys: rounds x (degree+1) FULL coefficients are
 * NOT trusted: only the compressed part (all but the linear term) is read, the linear term is
 * re-derived from the running claim exactly as CompressedUniPoly::decompress does
 * (univariate.rs:166-179). */
EXPORT int orc_sumcheck_verify(const fe *polys, size_t rounds, int degree, const fe *claim, transcript *ts,
                               fe *e_out, fe *r_out) {
  int nc = degree + 1; fe e = *claim;
  for (size_t i = 0; i < rounds; i++) {
    fe co[4]; const fe *pc = &polys[nc * i];
    fe lin; f_sub(&FQ, &lin, &e, &pc[0]); f_sub(&FQ, &lin, &lin, &pc[0]);
    for (int k = 2; k < nc; k++) f_sub(&FQ, &lin, &lin, &pc[k]);
    co[0] = pc[0]; co[1] = lin; for (int k = 2; k < nc; k++) co[k] = pc[k];
    ts_absorb_unipoly(ts, co, nc);
    fe r; ts_squeeze(ts, &FQ, "c", &r); r_out[i] = r;
    unipoly_eval(co, nc, &r, &e);
  }
  *e_out = e; return 0;
}

/* ------------------------------------------------------------------------------------
 * R1CS: classified SpMV, ABC builder, matrix MLE evaluation   (r1cs/sparse.rs, r1cs/mod.rs)
 * ---------------------------------------------------------------------------------- */
typedef struct {                      /* PrecomputedSparseMatrix (sparse.rs:29-45) */
  size_t num_rows, num_cols;
  uint32_t *off_up, *off_un, *off_sm, *off_ge;
  uint32_t *up_cols, *un_cols, *sm_cols, *ge_cols; int8_t *sm_coef; fe *ge_vals;
} pmat;
typedef struct { size_t n; uint32_t *rows, *cols; fe *vals; } fcoo;   /* FilteredSpmv (sparse.rs:305-380): COO of the columns >= first_col */
typedef struct {
  size_t num_cons, num_cons_unpadded, num_vars, num_shared, num_precommitted, num_rest, num_public, num_challenges;
  pmat M[3];
  fcoo F[3];
} shape;

static void pmat_build(pmat *P, size_t rows, size_t cols, const fe *data, const uint32_t *indices, const uint32_t *indptr) {
  /* from_sparse (sparse.rs:49-134) */
  memset(P, 0, sizeof(*P)); P->num_rows = rows; P->num_cols = cols;
  size_t nnz = indptr[rows];
  P->off_up = (uint32_t *)calloc(rows + 1, 4); P->off_un = (uint32_t *)calloc(rows + 1, 4);
  P->off_sm = (uint32_t *)calloc(rows + 1, 4); P->off_ge = (uint32_t *)calloc(rows + 1, 4);
  P->up_cols = (uint32_t *)malloc((nnz + 1) * 4); P->un_cols = (uint32_t *)malloc((nnz + 1) * 4);
  P->sm_cols = (uint32_t *)malloc((nnz + 1) * 4); P->ge_cols = (uint32_t *)malloc((nnz + 1) * 4);
  P->sm_coef = (int8_t *)malloc(nnz + 1); P->ge_vals = (fe *)malloc((nnz + 1) * sizeof(fe));
  fe one, neg1, sp[6], sn[6]; f_one(&FQ, &one); f_neg(&FQ, &neg1, &one);
  for (int k = 0; k < 6; k++) { f_from_u64(&FQ, &sp[k], (uint64_t)(k + 2)); f_neg(&FQ, &sn[k], &sp[k]); }
  uint32_t nup = 0, nun = 0, nsm = 0, nge = 0;
  for (size_t r = 0; r < rows; r++) {
    P->off_up[r] = nup; P->off_un[r] = nun; P->off_sm[r] = nsm; P->off_ge[r] = nge;
    for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) {
      const fe *v = &data[e]; uint32_t c = indices[e]; int done = 0;
      if (f_eq(v, &one)) { P->up_cols[nup++] = c; continue; }
      if (f_eq(v, &neg1)) { P->un_cols[nun++] = c; continue; }
      for (int k = 0; k < 6 && !done; k++) if (f_eq(v, &sp[k])) { P->sm_cols[nsm] = c; P->sm_coef[nsm++] = (int8_t)(k + 2); done = 1; }
      for (int k = 0; k < 6 && !done; k++) if (f_eq(v, &sn[k])) { P->sm_cols[nsm] = c; P->sm_coef[nsm++] = (int8_t)(-(k + 2)); done = 1; }
      if (!done) { P->ge_cols[nge] = c; P->ge_vals[nge++] = *v; }
    }
  }
  P->off_up[rows] = nup; P->off_un[rows] = nun; P->off_sm[rows] = nsm; P->off_ge[rows] = nge;
}
static void pmat_free(pmat *P) {
  free(P->off_up); free(P->off_un); free(P->off_sm); free(P->off_ge);
  free(P->up_cols); free(P->un_cols); free(P->sm_cols); free(P->ge_cols); free(P->sm_coef); free(P->ge_vals);
}
static inline void small_mul(int8_t coef, const fe *x, fe *out) {   /* sparse.rs:137-155 */
  int a = coef < 0 ? -coef : coef; fe d, r;
  switch (a) {
    case 2: f_dbl(&FQ, &r, x); break;
    case 3: f_dbl(&FQ, &r, x); f_add(&FQ, &r, &r, x); break;
    case 4: f_dbl(&FQ, &r, x); f_dbl(&FQ, &r, &r); break;
    case 5: f_dbl(&FQ, &r, x); f_dbl(&FQ, &r, &r); f_add(&FQ, &r, &r, x); break;
    case 6: f_dbl(&FQ, &d, x); f_dbl(&FQ, &r, &d); f_add(&FQ, &r, &r, &d); break;
    default: f_dbl(&FQ, &d, x); f_dbl(&FQ, &r, &d); f_add(&FQ, &r, &r, &d); f_add(&FQ, &r, &r, x); break;
  }
  if (coef < 0) f_neg(&FQ, out, &r); else *out = r;
}
static void pmat_row(const pmat *P, size_t row, const fe *v, fe *out) {  /* compute_row_single (sparse.rs:194-218) */
  fe sum, t; f_zero(&sum);
  for (uint32_t i = P->off_up[row]; i < P->off_up[row + 1]; i++) f_add(&FQ, &sum, &sum, &v[P->up_cols[i]]);
  for (uint32_t i = P->off_un[row]; i < P->off_un[row + 1]; i++) f_sub(&FQ, &sum, &sum, &v[P->un_cols[i]]);
  for (uint32_t i = P->off_sm[row]; i < P->off_sm[row + 1]; i++) { small_mul(P->sm_coef[i], &v[P->sm_cols[i]], &t); f_add(&FQ, &sum, &sum, &t); }
  for (uint32_t i = P->off_ge[row]; i < P->off_ge[row + 1]; i++) { f_mul(&FQ, &t, &P->ge_vals[i], &v[P->ge_cols[i]]); f_add(&FQ, &sum, &sum, &t); }
  *out = sum;
}
static void pmat_mulvec(const pmat *P, const fe *v, fe *out) {   /* multiply_vec (sparse.rs:221-233) */
#pragma omp parallel for schedule(dynamic, 1024) if (g_threads > 1 && P->num_rows > 4096)
  for (size_t r = 0; r < P->num_rows; r++) pmat_row(P, r, v, &out[r]);
}

/* SplitR1CSShape (r1cs/mod.rs:743-911): the caller passes already-padded CSR matrices. */
EXPORT shape *orc_shape_new(size_t num_cons, size_t num_cons_unpadded, size_t num_shared, size_t num_precommitted,
                            size_t num_rest, size_t num_public, size_t num_challenges,
                            const fe *dA, const uint32_t *iA, const uint32_t *pA,
                            const fe *dB, const uint32_t *iB, const uint32_t *pB,
                            const fe *dC, const uint32_t *iC, const uint32_t *pC) {
  shape *S = (shape *)calloc(1, sizeof(shape));
  S->num_cons = num_cons; S->num_cons_unpadded = num_cons_unpadded; S->num_shared = num_shared;
  S->num_precommitted = num_precommitted; S->num_rest = num_rest; S->num_public = num_public; S->num_challenges = num_challenges;
  S->num_vars = num_shared + num_precommitted + num_rest;
  size_t cols = S->num_vars + 1 + num_public + num_challenges;
  pmat_build(&S->M[0], num_cons, cols, dA, iA, pA);
  pmat_build(&S->M[1], num_cons, cols, dB, iB, pB);
  pmat_build(&S->M[2], num_cons, cols, dC, iC, pC);
  /* build_filtered (sparse.rs:305-364; SplitR1CSShape::precompute, r1cs/mod.rs:1059-1073): entries in the columns that prep_prove
   * does not know yet (rest witness, the constant 1, public IO, challenges) */
  const fe *ds[3] = {dA, dB, dC}; const uint32_t *is[3] = {iA, iB, iC}, *ps[3] = {pA, pB, pC};
  const uint32_t first_col = (uint32_t)(num_shared + num_precommitted);
  for (int k = 0; k < 3; k++) {
    size_t cnt = 0;
    for (uint32_t e = 0; e < ps[k][num_cons]; e++) if (is[k][e] >= first_col) cnt++;
    fcoo *F = &S->F[k]; F->n = cnt;
    F->rows = (uint32_t *)malloc((cnt + 1) * 4); F->cols = (uint32_t *)malloc((cnt + 1) * 4); F->vals = (fe *)malloc((cnt + 1) * sizeof(fe));
    size_t q = 0;
    for (size_t r = 0; r < num_cons; r++) for (uint32_t e = ps[k][r]; e < ps[k][r + 1]; e++) if (is[k][e] >= first_col) { F->rows[q] = (uint32_t)r; F->cols[q] = is[k][e]; F->vals[q++] = ds[k][e]; }
  }
  return S;
}
EXPORT void orc_shape_free(shape *S) { for (int k = 0; k < 3; k++) { pmat_free(&S->M[k]); free(S->F[k].rows); free(S->F[k].cols); free(S->F[k].vals); } free(S); }
/* multiply_vec_incremental_into (r1cs/mod.rs:1170-1211): out = cached (the products over the shared + precommitted columns,
 * computed by prep_prove) + FilteredSpmv::multiply_vec_add over the remaining columns (serial, sparse.rs:375-379) */
EXPORT void orc_shape_multiply_vec_incremental(const shape *S, const fe *z, const fe *caz, const fe *cbz, const fe *ccz, fe *az, fe *bz, fe *cz) {
  const fe *c[3] = {caz, cbz, ccz}; fe *o[3] = {az, bz, cz};
#pragma omp parallel for schedule(static, 1) if (g_threads > 1)          /* the three matrices in parallel (nested rayon::join) */
  for (int k = 0; k < 3; k++) {
    memcpy(o[k], c[k], S->num_cons * sizeof(fe));
    const fcoo *F = &S->F[k];
    for (size_t q = 0; q < F->n; q++) { fe t; f_mul(&FQ, &t, &F->vals[q], &z[F->cols[q]]); f_add(&FQ, &o[k][F->rows[q]], &o[k][F->rows[q]], &t); }
  }
}
/* SplitR1CSShape::multiply_vec (r1cs/mod.rs:1075-1107) */
EXPORT void orc_shape_multiply_vec(const shape *S, const fe *z, fe *az, fe *bz, fe *cz) {
  pmat_mulvec(&S->M[0], z, az); pmat_mulvec(&S->M[1], z, bz); pmat_mulvec(&S->M[2], z, cz);
}
/* plain CSR SpMV (SparseMatrix::multiply_vec, sparse.rs:476-520) for the [25,9,4] KAT */
EXPORT void orc_csr_multiply_vec(size_t rows, const fe *data, const uint32_t *indices, const uint32_t *indptr, const fe *z, fe *out) {
  for (size_t r = 0; r < rows; r++) {
    fe s, t; f_zero(&s);
    for (uint32_t e = indptr[r]; e < indptr[r + 1]; e++) { f_mul(&FQ, &t, &data[e], &z[indices[e]]); f_add(&FQ, &s, &s, &t); }
    out[r] = s;
  }
}

/* accumulate_rows (r1cs/mod.rs:1324-1398) */
static void abc_rows(const shape *S, const fe *rx, const fe *r, const fe *r2, size_t s, size_t e, fe *out) {
  for (size_t row = s; row < e; row++) {
    fe w[3], t; w[0] = rx[row]; f_mul(&FQ, &w[1], &rx[row], r); f_mul(&FQ, &w[2], &rx[row], r2);
    for (int k = 0; k < 3; k++) {
      const pmat *P = &S->M[k];
      for (uint32_t i = P->off_up[row]; i < P->off_up[row + 1]; i++) f_add(&FQ, &out[P->up_cols[i]], &out[P->up_cols[i]], &w[k]);
      for (uint32_t i = P->off_un[row]; i < P->off_un[row + 1]; i++) f_sub(&FQ, &out[P->un_cols[i]], &out[P->un_cols[i]], &w[k]);
      for (uint32_t i = P->off_sm[row]; i < P->off_sm[row + 1]; i++) { small_mul(P->sm_coef[i], &w[k], &t); f_add(&FQ, &out[P->sm_cols[i]], &out[P->sm_cols[i]], &t); }
      for (uint32_t i = P->off_ge[row]; i < P->off_ge[row + 1]; i++) { f_mul(&FQ, &t, &P->ge_vals[i], &w[k]); f_add(&FQ, &out[P->ge_cols[i]], &out[P->ge_cols[i]], &t); }
    }
  }
}
/* bind_and_prepare_poly_ABC_inner (r1cs/mod.rs:1272-1321); out_len = num_vars + num_extra */
EXPORT void orc_shape_abc(const shape *S, const fe *rx, const fe *r, fe *out, size_t out_len) {
  fe r2; f_mul(&FQ, &r2, r, r);
  size_t rows = S->num_cons_unpadded;
  memset(out, 0, out_len * sizeof(fe));
  if (g_threads <= 1 || rows <= 4096) { abc_rows(S, rx, r, &r2, 0, rows, out); return; }
  int T = g_threads; size_t chunk = (rows + T - 1) / T;
  fe *loc = (fe *)calloc((size_t)T * out_len, sizeof(fe));
#pragma omp parallel for schedule(static, 1)
  for (int t = 0; t < T; t++) {
    size_t s = t * chunk, e = s + chunk < rows ? s + chunk : rows;
    if (s < e) abc_rows(S, rx, r, &r2, s, e, loc + (size_t)t * out_len);
  }
#pragma omp parallel for schedule(static)
  for (size_t j = 0; j < out_len; j++) { fe a; f_zero(&a); for (int t = 0; t < T; t++) f_add(&FQ, &a, &a, &loc[(size_t)t * out_len + j]); out[j] = a; }
  free(loc);
}
/* evaluate_with_tables_fast (r1cs/mod.rs:36-146, 1216-1226) — verifier side */
static void row_dot_ty(const pmat *P, size_t row, const fe *ty, fe *out) { pmat_row(P, row, ty, out); }
EXPORT void orc_shape_eval_tables(const shape *S, const fe *tx, const fe *ty, fe *evals3) {
  for (int k = 0; k < 3; k++) {
    acc9 tot; memset(&tot, 0, sizeof(tot));
#pragma omp parallel if (g_threads > 1)
    {
      acc9 a; memset(&a, 0, sizeof(a));
#pragma omp for schedule(static) nowait
      for (size_t row = 0; row < S->num_cons; row++) { fe rs; row_dot_ty(&S->M[k], row, ty, &rs); f_mul_acc(&a, &tx[row], &rs); }
#pragma omp critical
      f_acc_add(&tot, &a);
    }
    f_reduce9(&FQ, &evals3[k], &tot);
  }
}

/* ------------------------------------------------------------------------------------
 * MSM family  (provider/msm.rs)
 * ---------------------------------------------------------------------------------- */
typedef struct { int kind; apt a; pt p; } bucket;   /* Bucket::{None,Affine,Projective} (msm.rs:25-57) */
static void bucket_add_assign(bucket *b, const apt *o) {
  if (b->kind == 0) { b->kind = 1; b->a = *o; }
  else if (b->kind == 1) { pt t; pt_from_affine(&CV, &t, &b->a); pt_add_mixed(&CV, &b->p, &t, o); b->kind = 2; }
  else pt_add_mixed(&CV, &b->p, &b->p, o);
}
static void bucket_add(const bucket *b, pt *acc) {   /* other + bucket */
  if (b->kind == 1) pt_add_mixed(&CV, acc, acc, &b->a);
  else if (b->kind == 2) pt_add(&CV, acc, acc, &b->p);
}
static size_t get_at(size_t segment, size_t c, const uint8_t bytes[32]) {   /* msm.rs:68-86 */
  size_t skip_bits = segment * c, skip_bytes = skip_bits / 8;
  if (skip_bytes >= 32) return 0;
  uint8_t v[8] = { 0 };
  for (size_t i = 0; i < 8 && skip_bytes + i < 32; i++) v[i] = bytes[skip_bytes + i];
  uint64_t tmp; memcpy(&tmp, v, 8);
  tmp >>= skip_bits - skip_bytes * 8;
  tmp %= (uint64_t)1 << c;
  return (size_t)tmp;
}
/* cpu_msm_serial (msm.rs:59-178) */
static void msm_serial(const fe *coeffs, const apt *bases, size_t nin, pt *out) {
  size_t c = nin < 4 ? 1 : nin < 32 ? 3 : (size_t)ceil(log((double)nin));
  pt boolean_sum; pt_set_inf(&CV, &boolean_sum);
  uint8_t (*reprs)[32] = (uint8_t (*)[32])malloc((nin + 1) * 32);
  apt *nb = (apt *)malloc((nin + 1) * sizeof(apt));
  fe one; f_one(&FQ, &one); size_t n = 0;
  for (size_t i = 0; i < nin; i++) {
    if (f_eq(&coeffs[i], &one)) pt_add_mixed(&CV, &boolean_sum, &boolean_sum, &bases[i]);
    else if (!f_is_zero(&coeffs[i])) { fe_to_le_bytes(&FQ, &coeffs[i], reprs[n]); nb[n++] = bases[i]; }
  }
  if (n == 0) { *out = boolean_sum; free(reprs); free(nb); return; }
  size_t segments = 256 / c + 1, half = (size_t)1 << (c - 1), full = (size_t)1 << c;
  int16_t *digits = (int16_t *)calloc((segments + 1) * n, sizeof(int16_t));
  uint8_t *carry = (uint8_t *)calloc(n, 1);
  for (size_t seg = 0; seg < segments; seg++)
    for (size_t j = 0; j < n; j++) {
      size_t raw = get_at(seg, c, reprs[j]) + carry[j]; carry[j] = 0;
      if (raw >= half) { digits[seg * n + j] = (int16_t)(-(int)(full - raw)); carry[j] = 1; }
      else digits[seg * n + j] = (int16_t)raw;
    }
  size_t total = segments; int any = 0;
  for (size_t j = 0; j < n; j++) any |= carry[j];
  if (any) { for (size_t j = 0; j < n; j++) digits[segments * n + j] = carry[j]; total = segments + 1; }
  bucket *buckets = (bucket *)malloc(half * sizeof(bucket));
  pt acc; pt_set_inf(&CV, &acc);
  for (size_t seg = total; seg-- > 0;) {
    for (size_t k = 0; k < c; k++) pt_dbl(&CV, &acc, &acc);
    for (size_t b = 0; b < half; b++) buckets[b].kind = 0;
    for (size_t j = 0; j < n; j++) {
      int d = digits[seg * n + j];
      if (d > 0) bucket_add_assign(&buckets[d - 1], &nb[j]);
      else if (d < 0) { apt neg; apt_neg(&CV, &neg, &nb[j]); bucket_add_assign(&buckets[-d - 1], &neg); }
    }
    pt run; pt_set_inf(&CV, &run);
    for (size_t b = half; b-- > 0;) { bucket_add(&buckets[b], &run); pt_add(&CV, &acc, &acc, &run); }
  }
  pt_add(&CV, out, &boolean_sum, &acc);
  free(reprs); free(nb); free(digits); free(carry); free(buckets);
}
/* msm (msm.rs:187-222) */
static void msm(const fe *coeffs, const apt *bases, size_t n, int par, pt *out) {
  int T = (par && n >= 1024) ? g_threads : 1;
  if (n > (size_t)T && T > 1) {
    size_t chunk = n / T, nch = (n + chunk - 1) / chunk;
    pt *parts = (pt *)malloc(nch * sizeof(pt));
#pragma omp parallel for schedule(static, 1)
    for (size_t k = 0; k < nch; k++) { size_t s = k * chunk, e = s + chunk < n ? s + chunk : n; msm_serial(coeffs + s, bases + s, e - s, &parts[k]); }
    pt acc; pt_set_inf(&CV, &acc);
    for (size_t k = 0; k < nch; k++) pt_add(&CV, &acc, &acc, &parts[k]);
    *out = acc; free(parts);
  } else msm_serial(coeffs, bases, n, out);
}
EXPORT void orc_msm(const fe *coeffs, const apt *bases, size_t n, apt *out) { pt r; msm(coeffs, bases, n, 1, &r); pt_to_affine(&CV, out, &r); }

/* msm_small (msm.rs:367-620): binary / <=10-bit / windowed */
static void msm_small_serial(const uint64_t *s, const apt *bases, size_t n, pt *out) {
  uint64_t mx = 0; for (size_t i = 0; i < n; i++) if (s[i] > mx) mx = s[i];
  size_t bits = 0; while (bits < 64 && (mx >> bits)) bits++;
  pt res; pt_set_inf(&CV, &res);
  if (bits == 0) { *out = res; return; }
  if (bits == 1) {                                  /* msm_binary (msm.rs:418-451) */
    for (size_t i = 0; i < n; i++) if (s[i]) pt_add_mixed(&CV, &res, &res, &bases[i]);
    *out = res; return;
  }
  if (bits <= 10) {                                 /* msm_10 (msm.rs:455-502) */
    size_t nb = (size_t)1 << bits; bucket *bk = (bucket *)calloc(nb, sizeof(bucket));
    for (size_t i = 0; i < n; i++) if (s[i]) bucket_add_assign(&bk[s[i]], &bases[i]);
    pt run; pt_set_inf(&CV, &run);
    for (size_t b = nb; b-- > 1;) { bucket_add(&bk[b], &run); pt_add(&CV, &res, &res, &run); }
    free(bk); *out = res; return;
  }
  /* msm_small_rest (msm.rs:505-620) */
  size_t c = 3;
  if (n >= 32) { size_t lg = 0; while (((size_t)1 << (lg + 1)) <= n) lg++; c = lg * 69 / 100 + 2; }
  size_t nwin = (bits + c - 1) / c; pt *ws = (pt *)malloc(nwin * sizeof(pt));
  size_t nb = ((size_t)1 << c) - 1; pt *bk = (pt *)malloc(nb * sizeof(pt));
  for (size_t w = 0; w < nwin; w++) {
    size_t ws0 = w * c; pt r; pt_set_inf(&CV, &r);
    for (size_t b = 0; b < nb; b++) pt_set_inf(&CV, &bk[b]);
    for (size_t i = 0; i < n; i++) {
      if (!s[i]) continue;
      if (s[i] == 1) { if (ws0 == 0) pt_add_mixed(&CV, &r, &r, &bases[i]); }
      else { uint64_t v = (s[i] >> ws0) % ((uint64_t)1 << c); if (v) pt_add_mixed(&CV, &bk[v - 1], &bk[v - 1], &bases[i]); }
    }
    pt run; pt_set_inf(&CV, &run);
    for (size_t b = nb; b-- > 0;) { pt_add(&CV, &run, &run, &bk[b]); pt_add(&CV, &r, &r, &run); }
    ws[w] = r;
  }
  pt total; pt_set_inf(&CV, &total);
  for (size_t w = nwin; w-- > 1;) { pt_add(&CV, &total, &total, &ws[w]); for (size_t k = 0; k < c; k++) pt_dbl(&CV, &total, &total); }
  pt_add(&CV, out, &ws[0], &total);
  free(ws); free(bk);
}
EXPORT void orc_msm_small(const uint64_t *s, const apt *bases, size_t n, apt *out) { pt r; msm_small_serial(s, bases, n, &r); pt_to_affine(&CV, out, &r); }

/* scalar * point (FixedBaseMul::mul msm.rs:691 / vartime_scalar_mul msm.rs:779 produce the same
 * group element as plain double-and-add; the table layout is an implementation detail) */
static void pt_mul_fe(pt *out, const apt *base, const fe *k) {
  uint64_t raw[4]; f_to_raw(&FQ, raw, k); pt b; pt_from_affine(&CV, &b, base); pt_mul_raw(&CV, out, &b, raw);
}
/* FixedBaseMul (msm.rs:637-774): 8-bit windows, 32 tables of 255 affine multiples d * 2^(8w) * B, batch-normalised; mul = at
 * most 32 mixed additions.  The reference keeps one for h in the commitment key (ck.h_table, hyrax_pc.rs:84, 222-227) and uses it
 * in commit / commit_zeros / rerandomize_commitment / commit_incremental / fold_commitments_partial.  Cached per base here. */
typedef struct { apt base; apt *tab; } fbm;       /* tab[w * 255 + d - 1] */
#define FBM_CACHE 8
static fbm g_fbm[FBM_CACHE]; static int g_fbm_n = 0;
static const fbm *fbm_get(const apt *base) {
  const fbm *hit = NULL;
#pragma omp critical(fbm_cache)
  {
    for (int i = 0; i < g_fbm_n; i++) if (memcmp(&g_fbm[i].base, base, sizeof(apt)) == 0) hit = &g_fbm[i];
    if (!hit) {
      fbm *f = &g_fbm[g_fbm_n < FBM_CACHE ? g_fbm_n++ : 0];
      if (f->tab) free(f->tab);
      f->base = *base; f->tab = (apt *)malloc(32 * 255 * sizeof(apt));
      pt *tmp = (pt *)malloc(32 * 255 * sizeof(pt));
      pt wbase; pt_from_affine(&CV, &wbase, base);
      for (int w = 0; w < 32; w++) {                       /* multiples 1..255 of 2^(8w) B (msm.rs:651-667) */
        pt m = wbase;
        for (int d = 0; d < 255; d++) { tmp[w * 255 + d] = m; pt_add(&CV, &m, &m, &wbase); }
        wbase = m;                                          /* 256 * previous window base */
      }
      /* batch_normalize (msm.rs:669-676): Montgomery's trick, one inversion */
      fe *pre = (fe *)malloc((32 * 255 + 1) * sizeof(fe)); fe acc; f_one(&FPB, &acc);
      for (int i = 0; i < 32 * 255; i++) { pre[i] = acc; if (!pt_is_inf(&tmp[i])) f_mul(&FPB, &acc, &acc, &tmp[i].z); }
      fe inv; f_inv(&FPB, &inv, &acc);
      for (int i = 32 * 255; i-- > 0;) {
        if (pt_is_inf(&tmp[i])) { f_zero(&f->tab[i].x); f_zero(&f->tab[i].y); continue; }
        fe zi, zi2, zi3; f_mul(&FPB, &zi, &inv, &pre[i]); f_mul(&FPB, &inv, &inv, &tmp[i].z);
        f_sqr(&FPB, &zi2, &zi); f_mul(&FPB, &zi3, &zi2, &zi);
        f_mul(&FPB, &f->tab[i].x, &tmp[i].x, &zi2); f_mul(&FPB, &f->tab[i].y, &tmp[i].y, &zi3);
      }
      free(pre); free(tmp);
      hit = f;
    }
  }
  return hit;
}
/* FixedBaseMul::mul (msm.rs:691-726) */
static void pt_mul_fixed(pt *out, const apt *base, const fe *k) {
  const fbm *f = fbm_get(base);
  uint8_t bytes[32]; fe_to_le_bytes(&FQ, k, bytes);
  pt acc; pt_set_inf(&CV, &acc);
  for (int w = 0; w < 32; w++) if (bytes[w]) pt_add_mixed(&CV, &acc, &acc, &f->tab[w * 255 + bytes[w] - 1]);
  *out = acc;
}
EXPORT void orc_fixed_base_mul(const apt *base, const fe *k, apt *out) { pt r; pt_mul_fixed(&r, base, k); pt_to_affine(&CV, out, &r); }
EXPORT void orc_scalar_mul(const apt *base, const fe *k, apt *out) { pt r; pt_mul_fe(&r, base, k); pt_to_affine(&CV, out, &r); }
EXPORT int orc_on_curve(const apt *p) { return apt_on_curve(&CV, p); }
EXPORT void orc_point_add(const apt *a, const apt *b, apt *out) { pt p, r; pt_from_affine(&CV, &p, a); pt_add_mixed(&CV, &r, &p, b); pt_to_affine(&CV, out, &r); }

/* ------------------------------------------------------------------------------------
 * Hyrax  (provider/pcs/hyrax_pc.rs, ipa.rs)
 * ---------------------------------------------------------------------------------- */
/* HyraxPCS::commit (hyrax_pc.rs:207-303); commit_zeros (:305-319) is the all-zero case of it.
 * ck: num_cols bases, h: blinding base. out: ceil(n/num_cols) affine rows. */
EXPORT void orc_hyrax_commit(const apt *ck, size_t num_cols, const apt *h, const fe *v, size_t n,
                             const fe *blinds, int is_small, apt *out) {
  size_t rows = (n + num_cols - 1) / num_cols;
#pragma omp parallel for schedule(dynamic, 1) if (g_threads > 1)
  for (size_t i = 0; i < rows; i++) {
    size_t lo = i * num_cols, hi = lo + num_cols < n ? lo + num_cols : n, len = hi - lo;
    const fe *s = v + lo; pt hb, acc; pt_mul_fixed(&hb, h, &blinds[i]);      /* h_table.mul (hyrax_pc.rs:222-227, 296) */
    size_t eff = len; while (eff > 0 && f_is_zero(&s[eff - 1])) eff--;
    if (eff == 0) { pt_to_affine(&CV, &out[i], &hb); continue; }
    int all_small = is_small;
    uint64_t *sm = (uint64_t *)malloc(eff * 8);
    if (eff <= 16) all_small = 0;
    else {
      if (!is_small) all_small = 1;
      for (size_t j = 0; j < eff; j++) {
        uint64_t raw[4]; f_to_raw(&FQ, raw, &s[j]);
        if (!is_small && (raw[1] | raw[2] | raw[3])) { all_small = 0; break; }
        sm[j] = raw[0];
      }
    }
    if (all_small) msm_small_serial(sm, ck, eff, &acc); else msm(s, ck, eff, 0, &acc);
    free(sm);
    pt_add(&CV, &acc, &acc, &hb); pt_to_affine(&CV, &out[i], &acc);
  }
}
/* bind_with_delayed (hyrax_pc.rs:38-54): LZ = L^T * W */
static void hyrax_bind(const fe *poly, const fe *L, size_t rows, size_t r_len, fe *out) {
  acc9 *acc = (acc9 *)calloc(r_len, sizeof(acc9));
  for (size_t j = 0; j < rows; j++) for (size_t i = 0; i < r_len; i++) f_mul_acc(&acc[i], &L[j], &poly[j * r_len + i]);
  for (size_t i = 0; i < r_len; i++) f_reduce9(&FQ, &out[i], &acc[i]);
  free(acc);
}
EXPORT void orc_hyrax_bind(const fe *poly, const fe *L, size_t rows, size_t r_len, fe *out) { hyrax_bind(poly, L, rows, r_len, out); }

typedef struct {           /* flat proof view shared with the product ABI (include/spartan2_b200.h) */
  uint64_t num_rounds_x, num_rounds_y, num_comm_rows, num_cols;
  apt *comm_W;             /* num_comm_rows */
  fe *outer_polys;         /* num_rounds_x * 3 : compressed [c0, c2, c3] */
  fe *claims_outer;        /* 3 */
  fe *inner_polys;         /* num_rounds_y * 2 : compressed [c0, c2] */
  fe *eval_W, *blind_eval_W;
  apt *delta, *beta;
  fe *z_vec;               /* num_cols */
  fe *z_delta, *z_beta;
} proof_view;

typedef struct {           /* commitment keys: ck (num_cols bases + h) and ck_s (1 base + h) */
  const apt *ck; size_t num_cols; const apt *h; const apt *ck_s; const apt *h_s;
} keys_view;

typedef struct {           /* prover randomness, supplied by the caller so runs are reproducible */
  const fe *blinds_W;      /* one per commitment row */
  const fe *blind_eval_W, *d_vec, *r_delta, *r_beta;
} rand_view;

/* HyraxPCS::prove (hyrax_pc.rs:387-478) + InnerProductArgumentLinear::prove (ipa.rs:125-170) */
static void hyrax_prove(const keys_view *K, transcript *ts, const apt *comm, size_t nrows_comm, const fe *poly, size_t n,
                        const fe *blind, const fe *point, size_t npoint, const apt *comm_eval, const rand_view *R,
                        proof_view *P, fe *LZ_out) {
  ts_absorb_commitment(ts, "poly_com", comm, nrows_comm);
  size_t num_cols = K->num_cols, rows = (n + num_cols - 1) / num_cols;
  size_t nvr = 0; while (((size_t)1 << nvr) < rows) nvr++;
  fe *Rv, *LZ = (fe *)malloc(num_cols * sizeof(fe)); fe r_LZ; apt comm_LZ; size_t rlen;
  if (nvr == 0) {
    rlen = (size_t)1 << npoint; Rv = (fe *)malloc(rlen * sizeof(fe)); eq_evals(point, npoint, Rv);
    memcpy(LZ, poly, rlen * sizeof(fe)); r_LZ = blind[0]; comm_LZ = comm[0];
  } else {
    fe *L = (fe *)malloc(rows * sizeof(fe)); rlen = (size_t)1 << (npoint - nvr); Rv = (fe *)malloc(rlen * sizeof(fe));
    eq_evals(point, nvr, L); eq_evals(point + nvr, npoint - nvr, Rv);
    hyrax_bind(poly, L, rows, rlen, LZ);
    f_zero(&r_LZ);
    for (size_t i = 0; i < rows; i++) { fe t; f_mul(&FQ, &t, &L[i], &blind[i]); f_add(&FQ, &r_LZ, &r_LZ, &t); }
    pt c, hb; msm(LZ, K->ck, rlen, 1, &c); pt_mul_fe(&hb, K->h, &r_LZ); pt_add(&CV, &c, &c, &hb); pt_to_affine(&CV, &comm_LZ, &c);
    free(L);
  }
  if (LZ_out) memcpy(LZ_out, LZ, rlen * sizeof(fe));
  /* IPA */
  ts_dom_sep(ts, "inner product argument (linear)");
  { uint8_t b[128]; apt_to_bytes(&comm_LZ, b); apt_to_bytes(comm_eval, b + 64); ts_absorb_bytes(ts, "U", b, 128); }
  pt d, hb, be; msm(R->d_vec, K->ck, rlen, 1, &d); pt_mul_fe(&hb, K->h, R->r_delta); pt_add(&CV, &d, &d, &hb);
  pt_to_affine(&CV, P->delta, &d);
  fe ip; f_zero(&ip);
  for (size_t i = 0; i < rlen; i++) { fe t; f_mul(&FQ, &t, &Rv[i], &R->d_vec[i]); f_add(&FQ, &ip, &ip, &t); }   /* inner_product (ipa.rs:22-27) */
  pt_mul_fe(&be, K->ck_s, &ip); pt_mul_fe(&hb, K->h_s, R->r_beta); pt_add(&CV, &be, &be, &hb);
  pt_to_affine(&CV, P->beta, &be);
  ts_absorb_point(ts, "delta", P->delta); ts_absorb_point(ts, "beta", P->beta);
  fe r; ts_squeeze(ts, &FQ, "r", &r);
  for (size_t i = 0; i < rlen; i++) { fe t; f_mul(&FQ, &t, &r, &LZ[i]); f_add(&FQ, &P->z_vec[i], &t, &R->d_vec[i]); }
  fe t; f_mul(&FQ, &t, &r, &r_LZ); f_add(&FQ, P->z_delta, &t, R->r_delta);
  f_mul(&FQ, &t, &r, R->blind_eval_W); f_add(&FQ, P->z_beta, &t, R->r_beta);
  free(Rv); free(LZ);
}

/* HyraxPCS::verify (hyrax_pc.rs:480-531) + InnerProductArgumentLinear::verify (ipa.rs:173-221) */
static int hyrax_verify(const keys_view *K, transcript *ts, const apt *comm, size_t nrows_comm, const fe *point, size_t npoint,
                        const apt *comm_eval, const proof_view *P) {
  ts_absorb_commitment(ts, "poly_com", comm, nrows_comm);
  size_t n = (size_t)1 << npoint, num_cols = K->num_cols, rows = (n + num_cols - 1) / num_cols;
  size_t nvr = 0; while (((size_t)1 << nvr) < rows) nvr++;
  fe *Rv; size_t rlen; pt comm_LZ;
  if (nvr == 0) { rlen = n; Rv = (fe *)malloc(rlen * sizeof(fe)); eq_evals(point, npoint, Rv); pt_from_affine(&CV, &comm_LZ, &comm[0]); }
  else {
    fe *L = (fe *)malloc(rows * sizeof(fe)); rlen = (size_t)1 << (npoint - nvr); Rv = (fe *)malloc(rlen * sizeof(fe));
    eq_evals(point, nvr, L); eq_evals(point + nvr, npoint - nvr, Rv);
    msm(L, comm, rows, 1, &comm_LZ); free(L);
  }
  ts_dom_sep(ts, "inner product argument (linear)");
  { apt a; pt_to_affine(&CV, &a, &comm_LZ); uint8_t b[128]; apt_to_bytes(&a, b); apt_to_bytes(comm_eval, b + 64); ts_absorb_bytes(ts, "U", b, 128); }
  ts_absorb_point(ts, "delta", P->delta); ts_absorb_point(ts, "beta", P->beta);
  fe r; ts_squeeze(ts, &FQ, "r", &r);
  int ok = 1;
  { pt lhs, rhs, hb, dl; uint64_t raw[4]; f_to_raw(&FQ, raw, &r);
    pt_mul_raw(&CV, &lhs, &comm_LZ, raw); pt_from_affine(&CV, &dl, P->delta); pt_add(&CV, &lhs, &lhs, &dl);
    msm(P->z_vec, K->ck, rlen, 1, &rhs); pt_mul_fe(&hb, K->h, P->z_delta); pt_add(&CV, &rhs, &rhs, &hb);
    if (!pt_eq(&CV, &lhs, &rhs)) ok = 0; }
  { pt lhs, rhs, hb, bt, ce; uint64_t raw[4]; f_to_raw(&FQ, raw, &r);
    pt_from_affine(&CV, &ce, comm_eval); pt_mul_raw(&CV, &lhs, &ce, raw); pt_from_affine(&CV, &bt, P->beta); pt_add(&CV, &lhs, &lhs, &bt);
    fe ip; f_zero(&ip);
    for (size_t i = 0; i < rlen; i++) { fe t; f_mul(&FQ, &t, &P->z_vec[i], &Rv[i]); f_add(&FQ, &ip, &ip, &t); }
    pt_mul_fe(&rhs, K->ck_s, &ip); pt_mul_fe(&hb, K->h_s, P->z_beta); pt_add(&CV, &rhs, &rhs, &hb);
    if (!pt_eq(&CV, &lhs, &rhs)) ok = 0; }
  free(Rv);
  return ok ? 0 : -1;
}

/* SparsePolynomial::evaluate (polys/multilinear.rs:190-207) */
static void sparse_poly_eval(size_t num_vars, const fe *Z, size_t zlen, const fe *r, fe *out) {
  size_t p2 = 1, nvz = 0; while (p2 < zlen) { p2 <<= 1; nvz++; }
  size_t skip = num_vars - 1 - nvz;       /* mirrors r[self.num_vars - 1 - num_vars_z ..] */
  size_t k = num_vars - skip; fe *chis = (fe *)malloc(((size_t)1 << k) * sizeof(fe));
  eq_evals(r + skip, k, chis);
  fe acc, t; f_zero(&acc);
  for (size_t i = 0; i < zlen; i++) { f_mul(&FQ, &t, &Z[i], &chis[i]); f_add(&FQ, &acc, &acc, &t); }
  fe common, one; f_one(&FQ, &common); f_one(&FQ, &one);
  for (size_t i = 0; i < skip; i++) { f_sub(&FQ, &t, &one, &r[i]); f_mul(&FQ, &common, &common, &t); }
  f_mul(&FQ, out, &common, &acc); free(chis);
}

/* ------------------------------------------------------------------------------------
 * SpartanSNARK::prove  (spartan.rs:219-466) — non-ZK, no challenges (num_challenges == 0)
 *   W          : num_vars witness (shared | precommitted | rest), Montgomery limbs
 *   comm_pre   : commitment rows of the shared+precommitted sections (from prep_prove)
 *   cached_*   : optional cached partial products (prep_prove, spartan.rs:184-187); pass NULL to
 *                recompute the full SpMV (same values).
 * debug_out (optional): [tau (l) | r_x (l) | r_y (m+1)] challenges for cross-checking.
 * ---------------------------------------------------------------------------------- */
EXPORT int orc_spartan_prove_cached(const shape *S, const keys_view *K, const uint8_t vk_digest[32], const fe *public_values,
                                    const fe *W, const apt *comm_pre, size_t comm_pre_rows, const rand_view *R,
                                    proof_view *P, fe *debug_out, double *phase_ms, const fe *caz, const fe *cbz, const fe *ccz);
EXPORT int orc_spartan_prove(const shape *S, const keys_view *K, const uint8_t vk_digest[32], const fe *public_values,
                             const fe *W, const apt *comm_pre, size_t comm_pre_rows, const rand_view *R,
                             proof_view *P, fe *debug_out, double *phase_ms) {
  return orc_spartan_prove_cached(S, K, vk_digest, public_values, W, comm_pre, comm_pre_rows, R, P, debug_out, phase_ms, NULL, NULL, NULL);
}
/* caz / cbz / ccz: the cached partial products of prep_prove (spartan.rs:184-187), or NULL to recompute the full SpMV */
EXPORT int orc_spartan_prove_cached(const shape *S, const keys_view *K, const uint8_t vk_digest[32], const fe *public_values,
                                    const fe *W, const apt *comm_pre, size_t comm_pre_rows, const rand_view *R,
                                    proof_view *P, fe *debug_out, double *phase_ms, const fe *caz, const fe *cbz, const fe *ccz) {
  double t_last = 0; (void)t_last;
#ifdef _OPENMP
#define TICK(idx) do { double now = omp_get_wtime(); if (phase_ms) phase_ms[idx] += (now - t_last) * 1e3; t_last = now; } while (0)
  t_last = omp_get_wtime();
#else
#define TICK(idx) do {} while (0)
#endif
  size_t num_vars = S->num_vars, N = S->num_cons, num_cols = K->num_cols;
  size_t l = 0; while (((size_t)1 << l) < N) l++;
  size_t m = 0; while (((size_t)1 << m) < num_vars) m++;
  size_t nry = m + 1, num_extra = 1 + S->num_public + S->num_challenges;
  size_t rows_total = num_vars / num_cols, rest_rows = rows_total - comm_pre_rows;
  transcript ts; ts_new(&ts, "SpartanSNARK");
  ts_absorb_bytes(&ts, "vk", vk_digest, 32);
  ts_absorb_scalars(&ts, &FQ, "public_values", public_values, S->num_public);
  /* r1cs_instance_and_witness (bellpepper/r1cs.rs:411-537) */
  size_t pre_rows_shared = S->num_shared / num_cols;
  if (S->num_shared) ts_absorb_commitment(&ts, "comm_W_shared", comm_pre, pre_rows_shared);
  if (S->num_precommitted) ts_absorb_commitment(&ts, "comm_W_precommitted", comm_pre + pre_rows_shared, comm_pre_rows - pre_rows_shared);
  memcpy(P->comm_W, comm_pre, comm_pre_rows * sizeof(apt));
  orc_hyrax_commit(K->ck, num_cols, K->h, W + comm_pre_rows * num_cols, rest_rows * num_cols, R->blinds_W + comm_pre_rows, 0, P->comm_W + comm_pre_rows);
  ts_absorb_commitment(&ts, "comm_W_rest", P->comm_W + comm_pre_rows, rest_rows);
  P->num_rounds_x = l; P->num_rounds_y = nry; P->num_comm_rows = rows_total; P->num_cols = num_cols;
  TICK(0);
  /* z = W | 1 | X  (spartan.rs:248-253) */
  size_t zlen = num_vars + num_extra;
  fe *z = (fe *)malloc(2 * num_vars * sizeof(fe)); memset(z, 0, 2 * num_vars * sizeof(fe));
  memcpy(z, W, num_vars * sizeof(fe)); f_one(&FQ, &z[num_vars]); memcpy(&z[num_vars + 1], public_values, S->num_public * sizeof(fe));
  fe *tau = (fe *)malloc(l * sizeof(fe));
  for (size_t i = 0; i < l; i++) ts_squeeze(&ts, &FQ, "t", &tau[i]);
  fe *az = (fe *)malloc(N * sizeof(fe)), *bz = (fe *)malloc(N * sizeof(fe)), *cz = (fe *)malloc(N * sizeof(fe));
  if (caz) orc_shape_multiply_vec_incremental(S, z, caz, cbz, ccz, az, bz, cz);     /* spartan.rs:271 */
  else orc_shape_multiply_vec(S, z, az, bz, cz);
  TICK(1);
  fe zero; f_zero(&zero);
  fe *polys = (fe *)malloc(l * 4 * sizeof(fe)), *rx = (fe *)malloc(l * sizeof(fe)), claims[3];
  orc_sumcheck_cubic_prove(&zero, tau, l, az, bz, cz, &ts, polys, rx, claims, NULL);
  for (size_t i = 0; i < l; i++) { P->outer_polys[3 * i] = polys[4 * i]; P->outer_polys[3 * i + 1] = polys[4 * i + 2]; P->outer_polys[3 * i + 2] = polys[4 * i + 3]; }
  memcpy(P->claims_outer, claims, sizeof(claims));
  ts_absorb_scalars(&ts, &FQ, "claims_outer", claims, 3);
  TICK(2);
  fe r; ts_squeeze(&ts, &FQ, "r", &r);
  fe joint, t, r2; f_mul(&FQ, &r2, &r, &r);
  f_mul(&FQ, &t, &r, &claims[1]); f_add(&FQ, &joint, &claims[0], &t); f_mul(&FQ, &t, &r2, &claims[2]); f_add(&FQ, &joint, &joint, &t);
  fe *evals_rx = az;                                    /* reuse allocation (N entries) */
  eq_evals_serial(rx, l, evals_rx);
  fe *abc = (fe *)malloc(2 * num_vars * sizeof(fe));
  orc_shape_abc(S, evals_rx, &r, abc, zlen);
  TICK(3);
  /* inner sum-check, manual round 0 (spartan.rs:330-384) */
  acc9 a0; memset(&a0, 0, sizeof(a0));
  for (size_t j = 0; j < num_vars; j++) f_mul_acc(&a0, &abc[j], &z[j]);
  fe eval0; f_reduce9(&FQ, &eval0, &a0);
  fe corr_low, corr_cross; f_zero(&corr_low); f_zero(&corr_cross);
  for (size_t j = 0; j < num_extra; j++) {
    fe d1, d2;
    f_mul(&FQ, &t, &abc[j], &z[j]); f_add(&FQ, &corr_low, &corr_low, &t);
    f_sub(&FQ, &d1, &abc[num_vars + j], &abc[j]); f_sub(&FQ, &d2, &z[num_vars + j], &z[j]);
    f_mul(&FQ, &t, &d1, &d2); f_add(&FQ, &corr_cross, &corr_cross, &t);
  }
  fe tinf; f_sub(&FQ, &tinf, &eval0, &corr_low); f_add(&FQ, &tinf, &tinf, &corr_cross);
  fe co0[3]; quad_round_poly(&eval0, &tinf, &joint, co0);
  ts_absorb_unipoly(&ts, co0, 3);
  fe r0; ts_squeeze(&ts, &FQ, "c", &r0);
  fe claim1; unipoly_eval(co0, 3, &r0, &claim1);
  fe one, omr; f_one(&FQ, &one); f_sub(&FQ, &omr, &one, &r0);
  for (size_t j = 0; j < num_extra; j++) {
    fe d;
    f_sub(&FQ, &d, &abc[num_vars + j], &abc[j]); f_mul(&FQ, &d, &d, &r0); f_add(&FQ, &abc[j], &abc[j], &d);
    f_sub(&FQ, &d, &z[num_vars + j], &z[j]); f_mul(&FQ, &d, &d, &r0); f_add(&FQ, &z[j], &z[j], &d);
  }
  for (size_t j = num_extra; j < num_vars; j++) { f_mul(&FQ, &abc[j], &abc[j], &omr); f_mul(&FQ, &z[j], &z[j], &omr); }
  fe *ipolys = (fe *)malloc(nry * 3 * sizeof(fe)), *ry = (fe *)malloc(nry * sizeof(fe)), iclaims[2];
  orc_sumcheck_quad_prove(&claim1, m, abc, z, &ts, ipolys + 3, ry + 1, iclaims);
  memcpy(ipolys, co0, sizeof(co0)); ry[0] = r0;
  for (size_t i = 0; i < nry; i++) { P->inner_polys[2 * i] = ipolys[3 * i]; P->inner_polys[2 * i + 1] = ipolys[3 * i + 2]; }
  TICK(4);
  /* eval_W (spartan.rs:411-421) */
  fe *X = (fe *)malloc(num_extra * sizeof(fe)); X[0] = one; memcpy(&X[1], public_values, S->num_public * sizeof(fe));
  fe eval_X; sparse_poly_eval(nry - 1, X, num_extra, ry + 1, &eval_X);
  fe inv; f_sub(&FQ, &t, &one, &ry[0]);
  if (!f_inv(&FQ, &inv, &t)) return -5;                  /* SpartanError::DivisionByZero */
  fe evW; f_mul(&FQ, &t, &ry[0], &eval_X); f_sub(&FQ, &evW, &iclaims[1], &t); f_mul(&FQ, &evW, &evW, &inv);
  *P->eval_W = evW; *P->blind_eval_W = *R->blind_eval_W;
  /* comm_eval_W = commit(ck_s, [eval_W], blind) */
  apt comm_eval; { pt a, b; pt_mul_fe(&a, K->ck_s, &evW); pt_mul_fe(&b, K->h_s, R->blind_eval_W); pt_add(&CV, &a, &a, &b); pt_to_affine(&CV, &comm_eval, &a); }
  hyrax_prove(K, &ts, P->comm_W, rows_total, W, num_vars, R->blinds_W, ry + 1, m, &comm_eval, R, P, NULL);
  TICK(5);
  if (debug_out) { memcpy(debug_out, tau, l * sizeof(fe)); memcpy(debug_out + l, rx, l * sizeof(fe)); memcpy(debug_out + 2 * l, ry, nry * sizeof(fe)); }
  free(z); free(tau); free(az); free(bz); free(cz); free(polys); free(rx); free(abc); free(ipolys); free(ry); free(X);
  ts_free(&ts);
  return 0;
}

/* SpartanSNARK::verify (spartan.rs:469-578).  Returns 0 on success, negative codes otherwise. */
EXPORT int orc_spartan_verify(const shape *S, const keys_view *K, const uint8_t vk_digest[32], const fe *public_values,
                              const proof_view *P) {
  size_t num_vars = S->num_vars, N = S->num_cons, num_cols = K->num_cols;
  size_t l = 0; while (((size_t)1 << l) < N) l++;
  size_t m = 0; while (((size_t)1 << m) < num_vars) m++;
  size_t nry = m + 1, num_extra = 1 + S->num_public + S->num_challenges;
  if (P->num_rounds_x != l || P->num_rounds_y != nry || P->num_comm_rows != num_vars / num_cols) return -1;
  transcript ts; ts_new(&ts, "SpartanSNARK");
  ts_absorb_bytes(&ts, "vk", vk_digest, 32);
  ts_absorb_scalars(&ts, &FQ, "public_values", public_values, S->num_public);
  size_t sh_rows = S->num_shared / num_cols, pre_rows = S->num_precommitted / num_cols, rest_rows = S->num_rest / num_cols;
  if (sh_rows) ts_absorb_commitment(&ts, "comm_W_shared", P->comm_W, sh_rows);
  if (pre_rows) ts_absorb_commitment(&ts, "comm_W_precommitted", P->comm_W + sh_rows, pre_rows);
  ts_absorb_commitment(&ts, "comm_W_rest", P->comm_W + sh_rows + pre_rows, rest_rows);
  fe *tau = (fe *)malloc(l * sizeof(fe));
  for (size_t i = 0; i < l; i++) ts_squeeze(&ts, &FQ, "t", &tau[i]);
  /* expand compressed polys into [c0, (linear unused), c2, c3] */
  fe *op = (fe *)calloc(l * 4, sizeof(fe));
  for (size_t i = 0; i < l; i++) { op[4 * i] = P->outer_polys[3 * i]; op[4 * i + 2] = P->outer_polys[3 * i + 1]; op[4 * i + 3] = P->outer_polys[3 * i + 2]; }
  fe zero, e_outer, *rx = (fe *)malloc(l * sizeof(fe)); f_zero(&zero);
  orc_sumcheck_verify(op, l, 3, &zero, &ts, &e_outer, rx);
  fe tb, one, t, u; f_one(&FQ, &tb); f_one(&FQ, &one);
  for (size_t i = 0; i < l; i++) {   /* EqPolynomial::evaluate (eq.rs:41-46) */
    fe a, b; f_mul(&FQ, &a, &rx[i], &tau[i]); f_sub(&FQ, &t, &one, &rx[i]); f_sub(&FQ, &u, &one, &tau[i]); f_mul(&FQ, &b, &t, &u);
    f_add(&FQ, &a, &a, &b); f_mul(&FQ, &tb, &tb, &a);
  }
  const fe *cA = &P->claims_outer[0], *cB = &P->claims_outer[1], *cC = &P->claims_outer[2];
  fe exp; f_mul(&FQ, &exp, cA, cB); f_sub(&FQ, &exp, &exp, cC); f_mul(&FQ, &exp, &exp, &tb);
  int rc = 0;
  if (!f_eq(&exp, &e_outer)) rc = -2;
  ts_absorb_scalars(&ts, &FQ, "claims_outer", P->claims_outer, 3);
  fe r; ts_squeeze(&ts, &FQ, "r", &r);
  fe r2, joint; f_mul(&FQ, &r2, &r, &r);
  f_mul(&FQ, &t, &r, cB); f_add(&FQ, &joint, cA, &t); f_mul(&FQ, &t, &r2, cC); f_add(&FQ, &joint, &joint, &t);
  fe *ip = (fe *)calloc(nry * 3, sizeof(fe));
  for (size_t i = 0; i < nry; i++) { ip[3 * i] = P->inner_polys[2 * i]; ip[3 * i + 2] = P->inner_polys[2 * i + 1]; }
  fe e_inner, *ry = (fe *)malloc(nry * sizeof(fe));
  orc_sumcheck_verify(ip, nry, 2, &joint, &ts, &e_inner, ry);
  fe *X = (fe *)malloc(num_extra * sizeof(fe)); X[0] = one; memcpy(&X[1], public_values, S->num_public * sizeof(fe));
  fe eval_X; sparse_poly_eval(m, X, num_extra, ry + 1, &eval_X);
  fe eval_Z; f_sub(&FQ, &t, &one, &ry[0]); f_mul(&FQ, &eval_Z, &t, P->eval_W); f_mul(&FQ, &t, &ry[0], &eval_X); f_add(&FQ, &eval_Z, &eval_Z, &t);
  fe *Tx = (fe *)malloc(N * sizeof(fe)), *Ty = (fe *)malloc(2 * num_vars * sizeof(fe)), ev[3];
  eq_evals(rx, l, Tx); eq_evals(ry, nry, Ty);
  orc_shape_eval_tables(S, Tx, Ty, ev);
  fe comb; f_mul(&FQ, &t, &r, &ev[1]); f_add(&FQ, &comb, &ev[0], &t); f_mul(&FQ, &t, &r2, &ev[2]); f_add(&FQ, &comb, &comb, &t);
  f_mul(&FQ, &comb, &comb, &eval_Z);
  if (!rc && !f_eq(&comb, &e_inner)) rc = -3;
  apt comm_eval; { pt a, b; pt_mul_fe(&a, K->ck_s, P->eval_W); pt_mul_fe(&b, K->h_s, P->blind_eval_W); pt_add(&CV, &a, &a, &b); pt_to_affine(&CV, &comm_eval, &a); }
  if (!rc && hyrax_verify(K, &ts, P->comm_W, P->num_comm_rows, ry + 1, m, &comm_eval, P) != 0) rc = -4;
  free(tau); free(op); free(rx); free(ip); free(ry); free(X); free(Tx); free(Ty); ts_free(&ts);
  return rc;
}

/* ------------------------------------------------------------------------------------
 * NeutronNova building blocks (neutronnova_zk.rs, polys/power.rs, r1cs/mod.rs, sumcheck.rs)
 * Function-level restatements: the ZK driver around them (verifier circuit, process_round) is
 * out of scope; challenges are inputs.
 * ---------------------------------------------------------------------------------- */
/* PowPolynomial::split_evals (polys/power.rs:65-86): [1,t,..,t^(left-1)] || [1,t^left,t^(2 left),..] */
EXPORT void orc_pow_split_evals(const fe *t, size_t left, size_t right, fe *out) {
  f_one(&FQ, &out[0]);
  for (size_t i = 1; i < left; i++) f_mul(&FQ, &out[i], &out[i - 1], t);
  fe step; f_mul(&FQ, &step, &out[left - 1], t);
  f_one(&FQ, &out[left]);
  for (size_t i = 1; i < right; i++) f_mul(&FQ, &out[left + i], &out[left + i - 1], &step);
}
/* suffix_weight_full (neutronnova_zk.rs:78-87) */
static void suffix_weight_full(size_t t, size_t ell_b, size_t pair_idx, const fe *rhos, fe *w) {
  fe one; f_one(&FQ, &one); *w = one; size_t k = pair_idx;
  for (size_t s = t + 1; s < ell_b; s++) {
    fe f; if (k & 1) f = rhos[s]; else f_sub(&FQ, &f, &one, &rhos[s]);
    f_mul(&FQ, w, w, &f); k >>= 1;
  }
}
/* NeutronNovaNIFS::prove_helper (neutronnova_zk.rs:98-178), one pair of layers */
static void nifs_prove_helper(size_t round, size_t left, size_t right, const fe *e, const fe *Az1, const fe *Bz1, const fe *Cz1,
                              const fe *Az2, const fe *Bz2, fe *e0, fe *quad) {
  const fe *f = e + left, *el = e;
  acc9 a0, aq; memset(&a0, 0, sizeof(a0)); memset(&aq, 0, sizeof(aq));
  for (size_t i = 0; i < right; i++) {
    acc9 i0, iq; memset(&i0, 0, sizeof(i0)); memset(&iq, 0, sizeof(iq));
    for (size_t j = 0; j < left; j++) {
      size_t k = i * left + j; fe v, da, db;
      if (round != 0) { f_mul(&FQ, &v, &Az1[k], &Bz1[k]); f_sub(&FQ, &v, &v, &Cz1[k]); f_mul_acc(&i0, &el[j], &v); }
      f_sub(&FQ, &da, &Az2[k], &Az1[k]); f_sub(&FQ, &db, &Bz2[k], &Bz1[k]); f_mul(&FQ, &v, &da, &db);
      f_mul_acc(&iq, &el[j], &v);
    }
    fe r0, rq; f_reduce9(&FQ, &r0, &i0); f_reduce9(&FQ, &rq, &iq);
    f_mul_acc(&a0, &f[i], &r0); f_mul_acc(&aq, &f[i], &rq);
  }
  f_reduce9(&FQ, e0, &a0); f_reduce9(&FQ, quad, &aq);
}
/* one NIFS round evaluation over m live layers (standard field path, neutronnova_zk.rs:809-833 / fallback branches):
 * layers are stored layer-major (layer q at offset q*N); returns (e0, quad_coeff) = sum_p w_p * prove_helper(pair p) */
/* pair_offset: index of the first pair among ALL live pairs when A, B, Cm hold only a contiguous block of the layers (the
 * multi-GPU instance sharding of SURVEY.md §8e; 0 for the whole set) */
EXPORT void orc_nifs_round_block(size_t t, size_t ell_b, const fe *rhos, size_t left, size_t right, const fe *E, const fe *A, const fe *B,
                                 const fe *Cm, size_t N, size_t m, size_t pair_offset, fe *out2);
EXPORT void orc_nifs_round(size_t t, size_t ell_b, const fe *rhos, size_t left, size_t right, const fe *E, const fe *A, const fe *B,
                           const fe *Cm, size_t N, size_t m, fe *out2) {
  orc_nifs_round_block(t, ell_b, rhos, left, right, E, A, B, Cm, N, m, 0, out2);
}
EXPORT void orc_nifs_round_block(size_t t, size_t ell_b, const fe *rhos, size_t left, size_t right, const fe *E, const fe *A, const fe *B,
                                 const fe *Cm, size_t N, size_t m, size_t pair_offset, fe *out2) {
  fe e0, q; f_zero(&e0); f_zero(&q);
  fe *pes = (fe *)malloc((m / 2 + 1) * 2 * sizeof(fe));
#pragma omp parallel for schedule(dynamic, 1) if (g_threads > 1)      /* par_iter over the pairs (neutronnova_zk.rs:784-834) */
  for (size_t p = 0; p < m / 2; p++)
    nifs_prove_helper(t, left, right, E, A + 2 * p * N, B + 2 * p * N, Cm + 2 * p * N, A + (2 * p + 1) * N, B + (2 * p + 1) * N, &pes[2 * p], &pes[2 * p + 1]);
  for (size_t p = 0; p < m / 2; p++) {
    fe w, tmp;
    suffix_weight_full(t, ell_b, p + pair_offset, rhos, &w);
    f_mul(&FQ, &tmp, &pes[2 * p], &w); f_add(&FQ, &e0, &e0, &tmp);
    f_mul(&FQ, &tmp, &pes[2 * p + 1], &w); f_add(&FQ, &q, &q, &tmp);
  }
  free(pes);
  out2[0] = e0; out2[1] = q;
}
/* fold_abc_pair over all pairs (neutronnova_zk.rs:738-776): layer p <- lo + r_b (hi - lo); compacts to m/2 layers */
EXPORT void orc_nifs_fold(fe *L, size_t N, size_t m, const fe *r_b) {
  /* pairs are folded in order (layer p <- pair p, compacting in place), each pair's N entries in parallel */
  for (size_t p = 0; p < m / 2; p++)
#pragma omp parallel for schedule(static) if (g_threads > 1 && N >= 4096)
    for (size_t k = 0; k < N; k++) {
      fe d; f_sub(&FQ, &d, &L[(2 * p + 1) * N + k], &L[2 * p * N + k]); f_mul(&FQ, &d, &d, r_b);
      f_add(&FQ, &L[p * N + k], &L[2 * p * N + k], &d);
    }
}
/* weights_from_r (r1cs/mod.rs:153-166): eq(i, r) with LSB-first bits */
EXPORT void orc_weights_from_r(const fe *r_bs, size_t ell, size_t n, fe *w) {
  fe one; f_one(&FQ, &one);
  for (size_t i = 0; i < n; i++) {
    fe wi = one; size_t k = i;
    for (size_t t = 0; t < ell; t++) { fe f; if (k & 1) f = r_bs[t]; else f_sub(&FQ, &f, &one, &r_bs[t]); f_mul(&FQ, &wi, &wi, &f); k >>= 1; }
    w[i] = wi;
  }
}
/* R1CSWitness::fold_multiple, W part (r1cs/mod.rs:570-660): out[j] = sum_i w_i * Ws[i][j] */
EXPORT void orc_fold_vectors(const fe *Ws, size_t n, size_t dim, const fe *w, fe *out) {
  for (size_t j = 0; j < dim; j++) {
    acc9 a; memset(&a, 0, sizeof(a));
    for (size_t i = 0; i < n; i++) f_mul_acc(&a, &w[i], &Ws[i * dim + j]);
    f_reduce9(&FQ, &out[j], &a);
  }
}
/* compute_eval_points_cubic_with_additive_term_with_outer_pow (sumcheck.rs:366-498) incl. the len < left case
 * (:262-342 with the weight table as first polynomial).  table_len = current length of A/B/C. */
EXPORT void orc_pow_cubic_eval(const fe *pl, size_t left, const fe *pr, const fe *A, const fe *B, const fe *Cm, size_t table_len, fe *out3) {
  size_t len = table_len / 2;
  acc9 a0, a2, a3; memset(&a0, 0, sizeof(a0)); memset(&a2, 0, sizeof(a2)); memset(&a3, 0, sizeof(a3));
  if (len < left) {
    for (size_t i = 0; i < len; i++) {
      fe v, wb, ab, bb, cb, t;
      f_mul(&FQ, &v, &A[i], &B[i]); f_sub(&FQ, &v, &v, &Cm[i]); f_mul_acc(&a0, &pl[i], &v);
      f_dbl(&FQ, &wb, &pl[i + len]); f_sub(&FQ, &wb, &wb, &pl[i]);
      f_dbl(&FQ, &ab, &A[i + len]); f_sub(&FQ, &ab, &ab, &A[i]);
      f_dbl(&FQ, &bb, &B[i + len]); f_sub(&FQ, &bb, &bb, &B[i]);
      f_dbl(&FQ, &cb, &Cm[i + len]); f_sub(&FQ, &cb, &cb, &Cm[i]);
      f_mul(&FQ, &v, &ab, &bb); f_sub(&FQ, &v, &v, &cb); f_mul_acc(&a2, &wb, &v);
      f_add(&FQ, &wb, &wb, &pl[i + len]); f_sub(&FQ, &wb, &wb, &pl[i]);
      f_add(&FQ, &ab, &ab, &A[i + len]); f_sub(&FQ, &ab, &ab, &A[i]);
      f_add(&FQ, &bb, &bb, &B[i + len]); f_sub(&FQ, &bb, &bb, &B[i]);
      f_add(&FQ, &cb, &cb, &Cm[i + len]); f_sub(&FQ, &cb, &cb, &Cm[i]);
      f_mul(&FQ, &v, &ab, &bb); f_sub(&FQ, &v, &v, &cb); f_mul_acc(&a3, &wb, &v); (void)t;
    }
  } else {
    size_t right = len / left;
    for (size_t i = 0; i < left; i++) {
      acc9 i0, i2, i3; memset(&i0, 0, sizeof(i0)); memset(&i2, 0, sizeof(i2)); memset(&i3, 0, sizeof(i3));
      for (size_t j = 0; j < right; j++) {
        size_t low = i + j * left, high = low + len;
        const fe *tl = &pr[j], *th = &pr[j + right];
        fe v, tb, ab, bb, cb;
        f_mul(&FQ, &v, &A[low], &B[low]); f_sub(&FQ, &v, &v, &Cm[low]); f_mul_acc(&i0, tl, &v);
        f_dbl(&FQ, &tb, th); f_sub(&FQ, &tb, &tb, tl);
        f_dbl(&FQ, &ab, &A[high]); f_sub(&FQ, &ab, &ab, &A[low]);
        f_dbl(&FQ, &bb, &B[high]); f_sub(&FQ, &bb, &bb, &B[low]);
        f_dbl(&FQ, &cb, &Cm[high]); f_sub(&FQ, &cb, &cb, &Cm[low]);
        f_mul(&FQ, &v, &ab, &bb); f_sub(&FQ, &v, &v, &cb); f_mul_acc(&i2, &tb, &v);
        f_add(&FQ, &tb, &tb, th); f_sub(&FQ, &tb, &tb, tl);
        f_add(&FQ, &ab, &ab, &A[high]); f_sub(&FQ, &ab, &ab, &A[low]);
        f_add(&FQ, &bb, &bb, &B[high]); f_sub(&FQ, &bb, &bb, &B[low]);
        f_add(&FQ, &cb, &cb, &Cm[high]); f_sub(&FQ, &cb, &cb, &Cm[low]);
        f_mul(&FQ, &v, &ab, &bb); f_sub(&FQ, &v, &v, &cb); f_mul_acc(&i3, &tb, &v);
      }
      fe r0, r2, r3; f_reduce9(&FQ, &r0, &i0); f_reduce9(&FQ, &r2, &i2); f_reduce9(&FQ, &r3, &i3);
      f_mul_acc(&a0, &pl[i], &r0); f_mul_acc(&a2, &pl[i], &r2); f_mul_acc(&a3, &pl[i], &r3);
    }
  }
  f_reduce9(&FQ, &out3[0], &a0); f_reduce9(&FQ, &out3[1], &a2); f_reduce9(&FQ, &out3[2], &a3);
}
/* compute_eval_points_quad (sumcheck.rs:128-174) */
EXPORT void orc_quad_eval(const fe *A, const fe *B, size_t table_len, fe *out2) { quad_points(A, B, table_len, &out2[0], &out2[1]); }
/* HyraxPCS::fold_commitments (hyrax_pc.rs:737-793) as group elements: out[row] = sum_i w_i * comms[i][row] */
EXPORT void orc_fold_commitments(const apt *comms, size_t n, size_t rows, const fe *w, apt *out) {
#pragma omp parallel for schedule(dynamic, 1) if (g_threads > 1)      /* msm_shared_weights: par_iter over rows (msm.rs:312-353) */
  for (size_t r = 0; r < rows; r++) {
    pt acc;
    if (n == 2) {                                            /* fast path: p + vartime_scalar_mul(q, w) (hyrax_pc.rs:758-777) */
      pt t; pt_mul_fe(&acc, &comms[r], &w[0]); pt_mul_fe(&t, &comms[rows + r], &w[1]); pt_add(&CV, &acc, &acc, &t);
    } else {                                                 /* msm_shared_weights (msm.rs:228-356): a signed-digit bucket MSM per row */
      apt *rb = (apt *)malloc(n * sizeof(apt));
      for (size_t i = 0; i < n; i++) rb[i] = comms[i * rows + r];
      msm_serial(w, rb, n, &acc); free(rb);
    }
    pt_to_affine(&CV, &out[r], &acc);
  }
}

/* ------------------------------------------------------------------------------------
 * Small-value path (big_num/small_value.rs; its users in neutronnova_zk.rs)
 * ---------------------------------------------------------------------------------- */
#define SMALL_VALUE_MAX ((((uint64_t)1) << 62) - 1)          /* small_value.rs:32 */
/* to_small_vec_or_zero (small_value.rs:42-85): canonical value in [0, 2^62-1] -> v, in [p-(2^62-1), p-1] -> -(p-v),
 * anything else -> 0 with its index recorded.  Returns the number of large positions (ascending). */
EXPORT size_t orc_to_small_vec_or_zero(const fe *poly, size_t n, int64_t *out, uint64_t *large_positions) {
  size_t nl = 0;
  for (size_t idx = 0; idx < n; idx++) {
    fe c; f_to_raw(&FQ, c.l, &poly[idx]);                      /* to_repr(): canonical little-endian limbs */
    if (c.l[1] == 0 && c.l[2] == 0 && c.l[3] == 0 && c.l[0] <= SMALL_VALUE_MAX) { out[idx] = (int64_t)c.l[0]; continue; }
    uint64_t d[4];
    f_sub4(d, FQ.mod, c.l);
    if (d[1] == 0 && d[2] == 0 && d[3] == 0 && d[0] > 0 && d[0] <= SMALL_VALUE_MAX) { out[idx] = -(int64_t)d[0]; continue; }
    out[idx] = 0; large_positions[nl++] = idx;
  }
  return nl;
}
/* SmallAccumulator (small_value.rs:96-196): separate positive / negative 448-bit sums of field_mont * |i128| */
typedef struct { uint64_t pos[7], neg[7]; } small_acc;
static void small_acc_accumulate(small_acc *a, const fe *f, __int128 val) {
  if (val == 0) return;
  const u128 av = val > 0 ? (u128)val : (u128)(-val);
  uint64_t *t = val > 0 ? a->pos : a->neg;
  const uint64_t lo = (uint64_t)av, hi = (uint64_t)(av >> 64);
  u128 carry = 0;
  for (int j = 0; j < 4; j++) { u128 p = (u128)f->l[j] * lo + t[j] + carry; t[j] = (uint64_t)p; carry = p >> 64; }
  for (int j = 4; j < 7 && carry; j++) { u128 s = (u128)t[j] + carry; t[j] = (uint64_t)s; carry = s >> 64; }
  if (hi) {
    carry = 0;
    for (int j = 0; j < 4; j++) { u128 p = (u128)f->l[j] * hi + t[j + 1] + carry; t[j + 1] = (uint64_t)p; carry = p >> 64; }
    for (int j = 5; j < 7 && carry; j++) { u128 s = (u128)t[j] + carry; t[j] = (uint64_t)s; carry = s >> 64; }
  }
}
/* reduce_7_to_field (small_value.rs:204-222): acc mod p, read as Montgomery limbs.  The general case is the
 * reference's binary long division (limbs.rs:411-470), restated as shift-and-subtract from the top bit. */
static void reduce_7_to_field(fe *out, const uint64_t acc[7]) {
  if (acc[4] == 0 && acc[5] == 0 && acc[6] == 0) {
    uint64_t b[4] = {acc[0], acc[1], acc[2], acc[3]};
    for (int i = 0; i < FQ.max_sub; i++) if (f_gte4(b, FQ.mod)) f_sub4(b, b, FQ.mod);
    memcpy(out->l, b, 32);
    return;
  }
  uint64_t r[5] = {0, 0, 0, 0, 0};                             /* remainder < 2p fits 5 limbs */
  for (int bit = 447; bit >= 0; bit--) {
    for (int i = 4; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 63);
    r[0] = (r[0] << 1) | ((acc[bit >> 6] >> (bit & 63)) & 1);
    if (r[4] || f_gte4(r, FQ.mod)) { uint64_t bw = f_sub4(r, r, FQ.mod); r[4] -= bw; }
  }
  memcpy(out->l, r, 32);
}
static void small_acc_reduce(fe *out, const small_acc *a) {
  fe p, n; reduce_7_to_field(&p, a->pos); reduce_7_to_field(&n, a->neg);
  f_sub(&FQ, out, &p, &n);
}
/* test hook: sum_j f[j] * vals[j] through SmallAccumulator (vals as (lo, hi) two's-complement i128 halves) */
EXPORT void orc_small_acc_dot(const fe *f, const uint64_t *vals_lo, const int64_t *vals_hi, size_t n, fe *out) {
  small_acc a; memset(&a, 0, sizeof(a));
  for (size_t j = 0; j < n; j++) small_acc_accumulate(&a, &f[j], (__int128)(((u128)(uint64_t)vals_hi[j] << 64) | vals_lo[j]));
  small_acc_reduce(out, &a);
}
/* NeutronNovaNIFS::prove_helper_small (neutronnova_zk.rs:255-325): round 0, quad coefficient of one pair */
static void nifs_prove_helper_small(size_t left, size_t right, const fe *e, const fe *Az1, const fe *Bz1, const fe *Az2, const fe *Bz2,
                                    const int64_t *a1, const int64_t *b1, const int64_t *a2, const int64_t *b2,
                                    const uint64_t *large_positions, size_t n_large, fe *quad) {
  const fe *f = e + left, *el = e;
  const size_t total = left * right;
  acc9 aq; memset(&aq, 0, sizeof(aq));
  for (size_t i = 0; i < right; i++) {
    small_acc in; memset(&in, 0, sizeof(in));
    for (size_t j = 0; j < left; j++) {
      const size_t k = i * left + j;
      const __int128 da = (__int128)a2[k] - (__int128)a1[k], db = (__int128)b2[k] - (__int128)b1[k];
      small_acc_accumulate(&in, &el[j], da * db);
    }
    fe red; small_acc_reduce(&red, &in);
    f_mul_acc(&aq, &f[i], &red);
  }
  f_reduce9(&FQ, quad, &aq);
  for (size_t q = 0; q < n_large; q++) {                       /* field correction at the zeroed positions */
    const size_t k = large_positions[q];
    if (k >= total) continue;
    const size_t i = k / left, j = k % left;
    fe da, db, t;
    f_sub(&FQ, &da, &Az2[k], &Az1[k]); f_sub(&FQ, &db, &Bz2[k], &Bz1[k]);
    f_mul(&FQ, &t, &f[i], &el[j]); f_mul(&FQ, &t, &t, &da); f_mul(&FQ, &t, &t, &db);
    f_add(&FQ, quad, quad, &t);
  }
}
/* NIFS round 0 on the i64 layers (neutronnova_zk.rs:781-810): e0 = 0, quad = sum_p w_p * prove_helper_small(pair p) */
EXPORT void orc_nifs_round0_small(size_t ell_b, const fe *rhos, size_t left, size_t right, const fe *E, const fe *A, const fe *B,
                                  const int64_t *A64, const int64_t *B64, const uint64_t *large_positions, size_t n_large, size_t N, size_t m,
                                  fe *out2) {
  fe q; f_zero(&q); f_zero(&out2[0]);
  fe *pqs = (fe *)malloc((m / 2 + 1) * sizeof(fe));
#pragma omp parallel for schedule(dynamic, 1) if (g_threads > 1)      /* par_iter over the pairs (neutronnova_zk.rs:781-810) */
  for (size_t p = 0; p < m / 2; p++)
    nifs_prove_helper_small(left, right, E, A + 2 * p * N, B + 2 * p * N, A + (2 * p + 1) * N, B + (2 * p + 1) * N,
                            A64 + 2 * p * N, B64 + 2 * p * N, A64 + (2 * p + 1) * N, B64 + (2 * p + 1) * N, large_positions, n_large, &pqs[p]);
  for (size_t p = 0; p < m / 2; p++) {
    fe w, tmp;
    suffix_weight_full(0, ell_b, p, rhos, &w);
    f_mul(&FQ, &tmp, &pqs[p], &w); f_add(&FQ, &q, &q, &tmp);
  }
  free(pqs);
  out2[1] = q;
}
/* c_vals (neutronnova_zk.rs:649-693): c_vals[b] = sum_k E[k] * Cz_b[k] from the i64 layer, corrected at large positions */
EXPORT void orc_nifs_cvals_small(size_t left, size_t right, const fe *E, const fe *Cl, const int64_t *C64, const uint64_t *large_positions,
                                 size_t n_large, size_t N, size_t n, fe *vals) {
  const fe *f = E + left, *el = E;
  for (size_t b = 0; b < n; b++) {
    acc9 acc; memset(&acc, 0, sizeof(acc));
    for (size_t i = 0; i < right; i++) {
      small_acc in; memset(&in, 0, sizeof(in));
      for (size_t j = 0; j < left; j++) small_acc_accumulate(&in, &el[j], (__int128)C64[b * N + i * left + j]);
      fe red; small_acc_reduce(&red, &in);
      f_mul_acc(&acc, &f[i], &red);
    }
    f_reduce9(&FQ, &vals[b], &acc);
  }
  for (size_t q = 0; q < n_large; q++) {
    const size_t k = large_positions[q];
    if (k >= left * right) continue;
    fe ef; f_mul(&FQ, &ef, &el[k % left], &f[k / left]);
    for (size_t b = 0; b < n; b++) { fe t; f_mul(&FQ, &t, &ef, &Cl[b * N + k]); f_add(&FQ, &vals[b], &vals[b], &t); }
  }
}

/* ------------------------------------------------------------------------------------
 * Hyrax commitment family on group elements: commit_without_blind / commit_incremental (hyrax_pc.rs:533-607),
 * rerandomize_commitment (:321-344), fold_blinds (:795-819), fold_commitments_partial (:821-874)
 * ---------------------------------------------------------------------------------- */
/* commit_without_blind (hyrax_pc.rs:533-567): raw row points, identity (0,0) for an all-zero row */
EXPORT void orc_hyrax_commit_without_blind(const apt *ck, size_t num_cols, const fe *v, size_t n, int is_small, apt *out) {
  size_t rows = (n + num_cols - 1) / num_cols;
#pragma omp parallel for schedule(dynamic, 1) if (g_threads > 1)
  for (size_t i = 0; i < rows; i++) {
    size_t lo = i * num_cols, hi = lo + num_cols < n ? lo + num_cols : n, len = hi - lo;
    const fe *s = v + lo; pt acc; int all_zero = 1;
    for (size_t j = 0; j < len; j++) if (!f_is_zero(&s[j])) { all_zero = 0; break; }
    if (all_zero) { f_zero(&out[i].x); f_zero(&out[i].y); continue; }
    if (is_small) {
      uint64_t *sm = (uint64_t *)malloc(len * 8);
      for (size_t j = 0; j < len; j++) { uint64_t raw[4]; f_to_raw(&FQ, raw, &s[j]); sm[j] = raw[0]; }
      msm_small_serial(sm, ck, len, &acc); free(sm);
    } else msm(s, ck, len, 0, &acc);
    pt_to_affine(&CV, &out[i], &acc);
  }
}
/* commit_incremental (hyrax_pc.rs:569-607): raw[i] + <delta_row_i, ck> + blind_i * h */
EXPORT void orc_hyrax_commit_incremental(const apt *ck, size_t num_cols, const apt *h, const apt *raw, size_t n_raw, const fe *delta, size_t n,
                                         const fe *blinds, apt *out) {
  size_t rows = (n + num_cols - 1) / num_cols;
#pragma omp parallel for schedule(dynamic, 1) if (g_threads > 1)
  for (size_t i = 0; i < rows; i++) {
    size_t lo = i * num_cols, hi = lo + num_cols < n ? lo + num_cols : n, len = hi - lo;
    const fe *s = delta + lo; pt acc, hb; int all_zero = 1;
    for (size_t j = 0; j < len; j++) if (!f_is_zero(&s[j])) { all_zero = 0; break; }
    if (i < n_raw) pt_from_affine(&CV, &acc, &raw[i]); else pt_set_inf(&CV, &acc);
    if (!all_zero) { pt d; msm(s, ck, len, 0, &d); pt_add(&CV, &acc, &acc, &d); }
    pt_mul_fixed(&hb, h, &blinds[i]); pt_add(&CV, &acc, &acc, &hb);
    pt_to_affine(&CV, &out[i], &acc);
  }
}
/* rerandomize_commitment (hyrax_pc.rs:321-344): comm[i] + (r_new[i] - r_old[i]) * h */
EXPORT void orc_hyrax_rerandomize(const apt *h, const apt *comm, const fe *r_old, const fe *r_new, size_t rows, apt *out) {
#pragma omp parallel for schedule(static) if (g_threads > 1)
  for (size_t i = 0; i < rows; i++) {
    fe d; f_sub(&FQ, &d, &r_new[i], &r_old[i]);
    pt a, hb; pt_from_affine(&CV, &a, &comm[i]); pt_mul_fixed(&hb, h, &d); pt_add(&CV, &a, &a, &hb);
    pt_to_affine(&CV, &out[i], &a);
  }
}
/* fold_blinds (hyrax_pc.rs:795-819): out[row] = sum_k w_k * blinds[k][row] */
EXPORT void orc_fold_blinds(const fe *blinds, size_t n, size_t rows, const fe *w, fe *out) {
  for (size_t r = 0; r < rows; r++) {
    fe acc; f_zero(&acc);
    for (size_t k = 0; k < n; k++) { fe t; f_mul(&FQ, &t, &blinds[k * rows + r], &w[k]); f_add(&FQ, &acc, &acc, &t); }
    out[r] = acc;
  }
}
/* fold_commitments_partial (hyrax_pc.rs:821-874): MSM-fold the data rows, rest rows = folded_blind[row] * h */
EXPORT void orc_fold_commitments_partial(const apt *comms, size_t n, size_t rows, const fe *w, size_t num_data_rows, const fe *folded_blind,
                                         const apt *h, apt *out) {
  if (num_data_rows >= rows) { orc_fold_commitments(comms, n, rows, w, out); return; }
#pragma omp parallel for schedule(dynamic, 1) if (g_threads > 1)
  for (size_t r = 0; r < rows; r++) {
    pt acc; pt_set_inf(&CV, &acc);
    if (r < num_data_rows) { apt *rb = (apt *)malloc(n * sizeof(apt)); for (size_t i = 0; i < n; i++) rb[i] = comms[i * rows + r]; msm_serial(w, rb, n, &acc); free(rb); }
    else pt_mul_fixed(&acc, h, &folded_blind[r]);
    pt_to_affine(&CV, &out[r], &acc);
  }
}

/* ------------------------------------------------------------------------------------
 * NeutronNova, NON-ZK variant: NeutronNovaZkSNARK::{prove, verify} (neutronnova_zk.rs:1609-2093, 2096-2330) with the ZK
 * wrapper removed.  The data path, the per-round scalar algebra, the instance / witness / commitment folds and the
 * closing PCS::prove are the reference's; what differs is WHERE the challenges come from: the reference commits every
 * round message inside its in-circuit verifier (process_round, bellpepper/r1cs.rs:735-816: out of scope) and
 * squeezes after that commitment; here the round polynomials are absorbed directly (b"p", all coefficients as scalars)
 * and the challenge squeezed (b"c"), as the non-ZK SpartanSNARK does (sumcheck.rs:536-548), eval_W_{step,core} are
 * revealed with their blinds (as SpartanSNARK reveals eval_W, spartan.rs:126-139) and their commitments absorbed before
 * c_eval.  The verifier below checks exactly what NeutronNovaVerifierCircuit (zk.rs:600-940) + verify (:2096-2330) check.
 * Requires num_shared == 0 and num_challenges == 0 (the SHA-256 chain of benches/sha256_neutronnova.rs).
 * ---------------------------------------------------------------------------------- */
typedef struct {
  uint64_t n_steps, ell_b, ell, rounds_y, rows, num_cols;
  apt *comm_W_steps;       /* n_steps * rows : U_i.comm_W (precommitted rows rerandomised, rest rows = blind * h) */
  apt *comm_W_core;        /* rows */
  fe *nifs_polys;          /* ell_b * 4 */
  fe *outer_polys;         /* ell * 8 : step | core cubic per round */
  fe *claims_outer;        /* 6 */
  fe *inner_polys;         /* rounds_y * 6 : step | core quadratic per round */
  fe *eval_W, *blind_eval_W;   /* 2 each: step, core */
  apt *delta, *beta; fe *z_vec; fe *z_delta, *z_beta;
} nn_proof_view;
typedef struct {
  const fe *blinds_steps;  /* n_steps * rows: the blinds of U_i.comm_W (fresh per prove: rerandomisation + rest rows) */
  const fe *blinds_core;   /* rows */
  const fe *blind_eval_W;  /* 2 */
  const fe *d_vec, *r_delta, *r_beta;
} nn_rand_view;

static void nn_absorb_instance(transcript *ts, const char *label, const apt *comm, size_t rows, const fe *X, size_t nx) {
  /* R1CSInstance::to_transcript_bytes (r1cs/mod.rs:728-736): comm_W bytes then X */
  ts_push(ts, label, strlen(label));
  ts_push(ts, "poly_commitment_begin", 21);
  for (size_t i = 0; i < rows; i++) { uint8_t b[64]; apt_to_bytes(&comm[i], b); ts_push(ts, b, 64); }
  ts_push(ts, "poly_commitment_end", 19);
  for (size_t i = 0; i < nx; i++) { uint8_t b[32]; fe_to_be_bytes(&FQ, &X[i], b); ts_push(ts, b, 32); }
}
static void nn_pow_eval(const fe *tau, const fe *rx, size_t ell, fe *out) {
  /* PowPolynomial::evaluate (polys/power.rs): prod_i (1 + (tau^(2^(ell-1-i)) - 1) r_i), first challenge = top variable */
  fe *tp = (fe *)malloc((ell + 1) * sizeof(fe)); tp[0] = *tau;
  for (size_t k = 1; k <= ell; k++) f_sqr(&FQ, &tp[k], &tp[k - 1]);
  fe one, acc, t; f_one(&FQ, &one); acc = one;
  for (size_t i = 0; i < ell; i++) { f_sub(&FQ, &t, &tp[ell - 1 - i], &one); f_mul(&FQ, &t, &t, &rx[i]); f_add(&FQ, &t, &t, &one); f_mul(&FQ, &acc, &acc, &t); }
  *out = acc; free(tp);
}

/* zs: n * num_cols (z_i = [W_i | 1 | X_i]); zc: num_cols.  comm_pre_*: the PrecommittedState commitments of prep_prove
 * (pre_rows per instance) with their blinds.  debug (optional, 8 fe): T_out, tau_at_rx, r, c_eval, eq_rho_at_rb.
 * phase_ms (optional, 8 doubles): rerandomize+commit_zeros, transcript absorb, NIFS, fold (witness, blinds, instances),
 * outer, poly_ABC, inner, pcs. */
EXPORT int orc_neutronnova_prove(const shape *S, const keys_view *K, const uint8_t vk_digest[32], size_t n, const fe *zs, const fe *zc,
                                 const apt *comm_pre_steps, const fe *blinds_pre_steps, const apt *comm_pre_core, const fe *blinds_pre_core,
                                 const nn_rand_view *R, nn_proof_view *P, fe *debug, double *phase_ms) {
  double t_last = 0; (void)t_last;
#ifdef _OPENMP
  t_last = omp_get_wtime();
#endif
  if (S->num_shared || S->num_challenges) return -7;
  const size_t N = S->num_cons, M = S->num_vars, ncols = S->num_vars + 1 + S->num_public, width = K->num_cols;
  const size_t rows = M / width, pre_rows = S->num_precommitted / width, np = S->num_public;
  size_t ell_b = 0; while (((size_t)1 << ell_b) < n) ell_b++;
  size_t ell = 0; while (((size_t)1 << ell) < N) ell++;
  size_t my = 0; while (((size_t)1 << my) < 2 * M) my++;
  const size_t left = (size_t)1 << ((ell + 1) / 2), right = (size_t)1 << (ell / 2);
  if (((size_t)1 << ell_b) != n || n < 2) return -2;
  P->n_steps = n; P->ell_b = ell_b; P->ell = ell; P->rounds_y = my; P->rows = rows; P->num_cols = width;
  fe one, zero; f_one(&FQ, &one); f_zero(&zero);
  /* rerandomize_in_place (bellpepper/r1cs.rs:568-602 -> hyrax_pc.rs:321-344) + commit_zeros of the rest rows (r1cs.rs:467-470) */
  for (size_t i = 0; i <= n; i++) {
    const apt *cp = i < n ? comm_pre_steps + i * pre_rows : comm_pre_core;
    const fe *bo = i < n ? blinds_pre_steps + i * pre_rows : blinds_pre_core;
    const fe *bn = i < n ? R->blinds_steps + i * rows : R->blinds_core;
    apt *out = i < n ? P->comm_W_steps + i * rows : P->comm_W_core;
    orc_hyrax_rerandomize(K->h, cp, bo, bn, pre_rows, out);
    for (size_t r = pre_rows; r < rows; r++) { pt hb; pt_mul_fixed(&hb, K->h, &bn[r]); pt_to_affine(&CV, &out[r], &hb); }   /* commit_zeros (hyrax_pc.rs:305-319) */
  }
  TICK(0);
  transcript ts; ts_new(&ts, "neutronnova_prove");
  ts_absorb_bytes(&ts, "vk", vk_digest, 32);
  nn_absorb_instance(&ts, "core_instance", P->comm_W_core, rows, zc + M + 1, np);
  for (size_t i = 0; i < n; i++) nn_absorb_instance(&ts, "U", P->comm_W_steps + i * rows, rows, zs + i * ncols + M + 1, np);
  ts_absorb_scalars(&ts, &FQ, "T", &zero, 1);
  fe tau; ts_squeeze(&ts, &FQ, "tau", &tau);
  fe *E = (fe *)malloc((left + right) * sizeof(fe)); orc_pow_split_evals(&tau, left, right, E);
  fe *rhos = (fe *)malloc((ell_b + 1) * sizeof(fe));
  for (size_t t = 0; t < ell_b; t++) ts_squeeze(&ts, &FQ, "rho", &rhos[t]);
  TICK(1);
  /* layers (prep_prove's cached matvec, neutronnova_zk.rs:1538-1549) */
  fe *A = (fe *)malloc(n * N * sizeof(fe)), *B = (fe *)malloc(n * N * sizeof(fe)), *Cm = (fe *)malloc(n * N * sizeof(fe));
  for (size_t i = 0; i < n; i++) orc_shape_multiply_vec(S, zs + i * ncols, A + i * N, B + i * N, Cm + i * N);
  fe *Ac = (fe *)malloc(N * sizeof(fe)), *Bc = (fe *)malloc(N * sizeof(fe)), *Cc = (fe *)malloc(N * sizeof(fe));
  orc_shape_multiply_vec(S, zc, Ac, Bc, Cc);
  /* i64 copies of the A / B layers with the union of the large positions zeroed (prep_prove, neutronnova_zk.rs:1551-1584) */
  int64_t *A64 = (int64_t *)malloc(n * N * sizeof(int64_t)), *B64 = (int64_t *)malloc(n * N * sizeof(int64_t));
  uint64_t *lp = (uint64_t *)malloc(N * sizeof(uint64_t)), *tmp_pos = (uint64_t *)malloc(N * sizeof(uint64_t));
  uint8_t *is_large = (uint8_t *)calloc(N, 1); size_t n_large = 0;
  for (size_t i = 0; i < n; i++) {
    size_t k = orc_to_small_vec_or_zero(A + i * N, N, A64 + i * N, tmp_pos); for (size_t q = 0; q < k; q++) is_large[tmp_pos[q]] = 1;
    k = orc_to_small_vec_or_zero(B + i * N, N, B64 + i * N, tmp_pos); for (size_t q = 0; q < k; q++) is_large[tmp_pos[q]] = 1;
  }
  for (size_t k = 0; k < N; k++) if (is_large[k]) { lp[n_large++] = k; for (size_t i = 0; i < n; i++) { A64[i * N + k] = 0; B64[i * N + k] = 0; } }
  free(tmp_pos); free(is_large);
#ifdef _OPENMP
  t_last = omp_get_wtime();                                  /* the matvec and the i64 conversion are prep_prove work: untimed */
#endif
  /* HOT LOOP A: NIFS rounds (neutronnova_zk.rs:778-1168; round 0 on the i64 layers, prove_helper_small :255-325), finish_round! (:703-735) */
  fe T_cur = zero, acc_eq = one, *r_b = (fe *)malloc((ell_b + 1) * sizeof(fe));
  size_t m = n;
  for (size_t t = 0; t < ell_b; t++) {
    fe e0q[2];
    if (t == 0) orc_nifs_round0_small(ell_b, rhos, left, right, E, A, B, A64, B64, lp, n_large, N, m, e0q);
    else orc_nifs_round(t, ell_b, rhos, left, right, E, A, B, Cm, N, m, e0q);
    fe rho = rhos[t], omr, trm, c, a, abc, b, rinv, tmp, tmp2, co[4];
    f_sub(&FQ, &omr, &one, &rho); f_sub(&FQ, &trm, &rho, &omr);
    f_mul(&FQ, &c, &e0q[0], &acc_eq); f_mul(&FQ, &a, &e0q[1], &acc_eq);
    if (!f_inv(&FQ, &rinv, &rho)) return -5;
    f_mul(&FQ, &tmp, &c, &omr); f_sub(&FQ, &tmp, &T_cur, &tmp); f_mul(&FQ, &abc, &tmp, &rinv);
    f_sub(&FQ, &b, &abc, &a); f_sub(&FQ, &b, &b, &c);
    f_mul(&FQ, &co[0], &c, &omr);
    f_mul(&FQ, &tmp, &c, &trm); f_mul(&FQ, &tmp2, &b, &omr); f_add(&FQ, &co[1], &tmp, &tmp2);
    f_mul(&FQ, &tmp, &b, &trm); f_mul(&FQ, &tmp2, &a, &omr); f_add(&FQ, &co[2], &tmp, &tmp2);
    f_mul(&FQ, &co[3], &a, &trm);
    memcpy(&P->nifs_polys[4 * t], co, sizeof(co));
    ts_absorb_scalars(&ts, &FQ, "p", co, 4);
    ts_squeeze(&ts, &FQ, "c", &r_b[t]);
    f_sub(&FQ, &tmp, &one, &r_b[t]); f_mul(&FQ, &tmp, &tmp, &omr); f_mul(&FQ, &tmp2, &r_b[t], &rho); f_add(&FQ, &tmp, &tmp, &tmp2);
    f_mul(&FQ, &acc_eq, &acc_eq, &tmp);
    unipoly_eval(co, 4, &r_b[t], &T_cur);
    orc_nifs_fold(A, N, m, &r_b[t]); orc_nifs_fold(B, N, m, &r_b[t]); orc_nifs_fold(Cm, N, m, &r_b[t]);
    m /= 2;
  }
  fe T_out, ainv; if (!f_inv(&FQ, &ainv, &acc_eq)) return -5;
  f_mul(&FQ, &T_out, &T_cur, &ainv);
  TICK(2);
  /* fold_multiple (r1cs/mod.rs:570-660), fold_blinds, X fold, fold_commitments_partial (neutronnova_zk.rs:1212-1262) */
  fe *w = (fe *)malloc(n * sizeof(fe)); orc_weights_from_r(r_b, ell_b, n, w);
  fe *Wf = (fe *)calloc(2 * M, sizeof(fe)), *Wc = (fe *)calloc(2 * M, sizeof(fe));     /* z tables of the inner sum-check */
#pragma omp parallel for schedule(static) if (g_threads > 1)
  for (size_t j = 0; j < M; j++) { acc9 a; memset(&a, 0, sizeof(a)); for (size_t i = 0; i < n; i++) f_mul_acc(&a, &w[i], &zs[i * ncols + j]); f_reduce9(&FQ, &Wf[j], &a); }
  fe *blind_fold = (fe *)malloc(rows * sizeof(fe)); orc_fold_blinds(R->blinds_steps, n, rows, w, blind_fold);
  fe *X_acc = (fe *)calloc(np + 1, sizeof(fe));
  for (size_t i = 0; i < n; i++) for (size_t j = 0; j < np; j++) { fe t; f_mul(&FQ, &t, &w[i], &zs[i * ncols + M + 1 + j]); f_add(&FQ, &X_acc[j], &X_acc[j], &t); }
  apt *comm_fold = (apt *)malloc(rows * sizeof(apt));
  orc_fold_commitments_partial(P->comm_W_steps, n, rows, w, pre_rows, blind_fold, K->h, comm_fold);
  Wf[M] = one; memcpy(&Wf[M + 1], X_acc, np * sizeof(fe));
  memcpy(Wc, zc, M * sizeof(fe)); Wc[M] = one; memcpy(&Wc[M + 1], zc + M + 1, np * sizeof(fe));
  fe *W_fold = (fe *)malloc(M * sizeof(fe)); memcpy(W_fold, Wf, M * sizeof(fe));
  TICK(3);
  /* HOT LOOP B: prove_cubic_with_additive_term_batched (sumcheck.rs:786-917) */
  fe *tau_pow = (fe *)malloc((ell + 1) * sizeof(fe)); tau_pow[0] = tau;
  for (size_t k = 1; k <= ell; k++) f_sqr(&FQ, &tau_pow[k], &tau_pow[k - 1]);
  fe base_tau = one, claim_s = T_out, claim_c = zero, *r_x = (fe *)malloc(ell * sizeof(fe));
  size_t tl = N;
  for (size_t i = 0; i < ell; i++) {
    fe ev[6], evl[8], co[8], tmp;
    orc_pow_cubic_eval(E, left, E + left, A, B, Cm, tl, ev); orc_pow_cubic_eval(E, left, E + left, Ac, Bc, Cc, tl, ev + 3);
    for (int k = 0; k < 6; k++) f_mul(&FQ, &ev[k], &ev[k], &base_tau);
    evl[0] = ev[0]; f_sub(&FQ, &evl[1], &claim_s, &ev[0]); evl[2] = ev[1]; evl[3] = ev[2];
    evl[4] = ev[3]; f_sub(&FQ, &evl[5], &claim_c, &ev[3]); evl[6] = ev[4]; evl[7] = ev[5];
    unipoly_from_evals_deg3(evl, co); unipoly_from_evals_deg3(evl + 4, co + 4);
    memcpy(&P->outer_polys[8 * i], co, sizeof(co));
    ts_absorb_scalars(&ts, &FQ, "p", co, 8);
    ts_squeeze(&ts, &FQ, "c", &r_x[i]);
    unipoly_eval(co, 4, &r_x[i], &claim_s); unipoly_eval(co + 4, 4, &r_x[i], &claim_c);
    bind_top(A, tl, &r_x[i]); bind_top(B, tl, &r_x[i]); bind_top(Cm, tl, &r_x[i]);
    bind_top(Ac, tl, &r_x[i]); bind_top(Bc, tl, &r_x[i]); bind_top(Cc, tl, &r_x[i]);
    tl /= 2;
    f_sub(&FQ, &tmp, &tau_pow[ell - 1 - i], &one); f_mul(&FQ, &tmp, &tmp, &r_x[i]); f_add(&FQ, &tmp, &tmp, &one);
    f_mul(&FQ, &base_tau, &base_tau, &tmp);
  }
  fe cl[6] = { A[0], B[0], Cm[0], Ac[0], Bc[0], Cc[0] };
  memcpy(P->claims_outer, cl, sizeof(cl));
  TICK(4);
  ts_absorb_scalars(&ts, &FQ, "claims_outer", cl, 6);
  fe r, r2, claim_js, claim_jc, t1; ts_squeeze(&ts, &FQ, "r", &r); f_sqr(&FQ, &r2, &r);
  f_mul(&FQ, &t1, &r, &cl[1]); f_add(&FQ, &claim_js, &cl[0], &t1); f_mul(&FQ, &t1, &r2, &cl[2]); f_add(&FQ, &claim_js, &claim_js, &t1);
  f_mul(&FQ, &t1, &r, &cl[4]); f_add(&FQ, &claim_jc, &cl[3], &t1); f_mul(&FQ, &t1, &r2, &cl[5]); f_add(&FQ, &claim_jc, &claim_jc, &t1);
  fe *rx = (fe *)malloc(N * sizeof(fe)); eq_evals(r_x, ell, rx);
  fe *abc_s = (fe *)malloc(2 * M * sizeof(fe)), *abc_c = (fe *)malloc(2 * M * sizeof(fe));
  orc_shape_abc(S, rx, &r, abc_s, 2 * M); orc_shape_abc(S, rx, &r, abc_c, 2 * M);
  TICK(5);
  /* HOT LOOP C: prove_quad_batched (sumcheck.rs:702-782) */
  fe *r_y = (fe *)malloc(my * sizeof(fe));
  tl = 2 * M;
  for (size_t j = 0; j < my; j++) {
    fe e[4], co[6];
    quad_points(abc_s, Wf, tl, &e[0], &e[1]); quad_points(abc_c, Wc, tl, &e[2], &e[3]);
    quad_round_poly(&e[0], &e[1], &claim_js, co); quad_round_poly(&e[2], &e[3], &claim_jc, co + 3);
    memcpy(&P->inner_polys[6 * j], co, sizeof(co));
    ts_absorb_scalars(&ts, &FQ, "p", co, 6);
    ts_squeeze(&ts, &FQ, "c", &r_y[j]);
    bind_top(abc_s, tl, &r_y[j]); bind_top(Wf, tl, &r_y[j]); bind_top(abc_c, tl, &r_y[j]); bind_top(Wc, tl, &r_y[j]);
    tl /= 2;
    unipoly_eval(co, 3, &r_y[j], &claim_js); unipoly_eval(co + 3, 3, &r_y[j], &claim_jc);
  }
  /* eval_W (neutronnova_zk.rs:1930-1951) */
  fe *Xs = (fe *)malloc((np + 1) * sizeof(fe)), *Xc = (fe *)malloc((np + 1) * sizeof(fe));
  Xs[0] = one; memcpy(&Xs[1], X_acc, np * sizeof(fe)); Xc[0] = one; memcpy(&Xc[1], zc + M + 1, np * sizeof(fe));
  fe eXs, eXc, den, dinv; sparse_poly_eval(my - 1, Xs, np + 1, r_y + 1, &eXs); sparse_poly_eval(my - 1, Xc, np + 1, r_y + 1, &eXc);
  f_sub(&FQ, &den, &one, &r_y[0]); if (!f_inv(&FQ, &dinv, &den)) return -5;
  f_mul(&FQ, &t1, &r_y[0], &eXs); f_sub(&FQ, &P->eval_W[0], &Wf[0], &t1); f_mul(&FQ, &P->eval_W[0], &P->eval_W[0], &dinv);
  f_mul(&FQ, &t1, &r_y[0], &eXc); f_sub(&FQ, &P->eval_W[1], &Wc[0], &t1); f_mul(&FQ, &P->eval_W[1], &P->eval_W[1], &dinv);
  P->blind_eval_W[0] = R->blind_eval_W[0]; P->blind_eval_W[1] = R->blind_eval_W[1];
  TICK(6);
  /* commitments to the two evaluations (the reference's per-round commits of eval_W_step / eval_W_core), c_eval, folds
   * (neutronnova_zk.rs:2019-2051), PCS::prove (:2053-2064) */
  apt ce[2];
  for (int b = 0; b < 2; b++) { pt a, hb; pt_mul_fe(&a, K->ck_s, &P->eval_W[b]); pt_mul_fe(&hb, K->h_s, &R->blind_eval_W[b]); pt_add(&CV, &a, &a, &hb); pt_to_affine(&CV, &ce[b], &a); }
  ts_absorb_point(&ts, "comm_eval_W_step", &ce[0]); ts_absorb_point(&ts, "comm_eval_W_core", &ce[1]);
  fe c_eval; ts_squeeze(&ts, &FQ, "c_eval", &c_eval);
  fe wts[2] = { one, c_eval };
  apt *pair = (apt *)malloc(2 * rows * sizeof(apt)), *comm = (apt *)malloc(rows * sizeof(apt));
  memcpy(pair, comm_fold, rows * sizeof(apt)); memcpy(pair + rows, P->comm_W_core, rows * sizeof(apt));
  orc_fold_commitments(pair, 2, rows, wts, comm);
  fe *blind = (fe *)malloc(rows * sizeof(fe));
  for (size_t i = 0; i < rows; i++) { f_mul(&FQ, &t1, &c_eval, &R->blinds_core[i]); f_add(&FQ, &blind[i], &blind_fold[i], &t1); }
  fe *Wfin = (fe *)malloc(M * sizeof(fe));
  for (size_t j = 0; j < M; j++) { f_mul(&FQ, &t1, &c_eval, &zc[j]); f_add(&FQ, &Wfin[j], &W_fold[j], &t1); }
  apt comm_eval; orc_fold_commitments(ce, 2, 1, wts, &comm_eval);
  fe blind_eval; f_mul(&FQ, &t1, &c_eval, &R->blind_eval_W[1]); f_add(&FQ, &blind_eval, &R->blind_eval_W[0], &t1);
  rand_view RV = { blind, &blind_eval, R->d_vec, R->r_delta, R->r_beta };
  proof_view PV; memset(&PV, 0, sizeof(PV)); PV.delta = P->delta; PV.beta = P->beta; PV.z_vec = P->z_vec; PV.z_delta = P->z_delta; PV.z_beta = P->z_beta;
  hyrax_prove(K, &ts, comm, rows, Wfin, M, blind, r_y + 1, my - 1, &comm_eval, &RV, &PV, NULL);
  TICK(7);
  if (debug) { debug[0] = T_out; debug[1] = base_tau; debug[2] = r; debug[3] = c_eval; debug[4] = acc_eq; debug[5] = eXs; debug[6] = eXc; }
  free(E); free(rhos); free(A); free(B); free(Cm); free(Ac); free(Bc); free(Cc); free(r_b); free(w); free(Wf); free(Wc); free(blind_fold);
  free(X_acc); free(comm_fold); free(W_fold); free(tau_pow); free(r_x); free(rx); free(abc_s); free(abc_c); free(r_y); free(Xs); free(Xc);
  free(pair); free(comm); free(blind); free(Wfin); free(A64); free(B64); free(lp);
  ts_free(&ts);
  return 0;
}

/* step_X: n * num_public public IO of the step instances; core_X: num_public.  0 = accept; negative = the failing check:
 * -2 NIFS round / final, -3 outer rounds / final, -4 inner rounds / final (matrix evaluations), -5 PCS. */
EXPORT int orc_neutronnova_verify(const shape *S, const keys_view *K, const uint8_t vk_digest[32], const fe *step_X, const fe *core_X,
                                  const nn_proof_view *P) {
  if (S->num_shared || S->num_challenges) return -7;
  const size_t N = S->num_cons, M = S->num_vars, width = K->num_cols, rows = M / width, np = S->num_public, n = P->n_steps;
  size_t ell_b = 0; while (((size_t)1 << ell_b) < n) ell_b++;
  size_t ell = 0; while (((size_t)1 << ell) < N) ell++;
  size_t my = 0; while (((size_t)1 << my) < 2 * M) my++;
  if (P->ell_b != ell_b || P->ell != ell || P->rounds_y != my || P->rows != rows || n < 2 || ((size_t)1 << ell_b) != n) return -1;
  fe one, zero; f_one(&FQ, &one); f_zero(&zero);
  transcript ts; ts_new(&ts, "neutronnova_prove");
  ts_absorb_bytes(&ts, "vk", vk_digest, 32);
  nn_absorb_instance(&ts, "core_instance", P->comm_W_core, rows, core_X, np);
  for (size_t i = 0; i < n; i++) nn_absorb_instance(&ts, "U", P->comm_W_steps + i * rows, rows, step_X + i * np, np);
  ts_absorb_scalars(&ts, &FQ, "T", &zero, 1);
  fe tau; ts_squeeze(&ts, &FQ, "tau", &tau);
  fe *rhos = (fe *)malloc((ell_b + 1) * sizeof(fe)), *r_b = (fe *)malloc((ell_b + 1) * sizeof(fe));
  for (size_t t = 0; t < ell_b; t++) ts_squeeze(&ts, &FQ, "rho", &rhos[t]);
  int rc = 0;
  /* NIFS rounds: p_t(0) + p_t(1) = claim (zk.rs:608-637, enforce_sc_claim); final: eq_rho_at_rb * T_out = p_last(r_last) (:638-661) */
  fe claim = zero, eq_rho = one, t1, t2;
  for (size_t t = 0; t < ell_b; t++) {
    const fe *co = &P->nifs_polys[4 * t];
    fe s; f_dbl(&FQ, &s, &co[0]); f_add(&FQ, &s, &s, &co[1]); f_add(&FQ, &s, &s, &co[2]); f_add(&FQ, &s, &s, &co[3]);
    if (!f_eq(&s, &claim)) rc = rc ? rc : -2;
    ts_absorb_scalars(&ts, &FQ, "p", co, 4);
    ts_squeeze(&ts, &FQ, "c", &r_b[t]);
    unipoly_eval(co, 4, &r_b[t], &claim);
    f_sub(&FQ, &t1, &one, &r_b[t]); f_sub(&FQ, &t2, &one, &rhos[t]); f_mul(&FQ, &t1, &t1, &t2); f_mul(&FQ, &t2, &r_b[t], &rhos[t]); f_add(&FQ, &t1, &t1, &t2);
    f_mul(&FQ, &eq_rho, &eq_rho, &t1);                       /* EqPolynomial::new(r_b).evaluate(&rhos) (:2283) */
  }
  fe T_out, einv; if (!f_inv(&FQ, &einv, &eq_rho)) return -2;
  f_mul(&FQ, &T_out, &claim, &einv);
  /* outer: step claim_0 = T_out, core claim_0 = 0 (zk.rs:674-700); final tau(r_x) (Az Bz - Cz) = claim (:745-760) */
  fe cs = T_out, cc = zero, *r_x = (fe *)malloc(ell * sizeof(fe));
  for (size_t i = 0; i < ell; i++) {
    const fe *co = &P->outer_polys[8 * i];
    for (int b = 0; b < 2; b++) {
      const fe *c = co + 4 * b; fe s; f_dbl(&FQ, &s, &c[0]); f_add(&FQ, &s, &s, &c[1]); f_add(&FQ, &s, &s, &c[2]); f_add(&FQ, &s, &s, &c[3]);
      if (!f_eq(&s, b ? &cc : &cs)) rc = rc ? rc : -3;
    }
    ts_absorb_scalars(&ts, &FQ, "p", co, 8);
    ts_squeeze(&ts, &FQ, "c", &r_x[i]);
    unipoly_eval(co, 4, &r_x[i], &cs); unipoly_eval(co + 4, 4, &r_x[i], &cc);
  }
  fe tau_rx; nn_pow_eval(&tau, r_x, ell, &tau_rx);
  const fe *cl = P->claims_outer;
  f_mul(&FQ, &t1, &cl[0], &cl[1]); f_sub(&FQ, &t1, &t1, &cl[2]); f_mul(&FQ, &t1, &t1, &tau_rx); if (!f_eq(&t1, &cs)) rc = rc ? rc : -3;
  f_mul(&FQ, &t1, &cl[3], &cl[4]); f_sub(&FQ, &t1, &t1, &cl[5]); f_mul(&FQ, &t1, &t1, &tau_rx); if (!f_eq(&t1, &cc)) rc = rc ? rc : -3;
  ts_absorb_scalars(&ts, &FQ, "claims_outer", cl, 6);
  fe r, r2, js, jc; ts_squeeze(&ts, &FQ, "r", &r); f_sqr(&FQ, &r2, &r);
  f_mul(&FQ, &t1, &r, &cl[1]); f_add(&FQ, &js, &cl[0], &t1); f_mul(&FQ, &t1, &r2, &cl[2]); f_add(&FQ, &js, &js, &t1);
  f_mul(&FQ, &t1, &r, &cl[4]); f_add(&FQ, &jc, &cl[3], &t1); f_mul(&FQ, &t1, &r2, &cl[5]); f_add(&FQ, &jc, &jc, &t1);
  fe *r_y = (fe *)malloc(my * sizeof(fe));
  for (size_t j = 0; j < my; j++) {
    const fe *co = &P->inner_polys[6 * j];
    for (int b = 0; b < 2; b++) {
      const fe *c = co + 3 * b; fe s; f_dbl(&FQ, &s, &c[0]); f_add(&FQ, &s, &s, &c[1]); f_add(&FQ, &s, &s, &c[2]);
      if (!f_eq(&s, b ? &jc : &js)) rc = rc ? rc : -4;
    }
    ts_absorb_scalars(&ts, &FQ, "p", co, 6);
    ts_squeeze(&ts, &FQ, "c", &r_y[j]);
    unipoly_eval(co, 3, &r_y[j], &js); unipoly_eval(co + 3, 3, &r_y[j], &jc);
  }
  /* R1CSInstance::fold_multiple (X part), eval_X, matrix evaluations, inner final check (zk.rs:166-230; :2244-2300) */
  fe *w = (fe *)malloc(n * sizeof(fe)); orc_weights_from_r(r_b, ell_b, n, w);
  fe *Xs = (fe *)calloc(np + 1, sizeof(fe)), *Xc = (fe *)calloc(np + 1, sizeof(fe));
  Xs[0] = one; Xc[0] = one; memcpy(&Xc[1], core_X, np * sizeof(fe));
  for (size_t i = 0; i < n; i++) for (size_t j = 0; j < np; j++) { f_mul(&FQ, &t1, &w[i], &step_X[i * np + j]); f_add(&FQ, &Xs[1 + j], &Xs[1 + j], &t1); }
  fe eXs, eXc; sparse_poly_eval(my - 1, Xs, np + 1, r_y + 1, &eXs); sparse_poly_eval(my - 1, Xc, np + 1, r_y + 1, &eXc);
  fe *Tx = (fe *)malloc(N * sizeof(fe)), *Ty = (fe *)malloc(2 * M * sizeof(fe)), ev[3];
  eq_evals(r_x, ell, Tx); eq_evals(r_y, my, Ty);
  orc_shape_eval_tables(S, Tx, Ty, ev);
  fe quot; f_mul(&FQ, &t1, &r, &ev[1]); f_add(&FQ, &quot, &ev[0], &t1); f_mul(&FQ, &t1, &r2, &ev[2]); f_add(&FQ, &quot, &quot, &t1);
  fe omr; f_sub(&FQ, &omr, &one, &r_y[0]);
  f_mul(&FQ, &t1, &omr, &P->eval_W[0]); f_mul(&FQ, &t2, &r_y[0], &eXs); f_add(&FQ, &t1, &t1, &t2); f_mul(&FQ, &t1, &t1, &quot); if (!f_eq(&t1, &js)) rc = rc ? rc : -4;
  f_mul(&FQ, &t1, &omr, &P->eval_W[1]); f_mul(&FQ, &t2, &r_y[0], &eXc); f_add(&FQ, &t1, &t1, &t2); f_mul(&FQ, &t1, &t1, &quot); if (!f_eq(&t1, &jc)) rc = rc ? rc : -4;
  /* commitments: fold the step instances (fold_multiple: comm part as group elements), c_eval, PCS::verify (:2302-2326) */
  apt ce[2];
  for (int b = 0; b < 2; b++) { pt a, hb; pt_mul_fe(&a, K->ck_s, &P->eval_W[b]); pt_mul_fe(&hb, K->h_s, &P->blind_eval_W[b]); pt_add(&CV, &a, &a, &hb); pt_to_affine(&CV, &ce[b], &a); }
  ts_absorb_point(&ts, "comm_eval_W_step", &ce[0]); ts_absorb_point(&ts, "comm_eval_W_core", &ce[1]);
  fe c_eval; ts_squeeze(&ts, &FQ, "c_eval", &c_eval);
  apt *comm_fold = (apt *)malloc(rows * sizeof(apt)), *pair = (apt *)malloc(2 * rows * sizeof(apt)), *comm = (apt *)malloc(rows * sizeof(apt));
  orc_fold_commitments(P->comm_W_steps, n, rows, w, comm_fold);
  fe wts[2] = { one, c_eval };
  memcpy(pair, comm_fold, rows * sizeof(apt)); memcpy(pair + rows, P->comm_W_core, rows * sizeof(apt));
  orc_fold_commitments(pair, 2, rows, wts, comm);
  apt comm_eval; orc_fold_commitments(ce, 2, 1, wts, &comm_eval);
  proof_view PV; memset(&PV, 0, sizeof(PV)); PV.delta = P->delta; PV.beta = P->beta; PV.z_vec = P->z_vec; PV.z_delta = P->z_delta; PV.z_beta = P->z_beta;
  if (!rc && hyrax_verify(K, &ts, comm, rows, r_y + 1, my - 1, &comm_eval, &PV) != 0) rc = -5;
  free(rhos); free(r_b); free(r_x); free(r_y); free(w); free(Xs); free(Xc); free(Tx); free(Ty); free(comm_fold); free(pair); free(comm);
  ts_free(&ts);
  return rc;
}
