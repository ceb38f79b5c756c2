/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Keccak-256 (original Keccak padding 0x01, NOT SHA3-256) and the reference's
 * Fiat-Shamir transcript (reference src/provider/keccak.rs:18-105).
 * Keccak itself lives in the un-vendored dependency sha3 0.10 (reference
 * Cargo.toml:18); its published algorithm (FIPS-202 permutation, rate 136) is
 * restated here and pinned by the reference's known-answer test
 * (src/provider/keccak.rs:155-163).
 */
#ifndef ORACLE_KECCAK_H
#define ORACLE_KECCAK_H
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "field.h"

static const uint64_t KECCAK_RC[24] = {
  0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
  0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
  0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
  0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
  0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
  0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL };
static const int KECCAK_ROT[25] = { 0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39,
                                    41, 45, 15, 21, 8, 18, 2, 61, 56, 14 };

static inline uint64_t k_rotl(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static inline void keccak_f1600(uint64_t s[25]) {
  for (int round = 0; round < 24; round++) {
    uint64_t C[5], D[5], B[25];
    for (int x = 0; x < 5; x++) C[x] = s[x] ^ s[x + 5] ^ s[x + 10] ^ s[x + 15] ^ s[x + 20];
    for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ k_rotl(C[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) s[i] ^= D[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++)
        B[y + 5 * ((2 * x + 3 * y) % 5)] = k_rotl(s[x + 5 * y], KECCAK_ROT[x + 5 * y]);
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++)
        s[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
    s[0] ^= KECCAK_RC[round];
  }
}

static inline void keccak256(const uint8_t *in, size_t len, uint8_t out[32]) {
  uint64_t s[25]; memset(s, 0, sizeof(s));
  const size_t rate = 136;
  while (len >= rate) {
    for (size_t i = 0; i < rate / 8; i++) { uint64_t w; memcpy(&w, in + 8 * i, 8); s[i] ^= w; }
    keccak_f1600(s); in += rate; len -= rate;
  }
  uint8_t blk[136]; memset(blk, 0, sizeof(blk));
  memcpy(blk, in, len);
  blk[len] ^= 0x01; blk[rate - 1] ^= 0x80;
  for (size_t i = 0; i < rate / 8; i++) { uint64_t w; memcpy(&w, blk + 8 * i, 8); s[i] ^= w; }
  keccak_f1600(s);
  memcpy(out, s, 32);
}

/* ---- transcript (keccak.rs:26-105) ---- */
typedef struct {
  uint16_t round;
  uint8_t state[64];
  uint8_t *buf; size_t len, cap;     /* bytes absorbed since the last squeeze */
} transcript;

static inline void ts_push(transcript *t, const void *p, size_t n) {
  if (t->len + n > t->cap) {
    t->cap = (t->len + n) * 2 + 256;
    t->buf = (uint8_t *)realloc(t->buf, t->cap);
  }
  memcpy(t->buf + t->len, p, n); t->len += n;
}
/* compute_updated_state (keccak.rs:33-54): K(input||0x00) || K(input||0x01) */
static inline void ts_updated_state(const uint8_t *input, size_t n, uint8_t out[64]) {
  uint8_t *tmp = (uint8_t *)malloc(n + 1);
  memcpy(tmp, input, n);
  tmp[n] = 0; keccak256(tmp, n + 1, out);
  tmp[n] = 1; keccak256(tmp, n + 1, out + 32);
  free(tmp);
}
static inline void ts_new(transcript *t, const char *label) {      /* keccak.rs:57-68 */
  memset(t, 0, sizeof(*t));
  size_t n = strlen(label);
  uint8_t *in = (uint8_t *)malloc(4 + n);
  memcpy(in, "NoTR", 4); memcpy(in + 4, label, n);
  ts_updated_state(in, 4 + n, t->state);
  free(in);
}
static inline void ts_free(transcript *t) { free(t->buf); t->buf = NULL; t->len = t->cap = 0; }
static inline void ts_absorb_bytes(transcript *t, const char *label, const void *p, size_t n) {
  ts_push(t, label, strlen(label)); ts_push(t, p, n);              /* keccak.rs:96-99 */
}
static inline void ts_dom_sep(transcript *t, const char *bytes) {   /* keccak.rs:101-104 */
  ts_push(t, "NoDS", 4); ts_push(t, bytes, strlen(bytes));
}
static inline void ts_squeeze_bytes(transcript *t, const char *label, uint8_t out[64]) {
  uint8_t le[2] = { (uint8_t)(t->round & 0xff), (uint8_t)(t->round >> 8) };   /* keccak.rs:70-94 */
  ts_push(t, "NoDS", 4); ts_push(t, le, 2); ts_push(t, t->state, 64); ts_push(t, label, strlen(label));
  ts_updated_state(t->buf, t->len, out);
  t->round++;
  memcpy(t->state, out, 64);
  t->len = 0;
}
static inline void ts_squeeze(transcript *t, const fctx *F, const char *label, fe *out) {
  uint8_t o[64]; ts_squeeze_bytes(t, label, o); f_from_uniform(F, out, o);
}
/* scalar -> 32 bytes big-endian canonical (provider/traits.rs:282-286) */
static inline void fe_to_be_bytes(const fctx *F, const fe *a, uint8_t out[32]) {
  uint64_t raw[4]; f_to_raw(F, raw, a);
  for (int i = 0; i < 32; i++) out[31 - i] = (uint8_t)(raw[i >> 3] >> (8 * (i & 7)));
}
/* scalar -> 32 bytes little-endian canonical (to_repr; polys/univariate.rs:182-190) */
static inline void fe_to_le_bytes(const fctx *F, const fe *a, uint8_t out[32]) {
  uint64_t raw[4]; f_to_raw(F, raw, a); memcpy(out, raw, 32);
}
static inline void ts_absorb_scalars(transcript *t, const fctx *F, const char *label, const fe *v, size_t n) {
  ts_push(t, label, strlen(label));
  for (size_t i = 0; i < n; i++) { uint8_t b[32]; fe_to_be_bytes(F, &v[i], b); ts_push(t, b, 32); }
}
#endif
