// Host side of the Fiat-Shamir transcript (product code; the oracle has its own copy for checking).
// Restates reference src/provider/keccak.rs:18-105 (Keccak256Transcript) and the byte encodings of
// src/provider/traits.rs:275-305 (scalars big-endian, points x_BE || y_BE) and
// src/provider/pcs/hyrax_pc.rs:714-729 (commitment framing).  Used by the fused prover for the parts
// of the transcript that hash bulk data (commitment rows): a serial Keccak over tens of KB is a
// host-speed job, while the per-round sum-check squeezes stay on the device (keccak.cuh).
#pragma once
#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>

namespace sp2h {

typedef unsigned __int128 u128;

// ---- Keccak-256 (0x01 padding, not SHA-3) -------------------------------------------------------
inline uint64_t rotl(uint64_t v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; }
inline void keccak_f(uint64_t a[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
      0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
      0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
      0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  for (int rnd = 0; rnd < 24; rnd++) {
    uint64_t c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl(c[(x + 1) % 5], 1);
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl(a[x + 5 * y] ^ d[x], ROT[x + 5 * y]);
    for (int y = 0; y < 5; y++)
      for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= RC[rnd];
  }
}
inline void keccak256(const uint8_t *in, size_t n, uint8_t suffix_byte, bool with_suffix, uint8_t out[32]) {
  uint64_t a[25]; memset(a, 0, sizeof(a));
  const size_t total = n + (with_suffix ? 1 : 0);
  uint8_t blk[136];
  size_t off = 0;
  for (;;) {
    const size_t rem = total - off;
    const bool last = rem < 136;
    memset(blk, 0, 136);
    const size_t take = last ? rem : 136;
    for (size_t i = 0; i < take; i++) { const size_t p = off + i; blk[i] = p < n ? in[p] : suffix_byte; }
    if (last) { blk[take] ^= 0x01; blk[135] ^= 0x80; }
    for (int i = 0; i < 17; i++) { uint64_t w; memcpy(&w, blk + 8 * i, 8); a[i] ^= w; }
    keccak_f(a);
    if (last) break;
    off += 136;
  }
  memcpy(out, a, 32);
}

// ---- T256 scalar field helpers on the host: canonical bytes of Montgomery-form elements ---------
static const uint64_t FQ_MOD[4] = {0xffffffffffffffffULL, 0x00000000ffffffffULL, 0x0ULL, 0xffffffff00000001ULL};
static const uint64_t FP_MOD[4] = {0x93135661b1c4b117ULL, 0x7e72b42b30e73177ULL, 0x1ULL, 0xffffffff00000001ULL};
static const uint64_t FQ_INV = 1ULL, FP_INV = 0xe0a2f6a60f646959ULL;   // -p^-1 mod 2^64

// t / 2^256 mod p for t < p * 2^256 given as 4 low limbs (high limbs zero): Montgomery -> canonical
inline void from_mont(const uint64_t a[4], const uint64_t mod[4], uint64_t inv, uint64_t out[4]) {
  uint64_t t[9] = {a[0], a[1], a[2], a[3], 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    const uint64_t m = t[i] * inv;
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)m * mod[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
    for (int j = i + 4; j < 9 && c; j++) { c += t[j]; t[j] = (uint64_t)c; c >>= 64; }
  }
  uint64_t r[4] = {t[4], t[5], t[6], t[7]};
  // one conditional subtraction
  uint64_t d[4]; unsigned borrow = 0;
  for (int i = 0; i < 4; i++) { u128 x = (u128)r[i] - mod[i] - borrow; d[i] = (uint64_t)x; borrow = (unsigned)((x >> 64) & 1); }
  const bool ge = t[8] || !borrow;
  for (int i = 0; i < 4; i++) out[i] = ge ? d[i] : r[i];
}
inline void limbs_to_be(const uint64_t c[4], uint8_t out[32]) {
  for (int i = 0; i < 32; i++) out[31 - i] = (uint8_t)(c[i >> 3] >> (8 * (i & 7)));
}


// ---- host base-field arithmetic (T256 Fp, Montgomery form): only to normalise the handful of Jacobian points
// an MSM batch returns — a serial 256-bit inversion is ~15 us on a host core and ~130 us on a GPU thread ----
inline void mont_mul(const uint64_t a[4], const uint64_t b[4], const uint64_t mod[4], uint64_t inv, uint64_t out[4]) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    const uint64_t m = t[0] * inv;
    c = ((u128)m * mod[0] + t[0]) >> 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * mod[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  uint64_t d[4]; unsigned borrow = 0;
  for (int i = 0; i < 4; i++) { u128 x = (u128)t[i] - mod[i] - borrow; d[i] = (uint64_t)x; borrow = (unsigned)((x >> 64) & 1); }
  const bool ge = t[4] || !borrow;
  for (int i = 0; i < 4; i++) out[i] = ge ? d[i] : t[i];
}
static const uint64_t FP_ONE[4] = {0x6ceca99e4e3b4ee9ULL, 0x818d4bd4cf18ce88ULL, 0xfffffffffffffffeULL, 0x00000000fffffffeULL};   // R mod p
inline void fp_mul(const uint64_t a[4], const uint64_t b[4], uint64_t o[4]) { mont_mul(a, b, FP_MOD, FP_INV, o); }
inline void fp_inv(const uint64_t a[4], uint64_t o[4]) {       // a^(p-2); inv(0) = 0
  uint64_t e[4] = {FP_MOD[0] - 2, FP_MOD[1], FP_MOD[2], FP_MOD[3]};
  uint64_t r[4]; memcpy(r, FP_ONE, 32);
  for (int i = 255; i >= 0; i--) {
    fp_mul(r, r, r);
    if ((e[i >> 6] >> (i & 63)) & 1) fp_mul(r, a, r);
  }
  memcpy(o, r, 32);
}
// n Jacobian points (x,y,z: 12 limbs each) -> affine (x,y: 8 limbs each), identity (z = 0) -> all zero;
// one inversion for the whole batch (Montgomery's trick)
inline void batch_normalize(const uint64_t *jac, size_t n, uint64_t *aff) {
  std::vector<uint64_t> pre(4 * (n + 1));
  uint64_t acc[4]; memcpy(acc, FP_ONE, 32);
  auto is_zero = [](const uint64_t *z) { return !(z[0] | z[1] | z[2] | z[3]); };
  for (size_t i = 0; i < n; i++) {
    memcpy(&pre[4 * i], acc, 32);
    const uint64_t *z = jac + 12 * i + 8;
    if (!is_zero(z)) fp_mul(acc, z, acc);
  }
  uint64_t inv[4]; fp_inv(acc, inv);
  for (size_t i = n; i-- > 0;) {
    const uint64_t *p = jac + 12 * i, *z = p + 8;
    uint64_t *o = aff + 8 * i;
    if (is_zero(z)) { memset(o, 0, 64); continue; }
    uint64_t zi[4], zi2[4], zi3[4];
    fp_mul(inv, &pre[4 * i], zi);
    fp_mul(inv, z, inv);
    fp_mul(zi, zi, zi2); fp_mul(zi2, zi, zi3);
    fp_mul(p, zi2, o); fp_mul(p + 4, zi3, o + 4);
  }
}

// ---- Keccak256Transcript --------------------------------------------------------------------------
struct Transcript {
  uint16_t round = 0;
  uint8_t state[64];
  std::vector<uint8_t> buf;

  // keccak.rs:33-54: K(in || 0x00) || K(in || 0x01).  The two messages share every full block of `in`: absorb
  // those once, then fork the sponge for the two one-byte suffixes.
  static void updated_state(const uint8_t *in, size_t n, uint8_t out[64]) {
    uint64_t a[25]; memset(a, 0, sizeof(a));
    const size_t full = n / 136;
    for (size_t b = 0; b < full; b++) {
      for (int i = 0; i < 17; i++) { uint64_t w; memcpy(&w, in + 136 * b + 8 * i, 8); a[i] ^= w; }
      keccak_f(a);
    }
    const size_t rem = n - 136 * full;              // < 136 bytes left, then the suffix byte, then padding
    for (int sfx = 0; sfx < 2; sfx++) {
      uint64_t s[25]; memcpy(s, a, sizeof(s));
      uint8_t blk[272]; memset(blk, 0, sizeof(blk));
      memcpy(blk, in + 136 * full, rem);
      blk[rem] = (uint8_t)sfx;
      const size_t total = rem + 1, nb = total / 136 + 1;
      blk[total] ^= 0x01; blk[136 * nb - 1] ^= 0x80;
      for (size_t b = 0; b < nb; b++) {
        for (int i = 0; i < 17; i++) { uint64_t w; memcpy(&w, blk + 136 * b + 8 * i, 8); s[i] ^= w; }
        keccak_f(s);
      }
      memcpy(out + 32 * sfx, s, 32);
    }
  }
  explicit Transcript(const char *label) {                                       // keccak.rs:57-68
    std::vector<uint8_t> in; const char *p = "NoTR"; in.insert(in.end(), p, p + 4); in.insert(in.end(), label, label + strlen(label));
    updated_state(in.data(), in.size(), state);
  }
  Transcript(uint16_t rnd, const uint8_t st[64]) : round(rnd) { memcpy(state, st, 64); }
  void push(const void *p, size_t n) { const uint8_t *q = (const uint8_t *)p; buf.insert(buf.end(), q, q + n); }
  void absorb_bytes(const char *label, const void *p, size_t n) { push(label, strlen(label)); push(p, n); }   // keccak.rs:96-99
  void dom_sep(const char *b) { push("NoDS", 4); push(b, strlen(b)); }                                       // keccak.rs:101-104
  void absorb_scalars(const char *label, const uint64_t *mont, size_t n) {       // traits.rs:282-286: 32 bytes big-endian each
    push(label, strlen(label));
    for (size_t i = 0; i < n; i++) { uint64_t c[4]; uint8_t b[32]; from_mont(mont + 4 * i, FQ_MOD, FQ_INV, c); limbs_to_be(c, b); push(b, 32); }
  }
  void push_point(const uint64_t *xy) {                                          // traits.rs:288-305: x_BE || y_BE
    uint64_t c[4]; uint8_t b[32];
    from_mont(xy, FP_MOD, FP_INV, c); limbs_to_be(c, b); push(b, 32);
    from_mont(xy + 4, FP_MOD, FP_INV, c); limbs_to_be(c, b); push(b, 32);
  }
  void absorb_point(const char *label, const uint64_t *xy) { push(label, strlen(label)); push_point(xy); }
  void absorb_commitment(const char *label, const uint64_t *rows_xy, size_t rows) {   // hyrax_pc.rs:714-729
    push(label, strlen(label));
    push("poly_commitment_begin", 21);
    for (size_t i = 0; i < rows; i++) push_point(rows_xy + 8 * i);
    push("poly_commitment_end", 19);
  }
  // same framing from precomputed point bytes (64 per row)
  void absorb_commitment_be(const char *label, const uint8_t *rows_be, size_t rows) {
    push(label, strlen(label));
    push("poly_commitment_begin", 21);
    push(rows_be, 64 * rows);
    push("poly_commitment_end", 19);
  }
  // squeeze (keccak.rs:70-94): 64 uniform bytes; the caller reduces them mod p (on the device)
  void squeeze(const char *label, uint8_t out[64]) {
    const uint8_t le[2] = {(uint8_t)(round & 0xff), (uint8_t)(round >> 8)};
    push("NoDS", 4); push(le, 2); push(state, 64); push(label, strlen(label));
    updated_state(buf.data(), buf.size(), out);
    round++;
    memcpy(state, out, 64);
    buf.clear();
  }
};

}  // namespace sp2h
