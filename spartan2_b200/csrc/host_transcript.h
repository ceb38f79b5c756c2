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
#define SP2H_ROL(v, n) (((v) << (n)) | ((v) >> (64 - (n))))
// Keccak-f[1600], fully unrolled rounds with the 25 lanes in locals (this hashes the tens of KB of commitment rows a
// prove absorbs: ~0.3 us per permutation instead of ~3 us for a table-driven loop)
inline void keccak_f(uint64_t st[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL, 0x000000000000808bULL, 0x0000000080000001ULL,
      0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
      0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL,
      0x000000000000800aULL, 0x800000008000000aULL, 0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  uint64_t a00 = st[0], a01 = st[1], a02 = st[2], a03 = st[3], a04 = st[4], a05 = st[5], a06 = st[6], a07 = st[7], a08 = st[8], a09 = st[9],
           a10 = st[10], a11 = st[11], a12 = st[12], a13 = st[13], a14 = st[14], a15 = st[15], a16 = st[16], a17 = st[17], a18 = st[18],
           a19 = st[19], a20 = st[20], a21 = st[21], a22 = st[22], a23 = st[23], a24 = st[24];
  for (int rnd = 0; rnd < 24; rnd++) {
    // theta
    const uint64_t c0 = a00 ^ a05 ^ a10 ^ a15 ^ a20, c1 = a01 ^ a06 ^ a11 ^ a16 ^ a21, c2 = a02 ^ a07 ^ a12 ^ a17 ^ a22,
                   c3 = a03 ^ a08 ^ a13 ^ a18 ^ a23, c4 = a04 ^ a09 ^ a14 ^ a19 ^ a24;
    const uint64_t d0 = c4 ^ SP2H_ROL(c1, 1), d1 = c0 ^ SP2H_ROL(c2, 1), d2 = c1 ^ SP2H_ROL(c3, 1), d3 = c2 ^ SP2H_ROL(c4, 1), d4 = c3 ^ SP2H_ROL(c0, 1);
    a00 ^= d0; a05 ^= d0; a10 ^= d0; a15 ^= d0; a20 ^= d0;
    a01 ^= d1; a06 ^= d1; a11 ^= d1; a16 ^= d1; a21 ^= d1;
    a02 ^= d2; a07 ^= d2; a12 ^= d2; a17 ^= d2; a22 ^= d2;
    a03 ^= d3; a08 ^= d3; a13 ^= d3; a18 ^= d3; a23 ^= d3;
    a04 ^= d4; a09 ^= d4; a14 ^= d4; a19 ^= d4; a24 ^= d4;
    // rho + pi: b[y + 5*((2x+3y)%5)] = rot(a[x + 5y])
    const uint64_t b00 = a00, b10 = SP2H_ROL(a01, 1), b20 = SP2H_ROL(a02, 62), b05 = SP2H_ROL(a03, 28), b15 = SP2H_ROL(a04, 27),
                   b16 = SP2H_ROL(a05, 36), b01 = SP2H_ROL(a06, 44), b11 = SP2H_ROL(a07, 6), b21 = SP2H_ROL(a08, 55), b06 = SP2H_ROL(a09, 20),
                   b07 = SP2H_ROL(a10, 3), b17 = SP2H_ROL(a11, 10), b02 = SP2H_ROL(a12, 43), b12 = SP2H_ROL(a13, 25), b22 = SP2H_ROL(a14, 39),
                   b23 = SP2H_ROL(a15, 41), b08 = SP2H_ROL(a16, 45), b18 = SP2H_ROL(a17, 15), b03 = SP2H_ROL(a18, 21), b13 = SP2H_ROL(a19, 8),
                   b14 = SP2H_ROL(a20, 18), b24 = SP2H_ROL(a21, 2), b09 = SP2H_ROL(a22, 61), b19 = SP2H_ROL(a23, 56), b04 = SP2H_ROL(a24, 14);
    // chi
    a00 = b00 ^ (~b01 & b02); a01 = b01 ^ (~b02 & b03); a02 = b02 ^ (~b03 & b04); a03 = b03 ^ (~b04 & b00); a04 = b04 ^ (~b00 & b01);
    a05 = b05 ^ (~b06 & b07); a06 = b06 ^ (~b07 & b08); a07 = b07 ^ (~b08 & b09); a08 = b08 ^ (~b09 & b05); a09 = b09 ^ (~b05 & b06);
    a10 = b10 ^ (~b11 & b12); a11 = b11 ^ (~b12 & b13); a12 = b12 ^ (~b13 & b14); a13 = b13 ^ (~b14 & b10); a14 = b14 ^ (~b10 & b11);
    a15 = b15 ^ (~b16 & b17); a16 = b16 ^ (~b17 & b18); a17 = b17 ^ (~b18 & b19); a18 = b18 ^ (~b19 & b15); a19 = b19 ^ (~b15 & b16);
    a20 = b20 ^ (~b21 & b22); a21 = b21 ^ (~b22 & b23); a22 = b22 ^ (~b23 & b24); a23 = b23 ^ (~b24 & b20); a24 = b24 ^ (~b20 & b21);
    a00 ^= RC[rnd];
  }
  st[0] = a00; st[1] = a01; st[2] = a02; st[3] = a03; st[4] = a04; st[5] = a05; st[6] = a06; st[7] = a07; st[8] = a08; st[9] = a09;
  st[10] = a10; st[11] = a11; st[12] = a12; st[13] = a13; st[14] = a14; st[15] = a15; st[16] = a16; st[17] = a17; st[18] = a18; st[19] = a19;
  st[20] = a20; st[21] = a21; st[22] = a22; st[23] = a23; st[24] = a24;
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
inline void fp_inv_fermat(const uint64_t a[4], uint64_t o[4]) {       // a^(p-2); inv(0) = 0 (cross-check of fp_inv)
  uint64_t e[4] = {FP_MOD[0] - 2, FP_MOD[1], FP_MOD[2], FP_MOD[3]};
  uint64_t r[4]; memcpy(r, FP_ONE, 32);
  for (int i = 255; i >= 0; i--) {
    fp_mul(r, r, r);
    if ((e[i >> 6] >> (i & 63)) & 1) fp_mul(r, a, r);
  }
  memcpy(o, r, 32);
}
// ---- 256-bit helpers for the binary extended Euclid below ----
inline bool u256_is_one(const uint64_t a[4]) { return a[0] == 1 && !(a[1] | a[2] | a[3]); }
inline bool u256_ge(const uint64_t a[4], const uint64_t b[4]) {
  for (int i = 3; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
  return true;
}
inline void u256_sub(uint64_t a[4], const uint64_t b[4]) {                      // a -= b (a >= b)
  unsigned borrow = 0;
  for (int i = 0; i < 4; i++) { u128 x = (u128)a[i] - b[i] - borrow; a[i] = (uint64_t)x; borrow = (unsigned)((x >> 64) & 1); }
}
inline void u256_shr1(uint64_t a[4], uint64_t top) {                            // (top:a) >> 1
  a[0] = (a[0] >> 1) | (a[1] << 63); a[1] = (a[1] >> 1) | (a[2] << 63); a[2] = (a[2] >> 1) | (a[3] << 63); a[3] = (a[3] >> 1) | (top << 63);
}
inline void mod_half(uint64_t x[4], const uint64_t mod[4]) {                    // x / 2 mod p, x < p
  uint64_t top = 0;
  if (x[0] & 1) { unsigned carry = 0; for (int i = 0; i < 4; i++) { u128 s = (u128)x[i] + mod[i] + carry; x[i] = (uint64_t)s; carry = (unsigned)(s >> 64); } top = carry; }
  u256_shr1(x, top);
}
inline void mod_sub(uint64_t x[4], const uint64_t y[4], const uint64_t mod[4]) { // x = x - y mod p, both < p
  unsigned borrow = 0;
  for (int i = 0; i < 4; i++) { u128 d = (u128)x[i] - y[i] - borrow; x[i] = (uint64_t)d; borrow = (unsigned)((d >> 64) & 1); }
  if (borrow) { unsigned carry = 0; for (int i = 0; i < 4; i++) { u128 s = (u128)x[i] + mod[i] + carry; x[i] = (uint64_t)s; carry = (unsigned)(s >> 64); } }
}
// Montgomery inverse a R -> a^-1 R by the binary extended Euclid on plain integers (~4 us on a host core; the Fermat ladder is
// 384 Montgomery multiplications, ~15 us — and this inversion sits on the prove's critical path twice: the rest-commitment
// rows before the taus, the PCS points before the IPA challenge).  x = (aR)^-1 as an integer, then x * R^3 / R = a^-1 R.  inv(0) = 0.
inline void pow2_mod(int k, const uint64_t mod[4], uint64_t out[4]) {         // 2^k mod p by doubling
  uint64_t r[4] = {1, 0, 0, 0};
  for (int i = 0; i < k; i++) {
    uint64_t d[4]; unsigned carry = 0;
    for (int j = 0; j < 4; j++) { const uint64_t v = r[j]; d[j] = (v << 1) | carry; carry = (unsigned)(v >> 63); }
    uint64_t e[4]; unsigned borrow = 0;
    for (int j = 0; j < 4; j++) { u128 x = (u128)d[j] - mod[j] - borrow; e[j] = (uint64_t)x; borrow = (unsigned)((x >> 64) & 1); }
    const bool ge = carry || !borrow;
    for (int j = 0; j < 4; j++) r[j] = ge ? e[j] : d[j];
  }
  memcpy(out, r, 32);
}
inline void mont_inv(const uint64_t a[4], const uint64_t mod[4], uint64_t minv, const uint64_t R3[4], uint64_t o[4]) {
  if (!(a[0] | a[1] | a[2] | a[3])) { memset(o, 0, 32); return; }
  uint64_t u[4], v[4], x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
  memcpy(u, a, 32); memcpy(v, mod, 32);
  while (!u256_is_one(u) && !u256_is_one(v)) {
    while (!(u[0] & 1)) { u256_shr1(u, 0); mod_half(x1, mod); }
    while (!(v[0] & 1)) { u256_shr1(v, 0); mod_half(x2, mod); }
    if (u256_ge(u, v)) { u256_sub(u, v); mod_sub(x1, x2, mod); }
    else { u256_sub(v, u); mod_sub(x2, x1, mod); }
  }
  mont_mul(u256_is_one(u) ? x1 : x2, R3, mod, minv, o);
}
inline void fp_inv(const uint64_t a[4], uint64_t o[4]) {
  struct Consts { uint64_t R3[4]; Consts() { pow2_mod(768, FP_MOD, R3); } };   // (function-local static: thread-safe one-time initialisation)
  static const Consts K;
  mont_inv(a, FP_MOD, FP_INV, K.R3, o);
}
// the same for the scalar field (the prover's tau inverses: sumcheck.cu derives t(1) from the running claim, sumcheck.rs:1277-1324)
inline void fq_inv(const uint64_t a[4], uint64_t o[4]) {
  struct Consts { uint64_t R3[4]; Consts() { pow2_mod(768, FQ_MOD, R3); } };
  static const Consts K;
  mont_inv(a, FQ_MOD, FQ_INV, K.R3, o);
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

// ---- streaming Keccak-256 sponge: data is permuted as it is pushed, so bulk absorbs (commitment rows) are hashed
// while the GPU is busy; squeeze only finishes the last block(s) ------------------------------------------------
struct Sponge {
  uint64_t a[25]; uint8_t blk[136]; size_t fill = 0;
  Sponge() { memset(a, 0, sizeof(a)); }
  void absorb(const void *p, size_t n) {
    const uint8_t *q = (const uint8_t *)p;
    while (n) {
      const size_t take = n < 136 - fill ? n : 136 - fill;
      memcpy(blk + fill, q, take); fill += take; q += take; n -= take;
      if (fill == 136) { for (int i = 0; i < 17; i++) { uint64_t w; memcpy(&w, blk + 8 * i, 8); a[i] ^= w; } keccak_f(a); fill = 0; }
    }
  }
  // digest of (everything absorbed || suffix byte), without disturbing this sponge
  void finish_with_suffix(uint8_t sfx, uint8_t out[32]) const {
    Sponge s = *this;
    s.absorb(&sfx, 1);
    memset(s.blk + s.fill, 0, 136 - s.fill);
    s.blk[s.fill] ^= 0x01; s.blk[135] ^= 0x80;
    for (int i = 0; i < 17; i++) { uint64_t w; memcpy(&w, s.blk + 8 * i, 8); s.a[i] ^= w; }
    keccak_f(s.a);
    memcpy(out, s.a, 32);
  }
};
inline void point_be(const uint64_t *xy, uint8_t out[64]) {   // traits.rs:288-305: x_BE || y_BE
  uint64_t c[4];
  from_mont(xy, FP_MOD, FP_INV, c); limbs_to_be(c, out);
  from_mont(xy + 4, FP_MOD, FP_INV, c); limbs_to_be(c, out + 32);
}

// ---- Keccak256Transcript --------------------------------------------------------------------------
struct Transcript {
  uint16_t round = 0;
  uint8_t state[64];
  Sponge sp;                       // absorbs since the last squeeze (keccak.rs keeps them in a buffer and hashes at squeeze)

  // keccak.rs:33-54: K(in || 0x00) || K(in || 0x01)
  static void updated_state(const uint8_t *in, size_t n, uint8_t out[64]) {
    Sponge s; s.absorb(in, n);
    s.finish_with_suffix(0x00, out); s.finish_with_suffix(0x01, out + 32);
  }
  explicit Transcript(const char *label) {                                       // keccak.rs:57-68
    std::vector<uint8_t> in; const char *p = "NoTR"; in.insert(in.end(), p, p + 4); in.insert(in.end(), label, label + strlen(label));
    updated_state(in.data(), in.size(), state);
  }
  // a transcript whose (round, state) are not known yet (they are spliced in at squeeze time, after the absorbed data)
  Transcript() { memset(state, 0, 64); }
  Transcript(uint16_t rnd, const uint8_t st[64]) : round(rnd) { memcpy(state, st, 64); }
  void set_state(uint16_t rnd, const uint8_t st[64]) { round = rnd; memcpy(state, st, 64); }
  void push(const void *p, size_t n) { sp.absorb(p, n); }
  void absorb_bytes(const char *label, const void *p, size_t n) { push(label, strlen(label)); push(p, n); }   // keccak.rs:96-99
  void dom_sep(const char *b) { push("NoDS", 4); push(b, strlen(b)); }                                       // keccak.rs:101-104
  void absorb_scalars(const char *label, const uint64_t *mont, size_t n) {       // traits.rs:282-286: 32 bytes big-endian each
    push(label, strlen(label));
    for (size_t i = 0; i < n; i++) { uint64_t c[4]; uint8_t b[32]; from_mont(mont + 4 * i, FQ_MOD, FQ_INV, c); limbs_to_be(c, b); push(b, 32); }
  }
  void push_point(const uint64_t *xy) { uint8_t b[64]; point_be(xy, b); push(b, 64); }
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
    sp.finish_with_suffix(0x00, out); sp.finish_with_suffix(0x01, out + 32);
    round++;
    memcpy(state, out, 64);
    sp = Sponge();
  }
};

// ---- challenge = from_uniform(64 squeezed bytes): 512-bit little-endian integer mod q, Montgomery form
// (provider/traits.rs:275-280; halo2curves from_uniform_bytes) ---------------------------------------------------
inline void fq_pow2_mont(int k, uint64_t out[4]) {              // 2^k mod q (canonical) by doubling
  uint64_t r[4] = {1, 0, 0, 0};
  for (int i = 0; i < k; i++) {
    uint64_t d[4]; unsigned carry = 0;
    for (int j = 0; j < 4; j++) { const uint64_t v = r[j]; d[j] = (v << 1) | carry; carry = (unsigned)(v >> 63); }
    uint64_t e[4]; unsigned borrow = 0;
    for (int j = 0; j < 4; j++) { u128 x = (u128)d[j] - FQ_MOD[j] - borrow; e[j] = (uint64_t)x; borrow = (unsigned)((x >> 64) & 1); }
    const bool ge = carry || !borrow;
    for (int j = 0; j < 4; j++) r[j] = ge ? e[j] : d[j];
  }
  memcpy(out, r, 32);
}
inline void fq_from_uniform(const uint8_t dg[64], uint64_t out_mont[4]) {
  // function-local static with a dynamic initialiser: C++11 guarantees a thread-safe one-time initialisation (contexts on
  // several host threads squeeze concurrently)
  struct Consts { uint64_t R2[4], R3[4]; Consts() { fq_pow2_mont(512, R2); fq_pow2_mont(768, R3); } };
  static const Consts K;
  const uint64_t *R2 = K.R2, *R3 = K.R3;
  uint64_t lo[4], hi[4], a[4], b[4];
  memcpy(lo, dg, 32); memcpy(hi, dg + 32, 32);
  mont_mul(lo, R2, FQ_MOD, FQ_INV, a);                           // lo * R
  mont_mul(hi, R3, FQ_MOD, FQ_INV, b);                           // hi * 2^256 * R
  unsigned carry = 0; uint64_t s4[4];
  for (int j = 0; j < 4; j++) { u128 x = (u128)a[j] + b[j] + carry; s4[j] = (uint64_t)x; carry = (unsigned)(x >> 64); }
  uint64_t e[4]; unsigned borrow = 0;
  for (int j = 0; j < 4; j++) { u128 x = (u128)s4[j] - FQ_MOD[j] - borrow; e[j] = (uint64_t)x; borrow = (unsigned)((x >> 64) & 1); }
  const bool ge = carry || !borrow;
  for (int j = 0; j < 4; j++) out_mont[j] = ge ? e[j] : s4[j];
}

}  // namespace sp2h

// the C-ABI handle of the host transcript (include/spartan2_b200.h: sp2_transcript_*)
struct sp2_transcript { sp2h::Transcript t; explicit sp2_transcript(const char *label) : t(label) {} };
