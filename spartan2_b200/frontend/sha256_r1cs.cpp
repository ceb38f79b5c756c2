// SHA-256 -> R1CS generator: the benchmark workload of microsoft/Spartan2 (benches/sha256_spartan.rs:37-152
// Sha256Circuit; benches/sha256_neutronnova.rs:52-128 Sha256StepCircuit) without the bellpepper frontend.
//
// The reference builds these circuits with the third-party `bellpepper` gadget library (Cargo.toml:14-15,
// un-vendored; src/bellpepper/ is OUT OF SCOPE for the hot path).  To have the same
// workload on a box without Rust, the gadget's published structure is restated here from scratch: Boolean
// {Constant, Is, Not} with constant propagation, AllocatedBit xor/and/and_not/nor (1 constraint each),
// sha256_ch (1 constraint/bit), sha256_maj (2 constraints/bit), UInt32 rotr/shr (free), addmany with deferred
// additions (result bits + ONE long linear row).  Checkpoint: one compression with the constant IV costs 25,840
// constraints beyond its 512 input-bit constraints — the count quoted at benches/sha256_neutronnova.rs:159-160
// (26,352 = 25,840 + 512); tests/test_sha256_r1cs.py asserts it, and that the witness satisfies every
// constraint with the digest hashlib computes.  Variable/constraint ORDER inside the gadget is not pinned by
// anything in the reference tree ("parity unpinned"); shapes and statistics are.
//
// Output: the three matrices in the padded layout SplitR1CSShape::new produces (src/r1cs/mod.rs:810-911):
// columns [W_precommitted (padded to the commitment width) | W_rest (zero padding up to a power of two) | 1 |
// public], rows padded to a power of two; coefficients dictionary-coded as exact integers (MultiEq rows reach 2^255).
#include <stdint.h>
#include <string.h>
#include <algorithm>
#include <vector>

namespace {

typedef int32_t Var;                       // >= 0: aux index; -1: ONE; <= -2: public input (-2 - k)
const Var ONE = -1;
struct Term { Var v; int64_t c; uint16_t sh; };     // coefficient c * 2^sh
typedef std::vector<Term> LC;

struct Bool { int kind; Var v; bool val; };   // kind 0 = Constant(val), 1 = Is(v), 2 = Not(v); val = current value
Bool bconst(bool b) { return Bool{0, 0, b}; }
Bool bnot(const Bool &b) { if (b.kind == 0) return bconst(!b.val); return Bool{b.kind == 1 ? 2 : 1, b.v, !b.val}; }

// signed 384-bit integers: exact coefficient and row-value arithmetic (MultiEq packs ~7 additions of 35 bits each
// into one constraint, so coefficients reach 2^255)
struct Big {
  uint64_t l[6];
  Big() { memset(l, 0, sizeof(l)); }
  static Big from_shifted(int64_t c, unsigned sh) {
    Big r; const uint64_t ext = c < 0 ? ~0ull : 0ull;
    for (int i = 0; i < 6; i++) r.l[i] = ext;
    r.l[0] = (uint64_t)c;
    for (unsigned k = 0; k < sh / 64; k++) { for (int i = 5; i > 0; i--) r.l[i] = r.l[i - 1]; r.l[0] = 0; }
    const unsigned b = sh % 64;
    if (b) { for (int i = 5; i > 0; i--) r.l[i] = (r.l[i] << b) | (r.l[i - 1] >> (64 - b)); r.l[0] <<= b; }
    return r;
  }
  void add(const Big &o) { unsigned __int128 c = 0; for (int i = 0; i < 6; i++) { c += (unsigned __int128)l[i] + o.l[i]; l[i] = (uint64_t)c; c >>= 64; } }
  bool is_zero() const { for (int i = 0; i < 6; i++) if (l[i]) return false; return true; }
  bool neg() const { return l[5] >> 63; }
  Big negated() const { Big r; unsigned __int128 c = 1; for (int i = 0; i < 6; i++) { c += (uint64_t)~l[i]; r.l[i] = (uint64_t)c; c >>= 64; } return r; }
  bool fits_i64() const { const uint64_t ext = (l[0] >> 63) ? ~0ull : 0ull; for (int i = 1; i < 6; i++) if (l[i] != ext) return false; return true; }
  Big times_i64(int64_t m) const {      // exact as long as the result fits 384 bits
    Big a = neg() ? negated() : *this; const bool sgn = neg() != (m < 0);
    const uint64_t mm = m < 0 ? (uint64_t)(-(m + 1)) + 1 : (uint64_t)m;
    Big r; unsigned __int128 c = 0;
    for (int i = 0; i < 6; i++) { c += (unsigned __int128)a.l[i] * mm; r.l[i] = (uint64_t)c; c >>= 64; }
    return sgn ? r.negated() : r;
  }
  bool operator==(const Big &o) const { return !memcmp(l, o.l, sizeof(l)); }
  bool operator<(const Big &o) const { return memcmp(l, o.l, sizeof(l)) < 0; }
};

struct CS {
  // bellpepper's MultiEq (gadgets/multieq.rs): equalities of `num_bits`-bit quantities are packed, shifted by the
  // bits already used, into one constraint lhs * 1 = rhs until the field capacity (255 bits here) would be exceeded
  LC me_lhs, me_rhs; unsigned me_bits = 0;
  void me_flush() { if (me_bits) { LC one; one.push_back(Term{-1, 1, 0}); enforce(me_lhs, one, me_rhs); me_lhs.clear(); me_rhs.clear(); me_bits = 0; } }
  void enforce_equal(unsigned num_bits, const LC &lhs, const LC &rhs) {
    const unsigned CAPACITY = 255;                    // T256 scalar field: NUM_BITS - 1
    if (CAPACITY <= me_bits + num_bits) me_flush();
    for (const Term &t : lhs) me_lhs.push_back(Term{t.v, t.c, (uint16_t)(t.sh + me_bits)});
    for (const Term &t : rhs) me_rhs.push_back(Term{t.v, t.c, (uint16_t)(t.sh + me_bits)});
    me_bits += num_bits;
  }
  std::vector<uint8_t> aux;                // every auxiliary variable of these circuits is a bit
  std::vector<uint8_t> inputs;             // public inputs (0/1)
  std::vector<LC> A, B, C;
  Var alloc(bool v) { aux.push_back(v ? 1 : 0); return (Var)aux.size() - 1; }
  Var alloc_input(bool v) { inputs.push_back(v ? 1 : 0); return (Var)(-2 - ((Var)inputs.size() - 1)); }
  void enforce(const LC &a, const LC &b, const LC &c) { A.push_back(a); B.push_back(b); C.push_back(c); }
};

void lc_add(LC &lc, Var v, int64_t c) { if (c) lc.push_back(Term{v, c, 0}); }
// Boolean::lc(one, coeff)
void lc_add_bool(LC &lc, const Bool &b, int64_t coeff) {
  if (b.kind == 0) { if (b.val) lc_add(lc, ONE, coeff); }
  else if (b.kind == 1) lc_add(lc, b.v, coeff);
  else { lc_add(lc, ONE, coeff); lc_add(lc, b.v, -coeff); }
}
LC lc_of_bool(const Bool &b, int64_t coeff) { LC l; lc_add_bool(l, b, coeff); return l; }
LC lc_var(Var v, int64_t c = 1) { LC l; lc_add(l, v, c); return l; }

// AllocatedBit::alloc: (1 - a) * a = 0
Bool alloc_bit(CS &cs, bool v) {
  Var a = cs.alloc(v);
  LC l1; lc_add(l1, ONE, 1); lc_add(l1, a, -1);
  cs.enforce(l1, lc_var(a), LC());
  return Bool{1, a, v};
}
// AllocatedBit::{xor, and, and_not, nor}: one constraint each, result is a raw allocation
Var ab_xor(CS &cs, Var a, bool av, Var b, bool bv) {
  Var c = cs.alloc(av ^ bv);
  LC l; lc_add(l, a, 1); lc_add(l, b, 1); lc_add(l, c, -1);
  cs.enforce(lc_var(a, 2), lc_var(b), l);                       // (a + a) * b = a + b - c
  return c;
}
Var ab_and(CS &cs, Var a, bool av, Var b, bool bv) { Var c = cs.alloc(av && bv); cs.enforce(lc_var(a), lc_var(b), lc_var(c)); return c; }
Var ab_and_not(CS &cs, Var a, bool av, Var b, bool bv) {      // a AND (NOT b)
  Var c = cs.alloc(av && !bv);
  LC nb; lc_add(nb, ONE, 1); lc_add(nb, b, -1);
  cs.enforce(lc_var(a), nb, lc_var(c));
  return c;
}
Var ab_nor(CS &cs, Var a, bool av, Var b, bool bv) {          // (NOT a) AND (NOT b)
  Var c = cs.alloc(!av && !bv);
  LC na, nb; lc_add(na, ONE, 1); lc_add(na, a, -1); lc_add(nb, ONE, 1); lc_add(nb, b, -1);
  cs.enforce(na, nb, lc_var(c));
  return c;
}
bool raw(const Bool &b) { return b.kind == 2 ? !b.val : b.val; }   // value of the underlying allocated bit

Bool b_xor(CS &cs, const Bool &a, const Bool &b) {
  if (a.kind == 0) return a.val ? bnot(b) : b;
  if (b.kind == 0) return b.val ? bnot(a) : a;
  Var c = ab_xor(cs, a.v, raw(a), b.v, raw(b));
  const bool cv = raw(a) ^ raw(b);
  if (a.kind != b.kind) return Bool{2, c, !cv};                   // Is ^ Not = Not(xor)
  return Bool{1, c, cv};
}
Bool b_and(CS &cs, const Bool &a, const Bool &b) {
  if (a.kind == 0) return a.val ? b : bconst(false);
  if (b.kind == 0) return b.val ? a : bconst(false);
  if (a.kind == 1 && b.kind == 1) { Var c = ab_and(cs, a.v, a.val, b.v, b.val); return Bool{1, c, a.val && b.val}; }
  if (a.kind == 1 && b.kind == 2) { Var c = ab_and_not(cs, a.v, a.val, b.v, raw(b)); return Bool{1, c, a.val && b.val}; }
  if (a.kind == 2 && b.kind == 1) { Var c = ab_and_not(cs, b.v, b.val, a.v, raw(a)); return Bool{1, c, a.val && b.val}; }
  Var c = ab_nor(cs, a.v, raw(a), b.v, raw(b));
  return Bool{1, c, a.val && b.val};
}
// Boolean::sha256_ch: (a and b) xor ((not a) and c)
Bool b_ch(CS &cs, const Bool &a, const Bool &b, const Bool &c) {
  const bool v = (a.val && b.val) ^ (!a.val && c.val);
  if (a.kind == 0 && b.kind == 0 && c.kind == 0) return bconst(v);
  if (a.kind == 0) return a.val ? b : c;
  if (b.kind == 0 && !b.val) return b_and(cs, bnot(a), c);
  if (c.kind == 0 && !c.val) return b_and(cs, a, b);
  if (c.kind == 0 && c.val) return bnot(b_and(cs, a, bnot(b)));
  if (b.kind == 0 && b.val) return bnot(b_and(cs, bnot(a), bnot(c)));
  Var ch = cs.alloc(v);                                           // a (b - c) = ch - c
  LC l1 = lc_of_bool(b, 1); lc_add_bool(l1, c, -1);
  LC l3 = lc_var(ch); lc_add_bool(l3, c, -1);
  cs.enforce(l1, lc_of_bool(a, 1), l3);
  return Bool{1, ch, v};
}
// Boolean::sha256_maj: (a and b) xor (a and c) xor (b and c)
Bool b_maj(CS &cs, const Bool &a, const Bool &b, const Bool &c) {
  const bool v = (a.val && b.val) ^ (a.val && c.val) ^ (b.val && c.val);
  if (a.kind == 0 && b.kind == 0 && c.kind == 0) return bconst(v);
  if (a.kind == 0 && !a.val) return b_and(cs, b, c);
  if (b.kind == 0 && !b.val) return b_and(cs, a, c);
  if (c.kind == 0 && !c.val) return b_and(cs, a, b);
  if (c.kind == 0 && c.val) return bnot(b_and(cs, bnot(a), bnot(b)));
  if (b.kind == 0 && b.val) return bnot(b_and(cs, bnot(a), bnot(c)));
  if (a.kind == 0 && a.val) return bnot(b_and(cs, bnot(b), bnot(c)));
  const Bool bc = b_and(cs, b, c);
  Var maj = cs.alloc(v);                                          // (2bc - b - c) * a = bc - maj
  LC l1 = lc_of_bool(bc, 2); lc_add_bool(l1, b, -1); lc_add_bool(l1, c, -1);
  LC l3 = lc_of_bool(bc, 1); lc_add(l3, maj, -1);
  cs.enforce(l1, lc_of_bool(a, 1), l3);
  return Bool{1, maj, v};
}

struct U32 { Bool bits[32]; };             // little-endian bits
U32 u_const(uint32_t v) { U32 r; for (int i = 0; i < 32; i++) r.bits[i] = bconst((v >> i) & 1); return r; }
U32 u_from_bits_be(const Bool *b) { U32 r; for (int i = 0; i < 32; i++) r.bits[i] = b[31 - i]; return r; }
U32 u_rotr(const U32 &a, int by) { U32 r; for (int i = 0; i < 32; i++) r.bits[i] = a.bits[(i + by) % 32]; return r; }
U32 u_shr(const U32 &a, int by) { U32 r; for (int i = 0; i < 32; i++) r.bits[i] = i + by < 32 ? a.bits[i + by] : bconst(false); return r; }
U32 u_xor(CS &cs, const U32 &a, const U32 &b) { U32 r; for (int i = 0; i < 32; i++) r.bits[i] = b_xor(cs, a.bits[i], b.bits[i]); return r; }
U32 u_ch(CS &cs, const U32 &a, const U32 &b, const U32 &c) { U32 r; for (int i = 0; i < 32; i++) r.bits[i] = b_ch(cs, a.bits[i], b.bits[i], c.bits[i]); return r; }
U32 u_maj(CS &cs, const U32 &a, const U32 &b, const U32 &c) { U32 r; for (int i = 0; i < 32; i++) r.bits[i] = b_maj(cs, a.bits[i], b.bits[i], c.bits[i]); return r; }
uint32_t u_value(const U32 &a) { uint32_t v = 0; for (int i = 0; i < 32; i++) v |= (uint32_t)a.bits[i].val << i; return v; }

// UInt32::addmany: result bits (booleanity each) + one linear "modular addition" row
U32 u_addmany(CS &cs, const std::vector<U32> &ops) {
  uint64_t max_value = (uint64_t)ops.size() * 0xffffffffull, sum = 0;
  bool all_const = true;
  LC lc;
  for (const U32 &op : ops) {
    sum += u_value(op);
    for (int i = 0; i < 32; i++) { lc_add_bool(lc, op.bits[i], (int64_t)1 << i); all_const &= op.bits[i].kind == 0; }
  }
  if (all_const) return u_const((uint32_t)sum);
  U32 r; LC rlc; int i = 0;
  while (max_value) {
    const Bool b = alloc_bit(cs, (sum >> i) & 1);
    lc_add(rlc, b.v, (int64_t)1 << i);
    if (i < 32) r.bits[i] = b;
    max_value >>= 1; i++;
  }
  cs.enforce_equal((unsigned)i, lc, rlc);
  return r;
}

const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3,
    0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
const uint32_t IV256[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};

struct Maybe { bool deferred; std::vector<U32> ops; };           // Deferred(ops) | Concrete(ops[0])
U32 maybe_compute(CS &cs, Maybe m, const std::vector<U32> &others) {
  if (!m.deferred && others.empty()) return m.ops[0];
  std::vector<U32> v = m.ops; v.insert(v.end(), others.begin(), others.end());
  return u_addmany(cs, v);
}

void compression(CS &cs, const Bool *input512, U32 cur[8]) {
  std::vector<U32> w;
  for (int i = 0; i < 16; i++) w.push_back(u_from_bits_be(input512 + 32 * i));
  for (int i = 16; i < 64; i++) {
    U32 s0 = u_rotr(w[i - 15], 7); s0 = u_xor(cs, s0, u_rotr(w[i - 15], 18)); s0 = u_xor(cs, s0, u_shr(w[i - 15], 3));
    U32 s1 = u_rotr(w[i - 2], 17); s1 = u_xor(cs, s1, u_rotr(w[i - 2], 19)); s1 = u_xor(cs, s1, u_shr(w[i - 2], 10));
    w.push_back(u_addmany(cs, {w[i - 16], s0, w[i - 7], s1}));
  }
  Maybe a{false, {cur[0]}}, e{false, {cur[4]}};
  U32 b = cur[1], c = cur[2], d = cur[3], f = cur[5], g = cur[6], h = cur[7];
  for (int i = 0; i < 64; i++) {
    const U32 new_e = maybe_compute(cs, e, {});
    U32 s1 = u_rotr(new_e, 6); s1 = u_xor(cs, s1, u_rotr(new_e, 11)); s1 = u_xor(cs, s1, u_rotr(new_e, 25));
    const U32 ch = u_ch(cs, new_e, f, g);
    const std::vector<U32> temp1 = {h, s1, ch, u_const(K256[i]), w[i]};
    const U32 new_a = maybe_compute(cs, a, {});
    U32 s0 = u_rotr(new_a, 2); s0 = u_xor(cs, s0, u_rotr(new_a, 13)); s0 = u_xor(cs, s0, u_rotr(new_a, 22));
    const U32 maj = u_maj(cs, new_a, b, c);
    h = g; g = f; f = new_e;
    e.deferred = true; e.ops = temp1; e.ops.push_back(d);
    d = c; c = b; b = new_a;
    a.deferred = true; a.ops = temp1; a.ops.push_back(s0); a.ops.push_back(maj);
  }
  U32 out[8];
  out[0] = maybe_compute(cs, a, {cur[0]});
  out[1] = u_addmany(cs, {cur[1], b});
  out[2] = u_addmany(cs, {cur[2], c});
  out[3] = u_addmany(cs, {cur[3], d});
  out[4] = maybe_compute(cs, e, {cur[4]});
  out[5] = u_addmany(cs, {cur[5], f});
  out[6] = u_addmany(cs, {cur[6], g});
  out[7] = u_addmany(cs, {cur[7], h});
  for (int i = 0; i < 8; i++) cur[i] = out[i];
  cs.me_flush();                                                  // MultiEq is dropped at the end of sha256_compression_function
}

// bellpepper gadgets::sha256::sha256: pad, iterate the compression function, big-endian output bits
std::vector<Bool> sha256_gadget(CS &cs, const std::vector<Bool> &input) {
  std::vector<Bool> padded = input;
  const uint64_t plen = padded.size();
  padded.push_back(bconst(true));
  while ((padded.size() + 64) % 512) padded.push_back(bconst(false));
  for (int i = 63; i >= 0; i--) padded.push_back(bconst((plen >> i) & 1));
  U32 cur[8];
  for (int i = 0; i < 8; i++) cur[i] = u_const(IV256[i]);
  for (size_t blk = 0; blk < padded.size() / 512; blk++) compression(cs, padded.data() + 512 * blk, cur);
  std::vector<Bool> out;
  for (int i = 0; i < 8; i++) for (int j = 31; j >= 0; j--) out.push_back(cur[i].bits[j]);
  return out;
}

struct Built {
  CS cs;
  uint64_t num_cons = 0, num_aux = 0, num_public = 0, width = 2048;
  uint64_t num_pre_padded = 0, num_rest_padded = 0, num_vars_padded = 0, num_cons_padded = 0;
  std::vector<uint32_t> cid[3], idx[3], ptr[3];      // per nnz: coefficient id (into dict), column; row pointers
  std::vector<Big> dict;                              // distinct coefficient values (exact integers)
  uint8_t digest_bits[256];
};

uint64_t next_pow2(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return p; }

void finalize(Built &B) {
  CS &cs = B.cs;
  B.num_cons = cs.A.size(); B.num_aux = cs.aux.size(); B.num_public = cs.inputs.size();
  const uint64_t w = B.width;
  B.num_pre_padded = (B.num_aux + w - 1) / w * w;                 // everything is allocated in `precommitted`
  uint64_t rest = 0, vars = B.num_pre_padded;
  if (vars < B.num_public + 1) { rest = B.num_public + 1 - vars; vars += rest; }
  if (next_pow2(vars) != vars) { rest = next_pow2(vars) - B.num_pre_padded; vars = B.num_pre_padded + rest; }
  B.num_rest_padded = rest; B.num_vars_padded = vars; B.num_cons_padded = next_pow2(B.num_cons);
  const std::vector<LC> *M[3] = {&cs.A, &cs.B, &cs.C};
  std::vector<std::pair<Big, uint32_t>> dict_sorted;              // built lazily: small values are by far the most common
  std::vector<std::pair<int64_t, uint32_t>> small_ids;            // cache for coefficients that fit an int64
  auto dict_id = [&](const Big &v) -> uint32_t {
    if (v.fits_i64()) {
      const int64_t k = (int64_t)v.l[0];
      for (auto &e : small_ids) if (e.first == k) return e.second;
      if (small_ids.size() < 64) { B.dict.push_back(v); small_ids.push_back({k, (uint32_t)B.dict.size() - 1}); return small_ids.back().second; }
    }
    auto it = std::lower_bound(dict_sorted.begin(), dict_sorted.end(), std::make_pair(v, (uint32_t)0),
                               [](const std::pair<Big, uint32_t> &x, const std::pair<Big, uint32_t> &y) { return x.first < y.first; });
    if (it != dict_sorted.end() && it->first == v) return it->second;
    B.dict.push_back(v);
    dict_sorted.insert(it, {v, (uint32_t)B.dict.size() - 1});
    return (uint32_t)B.dict.size() - 1;
  };
  for (int k = 0; k < 3; k++) {
    B.ptr[k].assign(1, 0);
    std::vector<std::pair<uint32_t, Big>> row;
    for (const LC &lc : *M[k]) {
      row.clear();
      for (const Term &t : lc) {
        const uint32_t col = t.v >= 0 ? (uint32_t)t.v : (uint32_t)(B.num_vars_padded + (uint64_t)(-1 - t.v));   // ONE -> vars, input k -> vars+1+k
        row.push_back({col, Big::from_shifted(t.c, t.sh)});
      }
      std::stable_sort(row.begin(), row.end(), [](const std::pair<uint32_t, Big> &x, const std::pair<uint32_t, Big> &y) { return x.first < y.first; });
      for (size_t i = 0; i < row.size();) {                       // merge duplicate variables, drop zeros
        size_t j = i; Big c;
        while (j < row.size() && row[j].first == row[i].first) c.add(row[j++].second);
        if (!c.is_zero()) { B.idx[k].push_back(row[i].first); B.cid[k].push_back(dict_id(c)); }
        i = j;
      }
      B.ptr[k].push_back((uint32_t)B.idx[k].size());
    }
    for (uint64_t r = B.num_cons; r < B.num_cons_padded; r++) B.ptr[k].push_back((uint32_t)B.idx[k].size());
  }
}

// exact integer check of Az o Bz = Cz for the given assignment (all variables are bits)
bool satisfied(const Built &B, const uint8_t *aux, const uint8_t *pub) {
  auto zval = [&](uint32_t col) -> int {
    if (col < B.num_vars_padded) return col < B.num_aux ? aux[col] : 0;
    if (col == B.num_vars_padded) return 1;
    return pub[col - B.num_vars_padded - 1];
  };
  for (uint64_t r = 0; r < B.num_cons; r++) {
    Big v[3];
    for (int k = 0; k < 3; k++)
      for (uint32_t e = B.ptr[k][r]; e < B.ptr[k][r + 1]; e++) if (zval(B.idx[k][e])) v[k].add(B.dict[B.cid[k][e]]);
    Big prod;
    if (v[1].fits_i64()) prod = v[0].times_i64((int64_t)v[1].l[0]);
    else if (v[0].fits_i64()) prod = v[1].times_i64((int64_t)v[0].l[0]);
    else return false;
    if (!(prod == v[2])) return false;
  }
  return true;
}

std::vector<Bool> alloc_bytes_be(CS &cs, const uint8_t *p, size_t n) {
  std::vector<Bool> bits;
  for (size_t i = 0; i < n; i++) for (int j = 7; j >= 0; j--) bits.push_back(alloc_bit(cs, (p[i] >> j) & 1));
  return bits;
}

}  // namespace

extern "C" {

// kind 0: Sha256Circuit(preimage) of benches/sha256_spartan.rs (hash bits exposed as 256 public inputs)
// kind 1: one sha256_compression_function over 512 allocated input bits with the constant IV, no public IO
//         (the unit whose constraint count the reference quotes)
void *sha_r1cs_build(int kind, const uint8_t *data, uint64_t len, uint64_t width) {
  Built *B = new Built();
  B->width = width ? width : 2048;
  CS &cs = B->cs;
  if (kind == 0) {
    std::vector<Bool> pre = alloc_bytes_be(cs, data, len);
    std::vector<Bool> hash = sha256_gadget(cs, pre);
    for (int i = 0; i < 256; i++) {
      B->digest_bits[i] = hash[i].val;
      Var n = cs.alloc_input(hash[i].val);                        // "bit == num i": bit * 1 = n
      cs.enforce(lc_of_bool(hash[i], 1), lc_var(ONE), lc_var(n));
    }
  } else {
    std::vector<Bool> in = alloc_bytes_be(cs, data, 64);
    U32 cur[8];
    for (int i = 0; i < 8; i++) cur[i] = u_const(IV256[i]);
    compression(cs, in.data(), cur);
    for (int i = 0; i < 8; i++) for (int j = 31; j >= 0; j--) B->digest_bits[32 * i + (31 - j)] = cur[i].bits[j].val;
  }
  finalize(*B);
  return B;
}
void sha_r1cs_free(void *h) { delete (Built *)h; }
// [num_cons, num_aux, num_public, num_cons_padded, num_precommitted_padded, num_rest_padded, nnzA, nnzB, nnzC, dict size]
void sha_r1cs_sizes(void *h, uint64_t *out10) {
  Built *B = (Built *)h;
  out10[0] = B->num_cons; out10[1] = B->num_aux; out10[2] = B->num_public; out10[3] = B->num_cons_padded;
  out10[4] = B->num_pre_padded; out10[5] = B->num_rest_padded;
  for (int k = 0; k < 3; k++) out10[6 + k] = B->idx[k].size();
  out10[9] = B->dict.size();
}
void sha_r1cs_export(void *h, int k, uint32_t *coef_id, uint32_t *indices, uint32_t *indptr) {
  Built *B = (Built *)h;
  memcpy(coef_id, B->cid[k].data(), B->cid[k].size() * 4);
  memcpy(indices, B->idx[k].data(), B->idx[k].size() * 4);
  memcpy(indptr, B->ptr[k].data(), B->ptr[k].size() * 4);
}
// distinct coefficients as sign (1 = negative) + magnitude (6 little-endian u64 limbs)
void sha_r1cs_dict(void *h, uint8_t *sign, uint64_t *mag6) {
  Built *B = (Built *)h;
  for (size_t i = 0; i < B->dict.size(); i++) {
    const Big &v = B->dict[i]; sign[i] = v.neg() ? 1 : 0;
    const Big m = v.neg() ? v.negated() : v;
    memcpy(mag6 + 6 * i, m.l, 48);
  }
}
// witness bits: aux (num_aux entries), public inputs (num_public entries), digest bits (256, big-endian order)
void sha_r1cs_witness(void *h, uint8_t *aux, uint8_t *pub, uint8_t *digest_bits) {
  Built *B = (Built *)h;
  memcpy(aux, B->cs.aux.data(), B->cs.aux.size());
  if (B->cs.inputs.size()) memcpy(pub, B->cs.inputs.data(), B->cs.inputs.size());
  memcpy(digest_bits, B->digest_bits, 256);
}
int sha_r1cs_check(void *h, const uint8_t *aux, const uint8_t *pub) { return satisfied(*(Built *)h, aux, pub) ? 1 : 0; }

}  // extern "C"
