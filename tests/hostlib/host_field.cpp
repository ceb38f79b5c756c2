// Host build of the device arithmetic headers (prim.cuh emulates the PTX carry chain), exported
// through a tiny C ABI so pytest can differential-test the exact kernel arithmetic on CPU.
#include <cstddef>
#include "../../spartan2_b200/csrc/field.cuh"
using namespace sp2;
extern "C" {
#define BIN(name, F, op) void name(const fe *a, const fe *b, fe *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = F::op(a[i], b[i]); }
BIN(ht_fq_mul, Fq, mul) BIN(ht_fq_add, Fq, add) BIN(ht_fq_sub, Fq, sub)
BIN(ht_fp_mul, Fp, mul) BIN(ht_fp_add, Fp, add) BIN(ht_fp_sub, Fp, sub)
void ht_fq_mul_il(const fe *a, const fe *b, fe *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = mont_mul_interleaved<FqParams>(a[i], b[i]); }
void ht_fp_mul_il(const fe *a, const fe *b, fe *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = mont_mul_interleaved<FpParams>(a[i], b[i]); }
void ht_fp_mul_cios(const fe *a, const fe *b, fe *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = Fp::mul_cios(a[i], b[i]); }
void ht_fq_inv(const fe *a, fe *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = Fq::inv(a[i]); }
void ht_fp_inv(const fe *a, fe *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = Fp::inv(a[i]); }
void ht_fq_half(const fe *a, fe *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = Fq::half(a[i]); }
void ht_fq_from_mont(const fe *a, fe *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = Fq::from_mont(a[i]); }
void ht_fq_to_mont(const fe *a, fe *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = Fq::to_mont(a[i]); }
void ht_fp_from_mont(const fe *a, fe *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = Fp::from_mont(a[i]); }
void ht_fq_from_uniform(const fe *lohi, fe *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = Fq::from_uniform(lohi[2 * i], lohi[2 * i + 1]); }
void ht_fq_dot(const fe *a, const fe *b, size_t n, fe *o) {
  Fq::acc acc = Fq::acc_zero();
  for (size_t i = 0; i < n; i++) Fq::mul_acc(acc, a[i], b[i]);
  *o = Fq::acc_reduce(acc);
}
void ht_fq_acc_reduce(const u32 *limbs17, fe *o) { Fq::acc a; for (int i = 0; i < 17; i++) a.v[i] = limbs17[i]; *o = Fq::acc_reduce(a); }
void ht_mul_wide(const fe *a, const fe *b, u32 *o16) { u32 w[16]; mul_wide(w, *a, *b); for (int i = 0; i < 16; i++) o16[i] = w[i]; }
}
