// Host build of the device curve arithmetic (curve.cuh) for CPU differential tests against the oracle.
#include <cstddef>
#include "../../spartan2_b200/csrc/curve.cuh"
using namespace sp2;
extern "C" {
void ht_add_mixed(const aff *a, const aff *b, aff *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = jac_to_aff(jac_add_mixed(jac_from_aff(a[i]), b[i])); }
// (2a) + (2b) through the full Jacobian add with non-trivial Z on both sides
void ht_add_full_of_doubles(const aff *a, const aff *b, aff *o, size_t n) {
  for (size_t i = 0; i < n; i++) o[i] = jac_to_aff(jac_add(jac_dbl(jac_from_aff(a[i])), jac_dbl(jac_from_aff(b[i]))));
}
void ht_add_full(const aff *a, const aff *b, aff *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = jac_to_aff(jac_add(jac_from_aff(a[i]), jac_from_aff(b[i]))); }
void ht_dbl(const aff *a, aff *o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = jac_to_aff(jac_dbl(jac_from_aff(a[i]))); }
// k * a for a canonical 256-bit little-endian scalar (8 x u32), double-and-add with mixed adds
void ht_scalar_mul(const aff *a, const u32 *k, aff *o) {
  jac acc = jac_inf();
  for (int i = 255; i >= 0; i--) { acc = jac_dbl(acc); if ((k[i >> 5] >> (i & 31)) & 1) acc = jac_add_mixed(acc, *a); }
  *o = jac_to_aff(acc);
}
int ht_on_curve(const aff *a) { return aff_on_curve(*a) ? 1 : 0; }
}
