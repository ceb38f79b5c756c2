// Multilinear-polynomial building blocks shared by the kernels: eq~ tables and top-variable bind.
// Restates reference src/polys/eq.rs:59-117 (EqPolynomial::evals_from_points, MSB-first) and
// src/polys/multilinear.rs:95-164 (bind_poly_var_top: Z[i] <- Z[i] + r*(Z[i+n] - Z[i])).
#pragma once
#include "devutil.cuh"

namespace sp2 {

// All prefix tables of eq~ over `pts[0..n)`, built by ONE block:
//   tab_k = eq~(pts[n-k..n), .)  (2^k entries, MSB-first), stored at out + (2^k - 1), k = 0..n.
// This is the doubling recurrence of eq.rs:66-90 (y = x*r; x -= y) with every intermediate kept —
// exactly the prefix family EqSumCheckInstance::new builds (sumcheck.rs:960-987).
__device__ __forceinline__ void eq_prefix_block(const fe *pts, int n, fe *out) {
  if (threadIdx.x == 0) stg_fe(out, Fq::one());
  __syncthreads();
  for (int i = 0; i < n; i++) {
    const size_t sz = (size_t)1 << i;
    const fe *prev = out + (sz - 1);
    fe *next = out + (2 * sz - 1);
    const fe t = ldg_fe(pts + (n - 1 - i));
    for (size_t j = threadIdx.x; j < sz; j += blockDim.x) {
      fe x = ldg_fe(prev + j);
      fe y = Fq::mul(x, t);
      stg_fe(next + sz + j, y);
      stg_fe(next + j, Fq::sub(x, y));
    }
    __syncthreads();
  }
}

// lo + r*(hi - lo)
__device__ __forceinline__ fe bind_pair(const fe &lo, const fe &hi, const fe &r) {
  return Fq::add(lo, Fq::mul(Fq::sub(hi, lo), r));
}

}  // namespace sp2
