// Commitment-provider device state and batched MSM interface (reference src/provider/msm.rs,
// src/provider/pcs/hyrax_pc.rs).
#pragma once
#include <vector>
#include "ctx.cuh"
#include "curve.cuh"

namespace sp2 {

constexpr int MSM_C = 8;                 // window bits (reference: c = ceil(ln n) = 8 at n = 2048, msm.rs:60-66)
constexpr int MSM_NW = 33;               // 32 byte windows + the signed-digit carry window (msm.rs:110, 122-148)
constexpr int MSM_ND = 128;              // signed digits: multiples 1..128 per window (msm.rs:114-116)
constexpr int MSM_TERMS = 32;            // terms per gather CTA
constexpr int MSM_THREADS = 256;

// One linear combination sum_i s_i * base[base0 + i]  (+ up to two extra terms, e.g. blind * h).
struct MsmJob {
  const fe *scalars;        // device pointer, Montgomery form, `len` entries (may be null when len == 0)
  u32 len;
  u32 base0;                // index of the first base in the key's table
  u32 nextra;
  u32 extra_base[2];
  const fe *extra_scalar[2];
  u32 nblk;                 // filled by msm_run
  u32 pad;
};

}  // namespace sp2

struct sp2_ck {
  sp2_ctx *ctx = nullptr;
  uint32_t n = 0;           // number of row bases ck[0..n)
  uint32_t nbase = 0;       // n + 3: [ck_0 .. ck_{n-1}, h, ck_s, h_s]
  sp2::aff *table = nullptr;   // [nbase][MSM_NW][MSM_ND]: table[b][w][d-1] = d * 2^(8w) * base_b, affine
  uint32_t idx_h() const { return n; }
  uint32_t idx_ck_s() const { return n + 1; }
  uint32_t idx_h_s() const { return n + 2; }
};

namespace sp2 {
// run `jobs` (host array; pointers inside are device pointers); d_out[njobs] JACOBIAN results
// (normalise on the host: sp2h::batch_normalize)
int msm_run(sp2_ctx *ctx, const sp2_ck *ck, const std::vector<MsmJob> &jobs, jac *d_out);
int hyrax_bind_dev(sp2_ctx *ctx, const fe *d_poly, const fe *d_L, uint64_t rows, uint64_t r_len, fe *d_out);
}  // namespace sp2
