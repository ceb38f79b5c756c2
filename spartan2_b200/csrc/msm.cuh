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
constexpr int MSM_WARP_LOOKUPS = 128;   // table entries summed by one gather warp (4 per lane), see msm.cu
constexpr int MSM_THREADS = 256;

// One linear combination  sum_i s_i * base[base0 + i]  (+ up to three extra terms, e.g. blind * h, each from the key's table
// or from an auxiliary table of the same [base][33][128] layout)  (+ an optional addend point).
struct MsmJob {
  const fe *scalars;        // device pointer, Montgomery form, `len` entries (may be null when len == 0)
  u32 len;
  u32 base0;                // index of the first base in the key's table
  u32 nextra;               // <= 3
  u32 extra_base[3];
  const fe *extra_scalar[3];
  const aff *extra_tab[3];  // nullptr: the key's table
  const aff *add_aff;       // optional addend, affine (identity = (0,0)) ...
  const jac *add_jac;       // ... or Jacobian
  u32 first_part, nparts;   // filled by msm_run: the job's partial sums (one per gather warp)
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
// stream / scratch slots: a second MSM batch in flight on a side stream needs its own job and partial buffers
// max_ctas (0 = no cap): grid size limit of the gather kernel
int msm_run(sp2_ctx *ctx, const sp2_ck *ck, const std::vector<MsmJob> &jobs, jac *d_out, cudaStream_t stream = nullptr, int slot_jobs = 10,
            int slot_partials = 11, unsigned max_ctas = 0);
int hyrax_commit_rows(sp2_ctx *ctx, const sp2_ck *ck, const fe *d_v, uint64_t len, const fe *d_blinds, uint64_t rows, jac *d_out, cudaStream_t stream,
                      int slot_jobs, int slot_partials);
int hyrax_bind_dev(sp2_ctx *ctx, const fe *d_poly, const fe *d_L, uint64_t rows, uint64_t r_len, fe *d_out, cudaStream_t stream = nullptr, int slot = 12);
// window tables [nbase][MSM_NW][MSM_ND] of arbitrary device-resident affine bases (identity allowed): *table_out is cudaMalloc'ed
int msm_build_tables(sp2_ctx *ctx, const aff *d_bases, uint32_t nbase, aff **table_out);
// n Jacobian points -> affine on the device (one inversion per 16-point chunk, Montgomery's trick), identity -> (0,0)
int batch_normalize_dev(sp2_ctx *ctx, const jac *d_in, uint64_t n, aff *d_out);
}  // namespace sp2
