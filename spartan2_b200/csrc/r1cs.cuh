// Device-resident R1CS shape (reference src/r1cs/mod.rs:743-911 SplitR1CSShape, src/r1cs/sparse.rs).
#pragma once
#include <vector>
#include "ctx.cuh"
#include "devutil.cuh"

namespace sp2 {

// One sparse matrix in compressed form, either row-major (CSR, for M*z) or column-major (the transpose
// in CSR form, for M^T * eq(rx)).  Entry = (index of the gathered vector element, coefficient id).
// Coefficients are dictionary-coded: R1CS matrices hold a handful of distinct values (+-1, small
// integers, powers of two) — the observation behind the reference's coefficient classes
// (sparse.rs:29-45: unit_pos / unit_neg / small / general) — so an entry is 8 bytes instead of 36.
// dict[0] = 1, dict[1] = -1 (adds/subs, no multiply); anything else is one Montgomery multiply.
struct DevMatrix {
  u32 nseg = 0;            // rows (CSR) or columns (transpose)
  u32 nnz = 0;
  u32 *ptr = nullptr;      // nseg + 1
  uint2 *ent = nullptr;    // nnz: .x = gathered index, .y = coefficient id
  fe *dict = nullptr;
  u32 ndict = 0;
  u32 *long_seg = nullptr; // segments with more than LONG_SEG entries (handled by a block each)
  u32 nlong = 0;
};
constexpr u32 LONG_SEG = 48;
constexpr u32 ABC_CHUNK = 2048;
struct uint4_ { u32 k, start, end, pad; };

}  // namespace sp2

struct sp2_shape {
  sp2_ctx *ctx = nullptr;
  uint64_t num_cons = 0, num_cons_unpadded = 0, num_shared = 0, num_precommitted = 0, num_rest = 0, num_public = 0, num_challenges = 0;
  uint64_t num_vars = 0, num_cols = 0;       // num_cols = num_vars + 1 + num_public + num_challenges
  // Multi-GPU shard of the shape: rank g of G = 2^shard_k keeps the rows i = g (mod G) of M and F
  // (local row i >> shard_k: Az/Bz/Cz land directly in the cyclic sum-check layout) and the columns j = g (mod G) of
  // the transposes T (local column j >> shard_k: poly_ABC lands in the inner sum-check's cyclic layout).  Gathered
  // vectors (z for M/F, eq(r_x) for T) are replicated and indexed globally.  rank 0 of 1 = the whole shape.
  int rank = 0, nranks = 1, shard_k = 0;
  uint64_t rows_local = 0, cols_local = 0;   // num_cons >> shard_k; number of columns j < num_cols with j = rank (mod G)
  sp2::DevMatrix M[3];                       // A, B, C row-major
  sp2::DevMatrix T[3];                       // transposes (column-major) for bind_and_prepare_poly_ABC
  sp2::DevMatrix F[3];                       // row-major, columns >= num_shared + num_precommitted only (FilteredSpmv)
  sp2::u32 *long_cols = nullptr; sp2::u32 nlong_cols = 0;   // columns whose A+B+C degree exceeds LONG_SEG
  // columns in order of decreasing A+B+C degree: thread t of k_abc takes column col_order[t], so the lanes of a warp
  // walk segments of (nearly) equal length — a thread-per-column gather is bound by the LONGEST column of each warp
  sp2::u32 *col_order = nullptr;
  // long columns (the constant-one column of SHA-256 has ~10^6 entries) are cut into chunks of ABC_CHUNK entries,
  // one CTA each: chunk = (matrix, first entry, last entry); chunk_first[3*lc + k] .. chunk_first[3*lc + k + 1]
  // are the chunks of long column lc in matrix k
  sp2::uint4_ *chunks = nullptr; sp2::u32 nchunks = 0; sp2::u32 *chunk_first = nullptr; sp2::fe *chunk_partial = nullptr;
  std::vector<void *> owned;
  uint64_t nnz_total = 0, nnz_general = 0;
};

namespace sp2 {
int spmv3_dev(sp2_ctx *ctx, const sp2_shape *S, const DevMatrix *mats, const fe *d_z, const fe *const *base, fe *const *out);
int abc_dev(sp2_ctx *ctx, const sp2_shape *S, const fe *d_rx, const fe *d_r, fe *d_out, uint64_t out_len, cudaStream_t side = nullptr,
            cudaEvent_t ev_rx = nullptr, cudaEvent_t ev_chunks = nullptr);
}  // namespace sp2
