// R1CS sparse-matrix kernels: Az/Bz/Cz = M*z and the fused M^T*eq(rx) builder.
//
// Restates reference src/r1cs/sparse.rs:194-302 (classified SpMV), :305-380 (FilteredSpmv) and
// src/r1cs/mod.rs:1075-1211 (drivers), :1235-1398 (bind_and_prepare_poly_ABC / accumulate_rows).
//
// B200 design: the matrices are uploaded once (sp2_shape_upload == SplitR1CSShape::precompute) into
// dictionary-coded 8-byte entries; the transpose is built at upload so that the reference's
// scatter-add over per-thread private copies of poly_ABC (mod.rs:1299-1319) becomes a gather with
// no atomics; short segments get a thread each (coalesced 8-byte entry loads across the warp are
// not needed: the gathered 32-byte z / rx elements dominate and live in the 126 MB L2), long
// segments (SHA-256 `addmany` rows, the constant-one column) get a CTA each.
#include <string.h>
#include <algorithm>
#include <unordered_map>
#include "r1cs.cuh"

using namespace sp2;

namespace {

struct Key { uint64_t l[4]; bool operator==(const Key &o) const { return !memcmp(l, o.l, 32); } };
struct KeyHash { size_t operator()(const Key &k) const { return (size_t)(k.l[0] * 0x9E3779B97F4A7C15ull ^ (k.l[1] + 0x7F4A7C15ull) ^ (k.l[2] << 7) ^ k.l[3]); } };

// host: Montgomery one / minus one of Fq
const uint64_t H_ONE[4] = {0x1ull, 0xffffffff00000000ull, 0xffffffffffffffffull, 0x00000000fffffffeull};
const uint64_t H_MOD[4] = {0xffffffffffffffffull, 0x00000000ffffffffull, 0x0ull, 0xffffffff00000001ull};

struct HostMatrix {
  std::vector<u32> ptr; std::vector<uint2> ent; std::vector<Key> dict; std::vector<u32> long_seg;
};

// classify coefficients (PrecomputedSparseMatrix::from_sparse, sparse.rs:49-134) into a dictionary
void build_dict(const uint64_t *data, size_t nnz, std::vector<Key> &dict, std::vector<u32> &cid) {
  Key one, mone;
  memcpy(one.l, H_ONE, 32);
  { unsigned __int128 b = 0; for (int i = 0; i < 4; i++) { unsigned __int128 d = (unsigned __int128)H_MOD[i] - H_ONE[i] - (uint64_t)b; mone.l[i] = (uint64_t)d; b = (d >> 64) & 1; } }
  std::unordered_map<Key, u32, KeyHash> map;
  dict.clear(); dict.push_back(one); dict.push_back(mone);
  map[one] = 0; map[mone] = 1;
  cid.resize(nnz);
  for (size_t e = 0; e < nnz; e++) {
    Key k; memcpy(k.l, data + 4 * e, 32);
    auto it = map.find(k);
    if (it == map.end()) { u32 id = (u32)dict.size(); dict.push_back(k); map[k] = id; cid[e] = id; }
    else cid[e] = it->second;
  }
}

int upload_matrix(sp2_shape *S, const HostMatrix &h, DevMatrix *d) {
  sp2_ctx *ctx = S->ctx;
  d->nseg = (u32)h.ptr.size() - 1; d->nnz = (u32)h.ent.size(); d->ndict = (u32)h.dict.size(); d->nlong = (u32)h.long_seg.size();
  auto up = [&](void **dst, const void *src, size_t bytes) -> int {
    SP2_CUDA_OK(cudaMalloc(dst, bytes ? bytes : 32));
    S->owned.push_back(*dst);
    if (bytes) SP2_CUDA_OK(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return SP2_OK;
  };
  SP2_TRY(up((void **)&d->ptr, h.ptr.data(), h.ptr.size() * 4));
  SP2_TRY(up((void **)&d->ent, h.ent.data(), h.ent.size() * 8));
  SP2_TRY(up((void **)&d->dict, h.dict.data(), h.dict.size() * 32));
  SP2_TRY(up((void **)&d->long_seg, h.long_seg.data(), h.long_seg.size() * 4));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));    // host vectors die with the caller
  return SP2_OK;
}

void mark_long(HostMatrix &h) {
  h.long_seg.clear();
  for (size_t s = 0; s + 1 < h.ptr.size(); s++) if (h.ptr[s + 1] - h.ptr[s] > LONG_SEG) h.long_seg.push_back((u32)s);
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void accum_entry(fe &sum, const uint2 e, const fe *vec, const fe *dict) {
  const fe v = ldg_fe_ro(vec + e.x);
  if (e.y == 0) sum = Fq::add(sum, v);
  else if (e.y == 1) sum = Fq::sub(sum, v);
  else sum = Fq::add(sum, Fq::mul(ldg_fe_ro(dict + e.y), v));
}

struct Spmv3Args {
  const u32 *ptr[3]; const uint2 *ent[3]; const fe *dict[3]; const u32 *long_seg[3]; u32 nlong[3];
  const fe *base[3]; fe *out[3];
};

// short rows: one thread per row, blockIdx.y = matrix
__global__ void __launch_bounds__(256) k_spmv3(Spmv3Args a, u32 nrows, const fe *z) {
  const int k = blockIdx.y;
  const u32 row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nrows) return;
  const u32 s = a.ptr[k][row], e = a.ptr[k][row + 1];
  if (e - s > LONG_SEG) return;
  fe sum = a.base[k] ? ldg_fe(a.base[k] + row) : Fq::zero();
  for (u32 i = s; i < e; i++) accum_entry(sum, a.ent[k][i], z, a.dict[k]);
  stg_fe(a.out[k] + row, sum);
}
// long rows: one CTA of 128 threads per row
__global__ void __launch_bounds__(128) k_spmv3_long(Spmv3Args a, const fe *z) {
  __shared__ fe red[32];
  const int k = blockIdx.y;
  if (blockIdx.x >= a.nlong[k]) return;
  const u32 row = a.long_seg[k][blockIdx.x];
  const u32 s = a.ptr[k][row], e = a.ptr[k][row + 1];
  fe x[1] = {Fq::zero()};
  for (u32 i = s + threadIdx.x; i < e; i += blockDim.x) accum_entry(x[0], a.ent[k][i], z, a.dict[k]);
  block_sum_fq<1>(x, red);
  if (threadIdx.x == 0) {
    if (a.base[k]) x[0] = Fq::add(x[0], ldg_fe(a.base[k] + row));
    stg_fe(a.out[k] + row, x[0]);
  }
}

struct AbcArgs { const u32 *ptr[3]; const uint2 *ent[3]; const fe *dict[3]; };

// poly_ABC[col] = sum_A a*rx[row] + r * sum_B b*rx[row] + r^2 * sum_C c*rx[row]   (mod.rs:1324-1398)
__global__ void __launch_bounds__(256) k_abc(AbcArgs a, u32 ncols, const u32 *order, const fe *rx, const fe *r, fe *out) {
  const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ncols) return;
  const u32 col = order[t];                     // degree-sorted assignment (sp2_shape::col_order)
  u32 tot = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) tot += a.ptr[k][col + 1] - a.ptr[k][col];
  if (tot > LONG_SEG) return;
  fe s[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    s[k] = Fq::zero();
    for (u32 i = a.ptr[k][col], e = a.ptr[k][col + 1]; i < e; i++) accum_entry(s[k], a.ent[k][i], rx, a.dict[k]);
  }
  fe res = s[0];
  if (tot) {
    const fe rr = ldg_fe_ro(r);
    if (!Fq::is_zero(s[1]) || !Fq::is_zero(s[2])) res = Fq::add(s[0], Fq::mul(rr, Fq::add(s[1], Fq::mul(rr, s[2]))));
  }
  stg_fe(out + col, res);
}
// long columns: one CTA per chunk of ABC_CHUNK entries of one matrix, then one warp per long column
__global__ void __launch_bounds__(256) k_abc_long_chunks(AbcArgs a, const uint4_ *chunks, const fe *rx, fe *partial) {
  __shared__ fe red[32];
  const uint4_ c = chunks[blockIdx.x];
  fe s[1] = {Fq::zero()};
  for (u32 i = c.start + threadIdx.x; i < c.end; i += blockDim.x) accum_entry(s[0], a.ent[c.k][i], rx, a.dict[c.k]);
  block_sum_fq<1>(s, red);
  if (threadIdx.x == 0) stg_fe(partial + blockIdx.x, s[0]);
}
__global__ void __launch_bounds__(32) k_abc_long_finish(const u32 *long_cols, const u32 *chunk_first, const fe *partial, const fe *r, fe *out) {
  const u32 lc = blockIdx.x;
  fe s[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    fe v = Fq::zero();
    for (u32 c = chunk_first[3 * lc + k] + threadIdx.x; c < chunk_first[3 * lc + k + 1]; c += 32) v = Fq::add(v, ldg_fe(partial + c));
    s[k] = warp_sum_fq(v);
  }
  if (threadIdx.x == 0) {
    const fe rr = ldg_fe_ro(r);
    stg_fe(out + long_cols[lc], Fq::add(s[0], Fq::mul(rr, Fq::add(s[1], Fq::mul(rr, s[2])))));
  }
}

namespace sp2 {

int spmv3_dev(sp2_ctx *ctx, const sp2_shape *S, const DevMatrix *mats, const fe *d_z, const fe *const *base, fe *const *out) {
  Spmv3Args a; u32 maxlong = 0;
  for (int k = 0; k < 3; k++) {
    a.ptr[k] = mats[k].ptr; a.ent[k] = mats[k].ent; a.dict[k] = mats[k].dict; a.long_seg[k] = mats[k].long_seg; a.nlong[k] = mats[k].nlong;
    a.base[k] = base ? base[k] : nullptr; a.out[k] = out[k];
    maxlong = std::max(maxlong, mats[k].nlong);
  }
  const u32 nrows = (u32)S->rows_local;
  k_spmv3<<<dim3((nrows + 255) / 256, 3), 256, 0, ctx->stream>>>(a, nrows, d_z);
  SP2_LAUNCH_CHECK();
  if (maxlong) {
    k_spmv3_long<<<dim3(maxlong, 3), 128, 0, ctx->stream>>>(a, d_z);
    SP2_LAUNCH_CHECK();
  }
  return SP2_OK;
}

// side / ev_rx / ev_chunks (optional): d_rx is produced on `side` (ev_rx recorded there); the partial sums of the long columns, which
// need only d_rx, then run on `side` beside k_abc, and the main stream joins before the finish kernel
int abc_dev(sp2_ctx *ctx, const sp2_shape *S, const fe *d_rx, const fe *d_r, fe *d_out, uint64_t out_len, cudaStream_t side, cudaEvent_t ev_rx,
            cudaEvent_t ev_chunks) {
  AbcArgs a;
  for (int k = 0; k < 3; k++) { a.ptr[k] = S->T[k].ptr; a.ent[k] = S->T[k].ent; a.dict[k] = S->T[k].dict; }
  const u32 ncols = (u32)S->cols_local;                    // == num_cols on a single GPU
  if (out_len < ncols) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "abc: output shorter than num_vars + num_extra");
  if (out_len > ncols) SP2_CUDA_OK(cudaMemsetAsync(d_out + ncols, 0, (out_len - ncols) * sizeof(fe), ctx->stream));
  const bool chunks = S->nlong_cols && S->nchunks;
  if (side) {
    if (chunks) {
      k_abc_long_chunks<<<S->nchunks, 256, 0, side>>>(a, S->chunks, d_rx, S->chunk_partial);
      SP2_LAUNCH_CHECK();
      SP2_CUDA_OK(cudaEventRecord(ev_chunks, side));
    }
    SP2_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ev_rx, 0));
  }
  k_abc<<<(ncols + 255) / 256, 256, 0, ctx->stream>>>(a, ncols, S->col_order, d_rx, d_r, d_out);
  SP2_LAUNCH_CHECK();
  if (S->nlong_cols) {
    if (chunks) {
      if (side) SP2_CUDA_OK(cudaStreamWaitEvent(ctx->stream, ev_chunks, 0));
      else { k_abc_long_chunks<<<S->nchunks, 256, 0, ctx->stream>>>(a, S->chunks, d_rx, S->chunk_partial); SP2_LAUNCH_CHECK(); }
    }
    k_abc_long_finish<<<S->nlong_cols, 32, 0, ctx->stream>>>(S->long_cols, S->chunk_first, S->chunk_partial, d_r, d_out);
    SP2_LAUNCH_CHECK();
  }
  return SP2_OK;
}

}  // namespace sp2

extern "C" {

static int32_t shape_upload_impl(sp2_ctx *ctx, int rank, int nranks, uint64_t num_cons, uint64_t num_cons_unpadded, uint64_t num_shared,
                                 uint64_t num_precommitted, uint64_t num_rest, uint64_t num_public, uint64_t num_challenges,
                                 const uint64_t *const *datas, const uint32_t *const *inds, const uint32_t *const *ptrs, sp2_shape **out) {
  cudaSetDevice(ctx->device);
  if (!out) return SP2_ERR_INTERNAL;
  *out = nullptr;
  if (num_cons == 0 || (num_cons & (num_cons - 1)) || num_cons > (1ull << 31)) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "shape: num_cons must be a power of two");
  int sk = 0; while ((1 << sk) < nranks) sk++;
  if (nranks < 1 || (1 << sk) != nranks || rank < 0 || rank >= nranks || (uint64_t)nranks > num_cons)
    return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "shape: the number of ranks must be a power of two <= num_cons");
  sp2_shape *S = new sp2_shape();
  S->ctx = ctx;
  S->num_cons = num_cons; S->num_cons_unpadded = num_cons_unpadded; S->num_shared = num_shared; S->num_precommitted = num_precommitted;
  S->num_rest = num_rest; S->num_public = num_public; S->num_challenges = num_challenges;
  S->num_vars = num_shared + num_precommitted + num_rest;
  S->num_cols = S->num_vars + 1 + num_public + num_challenges;
  S->rank = rank; S->nranks = nranks; S->shard_k = sk;
  const u32 G = (u32)nranks, gmask = G - 1;
  const size_t rows_local = num_cons >> sk;
  const size_t cols_local = S->num_cols > (uint64_t)rank ? (S->num_cols - rank + G - 1) / G : 0;
  S->rows_local = rows_local; S->cols_local = cols_local;
  const u32 col_min = (u32)(num_shared + num_precommitted);
  std::vector<u32> coldeg(cols_local, 0);
  std::vector<u32> tptr[3];
  int rc = SP2_OK;
  for (int k = 0; k < 3 && rc == SP2_OK; k++) {
    const size_t nnz = ptrs[k][num_cons];
    for (size_t e = 0; e < nnz; e++) if (inds[k][e] >= S->num_cols) { rc = set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "shape: column index out of range"); break; }
    if (rc != SP2_OK) break;
    std::vector<u32> cid;
    HostMatrix hm, ht, hf;
    build_dict(datas[k], nnz, hm.dict, cid);
    ht.dict = hm.dict; hf.dict = hm.dict;
    S->nnz_total += nnz;
    for (size_t e = 0; e < nnz; e++) if (cid[e] > 1) S->nnz_general++;
    // row-major + filtered row-major, this rank's rows only (all rows on a single GPU)
    hm.ptr.resize(rows_local + 1); hf.ptr.resize(rows_local + 1);
    for (size_t lr = 0; lr < rows_local; lr++) {
      const size_t row = (lr << sk) | (size_t)rank;
      hm.ptr[lr] = (u32)hm.ent.size(); hf.ptr[lr] = (u32)hf.ent.size();
      for (u32 e = ptrs[k][row]; e < ptrs[k][row + 1]; e++) {
        const uint2 en = make_uint2(inds[k][e], cid[e]);
        hm.ent.push_back(en);
        if (inds[k][e] >= col_min) hf.ent.push_back(en);
      }
    }
    hm.ptr[rows_local] = (u32)hm.ent.size(); hf.ptr[rows_local] = (u32)hf.ent.size();
    // transpose by counting sort, this rank's columns only; entries gather eq(r_x) at the GLOBAL row
    ht.ptr.assign(cols_local + 1, 0);
    for (size_t e = 0; e < nnz; e++) if ((inds[k][e] & gmask) == (u32)rank) ht.ptr[(inds[k][e] >> sk) + 1]++;
    for (size_t c = 0; c < cols_local; c++) { coldeg[c] += ht.ptr[c + 1]; ht.ptr[c + 1] += ht.ptr[c]; }
    ht.ent.resize(ht.ptr[cols_local]);
    { std::vector<u32> cur(ht.ptr.begin(), ht.ptr.end() - 1);
      for (size_t row = 0; row < num_cons; row++)
        for (u32 e = ptrs[k][row]; e < ptrs[k][row + 1]; e++)
          if ((inds[k][e] & gmask) == (u32)rank) ht.ent[cur[inds[k][e] >> sk]++] = make_uint2((u32)row, cid[e]); }
    tptr[k] = ht.ptr;
    mark_long(hm); mark_long(hf);
    rc = upload_matrix(S, hm, &S->M[k]);
    if (rc == SP2_OK) rc = upload_matrix(S, ht, &S->T[k]);
    if (rc == SP2_OK) rc = upload_matrix(S, hf, &S->F[k]);
  }
  if (rc == SP2_OK) {
    { // degree-sorted column order for k_abc (stable; SP2_ABC_UNSORTED=1 keeps the natural order for A/B measurements)
      std::vector<u32> ord(cols_local);
      for (size_t c = 0; c < cols_local; c++) ord[c] = (u32)c;
      const char *env = getenv("SP2_ABC_UNSORTED");
      if (!(env && env[0] == '1')) std::stable_sort(ord.begin(), ord.end(), [&](u32 x, u32 y) { return coldeg[x] > coldeg[y]; });
      cudaError_t e0 = cudaMalloc((void **)&S->col_order, cols_local * 4 + 32);
      if (e0 != cudaSuccess) rc = set_cuda_error(ctx, e0, "cudaMalloc", __LINE__);
      else { S->owned.push_back(S->col_order); cudaMemcpy(S->col_order, ord.data(), cols_local * 4, cudaMemcpyHostToDevice); }
    }
    std::vector<u32> lc;
    for (size_t c = 0; c < cols_local; c++) if (coldeg[c] > LONG_SEG) lc.push_back((u32)c);
    S->nlong_cols = (u32)lc.size();
    cudaError_t e = cudaMalloc((void **)&S->long_cols, lc.size() * 4 + 32);
    if (e != cudaSuccess) rc = set_cuda_error(ctx, e, "cudaMalloc", __LINE__);
    else {
      S->owned.push_back(S->long_cols);
      if (!lc.empty()) cudaMemcpy(S->long_cols, lc.data(), lc.size() * 4, cudaMemcpyHostToDevice);
    }
    std::vector<uint4_> ch; std::vector<u32> first;
    for (size_t i = 0; i < lc.size(); i++)
      for (int k = 0; k < 3; k++) {
        first.push_back((u32)ch.size());
        for (u32 s0 = tptr[k][lc[i]], e0 = tptr[k][lc[i] + 1]; s0 < e0; s0 += ABC_CHUNK) ch.push_back(uint4_{(u32)k, s0, std::min(e0, s0 + ABC_CHUNK), 0});
      }
    first.push_back((u32)ch.size());
    S->nchunks = (u32)ch.size();
    if (rc == SP2_OK) {
      cudaError_t e2 = cudaMalloc((void **)&S->chunks, ch.size() * sizeof(uint4_) + 32);
      if (e2 == cudaSuccess) { S->owned.push_back(S->chunks); e2 = cudaMalloc((void **)&S->chunk_first, first.size() * 4 + 32); }
      if (e2 == cudaSuccess) { S->owned.push_back(S->chunk_first); e2 = cudaMalloc((void **)&S->chunk_partial, ch.size() * sizeof(fe) + 32); }
      if (e2 == cudaSuccess) {
        S->owned.push_back(S->chunk_partial);
        if (!ch.empty()) cudaMemcpy(S->chunks, ch.data(), ch.size() * sizeof(uint4_), cudaMemcpyHostToDevice);
        cudaMemcpy(S->chunk_first, first.data(), first.size() * 4, cudaMemcpyHostToDevice);
      } else rc = set_cuda_error(ctx, e2, "cudaMalloc", __LINE__);
    }
  }
  if (rc != SP2_OK) { for (void *p : S->owned) cudaFree(p); delete S; return rc; }
  *out = S;
  return SP2_OK;
}

int32_t sp2_shape_upload(sp2_ctx *ctx, uint64_t num_cons, uint64_t num_cons_unpadded, uint64_t num_shared, uint64_t num_precommitted,
                         uint64_t num_rest, uint64_t num_public, uint64_t num_challenges,
                         const uint64_t *dataA, const uint32_t *indicesA, const uint32_t *indptrA,
                         const uint64_t *dataB, const uint32_t *indicesB, const uint32_t *indptrB,
                         const uint64_t *dataC, const uint32_t *indicesC, const uint32_t *indptrC, sp2_shape **out) {
  const uint64_t *datas[3] = {dataA, dataB, dataC};
  const uint32_t *inds[3] = {indicesA, indicesB, indicesC};
  const uint32_t *ptrs[3] = {indptrA, indptrB, indptrC};
  return shape_upload_impl(ctx, 0, 1, num_cons, num_cons_unpadded, num_shared, num_precommitted, num_rest, num_public, num_challenges,
                           datas, inds, ptrs, out);
}

/* One rank's shard of the shape for the multi-GPU prover: same (whole) CSR inputs on every rank; rank g keeps the rows
 * i = g (mod nranks) of A, B, C and the columns j = g (mod nranks) of their transposes. */
int32_t sp2_shape_upload_sharded(sp2_ctx *ctx, int32_t rank, int32_t nranks, uint64_t num_cons, uint64_t num_cons_unpadded, uint64_t num_shared,
                                 uint64_t num_precommitted, uint64_t num_rest, uint64_t num_public, uint64_t num_challenges,
                                 const uint64_t *dataA, const uint32_t *indicesA, const uint32_t *indptrA,
                                 const uint64_t *dataB, const uint32_t *indicesB, const uint32_t *indptrB,
                                 const uint64_t *dataC, const uint32_t *indicesC, const uint32_t *indptrC, sp2_shape **out) {
  const uint64_t *datas[3] = {dataA, dataB, dataC};
  const uint32_t *inds[3] = {indicesA, indicesB, indicesC};
  const uint32_t *ptrs[3] = {indptrA, indptrB, indptrC};
  return shape_upload_impl(ctx, rank, nranks, num_cons, num_cons_unpadded, num_shared, num_precommitted, num_rest, num_public, num_challenges,
                           datas, inds, ptrs, out);
}

void sp2_shape_free(sp2_shape *S) {
  if (!S) return;
  cudaSetDevice(S->ctx->device);
  cudaStreamSynchronize(S->ctx->stream);
  for (void *p : S->owned) cudaFree(p);
  delete S;
}

/* sizes: [num_cons, num_vars, num_cols, nnz(A+B+C), nnz with a general coefficient, long rows, long columns] */
int32_t sp2_shape_sizes(const sp2_shape *S, uint64_t *out7) {
  out7[0] = S->num_cons; out7[1] = S->num_vars; out7[2] = S->num_cols; out7[3] = S->nnz_total; out7[4] = S->nnz_general;
  out7[5] = S->M[0].nlong + S->M[1].nlong + S->M[2].nlong; out7[6] = S->nlong_cols;
  return SP2_OK;
}

int32_t sp2_spmv3_dev(sp2_ctx *ctx, const sp2_shape *S, const void *d_z, void *d_az, void *d_bz, void *d_cz) {
  cudaSetDevice(ctx->device);
  fe *out[3] = {(fe *)d_az, (fe *)d_bz, (fe *)d_cz};
  return spmv3_dev(ctx, S, S->M, (const fe *)d_z, nullptr, out);
}

/* multiply_vec_incremental_into (mod.rs:1170-1211): az = cached_az + A[:, cols >= shared+precommitted] * z */
int32_t sp2_spmv3_incremental_dev(sp2_ctx *ctx, const sp2_shape *S, const void *d_z, const void *d_cached_az, const void *d_cached_bz,
                                  const void *d_cached_cz, void *d_az, void *d_bz, void *d_cz) {
  cudaSetDevice(ctx->device);
  const fe *base[3] = {(const fe *)d_cached_az, (const fe *)d_cached_bz, (const fe *)d_cached_cz};
  fe *out[3] = {(fe *)d_az, (fe *)d_bz, (fe *)d_cz};
  return spmv3_dev(ctx, S, S->F, (const fe *)d_z, base, out);
}

int32_t sp2_spmv3(sp2_ctx *ctx, const sp2_shape *S, const uint64_t *z, uint64_t z_len, uint64_t *az, uint64_t *bz, uint64_t *cz) {
  cudaSetDevice(ctx->device);
  if (z_len != S->num_cols) return set_error(ctx, SP2_ERR_INVALID_WITNESS_LENGTH, "multiply_vec: z has the wrong length");
  const size_t nb = S->rows_local * sizeof(fe);
  void *dz, *da, *db, *dc;
  SP2_TRY(scratch(ctx, 0, z_len * sizeof(fe), &dz)); SP2_TRY(scratch(ctx, 1, nb, &da)); SP2_TRY(scratch(ctx, 2, nb, &db)); SP2_TRY(scratch(ctx, 3, nb, &dc));
  SP2_CUDA_OK(cudaMemcpyAsync(dz, z, z_len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_TRY(sp2_spmv3_dev(ctx, S, dz, da, db, dc));
  SP2_CUDA_OK(cudaMemcpyAsync(az, da, nb, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(bz, db, nb, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(cz, dc, nb, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

int32_t sp2_spmv3_incremental(sp2_ctx *ctx, const sp2_shape *S, const uint64_t *z, uint64_t z_len, const uint64_t *cached_az,
                              const uint64_t *cached_bz, const uint64_t *cached_cz, uint64_t *az, uint64_t *bz, uint64_t *cz) {
  cudaSetDevice(ctx->device);
  if (z_len != S->num_cols) return set_error(ctx, SP2_ERR_INVALID_WITNESS_LENGTH, "multiply_vec_incremental: z has the wrong length");
  const size_t nb = S->rows_local * sizeof(fe);
  void *dz, *d[3];
  SP2_TRY(scratch(ctx, 0, z_len * sizeof(fe), &dz));
  for (int k = 0; k < 3; k++) SP2_TRY(scratch(ctx, 1 + k, nb, &d[k]));
  SP2_CUDA_OK(cudaMemcpyAsync(dz, z, z_len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  const uint64_t *cached[3] = {cached_az, cached_bz, cached_cz};
  for (int k = 0; k < 3; k++) SP2_CUDA_OK(cudaMemcpyAsync(d[k], cached[k], nb, cudaMemcpyHostToDevice, ctx->stream));
  SP2_TRY(sp2_spmv3_incremental_dev(ctx, S, dz, d[0], d[1], d[2], d[0], d[1], d[2]));
  uint64_t *outs[3] = {az, bz, cz};
  for (int k = 0; k < 3; k++) SP2_CUDA_OK(cudaMemcpyAsync(outs[k], d[k], nb, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

int32_t sp2_abc_dev(sp2_ctx *ctx, const sp2_shape *S, const void *d_rx, const void *d_r, void *d_out, uint64_t out_len) {
  cudaSetDevice(ctx->device);
  return abc_dev(ctx, S, (const fe *)d_rx, (const fe *)d_r, (fe *)d_out, out_len);
}

int32_t sp2_abc(sp2_ctx *ctx, const sp2_shape *S, const uint64_t *rx, uint64_t rx_len, const uint64_t *r, uint64_t *out, uint64_t out_len) {
  cudaSetDevice(ctx->device);
  if (rx_len != S->num_cons) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "bind_and_prepare_poly_ABC: rx must have num_cons entries");
  void *drx, *dr, *dout;
  SP2_TRY(scratch(ctx, 0, rx_len * sizeof(fe), &drx)); SP2_TRY(scratch(ctx, 1, sizeof(fe), &dr)); SP2_TRY(scratch(ctx, 2, out_len * sizeof(fe), &dout));
  SP2_CUDA_OK(cudaMemcpyAsync(drx, rx, rx_len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(dr, r, sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_TRY(abc_dev(ctx, S, (const fe *)drx, (const fe *)dr, (fe *)dout, out_len));
  SP2_CUDA_OK(cudaMemcpyAsync(out, dout, out_len * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

}  // extern "C"
