// Multi-scalar multiplication and Hyrax row commitments on the device.
//
// Restates (as group elements — the affine results are what the reference emits and absorbs):
//   src/provider/msm.rs:59-222   cpu_msm_serial / msm  (signed-digit Pippenger, c = 8 at n = 2048)
//   src/provider/msm.rs:367-620  msm_small family (binary / <=10 bit / windowed) — same sums
//   src/provider/msm.rs:637-774  FixedBaseMul (8-bit window tables) — blind * h terms
//   src/provider/pcs/hyrax_pc.rs:207-319  HyraxPCS::commit / commit_zeros (one Pedersen row per 2048)
//   src/provider/pcs/hyrax_pc.rs:38-54    bind_with_delayed (LZ = L^T W)
//
// B200 design: latency, not arithmetic, bounds these MSMs (2048 terms, 2 per proof on the critical path), so the
// bucket method is dropped altogether.  The commitment key is fixed at setup and HBM is plentiful, so the key
// upload precomputes, for EVERY base, the reference's FixedBaseMul table (msm.rs:651-689, which the reference
// builds only for h and for <=64-wide keys): table[b][w][d] = d * 2^(8w) * base_b for the 33 signed byte-digit
// windows and d = 1..128, affine (554 MB at 2051 bases).  An MSM is then a pure gather-and-sum of
// <= 33 * n affine points: no buckets, no bucket reduction, no doublings.
//   k_msm_gather: warp-granular — a warp sums 128 consecutive (term, window) table entries of ONE job: every lane
//                 adds ~4 entries (affine + affine, then mixed Jacobian adds), then the warp tree; no barriers, no shared points;
//   k_msm_final:  one CTA per job sums the warp partials, adds the job's extra terms' partials and its optional
//                 addend and emits ONE Jacobian point.
//   Both trees run LANE-QUAD additions (curve_quad.cuh): the multiplications of one Jacobian addition are spread over
//   four lanes (5 dependent multiplication levels instead of 16), because in a tree most lanes are idle anyway.
// A job (MsmJob, msm.cuh) = a scalar vector over consecutive key bases + up to 3 extra (base, scalar) terms taken from
// the key's or an auxiliary table (blind * h, the folded-commitment rows of the NeutronNova prover) + an optional
// affine / Jacobian addend, so rerandomisation (U + r h), commit_zeros, the IPA's two-term commitments and the
// "fold by linearity" commitments are all one batch of table lookups.
// Normalisation to affine (one field inversion) is done by the caller on the host for the whole batch
// (host_transcript.h: batch_normalize): ~15 us there vs ~130 us for a serial inversion on a GPU thread.
// Many MSMs (Hyrax rows, the prover's blinded terms) are batched into one pair of launches.
#include <string.h>
#include "msm.cuh"
#include "devutil.cuh"
#include "curve_quad.cuh"
#include "host_transcript.h"

using namespace sp2;

namespace {

__device__ __forceinline__ jac ld_jac(const jac *p) { jac r; r.x = ldg_fe(&p->x); r.y = ldg_fe(&p->y); r.z = ldg_fe(&p->z); return r; }
__device__ __forceinline__ void st_jac(jac *p, const jac &v) { stg_fe(&p->x, v.x); stg_fe(&p->y, v.y); stg_fe(&p->z, v.z); }
__device__ __forceinline__ aff ld_aff_ro(const aff *p) { aff r; r.x = ldg_fe_ro(&p->x); r.y = ldg_fe_ro(&p->y); return r; }

// ---- key precompute ----------------------------------------------------------------------------
// thread (b, w): the 128 multiples d * 2^(8w) * base_b, Jacobian, plus the running product of their Z's
__global__ void __launch_bounds__(128) k_ck_multiples(const aff *bases, u32 nbase, jac *tmp, fe *pre) {
  const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nbase * MSM_NW) return;
  const u32 b = idx / MSM_NW, w = idx % MSM_NW;
  jac p = jac_from_aff(ld_aff_ro(bases + b));
#pragma unroll 1
  for (u32 k = 0; k < 8 * w; k++) p = jac_dbl(p);
  jac m = p;
  fe acc = Fp::one();
  jac *out = tmp + (size_t)idx * MSM_ND;
  fe *po = pre + (size_t)idx * (MSM_ND + 1);
#pragma unroll 1
  for (int d = 0; d < MSM_ND; d++) {
    st_jac(out + d, m);
    stg_fe(po + d, acc);
    if (!Fp::is_zero(m.z)) acc = Fp::mul(acc, m.z);
    m = jac_add(m, p);
  }
  stg_fe(po + MSM_ND, acc);     // (one spare slot per thread: total product)
}
// batch normalisation over the 128 multiples of one (b, w) (Montgomery's trick; cf. batch_normalize, msm.rs:669-676)
__global__ void __launch_bounds__(128) k_ck_normalize(const jac *tmp, const fe *pre, u32 nbase, aff *table) {
  const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nbase * MSM_NW) return;
  const jac *in = tmp + (size_t)idx * MSM_ND;
  const fe *pi = pre + (size_t)idx * (MSM_ND + 1);
  aff *out = table + (size_t)idx * MSM_ND;
  fe inv = Fp::inv(ldg_fe(pi + MSM_ND));
#pragma unroll 1
  for (int d = MSM_ND - 1; d >= 0; d--) {
    const jac p = ld_jac(in + d);
    aff a;
    if (Fp::is_zero(p.z)) { a.x = Fp::zero(); a.y = Fp::zero(); }
    else { a = jac_to_aff_with_inv(p, Fp::mul(inv, ldg_fe(pi + d))); inv = Fp::mul(inv, p.z); }
    stg_fe(&out[d].x, a.x);
    stg_fe(&out[d].y, a.y);
  }
}

// ---- gather + sum ------------------------------------------------------------------------------
// Warp-granular: gather warp g sums MSM_WARP_LOOKUPS consecutive (term, window) table entries of ONE job — lanes
// accumulate ~4 entries each with mixed additions, then a 5-level shuffle tree — and writes one partial sum.  No
// __syncthreads, no shared-memory points: a job of one term (commit_zeros, rerandomisation, blind * h: 33 entries) costs
// one warp and ~1 mixed + 5 full additions of latency instead of a 256-thread CTA walking an 8-level barrier tree, and
// the thousands of row jobs of a NeutronNova prove pack eight to a CTA.  k_msm_final: one CTA per job sums its partials
// (strided serial adds, warp shuffle tree, 4-warp smem step) and adds the job's optional addend.
constexpr int GW = MSM_THREADS / 32;            // gather warps per CTA

// job that owns partial `part` (jobs are sorted by first_part): binary search
__device__ __forceinline__ u32 job_of_part(const MsmJob *jobs, u32 njobs, u32 part) {
  u32 lo = 0, hi = njobs - 1;
  while (lo < hi) { const u32 mid = (lo + hi + 1) >> 1; if (jobs[mid].first_part <= part) lo = mid; else hi = mid - 1; }
  return lo;
}

struct GatherSmem {
  short dig[GW][8][MSM_NW + 1];      // signed byte digits of the <= 6 terms a warp's 128 entries span
  u32 bidx[GW][8];
  const aff *tab[GW][8];
};

__global__ void __launch_bounds__(MSM_THREADS) k_msm_gather(const MsmJob *jobs, u32 njobs, u32 total_parts, const aff *table, jac *partial, jac *out) {
  __shared__ GatherSmem sm;
  const u32 wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid-stride over the partials: a capped grid (msm_run's max_ctas) keeps a background batch on a side stream from taking
  // every SM away from the latency-bound kernels of the main stream
  for (u32 part = blockIdx.x * GW + wib; part < total_parts; part += gridDim.x * GW) {
  const u32 jid = job_of_part(jobs, njobs, part);
  const MsmJob job = jobs[jid];
  const u32 total = (job.len + job.nextra) * MSM_NW;
  const u32 p0 = (part - job.first_part) * MSM_WARP_LOOKUPS, p1 = min(total, p0 + MSM_WARP_LOOKUPS);
  const u32 t0 = p0 / MSM_NW, nterm = p1 > p0 ? (p1 - 1) / MSM_NW - t0 + 1 : 0;
  // signed byte digits of each scalar (to_repr() little-endian bytes, msm.rs:97-100; digit recoding :122-148)
  if (lane < nterm) {
    const u32 g = t0 + lane;
    fe s; u32 b; const aff *tb = table;
    if (g < job.len) { s = ldg_fe(job.scalars + g); b = job.base0 + g; }
    else { const u32 e = g - job.len; s = ldg_fe(job.extra_scalar[e]); b = job.extra_base[e]; if (job.extra_tab[e]) tb = job.extra_tab[e]; }
    s = Fq::from_mont(s);
    sm.bidx[wib][lane] = b; sm.tab[wib][lane] = tb;
    int carry = 0;
#pragma unroll
    for (int w = 0; w < 32; w++) {
      int d = (int)((s.v[w >> 2] >> (8 * (w & 3))) & 0xffu) + carry;
      carry = 0;
      if (d > 128) { d -= 256; carry = 1; }
      sm.dig[wib][lane][w] = (short)d;
    }
    sm.dig[wib][lane][32] = (short)carry;
  }
  __syncwarp();
  jac acc = jac_inf();
  aff first; first.x = Fp::zero(); first.y = Fp::zero();
  u32 cnt = 0;
  for (u32 p = p0 + lane; p < p1; p += 32) {
    const u32 t = p / MSM_NW, w = p - t * MSM_NW, lt = t - t0;
    const int d = sm.dig[wib][lt][w];
    if (d) {
      aff pt = ld_aff_ro(sm.tab[wib][lt] + ((size_t)sm.bidx[wib][lt] * MSM_NW + w) * MSM_ND + (u32)((d < 0 ? -d : d) - 1));
      if (d < 0) pt.y = Fp::neg(pt.y);
      // the lane's first two entries are both affine: mmadd (6 multiplications) instead of a copy + madd (11)
      if (cnt == 0) first = pt;
      else if (cnt == 1) acc = aff_add_to_jac(first, pt);
      else acc = jac_add_mixed(acc, pt);
      cnt++;
    }
  }
  if (cnt == 1) acc = jac_from_aff(first);
  __syncwarp();
  acc = warp_sum_jac_quad(acc);
  if (lane == 0) {
    if (job.nparts == 1 && !job.add_jac) {
      // a job of one partial (<= 3 terms: commit_zeros rows, rerandomisation, the two-term commitments at the end of a prove) is
      // complete here: no k_msm_final launch for it
      if (job.add_aff) acc = jac_add_mixed(acc, ld_aff_ro(job.add_aff));
      st_jac(out + jid, acc);
    } else st_jac(partial + part, acc);
  }
  __syncwarp();
  }
}

// One CTA of NT threads = NT/4 lane quads per job (curve_quad.cuh): quad k sums the partials k, k + NQ, ... (quad additions),
// then a tree over the quads — through shared memory while it spans warps, by shuffles inside warp 0 — and thread 0 adds
// the job's optional addends.  A 2048-term job (528 partials) is 4 + 7 quad additions = 55 multiplication latencies; the
// single-lane version (4-5 strided adds, 5-level warp tree, 4-warp step) was ~190.  `list`: the jobs of this size class.
template <int NT>
__global__ void __launch_bounds__(NT) k_msm_final(const MsmJob *jobs, const u32 *list, const jac *partial, jac *out) {
  constexpr u32 NQ = NT / 4;
  __shared__ jac pts[NQ > 8 ? NQ : 1];
  const u32 jid = list[blockIdx.x];
  const MsmJob job = jobs[jid];
  const u32 n = job.nparts, tid = threadIdx.x, lane = tid & 31, wib = tid >> 5, k = tid >> 2;
  const u32 m = n < NQ ? n : NQ;                               // quads that hold a sum
  const jac *src = partial + job.first_part;
  jac acc = k < n ? ld_jac(src + k) : jac_inf();
  for (u32 base = NQ; base < n; base += NQ) {
    if (base + wib * 8 < n) {                                  // warp-uniform: some quad of this warp has a partial
      const u32 idx = base + k;
      acc = quad_jac_add(acc, idx < n ? ld_jac(src + idx) : jac_inf());
    }
  }
  if (NQ > 8) {
#pragma unroll 1
    for (u32 stride = NQ / 2; stride >= 8; stride >>= 1) {
      if (stride < m) {                                        // CTA-uniform
        if (k >= stride && k < 2 * stride && (tid & 3) == 0) pts[k] = acc;
        __syncthreads();
        if (k < stride) acc = quad_jac_add(acc, pts[k + stride]);   // whole warps (stride >= 8 quads)
      }
    }
  }
  if (wib == 0) {
#pragma unroll 1
    for (u32 d = 16; d >= 4; d >>= 1) {
      if ((d >> 2) < m) {
        jac o = shfl_idx_jac(acc, (lane + d) & 31);
        if (lane + d >= 32) o = jac_inf();
        acc = quad_jac_add(acc, o);
      }
    }
  }
  if (tid == 0) {
    if (job.add_aff) acc = jac_add_mixed(acc, ld_aff_ro(job.add_aff));
    if (job.add_jac) acc = jac_add(acc, ld_jac(job.add_jac));
    st_jac(out + jid, acc);
  }
}

// n Jacobian points -> affine: one thread per chunk of <= 16 points, one inversion per chunk (Montgomery's trick)
constexpr int NORM_CHUNK = 16;
__global__ void __launch_bounds__(64) k_batch_normalize(const jac *in, u64 n, aff *out) {
  const u64 c = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  const u64 i0 = c * NORM_CHUNK; if (i0 >= n) return;
  const u64 i1 = min(n, i0 + NORM_CHUNK);
  fe pre[NORM_CHUNK];
  fe acc = Fp::one();
#pragma unroll 1
  for (u64 i = i0; i < i1; i++) { pre[i - i0] = acc; const fe z = ldg_fe(&in[i].z); if (!Fp::is_zero(z)) acc = Fp::mul(acc, z); }
  fe inv = Fp::inv(acc);
#pragma unroll 1
  for (u64 i = i1; i-- > i0;) {
    const jac p = ld_jac(in + i);
    aff a;
    if (Fp::is_zero(p.z)) { a.x = Fp::zero(); a.y = Fp::zero(); }
    else { a = jac_to_aff_with_inv(p, Fp::mul(inv, pre[i - i0])); inv = Fp::mul(inv, p.z); }
    stg_fe(&out[i].x, a.x); stg_fe(&out[i].y, a.y);
  }
}

// ---- Hyrax bind: LZ[i] = sum_j L[j] * W[j * r_len + i], delayed reduction (hyrax_pc.rs:38-54) ----
constexpr int BIND_RG = 32;
__global__ void __launch_bounds__(128) k_hyrax_bind(const fe *poly, const fe *L, u64 rows, u64 r_len, fe *partial) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r_len) return;
  Fq::acc acc = Fq::acc_zero();
  for (u64 j = blockIdx.y; j < rows; j += gridDim.y) Fq::mul_acc(acc, ldg_fe_ro(L + j), ldg_fe(poly + j * r_len + i));
  stg_fe(partial + (u64)blockIdx.y * r_len + i, Fq::acc_reduce(acc));
}
__global__ void __launch_bounds__(128) k_sum_rows(const fe *partial, u64 nparts, u64 r_len, fe *out) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r_len) return;
  // eight loads in flight per step: the loads and the carry chains are volatile asm and keep their program order, so a
  // load-add-load-add loop pays one L2 round trip per partial (15 us for 32 partials)
  fe s = Fq::zero();
  for (u64 p = 0; p < nparts; p += 8) {
    fe v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) v[q] = p + q < nparts ? ldg_fe(partial + (p + q) * r_len + i) : Fq::zero();
#pragma unroll
    for (int q = 0; q < 8; q++) s = Fq::add(s, v[q]);
  }
  stg_fe(out + i, s);
}


// ---- harness support: n pseudo-random curve points s_i * G (no hash-to-curve on this side of the ABI; the
// reference derives its generators in halo2curves, src/provider/traits.rs:205-249, and ships them as bases) ----
__device__ __forceinline__ u64 splitmix64(u64 &x) { x += 0x9E3779B97F4A7C15ull; u64 z = x; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
__global__ void __launch_bounds__(64) k_test_points(u64 seed, u32 n, aff *out) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // generator (3, 0x5a6dd32d...f02d) of T256 (verified there)
  fe gx, gy;
  const u32 GY[8] = {0x25b1f02du, 0x30b49549u, 0xa351bb3cu, 0xecd9d538u, 0xbe66600du, 0x4e97345cu, 0xf58708e6u, 0x5a6dd32du};
#pragma unroll
  for (int k = 0; k < 8; k++) { gx.v[k] = k == 0 ? 3u : 0u; gy.v[k] = GY[k]; }
  aff g; g.x = Fp::to_mont(gx); g.y = Fp::to_mont(gy);
  u64 st = seed ^ (0xD1B54A32D192ED03ull * (u64)(i + 1));
  u32 k[8];
#pragma unroll
  for (int j = 0; j < 4; j++) { const u64 w = splitmix64(st); k[2 * j] = (u32)w; k[2 * j + 1] = (u32)(w >> 32); }
  k[7] &= 0x7fffffffu; k[0] |= 1u;                   // 0 < k < group order
  jac acc = jac_inf();
  for (int b = 254; b >= 0; b--) {
    acc = jac_dbl(acc);
    if ((k[b >> 5] >> (b & 31)) & 1u) acc = jac_add_mixed(acc, g);
  }
  const aff a = jac_to_aff(acc);
  stg_fe(&out[i].x, a.x); stg_fe(&out[i].y, a.y);
}

}  // namespace

namespace sp2 {

// d_out[njobs]: Jacobian results (the caller normalises: sp2h::batch_normalize on the host, batch_normalize_dev here)
int msm_run(sp2_ctx *ctx, const sp2_ck *ck, const std::vector<MsmJob> &jobs_in, jac *d_out, cudaStream_t stream, int slot_jobs, int slot_partials,
            unsigned max_ctas) {
  if (jobs_in.empty()) return SP2_OK;
  if (!stream) stream = ctx->stream;
  std::vector<MsmJob> jobs(jobs_in);
  u32 parts = 0;
  for (auto &j : jobs) {
    if (j.nextra > 3) return set_error(ctx, SP2_ERR_INTERNAL, "msm: at most three extra terms per job");
    if (j.base0 + j.len > ck->nbase) return set_error(ctx, SP2_ERR_INVALID_CK_LENGTH, "msm: more scalars than commitment-key bases");
    const u32 total = (j.len + j.nextra) * MSM_NW;
    j.first_part = parts; j.nparts = std::max<u32>(1, (total + MSM_WARP_LOOKUPS - 1) / MSM_WARP_LOOKUPS);
    parts += j.nparts;
  }
  const size_t nj = jobs.size();
  // size classes of the final reduction (k_msm_final<NT>): <= 8 partials one warp, <= 32 four warps, else 128 quads
  std::vector<u32> cls[3];
  for (size_t i = 0; i < nj; i++) {
    if (jobs[i].nparts == 1 && !jobs[i].add_jac) continue;          // finished by the gather warp itself
    cls[jobs[i].nparts <= 8 ? 0 : jobs[i].nparts <= 32 ? 1 : 2].push_back((u32)i);
  }
  std::vector<unsigned char> blob(nj * sizeof(MsmJob) + nj * sizeof(u32));
  memcpy(blob.data(), jobs.data(), nj * sizeof(MsmJob));
  u32 *lists = (u32 *)(blob.data() + nj * sizeof(MsmJob));
  size_t off[3], o = 0;
  for (int c = 0; c < 3; c++) { off[c] = o; if (!cls[c].empty()) memcpy(lists + o, cls[c].data(), cls[c].size() * sizeof(u32)); o += cls[c].size(); }
  void *d_jobs, *d_partial;
  SP2_TRY(scratch(ctx, slot_jobs, blob.size(), &d_jobs));
  SP2_TRY(scratch(ctx, slot_partials, (size_t)parts * sizeof(jac), &d_partial));
  // pageable source: the runtime stages it before cudaMemcpyAsync returns, so the local vector may die
  SP2_CUDA_OK(cudaMemcpyAsync(d_jobs, blob.data(), blob.size(), cudaMemcpyHostToDevice, stream));
  const u32 *d_lists = (const u32 *)((const unsigned char *)d_jobs + nj * sizeof(MsmJob));
  unsigned grid = (parts + GW - 1) / GW; if (max_ctas && grid > max_ctas) grid = max_ctas;
  k_msm_gather<<<grid, MSM_THREADS, 0, stream>>>((const MsmJob *)d_jobs, (u32)nj, parts, ck->table, (jac *)d_partial, d_out);
  SP2_LAUNCH_CHECK();
  if (!cls[0].empty()) { k_msm_final<32><<<(unsigned)cls[0].size(), 32, 0, stream>>>((const MsmJob *)d_jobs, d_lists + off[0], (const jac *)d_partial, d_out); SP2_LAUNCH_CHECK(); }
  if (!cls[1].empty()) { k_msm_final<128><<<(unsigned)cls[1].size(), 128, 0, stream>>>((const MsmJob *)d_jobs, d_lists + off[1], (const jac *)d_partial, d_out); SP2_LAUNCH_CHECK(); }
  if (!cls[2].empty()) { k_msm_final<512><<<(unsigned)cls[2].size(), 512, 0, stream>>>((const MsmJob *)d_jobs, d_lists + off[2], (const jac *)d_partial, d_out); SP2_LAUNCH_CHECK(); }
  return SP2_OK;
}

int batch_normalize_dev(sp2_ctx *ctx, const jac *d_in, uint64_t n, aff *d_out) {
  if (!n) return SP2_OK;
  const u64 chunks = (n + NORM_CHUNK - 1) / NORM_CHUNK;
  k_batch_normalize<<<(unsigned)((chunks + 63) / 64), 64, 0, ctx->stream>>>(d_in, n, d_out);
  SP2_LAUNCH_CHECK();
  return SP2_OK;
}

// window tables of `nbase` device-resident affine bases: table[b][w][d-1] = d * 2^(8w) * base_b
int msm_build_tables(sp2_ctx *ctx, const aff *d_bases, uint32_t nbase, aff **table_out) {
  *table_out = nullptr;
  jac *tmp = nullptr; fe *pre = nullptr; aff *table = nullptr;
  const size_t nent = (size_t)nbase * MSM_NW * MSM_ND;
  cudaError_t e = cudaMalloc((void **)&tmp, nent * sizeof(jac));
  if (e == cudaSuccess) e = cudaMalloc((void **)&pre, (size_t)nbase * MSM_NW * (MSM_ND + 1) * sizeof(fe));
  if (e == cudaSuccess) e = cudaMalloc((void **)&table, nent * sizeof(aff));
  if (e == cudaSuccess) {
    const unsigned blocks = (nbase * MSM_NW + 127) / 128;
    k_ck_multiples<<<blocks, 128, 0, ctx->stream>>>(d_bases, nbase, tmp, pre);
    k_ck_normalize<<<blocks, 128, 0, ctx->stream>>>(tmp, pre, nbase, table);
    ctx->launches += 2;
    e = cudaStreamSynchronize(ctx->stream);
  }
  if (pre) cudaFree(pre);
  if (tmp) cudaFree(tmp);
  if (e != cudaSuccess) { if (table) cudaFree(table); return set_cuda_error(ctx, e, "msm table build", __LINE__); }
  *table_out = table;
  return SP2_OK;
}

// HyraxPCS::commit rows on a given stream (see sp2_hyrax_commit_dev)
int hyrax_commit_rows(sp2_ctx *ctx, const sp2_ck *ck, const fe *d_v, uint64_t len, const fe *d_blinds, uint64_t rows, jac *d_out, cudaStream_t stream,
                      int slot_jobs, int slot_partials) {
  if (rows < (len + ck->n - 1) / ck->n) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "hyrax commit: too few rows for the vector");
  std::vector<MsmJob> jobs(rows);
  for (uint64_t i = 0; i < rows; i++) {
    MsmJob &j = jobs[i]; memset(&j, 0, sizeof(j));
    const uint64_t lo = i * ck->n, hi = std::min<uint64_t>(len, lo + ck->n);
    j.scalars = d_v + lo; j.len = hi > lo ? (u32)(hi - lo) : 0; j.base0 = 0;
    j.nextra = 1; j.extra_base[0] = ck->idx_h(); j.extra_scalar[0] = d_blinds + i;
  }
  return msm_run(ctx, ck, jobs, d_out, stream, slot_jobs, slot_partials);
}

int hyrax_bind_dev(sp2_ctx *ctx, const fe *d_poly, const fe *d_L, uint64_t rows, uint64_t r_len, fe *d_out, cudaStream_t stream, int slot) {
  if (!stream) stream = ctx->stream;
  const unsigned rg = (unsigned)std::min<uint64_t>(rows, BIND_RG);
  void *part;
  SP2_TRY(scratch(ctx, slot, (size_t)rg * r_len * sizeof(fe), &part));
  k_hyrax_bind<<<dim3((unsigned)((r_len + 127) / 128), rg), 128, 0, stream>>>(d_poly, d_L, rows, r_len, (fe *)part);
  SP2_LAUNCH_CHECK();
  k_sum_rows<<<(unsigned)((r_len + 127) / 128), 128, 0, stream>>>((const fe *)part, rg, r_len, d_out);
  SP2_LAUNCH_CHECK();
  return SP2_OK;
}

}  // namespace sp2

extern "C" {

int32_t sp2_ck_upload(sp2_ctx *ctx, const uint64_t *bases_xy, uint32_t n, const uint64_t *h_xy, const uint64_t *ck_s_xy,
                      const uint64_t *h_s_xy, sp2_ck **out) {
  cudaSetDevice(ctx->device);
  if (!out) return SP2_ERR_INTERNAL;
  *out = nullptr;
  if (n == 0 || n > (1u << 24)) return set_error(ctx, SP2_ERR_INVALID_CK_LENGTH, "ck: bad number of bases");
  sp2_ck *ck = new sp2_ck();
  ck->ctx = ctx; ck->n = n; ck->nbase = n + 3;
  aff *d_bases = nullptr;
  cudaError_t e = cudaMalloc((void **)&d_bases, (size_t)ck->nbase * sizeof(aff));
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_bases, bases_xy, (size_t)n * sizeof(aff), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_bases + n, h_xy, sizeof(aff), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_bases + n + 1, ck_s_xy, sizeof(aff), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_bases + n + 2, h_s_xy, sizeof(aff), cudaMemcpyHostToDevice, ctx->stream);
  int rc = e == cudaSuccess ? msm_build_tables(ctx, d_bases, ck->nbase, &ck->table) : set_cuda_error(ctx, e, "ck upload", __LINE__);
  if (d_bases) cudaFree(d_bases);
  if (rc != SP2_OK) { delete ck; return rc; }
  *out = ck;
  return SP2_OK;
}

void sp2_ck_free(sp2_ck *ck) {
  if (!ck) return;
  cudaSetDevice(ck->ctx->device);
  cudaStreamSynchronize(ck->ctx->stream);
  if (ck->table) cudaFree(ck->table);
  delete ck;
}

/* DlogGroupExt::vartime_multiscalar_mul(scalars, ck[..n]) (provider/traits.rs:118-134 -> msm.rs:187) */
int32_t sp2_msm(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *scalars, uint32_t n, uint64_t *out_xy) {
  cudaSetDevice(ctx->device);
  if (n > ck->n) return set_error(ctx, SP2_ERR_INVALID_CK_LENGTH, "msm: more scalars than commitment-key bases");
  void *d_s, *d_o;
  SP2_TRY(scratch(ctx, 0, (size_t)n * sizeof(fe) + 32, &d_s)); SP2_TRY(scratch(ctx, 1, sizeof(jac), &d_o));
  if (n) SP2_CUDA_OK(cudaMemcpyAsync(d_s, scalars, (size_t)n * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  MsmJob j; memset(&j, 0, sizeof(j));
  j.scalars = (const fe *)d_s; j.len = n; j.base0 = 0; j.nextra = 0;
  SP2_TRY(msm_run(ctx, ck, std::vector<MsmJob>{j}, (jac *)d_o));
  uint64_t hj[12];
  SP2_CUDA_OK(cudaMemcpyAsync(hj, d_o, sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  sp2h::batch_normalize(hj, 1, out_xy);
  return SP2_OK;
}

/* HyraxPCS::commit (hyrax_pc.rs:207-303; commit_zeros :305-319 is the all-zero v): rows of ck->n scalars,
 * row i = sum_j v[i*n + j] * ck_j + blinds[i] * h.  d_v may be shorter than rows*n (last row ragged).
 * d_out_rows_jac: rows Jacobian points (x, y, z: 12 limbs each; z = 0 is the identity) — normalise on the host. */
int32_t sp2_hyrax_commit_dev(sp2_ctx *ctx, const sp2_ck *ck, const void *d_v, uint64_t len, const void *d_blinds, uint64_t rows,
                             void *d_out_rows_jac) {
  cudaSetDevice(ctx->device);
  return sp2::hyrax_commit_rows(ctx, ck, (const fe *)d_v, len, (const fe *)d_blinds, rows, (jac *)d_out_rows_jac, nullptr, 10, 11);
}

int32_t sp2_hyrax_commit(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *v, uint64_t len, const uint64_t *blinds, uint64_t rows,
                         int32_t is_small, uint64_t *out_rows) {
  (void)is_small;   // the reference's small-scalar hint only selects a faster CPU path (hyrax_pc.rs:266-290); same group element
  cudaSetDevice(ctx->device);
  void *d_v, *d_b, *d_o;
  SP2_TRY(scratch(ctx, 0, len * sizeof(fe) + 32, &d_v)); SP2_TRY(scratch(ctx, 1, rows * sizeof(fe) + 32, &d_b));
  SP2_TRY(scratch(ctx, 2, rows * sizeof(jac) + 32, &d_o));
  if (len) SP2_CUDA_OK(cudaMemcpyAsync(d_v, v, len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(d_b, blinds, rows * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_TRY(sp2_hyrax_commit_dev(ctx, ck, d_v, len, d_b, rows, d_o));
  std::vector<uint64_t> hj(rows * 12);
  SP2_CUDA_OK(cudaMemcpyAsync(hj.data(), d_o, rows * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  sp2h::batch_normalize(hj.data(), rows, out_rows);
  return SP2_OK;
}

/* HyraxPCS::commit_without_blind (hyrax_pc.rs:533-567): the raw row points <v_row_i, ck>, identity (all zero) for an all-zero row */
int32_t sp2_hyrax_commit_without_blind(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *v, uint64_t len, int32_t is_small, uint64_t *out_rows) {
  (void)is_small;
  cudaSetDevice(ctx->device);
  const uint64_t rows = (len + ck->n - 1) / ck->n;
  if (!rows) return SP2_OK;
  void *d_v, *d_o;
  SP2_TRY(scratch(ctx, 0, len * sizeof(fe) + 32, &d_v)); SP2_TRY(scratch(ctx, 2, rows * sizeof(jac) + 32, &d_o));
  SP2_CUDA_OK(cudaMemcpyAsync(d_v, v, len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<MsmJob> jobs(rows);
  for (uint64_t i = 0; i < rows; i++) {
    MsmJob &j = jobs[i]; memset(&j, 0, sizeof(j));
    const uint64_t lo = i * ck->n, hi = std::min<uint64_t>(len, lo + ck->n);
    j.scalars = (const fe *)d_v + lo; j.len = (u32)(hi - lo);
  }
  SP2_TRY(msm_run(ctx, ck, jobs, (jac *)d_o));
  std::vector<uint64_t> hj(rows * 12);
  SP2_CUDA_OK(cudaMemcpyAsync(hj.data(), d_o, rows * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  sp2h::batch_normalize(hj.data(), rows, out_rows);
  return SP2_OK;
}

/* HyraxPCS::commit_incremental (hyrax_pc.rs:569-607): out[i] = raw[i] (identity past n_raw) + <delta_row_i, ck> + blinds[i] * h */
int32_t sp2_hyrax_commit_incremental(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *raw_rows_xy, uint64_t n_raw, const uint64_t *delta, uint64_t len,
                                     const uint64_t *blinds, uint64_t *out_rows) {
  cudaSetDevice(ctx->device);
  const uint64_t rows = (len + ck->n - 1) / ck->n;
  if (!rows) return SP2_OK;
  void *d_v, *d_b, *d_o, *d_r;
  SP2_TRY(scratch(ctx, 0, len * sizeof(fe) + 32, &d_v)); SP2_TRY(scratch(ctx, 1, rows * sizeof(fe) + 32, &d_b));
  SP2_TRY(scratch(ctx, 2, rows * sizeof(jac) + 32, &d_o)); SP2_TRY(scratch(ctx, 3, rows * sizeof(aff) + 32, &d_r));
  SP2_CUDA_OK(cudaMemcpyAsync(d_v, delta, len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(d_b, blinds, rows * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemsetAsync(d_r, 0, rows * sizeof(aff), ctx->stream));
  if (n_raw) SP2_CUDA_OK(cudaMemcpyAsync(d_r, raw_rows_xy, std::min<uint64_t>(n_raw, rows) * sizeof(aff), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<MsmJob> jobs(rows);
  for (uint64_t i = 0; i < rows; i++) {
    MsmJob &j = jobs[i]; memset(&j, 0, sizeof(j));
    const uint64_t lo = i * ck->n, hi = std::min<uint64_t>(len, lo + ck->n);
    j.scalars = (const fe *)d_v + lo; j.len = (u32)(hi - lo);
    j.nextra = 1; j.extra_base[0] = ck->idx_h(); j.extra_scalar[0] = (const fe *)d_b + i; j.add_aff = (const aff *)d_r + i;
  }
  SP2_TRY(msm_run(ctx, ck, jobs, (jac *)d_o));
  std::vector<uint64_t> hj(rows * 12);
  SP2_CUDA_OK(cudaMemcpyAsync(hj.data(), d_o, rows * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  sp2h::batch_normalize(hj.data(), rows, out_rows);
  return SP2_OK;
}

/* HyraxPCS::rerandomize_commitment (hyrax_pc.rs:321-344): out[i] = comm[i] + (r_new[i] - r_old[i]) * h (fixed-base table of h) */
__global__ void k_fe_sub(const fe *a, const fe *b, fe *o, u64 n) { const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) stg_fe(o + i, Fq::sub(ldg_fe(a + i), ldg_fe(b + i))); }
int32_t sp2_hyrax_rerandomize(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *comm_rows_xy, const uint64_t *r_old, const uint64_t *r_new, uint64_t rows,
                              uint64_t *out_rows) {
  cudaSetDevice(ctx->device);
  if (!rows) return SP2_OK;
  void *d_a, *d_b, *d_o, *d_r;
  SP2_TRY(scratch(ctx, 0, 2 * rows * sizeof(fe) + 32, &d_a)); SP2_TRY(scratch(ctx, 1, rows * sizeof(fe) + 32, &d_b));
  SP2_TRY(scratch(ctx, 2, rows * sizeof(jac) + 32, &d_o)); SP2_TRY(scratch(ctx, 3, rows * sizeof(aff) + 32, &d_r));
  SP2_CUDA_OK(cudaMemcpyAsync(d_a, r_new, rows * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync((fe *)d_a + rows, r_old, rows * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(d_r, comm_rows_xy, rows * sizeof(aff), cudaMemcpyHostToDevice, ctx->stream));
  k_fe_sub<<<(unsigned)((rows + 127) / 128), 128, 0, ctx->stream>>>((const fe *)d_a, (const fe *)d_a + rows, (fe *)d_b, rows);
  SP2_LAUNCH_CHECK();
  std::vector<MsmJob> jobs(rows);
  for (uint64_t i = 0; i < rows; i++) {
    MsmJob &j = jobs[i]; memset(&j, 0, sizeof(j));
    j.nextra = 1; j.extra_base[0] = ck->idx_h(); j.extra_scalar[0] = (const fe *)d_b + i; j.add_aff = (const aff *)d_r + i;
  }
  SP2_TRY(msm_run(ctx, ck, jobs, (jac *)d_o));
  std::vector<uint64_t> hj(rows * 12);
  SP2_CUDA_OK(cudaMemcpyAsync(hj.data(), d_o, rows * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  sp2h::batch_normalize(hj.data(), rows, out_rows);
  return SP2_OK;
}

/* bind_with_delayed (hyrax_pc.rs:38-54): out[i] = sum_j L[j] * poly[j * r_len + i] */
int32_t sp2_hyrax_bind(sp2_ctx *ctx, const uint64_t *poly, const uint64_t *L, uint64_t rows, uint64_t r_len, uint64_t *out) {
  cudaSetDevice(ctx->device);
  void *d_p, *d_L, *d_o;
  SP2_TRY(scratch(ctx, 0, rows * r_len * sizeof(fe), &d_p)); SP2_TRY(scratch(ctx, 1, rows * sizeof(fe), &d_L)); SP2_TRY(scratch(ctx, 2, r_len * sizeof(fe), &d_o));
  SP2_CUDA_OK(cudaMemcpyAsync(d_p, poly, rows * r_len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(d_L, L, rows * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
  SP2_TRY(hyrax_bind_dev(ctx, (const fe *)d_p, (const fe *)d_L, rows, r_len, (fe *)d_o));
  SP2_CUDA_OK(cudaMemcpyAsync(out, d_o, r_len * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

/* harness support: n pseudo-random T256 points (seeded multiples of the generator), affine, to the host */
int32_t sp2_test_points(sp2_ctx *ctx, uint64_t seed, uint32_t n, uint64_t *out_xy) {
  cudaSetDevice(ctx->device);
  void *d; SP2_TRY(scratch(ctx, 0, (size_t)n * sizeof(aff) + 64, &d));
  k_test_points<<<(n + 63) / 64, 64, 0, ctx->stream>>>(seed, n, (aff *)d);
  SP2_LAUNCH_CHECK();
  SP2_CUDA_OK(cudaMemcpyAsync(out_xy, d, (size_t)n * sizeof(aff), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return SP2_OK;
}

}  // extern "C"
