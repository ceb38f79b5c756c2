// On-box verifier: SpartanSNARK::verify (reference src/spartan.rs:469-578) with its data-parallel parts on the device —
//   evaluate_with_tables_fast (src/r1cs/mod.rs:36-146, 1216-1226): the three matrix MLE evaluations
//       eval_M = sum_{(i,j) in M} M_ij T_x[i] T_y[j] = < T_x, M T_y >   =  the row-gather SpMV of the prover with z := T_y
//       (k_spmv3 over the dictionary-coded rows) followed by three dot products with T_x — O(nnz), no new matrix format;
//   HyraxPCS::verify (src/provider/pcs/hyrax_pc.rs:480-531): comm_LZ = MSM(L, commitment rows) — a VARIABLE-base MSM
//       (the rows are proof data), through the device Pippenger of msm_var.cu;
//   InnerProductArgumentLinear::verify (src/provider/pcs/ipa.rs:173-221): MSM(z_vec, ck) + z_delta h through the key's
//       window tables, r * comm_LZ + delta and r * comm_eval + beta as 2-term variable-base MSMs, <z_vec, R> on the device.
// The sum-check replay (l + m + 1 rounds of four scalars) and the transcript stay on the host, as in the prover.
// SURVEY.md §8 row f1.  Checked against oracle/oracle.c: orc_spartan_verify (accept + every tampered field rejected).
#include <string.h>
#include <vector>
#include "ctx.cuh"
#include "devutil.cuh"
#include "hostfield.h"
#include "msm.cuh"
#include "r1cs.cuh"

using namespace sp2;

namespace sp2 {
int eq_table_dev(sp2_ctx *ctx, const fe *d_r, uint32_t k, fe *d_out, cudaStream_t stream = nullptr, int slot = 15);
}

namespace {

// out[q] = <a_q, b> for q < 3 (one CTA per q)
__global__ void __launch_bounds__(256) k_dot3(const fe *a0, const fe *a1, const fe *a2, const fe *b, u64 n, fe *out) {
  __shared__ fe red[32];
  const fe *a = blockIdx.x == 0 ? a0 : blockIdx.x == 1 ? a1 : a2;
  Fq::acc acc = Fq::acc_zero();
  for (u64 i = threadIdx.x; i < n; i += blockDim.x) Fq::mul_acc(acc, ldg_fe(a + i), ldg_fe(b + i));
  fe x[1] = {Fq::acc_reduce(acc)};
  block_sum_fq<1>(x, red);
  if (threadIdx.x == 0) stg_fe(out + blockIdx.x, x[0]);
}

typedef NnHost H;
fe le_canon(const fe &m) { uint64_t a[4], c[4]; HF::ld(m, a); sp2h::from_mont(a, sp2h::FQ_MOD, sp2h::FQ_INV, c); return HF::st(c); }
// UniPoly::to_transcript_bytes (univariate.rs:182-190): every coefficient but the linear one, to_repr() little-endian
void absorb_unipoly(sp2h::Transcript &ts, const fe *co, int n) {
  ts.push("p", 1);
  for (int k = 0; k < n; k++) if (k != 1) { const fe c = le_canon(co[k]); ts.push(c.v, 32); }
}
fe squeeze(sp2h::Transcript &ts, const char *label) { uint8_t dg[64]; uint64_t o[4]; ts.squeeze(label, dg); sp2h::fq_from_uniform(dg, o); return H::load(o); }
// SumcheckProof::verify (sumcheck.rs:67-114) on compressed polynomials (CompressedUniPoly::decompress, univariate.rs:166-179)
void sumcheck_verify(sp2h::Transcript &ts, const uint64_t *polys, size_t rounds, int degree, fe claim, fe *e_out, std::vector<fe> &r) {
  const int nc = degree;                                    // stored coefficients per round: all but the linear one
  fe e = claim; r.resize(rounds);
  for (size_t i = 0; i < rounds; i++) {
    fe co[4];
    co[0] = H::load(polys + 4 * (nc * i));
    for (int k = 2; k <= degree; k++) co[k] = H::load(polys + 4 * (nc * i + k - 1));
    fe lin = HF::sub(HF::sub(e, co[0]), co[0]);
    for (int k = 2; k <= degree; k++) lin = HF::sub(lin, co[k]);
    co[1] = lin;
    absorb_unipoly(ts, co, degree + 1);
    r[i] = squeeze(ts, "c");
    e = H::eval(co, degree + 1, r[i]);
  }
  *e_out = e;
}
void eq_evals_host(const fe *r, size_t k, std::vector<fe> &out) {    // EqPolynomial::evals_from_points (eq.rs:59-92), MSB-first
  out.assign((size_t)1 << k, HF::zero()); out[0] = HF::one(); size_t size = 1;
  for (size_t t = k; t-- > 0;) { for (size_t i = 0; i < size; i++) { const fe hi = HF::mul(out[i], r[t]); out[size + i] = hi; out[i] = HF::sub(out[i], hi); } size *= 2; }
}

}  // namespace

extern "C" {

int32_t sp2_msm_var(sp2_ctx *ctx, const uint64_t *scalars, const uint64_t *bases_xy, uint32_t n, uint64_t *out_xy);

/* SpartanSNARK::verify (src/spartan.rs:469-578).  SP2_OK = accept; SP2_ERR_PROOF_VERIFY = reject (sp2_last_error names the failing
 * check: outer sum-check, inner sum-check / matrix evaluations, or the PCS argument); other codes = malformed input. */
int32_t sp2_spartan_verify(sp2_ctx *ctx, const sp2_shape *S, const sp2_ck *ck, const uint8_t *vk_digest, const uint64_t *public_values,
                           const sp2_spartan_proof *proof) {
  cudaSetDevice(ctx->device);
  if (S->nranks != 1) return set_error(ctx, SP2_ERR_UNSUPPORTED, "verify: whole shape expected");
  const uint64_t width = ck->n, nv = S->num_vars, N = S->num_cons;
  int l = 0; while (((uint64_t)1 << l) < N) l++;
  int m = 0; while (((uint64_t)1 << m) < nv) m++;
  const int nry = m + 1;
  const uint64_t num_extra = 1 + S->num_public + S->num_challenges, rows = nv / width;
  if (proof->num_rounds_x != (uint64_t)l || proof->num_rounds_y != (uint64_t)nry || proof->num_comm_rows != rows || proof->num_cols != width)
    return set_error(ctx, SP2_ERR_PROOF_VERIFY, "verify: proof dimensions do not match the shape");
  int nvr = 0; while (((uint64_t)1 << nvr) < rows) nvr++;
  if (((uint64_t)1 << nvr) != rows || (width & (width - 1)) || l > 40 || nry > 40) return set_error(ctx, SP2_ERR_INVALID_CK_LENGTH, "verify: rows and width must be powers of two");
  const fe one = HF::one(), zero = HF::zero();
  // ---- transcript head (spartan.rs:476-497) ---------------------------------------------------------------------------
  sp2h::Transcript ts("SpartanSNARK");
  ts.absorb_bytes("vk", vk_digest, 32);
  ts.absorb_scalars("public_values", public_values, S->num_public);
  const uint64_t sh_rows = S->num_shared / width, pre_rows = S->num_precommitted / width, rest_rows = S->num_rest / width;
  if (sh_rows) ts.absorb_commitment("comm_W_shared", proof->comm_W, sh_rows);
  if (pre_rows) ts.absorb_commitment("comm_W_precommitted", proof->comm_W + 8 * sh_rows, pre_rows);
  ts.absorb_commitment("comm_W_rest", proof->comm_W + 8 * (sh_rows + pre_rows), rest_rows);
  std::vector<fe> tau(l);
  for (int i = 0; i < l; i++) tau[i] = squeeze(ts, "t");
  // ---- outer sum-check (spartan.rs:499-517) ------------------------------------------------------------------------------
  fe e_outer; std::vector<fe> rx, ry;
  sumcheck_verify(ts, proof->outer_polys, (size_t)l, 3, zero, &e_outer, rx);
  fe tb = one;
  for (int i = 0; i < l; i++) tb = HF::mul(tb, HF::add(HF::mul(rx[i], tau[i]), HF::mul(HF::sub(one, rx[i]), HF::sub(one, tau[i]))));   // EqPolynomial::evaluate
  const fe cA = H::load(proof->claims_outer), cB = H::load(proof->claims_outer + 4), cC = H::load(proof->claims_outer + 8);
  if (!HF::eq(HF::mul(tb, HF::sub(HF::mul(cA, cB), cC)), e_outer)) return set_error(ctx, SP2_ERR_PROOF_VERIFY, "verify: outer sum-check final claim mismatch");
  ts.absorb_scalars("claims_outer", proof->claims_outer, 3);
  const fe r = squeeze(ts, "r"), r2 = HF::sqr(r);
  const fe joint = HF::add(HF::add(cA, HF::mul(r, cB)), HF::mul(r2, cC));
  // ---- inner sum-check (spartan.rs:519-551) ------------------------------------------------------------------------------
  fe e_inner;
  sumcheck_verify(ts, proof->inner_polys, (size_t)nry, 2, joint, &e_inner, ry);
  // eval_X = SparsePolynomial([1 | X]).evaluate(r_y[1..]) (polys/multilinear.rs:190-207), eval_Z
  std::vector<fe> Xv(num_extra); Xv[0] = one;
  for (uint64_t j = 0; j < S->num_public; j++) Xv[1 + j] = H::load(public_values + 4 * j);
  fe eval_X;
  { size_t p2 = 1, nvz = 0; while (p2 < num_extra) { p2 <<= 1; nvz++; }
    if ((int)nvz + 1 > m) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "verify: too many public inputs for the witness length");
    const size_t skip = (size_t)m - 1 - nvz, k = (size_t)m - skip;
    std::vector<fe> chis; eq_evals_host(&ry[1 + skip], k, chis);
    fe acc = zero;
    for (uint64_t i = 0; i < num_extra; i++) acc = HF::add(acc, HF::mul(Xv[i], chis[i]));
    fe common = one;
    for (size_t i = 0; i < skip; i++) common = HF::mul(common, HF::sub(one, ry[1 + i]));
    eval_X = HF::mul(common, acc); }
  const fe eval_W = H::load(proof->eval_W);
  const fe eval_Z = HF::add(HF::mul(HF::sub(one, ry[0]), eval_W), HF::mul(ry[0], eval_X));
  // ---- evaluate_with_tables_fast on the device: eval_M = < T_x, M T_y > ------------------------------------------------------
  void *d_r, *d_tx, *d_ty, *d_t3;
  // (slots 18-23 belong to the verifier: sp2_msm_var, called below, stages its operands in slots 0 / 1)
  SP2_TRY(scratch(ctx, 18, (size_t)(l + nry + 8) * sizeof(fe), &d_r));
  SP2_TRY(scratch(ctx, 19, N * sizeof(fe), &d_tx)); SP2_TRY(scratch(ctx, 20, 2 * nv * sizeof(fe), &d_ty)); SP2_TRY(scratch(ctx, 21, (3 * N + 8) * sizeof(fe), &d_t3));
  { std::vector<fe> up(rx); up.insert(up.end(), ry.begin(), ry.end());
    SP2_CUDA_OK(cudaMemcpyAsync(d_r, up.data(), up.size() * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream)); }
  SP2_TRY(eq_table_dev(ctx, (const fe *)d_r, (uint32_t)l, (fe *)d_tx));
  SP2_TRY(eq_table_dev(ctx, (const fe *)d_r + l, (uint32_t)nry, (fe *)d_ty));
  { fe *o[3] = {(fe *)d_t3, (fe *)d_t3 + N, (fe *)d_t3 + 2 * N};
    SP2_TRY(spmv3_dev(ctx, S, S->M, (const fe *)d_ty, nullptr, o));
    k_dot3<<<3, 256, 0, ctx->stream>>>(o[0], o[1], o[2], (const fe *)d_tx, N, (fe *)d_t3 + 3 * N);
    SP2_LAUNCH_CHECK(); }
  fe ev[3];
  SP2_CUDA_OK(cudaMemcpyAsync(ev, (fe *)d_t3 + 3 * N, 3 * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  const fe comb = HF::mul(HF::add(HF::add(ev[0], HF::mul(r, ev[1])), HF::mul(r2, ev[2])), eval_Z);
  if (!HF::eq(comb, e_inner)) return set_error(ctx, SP2_ERR_PROOF_VERIFY, "verify: inner sum-check final claim does not match the matrix evaluations");
  // ---- PCS::verify (hyrax_pc.rs:480-531) + IPA verify (ipa.rs:173-221) ---------------------------------------------------------
  ts.absorb_commitment("poly_com", proof->comm_W, rows);
  std::vector<fe> L; eq_evals_host(&ry[1], (size_t)nvr, L);
  uint64_t comm_LZ[8];
  if (nvr == 0) memcpy(comm_LZ, proof->comm_W, 64);
  else SP2_TRY(sp2_msm_var(ctx, (const uint64_t *)L.data(), proof->comm_W, (uint32_t)rows, comm_LZ));
  // table-driven part: comm_eval = eval_W ck_s + blind h_s ; rhs1 = <z_vec, ck> + z_delta h ; rhs2 = <z_vec, R> ck_s + z_beta h_s
  void *d_s, *d_R, *d_pts;
  SP2_TRY(scratch(ctx, 22, (width + 16) * sizeof(fe), &d_s)); SP2_TRY(scratch(ctx, 23, width * sizeof(fe) + 8 * sizeof(jac), &d_R)); d_pts = (fe *)d_R + width;
  fe *sm = (fe *)d_s + width;      // small scalars: [eval_W, blind_eval, z_delta, z_beta, ip]
  { std::vector<fe> up(width + 8);
    memcpy(up.data(), proof->z_vec, width * sizeof(fe));
    up[width] = eval_W; up[width + 1] = H::load(proof->blind_eval_W); up[width + 2] = H::load(proof->z_delta); up[width + 3] = H::load(proof->z_beta);
    SP2_CUDA_OK(cudaMemcpyAsync(d_s, up.data(), (width + 4) * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream)); }
  SP2_TRY(eq_table_dev(ctx, (const fe *)d_r + l + 1 + nvr, (uint32_t)(m - nvr), (fe *)d_R));
  k_dot3<<<1, 256, 0, ctx->stream>>>((const fe *)d_s, (const fe *)d_s, (const fe *)d_s, (const fe *)d_R, width, sm + 4);
  SP2_LAUNCH_CHECK();
  std::vector<MsmJob> jobs(3);
  for (auto &j : jobs) memset(&j, 0, sizeof(j));
  jobs[0].nextra = 2; jobs[0].extra_base[0] = ck->idx_ck_s(); jobs[0].extra_scalar[0] = sm; jobs[0].extra_base[1] = ck->idx_h_s(); jobs[0].extra_scalar[1] = sm + 1;
  jobs[1].scalars = (const fe *)d_s; jobs[1].len = (u32)width; jobs[1].nextra = 1; jobs[1].extra_base[0] = ck->idx_h(); jobs[1].extra_scalar[0] = sm + 2;
  jobs[2].nextra = 2; jobs[2].extra_base[0] = ck->idx_ck_s(); jobs[2].extra_scalar[0] = sm + 4; jobs[2].extra_base[1] = ck->idx_h_s(); jobs[2].extra_scalar[1] = sm + 3;
  SP2_TRY(msm_run(ctx, ck, jobs, (jac *)d_pts));
  uint64_t hj[36], tab[24];
  SP2_CUDA_OK(cudaMemcpyAsync(hj, d_pts, 3 * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  sp2h::batch_normalize(hj, 3, tab);
  const uint64_t *comm_eval = tab, *rhs1 = tab + 8, *rhs2 = tab + 16;
  ts.dom_sep("inner product argument (linear)");
  ts.push("U", 1); ts.push_point(comm_LZ); ts.push_point(comm_eval);
  ts.absorb_point("delta", proof->delta); ts.absorb_point("beta", proof->beta);
  const fe r_ipa = squeeze(ts, "r");
  // lhs1 = r comm_LZ + delta ; lhs2 = r comm_eval + beta   (two 2-term variable-base MSMs)
  const fe sc[2] = {r_ipa, one};
  uint64_t b1[16], b2[16], lhs1[8], lhs2[8];
  memcpy(b1, comm_LZ, 64); memcpy(b1 + 8, proof->delta, 64); memcpy(b2, comm_eval, 64); memcpy(b2 + 8, proof->beta, 64);
  SP2_TRY(sp2_msm_var(ctx, (const uint64_t *)sc, b1, 2, lhs1));
  SP2_TRY(sp2_msm_var(ctx, (const uint64_t *)sc, b2, 2, lhs2));
  if (memcmp(lhs1, rhs1, 64) != 0 || memcmp(lhs2, rhs2, 64) != 0) return set_error(ctx, SP2_ERR_PROOF_VERIFY, "verify: the inner-product argument does not hold");
  return SP2_OK;
}

}  // extern "C"
