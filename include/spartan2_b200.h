/* spartan2_b200 — C ABI of the B200-native Spartan2 prover hot path.
 *
 * Drop-in boundary for microsoft/Spartan2 (reference paths are relative to the reference
 * repository).  The reference has no FFI; these are the entry points a patched fork binds at
 * the seams listed in SURVEY.md §8(b) / INTEGRATION.md.  Every function is `extern "C"`, takes
 * plain pointers and sizes, returns an int32 status (0 = OK, negative = error; the text is
 * available from sp2_last_error) and never unwinds.
 *
 * Data layout (identical to the reference's in-memory types, so Rust slices cross without
 * conversion):
 *   scalar  F : 4 x uint64 little-endian limbs, Montgomery form R = 2^256, canonical [0,p)
 *               = `MontgomeryLimbs::to_limbs()`           (src/big_num/montgomery.rs:17-22)
 *   point   G : 8 x uint64 = affine x limbs then y limbs in the T256 base field (Montgomery),
 *               identity encoded as all-zero             (from `to_coordinates()`,
 *                                                         src/provider/traits.rs:107-108,259-268)
 *   matrices  : CSR with uint32 column indices / row pointers (src/r1cs/sparse.rs:385-394;
 *               the classified form at :29-45 already uses u32)
 *
 * Host pointers unless a parameter is documented as a device handle.  Callee never retains
 * caller pointers past return.  Engine: T256HyraxEngine (src/provider/mod.rs:76-82).
 */
#ifndef SPARTAN2_B200_H
#define SPARTAN2_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SP2_OK 0
#define SP2_ERR_CUDA (-1)               /* SpartanError::InternalError{reason}                    */
#define SP2_ERR_INVALID_INPUT_LENGTH (-2) /* SpartanError::InvalidInputLength{reason}             */
#define SP2_ERR_INVALID_WITNESS_LENGTH (-3) /* SpartanError::InvalidWitnessLength                 */
#define SP2_ERR_INVALID_CK_LENGTH (-4)  /* SpartanError::InvalidCommitmentKeyLength               */
#define SP2_ERR_DIVISION_BY_ZERO (-5)   /* SpartanError::DivisionByZero                           */
#define SP2_ERR_INTERNAL (-6)           /* SpartanError::InternalError                            */
#define SP2_ERR_UNSUPPORTED (-7)        /* path not offloaded: caller must use its CPU path       */
#define SP2_ERR_PROOF_VERIFY (-8)       /* SpartanError::ProofVerifyError{reason}                 */

typedef struct sp2_ctx sp2_ctx;         /* one per (process, GPU)                                 */
typedef struct sp2_ck sp2_ck;           /* device-resident commitment key                         */
typedef struct sp2_shape sp2_shape;     /* device-resident SplitR1CSShape                         */
typedef struct sp2_prep sp2_prep;       /* device-resident SpartanPrepSNARK                       */

/* ---- context ------------------------------------------------------------------------------- */
int32_t sp2_ctx_create(int32_t device, sp2_ctx **out);
void sp2_ctx_destroy(sp2_ctx *ctx);
const char *sp2_last_error(const sp2_ctx *ctx);
uint64_t sp2_launch_count(const sp2_ctx *ctx);      /* kernels launched so far through ctx        */
int32_t sp2_synchronize(sp2_ctx *ctx);
int32_t sp2_num_sms(const sp2_ctx *ctx);
/* Launch on a caller-owned CUDA stream (cudaStream_t) instead of the context's own.             */
int32_t sp2_ctx_set_stream(sp2_ctx *ctx, void *cuda_stream);
/* CUDA-event stopwatch on the context's stream (device time between the two calls).             */
int32_t sp2_timer_start(sp2_ctx *ctx);
int32_t sp2_timer_stop(sp2_ctx *ctx, float *ms);

/* ---- sum-check (src/sumcheck.rs) ------------------------------------------------------------ */
/* Transcript hand-off: the Keccak256Transcript (src/provider/keccak.rs:26-31) right after a
 * squeeze is fully described by (round, state[64]); the whole-loop provers take that pair,
 * run every round's absorb(b"p")/squeeze(b"c") on the device and hand the pair back.           */
typedef struct { uint16_t round; uint8_t state[64]; } sp2_transcript_state;

/* Replaces SumcheckProof::prove_cubic_with_three_inputs (src/sumcheck.rs:502-571):
 * claim, taus[l]; tables A,B,C of 2^l scalars (consumed: contents undefined after the call).
 * Outputs: polys[l*4] full coefficients (low->high; the proof keeps [0],[2],[3]), r[l], claims[3].
 * SP2_ERR_UNSUPPORTED when tau_i * prod eq = 0 (the reference's fallback, :1327-1396).          */
int32_t sp2_sumcheck_cubic_prove(sp2_ctx *ctx, const uint64_t *claim, const uint64_t *taus, uint32_t l,
                                 const uint64_t *A, const uint64_t *B, const uint64_t *C,
                                 sp2_transcript_state *ts, uint64_t *polys, uint64_t *r, uint64_t *claims);

/* Replaces SumcheckProof::prove_quad (src/sumcheck.rs:190-247): tables of 2^rounds scalars.
 * Outputs: polys[rounds*3], r[rounds], claims[2].                                               */
int32_t sp2_sumcheck_quad_prove(sp2_ctx *ctx, const uint64_t *claim, uint32_t rounds,
                                const uint64_t *A, const uint64_t *B,
                                sp2_transcript_state *ts, uint64_t *polys, uint64_t *r, uint64_t *claims);

/* Device-resident variants (tables already in HBM; used by the fused prover and by bench.py's
 * kernel-only timing).  dA/dB/dC are device pointers to 2^l scalars, bound in place.           */
int32_t sp2_sumcheck_cubic_prove_dev(sp2_ctx *ctx, const uint64_t *claim, const uint64_t *taus, uint32_t l,
                                     void *dA, void *dB, void *dC,
                                     sp2_transcript_state *ts, uint64_t *polys, uint64_t *r, uint64_t *claims);
int32_t sp2_sumcheck_quad_prove_dev(sp2_ctx *ctx, const uint64_t *claim, uint32_t rounds, void *dA, void *dB,
                                    sp2_transcript_state *ts, uint64_t *polys, uint64_t *r, uint64_t *claims);

/* EqSumCheckInstance::evaluation_points_zero_check_round0 (src/sumcheck.rs:1163-1271; used by the ZK prover's first round, :593-596):
 * only t(inf) is summed (no Cz reads); out3 = (eval_0 = 0, eval_2, eval_3).  dA, dB: device tables of 2^l scalars, not modified. */
int32_t sp2_sc_zero_check_round0_dev(sp2_ctx *ctx, const uint64_t *taus, uint32_t l, const void *dA, const void *dB, uint64_t *out3);

/* ---- multi-GPU sharding of the sum-checks -------------------------------------------------
 * One process per GPU.  The 2^l hypercube is split cyclically on the low index bits (rank g owns entries
 * i = g mod nranks), so every bind pair is local; per round the <= 3 partial sums are exchanged by the round
 * kernel itself through CUDA-IPC-mapped peer mailboxes over NVLink (no NCCL call, no extra launch); below
 * 2^16 entries the shards are all-gathered and every rank finishes redundantly.  Setup: each rank creates a
 * comm, the 64-byte handles are all-gathered by the host (torch.distributed, MPI, ...), then connect.       */
typedef struct sp2_comm sp2_comm;
int32_t sp2_comm_create(sp2_ctx *ctx, int32_t rank, int32_t nranks, sp2_comm **out);
int32_t sp2_comm_handle(sp2_comm *comm, uint8_t *out64);
int32_t sp2_comm_connect(sp2_comm *comm, const uint8_t *all_handles /* nranks x 64 bytes, rank order */);
void sp2_comm_destroy(sp2_comm *comm);
/* Ranks living in ONE process (several contexts on one GPU or on peer-enabled GPUs) connect without CUDA IPC:
 * mailboxes[q] = rank q's sp2_comm_mailbox() device pointer.                                                         */
int32_t sp2_comm_mailbox(sp2_comm *comm, void **out);
int32_t sp2_comm_connect_ptrs(sp2_comm *comm, void *const *mailboxes /* nranks device pointers, rank order */);
/* Every device-side wait on a peer is bounded (2 s): when a rank fails or dies mid-call the others return
 * SP2_ERR_INTERNAL instead of wedging their GPU.  Afterwards every rank calls sp2_comm_reset (host barrier before and
 * after) to bring the mailbox flags / epochs back in step.                                                           */
int32_t sp2_comm_reset(sp2_comm *comm);
/* dA, dB, dC: this rank's shards (2^l / nranks entries each, device pointers), bound in place; every rank passes
 * the same claim / taus / transcript and receives identical outputs (same meaning as the single-GPU provers). */
int32_t sp2_sumcheck_cubic_prove_sharded_dev(sp2_ctx *ctx, sp2_comm *comm, const uint64_t *claim, const uint64_t *taus, uint32_t l,
                                             void *dA, void *dB, void *dC, sp2_transcript_state *ts, uint64_t *polys, uint64_t *r,
                                             uint64_t *claims);
int32_t sp2_sumcheck_quad_prove_sharded_dev(sp2_ctx *ctx, sp2_comm *comm, const uint64_t *claim, uint32_t rounds, void *dA, void *dB,
                                            sp2_transcript_state *ts, uint64_t *polys, uint64_t *r, uint64_t *claims);

/* ---- polynomials (src/polys) ---------------------------------------------------------------- */
/* EqPolynomial::evals_from_points (src/polys/eq.rs:59-117): out[2^k], MSB-first.               */
int32_t sp2_eq_table(sp2_ctx *ctx, const uint64_t *r, uint32_t k, uint64_t *out);
/* MultilinearPolynomial::bind_poly_var_top (src/polys/multilinear.rs:95-164): Z[len] -> out[len/2] */
int32_t sp2_bind_top(sp2_ctx *ctx, const uint64_t *Z, uint64_t len, const uint64_t *r, uint64_t *out);
/* device-resident variants (d_* are device pointers; asynchronous on the context's stream)      */
int32_t sp2_eq_table_dev(sp2_ctx *ctx, const void *d_r, uint32_t k, void *d_out);
int32_t sp2_bind_top_dev(sp2_ctx *ctx, const void *d_Z, uint64_t len, const void *d_r, void *d_out);

/* ---- R1CS (src/r1cs/sparse.rs, src/r1cs/mod.rs) ---------------------------------------------- */
/* Replaces SplitR1CSShape::precompute (src/r1cs/mod.rs:1059-1073): uploads the three padded CSR
 * matrices (data: Montgomery scalars; indices/indptr: u32, src/r1cs/sparse.rs:385-394) and builds
 * the device forms (dictionary-coded rows, transposes, filtered rows).  Column order of z:
 * [W_shared | W_precommitted | W_rest | 1 | public | challenges]  (src/r1cs/mod.rs:830-853).     */
int32_t sp2_shape_upload(sp2_ctx *ctx, uint64_t num_cons, uint64_t num_cons_unpadded, uint64_t num_shared, uint64_t num_precommitted,
                         uint64_t num_rest, uint64_t num_public, uint64_t num_challenges,
                         const uint64_t *dataA, const uint32_t *indicesA, const uint32_t *indptrA,
                         const uint64_t *dataB, const uint32_t *indicesB, const uint32_t *indptrB,
                         const uint64_t *dataC, const uint32_t *indicesC, const uint32_t *indptrC, sp2_shape **out);
/* One rank's shard of the shape for the multi-GPU prover: every rank passes the same whole CSR matrices;
 * rank g of nranks = 2^k keeps rows i = g (mod nranks) of A, B, C (so Az/Bz/Cz land directly in the cyclic sum-check
 * layout, local row i >> k) and columns j = g (mod nranks) of the transposes (poly_ABC lands in the inner sum-check's
 * cyclic layout).  With such a handle sp2_spmv3* produce num_cons / nranks rows and sp2_abc* this rank's columns.   */
int32_t sp2_shape_upload_sharded(sp2_ctx *ctx, int32_t rank, int32_t nranks, uint64_t num_cons, uint64_t num_cons_unpadded,
                                 uint64_t num_shared, uint64_t num_precommitted, uint64_t num_rest, uint64_t num_public,
                                 uint64_t num_challenges,
                                 const uint64_t *dataA, const uint32_t *indicesA, const uint32_t *indptrA,
                                 const uint64_t *dataB, const uint32_t *indicesB, const uint32_t *indptrB,
                                 const uint64_t *dataC, const uint32_t *indicesC, const uint32_t *indptrC, sp2_shape **out);
void sp2_shape_free(sp2_shape *shape);
/* pk.sizes() analogue: [num_cons, num_vars, num_cols, nnz, general nnz, long rows, long columns]  */
int32_t sp2_shape_sizes(const sp2_shape *shape, uint64_t *out7);
/* Replaces SplitR1CSShape::multiply_vec (src/r1cs/mod.rs:1075-1107): z[num_cols] -> az,bz,cz[num_cons].
 * InvalidWitnessLength when z_len != num_cols.                                                   */
int32_t sp2_spmv3(sp2_ctx *ctx, const sp2_shape *shape, const uint64_t *z, uint64_t z_len, uint64_t *az, uint64_t *bz, uint64_t *cz);
int32_t sp2_spmv3_dev(sp2_ctx *ctx, const sp2_shape *shape, const void *d_z, void *d_az, void *d_bz, void *d_cz);
/* Replaces multiply_vec_incremental_into (src/r1cs/mod.rs:1170-1211): cached + filtered columns.  */
int32_t sp2_spmv3_incremental(sp2_ctx *ctx, const sp2_shape *shape, const uint64_t *z, uint64_t z_len, const uint64_t *cached_az,
                              const uint64_t *cached_bz, const uint64_t *cached_cz, uint64_t *az, uint64_t *bz, uint64_t *cz);
int32_t sp2_spmv3_incremental_dev(sp2_ctx *ctx, const sp2_shape *shape, const void *d_z, const void *d_cached_az, const void *d_cached_bz,
                                  const void *d_cached_cz, void *d_az, void *d_bz, void *d_cz);
/* Replaces bind_and_prepare_poly_ABC(_full) (src/r1cs/mod.rs:1235-1321): rx = eq(r_x) table of
 * num_cons scalars, r the batching challenge; out[j] = sum_i (A + r B + r^2 C)[i,j] rx[i] for
 * j < num_cols, zero up to out_len.                                                              */
int32_t sp2_abc(sp2_ctx *ctx, const sp2_shape *shape, const uint64_t *rx, uint64_t rx_len, const uint64_t *r, uint64_t *out, uint64_t out_len);
int32_t sp2_abc_dev(sp2_ctx *ctx, const sp2_shape *shape, const void *d_rx, const void *d_r, void *d_out, uint64_t out_len);

/* ---- commitment provider (src/provider/msm.rs, src/provider/pcs/hyrax_pc.rs) ----------------- */
/* Replaces PCSEngineTrait::precompute_ck (src/traits/pcs.rs:56-58): uploads the Hyrax key
 * (n row bases, blinding base h; the 1-wide evaluation key ck_s, h_s — hyrax_pc.rs:152-190) and
 * precomputes the 2^(8w) multiples of every base (cf. FixedBaseMul::precompute, msm.rs:651-689). */
int32_t sp2_ck_upload(sp2_ctx *ctx, const uint64_t *bases_xy, uint32_t n, const uint64_t *h_xy, const uint64_t *ck_s_xy,
                      const uint64_t *h_s_xy, sp2_ck **out);
void sp2_ck_free(sp2_ck *ck);
/* Replaces DlogGroupExt::vartime_multiscalar_mul (src/provider/traits.rs:118-134 -> msm.rs:187-222):
 * out = sum_i scalars[i] * ck_i, affine (identity = all zero).  InvalidCommitmentKeyLength if n > key. */
int32_t sp2_msm(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *scalars, uint32_t n, uint64_t *out_xy);
/* ---- DlogGroupExt for ARBITRARY bases (src/provider/traits.rs:118-162): what a `GE: DlogGroupExt` wrapper forwards to ----
 * Signed-digit Pippenger (c = 8, msm.rs:59-222) on the device; points affine (x, y limbs; identity all zero) in and out.
 * vartime_multiscalar_mul(scalars, bases, _)                                                                            */
int32_t sp2_msm_var(sp2_ctx *ctx, const uint64_t *scalars, const uint64_t *bases_xy, uint32_t n, uint64_t *out_xy);
/* vartime_multiscalar_mul_small(scalars: u64, bases, _)  (msm.rs:367-620)                                               */
int32_t sp2_msm_small_var(sp2_ctx *ctx, const uint64_t *scalars_u64, const uint64_t *bases_xy, uint32_t n, uint64_t *out_xy);
/* batch_vartime_multiscalar_mul(scalars: &[Vec<_>], bases): k vectors (concatenated, lens[j] scalars each), each against
 * bases[..lens[j]]; out_xy: k points                                                                                    */
int32_t sp2_msm_batch_var(sp2_ctx *ctx, const uint64_t *scalars, const uint32_t *lens, uint32_t k, const uint64_t *bases_xy, uint64_t *out_xy);
/* vartime_multiscalar_mul_shared_weights(weights, bases_rows) (msm.rs:228-356): bases_rows_xy = rows x n points, row-major;
 * out_xy[r] = sum_i weights[i] * bases_rows[r][i]                                                                       */
int32_t sp2_msm_shared_weights(sp2_ctx *ctx, const uint64_t *weights, uint32_t n, const uint64_t *bases_rows_xy, uint32_t rows, uint64_t *out_xy);

/* Replaces HyraxPCS::commit / commit_zeros (hyrax_pc.rs:207-319): one Pedersen commitment per row of
 * ck-width scalars, out_rows[i] = <v_row_i, ck> + blinds[i] * h.  `is_small` is the reference's hint
 * (msm_small path); it does not change the result.                                               */
int32_t sp2_hyrax_commit(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *v, uint64_t len, const uint64_t *blinds, uint64_t rows,
                         int32_t is_small, uint64_t *out_rows);
/* HyraxPCS::commit_without_blind (hyrax_pc.rs:533-567): the raw row points (identity = all zero for an all-zero row)       */
int32_t sp2_hyrax_commit_without_blind(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *v, uint64_t len, int32_t is_small, uint64_t *out_rows);
/* HyraxPCS::commit_incremental (hyrax_pc.rs:569-607): out[i] = raw[i] (identity past n_raw) + <delta_row_i, ck> + blinds[i] h */
int32_t sp2_hyrax_commit_incremental(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *raw_rows_xy, uint64_t n_raw, const uint64_t *delta, uint64_t len,
                                     const uint64_t *blinds, uint64_t *out_rows);
/* HyraxPCS::rerandomize_commitment (hyrax_pc.rs:321-344): out[i] = comm[i] + (r_new[i] - r_old[i]) h                         */
int32_t sp2_hyrax_rerandomize(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *comm_rows_xy, const uint64_t *r_old, const uint64_t *r_new, uint64_t rows,
                              uint64_t *out_rows);
/* device-resident variant: d_out_rows_jac receives rows JACOBIAN points (x, y, z: 12 limbs; z = 0 identity) */
int32_t sp2_hyrax_commit_dev(sp2_ctx *ctx, const sp2_ck *ck, const void *d_v, uint64_t len, const void *d_blinds, uint64_t rows,
                             void *d_out_rows_jac);
/* Replaces bind_with_delayed (hyrax_pc.rs:38-54): out[i] = sum_j L[j] * poly[j * r_len + i].      */
int32_t sp2_hyrax_bind(sp2_ctx *ctx, const uint64_t *poly, const uint64_t *L, uint64_t rows, uint64_t r_len, uint64_t *out);

/* ---- SpartanSNARK (src/spartan.rs) ------------------------------------------------------------ */
/* Flat view of SpartanSNARK<E> (src/spartan.rs:126-139) over caller-owned buffers:
 *   comm_W        num_comm_rows points (HyraxCommitment rows: shared | precommitted | rest)
 *   outer_polys   num_rounds_x x 3 scalars  (CompressedUniPoly: c0, c2, c3 — univariate.rs:147-153)
 *   claims_outer  3 scalars (Az, Bz, Cz at r_x)
 *   inner_polys   num_rounds_y x 2 scalars  (c0, c2)
 *   eval_W, blind_eval_W; IPA argument: delta, beta (points), z_vec (num_cols scalars), z_delta, z_beta
 *   (HyraxEvaluationArgument / InnerProductArgumentLinear, hyrax_pc.rs:96-118, ipa.rs:104-121).           */
typedef struct {
  uint64_t num_rounds_x, num_rounds_y, num_comm_rows, num_cols;
  uint64_t *comm_W, *outer_polys, *claims_outer, *inner_polys, *eval_W, *blind_eval_W, *delta, *beta, *z_vec, *z_delta, *z_beta;
} sp2_spartan_proof;
/* The prover's randomness (the reference draws it from thread_rng inside prove: hyrax_pc.rs:192-205,
 * ipa.rs:140-145, bellpepper/r1cs.rs:467); the caller samples and passes it so runs are reproducible:
 * blinds_W: one per commitment row (rows of the shared / precommitted sections are IGNORED: those rows were committed
 * by sp2_spartan_prep_prove with `blinds_cached` and keep that blind — comm_W is reused, bellpepper/r1cs.rs:422-431; only
 * the rest rows are committed, with blinds_W[cached_rows..], inside prove, :467-470); d_vec: num_cols scalars.          */
typedef struct { const uint64_t *blinds_W, *blind_eval_W, *d_vec, *r_delta, *r_beta; } sp2_spartan_rand;

/* Replaces SpartanSNARK::prep_prove (src/spartan.rs:176-216): uploads the shared+precommitted witness
 * (num_shared + num_precommitted scalars), commits it row-wise (comm_out: rows x 8, may be NULL) and caches
 * its Az/Bz/Cz on the device.  The returned handle is the device half of SpartanPrepSNARK (:107-124).
 * Runtime: the handle owns one helper host thread (transcript hashing during sp2_spartan_prove; no CUDA calls, asleep between
 * proves, joined by sp2_prep_free), one side stream, and a page of pinned device-visible memory through which the host opens
 * the gate kernels of a prove ("runtime behaviour"); one prove at a time per handle.                        */
int32_t sp2_spartan_prep_prove(sp2_ctx *ctx, const sp2_shape *shape, const sp2_ck *ck, const uint64_t *W_cached,
                               const uint64_t *blinds_cached, int32_t is_small, uint64_t *comm_out, sp2_prep **out);
void sp2_prep_free(sp2_prep *prep);
/* Replaces SpartanSNARK::prove (src/spartan.rs:219-466).  W_rest: num_rest scalars, or NULL for an all-zero
 * rest section (pure padding, as in the SHA-256 bench circuit whose `synthesize` allocates nothing).
 * phase_ms: optional 8 floats of device time (commit+transcript, matrix_vector_multiply, outer_sumcheck,
 * prepare_poly_ABC, inner_sumcheck, pcs_prove, ipa response, total).  DivisionByZero as spartan.rs:417.  */
int32_t sp2_spartan_prove(sp2_ctx *ctx, const sp2_shape *shape, const sp2_ck *ck, sp2_prep *prep, const uint8_t *vk_digest,
                          const uint64_t *public_values, const uint64_t *W_rest, const sp2_spartan_rand *rand,
                          sp2_spartan_proof *proof, float *phase_ms);

/* Replaces SpartanSNARK::verify (src/spartan.rs:469-578): the matrix evaluations (evaluate_with_tables_fast, src/r1cs/mod.rs:36-146,
 * 1216-1226), the Hyrax row MSM (hyrax_pc.rs:480-531) and the IPA checks (ipa.rs:173-221) run on the device; SP2_OK = accept,
 * SP2_ERR_PROOF_VERIFY = reject (sp2_last_error names the failing check).                                                  */
int32_t sp2_spartan_verify(sp2_ctx *ctx, const sp2_shape *shape, const sp2_ck *ck, const uint8_t *vk_digest, const uint64_t *public_values,
                           const sp2_spartan_proof *proof);

/* SpartanSNARK::prove with the 2^l hypercube split across the GPUs of `comm` (one process per GPU).  Every
 * rank calls with the same inputs, its shard of the shape (sp2_shape_upload_sharded) and the prep state made from it,
 * and receives the identical proof.  Sharded: Az/Bz/Cz rows, both sum-checks (per round the <= 3 partial sums cross
 * NVLink inside the round kernel), poly_ABC columns.  Replicated: witness, commitments, eq(r_x), Hyrax bind, PCS MSMs. */
int32_t sp2_spartan_prove_sharded(sp2_ctx *ctx, sp2_comm *comm, const sp2_shape *shape, const sp2_ck *ck, sp2_prep *prep,
                                  const uint8_t *vk_digest, const uint64_t *public_values, const uint64_t *W_rest,
                                  const sp2_spartan_rand *rand, sp2_spartan_proof *proof, float *phase_ms);

/* ---- NeutronNova building blocks (src/neutronnova_zk.rs, src/polys/power.rs, src/r1cs/mod.rs) ------------
 * The ZK drivers draw one challenge per round from the in-circuit verifier (process_round, out of scope), so the
 * seams are per-round {evaluate, fold/bind} pairs on device-resident tables rather than whole loops.            */
/* PowPolynomial::split_evals (src/polys/power.rs:65-86): [1,t,..,t^(left-1)] || [1,t^left,..] (host out)       */
int32_t sp2_pow_split_evals(sp2_ctx *ctx, const uint64_t *t, uint32_t left, uint32_t right, uint64_t *out);
/* One NIFS round (src/neutronnova_zk.rs:98-178 prove_helper per pair, :78-87 suffix weights, summed over pairs):
 * m live layers of N = left*right entries, layer q at slot q*stride of dA/dB/dC; out2 = (e0, quad_coeff).       */
int32_t sp2_nifs_round_dev(sp2_ctx *ctx, uint32_t t, const uint64_t *rhos, uint32_t ell_b, uint32_t left, uint32_t right, const void *dE,
                           const void *dA, const void *dB, const void *dC, uint64_t N, uint64_t m, uint64_t stride, uint64_t *out2);
/* fold_abc_pair for every pair (src/neutronnova_zk.rs:738-776): slot 2p*stride <- lo + r_b (hi - lo), in place   */
int32_t sp2_nifs_fold_dev(sp2_ctx *ctx, void *dA, void *dB, void *dC, uint64_t N, uint64_t m, uint64_t stride, const uint64_t *r_b);
/* weights_from_r (src/r1cs/mod.rs:153-166), LSB-first                                                            */
int32_t sp2_weights_from_r(sp2_ctx *ctx, const uint64_t *r_bs, uint32_t ell, uint32_t n, uint64_t *out);
/* R1CSWitness::fold_multiple, W part (src/r1cs/mod.rs:570-660): d_out[j] = sum_i w[i] * dWs[i*dim + j]            */
int32_t sp2_fold_vectors_dev(sp2_ctx *ctx, const void *dWs, uint64_t n, uint64_t dim, const uint64_t *w, void *d_out);
/* compute_eval_points_cubic_with_additive_term(_with_outer_pow) (src/sumcheck.rs:262-342, 366-498): out3 =
 * evaluations at 0, 2, 3, unscaled by base_tau; tables of table_len entries                                      */
int32_t sp2_sc_pow_cubic_eval_dev(sp2_ctx *ctx, const void *d_pow_left, uint32_t left, const void *d_pow_right, const void *dA, const void *dB,
                                  const void *dC, uint64_t table_len, uint64_t *out3);
/* compute_eval_points_quad (src/sumcheck.rs:128-174): out2 = (eval_point_0, bound_coeff)                         */
int32_t sp2_sc_quad_eval_dev(sp2_ctx *ctx, const void *dA, const void *dB, uint64_t table_len, uint64_t *out2);
/* bind_poly_var_top on up to 8 device tables with one challenge, in place (the rayon::join bind fan-out of
 * src/sumcheck.rs:880-903)                                                                                        */
int32_t sp2_bind_tables_dev(sp2_ctx *ctx, void *const *d_tables, uint32_t ntables, uint64_t table_len, const uint64_t *r);
/* HyraxPCS::fold_commitments (src/provider/pcs/hyrax_pc.rs:737-793 -> msm_shared_weights, msm.rs:228-356):
 * out[row] = sum_i w[i] * comms[i*rows + row], affine in/out                                                      */
int32_t sp2_fold_commitments(sp2_ctx *ctx, const uint64_t *comms_xy, uint32_t n, uint32_t rows, const uint64_t *w, uint64_t *out_xy);
/* HyraxPCS::fold_blinds (hyrax_pc.rs:795-819): out[row] = sum_k w[k] * blinds[k*rows + row]                                        */
int32_t sp2_fold_blinds(sp2_ctx *ctx, const uint64_t *blinds, uint32_t n, uint32_t rows, const uint64_t *w, uint64_t *out);
/* HyraxPCS::fold_commitments_partial (hyrax_pc.rs:821-874): data rows folded as group elements, rest rows = folded_blind[row] h   */
int32_t sp2_fold_commitments_partial(sp2_ctx *ctx, const sp2_ck *ck, const uint64_t *comms_xy, uint32_t n, uint32_t rows, const uint64_t *w,
                                     uint32_t num_data_rows, const uint64_t *folded_blind, uint64_t *out_xy);

/* ---- small-value path (src/big_num/small_value.rs and its users in src/neutronnova_zk.rs) ---------------------------
 * to_small_vec_or_zero (small_value.rs:42-85) for up to 4 device tables of n_layers x N scalars (the Az, Bz, Cz layers of
 * prep_prove, neutronnova_zk.rs:1551-1584): values with |v| <= 2^62 - 1 become int64, all others 0; the UNION of the
 * large positions over all layers of all tables is zeroed in every int64 table and returned ascending in d_positions
 * (N uint64, device); *n_large = their number.                                                                       */
int32_t sp2_to_small_layers_dev(sp2_ctx *ctx, const void *const *d_tables, void *const *d_i64, uint32_t ntables, uint64_t n_layers, uint64_t N,
                                void *d_positions, uint64_t *n_large);
/* NIFS round 0 on the i64 layers: prove_helper_small (neutronnova_zk.rs:255-325: i64 differences, i128 product,
 * SmallAccumulator small_value.rs:96-196, reduce_7_to_field :204-222, field correction at the large positions) per
 * pair, suffix weights, summed over pairs (:781-810).  out2 = (0, quad_coeff).                                       */
int32_t sp2_nifs_round0_small_dev(sp2_ctx *ctx, const uint64_t *rhos, uint32_t ell_b, uint32_t left, uint32_t right, const void *dE, const void *dA64,
                                  const void *dB64, const void *dA, const void *dB, const void *d_positions, uint64_t n_large, uint64_t N, uint64_t m,
                                  uint64_t *out2);
/* c_vals[b] = sum_k E[k] * Cz_b[k] from the i64 C layers (neutronnova_zk.rs:649-693), n scalars to the host           */
int32_t sp2_nifs_cvals_small_dev(sp2_ctx *ctx, uint32_t left, uint32_t right, const void *dE, const void *dC, const void *dC64, const void *d_positions,
                                 uint64_t n_large, uint64_t N, uint64_t n, uint64_t *out_vals);

/* ---- fused NeutronNova hot path (BASELINE configs 3 / 5) ----------------------------------------------------------
 * NeutronNovaZkSNARK::{prep_prove, prove} (src/neutronnova_zk.rs:1477-1603, 1609-2093), data path: the per-step
 * Az/Bz/Cz of prep_prove, then HOT LOOP A (NeutronNovaNIFS::prove :511-1273 + R1CSWitness::fold_multiple,
 * src/r1cs/mod.rs:570-660), HOT LOOP B (prove_cubic_with_additive_term_batched, two branches, src/sumcheck.rs:786-917),
 * bind_and_prepare_poly_ABC_full x 2 (:1868-1875) and HOT LOOP C (prove_quad_batched, src/sumcheck.rs:702-782), with all
 * tables device-resident and the per-round scalar algebra + Keccak transcript on the host inside the library (one
 * host-mapped flag round trip per round).  Challenges are transcript-derived (absorb b"p" / squeeze b"c"); the
 * reference's ZK wrapper draws them from its in-circuit verifier (process_round), which is out of scope — a fork
 * that keeps the ZK wrapper uses the per-round seams above instead.  Every scalar below is 4 x u64 Montgomery limbs;
 * the caller allocates all arrays (sizes in scalars). */
typedef struct sp2_nn_prep sp2_nn_prep;
typedef struct sp2_transcript sp2_transcript;
typedef struct {
  uint32_t n_steps, ell_b, ell, rounds_y;   /* out: instances, log2(instances), log2(num_cons), log2(2 * num_vars) */
  uint64_t *nifs_evals;    /* ell_b x 2: (e0, quad_coeff) per round (prove_helper sums, :98-178)                  */
  uint64_t *nifs_polys;    /* ell_b x 4: finish_round! coefficients (:703-735)                                    */
  uint64_t *r_b;           /* ell_b                                                                                */
  uint64_t *T_out;         /* 1: T_cur / acc_eq (:1206-1208)                                                      */
  uint64_t *outer_evals;   /* ell x 6: evaluations at (0, 2, 3) of the step and the core branch, unscaled         */
  uint64_t *outer_polys;   /* ell x 8: the two cubic round polynomials                                            */
  uint64_t *r_x;           /* ell                                                                                  */
  uint64_t *claims_outer;  /* 6: Az, Bz, Cz (step), Az, Bz, Cz (core) at r_x                                      */
  uint64_t *tau_at_rx;     /* 1                                                                                    */
  uint64_t *r;             /* 1: batching challenge                                                               */
  uint64_t *inner_evals;   /* rounds_y x 4: (eval_0, bound_coeff) per branch                                      */
  uint64_t *inner_polys;   /* rounds_y x 6: the two quadratic round polynomials                                   */
  uint64_t *r_y;           /* rounds_y                                                                             */
  uint64_t *inner_final;   /* 4: poly_ABC (step, core), z (step, core) at r_y                                     */
  uint64_t *eval_W;        /* 2                                                                                    */
  uint64_t *heads;         /* optional (may be NULL), 28: first 4 entries of the folded Az, Bz, Cz; first 8 of the
                              folded witness; first 8 of poly_ABC (step) — parity probes                          */
  int32_t outer_ok, inner_ok;  /* the verifier's final equations of the two sum-checks hold                        */
} sp2_nn_proof;
/* step_zs: n_steps x num_cols scalars (z_i = [W_i | 1 | X_i]); core_z: num_cols.  n_steps: power of two in [2, 256]. */
int32_t sp2_neutronnova_prep_prove(sp2_ctx *ctx, const sp2_shape *shape, uint32_t n_steps, const uint64_t *step_zs, const uint64_t *core_z,
                                   sp2_nn_prep **out);
/* phase_ms (optional, 6 floats): nifs, fold_witness, outer_sumcheck_batched, compute_eval_table_sparse,
 * inner_sumcheck_batched, total (host wall clock; every phase ends in a host wait)                                  */
int32_t sp2_neutronnova_prove(sp2_ctx *ctx, sp2_nn_prep *prep, sp2_transcript *ts, sp2_nn_proof *proof, float *phase_ms);
void sp2_neutronnova_prep_free(sp2_nn_prep *prep);
/* Multi-GPU NeutronNova (one process per GPU): the step instances are the independent units — rank g of
 * nranks (a power of two) holds n_local instances [g * n_local, (g+1) * n_local); the NIFS rounds fold adjacent pairs
 * locally for log2(n_local) rounds (per round the two sums cross ranks inside the publish kernel through `comm`'s peer
 * mailboxes — sp2_comm_create/_handle/_connect; comm = NULL: one 64-byte host all-gather per round), then the surviving
 * layer triples are all-gathered and the remaining log2(nranks) rounds, the sum-checks and poly_ABC run replicated;
 * the witness fold is a local partial + all-gather + sum.  `allgather` is the host's collective (recv = contributions in
 * rank order; on_device = 1: both pointers are device memory, complete on return).  Every rank gets the identical proof. */
typedef int32_t (*sp2_allgather_fn)(void *user, const void *send, uint64_t bytes, void *recv, int32_t on_device);
int32_t sp2_neutronnova_prep_prove_sharded(sp2_ctx *ctx, const sp2_shape *shape, int32_t rank, int32_t nranks, uint32_t n_local,
                                           const uint64_t *local_step_zs, const uint64_t *core_z, sp2_nn_prep **out);
/* optional: map every rank's exchange buffer through CUDA IPC (handle = 64 bytes, all-gathered by the host in rank
 * order) so that the two bulk exchanges of a sharded prove become peer stores over NVLink (needs `comm` for the flag
 * barriers); without it they go through the `allgather` callback                                                     */
int32_t sp2_neutronnova_prep_ipc_handle(sp2_nn_prep *prep, uint8_t *out64);
int32_t sp2_neutronnova_prep_connect(sp2_nn_prep *prep, const uint8_t *all_handles);
/* in-process variant (ranks in one process, no CUDA IPC): xbufs[q] = rank q's sp2_neutronnova_prep_xbuf() pointer */
int32_t sp2_neutronnova_prep_xbuf(sp2_nn_prep *prep, void **out);
int32_t sp2_neutronnova_prep_connect_ptrs(sp2_nn_prep *prep, void *const *xbufs);
int32_t sp2_neutronnova_prove_sharded(sp2_ctx *ctx, sp2_nn_prep *prep, sp2_transcript *ts, sp2_comm *comm, sp2_allgather_fn allgather, void *user,
                                      sp2_nn_proof *proof, float *phase_ms);

/* ---- NeutronNova prove incl. its commitment half (non-ZK variant; checker: oracle/oracle.c orc_neutronnova_prove/_verify) ----
 * NeutronNovaZkSNARK::{prep_prove, prove} (src/neutronnova_zk.rs:1477-1603, 1609-2093) with the ZK wrapper removed: per-step
 * rerandomize_commitment (hyrax_pc.rs:321-344) and commit_zeros of the rest rows (:305-319), the transcript over all
 * instances (b"vk", b"core_instance", b"U"...), HOT LOOPS A-C as sp2_neutronnova_prove, fold_blinds + fold_commitments_partial
 * (:795-874), commitments to eval_W_step / eval_W_core (absorbed as b"comm_eval_W_step" / b"comm_eval_W_core"), b"c_eval",
 * the [1, c_eval] folds of commitments / blinds / witnesses (neutronnova_zk.rs:2019-2051) and PCS::prove on the folded witness
 * (:2053-2064).  Round polynomials are absorbed directly (b"p", all coefficients) instead of being committed inside the
 * reference's in-circuit verifier; eval_W_{step,core} are revealed with their blinds, as SpartanSNARK reveals eval_W.
 * Requires num_shared == 0 (the SHA-256 chain).  All arrays are caller-allocated.                                      */
typedef struct {
  sp2_nn_proof base;            /* nifs_polys, outer_polys, claims_outer, inner_polys, eval_W (+ the evals / challenges probes)     */
  uint64_t rows;                /* out: commitment rows per instance = num_vars / ck width                                         */
  uint64_t *comm_W_steps;       /* n_steps x rows points: U_i.comm_W (precommitted rows rerandomised, rest rows = blind * h)        */
  uint64_t *comm_W_core;        /* rows points                                                                                       */
  uint64_t *blind_eval_W;       /* 2                                                                                                 */
  uint64_t *delta, *beta;       /* 1 point each (InnerProductArgumentLinear, ipa.rs:104-121)                                         */
  uint64_t *z_vec;              /* ck width scalars                                                                                  */
  uint64_t *z_delta, *z_beta;   /* 1 each                                                                                            */
  uint64_t *comm_eval_W;        /* optional (may be NULL), 2 points: parity probes — the verifier recomputes them                    */
  uint64_t *c_eval;             /* optional, 1                                                                                       */
  uint64_t *comm_fold;          /* optional, rows points: fold([folded_U.comm_W, core.comm_W], [1, c_eval])                          */
} sp2_nn_snark;
/* the prover's randomness: blinds_steps n_steps x rows / blinds_core rows = the blinds every instance commitment ends up with
 * (rerandomised precommitted rows and fresh rest rows, bellpepper/r1cs.rs:467, 568-602); blind_eval_W: 2; d_vec: ck width   */
typedef struct { const uint64_t *blinds_steps, *blinds_core, *blind_eval_W, *d_vec, *r_delta, *r_beta; } sp2_nn_rand;
/* prep_prove's commitment half on an sp2_neutronnova_prep_prove state: commits the precommitted section of every instance
 * (blinds_pre_steps: n_steps x pre_rows, blinds_pre_core: pre_rows; outputs may be NULL) and caches the unblinded rows
 * (HyraxPCS::commit_without_blind, hyrax_pc.rs:533-567) that make every later commitment operation fixed-base.             */
int32_t sp2_neutronnova_prep_commit(sp2_ctx *ctx, sp2_nn_prep *prep, const sp2_ck *ck, const uint64_t *blinds_pre_steps,
                                    const uint64_t *blinds_pre_core, uint64_t *comm_pre_steps_out, uint64_t *comm_pre_core_out);
/* phase_ms (optional, 10 floats, host wall clock): rerandomize + commit_zeros, instance transcript, nifs, fold_witness,
 * outer_sumcheck_batched, compute_eval_table_sparse, inner_sumcheck_batched, eval commitments + c_eval, pcs_prove, total    */
int32_t sp2_neutronnova_snark_prove(sp2_ctx *ctx, sp2_nn_prep *prep, const uint8_t *vk_digest, const sp2_nn_rand *rand, sp2_nn_snark *snark,
                                    float *phase_ms);

/* instance-sharded variant (one process per GPU, prep state from sp2_neutronnova_prep_prove_sharded + sp2_neutronnova_prep_commit with
 * the LOCAL instances' blinds): every rank passes the same `rand` (blinds of all n_local * nranks instances) and gets the identical
 * proof; a rank rerandomises only its own instances, the rows are all-gathered through `allgather` (host buffers).            */
int32_t sp2_neutronnova_snark_prove_sharded(sp2_ctx *ctx, sp2_nn_prep *prep, sp2_comm *comm, sp2_allgather_fn allgather, void *user,
                                            const uint8_t *vk_digest, const sp2_nn_rand *rand, sp2_nn_snark *snark, float *phase_ms);

/* ---- host Keccak256Transcript (src/provider/keccak.rs:18-105 behind TranscriptEngineTrait, src/traits/transcript.rs) ----
 * Host-side Fiat-Shamir for drivers that interleave per-round device calls with transcript steps (a Rust caller keeps
 * using its own Keccak256Transcript; this is the same object for C/C++/Python hosts).  Scalars are Montgomery limbs. */
int32_t sp2_transcript_new(const char *label, sp2_transcript **out);
void sp2_transcript_free(sp2_transcript *t);
int32_t sp2_transcript_absorb_bytes(sp2_transcript *t, const char *label, const uint8_t *data, uint64_t n);
int32_t sp2_transcript_absorb_scalars(sp2_transcript *t, const char *label, const uint64_t *scalars, uint64_t n);
int32_t sp2_transcript_absorb_commitment(sp2_transcript *t, const char *label, const uint64_t *rows_xy, uint64_t rows);
int32_t sp2_transcript_dom_sep(sp2_transcript *t, const char *label);
int32_t sp2_transcript_squeeze(sp2_transcript *t, const char *label, uint64_t *out_scalar);
int32_t sp2_transcript_get_state(const sp2_transcript *t, sp2_transcript_state *out);

/* measurement hook: device time of the last NIFS round-0 kernel (the pass over the i64 layers) of a NeutronNova prove          */
int32_t sp2_neutronnova_last_round0_ms(sp2_nn_prep *prep, float *ms);
/* measurement hook: duration in ms (CUDA events on the library's stream) of the most recent persistent cubic
 * sum-check kernel (all multi-CTA rounds of prove_cubic_with_three_inputs in one launch)                              */
int32_t sp2_last_cubic_persist_ms(sp2_ctx *ctx, float *ms);
/* table length at or below which the single-CTA tail kernels take over from the persistent multi-CTA kernels          */
uint64_t sp2_sc_tail_len(void);
/* table length (going into a round >= 2) at or below which the persistent kernels hand over to the pipelined multi-CTA
 * kernels (k_*_mid_pipe); 0 when they are disabled (SP2_MID_PIPE=0 / SP2_TAIL_PIPE=0)                                   */
uint64_t sp2_sc_mid_len(void);

/* ---- raw device memory helpers (for device-resident callers) ------------------------------- */
int32_t sp2_dev_alloc(sp2_ctx *ctx, uint64_t bytes, void **out);
int32_t sp2_dev_free(sp2_ctx *ctx, void *p);
int32_t sp2_dev_upload(sp2_ctx *ctx, void *dst, const void *src, uint64_t bytes);
int32_t sp2_dev_download(sp2_ctx *ctx, void *dst, const void *src, uint64_t bytes);
int32_t sp2_dev_copy(sp2_ctx *ctx, void *dst, const void *src, uint64_t bytes);
int32_t sp2_dev_memset(sp2_ctx *ctx, void *dst, int32_t value, uint64_t bytes);
/* page-locked host buffers (so uploads/downloads through this ABI are true async DMA)           */
int32_t sp2_host_alloc(sp2_ctx *ctx, uint64_t bytes, void **out);
int32_t sp2_host_free(sp2_ctx *ctx, void *p);

/* debug: 7 clock64() stamps (SM cycles) of the last finalised sum-check round, then 4 %globaltimer
 * values (ns) of the last multi-CTA cubic round: [-, election, finalize end, first CTA entry], then the
 * clock64() start stamp of the previous tail round (12 values)                                    */
int32_t sp2_debug_sc_clocks(sp2_ctx *ctx, uint64_t *out11);
/* debug: rounds x 4 %globaltimer stamps (ns) of the last persistent cubic kernel (CTA 0): start, own compute done,
 * all CTAs arrived, finalised                                                                      */
int32_t sp2_debug_sc_round_profile(sp2_ctx *ctx, uint64_t *out, uint32_t rounds);

/* harness support: n pseudo-random T256 points (seeded multiples of the generator) for test/bench keys.
 * (The reference derives its generators by hash-to-curve on the Rust side and ships them as bases.)        */
int32_t sp2_test_points(sp2_ctx *ctx, uint64_t seed, uint32_t n, uint64_t *out_xy);

/* ---- test hooks: the device field layer, element-wise (tests/test_gpu_field.py) ------------- */
/* op: 0 mul, 1 add, 2 sub, 3 inv, 4 from_mont, 5 to_mont, 6 half; field: 0 = T256 scalar, 1 = T256 base */
int32_t sp2_test_field_op(sp2_ctx *ctx, int32_t field, int32_t op, const uint64_t *a, const uint64_t *b,
                          uint64_t *out, uint64_t n);
/* sum_i a_i*b_i through the 544-bit delayed-reduction accumulator (big_num/delayed_reduction.rs) */
int32_t sp2_test_dot_delayed(sp2_ctx *ctx, const uint64_t *a, const uint64_t *b, uint64_t n, uint64_t *out);
/* Keccak256Transcript on the device: absorb `data` under `label`, squeeze under `sq_label`.     */
int32_t sp2_test_transcript(sp2_ctx *ctx, sp2_transcript_state *ts, const uint8_t *pending, uint32_t pending_len,
                            const char *sq_label, uint64_t *challenge_out);

#ifdef __cplusplus
}
#endif
#endif
