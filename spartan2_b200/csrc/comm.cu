// Multi-GPU plumbing for the sharded sum-checks: one process per GPU, mailboxes exchanged as CUDA IPC handles
// (the bytes travel through whatever the host uses — torch.distributed in bench/tests), peers written with plain
// stores over NVLink from inside the round kernels (sumcheck.cu: exchange_sums, k_shard_gather).  SURVEY.md §8(e).
#include <string.h>
#include "ctx.cuh"
#include "sumcheck.cuh"

using namespace sp2;

namespace sp2 {
// after the calling function has synchronised the stream: did a bounded peer wait of this rank expire?
int comm_check(sp2_ctx *ctx, sp2_comm *c) {
  if (!c || c->dc.n <= 1) return SP2_OK;
  u32 e = 0;
  SP2_CUDA_OK(cudaMemcpyAsync(&e, &c->dc.peer[c->dc.rank]->err, sizeof(u32), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (e) return set_error(ctx, SP2_ERR_INTERNAL, "sharded sum-check: a peer did not publish its round sums within 2 s (call sp2_comm_reset on every rank)");
  return SP2_OK;
}
}  // namespace sp2

extern "C" {

int32_t sp2_comm_create(sp2_ctx *ctx, int32_t rank, int32_t nranks, sp2_comm **out) {
  cudaSetDevice(ctx->device);
  if (!out) return SP2_ERR_INTERNAL;
  *out = nullptr;
  int k = 0; while ((1 << k) < nranks) k++;
  if (nranks < 1 || nranks > SC_MAX_RANKS || (1 << k) != nranks || rank < 0 || rank >= nranks)
    return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "comm: the number of ranks must be 1, 2, 4 or 8");
  sp2_comm *c = new sp2_comm();
  c->ctx = ctx;
  memset(&c->dc, 0, sizeof(c->dc));
  c->dc.rank = rank; c->dc.n = nranks; c->dc.k = k; c->dc.epoch = 0;
  MailBox *mb = nullptr;
  cudaError_t e = cudaMalloc((void **)&mb, sizeof(MailBox));
  if (e == cudaSuccess) e = cudaMemset(mb, 0, sizeof(MailBox));
  if (e != cudaSuccess) { delete c; return set_cuda_error(ctx, e, "comm alloc", __LINE__); }
  c->dc.peer[rank] = mb;
  c->connected = nranks == 1;
  *out = c;
  return SP2_OK;
}

/* the 64-byte cudaIpcMemHandle of this rank's mailbox (to be all-gathered by the host) */
int32_t sp2_comm_handle(sp2_comm *c, uint8_t *out64) {
  sp2_ctx *ctx = c->ctx;
  cudaSetDevice(ctx->device);
  cudaIpcMemHandle_t h;
  SP2_CUDA_OK(cudaIpcGetMemHandle(&h, c->dc.peer[c->dc.rank]));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out64, &h, 64);
  return SP2_OK;
}

/* all_handles: nranks x 64 bytes in rank order */
int32_t sp2_comm_connect(sp2_comm *c, const uint8_t *all_handles) {
  sp2_ctx *ctx = c->ctx;
  cudaSetDevice(ctx->device);
  for (int q = 0; q < c->dc.n; q++) {
    if (q == c->dc.rank) continue;
    cudaIpcMemHandle_t h; memcpy(&h, all_handles + 64 * q, 64);
    void *p = nullptr;
    SP2_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->dc.peer[q] = (MailBox *)p; c->opened[q] = true;
  }
  c->connected = true;
  return SP2_OK;
}

/* In-process variant of connect for ranks that live in ONE process (several contexts on one GPU, or on peer-enabled
 * GPUs): mailboxes[q] = rank q's sp2_comm_mailbox() pointer, no CUDA IPC involved.  Used by the single-GPU multi-rank
 * tests of the NVLink mailbox protocol (tests/test_gpu_multirank.py). */
int32_t sp2_comm_mailbox(sp2_comm *c, void **out) { *out = c->dc.peer[c->dc.rank]; return SP2_OK; }
int32_t sp2_comm_connect_ptrs(sp2_comm *c, void *const *mailboxes) {
  for (int q = 0; q < c->dc.n; q++) if (q != c->dc.rank) c->dc.peer[q] = (MailBox *)mailboxes[q];
  c->in_process = true;
  c->connected = true;
  return SP2_OK;
}

/* Collective re-initialisation after a failed sharded call (a rank returned an error between the epoch bump and its
 * kernels, or a bounded device wait expired): every rank calls it, with a host barrier before and after, and the
 * mailbox flags / epoch start from zero again. */
int32_t sp2_comm_reset(sp2_comm *c) {
  sp2_ctx *ctx = c->ctx;
  cudaSetDevice(ctx->device);
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  MailBox *mb = c->dc.peer[c->dc.rank];
  SP2_CUDA_OK(cudaMemsetAsync(mb->flag, 0, sizeof(mb->flag), ctx->stream));
  SP2_CUDA_OK(cudaMemsetAsync(&mb->err, 0, sizeof(u32), ctx->stream));
  SP2_CUDA_OK(cudaMemsetAsync(&mb->late_flag, 0, sizeof(u32), ctx->stream));
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));
  c->dc.epoch = 0;
  return SP2_OK;
}

void sp2_comm_destroy(sp2_comm *c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  for (int q = 0; q < c->dc.n; q++) if (c->opened[q]) cudaIpcCloseMemHandle(c->dc.peer[q]);
  cudaFree(c->dc.peer[c->dc.rank]);
  delete c;
}

/* prove_cubic_with_three_inputs over tables sharded cyclically across the ranks of `comm`: dA, dB, dC are this rank's
 * shards (entries i = rank mod nranks of the global 2^l tables, 2^l / nranks each), bound in place.  Every rank passes
 * the same claim / taus / transcript and receives identical outputs. */
int32_t sp2_sumcheck_cubic_prove_sharded_dev(sp2_ctx *ctx, sp2_comm *c, const uint64_t *claim, const uint64_t *taus, uint32_t l,
                                             void *dA, void *dB, void *dC, sp2_transcript_state *ts, uint64_t *polys, uint64_t *r,
                                             uint64_t *claims) {
  cudaSetDevice(ctx->device);
  if (!c->connected) return set_error(ctx, SP2_ERR_INTERNAL, "comm: not connected");
  if (l < 1 || l > SC_MAX_ROUNDS) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "sumcheck_cubic: 1 <= num_rounds <= 40");
  ScState *st;
  SP2_TRY(sc_state_upload(ctx, &st, claim, taus, l, ts));
  c->dc.epoch++;                                    // only once the fallible host steps in front of the kernels are done
  SP2_TRY(sumcheck_cubic_enqueue(ctx, st, l, (fe *)dA, (fe *)dB, (fe *)dC, &c->dc));
  const int rc = sc_state_download(ctx, st, ts, polys, 4, r, claims, 3, l);
  SP2_TRY(comm_check(ctx, c));
  return rc;
}

int32_t sp2_sumcheck_quad_prove_sharded_dev(sp2_ctx *ctx, sp2_comm *c, const uint64_t *claim, uint32_t rounds, void *dA, void *dB,
                                            sp2_transcript_state *ts, uint64_t *polys, uint64_t *r, uint64_t *claims) {
  cudaSetDevice(ctx->device);
  if (!c->connected) return set_error(ctx, SP2_ERR_INTERNAL, "comm: not connected");
  if (rounds < 1 || rounds > SC_MAX_ROUNDS) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "sumcheck_quad: 1 <= num_rounds <= 40");
  ScState *st;
  SP2_TRY(sc_state_upload(ctx, &st, claim, nullptr, rounds, ts));
  c->dc.epoch++;
  SP2_TRY(sumcheck_quad_enqueue(ctx, st, rounds, (fe *)dA, (fe *)dB, ~0ull, nullptr, &c->dc));
  const int rc = sc_state_download(ctx, st, ts, polys, 3, r, claims, 2, rounds);
  SP2_TRY(comm_check(ctx, c));
  return rc;
}

}  // extern "C"
