// Device-resident sum-check state shared by sumcheck.cu and the fused prover (prover.cu).
#pragma once
#include "ctx.cuh"
#include "keccak.cuh"

namespace sp2 {

constexpr int SC_MAX_ROUNDS = 40;
constexpr int SC_MAX_BLOCKS = 2048;
constexpr int SC_DERIVE_MAX = 8;
constexpr int SC_THREADS = 256;
constexpr int SC_TAIL_THREADS = 384;   // 12 warps = 4 role trios
constexpr int SC_ROLE_THREADS = 384;
// thresholds swept on B200 at N = 2^20 (prove ms): (4096, 2^16) 1.93 | (8192, 2^16) 2.01 | (2048, 2^16) 1.88 | (1024, 2^16) 1.89 |
// (4096, 2^15) 1.91 | (4096, 2^17) 1.93 | (2048, 2^15) 1.88 | (2048, 2^14) 1.90 | (1024, 2^15) 1.87
#ifndef SP2_SC_TAIL_LEN
#define SP2_SC_TAIL_LEN 1024
#endif
#ifndef SP2_SC_ROLE_LOG
#define SP2_SC_ROLE_LOG 15
#endif
constexpr unsigned long long SC_TAIL_LEN = SP2_SC_TAIL_LEN;   // tables this short are finished by one CTA in one launch
constexpr unsigned long long SC_ROLE_LEN = 1ull << SP2_SC_ROLE_LOG;   // tables this short use the role-split (3 items per pair) rounds

// tau inverses for the derived t(1): written by the HOST into pinned, device-visible memory (one batch inversion per prove, off the
// critical path: the first streaming round takes ~60 us) and read by the streaming kernel's finaliser — a kernel of its own between the
// gate and the cooperative launch would hold that launch back until the inverses arrive (measured: +13..33 us)
struct ScTinvMail { u32 flag, n, pad[6]; u32 tinv[SC_DERIVE_MAX][8]; };   // flag: the prover's epoch, release-stored last

struct ScState {
  DevTranscript ts;           // transcript hand-off: (round, state) in, (round, state) out
  fe claim;                   // quad: running claim (cubic sums t(0), t(1), t(inf) directly)
  fe p;                       // eval_eq_left (sumcheck.rs:951)
  fe L0, SL;                  // p*(1-tau_i), p*(2tau_i-1) for the round being evaluated
  u32 ticket, l, flags, arrived;     // arrived / released: monotonic counters of the persistent kernels' grid barrier
  u32 released, err, r_ready, pad1;   // r_ready: rounds of the QUADRATIC prover whose challenge is in r[] (polled by the prover's side stream);
                                     // err: a bounded device wait (grid barrier / peer mailbox) timed out -> SP2_ERR_INTERNAL on the host
  u32 mid_arrive[SC_MAX_ROUNDS + 8];   // pipelined multi-CTA rounds: CTAs whose partial sums of a round are published
  u32 mid_released, pad2[7];           //   and the last round whose challenge is published
  u32 mid_acc0[12 * 16];               //   column accumulators of the first such round (3 direct + 9 coefficient sums), zero on entry
  // t(1) derived from the running claim in the streaming rounds of the cubic prover (sumcheck.rs:1277-1324): rounds 1..derive_rounds of
  // k_cubic_persist sum only t(0) and t(inf); the finaliser solves (1 - tau_i) t(0) + tau_i t(1) = t_{i-1}(r_{i-1}) with the host's
  // tau_i^-1 (ScTinvMail, below).  0 = all three sums directly (every other path, and any tau_i = 0).
  u32 derive_rounds, pad3[7];
  fe tclaim;                           // t_{i-1}(r_{i-1}) of the round being evaluated (round 1: the claim)
  ScTinvMail mail;                     // the standalone provers' inverses (the host has the taus at upload time); the fused prover uses pinned memory
  // flags bit 0: warp-shuffle Keccak (default; measured 12.6k cycles/squeeze vs 22k for the one-thread register version)
  fe taus[SC_MAX_ROUNDS];
  // ---- everything above is uploaded by the host; everything below is produced on the device ----
  fe r[SC_MAX_ROUNDS];
  fe polys[SC_MAX_ROUNDS * 4];
  fe claims[4];
  unsigned long long clk[16];   // debug: clock64() stamps of the last finalised round (thread 0)
  unsigned long long gt[8];     // debug: %globaltimer (ns) of the last multi-CTA round: [first CTA entry, election, finalize end]
  unsigned long long prof[SC_MAX_ROUNDS][4];   // debug: per persistent round on CTA 0: %globaltimer at start, own compute done, all arrived, finalised
  fe partial[3 * SC_MAX_BLOCKS];
};

// ---- multi-GPU sharding (one process per GPU; peers' mailboxes are CUDA-IPC mapped over NVLink) ----------------
// The 2^l hypercube is split CYCLICALLY on the low index bits: rank g of G = 2^k owns global indices i = g (mod G),
// stored densely at i >> k.  Both entries of every bind pair (i, i + len/2) share their low bits, so every round is
// local; only the <= 3 partial sums cross GPUs, written by the round kernel's last CTA straight into every peer's
// mailbox (st.global over NVLink + system fence + flag) — the exchange is fused into the compute kernel, no NCCL
// call and no extra launch per round.  Once the table has <= SC_ROLE_LEN entries the shards are all-gathered
// (again by peer stores) and every rank finishes the latency-bound small rounds redundantly.
constexpr int SC_MAX_RANKS = 8;
struct MailBox {
  fe sums[SC_MAX_ROUNDS + 2][SC_MAX_RANKS][4];
  u32 flag[SC_MAX_ROUNDS + 2][SC_MAX_RANKS];
  fe gather[3][SC_ROLE_LEN];
  u32 err, pad[7];            // set by a round / barrier kernel of the LOCAL rank whose bounded wait for a peer expired
  // the prover's transcript hand-over: rank 0 alone hashes the commitment rows on its host and forwards (round, state, tau digests) to
  // every peer's mailbox from inside its gate kernel (prover.cu: k_gate_taus) — no rank but 0 spends host time on the transcript head
  u32 late_flag, late_round, late_pad[6];
  u32 late_words[16 + SC_MAX_ROUNDS * 16];
};
// every device-side wait is bounded (a dead or failed peer must surface as an error code, not wedge the GPU)
constexpr unsigned long long SC_WAIT_NS = 2000000000ull;
struct DevComm {
  int rank, n, k;             // n = 2^k ranks
  u32 epoch;                  // distinguishes successive sum-checks (flags are compared to it, never reset)
  MailBox *peer[SC_MAX_RANKS];   // peer[rank] is the local mailbox
};

}  // namespace sp2
struct sp2_comm {                // host handle of one rank's endpoint (comm.cu)
  sp2_ctx *ctx = nullptr;
  sp2::DevComm dc;
  bool connected = false;
  bool in_process = false;       // peers are contexts of THIS process (sp2_comm_connect_ptrs): their streams share the hardware queues
  bool opened[sp2::SC_MAX_RANKS] = {false};
};
namespace sp2 {

int sc_state_upload(sp2_ctx *ctx, ScState **d_st, const uint64_t *claim, const uint64_t *taus, uint32_t l,
                    const sp2_transcript_state *ts);
int sc_state_download(sp2_ctx *ctx, ScState *d_st, sp2_transcript_state *ts, uint64_t *polys, int ncoef, uint64_t *r,
                      uint64_t *claims, int nclaims, uint32_t l);
// dc: nullptr for a single GPU; otherwise A, B, C are this rank's cyclic shards (2^(l-k) entries) of the global tables
int sumcheck_cubic_enqueue(sp2_ctx *ctx, ScState *st, uint32_t l, fe *A, fe *B, fe *C, const DevComm *dc = nullptr, const ScTinvMail *mail = nullptr,
                           uint32_t mail_epoch = 0);
// rounds of a single-GPU cubic sum-check over 2^l entries that run in the persistent streaming kernel (0: none) — the rounds that may
// derive t(1) from the claim (ScState::derive_rounds must not exceed it)
uint32_t sumcheck_cubic_persist_rounds(sp2_ctx *ctx, uint32_t l);
// before enqueuing behind a kernel that waits for the host: load the kernels / reserve the scratch (sumcheck.cu)
void sumcheck_cubic_preload();
int sumcheck_cubic_reserve(sp2_ctx *ctx, uint32_t l);
// nvalid: table entries at index >= nvalid are unmaterialised zeros (~0ull: dense tables)
// after_first (optional): recorded once the launch that produces r[0] has been enqueued
int sumcheck_quad_enqueue(sp2_ctx *ctx, ScState *st, uint32_t rounds, fe *A, fe *B, uint64_t nvalid, cudaEvent_t after_first,
                          const DevComm *dc = nullptr);
// after a stream synchronisation: SP2_ERR_INTERNAL if a bounded peer wait of this rank expired (comm.cu)
int comm_check(sp2_ctx *ctx, sp2_comm *c);

}  // namespace sp2
