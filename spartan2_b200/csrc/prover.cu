// Fused SpartanSNARK prover: prep_prove + prove on the device.
//
// Restates reference src/spartan.rs:176-216 (prep_prove) and :219-466 (prove), with the witness
// commitment of src/bellpepper/r1cs.rs:359-537 (commit of the precommitted section, rest section,
// transcript absorbs) and HyraxPCS::prove + the linear inner-product argument
// (src/provider/pcs/hyrax_pc.rs:387-478, src/provider/pcs/ipa.rs:125-170).
//
// Device/host split (B200 design): every field/group computation runs on the device; the host only
// sequences launches and hashes BULK transcript data (commitment rows: tens of KB through a serial
// Keccak is host-speed work).  The two sum-check loops run on the device end to end with the
// transcript handed over as (round, state); the outer->inner transition (absorb claims_outer,
// squeeze r, joint claim) also stays on the device.  Host syncs per prove: 3 (rest-commitment rows,
// sum-check results + PCS points, IPA response) instead of one per sum-check round.
#include <string.h>
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include "ctx.cuh"
#include "host_transcript.h"
#include "keccak.cuh"
#include "msm.cuh"
#include "polys.cuh"
#include "r1cs.cuh"
#include "sumcheck.cuh"

using namespace sp2;

namespace sp2 {
int eq_table_dev(sp2_ctx *ctx, const fe *d_r, uint32_t k, fe *d_out, cudaStream_t stream = nullptr, int slot = 15);
int eq_table_reserve(sp2_ctx *ctx, uint32_t k, int slot);
}

// One helper host thread per prep state: hashes the BULK head of the transcript (the cached commitment rows, ~30 KB of serial Keccak =
// ~60 us) while the calling thread enqueues the device work of the same phase — the commit + transcript phase of a prove is bound
// by the host, not by the GPU.  The worker makes no CUDA calls.
struct HostWorker {
  std::thread th; std::mutex m; std::condition_variable cv;
  std::deque<std::function<void()>> q; bool busy = false, quit = false;     // jobs run in submission order
  void start() {
    th = std::thread([this] {
      std::unique_lock<std::mutex> lk(m);
      for (;;) {
        cv.wait(lk, [this] { return !q.empty() || quit; });
        if (q.empty()) return;
        std::function<void()> j = std::move(q.front()); q.pop_front(); busy = true;
        lk.unlock(); j(); lk.lock();
        busy = false; cv.notify_all();
      }
    });
  }
  void submit(std::function<void()> j) { { std::lock_guard<std::mutex> lk(m); q.push_back(std::move(j)); } cv.notify_all(); }
  void wait() { std::unique_lock<std::mutex> lk(m); cv.wait(lk, [this] { return q.empty() && !busy; }); }
  void stop() { if (!th.joinable()) return; { std::lock_guard<std::mutex> lk(m); quit = true; } cv.notify_all(); th.join(); }
};

struct LateIn {
  sp2::u32 flag;                   // host: prove epoch, release-stored LAST; bit 31 = abort (the host gave up: error path)
  sp2::u32 round;
  unsigned char state[64];
  sp2::u32 flag2;                  // second gate (IPA challenge), same protocol
  unsigned char pad[52];
  unsigned char dg[SC_MAX_ROUNDS * 64];
  unsigned char rdg[64];           // the IPA challenge digest
  unsigned char pad3[32];
  sp2::ScTinvMail mail;            // third mailbox: the inverses of the first taus (t(1) from the claim, sumcheck.cuh), Montgomery limbs
};
struct sp2_prep {
  HostWorker worker;
  sp2_ctx *ctx = nullptr;
  const sp2_shape *S = nullptr;
  const sp2_ck *ck = nullptr;
  uint64_t cached_len = 0, cached_rows = 0, rows_total = 0;
  fe *W = nullptr;                   // num_vars (cached part filled by prep_prove, rest by prove)
  fe *cached[3] = {nullptr, nullptr, nullptr};   // Az/Bz/Cz over the shared+precommitted columns (spartan.rs:184-187)
  fe *work[3] = {nullptr, nullptr, nullptr};     // sum-check tables
  fe *z = nullptr, *abc = nullptr;   // num_cols each (abc: 2M/G, dense, when the shape is a multi-GPU shard)
  fe *zs = nullptr, *rx = nullptr;   // multi-GPU only: this rank's cyclic shard of the virtual 2M-entry z table; full eq(r_x)
  fe *blinds = nullptr;              // rows_total
  fe *small = nullptr;               // scalars scratch (see offsets below)
  jac *points = nullptr;             // device points scratch (Jacobian): [rows_total comm rows | 4 PCS points]
  fe *LZ = nullptr, *Ltab = nullptr, *Rtab = nullptr, *dvec = nullptr, *zvec = nullptr;
  std::vector<uint64_t> comm_cached; // host copy of the cached commitment rows (affine)
  std::vector<uint8_t> comm_cached_be;  // their transcript bytes (x_BE || y_BE per row), computed once
  std::vector<uint64_t> blinds_cached_host;   // the blinds prep_prove committed the cached rows with (prove reuses them)
  fe *inbox = nullptr;               // per-prove host inputs, one H2D copy: [small slots | d_vec | blinds | X]
  uint8_t *h_inbox = nullptr;        // pinned staging of the same layout (+ tau digests)
  size_t inbox_bytes = 0;
  cudaEvent_t ev[9] = {nullptr};
  cudaEvent_t ev_r1 = nullptr, ev_inv = nullptr, ev_in = nullptr, ev_lz = nullptr, ev_delta = nullptr, ev_q = nullptr, ev_evalw = nullptr, ev_eq = nullptr, ev_chunks = nullptr;
  LateIn *h_late = nullptr, *d_late = nullptr;   // pinned, device-visible (k_gate_taus)
  uint32_t epoch = 0;
  cudaStream_t side2 = nullptr;      // early comm_LZ chain (prove): runs under the inner sum-check's last rounds
  std::vector<void *> owned;
};

namespace {

enum SmallSlot { S_RJOINT = 0, S_EVALW = 1, S_EVALX = 2, S_RLZ = 3, S_IP = 4, S_BLIND_EVAL = 5, S_RDELTA = 6, S_RBETA = 7,
                 S_ZDELTA = 8, S_ZBETA = 9, S_RIPA = 10, S_ERR = 11, S_DENINV = 12, S_COUNT = 16 };

// ---- the gate: everything the outer sum-check needs from the HOST transcript arrives through pinned, device-visible memory ----
// The host hashes the commitment rows and squeezes the taus while the device commits / multiplies; the launches of the outer
// sum-check are enqueued BEFORE that hashing ends, behind k_gate_taus, which waits (bounded) for the host's flag, copies the
// transcript hand-over (round, state) into the sum-check state and turns the squeezed digests into taus (from_uniform: LE 512-bit
// mod p).  The ~30 us of launch calls and the H2D copy no longer sit between the last squeeze and the first round.
__device__ __forceinline__ u32 ld_sys_u32(const void *p) { u32 v; asm volatile("ld.volatile.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
// Sharded proves (dc.n > 1): rank 0 alone has the host transcript; its gate forwards the hand-over to every peer's mailbox over
// NVLink (plain stores + system fence + flag, like the round sums), and the peers' gates wait on their own mailbox in device memory —
// the other ranks' hosts never hash the commitment rows (with 8 processes on one host that hashing was the largest phase).
__global__ void __launch_bounds__(64) k_gate_taus(ScState *st, const LateIn *late, u32 epoch, int l, int derive_max, DevComm dc, u32 xflag) {
  __shared__ int tau_zero;
  __shared__ u32 w[16 + SC_MAX_ROUNDS * 16];      // transcript state (64 B), then the tau digests (64 B each)
  __shared__ u32 s_round, s_abort;
  const int tid = threadIdx.x;
  const int nw = 16 + 16 * l;
  const bool from_host = dc.n <= 1 || dc.rank == 0;
  MailBox *mine = dc.n > 1 ? dc.peer[dc.rank] : nullptr;
  if (tid == 0) {
    tau_zero = 0; s_abort = 0;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (;;) {
      const u32 v = from_host ? ld_sys_u32(&late->flag) : ld_sys_u32(&mine->late_flag);
      if ((v & 0x7fffffffu) == (from_host ? epoch : xflag)) { if (v >> 31) { atomicExch(&st->err, 1u); s_abort = 1; } break; }
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      if (t - t0 > SC_WAIT_NS) { atomicExch(&st->err, 1u); s_abort = 1; break; }
      __nanosleep(100);
    }
    __threadfence_system();
    s_round = from_host ? ld_sys_u32(&late->round) : ld_sys_u32(&mine->late_round);
  }
  __syncthreads();
  for (int k = tid; k < nw; k += 64)
    w[k] = from_host ? (k < 16 ? ld_sys_u32(late->state + 4 * k) : ld_sys_u32(late->dg + 4 * (k - 16))) : ld_sys_u32(&mine->late_words[k]);
  __syncthreads();
  if (dc.n > 1 && dc.rank == 0) {
    for (int q = 1; q < dc.n; q++) {
      MailBox *pm = dc.peer[q];
      for (int k = tid; k < nw; k += 64) pm->late_words[k] = w[k];
      if (tid == 0) pm->late_round = s_round;
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      for (int q = 1; q < dc.n; q++) *(volatile u32 *)&dc.peer[q]->late_flag = xflag | (s_abort ? 0x80000000u : 0u);
      __threadfence_system();
    }
  }
  if (tid < 16) ((u32 *)st->ts.state)[tid] = w[tid];
  if (tid == 0) { st->ts.round = s_round; st->ts.pending_len = 0; }
  if (tid < l) {
    fe lo, hi;
#pragma unroll
    for (int k = 0; k < 8; k++) { lo.v[k] = w[16 + 16 * tid + k]; hi.v[k] = w[16 + 16 * tid + 8 + k]; }
    const fe tau = Fq::from_uniform(lo, hi);
    stg_fe(&st->taus[tid], tau);
    if (tid < derive_max && Fq::is_zero(tau)) tau_zero = 1;
  }
  __syncthreads();
  // the streaming rounds derive t(1) from the claim unless one of their taus is zero (the host takes the same decision when it inverts them)
  if (tid == 0) { st->derive_rounds = tau_zero ? 0u : (u32)derive_max; stg_fe(&st->tclaim, ldg_fe(&st->claim)); }
}
// second gate: the IPA challenge digest (squeezed on the host from the PCS points) -> device memory; k_ipa_finish is enqueued behind it
__global__ void __launch_bounds__(32) k_gate_ipa(ScState *st, const LateIn *late, u32 epoch, unsigned char *d_dg) {
  const int tid = threadIdx.x;
  if (tid == 0) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (;;) {
      const u32 v = ld_sys_u32(&late->flag2);
      if ((v & 0x7fffffffu) == epoch) { if (v >> 31) atomicExch(&st->err, 1u); break; }
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      if (t - t0 > SC_WAIT_NS) { atomicExch(&st->err, 1u); break; }
      __nanosleep(100);
    }
    __threadfence_system();
  }
  __syncwarp();
  if (tid < 16) ((u32 *)d_dg)[tid] = ld_sys_u32(late->rdg + 4 * tid);
}

// z[num_vars ..] = 1 | X   (spartan.rs:248-253)
__global__ void k_z_tail(fe *ztail, const fe *X, u32 num_public) {
  const u32 i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) stg_fe(ztail, Fq::one());
  if (i < num_public) stg_fe(ztail + 1 + i, ldg_fe(X + i));
}
// multi-GPU: this rank's cyclic shard of the inner sum-check's virtual z table (2M entries: z | zeros)
__global__ void __launch_bounds__(256) k_z_shard(const fe *z, u64 nc, fe *zs, u64 len_local, int k, int rank) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= len_local) return;
  const u64 g = (i << k) | (u64)rank;
  stg_fe(zs + i, g < nc ? ldg_fe(z + g) : Fq::zero());
}
// 1 / (1 - r_y[0]) as soon as the first inner challenge exists (side stream, overlaps the remaining inner rounds)
__global__ void k_den_inv(const ScState *inner, fe *small) {
  if (threadIdx.x == 0) {
    const fe den = Fq::sub(Fq::one(), ldg_fe(&inner->r[0]));
    if (Fq::is_zero(den)) ((u32 *)&small[S_ERR])[0] = 5;   // SpartanError::DivisionByZero
    stg_fe(&small[S_DENINV], Fq::inv(den));
  }
}

// side stream: returns once the quadratic prover has published the challenges of rounds 1..round (ScState::r_ready), so that the
// work queued behind it (the Hyrax bind and the comm_LZ MSM, which need only r_y[1..log2 rows]) runs under the remaining rounds.
// Bounded like every device-side wait; on a timeout the sum-check state carries the error to the host.
__global__ void k_wait_round(ScState *st, u32 round) {
  if (threadIdx.x != 0) return;
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  for (;;) {
    u32 v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(&st->r_ready));
    if (v >= round) break;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    if (t - t0 > SC_WAIT_NS) { atomicExch(&st->err, 1u); break; }
    __nanosleep(200);
  }
  __threadfence();
}

// outer -> inner: absorb(b"claims_outer", [A(rx), B(rx), C(rx)]); r = squeeze(b"r");
// joint = A + r B + r^2 C (spartan.rs:305-316).  Initialises the inner sum-check state.
__global__ void __launch_bounds__(64) k_outer_to_inner(ScState *outer, ScState *inner, fe *small, int rounds_inner) {
  __shared__ unsigned char buf[2304];
  __shared__ fe ch;
  DevTranscript *ts = &outer->ts;
  if (threadIdx.x == 0) {
    const char lab[] = "claims_outer";
    ts->pending_len = 0;
    ts_push_bytes(ts, (const unsigned char *)lab, 12);
    for (int k = 0; k < 3; k++) ts_push_fe_be(ts, Fq::from_mont(outer->claims[k]));
  }
  __syncthreads();
  ts_squeeze_block(ts, "r", 1, buf, &ch);
  if (threadIdx.x == 0) {
    const fe r = ch;
    const fe joint = Fq::add(outer->claims[0], Fq::mul(r, Fq::add(outer->claims[1], Fq::mul(r, outer->claims[2]))));
    stg_fe(&small[S_RJOINT], r);
    inner->ts.round = ts->round; inner->ts.pending_len = 0;
    for (int i = 0; i < 64; i++) inner->ts.state[i] = ts->state[i];
    stg_fe(&inner->claim, joint);
    inner->ticket = 0; inner->l = (u32)rounds_inner; inner->flags = outer->flags; inner->arrived = 0; inner->released = 0; inner->err = 0; inner->r_ready = 0;
    for (int i = 0; i < SC_MAX_ROUNDS + 8; i++) inner->mid_arrive[i] = 0;
    inner->mid_released = 0;
    for (int i = 0; i < 12 * 16; i++) inner->mid_acc0[i] = 0;
  }
}

// eval_X = SparsePolynomial(X).evaluate(r_y[1..]) (polys/multilinear.rs:190-207);
// eval_W = (eval_Z - r_y[0] eval_X) / (1 - r_y[0])   (spartan.rs:411-421)
__global__ void __launch_bounds__(256) k_eval_w(const ScState *inner, const fe *X, u32 zlen, int m, fe *chis_scratch, fe *small) {
  __shared__ fe red[32];
  int nvz = 0; while (((u32)1 << nvz) < zlen) nvz++;
  const int skip = m - 1 - nvz, k = nvz + 1;
  const fe *ry1 = inner->r + 1;                       // r_y[1..]
  eq_prefix_block(ry1 + skip, k, chis_scratch);
  const fe *chis = chis_scratch + (((size_t)1 << k) - 1);
  fe x[1] = {Fq::zero()};
  for (u32 i = threadIdx.x; i < zlen; i += blockDim.x) x[0] = Fq::add(x[0], Fq::mul(ldg_fe(X + i), ldg_fe(chis + i)));
  block_sum_fq<1>(x, red);
  // prod_{i < skip} (1 - r_y[1 + i]): a warp tree (5 dependent multiplications) instead of a serial chain of m - 1 - nvz ~ 18
  fe common = Fq::one();
  if (threadIdx.x < 32) {
    for (int i = threadIdx.x; i < skip; i += 32) common = Fq::mul(common, Fq::sub(Fq::one(), ldg_fe(ry1 + i)));
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) common = Fq::mul(common, shfl_xor_fe(common, d));
  }
  if (threadIdx.x == 0) {
    const fe eval_X = Fq::mul(common, x[0]);
    const fe ry0 = ldg_fe(&inner->r[0]);
    const fe eval_W = Fq::mul(Fq::sub(ldg_fe(&inner->claims[1]), Fq::mul(ry0, eval_X)), ldg_fe(&small[S_DENINV]));
    stg_fe(&small[S_EVALX], eval_X);
    stg_fe(&small[S_EVALW], eval_W);
  }
}

// out = <a, b> (delayed reduction), one CTA
__global__ void __launch_bounds__(256) k_dot(const fe *a, const fe *b, u64 n, fe *out) {
  __shared__ fe red[32];
  Fq::acc acc = Fq::acc_zero();
  for (u64 i = threadIdx.x; i < n; i += blockDim.x) Fq::mul_acc(acc, ldg_fe(a + i), ldg_fe(b + i));
  fe x[1] = {Fq::acc_reduce(acc)};
  block_sum_fq<1>(x, red);
  if (threadIdx.x == 0) stg_fe(out, x[0]);
}

// IPA response (ipa.rs:155-167): r from the host-squeezed digest; z_vec = r*LZ + d; z_delta = r*r_LZ + r_delta;
// z_beta = r*blind_eval + r_beta
__global__ void __launch_bounds__(256) k_ipa_finish(const unsigned char *dg, const fe *LZ, const fe *d, u64 n, fe *zvec, fe *small) {
  fe lo, hi;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const unsigned char *p = dg + 4 * k;
    lo.v[k] = (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
    hi.v[k] = (u32)p[32] | ((u32)p[33] << 8) | ((u32)p[34] << 16) | ((u32)p[35] << 24);
  }
  const fe r = Fq::from_uniform(lo, hi);
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) stg_fe(zvec + i, Fq::add(Fq::mul(r, ldg_fe(LZ + i)), ldg_fe(d + i)));
  if (i == 0) {
    stg_fe(&small[S_RIPA], r);
    stg_fe(&small[S_ZDELTA], Fq::add(Fq::mul(r, ldg_fe(&small[S_RLZ])), ldg_fe(&small[S_RDELTA])));
    stg_fe(&small[S_ZBETA], Fq::add(Fq::mul(r, ldg_fe(&small[S_BLIND_EVAL])), ldg_fe(&small[S_RBETA])));
  }
}

template <class T>
int palloc(sp2_prep *P, T **p, size_t n) {
  sp2_ctx *ctx = P->ctx;
  SP2_CUDA_OK(cudaMalloc((void **)p, n * sizeof(T) + 64));
  P->owned.push_back(*p);
  return SP2_OK;
}

int log2_exact(uint64_t v) { int l = 0; while (((uint64_t)1 << l) < v) l++; return l; }

}  // namespace

extern "C" {

void sp2_prep_free(sp2_prep *P) {
  if (!P) return;
  P->worker.stop();
  cudaSetDevice(P->ctx->device);
  cudaStreamSynchronize(P->ctx->stream);
  for (void *p : P->owned) cudaFree(p);
  if (P->h_inbox) cudaFreeHost(P->h_inbox);
  if (P->h_late) cudaFreeHost(P->h_late);
  for (auto &e : P->ev) if (e) cudaEventDestroy(e);
  if (P->ev_r1) cudaEventDestroy(P->ev_r1);
  if (P->ev_inv) cudaEventDestroy(P->ev_inv);
  if (P->ev_in) cudaEventDestroy(P->ev_in);
  if (P->ev_lz) cudaEventDestroy(P->ev_lz);
  if (P->ev_delta) cudaEventDestroy(P->ev_delta);
  if (P->ev_q) cudaEventDestroy(P->ev_q);
  if (P->ev_evalw) cudaEventDestroy(P->ev_evalw);
  if (P->ev_eq) cudaEventDestroy(P->ev_eq);
  if (P->ev_chunks) cudaEventDestroy(P->ev_chunks);
  if (P->side2) { cudaStreamSynchronize(P->side2); cudaStreamDestroy(P->side2); }
  delete P;
}

/* SpartanSNARK::prep_prove (spartan.rs:176-216): commit the shared+precommitted witness sections
 * (bellpepper/r1cs.rs:306-408 -> HyraxPCS::commit) and cache their Az/Bz/Cz (multiply_vec_precommitted). */
int32_t sp2_spartan_prep_prove(sp2_ctx *ctx, const sp2_shape *S, const sp2_ck *ck, const uint64_t *W_cached, const uint64_t *blinds_cached,
                               int32_t is_small, uint64_t *comm_out, sp2_prep **out) {
  (void)is_small;
  cudaSetDevice(ctx->device);
  if (!out) return SP2_ERR_INTERNAL;
  *out = nullptr;
  const uint64_t width = ck->n, nv = S->num_vars;
  if (nv == 0 || (nv & (nv - 1))) return set_error(ctx, SP2_ERR_INVALID_WITNESS_LENGTH, "prep_prove: num_vars must be a power of two");
  if (nv % width || S->num_shared % width || S->num_precommitted % width || S->num_rest % width)
    return set_error(ctx, SP2_ERR_INVALID_WITNESS_LENGTH, "prep_prove: witness sections must be multiples of the commitment width");
  if (S->num_challenges) return set_error(ctx, SP2_ERR_UNSUPPORTED, "prep_prove: multi-round (challenge) circuits are not offloaded");
  const bool shard = S->nranks > 1;
  if (shard && ((2 * nv) >> S->shard_k) == 0) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "prep_prove: fewer variables than ranks");
  sp2_prep *P = new sp2_prep();
  P->ctx = ctx; P->S = S; P->ck = ck;
  P->cached_len = S->num_shared + S->num_precommitted; P->cached_rows = P->cached_len / width; P->rows_total = nv / width;
  const uint64_t N = S->num_cons, nc = S->num_cols, Nl = S->rows_local;
  int rc = SP2_OK;
  auto A = [&](int r) { if (rc == SP2_OK) rc = r; };
  A(palloc(P, &P->W, nv));
  for (int k = 0; k < 3; k++) { A(palloc(P, &P->cached[k], Nl)); A(palloc(P, &P->work[k], Nl)); }
  A(palloc(P, &P->z, nc)); A(palloc(P, &P->abc, shard ? (2 * nv) >> S->shard_k : nc));
  if (shard) { A(palloc(P, &P->zs, (2 * nv) >> S->shard_k)); A(palloc(P, &P->rx, N)); }
  A(palloc(P, &P->points, P->rows_total + 8));
  A(palloc(P, &P->LZ, width)); A(palloc(P, &P->Ltab, std::max<uint64_t>(P->rows_total, 1))); A(palloc(P, &P->Rtab, width));
  A(palloc(P, &P->zvec, width));
  P->inbox_bytes = ((size_t)S_COUNT + width + P->rows_total + S->num_public) * sizeof(fe);
  { fe *ib = nullptr; A(palloc(P, &ib, (size_t)S_COUNT + width + P->rows_total + S->num_public)); P->inbox = ib; }
  // pinned staging: [inbox | (spare) | rest-commitment rows (Jacobian) read back in prove, later the IPA response z_vec]
  if (rc == SP2_OK && cudaMallocHost((void **)&P->h_inbox, P->inbox_bytes + SC_MAX_ROUNDS * 64 + 64 + std::max<size_t>(P->rows_total * sizeof(jac), width * sizeof(fe))) != cudaSuccess) rc = set_error(ctx, SP2_ERR_CUDA, "cudaMallocHost");
  if (rc == SP2_OK) {
    P->small = P->inbox; P->dvec = P->inbox + S_COUNT; P->blinds = P->dvec + width;
    for (auto &e : P->ev) cudaEventCreate(&e);
    cudaEventCreateWithFlags(&P->ev_r1, cudaEventDisableTiming); cudaEventCreateWithFlags(&P->ev_inv, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&P->ev_in, cudaEventDisableTiming); cudaEventCreateWithFlags(&P->ev_lz, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&P->ev_delta, cudaEventDisableTiming); cudaEventCreateWithFlags(&P->ev_q, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&P->ev_evalw, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&P->ev_eq, cudaEventDisableTiming); cudaEventCreateWithFlags(&P->ev_chunks, cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&P->side2, cudaStreamNonBlocking);
    if (cudaHostAlloc((void **)&P->h_late, sizeof(LateIn), cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void **)&P->d_late, P->h_late, 0) != cudaSuccess) rc = set_error(ctx, SP2_ERR_CUDA, "cudaHostAlloc (gate)");
    else memset(P->h_late, 0, sizeof(LateIn));
  }
  if (rc != SP2_OK) { sp2_prep_free(P); return rc; }
  auto fail = [&](int r) { sp2_prep_free(P); return r; };
  cudaError_t e = cudaMemsetAsync(P->W, 0, nv * sizeof(fe), ctx->stream);
  if (e == cudaSuccess && P->cached_len) e = cudaMemcpyAsync(P->W, W_cached, P->cached_len * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream);
  if (P->cached_rows) P->blinds_cached_host.assign(blinds_cached, blinds_cached + 4 * P->cached_rows);
  if (e == cudaSuccess && P->cached_rows) e = cudaMemcpyAsync(P->blinds, blinds_cached, P->cached_rows * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream);
  if (e != cudaSuccess) return fail(set_cuda_error(ctx, e, "prep upload", __LINE__));
  // commitments of the cached rows
  if (P->cached_rows) {
    rc = sp2_hyrax_commit_dev(ctx, ck, P->W, P->cached_len, P->blinds, P->cached_rows, P->points);
    if (rc != SP2_OK) return fail(rc);
    P->comm_cached.resize(P->cached_rows * 8);
    std::vector<uint64_t> hj(P->cached_rows * 12);
    e = cudaMemcpyAsync(hj.data(), P->points, P->cached_rows * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return fail(set_cuda_error(ctx, e, "prep download", __LINE__));
    sp2h::batch_normalize(hj.data(), P->cached_rows, P->comm_cached.data());
    P->comm_cached_be.resize(64 * P->cached_rows);
    for (uint64_t i = 0; i < P->cached_rows; i++) sp2h::point_be(P->comm_cached.data() + 8 * i, P->comm_cached_be.data() + 64 * i);
  }
  // cached partial products: z = [W_cached | 0 ... 0]
  e = cudaMemsetAsync(P->z, 0, nc * sizeof(fe), ctx->stream);
  if (e == cudaSuccess && P->cached_len) e = cudaMemcpyAsync(P->z, P->W, P->cached_len * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream);
  if (e != cudaSuccess) return fail(set_cuda_error(ctx, e, "prep z", __LINE__));
  rc = spmv3_dev(ctx, S, S->M, P->z, nullptr, P->cached);
  if (rc != SP2_OK) return fail(rc);
  e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) return fail(set_cuda_error(ctx, e, "prep sync", __LINE__));
  if (comm_out && P->cached_rows) memcpy(comm_out, P->comm_cached.data(), P->cached_rows * sizeof(aff));
  P->worker.start();
  *out = P;
  return SP2_OK;
}

/* SpartanSNARK::prove (spartan.rs:219-466).  phase_ms (optional, 8 floats): device time of
 * [witness commit + transcript, matrix_vector_multiply, outer_sumcheck, prepare_poly_ABC, inner_sumcheck,
 *  pcs_prove (bind + MSMs), ipa response, total]. */
static int32_t spartan_prove_impl(sp2_ctx *ctx, sp2_comm *comm, const sp2_shape *S, const sp2_ck *ck, sp2_prep *P, const uint8_t *vk_digest,
                                  const uint64_t *public_values, const uint64_t *W_rest, const sp2_spartan_rand *rnd, sp2_spartan_proof *proof,
                                  float *phase_ms) {
  cudaSetDevice(ctx->device);
  if (!P || P->S != S || P->ck != ck) return set_error(ctx, SP2_ERR_INTERNAL, "prove: prep state does not belong to this shape/key");
  const bool shard = S->nranks > 1;
  if (shard && (!comm || !comm->connected || comm->dc.n != S->nranks || comm->dc.rank != S->rank))
    return set_error(ctx, SP2_ERR_INTERNAL, "prove: a sharded shape needs the connected comm of the same rank / size");
  if (!shard && comm && comm->dc.n > 1) return set_error(ctx, SP2_ERR_INTERNAL, "prove: comm given but the shape is not sharded");
  const uint64_t width = ck->n, nv = S->num_vars, N = S->num_cons, nc = S->num_cols;
  const int l = log2_exact(N), m = log2_exact(nv), nry = m + 1;
  const uint64_t num_extra = 1 + S->num_public;
  const uint64_t rows = P->rows_total, rest_rows = rows - P->cached_rows;
  int nvr = log2_exact(rows);
  if (((uint64_t)1 << nvr) != rows || (width & (width - 1))) return set_error(ctx, SP2_ERR_INVALID_CK_LENGTH, "prove: rows and width must be powers of two");
  int nvz = 0; while (((uint64_t)1 << nvz) < num_extra) nvz++;
  if (m - 1 - nvz < 0) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "prove: too many public inputs for the witness length");
  if (l > SC_MAX_ROUNDS || nry > SC_MAX_ROUNDS) return set_error(ctx, SP2_ERR_INVALID_INPUT_LENGTH, "prove: instance too large");

  cudaEvent_t *ev = P->ev;
  auto mark = [&](int i) { cudaEventRecord(ev[i], ctx->stream); };
  auto cleanup = [&]() {};
  mark(0);
  // SP2_PROVE_TRACE=1: host-side timeline of the prove (us since entry) on stderr — which side bounds a phase
  static const bool trace_on = [] { const char *e = getenv("SP2_PROVE_TRACE"); return e && e[0] == '1'; }();
  const auto t_entry = std::chrono::steady_clock::now();
  std::vector<std::pair<const char *, double>> trace;
  auto tp = [&](const char *what) { if (trace_on) trace.emplace_back(what, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_entry).count()); };

  // ---- transcript up to the taus (host; spartan.rs:226-264, bellpepper/r1cs.rs:422-431,491) -----
  // The ~30 KB of serial Keccak over the cached commitment rows run on the prep state's helper thread, beside the staging and the
  // enqueues of this phase.
  sp2h::Transcript ts("SpartanSNARK");
  auto absorb_head = [&]() {
    ts.absorb_bytes("vk", vk_digest, 32);
    ts.absorb_scalars("public_values", public_values, S->num_public);
    const uint64_t sh_rows = S->num_shared / width, pre_rows = S->num_precommitted / width;
    if (sh_rows) ts.absorb_commitment_be("comm_W_shared", P->comm_cached_be.data(), sh_rows);
    if (pre_rows) ts.absorb_commitment_be("comm_W_precommitted", P->comm_cached_be.data() + 64 * sh_rows, pre_rows);
    memcpy(proof->comm_W, P->comm_cached.data(), P->cached_rows * sizeof(aff));
  };
  // the head is hashed by the prep state's helper thread while this thread enqueues; joined before the rest rows are absorbed
  // (and on every early return: the job refers to locals)
  struct Join { HostWorker &w; ~Join() { w.wait(); } } join_head{P->worker};
  const bool hash_here = !shard || S->rank == 0;                             // sharded: rank 0 alone hashes; the taus reach the peers through
  if (hash_here) P->worker.submit(absorb_head);                             // its gate kernel and the NVLink mailboxes (k_gate_taus)
  else memcpy(proof->comm_W, P->comm_cached.data(), P->cached_rows * sizeof(aff));
  // ---- per-prove host inputs: one staged copy (randomness, blinds, public values) ------------------
  fe *small = P->small;
  { uint8_t *h = P->h_inbox;
    memset(h, 0, S_COUNT * sizeof(fe));
    memcpy(h + S_BLIND_EVAL * sizeof(fe), rnd->blind_eval_W, sizeof(fe));
    memcpy(h + S_RDELTA * sizeof(fe), rnd->r_delta, sizeof(fe));
    memcpy(h + S_RBETA * sizeof(fe), rnd->r_beta, sizeof(fe));
    uint8_t *q = h + S_COUNT * sizeof(fe);
    memcpy(q, rnd->d_vec, width * sizeof(fe)); q += width * sizeof(fe);
    // blinds: the shared / precommitted rows keep the blinds prep_prove committed them with (comm_W is reused from there, and
    // r_LZ = <L, blinds> must match it); only the rest rows take fresh blinds from `rand` (bellpepper/r1cs.rs:467-470)
    memcpy(q, P->blinds_cached_host.data(), P->cached_rows * sizeof(fe));
    memcpy(q + P->cached_rows * sizeof(fe), rnd->blinds_W + 4 * P->cached_rows, (rows - P->cached_rows) * sizeof(fe)); q += rows * sizeof(fe);
    if (S->num_public) memcpy(q, public_values, S->num_public * sizeof(fe));
    SP2_CUDA_OK(cudaMemcpyAsync(P->inbox, h, P->inbox_bytes, cudaMemcpyHostToDevice, ctx->stream)); }
  const fe *d_X = P->blinds + rows;
  // PCS points of this prove: [comm_LZ, delta, comm_eval_W, beta]
  jac *d_pts = P->points + rows;
  // delta = <d, ck> + r_delta h (ipa.rs:140-146) depends on the prover's randomness only: side stream, under the witness commitment
  // (enqueued after the work the host is about to wait for)
  auto enqueue_delta = [&]() -> int {
    MsmJob j; memset(&j, 0, sizeof(j)); j.scalars = P->dvec; j.len = (u32)width; j.nextra = 1; j.extra_base[0] = ck->idx_h(); j.extra_scalar[0] = &small[S_RDELTA];
    SP2_CUDA_OK(cudaStreamWaitEvent(ctx->side, P->ev_in, 0));
    SP2_TRY(msm_run(ctx, ck, std::vector<MsmJob>{j}, d_pts + 1, ctx->side, 16, 17));
    SP2_CUDA_OK(cudaEventRecord(P->ev_delta, ctx->side));
    return SP2_OK;
  };
  // rest section of the witness (NULL: all zero, e.g. pure padding as in the SHA-256 bench circuit)
  if (S->num_rest) {
    if (W_rest) SP2_CUDA_OK(cudaMemcpyAsync(P->W + P->cached_len, W_rest, S->num_rest * sizeof(fe), cudaMemcpyHostToDevice, ctx->stream));
    else SP2_CUDA_OK(cudaMemsetAsync(P->W + P->cached_len, 0, S->num_rest * sizeof(fe), ctx->stream));
  }
  SP2_CUDA_OK(cudaEventRecord(P->ev_in, ctx->stream));                      // this prove's inputs are on the device

  // commit the rest section (r1cs.rs:467-470: blind + commit / commit_zeros) first — the host waits for these rows; an all-zero rest
  // section is HyraxPCS::commit_zeros (hyrax_pc.rs:305-319): rows = blind_i * h, no row terms to walk.  (On a side stream beside the
  // SpMV the MSM's large CTAs starve behind the SpMV's 12k small ones: rows 45 us later, measured.)
  uint64_t *hj = (uint64_t *)(P->h_inbox + P->inbox_bytes + SC_MAX_ROUNDS * 64 + 64);      // pinned: a true async DMA
  if (rest_rows) {
    SP2_TRY(hyrax_commit_rows(ctx, ck, P->W + P->cached_len, W_rest ? S->num_rest : 0, P->blinds + P->cached_rows, rest_rows, P->points + P->cached_rows,
                              ctx->stream, 10, 11));
    SP2_CUDA_OK(cudaMemcpyAsync(hj, P->points + P->cached_rows, rest_rows * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
    SP2_CUDA_OK(cudaEventRecord(P->ev_r1, ctx->stream));
  }
  // z = W | 1 | X   (spartan.rs:248-253); Az, Bz, Cz (spartan.rs:271 -> multiply_vec_incremental_into) do not depend on the taus
  SP2_CUDA_OK(cudaMemcpyAsync(P->z, P->W, nv * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
  k_z_tail<<<(unsigned)(num_extra + 127) / 128, 128, 0, ctx->stream>>>(P->z + nv, d_X, (u32)S->num_public);
  SP2_LAUNCH_CHECK();
  { const fe *base[3] = {P->cached[0], P->cached[1], P->cached[2]};
    SP2_TRY(spmv3_dev(ctx, S, S->F, P->z, base, P->work)); }
  SP2_TRY(enqueue_delta());
  tp("commit, SpMV, delta enqueued");

  const uint32_t epoch = (++P->epoch) & 0x7fffffffu;
  LateIn *late = P->h_late;
  // ---- transcript: rest rows, taus (host) ----
  if (rest_rows) {
    SP2_CUDA_OK(cudaEventSynchronize(P->ev_r1));                            // host sync 1 (rows only)
    tp("rest rows on the host");
    sp2h::batch_normalize(hj, rest_rows, proof->comm_W + 8 * P->cached_rows);
    tp("rows normalised");
  }
  // the rest of the transcript up to the taus goes to the helper thread, behind the head: absorb the rest rows, squeeze the taus into
  // the gate's mailbox, open the gate — while this thread keeps enqueuing (poly_ABC, inner sum-check, PCS)
  // (SP2_NO_DERIVE=1: all three sums directly in every round — measurement switch)
  static const bool derive_on = [] { const char *e = getenv("SP2_NO_DERIVE"); return !(e && e[0] == '1'); }();
  const uint32_t derive_max = (shard || !derive_on) ? 0u : std::min<uint32_t>(SC_DERIVE_MAX, sumcheck_cubic_persist_rounds(ctx, (uint32_t)l));
  if (hash_here) P->worker.submit([&ts, late, epoch, l, proof, P, rest_rows, derive_max]() {
    ts.absorb_commitment("comm_W_rest", proof->comm_W + 8 * P->cached_rows, rest_rows);
    for (int i = 0; i < l; i++) ts.squeeze("t", late->dg + 64 * i);
    late->round = ts.round; memcpy(late->state, ts.state, 64);
    __atomic_store_n(&late->flag, epoch, __ATOMIC_RELEASE);                 // open the gate
    // then, off the critical path (the first streaming round takes ~60 us): the inverses of the first taus, one inversion for all
    uint32_t n = derive_max;
    uint64_t tau[SC_DERIVE_MAX][4], pre[SC_DERIVE_MAX][4], acc[4], inv[4];
    for (uint32_t i = 0; i < n; i++) { sp2h::fq_from_uniform(late->dg + 64 * i, tau[i]); if (!(tau[i][0] | tau[i][1] | tau[i][2] | tau[i][3])) n = 0; }
    if (n) {
      memcpy(acc, tau[0], 32);
      for (uint32_t i = 1; i < n; i++) { memcpy(pre[i], acc, 32); sp2h::mont_mul(acc, tau[i], sp2h::FQ_MOD, sp2h::FQ_INV, acc); }
      sp2h::fq_inv(acc, inv);
      for (uint32_t i = n; i-- > 1;) {
        uint64_t ti[4]; sp2h::mont_mul(inv, pre[i], sp2h::FQ_MOD, sp2h::FQ_INV, ti); memcpy(late->mail.tinv[i], ti, 32);
        sp2h::mont_mul(inv, tau[i], sp2h::FQ_MOD, sp2h::FQ_INV, inv);
      }
      memcpy(late->mail.tinv[0], inv, 32);
    }
    late->mail.n = n;
    __atomic_store_n(&late->mail.flag, epoch, __ATOMIC_RELEASE);
  });
  tp("transcript tail handed to the helper thread");


  // ---- outer sum-check (spartan.rs:293 -> sumcheck.rs:502), enqueued behind the gate while the host is still hashing ----
  proof->num_rounds_x = l; proof->num_rounds_y = nry; proof->num_comm_rows = rows; proof->num_cols = width;
  ScState *st_outer, *st_inner;
  { void *p; SP2_TRY(scratch(ctx, 14, sizeof(ScState), &p)); st_outer = (ScState *)p;
    SP2_TRY(scratch(ctx, 9, sizeof(ScState), &p)); st_inner = (ScState *)p; }
  { sp2_transcript_state hts0; memset(&hts0, 0, sizeof(hts0));               // (round, state) arrive through the gate
    const uint64_t zero4[4] = {0, 0, 0, 0};
    SP2_TRY(sc_state_upload(ctx, &st_outer, zero4, nullptr, (uint32_t)l, &hts0)); }
  // nothing between the gate's launch and its opening may need an idle device (module loading, allocations): do that now
  { static unsigned loaded_mask = 0;
    if (!(loaded_mask >> ctx->device & 1u)) { cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, (const void *)k_gate_taus); cudaFuncGetAttributes(&fa, (const void *)k_outer_to_inner);
      cudaFuncGetAttributes(&fa, (const void *)k_gate_ipa); cudaFuncGetAttributes(&fa, (const void *)k_ipa_finish); loaded_mask |= 1u << ctx->device; }
    sumcheck_cubic_preload();
    SP2_TRY(sumcheck_cubic_reserve(ctx, (uint32_t)l));
    SP2_TRY(eq_table_reserve(ctx, (uint32_t)l, 19)); }
  // (the helper thread's tail job, queued above, opens the gate; it cannot fail, and join_head waits for it on every return)
  // SP2_NO_GATES=1 (profilers that serialise kernel launches block the launching thread until a kernel ends: a gate opened by
  // that same thread would run into its 2 s bound): the gates are launched only once their flags are set
  static const bool gates = [] { const char *e = getenv("SP2_NO_GATES"); return !(e && e[0] == '1'); }();
  if (!gates && hash_here) P->worker.wait();
  { DevComm gdc; memset(&gdc, 0, sizeof(gdc)); gdc.n = 1;
    if (shard) gdc = comm->dc;
    k_gate_taus<<<1, 64, 0, ctx->stream>>>(st_outer, P->d_late, epoch, l, (int)derive_max, gdc, shard ? ((comm->dc.epoch + 1) & 0x7fffffffu) : 0u); }
  SP2_LAUNCH_CHECK();
  mark(1);
  mark(2);
  if (shard) comm->dc.epoch++;
  SP2_TRY(sumcheck_cubic_enqueue(ctx, st_outer, (uint32_t)l, P->work[0], P->work[1], P->work[2], shard ? &comm->dc : nullptr,
                                 derive_max ? &P->d_late->mail : nullptr, epoch));
  // eq(r_x) (spartan.rs:320) needs only the outer challenges: second side stream, beside the outer -> inner transition
  // multi-GPU: eq(r_x) is replicated (write-only, N entries); each rank builds its columns of poly_ABC
  fe *d_rx = shard ? P->rx : P->work[0];                                   // the sum-check consumed the tables
  SP2_CUDA_OK(cudaEventRecord(P->ev_q, ctx->stream));
  SP2_CUDA_OK(cudaStreamWaitEvent(P->side2, P->ev_q, 0));
  SP2_TRY(eq_table_dev(ctx, st_outer->r, (uint32_t)l, d_rx, P->side2, 19));
  SP2_CUDA_OK(cudaEventRecord(P->ev_eq, P->side2));
  k_outer_to_inner<<<1, 64, 0, ctx->stream>>>(st_outer, st_inner, small, nry);
  SP2_LAUNCH_CHECK();
  tp("outer sum-check enqueued (gated)");
  mark(3);

  // the PCS transcript (hyrax_pc.rs:410) re-absorbs all commitment rows; its (round, state) come from the device after the
  // sum-checks, but they enter the hash AFTER the absorbed data, so the rows are hashed now, under the sum-checks
  sp2h::Transcript t2;
  auto absorb_poly_com = [&]() {
    t2.push("poly_com", 8); t2.push("poly_commitment_begin", 21);
    t2.push(P->comm_cached_be.data(), P->comm_cached_be.size());
    for (uint64_t i = P->cached_rows; i < rows; i++) t2.push_point(proof->comm_W + 8 * i);
    t2.push("poly_commitment_end", 19);
  };
  // ---- poly_ABC (spartan.rs:321): the long columns' partial sums on the side stream beside k_abc ----
  const uint64_t inner_local = (2 * nv) >> S->shard_k;                     // this rank's share of the virtual 2M-entry tables
  SP2_TRY(abc_dev(ctx, S, d_rx, &small[S_RJOINT], P->abc, shard ? inner_local : nc, P->side2, P->ev_eq, P->ev_chunks));
  if (shard) {
    k_z_shard<<<(unsigned)((inner_local + 255) / 256), 256, 0, ctx->stream>>>(P->z, nc, P->zs, inner_local, S->shard_k, S->rank);
    SP2_LAUNCH_CHECK();
  }
  mark(4);

  // ---- inner sum-check: m+1 rounds over the virtual 2M tables (spartan.rs:330-404) ---------------
  if (shard) { comm->dc.epoch++; SP2_TRY(sumcheck_quad_enqueue(ctx, st_inner, (uint32_t)nry, P->abc, P->zs, ~0ull, P->ev_r1, &comm->dc)); }
  else SP2_TRY(sumcheck_quad_enqueue(ctx, st_inner, (uint32_t)nry, P->abc, P->z, nc, P->ev_r1));
  SP2_CUDA_OK(cudaEventRecord(P->ev_q, ctx->stream));                       // the inner sum-check is complete
  SP2_CUDA_OK(cudaStreamWaitEvent(ctx->side, P->ev_r1, 0));
  k_den_inv<<<1, 32, 0, ctx->side>>>(st_inner, small);
  SP2_LAUNCH_CHECK();
  // ---- comm_LZ early (hyrax_pc.rs:415-444): L = eq(r_y[1..log2 rows]) is complete after inner round 1 + log2(rows), long before the
  // sum-check ends — the bind LZ = L^T W, r_LZ = <L, blinds> and the 2048-term MSM run on a second side stream under the
  // remaining (latency-bound) rounds, on the SMs the pipelined round kernel leaves free
  const fe *ry1 = st_inner->r + 1;
  MsmJob j;
  if (nvr > 0) {
    if (shard && comm->in_process) {
      // ranks that are contexts of ONE process (the in-process tests): after the whole sum-check (still beside the eval_W / R-table
      // chain of the main stream) — no spinning kernel next to round kernels that wait for peers sharing the hardware queues.
      // One process per GPU (CUDA IPC, the production layout) takes the early path below like a single-GPU prove
      SP2_CUDA_OK(cudaStreamWaitEvent(P->side2, P->ev_q, 0));
    } else {
      SP2_CUDA_OK(cudaStreamWaitEvent(P->side2, P->ev_r1, 0));
      k_wait_round<<<1, 32, 0, P->side2>>>(st_inner, (u32)(1 + nvr));
      SP2_LAUNCH_CHECK();
    }
    SP2_TRY(eq_table_dev(ctx, ry1, (uint32_t)nvr, P->Ltab, P->side2, 19));
    SP2_TRY(hyrax_bind_dev(ctx, P->W, P->Ltab, rows, width, P->LZ, P->side2, 20));
    k_dot<<<1, 256, 0, P->side2>>>(P->Ltab, P->blinds, rows, &small[S_RLZ]);
    SP2_LAUNCH_CHECK();
    memset(&j, 0, sizeof(j)); j.scalars = P->LZ; j.len = (u32)width; j.nextra = 1; j.extra_base[0] = ck->idx_h(); j.extra_scalar[0] = &small[S_RLZ];
    SP2_TRY(msm_run(ctx, ck, std::vector<MsmJob>{j}, d_pts, P->side2, 21, 22));                        // comm_LZ
    SP2_CUDA_OK(cudaEventRecord(P->ev_lz, P->side2));
  }
  mark(5);

  // ---- eval_W, PCS::prove (hyrax_pc.rs:387-478; ipa.rs:125-153): two independent chains after the last inner round —
  // side stream: eval_W (needs 1 / (1 - r_y[0]) from the same stream) -> comm_eval_W;  main stream: R table -> <R, d> -> beta
  { void *chis; SP2_TRY(scratch(ctx, 12, ((size_t)4 << nvz) * sizeof(fe) + 64, &chis));
    SP2_CUDA_OK(cudaStreamWaitEvent(ctx->side, P->ev_q, 0));
    k_eval_w<<<1, 256, 0, ctx->side>>>(st_inner, P->z + nv, (u32)num_extra, m, (fe *)chis, small);
    SP2_LAUNCH_CHECK();
    memset(&j, 0, sizeof(j)); j.nextra = 2; j.extra_base[0] = ck->idx_ck_s(); j.extra_scalar[0] = &small[S_EVALW];
    j.extra_base[1] = ck->idx_h_s(); j.extra_scalar[1] = &small[S_BLIND_EVAL];
    SP2_TRY(msm_run(ctx, ck, std::vector<MsmJob>{j}, d_pts + 2, ctx->side, 16, 17));                   // comm_eval_W
    SP2_CUDA_OK(cudaEventRecord(P->ev_evalw, ctx->side)); }
  SP2_TRY(eq_table_dev(ctx, ry1 + nvr, (uint32_t)(m - nvr), P->Rtab));
  if (nvr == 0) {
    SP2_CUDA_OK(cudaMemcpyAsync(P->LZ, P->W, width * sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
    SP2_CUDA_OK(cudaMemcpyAsync(&small[S_RLZ], P->blinds, sizeof(fe), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  k_dot<<<1, 256, 0, ctx->stream>>>(P->Rtab, P->dvec, width, &small[S_IP]);
  SP2_LAUNCH_CHECK();
  memset(&j, 0, sizeof(j)); j.nextra = 2; j.extra_base[0] = ck->idx_ck_s(); j.extra_scalar[0] = &small[S_IP];
  j.extra_base[1] = ck->idx_h_s(); j.extra_scalar[1] = &small[S_RBETA];
  SP2_TRY(msm_run(ctx, ck, std::vector<MsmJob>{j}, d_pts + 3));                                         // beta
  SP2_CUDA_OK(cudaStreamWaitEvent(ctx->stream, P->ev_delta, 0));
  SP2_CUDA_OK(cudaStreamWaitEvent(ctx->stream, P->ev_evalw, 0));
  if (nvr > 0) SP2_CUDA_OK(cudaStreamWaitEvent(ctx->stream, P->ev_lz, 0));
  tp("everything up to the PCS points enqueued");
  mark(6);

  absorb_poly_com();                                                        // host hashing overlapped with the device work above
  // ---- results so far -> host -------------------------------------------------------------------
  void *hp; SP2_TRY(pinned(ctx, 2 * sizeof(ScState) + 4096, &hp));
  ScState *h_outer = (ScState *)hp, *h_inner = h_outer + 1;
  uint64_t *h_small = (uint64_t *)(h_inner + 1), *h_jac = h_small + 4 * S_COUNT;
  uint64_t h_pts[32];
  const size_t upto = offsetof(ScState, partial);
  SP2_CUDA_OK(cudaMemcpyAsync(h_outer, st_outer, upto, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(h_inner, st_inner, upto, cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(h_small, small, S_COUNT * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaMemcpyAsync(h_jac, d_pts, 4 * sizeof(jac), cudaMemcpyDeviceToHost, ctx->stream));
  tp("poly_com hashed");
  uint32_t *h_comm_err = (uint32_t *)(h_small + 4 * S_COUNT + 48);          // (pinned: between the PCS points and the second scalar read-back)
  *h_comm_err = 0;
  if (shard) SP2_CUDA_OK(cudaMemcpyAsync(h_comm_err, &comm->dc.peer[comm->dc.rank]->err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  SP2_CUDA_OK(cudaEventRecord(P->ev_q, ctx->stream));
  // the IPA response is enqueued now, behind the second gate: its launch and the read-back are already in the queue when the host
  // has the challenge
  void *d_dg; SP2_TRY(scratch(ctx, 8, (size_t)SC_MAX_ROUNDS * 64 + 64, &d_dg));
  uint64_t *h_zvec = (uint64_t *)(P->h_inbox + P->inbox_bytes + SC_MAX_ROUNDS * 64 + 64);   // pinned (the rest-row staging is free again)
  struct Gate2Guard { LateIn *late; uint32_t epoch; cudaStream_t st; bool open = false;
    ~Gate2Guard() { if (!open) { __atomic_store_n(&late->flag2, epoch | 0x80000000u, __ATOMIC_RELEASE); cudaStreamSynchronize(st); } } } gate2{late, epoch, ctx->stream};
  auto enqueue_ipa = [&]() -> int {
    k_gate_ipa<<<1, 32, 0, ctx->stream>>>(st_inner, P->d_late, epoch, (unsigned char *)d_dg);
    SP2_LAUNCH_CHECK();
    k_ipa_finish<<<(unsigned)((width + 255) / 256), 256, 0, ctx->stream>>>((const unsigned char *)d_dg, P->LZ, P->dvec, width, P->zvec, small);
    SP2_LAUNCH_CHECK();
    mark(7);
    SP2_CUDA_OK(cudaMemcpyAsync(h_zvec, P->zvec, width * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
    SP2_CUDA_OK(cudaMemcpyAsync(h_small + 4 * S_COUNT + 64, small, S_COUNT * sizeof(fe), cudaMemcpyDeviceToHost, ctx->stream));
    return SP2_OK;
  };
  if (gates) SP2_TRY(enqueue_ipa());
  SP2_CUDA_OK(cudaEventSynchronize(P->ev_q));                               // host sync 2
  tp("host sync 2 (sum-checks + PCS points)");
  if (h_outer->err || h_inner->err) {
    char msg[640];
    snprintf(msg, sizeof(msg), "prove: a device-side wait (grid barrier / gate) did not complete within 2 s [outer err %u arrived %u released %u mid %u derive %u | inner err %u arrived %u released %u mid %u r_ready %u]",
             h_outer->err, h_outer->arrived, h_outer->released, h_outer->mid_released, h_outer->derive_rounds,
             h_inner->err, h_inner->arrived, h_inner->released, h_inner->mid_released, h_inner->r_ready);
    cleanup(); return set_error(ctx, SP2_ERR_INTERNAL, msg); }
  if (*h_comm_err) { cleanup(); return set_error(ctx, SP2_ERR_INTERNAL, "sharded sum-check: a peer did not publish its round sums within 2 s (call sp2_comm_reset on every rank)"); }
  const int p_first = nvr > 0 ? 0 : 1;                                      // (no comm_LZ point for a one-row commitment)
  sp2h::batch_normalize(h_jac + 12 * p_first, 4 - p_first, h_pts + 8 * p_first);
  tp("PCS points normalised");
  if (((uint32_t *)(h_small + 4 * S_ERR))[0] == 5) { cleanup(); return set_error(ctx, SP2_ERR_DIVISION_BY_ZERO, "prove: 1 - r_y[0] = 0"); }
  for (int i = 0; i < l; i++) {                                            // compressed: [c0, c2, c3] (univariate.rs:147-153)
    memcpy(proof->outer_polys + 12 * i, &h_outer->polys[4 * i], 32);
    memcpy(proof->outer_polys + 12 * i + 4, &h_outer->polys[4 * i + 2], 64);
  }
  memcpy(proof->claims_outer, h_outer->claims, 96);
  for (int i = 0; i < nry; i++) {                                          // compressed: [c0, c2]
    memcpy(proof->inner_polys + 8 * i, &h_inner->polys[4 * i], 32);
    memcpy(proof->inner_polys + 8 * i + 4, &h_inner->polys[4 * i + 2], 32);
  }
  memcpy(proof->eval_W, h_small + 4 * S_EVALW, 32);
  memcpy(proof->blind_eval_W, rnd->blind_eval_W, 32);
  const uint64_t *comm_LZ, *p_delta, *p_ceval, *p_beta;
  comm_LZ = nvr > 0 ? h_pts : proof->comm_W; p_delta = h_pts + 8; p_ceval = h_pts + 16; p_beta = h_pts + 24;
  memcpy(proof->delta, p_delta, 64); memcpy(proof->beta, p_beta, 64);

  // ---- transcript tail on the host: poly_com rows, IPA absorbs, r (hyrax_pc.rs:410; ipa.rs:134-153) ----
  tp("proof fields copied");
  t2.set_state((uint16_t)h_inner->ts.round, h_inner->ts.state);
  t2.dom_sep("inner product argument (linear)");
  t2.push("U", 1); t2.push_point(comm_LZ); t2.push_point(p_ceval);
  t2.absorb_point("delta", p_delta);
  t2.absorb_point("beta", p_beta);
  uint8_t rdg[64];
  t2.squeeze("r", rdg);
  tp("IPA challenge squeezed");
  memcpy(late->rdg, rdg, 64);
  __atomic_store_n(&late->flag2, epoch, __ATOMIC_RELEASE);                  // open the second gate
  gate2.open = true;
  if (!gates) SP2_TRY(enqueue_ipa());
  tp("IPA challenge sent");
  SP2_CUDA_OK(cudaStreamSynchronize(ctx->stream));                          // host sync 3
  tp("host sync 3 (done)");
  if (trace_on) { for (auto &e : trace) fprintf(stderr, "[sp2 prove] %9.1f us  %s\n", e.second, e.first); }
  memcpy(proof->z_vec, h_zvec, width * sizeof(fe));
  h_small += 4 * S_COUNT + 64;
  memcpy(proof->z_delta, h_small + 4 * S_ZDELTA, 32);
  memcpy(proof->z_beta, h_small + 4 * S_ZBETA, 32);
  if (phase_ms) {
    for (int i = 0; i < 7; i++) cudaEventElapsedTime(&phase_ms[i], ev[i], ev[i + 1]);
    cudaEventElapsedTime(&phase_ms[7], ev[0], ev[7]);
  }
  cleanup();
  return SP2_OK;
}

int32_t sp2_spartan_prove(sp2_ctx *ctx, const sp2_shape *S, const sp2_ck *ck, sp2_prep *P, const uint8_t *vk_digest,
                          const uint64_t *public_values, const uint64_t *W_rest, const sp2_spartan_rand *rnd, sp2_spartan_proof *proof,
                          float *phase_ms) {
  return spartan_prove_impl(ctx, nullptr, S, ck, P, vk_digest, public_values, W_rest, rnd, proof, phase_ms);
}

/* SpartanSNARK::prove with the hypercube split across the GPUs of `comm`: every rank (one process per GPU) calls
 * this with the same inputs, its own shard of the shape (sp2_shape_upload_sharded) and the prep state made from it.  Sharded:
 * Az/Bz/Cz (rows i = rank mod G), both sum-checks (cyclic tables, <= 3 partial sums per round exchanged inside the round
 * kernels over NVLink), poly_ABC (columns j = rank mod G).  Replicated (latency-bound, cheaper to recompute than to exchange):
 * witness, commitments, eq(r_x), the Hyrax bind and the PCS MSMs.  Every rank returns the identical proof. */
int32_t sp2_spartan_prove_sharded(sp2_ctx *ctx, sp2_comm *comm, const sp2_shape *S, const sp2_ck *ck, sp2_prep *P, const uint8_t *vk_digest,
                                  const uint64_t *public_values, const uint64_t *W_rest, const sp2_spartan_rand *rnd, sp2_spartan_proof *proof,
                                  float *phase_ms) {
  if (!comm) return set_error(ctx, SP2_ERR_INTERNAL, "prove_sharded: comm is NULL");
  return spartan_prove_impl(ctx, comm, S, ck, P, vk_digest, public_values, W_rest, rnd, proof, phase_ms);
}

}  // extern "C"
