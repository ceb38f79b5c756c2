// Device-resident Fiat-Shamir transcript: Keccak-256 with one 64-bit state lane per GPU lane
// (25 lanes of a warp; theta / rho-pi / chi are warp shuffles), two warps hash the lo/hi
// challenge halves concurrently.  Restates reference src/provider/keccak.rs:18-105
// (Keccak256Transcript: "NoTR"/"NoDS" framing, LE16 round counter, 64-byte lo||hi state,
// challenge = from_uniform(lo||hi)).
//
// Keeping the transcript on the device removes the per-round host round trip that every
// sum-check round of the reference ends in (sumcheck.rs:536-548).
#pragma once
#include "field.cuh"

namespace sp2 {

struct DevTranscript {          // lives in global memory
  u32 round;
  u32 pending_len;
  unsigned char state[64];
  unsigned char pending[1976];  // bytes absorbed since the last squeeze
};
static_assert(sizeof(DevTranscript) == 2048, "DevTranscript layout");

#if defined(__CUDACC__)
static __device__ __constant__ u64 KECCAK_RC_D[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};

struct KeccakLane {             // per-lane constants of the permutation
  int c1, c2, c3, c4;           // same-column lanes for theta
  int xm1, xp1, xp2;            // row neighbours
  int src;                      // rho-pi source lane
  int src1, src2;               // rho-pi source lanes of the row neighbours (x+1, y), (x+2, y): chi reads them directly
  int rot;                      // rho rotation of this lane's own value
};
__device__ __forceinline__ KeccakLane keccak_lane_init(int lane) {
  const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  int i = lane < 25 ? lane : 0;
  int x = i % 5, y = i / 5;
  KeccakLane k;
  k.c1 = (i + 5) % 25; k.c2 = (i + 10) % 25; k.c3 = (i + 15) % 25; k.c4 = (i + 20) % 25;
  k.xm1 = (x + 4) % 5 + 5 * y; k.xp1 = (x + 1) % 5 + 5 * y; k.xp2 = (x + 2) % 5 + 5 * y;
  k.src = ((x + 3 * y) % 5) + 5 * x;
  { const int x1 = (x + 1) % 5, x2 = (x + 2) % 5; k.src1 = ((x1 + 3 * y) % 5) + 5 * x1; k.src2 = ((x2 + 3 * y) % 5) + 5 * x2; }
  k.rot = ROT[i];
  return k;
}
__device__ __forceinline__ u64 rotl64(u64 v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; }

// One Keccak-f[1600] on the state distributed over lanes 0..24 of the calling warp.
__device__ __forceinline__ u64 keccak_f_warp(u64 s, const KeccakLane &k, int lane) {
  const unsigned FULL = 0xffffffffu;
#pragma unroll 1
  for (int rnd = 0; rnd < 24; rnd++) {
    u64 c = s ^ __shfl_sync(FULL, s, k.c1) ^ __shfl_sync(FULL, s, k.c2) ^ __shfl_sync(FULL, s, k.c3) ^ __shfl_sync(FULL, s, k.c4);
    u64 d = __shfl_sync(FULL, c, k.xm1) ^ rotl64(__shfl_sync(FULL, c, k.xp1), 1);
    s ^= d;
    // (A variant with theta gathered in ONE stage — ten independent gathers straight from s instead of column parities
    // followed by a dependent neighbour exchange, two dependent shuffle stages per round instead of three — was measured
    // on B200: no faster, 2.045 vs 2.014 ms per prove.  A variant with the cross-lane traffic through SHARED MEMORY
    // (store s | sync | 10 LDS.64 -> theta, rho | store | sync | 3 LDS.64 -> pi, chi; 55 instructions per round) was
    // measured SLOWER: 14.7k vs 12.2k cycles per squeeze of two permutations.  One warp cannot hide its own ALU/LDS
    // latencies: the round is a ~250-cycle dependent chain either way.)
    // rho in place, then pi and chi's two neighbour reads as ONE shuffle stage (three independent gathers from the
    // rotated values) instead of pi followed by a dependent neighbour exchange
    const u64 t = rotl64(s, k.rot);
    const u64 b = __shfl_sync(FULL, t, k.src), b1 = __shfl_sync(FULL, t, k.src1), b2 = __shfl_sync(FULL, t, k.src2);
    s = b ^ (~b1 & b2);
    if (lane == 0) s ^= KECCAK_RC_D[rnd];
  }
  return s;
}

// Keccak-256 of msg[0..len) || suffix by one warp; msg in shared (or global) memory.
// Returns the digest word (u64) in lanes 0..3.
__device__ __forceinline__ u64 keccak256_warp(const unsigned char *msg, int len, unsigned char suffix, int lane) {
  const KeccakLane k = keccak_lane_init(lane);
  const int total = len + 1;                       // message plus the one suffix byte
  const int nblocks = total / 136 + 1;             // padding always adds >= 1 byte
  u64 s = 0;
  for (int blk = 0; blk < nblocks; blk++) {
    u64 w = 0;
    if (lane < 17) {
#pragma unroll
      for (int b = 0; b < 8; b++) {
        int pos = blk * 136 + lane * 8 + b;
        unsigned v = 0;
        if (pos < len) v = msg[pos];
        else if (pos == len) v = suffix;
        if (pos == total) v ^= 0x01;               // Keccak (not SHA-3) domain padding
        if (pos == nblocks * 136 - 1) v ^= 0x80;
        w |= (u64)v << (8 * b);
      }
    }
    s ^= w;
    s = keccak_f_warp(s, k, lane);
  }
  return s;
}

// ---- single-thread Keccak-f[1600]: the whole state in registers, rounds fully unrolled inside -------
// Used by the sum-check round finaliser, where the two challenge halves are hashed by two threads of
// two different warps: lower latency than the shuffle version (no SHFL dependency chains).
__device__ __forceinline__ u64 rotl64c(u64 v, int n) { return n == 0 ? v : ((v << n) | (v >> (64 - n))); }
__device__ __forceinline__ void keccak_f_regs(u64 (&a)[25]) {
  constexpr int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
#pragma unroll 1
  for (int rnd = 0; rnd < 24; rnd++) {
    u64 c[5], d[5], b[25];
#pragma unroll
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
#pragma unroll
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rotl64c(c[(x + 1) % 5], 1);
#pragma unroll
    for (int y = 0; y < 5; y++)
#pragma unroll
      for (int x = 0; x < 5; x++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64c(a[x + 5 * y] ^ d[x], ROT[x + 5 * y]);
#pragma unroll
    for (int y = 0; y < 5; y++)
#pragma unroll
      for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= KECCAK_RC_D[rnd];
  }
}
// Keccak-256 of an already padded message of `nblocks` 136-byte blocks held as u64 words; out: 4 words
__device__ __forceinline__ void keccak256_padded(const u64 *words, int nblocks, u64 (&out)[4]) {
  u64 a[25];
#pragma unroll
  for (int i = 0; i < 25; i++) a[i] = 0;
  for (int blk = 0; blk < nblocks; blk++) {
#pragma unroll
    for (int i = 0; i < 17; i++) a[i] ^= words[blk * 17 + i];
    keccak_f_regs(a);
  }
#pragma unroll
  for (int i = 0; i < 4; i++) out[i] = a[i];
}

// squeeze (keccak.rs:70-94) executed by a 64-thread block: warp 0 -> lo, warp 1 -> hi.
// `buf` is a shared-memory scratch of >= 2048 bytes.  Result: the challenge (Montgomery form)
// is returned to every thread of warp 0... all threads via shared `out`.
__device__ __forceinline__ void ts_squeeze_block(DevTranscript *ts, const char *label, int label_len,
                                                 unsigned char *buf, fe *out /*shared*/) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ u64 digest[8];
  const int plen = ts->pending_len;
  // input = pending || "NoDS" || round_le16 || state || label
  for (int i = tid; i < plen; i += blockDim.x) buf[i] = ts->pending[i];
  if (tid < 64) buf[plen + 6 + tid] = ts->state[tid];
  if (tid == 0) {
    buf[plen + 0] = 'N'; buf[plen + 1] = 'o'; buf[plen + 2] = 'D'; buf[plen + 3] = 'S';
    buf[plen + 4] = (unsigned char)(ts->round & 0xff); buf[plen + 5] = (unsigned char)(ts->round >> 8);
    for (int i = 0; i < label_len; i++) buf[plen + 70 + i] = (unsigned char)label[i];
  }
  __syncthreads();
  const int len = plen + 70 + label_len;
  if (warp < 2) {
    u64 dg = keccak256_warp(buf, len, (unsigned char)warp, lane);
    if (lane < 4) digest[warp * 4 + lane] = dg;
  }
  __syncthreads();
  if (tid < 64) ts->state[tid] = ((unsigned char *)digest)[tid];
  if (tid == 0) {
    ts->round += 1; ts->pending_len = 0;
    fe lo, hi;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      lo.v[2 * i] = (u32)digest[i]; lo.v[2 * i + 1] = (u32)(digest[i] >> 32);
      hi.v[2 * i] = (u32)digest[4 + i]; hi.v[2 * i + 1] = (u32)(digest[4 + i] >> 32);
    }
    *out = Fq::from_uniform(lo, hi);
  }
  __syncthreads();
}

// single-thread helpers to append to the pending buffer
__device__ __forceinline__ void ts_push_bytes(DevTranscript *ts, const unsigned char *p, int n) {
  int o = ts->pending_len;
  for (int i = 0; i < n; i++) ts->pending[o + i] = p[i];
  ts->pending_len = o + n;
}
__device__ __forceinline__ void ts_push_fe_le(DevTranscript *ts, const fe &canon) {   // to_repr(): little-endian
  int o = ts->pending_len;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u32 w = canon.v[i];
    ts->pending[o + 4 * i + 0] = (unsigned char)w; ts->pending[o + 4 * i + 1] = (unsigned char)(w >> 8);
    ts->pending[o + 4 * i + 2] = (unsigned char)(w >> 16); ts->pending[o + 4 * i + 3] = (unsigned char)(w >> 24);
  }
  ts->pending_len = o + 32;
}
__device__ __forceinline__ void ts_push_fe_be(DevTranscript *ts, const fe &canon) {   // to_transcript_bytes(): big-endian
  int o = ts->pending_len;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u32 w = canon.v[7 - i];
    ts->pending[o + 4 * i + 0] = (unsigned char)(w >> 24); ts->pending[o + 4 * i + 1] = (unsigned char)(w >> 16);
    ts->pending[o + 4 * i + 2] = (unsigned char)(w >> 8); ts->pending[o + 4 * i + 3] = (unsigned char)w;
  }
  ts->pending_len = o + 32;
}
#endif  // __CUDACC__
}  // namespace sp2
