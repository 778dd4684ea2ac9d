// ws_device.cuh — device-side building blocks shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define WS_KEY_MAX 0xFFFFFFFFFFFFFFFFull
#include "ws_args.h"        // WS_TEAM, WS_CTA_THREADS, WS_TOPK_BUF

// ---- (distance, id) keys ---------------------------------------------------------------
// fp32 distance mapped to an order-preserving uint32 in the high word, id in the low word:
// one 64-bit compare gives the reference's (dist, id) lexicographic order
// (beamSearch.h:59-61).
__device__ __forceinline__ uint32_t ws_ord(float d) {
  uint32_t u = __float_as_uint(d);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ws_unord(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}
__device__ __forceinline__ uint64_t ws_key(float d, uint32_t low) {
  return ((uint64_t)ws_ord(d) << 32) | low;
}

// ---- loads -----------------------------------------------------------------------------
__device__ __forceinline__ float4 ws_ldg_f4(const float4* p) {
  float4 r;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// ---- team-of-8 distance ----------------------------------------------------------------
// Lane t (= lane & 7) of a team owns float4 columns t, t+8, t+16, ... of the row; one pass
// of the team reads 128 contiguous bytes, so a d=128 fp32 row is 4 coalesced 128 B lines.
// Accumulation order (restated by oracle/wsann_oracle.cpp in "device order" mode):
//   per lane: one fp32 accumulator, columns ascending, x,y,z,w, fused multiply-add;
//   then butterfly adds over lane^4, lane^2, lane^1.
// L2 : sum (v-q)^2  (Euclidian_Point<float>::distance, euclidian_point.h:62-65)
// MIPS: -sum v*q    (Mips_Point<float>::distance, mips_point.h:60-66)
template <int KQ, int METRIC>
__device__ __forceinline__ float ws_team_dist(const float4* __restrict__ row, const float4 (&q)[KQ],
                                              int tl, int dpad4, bool valid) {
  float4 v[KQ];
#pragma unroll
  for (int i = 0; i < KQ; i++) {
    int c = tl + WS_TEAM * i;
    v[i] = (valid && c < dpad4) ? ws_ldg_f4(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < KQ; i++) {
    if (METRIC == 0) {
      float dx = v[i].x - q[i].x, dy = v[i].y - q[i].y, dz = v[i].z - q[i].z, dw = v[i].w - q[i].w;
      acc = __fmaf_rn(dx, dx, acc);
      acc = __fmaf_rn(dy, dy, acc);
      acc = __fmaf_rn(dz, dz, acc);
      acc = __fmaf_rn(dw, dw, acc);
    } else {
      acc = __fmaf_rn(v[i].x, q[i].x, acc);
      acc = __fmaf_rn(v[i].y, q[i].y, acc);
      acc = __fmaf_rn(v[i].z, q[i].z, acc);
      acc = __fmaf_rn(v[i].w, q[i].w, acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  return METRIC == 0 ? acc : -acc;
}

// Predicate-free variant for the hot loops: `rowtl` already points at this lane's first
// float4 (row + tl) of a row that is always safe to read (callers clamp the candidate
// index instead of predicating the loads).  EXACT = the padded row is exactly 8*KQ float4s,
// so no column check is needed either.  Same accumulation order as ws_team_dist.
template <int KQ, int METRIC, bool EXACT>
__device__ __forceinline__ float ws_team_dist_nv(const float4* __restrict__ rowtl, const float4 (&q)[KQ], int tl,
                                                 int dpad4) {
  float4 v[KQ];
#pragma unroll
  for (int i = 0; i < KQ; i++) {
    if (EXACT || tl + WS_TEAM * i < dpad4) v[i] = ws_ldg_f4(rowtl + WS_TEAM * i);
    else v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < KQ; i++) {
    if (METRIC == 0) {
      float dx = v[i].x - q[i].x, dy = v[i].y - q[i].y, dz = v[i].z - q[i].z, dw = v[i].w - q[i].w;
      acc = __fmaf_rn(dx, dx, acc);
      acc = __fmaf_rn(dy, dy, acc);
      acc = __fmaf_rn(dz, dz, acc);
      acc = __fmaf_rn(dw, dw, acc);
    } else {
      acc = __fmaf_rn(v[i].x, q[i].x, acc);
      acc = __fmaf_rn(v[i].y, q[i].y, acc);
      acc = __fmaf_rn(v[i].z, q[i].z, acc);
      acc = __fmaf_rn(v[i].w, q[i].w, acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  return METRIC == 0 ? acc : -acc;
}

// Load this lane's share of the (zero-padded, smem-staged) query.
template <int KQ>
__device__ __forceinline__ void ws_load_query(const float* qs, float4 (&q)[KQ], int tl, int dpad4) {
#pragma unroll
  for (int i = 0; i < KQ; i++) {
    int c = tl + WS_TEAM * i;
    q[i] = (c < dpad4) ? reinterpret_cast<const float4*>(qs)[c] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---- CTA-wide bitonic sort of n (power of two) 64-bit keys in shared memory --------------
__device__ __forceinline__ void ws_cta_sort(uint64_t* a, int n, int tid) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < n; i += WS_CTA_THREADS) {
        int ixj = i ^ j;
        if (ixj > i) {
          uint64_t x = a[i], y = a[ixj];
          bool up = ((i & k) == 0);
          if ((x > y) == up) { a[i] = y; a[ixj] = x; }
        }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int ws_pow2ceil(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// ---- streaming top-k in shared memory ----------------------------------------------------
// buf holds [0,nbest) = best keys so far (sorted) followed by `cnt` appended keys.
// Compaction sorts everything and keeps k.  Capacity WS_TOPK_BUF keys; k <= WS_TOPK_BUF/2.
struct WsTopk {
  uint64_t* buf;   // [WS_TOPK_BUF]
  int* cnt;        // appended since last compaction (shared)
  int* nbest;      // shared
  uint64_t* tau;   // shared: current k-th best key (WS_KEY_MAX while fewer than k)
};

__device__ __forceinline__ void ws_topk_init(WsTopk& t, int tid) {
  if (tid == 0) { *t.cnt = 0; *t.nbest = 0; *t.tau = WS_KEY_MAX; }
}
// all threads of the CTA must call; contains __syncthreads
__device__ __forceinline__ void ws_topk_compact(WsTopk& t, int k, int tid) {
  __syncthreads();
  int total = *t.nbest + min(*t.cnt, WS_TOPK_BUF - *t.nbest);
  int p = ws_pow2ceil(max(total, 2));
  for (int i = total + tid; i < p; i += WS_CTA_THREADS) t.buf[i] = WS_KEY_MAX;
  __syncthreads();
  ws_cta_sort(t.buf, p, tid);
  if (tid == 0) {
    int nb = min(total, k);
    *t.nbest = nb;
    *t.cnt = 0;
    *t.tau = (nb == k) ? t.buf[k - 1] : WS_KEY_MAX;
  }
  __syncthreads();
}
// any thread; caller guarantees room (compaction is triggered before the buffer can fill)
__device__ __forceinline__ void ws_topk_push(WsTopk& t, uint64_t key) {
  int pos = atomicAdd(t.cnt, 1);
  int idx = *t.nbest + pos;
  if (idx < WS_TOPK_BUF) t.buf[idx] = key;
}

// ---- lossy-on-overflow visited set in shared memory --------------------------------------
// Open addressing, bounded probing.  Never reports an unseen id as seen (no false
// positives => no recall loss); when a probe window is full the oldest slot is
// overwritten, which can only cause a recomputation (beamSearch.h:64-73 makes the
// same trade with a direct-mapped table).
#define WS_HASH_PROBES 8
__device__ __forceinline__ uint32_t ws_hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
// returns true if `id` was already present; inserts it otherwise
__device__ __forceinline__ bool ws_seen_smem(int* table, uint32_t mask, int id) {
  uint32_t h = ws_hash32((uint32_t)id) & mask;
#pragma unroll 1
  for (int p = 0; p < WS_HASH_PROBES; p++) {
    uint32_t s = (h + p) & mask;
    int old = atomicCAS(&table[s], -1, id);
    if (old == -1) return false;
    if (old == id) return true;
  }
  table[h] = id;  // evict
  return false;
}
// exact visited bitmap in global memory (large-beam tier)
__device__ __forceinline__ bool ws_seen_bitmap(uint32_t* bits, int id) {
  uint32_t m = 1u << (id & 31);
  uint32_t old = atomicOr(&bits[id >> 5], m);
  return (old & m) != 0;
}
