// ws_kernels.cuh — the sm_100a kernels of the window-search hot path.
//
//   K3  ws_decompose_kernel   window -> (query,node)/(query,slice) work items   (ws_decompose.h)
//   K2  ws_beam_warp_kernel   one warp per graph task (beams <= 1024): Vamana beam search + label
//       ws_beam_cta2_kernel   predicate + exponential doubling, all on device; one CTA per task above
//   K1  ws_scan_kernel        one CTA per slice task: streaming fp32 distances + top-k
//   K4  ws_merge_kernel       per query: merge partial top-k lists, decode ids, pad
//
// All four are HBM-bound integer/gather work: 128-bit coalesced loads by
// teams of 8 lanes, state in shared memory, persistent grids sized from the SM count.
#pragma once
#include "ws_args.h"
#include "ws_device.cuh"


// ------------------------------------------------------------------------------------------
// K3: decomposition
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(128) ws_decompose_kernel(WsDecompArgs A) {
  uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= A.nq) return;
  float lo = A.windows[2 * (size_t)q], hi = A.windows[2 * (size_t)q + 1];
  WsEmitter em;
  em.slots = A.tasks + (size_t)q * A.cap;
  em.cap = A.cap; em.count = 0; em.overflow = 0; em.query = q; em.beam = A.p.beam;
  em.scan_chunk = A.p.scan_chunk;
  switch (A.mode) {
    case 0: ws_decompose_fenwick(A.g, lo, hi, 0, em); break;
    case 1: ws_decompose_opt_postfilter(A.g, lo, hi, A.p, em); break;
    case 2: ws_decompose_three_split(A.g, lo, hi, A.p, em); break;
    case 3: ws_decompose_super(A.g, lo, hi, em); break;
    case WS_MODE_PREFILTER: {
      uint64_t s = ws_prefilter_bound(A.g.labels, A.g.pf_n, lo);
      uint64_t e = ws_prefilter_bound(A.g.labels, A.g.pf_n, hi);
      em.scan(s, e, lo, hi);
      break;
    }
    case WS_MODE_POSTFILTER: em.graph(A.node, lo, hi, 0); break;
  }
  A.counts[q] = em.count;
  if (em.count == 1) em.slots[0].flags |= WS_TF_SOLO;  // no merge needed: K1/K2 write the result rows
  if (em.overflow) atomicExch(A.overflow, 1u);
  uint32_t ng = 0, ns = 0;
  for (uint32_t i = 0; i < em.count; i++) (em.slots[i].node >= 0) ? ng++ : ns++;
  uint32_t gbase = ng ? atomicAdd(A.gq_count, ng) : 0;
  uint32_t sbase = ns ? atomicAdd(A.sq_count, ns) : 0;
  for (uint32_t i = 0; i < em.count; i++) {
    uint32_t slot = q * A.cap + i;
    if (em.slots[i].node >= 0) A.gq[gbase++] = slot; else A.sq[sbase++] = slot;
  }
  if (ng) atomicAdd(A.stats + WS_ST_GTASKS, (unsigned long long)ng);
  if (ns) atomicAdd(A.stats + WS_ST_STASKS, (unsigned long long)ns);
}

// ------------------------------------------------------------------------------------------
// K2: beam search
// ------------------------------------------------------------------------------------------

__device__ __forceinline__ int ws_lb_shift1(const uint64_t* a, int n, uint64_t v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if ((a[mid] >> 1) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ordered compaction rank across the CTA: returns rank of this thread's flag among all set
// flags of lower thread id (plus `base`), and the CTA total through *total.
// Contains two __syncthreads.
__device__ __forceinline__ int ws_cta_rank(bool flag, int* s_wc, int lane, int warp, int* total) {
  unsigned bal = __ballot_sync(0xffffffffu, flag);
  __syncthreads();  // previous readers of s_wc are done
  if (lane == 0) s_wc[warp] = __popc(bal);
  __syncthreads();
  int pre = 0;
#pragma unroll
  for (int w = 0; w < WS_CTA_THREADS / 32; w++) pre += (w < warp) ? s_wc[w] : 0;
  *total = s_wc[0] + s_wc[1] + s_wc[2] + s_wc[3];
  return pre + __popc(bal & ((1u << lane) - 1u));
}

// final result slot j of query q (range_filter_tree.h:84-92 / postfilter_vamana.h:207-215)
__device__ __forceinline__ void ws_write_result(uint32_t* ids, float* dists, const uint32_t* decode, size_t q, int K,
                                                int j, uint64_t key) {
  const uint32_t rank = (uint32_t)(key & 0xFFFFFFFFull);
  ids[q * K + j] = decode ? decode[rank] : rank;
  dists[q * K + j] = ws_unord((uint32_t)(key >> 32));
}
__device__ __forceinline__ void ws_write_pad(uint32_t* ids, float* dists, uint32_t pad_id, size_t q, int K, int j) {
  ids[q * K + j] = pad_id;
  dists[q * K + j] = 3.402823466e+38f;
}

// Shared-memory working set of one beam search (pointers into the CTA's dynamic smem)
struct WsBeamSmem {
  uint64_t* fr;    // [beam_cap] frontier, keys = ord(dist)<<32 | id<<1 | visited
  uint64_t* fo;    // [beam_cap] merge target (ping-pong)
  uint64_t* ck;    // [cand_cap] candidate keys
  uint64_t* ck2;   // [cand_cap] de-duplicated candidate keys
  int* cid;        // [cand_cap] kept neighbour ids
  int* cpos;       // [cand_cap] insertion ranks
  int* hash;       // [hash_mask+1] visited table (GLOBAL_SEEN == false)
  int* s_m;        // scalars
  int* s_npick;
  int* s_pick;     // [8]
  int* s_wc;       // [4]
};

struct WsSearchCfg {
  int R, E, dpad4;
  uint32_t hash_mask;
  long long limit, degree_limit;
  uint32_t* bitmap;     // GLOBAL_SEEN: this CTA's slab
};

// One beam_search (beamSearch.h:51-184) with QP.beamSize = QP.k = B from local id 0 of `node`.
// All WS_CTA_THREADS threads call it.  On return *frontier points at the final frontier
// (sorted keys) and the return value is its length.  When vis_list != nullptr the expanded
// nodes are appended to it in expansion order (keys as in the frontier) — the `visited`
// sequence the Vamana builder prunes (vamana/index.h:258-262).
template <int KQ, int METRIC, bool GLOBAL_SEEN>
__device__ __forceinline__ int ws_beam_search(const WsBeamSmem& S, const WsSearchCfg& C, const WsNode& node,
                                              const float4* vbase, const float4 (&q)[KQ], const int B,
                                              const int skip_id, uint64_t** frontier,
                                              unsigned long long* nvis_out, unsigned long long* ncmp_out,
                                              uint64_t* vis_list, const int vis_cap) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tl = lane & (WS_TEAM - 1), team = tid / WS_TEAM;
  const int NTEAMS = WS_CTA_THREADS / WS_TEAM;
  const int R = C.R, E = C.E, dpad4 = C.dpad4;

  if (!GLOBAL_SEEN) {
    for (int i = tid; i <= (int)C.hash_mask; i += WS_CTA_THREADS) S.hash[i] = -1;
  } else {
    const int words = (int)((node.count + 31u) >> 5);
    for (int i = tid; i < words; i += WS_CTA_THREADS) C.bitmap[i] = 0u;
  }
  {
    float d0 = ws_team_dist<KQ, METRIC>(vbase, q, tl, dpad4, team == 0);
    if (tid == 0) S.fr[0] = ws_key(d0, 0u);
  }
  __syncthreads();
  if (tid == 0) {
    if (!GLOBAL_SEEN) ws_seen_smem(S.hash, C.hash_mask, 0); else ws_seen_bitmap(C.bitmap, 0);
  }
  int n = 1;
  int scan_from = 0;
  unsigned long long nvis = 0, ncmp = 1;
  uint64_t* cur = S.fr;
  uint64_t* oth = S.fo;

  for (;;) {
    if ((long long)nvis >= C.limit) break;
    // ---- pick the first E unvisited frontier entries (E = 1: beamSearch.h:111)
    __syncthreads();
    if (tid == 0) *S.s_npick = 0;
    __syncthreads();
    for (int base = scan_from; base < n && *S.s_npick < E; base += WS_CTA_THREADS) {
      int i = base + tid;
      bool unv = i < n && !(cur[i] & 1ull);
      int tot;
      int prev = *S.s_npick;
      int r = prev + ws_cta_rank(unv, S.s_wc, lane, warp, &tot);
      if (unv && r < E) S.s_pick[r] = i;
      __syncthreads();
      if (tid == 0) *S.s_npick = min(E, prev + tot);
      __syncthreads();
    }
    const int npick = *S.s_npick;
    if (npick == 0) break;
    const int last_pick = S.s_pick[npick - 1];
    if (tid < npick) {
      uint64_t key = cur[S.s_pick[tid]];
      cur[S.s_pick[tid]] = key | 1ull;  // visited (beamSearch.h:114-117)
      if (vis_list != nullptr && (int)nvis + tid < vis_cap) vis_list[(int)nvis + tid] = key;
    }
    if (tid == 0) *S.s_m = 0;
    nvis += (unsigned long long)npick;
    __syncthreads();

    // ---- neighbours not seen before (beamSearch.h:123-131)
    const int items = npick * R;
    for (int base = 0; base < items; base += WS_CTA_THREADS) {
      int it = base + tid;
      int nb = -1;
      bool keep = false;
      if (it < items) {
        int e = it / R, j = it - e * R;
        uint32_t cur_id = (uint32_t)(cur[S.s_pick[e]] & 0xFFFFFFFFull) >> 1;
        if ((long long)j < C.degree_limit) nb = __ldg(node.adj + (size_t)cur_id * R + j);
        if (nb >= 0 && nb != skip_id)
          keep = GLOBAL_SEEN ? !ws_seen_bitmap(C.bitmap, nb) : !ws_seen_smem(S.hash, C.hash_mask, nb);
      }
      unsigned bal = __ballot_sync(0xffffffffu, keep);
      int wbase = 0;
      if (lane == 0 && bal) wbase = atomicAdd(S.s_m, __popc(bal));
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (keep) S.cid[wbase + __popc(bal & ((1u << lane) - 1u))] = nb;
    }
    __syncthreads();
    const int m = *S.s_m;
    if (m == 0) { scan_from = last_pick + 1; continue; }
    ncmp += (unsigned long long)m;

    // ---- distances; keep those under the cutoff (beamSearch.h:135-145)
    const float cutoff = (n < B) ? (float)2147483647 : ws_unord((uint32_t)(cur[n - 1] >> 32));
    const int mp = max(ws_pow2ceil(m), 2);
    for (int jb = 0; jb < m; jb += NTEAMS) {
      int j = jb + team;
      bool valid = j < m;
      int id = valid ? S.cid[j] : 0;
      float d = ws_team_dist<KQ, METRIC>(vbase + (size_t)id * dpad4, q, tl, dpad4, valid);
      if (valid && tl == 0) S.ck[j] = (d < cutoff) ? ws_key(d, (uint32_t)id << 1) : WS_KEY_MAX;
    }
    for (int i = m + tid; i < mp; i += WS_CTA_THREADS) S.ck[i] = WS_KEY_MAX;
    __syncthreads();
    ws_cta_sort(S.ck, mp, tid);  // beamSearch.h:148
    int mc;
    {
      int lo = 0, hi = mp;
      while (lo < hi) { int mid = (lo + hi) >> 1; if (S.ck[mid] != WS_KEY_MAX) lo = mid + 1; else hi = mid; }
      mc = lo;
    }
    if (mc == 0) { scan_from = last_pick + 1; continue; }

    // ---- rank candidates against the frontier, dropping ones already in it
    //      (the reference's set_union does the same de-duplication, beamSearch.h:151-154)
    int mc2 = 0;
    for (int base = 0; base < mc; base += WS_CTA_THREADS) {
      int j = base + tid;
      bool ok = false;
      int p = 0;
      uint64_t key = 0;
      if (j < mc) {
        key = S.ck[j];
        p = ws_lb_shift1(cur, n, key >> 1);
        ok = !(p < n && (cur[p] >> 1) == (key >> 1));
      }
      int tot;
      int r = mc2 + ws_cta_rank(ok, S.s_wc, lane, warp, &tot);
      if (ok) { S.ck2[r] = key; S.cpos[r] = p; }
      mc2 += tot;
    }
    __syncthreads();
    if (mc2 == 0) { scan_from = last_pick + 1; continue; }

    // ---- merge into the other buffer, trim to the beam (beamSearch.h:151-172)
    for (int i = tid; i < n; i += WS_CTA_THREADS) {
      uint64_t key = cur[i];
      int pos = i + ws_lb_shift1(S.ck2, mc2, key >> 1);
      if (pos < B) oth[pos] = key;
    }
    for (int j = tid; j < mc2; j += WS_CTA_THREADS) {
      int pos = S.cpos[j] + j;
      if (pos < B) oth[pos] = S.ck2[j];
    }
    const int first_new = S.cpos[0];
    n = min(n + mc2, B);
    scan_from = min(last_pick + 1, first_new);
    uint64_t* tmp = cur; cur = oth; oth = tmp;
    // loop top has the barrier that publishes `oth` writes
  }
  __syncthreads();
  *frontier = cur;
  *nvis_out = nvis;
  *ncmp_out = ncmp;
  return n;
}

// carve the dynamic shared memory of a beam-search CTA
__device__ __forceinline__ float* ws_carve_beam_smem(unsigned char* base, uint32_t beam_cap, uint32_t cand_cap,
                                                     uint32_t dpad, WsBeamSmem& S) {
  S.fr = reinterpret_cast<uint64_t*>(base);
  S.fo = S.fr + beam_cap;
  S.ck = S.fo + beam_cap;
  S.ck2 = S.ck + cand_cap;
  S.cid = reinterpret_cast<int*>(S.ck2 + cand_cap);
  S.cpos = S.cid + cand_cap;
  float* qs = reinterpret_cast<float*>(S.cpos + cand_cap);
  S.hash = reinterpret_cast<int*>(qs + dpad);
  return qs;
}

// ------------------------------------------------------------------------------------------
// K1: brute-force scan of a contiguous slice (prefiltering.h:154-204 hot loop 2,
//     range_filter_tree.h:386-397 edge scans)
// ------------------------------------------------------------------------------------------

#define WS_SCAN_UNROLL 2

template <int KQ, int METRIC>
__global__ void __launch_bounds__(WS_CTA_THREADS) ws_scan_kernel(WsScanArgs A) {
  extern __shared__ __align__(16) unsigned char ws_smem[];
  uint64_t* buf = reinterpret_cast<uint64_t*>(ws_smem);
  float* qs = reinterpret_cast<float*>(buf + WS_TOPK_BUF);
  __shared__ uint32_t s_task;
  __shared__ int s_cnt, s_nbest;
  __shared__ uint64_t s_tau;
  WsTopk tk;
  tk.buf = buf; tk.cnt = &s_cnt; tk.nbest = &s_nbest; tk.tau = &s_tau;

  const int tid = threadIdx.x, lane = tid & 31;
  const int tl = lane & (WS_TEAM - 1), team = tid / WS_TEAM;
  const int NTEAMS = WS_CTA_THREADS / WS_TEAM;
  const int dpad4 = A.dpad >> 2;
  const int K = (int)A.k;
  const int ROWS_PER_ITER = NTEAMS * WS_SCAN_UNROLL;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(A.q_head, 1u);
    __syncthreads();
    const uint32_t t = s_task;
    if (t >= *A.q_in_count) break;
    const uint32_t slot = A.q_in[t];
    const WsTask task = A.tasks[slot];
    for (int i = tid; i < (int)A.dpad; i += WS_CTA_THREADS)
      qs[i] = (i < (int)A.dim) ? A.queries[(size_t)task.query * A.dim + i] : 0.f;
    ws_topk_init(tk, tid);
    __syncthreads();
    float4 q[KQ];
    ws_load_query<KQ>(qs, q, tl, dpad4);

    for (uint32_t r0 = task.a; r0 < task.b; r0 += ROWS_PER_ITER) {
      // every pusher re-reads s_cnt after its own atomicAdd, so the last one sees the full
      // count and the OR is exact; one barrier per iteration
      const bool need = s_nbest + *(volatile int*)&s_cnt + ROWS_PER_ITER > WS_TOPK_BUF;
      if (__syncthreads_or(need)) ws_topk_compact(tk, K, tid);
      const uint64_t tau = s_tau;
      float d[WS_SCAN_UNROLL];
#pragma unroll
      for (int u = 0; u < WS_SCAN_UNROLL; u++) {
        uint32_t r = r0 + u * NTEAMS + team;
        d[u] = ws_team_dist<KQ, METRIC>(reinterpret_cast<const float4*>(A.vecs + (size_t)r * A.dpad),
                                        q, tl, dpad4, r < task.b);
      }
      if (tl == 0) {
#pragma unroll
        for (int u = 0; u < WS_SCAN_UNROLL; u++) {
          uint32_t r = r0 + u * NTEAMS + team;
          if (r < task.b) {
            uint64_t key = ws_key(d[u], r);
            if (key < tau) ws_topk_push(tk, key);
          }
        }
      }
    }
    ws_topk_compact(tk, K, tid);
    const int nb = s_nbest;
    for (int i = tid; i < nb; i += WS_CTA_THREADS) A.res_keys[(size_t)slot * K + i] = buf[i];
    if (task.flags & WS_TF_SOLO)
      for (int j = tid; j < K; j += WS_CTA_THREADS) {
        if (j < nb) ws_write_result(A.out_ids, A.out_dists, A.decode, task.query, K, j, buf[j]);
        else ws_write_pad(A.out_ids, A.out_dists, A.pad_id, task.query, K, j);
      }
    if (tid == 0) {
      A.res_cnt[slot] = (uint32_t)nb;
      atomicAdd(A.stats + WS_ST_SCANPTS, (unsigned long long)(task.b - task.a));
    }
  }
}

// ------------------------------------------------------------------------------------------
// K4: per-query merge (sort_and_truncate, range_filter_tree.h:542-549), decode
//     (range_filter_tree.h:84-92) and padding
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(WS_CTA_THREADS) ws_merge_kernel(WsMergeArgs A) {
  __shared__ uint64_t buf[WS_TOPK_BUF];
  __shared__ int s_cnt, s_nbest;
  __shared__ uint64_t s_tau;
  WsTopk tk;
  tk.buf = buf; tk.cnt = &s_cnt; tk.nbest = &s_nbest; tk.tau = &s_tau;
  const int tid = threadIdx.x;
  const int K = (int)A.k;
  for (uint32_t q = blockIdx.x; q < A.nq; q += gridDim.x) {
    __syncthreads();
    ws_topk_init(tk, tid);
    __syncthreads();
    const uint32_t nt = A.counts[q];
    if (nt == 1) continue;  // WS_TF_SOLO: the search kernel already wrote this row
    int appended = 0;
    int nbest = 0;
    for (uint32_t t = 0; t < nt; t++) {
      const size_t slot = (size_t)q * A.cap + t;
      const int c = (int)min(A.res_cnt[slot], A.k);
      if (nbest + appended + c > WS_TOPK_BUF) {
        if (tid == 0) s_cnt = appended;
        ws_topk_compact(tk, K, tid);
        nbest = s_nbest;
        appended = 0;
      }
      for (int i = tid; i < c; i += WS_CTA_THREADS) buf[nbest + appended + i] = A.res_keys[slot * K + i];
      appended += c;
    }
    __syncthreads();
    if (tid == 0) s_cnt = appended;
    ws_topk_compact(tk, K, tid);
    nbest = s_nbest;
    for (int j = tid; j < K; j += WS_CTA_THREADS) {
      uint32_t id = A.pad_id;
      float dist = 3.402823466e+38f;  // std::numeric_limits<float>::max()
      if (j < nbest) {
        uint64_t key = buf[j];
        uint32_t rank = (uint32_t)(key & 0xFFFFFFFFull);
        id = A.decode ? A.decode[rank] : rank;
        dist = ws_unord((uint32_t)(key >> 32));
      }
      A.ids[(size_t)q * K + j] = id;
      A.dists[(size_t)q * K + j] = dist;
    }
  }
}

// L2 eviction between timed steps (bench hygiene): plain streaming write
__global__ void ws_fill_kernel(uint4* p, size_t n16) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) p[i] = make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------------------------------
// K2w: warp-per-task beam search for small beams (B <= 128, R <= 64)
//
// Same algorithm, same results as ws_beam_kernel (bit-identical frontiers), but one WARP owns
// a task: no CTA barriers, every reduction is a ballot/shuffle, the 64 candidate keys are
// sorted in registers.  Many more searches are resident per SM (each needs ~10 KB of shared
// memory), which is what hides the two dependent HBM latencies of every expansion.
// ------------------------------------------------------------------------------------------
#ifndef WS_BEAM_ROWS
#define WS_BEAM_ROWS 2  // candidate rows each team of 8 lanes keeps in flight (x4 teams per warp)
#endif
#ifndef WS_BEAM_PREFETCH
#define WS_BEAM_PREFETCH 1  // 1: L2-prefetch the candidate rows beyond the first register batch; 2: also survivors' adjacency rows
#endif
#ifndef WS_BEAM_PREFETCH_NEXT
#define WS_BEAM_PREFETCH_NEXT 1  // L2-prefetch the adjacency row of the entry that will most likely be expanded next
#endif
#ifndef WS_WARP_MINBLOCKS
#define WS_WARP_MINBLOCKS 7  // resident CTAs per SM the warp kernels are register-budgeted for
#endif

__device__ __forceinline__ bool ws_seen_warp(volatile int* table, uint32_t mask, int id) {
  // plain loads/stores: lanes of one warp may race on a slot, which can only lose an
  // insertion (a later recomputation), never report an unseen id as seen
  uint32_t h = ws_hash32((uint32_t)id) & mask;
#pragma unroll 1
  for (int p = 0; p < WS_HASH_PROBES; p++) {
    uint32_t s = (h + p) & mask;
    int v = table[s];
    if (v == id) return true;
    if (v == -1) { table[s] = id; return false; }
  }
  table[h] = id;
  return false;
}

__device__ __forceinline__ uint64_t ws_shfl_up_u64(uint64_t v, int d) {
  uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)v, d);
  uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(v >> 32), d);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t ws_shfl_idx_u64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(0xffffffffu, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t ws_shfl_xor_u64(uint64_t v, int m) {
  uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m);
  uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
  return ((uint64_t)hi << 32) | lo;
}

// ascending bitonic sort of 64 keys held two per lane: element i lives in lane (i & 31),
// register (i >> 5)
__device__ __forceinline__ void ws_warp_sort64(uint64_t& k0, uint64_t& k1, int lane) {
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 32) {
        uint64_t lo = k0 < k1 ? k0 : k1, hi = k0 < k1 ? k1 : k0;
        k0 = lo; k1 = hi;
      } else {
        uint64_t o0 = ws_shfl_xor_u64(k0, j), o1 = ws_shfl_xor_u64(k1, j);
        const bool lower = (lane & j) == 0;
        const bool up0 = (lane & k) == 0;          // element lane
        const bool up1 = ((lane + 32) & k) == 0;   // element lane + 32
        k0 = (up0 == lower) ? (k0 < o0 ? k0 : o0) : (k0 < o0 ? o0 : k0);
        k1 = (up1 == lower) ? (k1 < o1 ? k1 : o1) : (k1 < o1 ? o1 : k1);
      }
    }
  }
}

// Branch-free lower bound: number of keys in a[0..n) whose (key >> 1) is < v.  STEPS = log2 of
// the capacity (7 -> n <= 128, 6 -> n <= 64); fixed trip count, predicated loads only, so the
// warp never diverges on it.
template <int STEPS>
__device__ __forceinline__ int ws_lb_fixed(const uint64_t* a, int n, uint64_t v) {
  int lo = 0;
#pragma unroll
  for (int step = 1 << (STEPS - 1); step > 0; step >>= 1) {
    const int mid = lo + step;
    const uint64_t x = (mid <= n) ? a[mid - 1] : WS_KEY_MAX;
    lo = (mid <= n && (x >> 1) < v) ? mid : lo;
  }
  // capacity itself (n == 1 << STEPS) needs one more probe
  if ((1 << STEPS) <= n) {
    const uint64_t x = a[(1 << STEPS) - 1];
    lo = (lo == (1 << STEPS) - 1 && (x >> 1) < v) ? (1 << STEPS) : lo;
  }
  return lo;
}

// Visited-set probe for two ids per lane with warp-uniform control flow: every lane walks
// the probe sequence together (votes decide when to stop), loads and stores are predicated.
// Same guarantees as ws_seen_warp: never reports an unseen id as seen; a lost insertion or
// an eviction can only cause a recomputation.
__device__ __forceinline__ void ws_seen_warp2(volatile int* table, uint32_t mask, int id0, bool& keep0, int id1,
                                              bool& keep1) {
  const uint32_t h0 = ws_hash32((uint32_t)id0) & mask, h1 = ws_hash32((uint32_t)id1) & mask;
  bool pend0 = keep0, pend1 = keep1;
#pragma unroll 1
  for (int p = 0; p < WS_HASH_PROBES; p++) {
    if (!__any_sync(0xffffffffu, pend0 || pend1)) break;
    const uint32_t s0 = (h0 + p) & mask, s1 = (h1 + p) & mask;
    const int v0 = pend0 ? table[s0] : 0;
    const bool hit0 = pend0 && v0 == id0, empty0 = pend0 && v0 == -1;
    if (empty0) table[s0] = id0;
    keep0 = keep0 && !hit0;
    pend0 = pend0 && !hit0 && !empty0;
    __syncwarp();
    const int v1 = pend1 ? table[s1] : 0;
    const bool hit1 = pend1 && v1 == id1, empty1 = pend1 && v1 == -1;
    if (empty1) table[s1] = id1;
    keep1 = keep1 && !hit1;
    pend1 = pend1 && !hit1 && !empty1;
    __syncwarp();
  }
  if (pend0) table[h0] = id0;  // probe window full: evict
  if (pend1) table[h1] = id1;
}

// 16-bit variant: the table is indexed by the id's low bits (identity hash), so a slot only has
// to remember the id's high bits plus the probe offset it was stored at — exact membership in
// half the shared memory.  Valid while id < 2^(bits+12), i.e. nodes of up to 8 M points with
// 2048 slots; larger nodes use the 32-bit table.
__device__ __forceinline__ void ws_seen_warp2_h16(volatile unsigned short* table, uint32_t mask, int bits, int id0,
                                                  bool& keep0, int id1, bool& keep1) {
  const uint32_t h0 = (uint32_t)id0 & mask, h1 = (uint32_t)id1 & mask;
  const uint32_t t0 = ((uint32_t)id0 >> bits) << 3, t1 = ((uint32_t)id1 >> bits) << 3;
  bool pend0 = keep0, pend1 = keep1;
#pragma unroll 1
  for (int p = 0; p < WS_HASH_PROBES; p++) {
    if (!__any_sync(0xffffffffu, pend0 || pend1)) break;
    const uint32_t s0 = (h0 + p) & mask, s1 = (h1 + p) & mask;
    const uint32_t v0 = pend0 ? table[s0] : 0u;
    const bool hit0 = pend0 && v0 == (t0 | p), empty0 = pend0 && v0 == 0xFFFFu;
    if (empty0) table[s0] = (unsigned short)(t0 | p);
    keep0 = keep0 && !hit0;
    pend0 = pend0 && !hit0 && !empty0;
    __syncwarp();
    const uint32_t v1 = pend1 ? table[s1] : 0u;
    const bool hit1 = pend1 && v1 == (t1 | p), empty1 = pend1 && v1 == 0xFFFFu;
    if (empty1) table[s1] = (unsigned short)(t1 | p);
    keep1 = keep1 && !hit1;
    pend1 = pend1 && !hit1 && !empty1;
    __syncwarp();
  }
  if (pend0) table[h0] = (unsigned short)t0;  // probe window full: evict
  if (pend1) table[h1] = (unsigned short)t1;
}

// ascending bitonic sort of 32 keys, one per lane
__device__ __forceinline__ void ws_warp_sort32(uint64_t& k0, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t o = ws_shfl_xor_u64(k0, j);
      const bool lower = (lane & j) == 0;
      const bool up = (lane & k) == 0 || k == 32;
      k0 = (up == lower) ? (k0 < o ? k0 : o) : (k0 < o ? o : k0);
    }
  }
}


template <int STEPS>
__device__ __forceinline__ int ws_lb_fixed_raw(const uint64_t* a, int n, uint64_t v) {
  int lo = 0;
#pragma unroll
  for (int step = 1 << (STEPS - 1); step > 0; step >>= 1) {
    const int mid = lo + step;
    const uint64_t x = (mid <= n) ? a[mid - 1] : WS_KEY_MAX;
    lo = (mid <= n && x < v) ? mid : lo;
  }
  if ((1 << STEPS) <= n) {
    const uint64_t x = a[(1 << STEPS) - 1];
    lo = (lo == (1 << STEPS) - 1 && x < v) ? (1 << STEPS) : lo;
  }
  return lo;
}

// this lane's share of query `qi` (rows of the batch are unpadded [nq][dim])
template <int KQ, bool EXACT>
__device__ __forceinline__ void ws_load_query_global(const float* queries, uint32_t dim, uint32_t dpad, uint32_t qi, int tl,
                                                     float4 (&q)[KQ]) {
  const float* qrow = queries + (size_t)qi * dim;
  if (EXACT && dim == dpad) {  // 16-byte aligned, unpadded rows
    const float4* q4 = reinterpret_cast<const float4*>(qrow) + tl;
#pragma unroll
    for (int i = 0; i < KQ; i++) q[i] = __ldg(q4 + WS_TEAM * i);
  } else {
#pragma unroll
    for (int i = 0; i < KQ; i++) {
      const int c = (tl + WS_TEAM * i) * 4;
      q[i].x = (c + 0 < (int)dim) ? __ldg(qrow + c + 0) : 0.f;
      q[i].y = (c + 1 < (int)dim) ? __ldg(qrow + c + 1) : 0.f;
      q[i].z = (c + 2 < (int)dim) ? __ldg(qrow + c + 2) : 0.f;
      q[i].w = (c + 3 < (int)dim) ? __ldg(qrow + c + 3) : 0.f;
    }
  }
}

// One brute-force scan task (query, rows [a,b)) by one warp: streams the slice with 16 rows in
// flight, keeps the running top-B in `fr` (in-place merges of cutoff survivors) and writes the
// partial (or, for WS_TF_SOLO tasks, final) result.  Shared by the scan kernel and by the beam
// warp kernel, which drains the scan queue with the same warps once the graph queue is empty.
struct WsScanOut {
  const float* vecs;
  uint32_t dpad;
  uint64_t* res_keys;
  uint32_t* res_cnt;
  unsigned long long* stats;
  uint32_t* out_ids;
  float* out_dists;
  const uint32_t* decode;
  uint32_t pad_id;
};

template <int KQ, int METRIC, bool EXACT>
__device__ __forceinline__ void ws_scan_task(const WsScanOut& A, const WsTask& task, uint32_t slot, const float4 (&q)[KQ],
                                             int B, uint64_t* fr, uint64_t* sk, uint64_t* sk2, int* cpos) {
  const int lane = threadIdx.x & 31;
  const int tl = lane & (WS_TEAM - 1), team = lane / WS_TEAM;
  const int dpad4 = A.dpad >> 2;
  const unsigned lt = (1u << lane) - 1u;
  const bool leader = tl == 0;
  const float4* vtl = reinterpret_cast<const float4*>(A.vecs) + tl;
  int n = 0, s = 0;
  uint64_t cutoff = WS_KEY_MAX;  // k-th best key once the list is full
  const uint32_t a = task.a, b = task.b;
  for (uint32_t r0 = a; r0 < b; r0 += 16) {
    uint32_t r[4];
    float d[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      r[u] = r0 + 4 * u + team;
      const uint32_t rc = r[u] < b ? r[u] : b - 1;
      d[u] = ws_team_dist_nv<KQ, METRIC, EXACT>(vtl + (size_t)rc * dpad4, q, tl, dpad4);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint64_t key = ws_key(d[u], r[u]);
      const bool pass = leader && r[u] < b && key < cutoff;
      const unsigned bal = __ballot_sync(0xffffffffu, pass);
      if (pass) sk[s + __popc(bal & lt)] = key;
      s += __popc(bal);
    }
    if (s <= 48 && r0 + 16 < b) continue;
    if (s == 0) continue;
    // ---- fold the survivors into the running top-k
    __syncwarp();
    uint64_t k0 = lane < s ? sk[lane] : WS_KEY_MAX, k1 = WS_KEY_MAX;
    if (s > 32) {
      k1 = lane + 32 < s ? sk[lane + 32] : WS_KEY_MAX;
      ws_warp_sort64(k0, k1, lane);
    } else {
      ws_warp_sort32(k0, lane);
    }
    const int p0 = ws_lb_fixed_raw<7>(fr, n, k0), p1 = ws_lb_fixed_raw<7>(fr, n, k1);
    const bool ok0 = k0 != WS_KEY_MAX, ok1 = k1 != WS_KEY_MAX;
    const int mc2 = s;  // survivors are distinct rows: nothing to de-duplicate
    if (ok0) { sk2[lane] = k0; cpos[lane] = p0; }
    if (ok1) { sk2[lane + 32] = k1; cpos[lane + 32] = p1; }
    __syncwarp();
    const int first_new = cpos[0];
    uint64_t e[4];
    int np[4];
#pragma unroll
    for (int rr = 0; rr < 4; rr++) {
      const int i = lane + 32 * rr;
      const bool mv = i >= first_new && i < n;
      e[rr] = fr[i];
      const int c = ws_lb_fixed_raw<6>(sk2, mc2, e[rr]);
      np[rr] = mv ? i + c : B;
    }
    const int j1 = lane + 32;
    const uint64_t c0 = sk2[lane], c1 = sk2[j1];
    const int q0 = lane < mc2 ? cpos[lane] + lane : B, q1 = j1 < mc2 ? cpos[j1] + j1 : B;
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < 4; rr++)
      if (np[rr] < B) fr[np[rr]] = e[rr];
    if (q0 < B) fr[q0] = c0;
    if (q1 < B) fr[q1] = c1;
    n = min(n + mc2, B);
    s = 0;
    __syncwarp();
    cutoff = (n == B) ? fr[B - 1] : WS_KEY_MAX;
  }
  __syncwarp();
  for (int i = lane; i < n; i += 32) A.res_keys[(size_t)slot * B + i] = fr[i];
  if (task.flags & WS_TF_SOLO)
    for (int j = lane; j < B; j += 32) {
      if (j < n) ws_write_result(A.out_ids, A.out_dists, A.decode, task.query, B, j, fr[j]);
      else ws_write_pad(A.out_ids, A.out_dists, A.pad_id, task.query, B, j);
    }
  if (lane == 0) {
    A.res_cnt[slot] = (uint32_t)n;
    atomicAdd(A.stats + WS_ST_SCANPTS, (unsigned long long)(b - a));
  }
  __syncwarp();
}

// CS = log2 of the largest beam the instantiation can hold (7 -> 128, 8 -> 256, 9 -> 512, 10 -> 1024).
// CS <= 8: the merge holds the whole frontier tail in registers; above, the tail is shifted in place from the end,
// 32 entries at a time.  The 512 / 1024 instantiations serve the doubling tail of optimized postfiltering
// (postfilter_vamana.h:161-172): those tasks are few and long, i.e. bound by the latency of one expansion, not by
// bytes (an expansion adds ~7 unseen neighbours at these beams), so a warp per task — thousands of searches in
// flight, no CTA barrier on the dependent chain — beats a CTA per task (round 1: one 256-thread CTA per SM, 3 % of
// SM-time active, profiles/r01_ncu_beam_cta2_kernel_raw.csv).
template <int KQ, int METRIC, bool EXACT, int CS>
__global__ void __launch_bounds__(WS_WARPS_PER_CTA * 32, (CS <= 7 ? WS_WARP_MINBLOCKS : 2)) ws_beam_warp_kernel(WsBeamArgs A) {
  extern __shared__ __align__(16) unsigned char ws_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane & (WS_TEAM - 1), team = lane / WS_TEAM;  // 4 teams per warp
  const uint32_t CAP = A.beam_cap;
  unsigned char* base = ws_smem + ws_warp_smem_bytes(CAP, A.hash_mask + 1, A.hash16) * warp;
  uint64_t* fr = reinterpret_cast<uint64_t*>(base);  // [CAP] frontier, updated in place
  uint64_t* sk = fr + CAP;                           // [64] candidates under the cutoff
  uint64_t* sk2 = sk + 64;                           // [64] sorted, de-duplicated
  int* cpos = reinterpret_cast<int*>(sk2 + 64);      // [64] insertion ranks
  int* cid = cpos + 64;                              // [64] kept neighbour ids
  volatile int* hash = cid + 64;                     // [hash_mask + 1] ids, or 16-bit tags (A.hash16)
  volatile unsigned short* hash16 = reinterpret_cast<volatile unsigned short*>(cid + 64);
  const bool h16 = A.hash16 != 0;
  const int hbits = 32 - __clz(A.hash_mask);         // log2(entries)

  const int dpad4 = A.dpad >> 2;
  const int K = (int)A.k;
  const int R = (int)A.R;
  const unsigned lt = (1u << lane) - 1u;
  const bool leader = tl == 0;

  // A handful of escalated (long) tasks is better served by the CTA-per-task tier, where eight
  // warps share one expansion; a warp per task only pays off when there are enough of them.
  if (A.min_tasks != 0 && A.q_out != nullptr && *A.q_in_count < A.min_tasks) {
    if (blockIdx.x == 0) {
      const uint32_t cnt = *A.q_in_count;
      for (uint32_t i = threadIdx.x; i < cnt; i += blockDim.x) A.q_out[atomicAdd(A.q_out_count, 1u)] = A.q_in[i];
    }
    return;
  }

  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(A.q_head, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= *A.q_in_count) break;
    const uint32_t slot = A.q_in[t];
    const WsTask task = A.tasks[slot];
    const WsNode node = A.nodes[task.node];
    float4 q[KQ];
    ws_load_query_global<KQ, EXACT>(A.queries, A.dim, A.dpad, task.query, tl, q);
    const float4* vbase_tl = reinterpret_cast<const float4*>(A.vecs + (size_t)node.start * A.dpad) + tl;
    const int skip_id = A.skip_query_id ? (int)(task.query + A.query_id_base) : -1;

    long long beam = task.beam;
    int phase = (task.flags & WS_TF_FINAL) ? 1 : 0;
    const long long mult = (task.flags & WS_TF_MULT1) ? 1 : A.final_mult;
    int have = 0;
    bool escalate = false;
    if (!(task.flags & WS_TF_RESUMED) && lane == 0) A.res_cnt[slot] = 0;

    for (;;) {  // PostfilterVamanaIndex::query (postfilter_vamana.h:141-188)
      if (phase == 0) {
        if (!(have < K && beam < A.max_beam)) {
          long long fin = beam * mult;
          if (fin > A.max_beam) fin = A.max_beam;
          if (fin > beam) { beam = fin; phase = 1; } else break;
        }
      }
      if (beam > (long long)CAP) { escalate = true; break; }
      const int B = (int)beam;

      // ---- beam_search (beamSearch.h:51-184), QP.beamSize = QP.k = B, start = local id 0
      if (h16) {  // 0xFFFF = empty; cleared two slots per store
        for (int i = lane; i <= (int)(A.hash_mask >> 1); i += 32) hash[i] = -1;
      } else {
        for (int i = lane; i <= (int)A.hash_mask; i += 32) hash[i] = -1;
      }
      {
        const float d0 = ws_team_dist_nv<KQ, METRIC, EXACT>(vbase_tl, q, tl, dpad4);
        if (lane == 0) fr[0] = ws_key(d0, 0u);
      }
      __syncwarp();
      if (lane == 0) {  // the start point counts as seen
        if (h16) hash16[0] = 0; else hash[ws_hash32(0u) & A.hash_mask] = 0;
      }
      __syncwarp();
      int n = 1, scan_from = 0;
      unsigned long long nvis = 0, ncmp = 1;

      for (;;) {
        if ((long long)nvis >= A.limit) break;
        // first unvisited frontier entry (beamSearch.h:111)
        int pick = -1;
        for (int b0 = scan_from; b0 < n; b0 += 32) {
          const int i = b0 + lane;
          const bool unv = i < n && !(fr[i] & 1ull);
          const unsigned bal = __ballot_sync(0xffffffffu, unv);
          if (bal) { pick = b0 + __ffs(bal) - 1; break; }
        }
        if (pick < 0) break;
        const uint64_t pkey = fr[pick];
        const uint32_t cur_id = (uint32_t)(pkey & 0xFFFFFFFFull) >> 1;
        __syncwarp();
        if (lane == 0) fr[pick] = pkey | 1ull;  // visited (beamSearch.h:114-117)
        nvis++;
        if (WS_BEAM_PREFETCH_NEXT) {
          // The entry expanded NEXT is, unless this expansion inserts something in front of it, the next
          // unvisited one: pull its adjacency row (2 x 128 B) into L2 now, so that the first of the two
          // dependent loads of the next expansion is an L2 hit.  A hint only: a wrong guess costs 256 B.
          int nxt = -1;
          for (int b0 = pick + 1; b0 < n && b0 < pick + 65; b0 += 32) {
            const int i = b0 + lane;
            const bool unv = i < n && !(fr[i] & 1ull);
            const unsigned bal = __ballot_sync(0xffffffffu, unv);
            if (bal) { nxt = b0 + __ffs(bal) - 1; break; }
          }
          if (nxt >= 0 && lane < 2) {
            const uint32_t nid = (uint32_t)(fr[nxt] & 0xFFFFFFFFull) >> 1;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(node.adj + (size_t)nid * R) + 128 * lane));
          }
        }

        // neighbours not seen before (beamSearch.h:123-131); two per lane
        int nb0 = -1, nb1 = -1;
        {
          const int* arow = node.adj + (size_t)cur_id * R;
          if (lane < R && (long long)lane < A.degree_limit) nb0 = __ldg(arow + lane);
          if (lane + 32 < R && (long long)(lane + 32) < A.degree_limit) nb1 = __ldg(arow + lane + 32);
        }
        bool keep0 = nb0 >= 0 && nb0 != skip_id;
        bool keep1 = nb1 >= 0 && nb1 != skip_id;
        if (h16) ws_seen_warp2_h16(hash16, A.hash_mask, hbits, nb0, keep0, nb1, keep1);
        else ws_seen_warp2(hash, A.hash_mask, nb0, keep0, nb1, keep1);
        const unsigned bal0 = __ballot_sync(0xffffffffu, keep0), bal1 = __ballot_sync(0xffffffffu, keep1);
        const int m0 = __popc(bal0), m = m0 + __popc(bal1);
        if (m == 0) { scan_from = pick + 1; continue; }
        if (keep0) cid[__popc(bal0 & lt)] = nb0;
        if (keep1) cid[m0 + __popc(bal1 & lt)] = nb1;
        __syncwarp();
        ncmp += (unsigned long long)m;

        // Only 4*WS_BEAM_ROWS candidate rows fit in registers at a time; pull the later ones
        // towards L2 now so that their loads find them there (4 x 128 B lines per 512 B row,
        // one line per lane; costs no registers and no extra DRAM traffic).
        if (WS_BEAM_PREFETCH) {
          const int lines = ((int)A.dpad * 4 + 127) >> 7;  // 128 B lines per row
          for (int j = 4 * WS_BEAM_ROWS + (lane >> 2); j < m; j += 8) {
            const char* rowp = reinterpret_cast<const char*>(vbase_tl - tl + (size_t)cid[j] * dpad4);
            for (int l = lane & 3; l < lines; l += 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + 128 * l));
          }
        }

        // distances (4 teams x 2 rows in flight); keep those under the cutoff (beamSearch.h:135-145)
        const float cutoff = (n < B) ? (float)2147483647 : ws_unord((uint32_t)(fr[n - 1] >> 32));
        int s = 0;
        for (int jb = 0; jb < m; jb += 4 * WS_BEAM_ROWS) {
          int idu[WS_BEAM_ROWS];
          float du[WS_BEAM_ROWS];
#pragma unroll
          for (int u = 0; u < WS_BEAM_ROWS; u++) {
            idu[u] = cid[min(jb + 4 * u + team, m - 1)];
            du[u] = ws_team_dist_nv<KQ, METRIC, EXACT>(vbase_tl + (size_t)idu[u] * dpad4, q, tl, dpad4);
          }
#pragma unroll
          for (int u = 0; u < WS_BEAM_ROWS; u++) {
            const bool pu = leader && jb + 4 * u + team < m && du[u] < cutoff;
            const unsigned bu = __ballot_sync(0xffffffffu, pu);
            if (pu) {
              sk[s + __popc(bu & lt)] = ws_key(du[u], (uint32_t)idu[u] << 1);
              if (WS_BEAM_PREFETCH >= 2) {  // a candidate that enters the beam is likely to be expanded: warm its adjacency row
                const char* ar = reinterpret_cast<const char*>(node.adj + (size_t)idu[u] * R);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ar));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ar + 128));
              }
            }
            s += __popc(bu);
          }
        }
        if (s == 0) { scan_from = pick + 1; continue; }
        __syncwarp();

        // sort the survivors (beamSearch.h:148), then rank them against the frontier, dropping
        // ones already in it (the reference's set_union de-duplicates the same way, :151-154)
        uint64_t k0 = lane < s ? sk[lane] : WS_KEY_MAX, k1 = WS_KEY_MAX;
        if (s > 32) {
          k1 = lane + 32 < s ? sk[lane + 32] : WS_KEY_MAX;
          ws_warp_sort64(k0, k1, lane);
        } else {
          ws_warp_sort32(k0, lane);
        }
        // a row can list a neighbour twice (graph.h:85-95 appends without de-duplication) and
        // the racy visited table may let both copies through: equal keys are adjacent now
        const uint64_t up0 = ws_shfl_up_u64(k0, 1);
        uint64_t up1 = ws_shfl_up_u64(k1, 1);
        const uint64_t last0 = ws_shfl_idx_u64(k0, 31);
        up1 = lane == 0 ? last0 : up1;
        const int p0 = ws_lb_fixed<CS>(fr, n, k0 >> 1), p1 = ws_lb_fixed<CS>(fr, n, k1 >> 1);
        const uint64_t f0 = fr[min(p0, n - 1)], f1 = fr[min(p1, n - 1)];
        const bool ok0 = k0 != WS_KEY_MAX && !(lane > 0 && up0 == k0) && !(p0 < n && (f0 >> 1) == (k0 >> 1));
        const bool ok1 = k1 != WS_KEY_MAX && up1 != k1 && !(p1 < n && (f1 >> 1) == (k1 >> 1));
        const unsigned bka = __ballot_sync(0xffffffffu, ok0), bkb = __ballot_sync(0xffffffffu, ok1);
        const int ca = __popc(bka), mc2 = ca + __popc(bkb);
        if (mc2 == 0) { scan_from = pick + 1; continue; }
        if (ok0) { const int r = __popc(bka & lt); sk2[r] = k0; cpos[r] = p0; }
        if (ok1) { const int r = ca + __popc(bkb & lt); sk2[r] = k1; cpos[r] = p1; }
        __syncwarp();

        // merge in place, trim to the beam (beamSearch.h:151-172): entries before the first
        // insertion point stay where they are
        const int first_new = cpos[0];
        const int j1 = lane + 32;
        const uint64_t c0 = sk2[lane], c1 = sk2[j1];
        const int q0 = lane < mc2 ? cpos[lane] + lane : B, q1 = j1 < mc2 ? cpos[j1] + j1 : B;
        if constexpr (CS <= 8) {
          constexpr int NE = (1 << CS) / 32;  // frontier entries per lane
          uint64_t e[NE];
          int np[NE];
#pragma unroll
          for (int r = 0; r < NE; r++) {
            const int i = lane + 32 * r;
            const bool mv = i >= first_new && i < n;
            e[r] = fr[min(i, (int)CAP - 1)];
            const int c = ws_lb_fixed<6>(sk2, mc2, e[r] >> 1);
            np[r] = mv ? i + c : B;
          }
          __syncwarp();
#pragma unroll
          for (int r = 0; r < NE; r++)
            if (np[r] < B) fr[np[r]] = e[r];
        } else {
          // Shift the tail [first_new, n) right, from the end, 32 entries at a time.  Entry i moves to
          // i + (survivors below it) >= i, so a chunk never writes below its own start: the chunks still to come
          // (lower indices) read untouched entries, and one barrier per chunk (reads before writes) suffices.
          for (int hi = n; hi > first_new; hi -= 32) {
            const int i = hi - 1 - lane;
            const bool valid = i >= first_new;
            uint64_t key = 0;
            int c = 0;
            if (valid) {
              key = fr[i];
              c = ws_lb_fixed<6>(sk2, mc2, key >> 1);
            }
            __syncwarp();
            if (valid && i + c < B) fr[i + c] = key;
          }
          __syncwarp();
        }
        if (q0 < B) fr[q0] = c0;
        if (q1 < B) fr[q1] = c1;
        n = min(n + mc2, B);
        scan_from = min(pick + 1, first_new);
        __syncwarp();
      }

      // raw_query's label predicate, closed interval (postfilter_vamana.h:234-251)
      have = 0;
      for (int b0 = 0; b0 < n && have < K; b0 += 32) {
        const int i = b0 + lane;
        bool in = false;
        uint64_t key = 0;
        uint32_t rank = 0;
        if (i < n) {
          key = fr[i];
          rank = node.start + ((uint32_t)(key & 0xFFFFFFFFull) >> 1);
          const float lab = __ldg(A.labels + rank);
          in = (lab >= task.lo) && (lab <= task.hi);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, in);
        const int r = have + __popc(bal & lt);
        if (in && r < K) {
          const uint64_t okey = (key & 0xFFFFFFFF00000000ull) | rank;
          A.res_keys[(size_t)slot * K + r] = okey;
          if (task.flags & WS_TF_SOLO) ws_write_result(A.out_ids, A.out_dists, A.decode, task.query, K, r, okey);
        }
        have += __popc(bal);
      }
      if (lane == 0) {
        A.res_cnt[slot] = (uint32_t)min(have, K);
        atomicAdd(A.stats + WS_ST_SEARCHES, 1ull);
        atomicAdd(A.stats + WS_ST_VISITED, nvis);
        atomicAdd(A.stats + WS_ST_DISTCMPS, ncmp);
        atomicAdd(A.stats + WS_ST_BEAMSUM, (unsigned long long)B);
      }
      __syncwarp();
      if (phase == 1) break;
      if (have < K) beam *= 2;
    }

    if (!escalate && (task.flags & WS_TF_SOLO))
      for (int j = min(have, K) + lane; j < K; j += 32) ws_write_pad(A.out_ids, A.out_dists, A.pad_id, task.query, K, j);
    if (escalate && lane == 0) {
      if (A.q_out != nullptr) {
        A.tasks[slot].beam = (uint32_t)beam;
        A.tasks[slot].flags = task.flags | WS_TF_RESUMED | (phase ? WS_TF_FINAL : 0u);
        const uint32_t pos = atomicAdd(A.q_out_count, 1u);
        A.q_out[pos] = slot;
        atomicAdd(A.stats + WS_ST_ESCALATED, 1ull);
      } else if (A.sticky != nullptr) {
        atomicOr(A.sticky, 2u);  // no tier left: the host reports it (ws_index_sync / end of a host-buffer batch)
      }
    }
  }

  // ---- graph queue empty: the same warps drain the batch's brute-force scan tasks (fenwick
  //      edges), so the short scans fill the tail of the graph searches instead of a launch of their own
  if (A.sq_in != nullptr) {
    WsScanOut so;
    so.vecs = A.vecs; so.dpad = A.dpad; so.res_keys = A.res_keys; so.res_cnt = A.res_cnt; so.stats = A.stats;
    so.out_ids = A.out_ids; so.out_dists = A.out_dists; so.decode = A.decode; so.pad_id = A.pad_id;
    for (;;) {
      uint32_t t = 0;
      if (lane == 0) t = atomicAdd(A.sq_head, 1u);
      t = __shfl_sync(0xffffffffu, t, 0);
      if (t >= *A.sq_count) break;
      const uint32_t slot = A.sq_in[t];
      const WsTask task = A.tasks[slot];
      float4 q[KQ];
      ws_load_query_global<KQ, EXACT>(A.queries, A.dim, A.dpad, task.query, tl, q);
      ws_scan_task<KQ, METRIC, EXACT>(so, task, slot, q, K, fr, sk, sk2, cpos);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1w: warp-per-task brute-force scan (k <= 128)
//
// Streams a contiguous slice with 16 rows (8 KB) in flight per warp and keeps the running
// top-k in shared memory with the same machinery as the beam kernel: rows whose key is under
// the current k-th key are ballot-compacted, sorted in registers and merged in place.
// No CTA barriers; ~2.5 KB of shared memory per warp.
// ------------------------------------------------------------------------------------------
#ifndef WS_SCAN_MINBLOCKS
#define WS_SCAN_MINBLOCKS 6
#endif

template <int KQ, int METRIC, bool EXACT>
__global__ void __launch_bounds__(WS_WARPS_PER_CTA * 32, WS_SCAN_MINBLOCKS) ws_scan_warp_kernel(WsScanArgs A) {
  __shared__ uint64_t s_fr[WS_WARPS_PER_CTA][128];
  __shared__ uint64_t s_sk[WS_WARPS_PER_CTA][64];
  __shared__ uint64_t s_sk2[WS_WARPS_PER_CTA][64];
  __shared__ int s_cpos[WS_WARPS_PER_CTA][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane & (WS_TEAM - 1), team = lane / WS_TEAM;
  uint64_t* fr = s_fr[warp];
  uint64_t* sk = s_sk[warp];
  uint64_t* sk2 = s_sk2[warp];
  int* cpos = s_cpos[warp];
  const int dpad4 = A.dpad >> 2;
  const int B = (int)A.k;
  const unsigned lt = (1u << lane) - 1u;
  const bool leader = tl == 0;

  for (;;) {
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(A.q_head, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= *A.q_in_count) break;
    const uint32_t slot = A.q_in[t];
    const WsTask task = A.tasks[slot];
    float4 q[KQ];
    ws_load_query_global<KQ, EXACT>(A.queries, A.dim, A.dpad, task.query, tl, q);
    WsScanOut so;
    so.vecs = A.vecs; so.dpad = A.dpad; so.res_keys = A.res_keys; so.res_cnt = A.res_cnt; so.stats = A.stats;
    so.out_ids = A.out_ids; so.out_dists = A.out_dists; so.decode = A.decode; so.pad_id = A.pad_id;
    ws_scan_task<KQ, METRIC, EXACT>(so, task, slot, q, B, fr, sk, sk2, cpos);
  }
}


// ------------------------------------------------------------------------------------------
// K1d: PrefilterIndex::batch_search in ONE launch for batches of small windows
//      (prefiltering.h:124-204).  With a few hundred points per window the batch is bound by the
//      launch chain (memset -> K3 -> K1 -> K4: ~60 us for 10 000 queries), not by bytes, so here a
//      warp does everything for its query: the two r = n-1 binary searches on the labels
//      (prefiltering.h:159-184), the streaming scan of [start, end) (ws_scan_task, the arithmetic
//      and folding of K1 — rows are bit-identical), decode, padding and the final row.  No task
//      slots, no queues, no control words.  Any window size is answered correctly (the warp
//      streams the whole window); the host only routes batches here when windows are small.
// ------------------------------------------------------------------------------------------

template <int KQ, int METRIC, bool EXACT>
__global__ void __launch_bounds__(WS_WARPS_PER_CTA * 32, WS_SCAN_MINBLOCKS) ws_prefilter_direct_kernel(WsPrefilterDirectArgs A) {
  __shared__ uint64_t s_fr[WS_WARPS_PER_CTA][128];
  __shared__ uint64_t s_sk[WS_WARPS_PER_CTA][64];
  __shared__ uint64_t s_sk2[WS_WARPS_PER_CTA][64];
  __shared__ int s_cpos[WS_WARPS_PER_CTA][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane & (WS_TEAM - 1);
  const int B = (int)A.s.k;
  WsScanOut so;
  so.vecs = A.s.vecs; so.dpad = A.s.dpad; so.res_keys = A.s.res_keys; so.res_cnt = A.s.res_cnt; so.stats = A.s.stats;
  so.out_ids = A.s.out_ids; so.out_dists = A.s.out_dists; so.decode = A.s.decode; so.pad_id = A.s.pad_id;
  const uint32_t nwarps = gridDim.x * WS_WARPS_PER_CTA;
  for (uint32_t qi = blockIdx.x * WS_WARPS_PER_CTA + warp; qi < A.nq; qi += nwarps) {
    const float lo = A.windows[2 * (size_t)qi], hi = A.windows[2 * (size_t)qi + 1];
    WsTask task;
    task.query = qi; task.node = -1; task.lo = lo; task.hi = hi; task.beam = 0; task.flags = WS_TF_SOLO;
    task.a = (uint32_t)ws_prefilter_bound(A.labels, A.n, lo);  // warp-uniform: every lane walks the same path
    task.b = (uint32_t)ws_prefilter_bound(A.labels, A.n, hi);
    if (task.b < task.a) task.b = task.a;  // inverted window: empty, the row is all pads
    float4 q[KQ];
    ws_load_query_global<KQ, EXACT>(A.s.queries, A.s.dim, A.s.dpad, qi, tl, q);
    ws_scan_task<KQ, METRIC, EXACT>(so, task, qi, q, B, s_fr[warp], s_sk[warp], s_sk2[warp], s_cpos[warp]);
  }
}

// ------------------------------------------------------------------------------------------
// K4b: merge of per-shard partial top-k lists after the all-gather of the label-sharded mode
//      (SURVEY.md §8e-2): parts x [nq][k] (id, dist) rows -> [nq][k], ascending (dist, id);
//      pad rows (dist == FLT_MAX) are ignored and re-created at the tail.
// ------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(WS_CTA_THREADS) ws_merge_parts_kernel(WsMergePartsArgs A) {
  __shared__ uint64_t buf[WS_TOPK_BUF];
  __shared__ int s_cnt, s_nbest;
  __shared__ uint64_t s_tau;
  WsTopk tk;
  tk.buf = buf; tk.cnt = &s_cnt; tk.nbest = &s_nbest; tk.tau = &s_tau;
  const int tid = threadIdx.x;
  const int K = (int)A.k;
  for (uint32_t qi = blockIdx.x; qi < A.nq; qi += gridDim.x) {
    const uint32_t q = A.q0 + qi;
    __syncthreads();
    ws_topk_init(tk, tid);
    __syncthreads();
    int nbest = 0, appended = 0;
    for (uint32_t p = 0; p < A.parts; p++) {
      if (nbest + appended + K > WS_TOPK_BUF) {
        if (tid == 0) s_cnt = appended;
        ws_topk_compact(tk, K, tid);
        nbest = s_nbest;
        appended = 0;
      }
      const uint32_t* pi = A.ids ? A.ids + (size_t)p * A.nq_total * K : A.part_ids[p];
      const float* pd = A.ids ? A.dists + (size_t)p * A.nq_total * K : A.part_dists[p];
      const size_t base = (size_t)q * K;
      for (int i = tid; i < K; i += WS_CTA_THREADS) {
        const float d = pd[base + i];
        buf[nbest + appended + i] = (d == 3.402823466e+38f) ? WS_KEY_MAX : ws_key(d, pi[base + i]);
      }
      appended += K;
    }
    __syncthreads();
    if (tid == 0) s_cnt = appended;
    ws_topk_compact(tk, K, tid);
    nbest = s_nbest;
    for (int j = tid; j < K; j += WS_CTA_THREADS) {
      const uint64_t key = j < nbest ? buf[j] : WS_KEY_MAX;
      if (key != WS_KEY_MAX) {
        A.out_ids[(size_t)q * K + j] = (uint32_t)(key & 0xFFFFFFFFull);
        A.out_dists[(size_t)q * K + j] = ws_unord((uint32_t)(key >> 32));
      } else {
        A.out_ids[(size_t)q * K + j] = A.pad_id;
        A.out_dists[(size_t)q * K + j] = 3.402823466e+38f;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K2c: CTA-per-task beam search for the large-beam tail (beams up to 1024 in shared memory,
//      up to 12288 with a global visited bitmap).
//
// Tail tasks are few and long (the doubling loop of postfilter_vamana.h:161-172 has driven
// them to beams in the hundreds or thousands), so what matters is the latency of ONE
// expansion, not occupancy: warp 0 is the control warp (pick, adjacency, visited filter,
// survivor sort, ranks) and all 8 warps gather distances together, so every candidate row of
// an expansion is in flight at once; the frontier is one array updated in place by all
// threads.  Same algorithm and same results as the warp kernel / ws_beam_kernel.
// ------------------------------------------------------------------------------------------

template <int KQ, int METRIC, bool EXACT, bool GLOBAL_SEEN, int CS>
__global__ void __launch_bounds__(WS_CTA2_THREADS, 1) ws_beam_cta2_kernel(WsBeamArgs A) {
  extern __shared__ __align__(16) unsigned char ws_smem[];
  const uint32_t CAP = A.beam_cap;
  uint64_t* fr = reinterpret_cast<uint64_t*>(ws_smem);   // [CAP]
  uint64_t* sk = fr + CAP;                                // [64]
  uint64_t* sk2 = sk + 64;                                // [64]
  int* cpos = reinterpret_cast<int*>(sk2 + 64);           // [64]
  int* cid = cpos + 64;                                   // [64]
  volatile int* hash = cid + 64;                          // [hash_mask + 1] (!GLOBAL_SEEN)
  __shared__ uint32_t s_task;
  __shared__ int s_m, s_cnt, s_mc2, s_first_new, s_pick, s_have;
  __shared__ float s_cutoff;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tl = lane & (WS_TEAM - 1), team = lane / WS_TEAM;
  const int dpad4 = A.dpad >> 2;
  const int K = (int)A.k;
  const int R = (int)A.R;
  const unsigned lt = (1u << lane) - 1u;
  const bool leader = tl == 0;
  uint32_t* bitmap = GLOBAL_SEEN ? A.bitmap + (size_t)blockIdx.x * A.bitmap_words : nullptr;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(A.q_head, 1u);
    __syncthreads();
    const uint32_t t = s_task;
    if (t >= *A.q_in_count) break;
    const uint32_t slot = A.q_in[t];
    const WsTask task = A.tasks[slot];
    const WsNode node = A.nodes[task.node];
    float4 q[KQ];
    ws_load_query_global<KQ, EXACT>(A.queries, A.dim, A.dpad, task.query, tl, q);
    const float4* vbase_tl = reinterpret_cast<const float4*>(A.vecs + (size_t)node.start * A.dpad) + tl;
    const int skip_id = A.skip_query_id ? (int)(task.query + A.query_id_base) : -1;

    long long beam = task.beam;
    int phase = (task.flags & WS_TF_FINAL) ? 1 : 0;
    const long long mult = (task.flags & WS_TF_MULT1) ? 1 : A.final_mult;
    int have = 0;
    bool escalate = false;
    if (!(task.flags & WS_TF_RESUMED) && tid == 0) A.res_cnt[slot] = 0;

    for (;;) {  // PostfilterVamanaIndex::query (postfilter_vamana.h:141-188)
      if (phase == 0) {
        if (!(have < K && beam < A.max_beam)) {
          long long fin = beam * mult;
          if (fin > A.max_beam) fin = A.max_beam;
          if (fin > beam) { beam = fin; phase = 1; } else break;
        }
      }
      if (beam > (long long)CAP) { escalate = true; break; }
      const int B = (int)beam;

      // ---- beam_search (beamSearch.h:51-184)
      if (!GLOBAL_SEEN) {
        for (int i = tid; i <= (int)A.hash_mask; i += WS_CTA2_THREADS) hash[i] = -1;
      } else {
        const int words = (int)((node.count + 31u) >> 5);
        for (int i = tid; i < words; i += WS_CTA2_THREADS) bitmap[i] = 0u;
      }
      if (warp == 0) {
        const float d0 = ws_team_dist_nv<KQ, METRIC, EXACT>(vbase_tl, q, tl, dpad4);
        if (lane == 0) fr[0] = ws_key(d0, 0u);
      }
      __syncthreads();
      if (tid == 0) {
        if (!GLOBAL_SEEN) ws_seen_warp(hash, A.hash_mask, 0); else ws_seen_bitmap(bitmap, 0);
      }
      int n = 1, scan_from = 0;
      unsigned long long nvis = 0, ncmp = 1;

      for (;;) {
        if ((long long)nvis >= A.limit) break;
        __syncthreads();
        // ---- A: control warp — pick, adjacency, visited filter (beamSearch.h:111-131)
        if (warp == 0) {
          int pick = -1;
          for (int b0 = scan_from; b0 < n; b0 += 32) {
            const int i = b0 + lane;
            const bool unv = i < n && !(fr[i] & 1ull);
            const unsigned bal = __ballot_sync(0xffffffffu, unv);
            if (bal) { pick = b0 + __ffs(bal) - 1; break; }
          }
          int m = 0;
          if (pick >= 0) {
            const uint64_t pkey = fr[pick];
            const uint32_t cur_id = (uint32_t)(pkey & 0xFFFFFFFFull) >> 1;
            __syncwarp();
            if (lane == 0) fr[pick] = pkey | 1ull;
            int nb0 = -1, nb1 = -1;
            const int* arow = node.adj + (size_t)cur_id * R;
            if (lane < R && (long long)lane < A.degree_limit) nb0 = __ldg(arow + lane);
            if (lane + 32 < R && (long long)(lane + 32) < A.degree_limit) nb1 = __ldg(arow + lane + 32);
            if (WS_BEAM_PREFETCH_NEXT) {  // adjacency row of the entry most likely expanded next -> L2 (see the warp kernel)
              int nxt = -1;
              for (int b0 = pick + 1; b0 < n && b0 < pick + 65; b0 += 32) {
                const int i = b0 + lane;
                const bool unv = i < n && !(fr[i] & 1ull);
                const unsigned bal = __ballot_sync(0xffffffffu, unv);
                if (bal) { nxt = b0 + __ffs(bal) - 1; break; }
              }
              if (nxt >= 0 && lane < 2) {
                const uint32_t nid = (uint32_t)(fr[nxt] & 0xFFFFFFFFull) >> 1;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(node.adj + (size_t)nid * R) + 128 * lane));
              }
            }
            bool keep0 = nb0 >= 0 && nb0 != skip_id;
            bool keep1 = nb1 >= 0 && nb1 != skip_id;
            if (GLOBAL_SEEN && WS_BEAM_PREFETCH_NEXT) {
              // The visited bitmap lives in global memory, so the probe below is a second dependent memory round trip
              // before the gathers can start.  The large tier is latency-bound (a handful of very long tasks), not
              // bandwidth-bound: pull every listed neighbour's row towards L2 while the bitmap answers.
              const int lines = ((int)A.dpad * 4 + 127) >> 7;
              const char* vb = reinterpret_cast<const char*>(vbase_tl - tl);
              if (keep0) for (int l = 0; l < lines; l++) asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + (size_t)nb0 * A.dpad * 4 + 128 * l));
              if (keep1) for (int l = 0; l < lines; l++) asm volatile("prefetch.global.L2 [%0];" ::"l"(vb + (size_t)nb1 * A.dpad * 4 + 128 * l));
            }
            if (!GLOBAL_SEEN) {
              ws_seen_warp2(hash, A.hash_mask, nb0, keep0, nb1, keep1);
            } else {
              if (keep0) keep0 = !ws_seen_bitmap(bitmap, nb0);
              if (keep1) keep1 = !ws_seen_bitmap(bitmap, nb1);
            }
            const unsigned bal0 = __ballot_sync(0xffffffffu, keep0), bal1 = __ballot_sync(0xffffffffu, keep1);
            const int m0 = __popc(bal0);
            m = m0 + __popc(bal1);
            if (keep0) cid[__popc(bal0 & lt)] = nb0;
            if (keep1) cid[m0 + __popc(bal1 & lt)] = nb1;
          }
          if (lane == 0) {
            s_pick = pick;
            s_m = m;
            s_cnt = 0;
            s_cutoff = (n < B) ? (float)2147483647 : ws_unord((uint32_t)(fr[n - 1] >> 32));
          }
        }
        __syncthreads();
        const int pick = s_pick;
        if (pick < 0) break;
        nvis++;
        const int m = s_m;
        if (m == 0) { scan_from = pick + 1; continue; }
        ncmp += (unsigned long long)m;

        // ---- B: every warp gathers its share of the candidate rows (beamSearch.h:135-145)
        {
          const float cutoff = s_cutoff;
          for (int jb = 0; jb < m; jb += (WS_CTA2_THREADS / WS_TEAM)) {
            const int j = jb + warp * 4 + team;
            const int id = cid[min(j, m - 1)];
            const float d = ws_team_dist_nv<KQ, METRIC, EXACT>(vbase_tl + (size_t)id * dpad4, q, tl, dpad4);
            const bool pu = leader && j < m && d < cutoff;
            const unsigned bu = __ballot_sync(0xffffffffu, pu);
            int wbase = 0;
            if (lane == 0 && bu) wbase = atomicAdd(&s_cnt, __popc(bu));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (pu) sk[wbase + __popc(bu & lt)] = ws_key(d, (uint32_t)id << 1);
          }
        }
        __syncthreads();
        const int s = s_cnt;
        if (s == 0) { scan_from = pick + 1; continue; }

        // ---- C: control warp — sort survivors, drop duplicates, rank against the frontier
        if (warp == 0) {
          uint64_t k0 = lane < s ? sk[lane] : WS_KEY_MAX, k1 = WS_KEY_MAX;
          if (s > 32) {
            k1 = lane + 32 < s ? sk[lane + 32] : WS_KEY_MAX;
            ws_warp_sort64(k0, k1, lane);
          } else {
            ws_warp_sort32(k0, lane);
          }
          const uint64_t up0 = ws_shfl_up_u64(k0, 1);
          uint64_t up1 = ws_shfl_up_u64(k1, 1);
          const uint64_t last0 = ws_shfl_idx_u64(k0, 31);
          up1 = lane == 0 ? last0 : up1;
          const int p0 = ws_lb_fixed<CS>(fr, n, k0 >> 1), p1 = ws_lb_fixed<CS>(fr, n, k1 >> 1);
          const uint64_t f0 = fr[min(p0, n - 1)], f1 = fr[min(p1, n - 1)];
          const bool ok0 = k0 != WS_KEY_MAX && !(lane > 0 && up0 == k0) && !(p0 < n && (f0 >> 1) == (k0 >> 1));
          const bool ok1 = k1 != WS_KEY_MAX && up1 != k1 && !(p1 < n && (f1 >> 1) == (k1 >> 1));
          const unsigned bka = __ballot_sync(0xffffffffu, ok0), bkb = __ballot_sync(0xffffffffu, ok1);
          const int ca = __popc(bka);
          if (ok0) { const int r = __popc(bka & lt); sk2[r] = k0; cpos[r] = p0; }
          if (ok1) { const int r = ca + __popc(bkb & lt); sk2[r] = k1; cpos[r] = p1; }
          __syncwarp();
          if (lane == 0) { s_mc2 = ca + __popc(bkb); s_first_new = (ca + __popc(bkb)) ? cpos[0] : n; }
        }
        __syncthreads();
        const int mc2 = s_mc2;
        if (mc2 == 0) { scan_from = pick + 1; continue; }
        const int first_new = s_first_new;

        // ---- D: all threads shift the tail of the frontier right, from the end, one chunk of
        //         256 entries at a time (a chunk's targets never fall below its own start)
        for (int hi = n; hi > first_new; hi -= WS_CTA2_THREADS) {
          const int i = hi - 1 - tid;
          const bool valid = i >= first_new;
          uint64_t key = 0;
          int c = 0;
          if (valid) {
            key = fr[i];
            if (mc2 <= 4) {  // the common case at large beams: a few broadcast compares instead of a binary search
              const uint64_t kk = key >> 1;
              c = (int)((sk2[0] >> 1) < kk);
              if (mc2 > 1) c += (int)((sk2[1] >> 1) < kk);
              if (mc2 > 2) c += (int)((sk2[2] >> 1) < kk);
              if (mc2 > 3) c += (int)((sk2[3] >> 1) < kk);
            } else {
              c = ws_lb_fixed<6>(sk2, mc2, key >> 1);
            }
          }
          __syncthreads();
          if (valid && i + c < B) fr[i + c] = key;
        }
        __syncthreads();
        if (tid < mc2) {
          const int pos = cpos[tid] + tid;
          if (pos < B) fr[pos] = sk2[tid];
        }
        n = min(n + mc2, B);
        scan_from = min(pick + 1, first_new);
      }

      // ---- raw_query's label predicate by the control warp (postfilter_vamana.h:234-251)
      __syncthreads();
      if (warp == 0) {
        int hv = 0;
        for (int b0 = 0; b0 < n && hv < K; b0 += 32) {
          const int i = b0 + lane;
          bool in = false;
          uint64_t key = 0;
          uint32_t rank = 0;
          if (i < n) {
            key = fr[i];
            rank = node.start + ((uint32_t)(key & 0xFFFFFFFFull) >> 1);
            const float lab = __ldg(A.labels + rank);
            in = (lab >= task.lo) && (lab <= task.hi);
          }
          const unsigned bal = __ballot_sync(0xffffffffu, in);
          const int r = hv + __popc(bal & lt);
          if (in && r < K) {
            const uint64_t okey = (key & 0xFFFFFFFF00000000ull) | rank;
            A.res_keys[(size_t)slot * K + r] = okey;
            if (task.flags & WS_TF_SOLO) ws_write_result(A.out_ids, A.out_dists, A.decode, task.query, K, r, okey);
          }
          hv += __popc(bal);
        }
        if (lane == 0) {
          s_have = hv;
          A.res_cnt[slot] = (uint32_t)min(hv, K);
          atomicAdd(A.stats + WS_ST_SEARCHES, 1ull);
          atomicAdd(A.stats + WS_ST_VISITED, nvis);
          atomicAdd(A.stats + WS_ST_DISTCMPS, ncmp);
          atomicAdd(A.stats + WS_ST_BEAMSUM, (unsigned long long)B);
        }
      }
      __syncthreads();
      have = s_have;
      if (phase == 1) break;
      if (have < K) beam *= 2;
    }

    if (!escalate && (task.flags & WS_TF_SOLO))
      for (int j = min(have, K) + tid; j < K; j += WS_CTA2_THREADS) ws_write_pad(A.out_ids, A.out_dists, A.pad_id, task.query, K, j);
    if (escalate && tid == 0) {
      if (A.q_out != nullptr) {
        A.tasks[slot].beam = (uint32_t)beam;
        A.tasks[slot].flags = task.flags | WS_TF_RESUMED | (phase ? WS_TF_FINAL : 0u);
        const uint32_t pos = atomicAdd(A.q_out_count, 1u);
        A.q_out[pos] = slot;
        atomicAdd(A.stats + WS_ST_ESCALATED, 1ull);
      } else if (A.sticky != nullptr) {
        atomicOr(A.sticky, 2u);
      }
    }
  }
}
