// ws_build.cuh — Vamana graph construction on the device, for indices whose graph cache is
// missing.  NOT on the query hot path: it exists so that BASELINE-sized
// trees (1M points x 11 rows = 2047 graphs) can be produced in seconds on the GPU box and
// saved in the reference's own .bin format, which the reference then loads — both
// implementations still search the identical graph.
//
// Algorithm = ParlayANN's batch insertion, restated for lock-step execution over many
// graphs at once (ParlayANN/algorithms/vamana/index.h):
//   batch schedule (prefix doubling, then 2% batches)         index.h:211-268
//   per point: beam search (L) from local id 0, robustPrune    index.h:254-262, 61-108
//   reverse edges grouped by target; append or re-prune         index.h:266-298
//   final sort of every adjacency list by distance              index.h:131-134
// The random insertion order comes from this file's own generator, so graphs differ from
// the reference builder's (as two runs of any randomized builder do); the format, the
// parameters (R, L, alpha) and the pruning rule are the same.
#pragma once
#include "ws_kernels.cuh"


__device__ __forceinline__ uint32_t ws_build_find_graph_by_task(const WsBuildGraph* g, uint32_t n, uint32_t t) {
  uint32_t lo = 0, hi = n;  // last graph with task_off <= t
  while (lo + 1 < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (g[mid].task_off <= t) lo = mid; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ uint32_t ws_build_find_graph_by_row(const WsBuildGraph* g, uint32_t n, uint32_t row) {
  uint32_t lo = 0, hi = n;
  while (lo + 1 < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (g[mid].row_off <= row) lo = mid; else hi = mid;
  }
  return lo;
}

template <int KQ>
__device__ __forceinline__ void ws_load_row(const float4* row, float4 (&q)[KQ], int tl, int dpad4) {
#pragma unroll
  for (int i = 0; i < KQ; i++) {
    int c = tl + WS_TEAM * i;
    q[i] = (c < dpad4) ? ws_ldg_f4(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// robustPrune (vamana/index.h:61-108).  cand[0..nc) holds keys ord(dist to p)<<32 | id<<1
// (low bit ignored).  Writes up to R neighbour ids to out[], returns the count.  All
// threads of the CTA call it; cand must have room for pow2ceil(nc) keys.
template <int KQ, int METRIC>
__device__ __forceinline__ int ws_robust_prune(uint64_t* cand, int nc, int p_local, const float4* vbase, int dpad4,
                                               double alpha, int R, int* out, int* s_cnt) {
  const int tid = threadIdx.x, lane = tid & 31;
  const int tl = lane & (WS_TEAM - 1), team = tid / WS_TEAM;
  const int NTEAMS = WS_CTA_THREADS / WS_TEAM;
  const int np2 = max(ws_pow2ceil(nc), 2);
  for (int i = tid; i < np2; i += WS_CTA_THREADS) {
    uint64_t k = i < nc ? (cand[i] & ~1ull) : WS_KEY_MAX;
    if (i < nc && (int)((uint32_t)(k & 0xFFFFFFFFull) >> 1) == p_local) k = WS_KEY_MAX;  // p itself (index.h:93)
    cand[i] = k;
  }
  __syncthreads();
  ws_cta_sort(cand, np2, tid);
  int cnt = 0;
  int idx = 0;
  while (cnt < R && idx < nc) {
    const uint64_t ks = cand[idx];
    idx++;
    if (ks == WS_KEY_MAX) continue;  // removed, or the padding tail (sorted last)
    const int p_star = (int)((uint32_t)(ks & 0xFFFFFFFFull) >> 1);
    if (tid == 0) out[cnt] = p_star;
    cnt++;
    float4 qs[KQ];
    ws_load_row<KQ>(vbase + (size_t)p_star * dpad4, qs, tl, dpad4);
    for (int ib = idx; ib < nc; ib += NTEAMS) {
      const int i = ib + team;
      uint64_t k = i < nc ? cand[i] : WS_KEY_MAX;
      const bool alive = k != WS_KEY_MAX;
      const int id = alive ? (int)((uint32_t)(k & 0xFFFFFFFFull) >> 1) : 0;
      float d = ws_team_dist<KQ, METRIC>(vbase + (size_t)id * dpad4, qs, tl, dpad4, alive);
      if (alive && tl == 0) {
        float dp = ws_unord((uint32_t)(k >> 32));
        if (alpha * (double)d <= (double)dp) cand[i] = WS_KEY_MAX;  // index.h:99-101
      }
    }
    __syncthreads();
  }
  if (tid == 0) *s_cnt = cnt;
  __syncthreads();
  return cnt;
}

// ---- round kernel 1: search + prune for every point of the round ---------------------------
template <int KQ, int METRIC>
__global__ void __launch_bounds__(WS_CTA_THREADS) ws_build_insert_kernel(WsBuildArgs A) {
  extern __shared__ __align__(16) unsigned char ws_smem[];
  WsBeamSmem S;
  float* qs_unused = ws_carve_beam_smem(ws_smem, A.beam_cap, A.cand_cap, A.dpad, S);
  (void)qs_unused;
  uint64_t* vis = reinterpret_cast<uint64_t*>(S.hash + (A.hash_mask + 1));  // [WS_BUILD_VCAP]
  __shared__ uint32_t s_task;
  __shared__ int s_m, s_npick, s_cnt;
  __shared__ int s_pick[8];
  __shared__ int s_wc[WS_CTA_THREADS / 32];
  __shared__ int s_out[128];
  S.s_m = &s_m; S.s_npick = &s_npick; S.s_pick = s_pick; S.s_wc = s_wc;

  const int tid = threadIdx.x, lane = tid & 31;
  const int tl = lane & (WS_TEAM - 1), team = tid / WS_TEAM;
  const int NTEAMS = WS_CTA_THREADS / WS_TEAM;
  const int dpad4 = A.dpad >> 2;
  const int R = (int)A.R;
  WsSearchCfg C;
  C.R = R; C.E = (int)A.expand; C.dpad4 = dpad4; C.hash_mask = A.hash_mask;
  C.limit = 1ll << 62; C.degree_limit = 1ll << 62; C.bitmap = nullptr;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(A.head, 1u);
    __syncthreads();
    const uint32_t t = s_task;
    if (t >= A.ntasks) break;
    const WsBuildGraph g = A.graphs[ws_build_find_graph_by_task(A.graphs, A.ngraphs, t)];
    const int p = A.perm[g.row_off + g.floor + (t - g.task_off)];
    WsNode node;
    node.adj = A.adj + (size_t)g.row_off * R;
    node.start = g.start;
    node.count = g.count;
    const float4* vbase = reinterpret_cast<const float4*>(A.vecs + (size_t)g.start * A.dpad);
    float4 q[KQ];
    ws_load_row<KQ>(vbase + (size_t)p * dpad4, q, tl, dpad4);

    uint64_t* cur;
    unsigned long long nvis, ncmp;
    ws_beam_search<KQ, METRIC, false>(S, C, node, vbase, q, (int)A.L, p, &cur, &nvis, &ncmp, vis,
                                      WS_BUILD_VCAP - R);
    int nc = (int)min(nvis, (unsigned long long)(WS_BUILD_VCAP - R));
    // add the point's current out-neighbours (robustPrune's add = true, index.h:69-74)
    const int dcur = A.deg[g.row_off + p];
    for (int jb = 0; jb < dcur; jb += NTEAMS) {
      const int j = jb + team;
      const bool valid = j < dcur;
      const int id = valid ? node.adj[(size_t)p * R + j] : 0;
      float d = ws_team_dist<KQ, METRIC>(vbase + (size_t)id * dpad4, q, tl, dpad4, valid);
      if (valid && tl == 0) vis[nc + j] = ws_key(d, (uint32_t)id << 1);
    }
    nc += dcur;
    __syncthreads();
    const int cnt = ws_robust_prune<KQ, METRIC>(vis, nc, p, vbase, dpad4, A.alpha, R, s_out, &s_cnt);
    for (int j = tid; j < R; j += WS_CTA_THREADS) A.new_out[(size_t)t * R + j] = j < cnt ? s_out[j] : -1;
    if (tid == 0) {
      A.new_cnt[t] = cnt;
      atomicAdd(A.stats + 0, 1ull);
      atomicAdd(A.stats + 1, nvis);
      atomicAdd(A.stats + 2, ncmp);
    }
  }
}

// ---- round kernel 2: install the new out-lists, emit reverse edges (index.h:266-281) --------
__global__ void __launch_bounds__(256) ws_build_apply_kernel(WsBuildArgs A) {
  const uint32_t t = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= A.ntasks) return;
  const WsBuildGraph g = A.graphs[ws_build_find_graph_by_task(A.graphs, A.ngraphs, t)];
  const int p = A.perm[g.row_off + g.floor + (t - g.task_off)];
  const int R = (int)A.R;
  const int cnt = A.new_cnt[t];
  uint32_t base = 0;
  if (lane == 0) {
    A.deg[g.row_off + p] = cnt;
    base = atomicAdd(A.pair_count, (uint32_t)cnt);
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  for (int j = lane; j < R; j += 32) {
    const int v = A.new_out[(size_t)t * R + j];
    A.adj[((size_t)g.row_off + p) * R + j] = v;
    if (j < cnt) A.pairs[base + j] = ((uint64_t)(g.row_off + (uint32_t)v) << 32) | (uint32_t)p;
  }
}

// ---- round kernel 3: segment heads of the sorted reverse-edge list ---------------------------
__global__ void ws_build_heads_kernel(const uint64_t* pairs, uint32_t n, uint32_t* heads, uint32_t* head_count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0 || (pairs[i] >> 32) != (pairs[i - 1] >> 32)) heads[atomicAdd(head_count, 1u)] = i;
}

// ---- round kernel 4: add reverse edges; re-prune rows that overflow (index.h:285-297) --------

template <int KQ, int METRIC>
__global__ void __launch_bounds__(WS_CTA_THREADS) ws_build_reverse_kernel(WsBuildRevArgs A) {
  __shared__ uint64_t cand[WS_BUILD_VCAP];
  __shared__ uint32_t s_task;
  __shared__ int s_len, s_cnt;
  __shared__ int s_out[128];
  const int tid = threadIdx.x, lane = tid & 31;
  const int tl = lane & (WS_TEAM - 1), team = tid / WS_TEAM;
  const int NTEAMS = WS_CTA_THREADS / WS_TEAM;
  const int dpad4 = A.dpad >> 2;
  const int R = (int)A.R;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_task = atomicAdd(A.head, 1u);
    __syncthreads();
    const uint32_t w = s_task;
    if (w >= A.nheads) break;
    const uint32_t h0 = A.heads[w];
    const uint32_t row = (uint32_t)(A.pairs[h0] >> 32);
    const WsBuildGraph g = A.graphs[ws_build_find_graph_by_row(A.graphs, A.ngraphs, row)];
    const int j_local = (int)(row - g.row_off);
    // segment length, capped so that candidates + current neighbours fit
    const int cap = WS_BUILD_VCAP - R;
    if (tid == 0) s_len = (int)min((uint32_t)cap, A.npairs - h0);
    __syncthreads();
    const int lim = s_len;
    for (int i = tid; i < lim; i += WS_CTA_THREADS)
      if ((uint32_t)(A.pairs[h0 + i] >> 32) != row) atomicMin(&s_len, i);
    __syncthreads();
    const int c = s_len;
    const int dcur = A.deg[row];
    if (dcur + c <= R) {  // append_neighbors
      for (int i = tid; i < c; i += WS_CTA_THREADS)
        A.adj[(size_t)row * R + dcur + i] = (int32_t)(A.pairs[h0 + i] & 0xFFFFFFFFull);
      if (tid == 0) A.deg[row] = dcur + c;
      continue;
    }
    // robustPrune(index, candidates) with add = true
    const float4* vbase = reinterpret_cast<const float4*>(A.vecs + (size_t)g.start * A.dpad);
    float4 q[KQ];
    ws_load_row<KQ>(vbase + (size_t)j_local * dpad4, q, tl, dpad4);
    const int nc = c + dcur;
    for (int ib = 0; ib < nc; ib += NTEAMS) {
      const int i = ib + team;
      const bool valid = i < nc;
      int id = 0;
      if (valid) id = i < c ? (int32_t)(A.pairs[h0 + i] & 0xFFFFFFFFull) : A.adj[(size_t)row * R + (i - c)];
      float d = ws_team_dist<KQ, METRIC>(vbase + (size_t)id * dpad4, q, tl, dpad4, valid);
      if (valid && tl == 0) cand[i] = ws_key(d, (uint32_t)id << 1);
    }
    __syncthreads();
    const int cnt = ws_robust_prune<KQ, METRIC>(cand, nc, j_local, vbase, dpad4, A.alpha, R, s_out, &s_cnt);
    for (int j = tid; j < R; j += WS_CTA_THREADS) A.adj[(size_t)row * R + j] = j < cnt ? s_out[j] : -1;
    if (tid == 0) { A.deg[row] = cnt; atomicAdd(A.stats + 3, 1ull); }
  }
}

// ---- final: sort every adjacency list by distance to its node (index.h:131-134) ---------------

template <int KQ, int METRIC>
__global__ void __launch_bounds__(WS_CTA_THREADS) ws_build_sortadj_kernel(WsBuildSortArgs A) {
  __shared__ uint64_t keys[128];
  const int tid = threadIdx.x, lane = tid & 31;
  const int tl = lane & (WS_TEAM - 1), team = tid / WS_TEAM;
  const int NTEAMS = WS_CTA_THREADS / WS_TEAM;
  const int dpad4 = A.dpad >> 2;
  const int R = (int)A.R;
  for (uint32_t row = blockIdx.x; row < A.rows; row += gridDim.x) {
    const WsBuildGraph g = A.graphs[ws_build_find_graph_by_row(A.graphs, A.ngraphs, row)];
    const int p = (int)(row - g.row_off);
    const int d = A.deg[row];
    const float4* vbase = reinterpret_cast<const float4*>(A.vecs + (size_t)g.start * A.dpad);
    float4 q[KQ];
    ws_load_row<KQ>(vbase + (size_t)p * dpad4, q, tl, dpad4);
    __syncthreads();
    for (int jb = 0; jb < d; jb += NTEAMS) {
      const int j = jb + team;
      const bool valid = j < d;
      const int id = valid ? A.adj[(size_t)row * R + j] : 0;
      float dist = ws_team_dist<KQ, METRIC>(vbase + (size_t)id * dpad4, q, tl, dpad4, valid);
      if (valid && tl == 0) keys[j] = ws_key(dist, (uint32_t)id);
    }
    const int np2 = max(ws_pow2ceil(d), 2);
    for (int i = d + tid; i < np2; i += WS_CTA_THREADS) keys[i] = WS_KEY_MAX;
    __syncthreads();
    ws_cta_sort(keys, np2, tid);
    for (int j = tid; j < d; j += WS_CTA_THREADS) A.adj[(size_t)row * R + j] = (int32_t)(keys[j] & 0xFFFFFFFFull);
  }
}

__global__ void ws_fill_i32_kernel(int32_t* p, size_t n, int32_t v) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}
