// ws_decompose.h — window -> task decomposition for the tree query methods.
//
// One function per reference query method; each turns a (lo,hi) label window into
//   * graph tasks  (query, node)        -> a PostfilterVamanaIndex::query on that node
//   * scan tasks   (query, [a,b))       -> brute force over a contiguous slice of the
//                                          label-sorted arena
// The same code is compiled for the device (K3 decomposition kernel) and for the host
// (ws_debug_decompose_host, a CPU-side testing hook that performs no search).
//
// Reference behaviour restated here (file:line into /root/reference):
//   first_greater_than_or_equal_to            src/tree_utils.h:19-37
//   check_empty                               src/range_filter_tree.h:191-203
//   find_range_containing_index               src/range_filter_tree.h:213-232
//   find_largest_ranges_within_query_range    src/range_filter_tree.h:234-295
//   fenwick_tree_search (cover construction)  src/range_filter_tree.h:297-361,386-397
//   optimized_postfiltering_search            src/range_filter_tree.h:403-471
//   three_split_search                        src/range_filter_tree.h:473-540
//   super_optimized_postfiltering_search      src/super_optimized_postfilter_tree.h:187-245
//   PrefilterIndex::query_knn window search   src/prefiltering.h:159-184
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define WS_HD __host__ __device__ __forceinline__
#else
#define WS_HD inline
#endif

// task flags
#define WS_TF_MULT1 1u     // final_beam_multiply forced to 1 (three_split centre buckets, range_filter_tree.h:490-498)
#define WS_TF_FINAL 2u     // (resumed task) the pending search is the final-multiply search
#define WS_TF_RESUMED 4u   // task was escalated from a smaller beam tier
#define WS_TF_SOLO 8u      // the query's only task: the search kernel writes the final ids/dists itself

struct WsTask {            // 32 bytes
  uint32_t query;          // index into the batch
  int32_t node;            // >=0: graph search on this node; -1: brute-force scan of [a,b)
  uint32_t a, b;           // scan slice in arena ranks (scan tasks only)
  float lo, hi;            // label window the task filters with (closed for graph tasks)
  uint32_t beam;           // beam to start (or resume) at
  uint32_t flags;
};

struct WsGeom {
  const float* labels;     // [n] sorted
  uint64_t n;
  uint64_t pf_n;           // the `n` PrefilterIndex::query_knn's bounds use: n (the reference's r = n-1 rule, the last
                           // point is never inside a window) or n+1 (plain lower bound — every label shard but the
                           // last one, whose last point is NOT the data set's last point)
  // ---- B-WST (range_filter_tree.h:129-189)
  uint32_t wst_rows;
  uint32_t split;
  int32_t cutoff;
  const uint32_t* wst_nb;        // [rows] buckets per row
  const uint32_t* wst_off_ptr;   // [rows] start of row r's offsets in wst_off (nb+1 entries)
  const uint64_t* wst_off;       // concatenated offsets
  const uint32_t* wst_node_ptr;  // [rows] start of row r's node handles in wst_nodes
  const int32_t* wst_nodes;
  uint32_t wst_prefilter_nodes;  // != 0: the buckets' sub-indices are PrefilterIndex-es, not graphs
  // ---- super-postfilter tree (super_optimized_postfilter_tree.h:118-171)
  uint32_t sup_rows;
  const uint64_t* sup_size;      // [rows]
  const uint64_t* sup_shift;     // [rows]
  const uint32_t* sup_nb;        // [rows]
  const uint32_t* sup_node_ptr;  // [rows]
  const int32_t* sup_nodes;
};

struct WsDecompParams {
  uint32_t beam;
  int32_t has_ratio;
  float min_ratio;
  uint32_t scan_chunk;     // rows per brute-force task (large slices are split)
};

// Emits tasks for one query into a fixed-capacity slot array.
struct WsEmitter {
  WsTask* slots;
  uint32_t cap;
  uint32_t count;
  uint32_t overflow;
  uint32_t query;
  uint32_t beam;
  uint32_t scan_chunk;

  WS_HD void graph(int32_t node, float lo, float hi, uint32_t flags) {
    if (count >= cap) { overflow = 1; return; }
    WsTask t;
    t.query = query; t.node = node; t.a = 0; t.b = 0; t.lo = lo; t.hi = hi;
    t.beam = beam; t.flags = flags;
    slots[count++] = t;
  }
  WS_HD void scan(uint64_t a, uint64_t b, float lo, float hi) {
    while (a < b) {
      uint64_t e = (b - a > scan_chunk) ? a + scan_chunk : b;
      if (count >= cap) { overflow = 1; return; }
      WsTask t;
      t.query = query; t.node = -1; t.a = (uint32_t)a; t.b = (uint32_t)e; t.lo = lo; t.hi = hi;
      t.beam = 0; t.flags = 0;
      slots[count++] = t;
      a = e;
    }
  }
};

// tree_utils.h:19-37
WS_HD uint64_t ws_lower_bound(const float* labels, uint64_t n, float v) {
  if (labels[0] >= v) return 0;
  uint64_t start = 0, end = n;
  while (start + 1 < end) {
    uint64_t mid = (start + end) / 2;
    if (labels[mid] >= v) end = mid; else start = mid;
  }
  return end;
}

// prefiltering.h:159-184 — lower bound with r initialised to n-1:
// the last sorted point can never be inside a window.
WS_HD uint64_t ws_prefilter_bound(const float* labels, uint64_t n, float v) {
  uint64_t l = 0, r = n - 1;
  while (l < r) {
    uint64_t mid = (l + r) / 2;
    if (labels[mid] < v) l = mid + 1; else r = mid;
  }
  return l;
}

// range_filter_tree.h:191-203 (the reference also prints a warning)
WS_HD bool ws_check_empty(const WsGeom& g, float lo, float hi) {
  return hi < g.labels[0] || lo > g.labels[g.n - 1];
}

WS_HD uint64_t ws_wst_off(const WsGeom& g, uint32_t row, uint64_t i) {
  return g.wst_off[g.wst_off_ptr[row] + i];
}
WS_HD int32_t ws_wst_node(const WsGeom& g, uint32_t row, uint64_t i) {
  return g.wst_nodes[g.wst_node_ptr[row] + i];
}

// One sub-index query of a B-WST bucket (range_filter_tree.h:346,353,372-381,468,490-498: `index->query`).
// With Vamana sub-indices that is a graph task.  With PrefilterIndex sub-indices
// (RangeFilterTreeIndex<T, Point>, python_bindings.cpp:119-127; range_filter_tree.h:32 default template
// argument) it is query_knn (prefiltering.h:154-204) over the bucket's own sorted labels: a scan of
// [lb(lo), lb(hi)) with `r` starting at count-1, so a bucket's last point is never returned.
WS_HD void ws_emit_wst_node(const WsGeom& g, uint32_t row, uint64_t b, float lo, float hi, uint32_t flags,
                            WsEmitter& em) {
  if (!g.wst_prefilter_nodes) {
    em.graph(ws_wst_node(g, row, b), lo, hi, flags);
    return;
  }
  uint64_t ns = ws_wst_off(g, row, b), cnt = ws_wst_off(g, row, b + 1) - ns;
  uint64_t a = ns + ws_prefilter_bound(g.labels + ns, cnt, lo);
  uint64_t e = ns + ws_prefilter_bound(g.labels + ns, cnt, hi);
  em.scan(a, e, lo, hi);
}

// range_filter_tree.h:213-232: bucket of `row` containing sorted rank `index` (< n).
WS_HD uint64_t ws_find_range_containing(const WsGeom& g, uint32_t row, uint64_t index) {
  uint64_t left = 0, right = g.wst_nb[row];  // invariant: off[left] <= index < off[right]
  while (left + 1 < right) {
    uint64_t mid = (left + right) / 2;
    if (ws_wst_off(g, row, mid) <= index) left = mid; else right = mid;
  }
  return left;
}

struct WsSeqBuckets {
  bool ok;
  uint32_t row;
  uint64_t first, last;      // bucket indices [first, last)
  uint64_t cover_s, cover_e; // ranks covered
};

// range_filter_tree.h:234-295.  Divergence (documented): where the
// reference indexes one past the row's offsets and throws for the whole batch (window
// starting inside the last bucket of the chosen row), we treat the row as having no
// whole bucket and continue exactly as its `end > exclusive_end` branch does.
WS_HD WsSeqBuckets ws_find_largest_ranges(const WsGeom& g, uint64_t s, uint64_t e) {
  WsSeqBuckets out;
  out.ok = false; out.row = 0; out.first = out.last = 0; out.cover_s = out.cover_e = 0;
  uint64_t range_size = e - s;
  uint32_t row = 0;
  bool found = false;
  for (uint32_t r = 0; r < g.wst_rows; r++) {
    uint64_t bucket_size = ws_wst_off(g, r, 1) - ws_wst_off(g, r, 0) - 1;
    if (bucket_size <= range_size) { row = r; found = true; break; }
  }
  if (!found) return out;
  uint64_t fri = (s == 0) ? 0 : ws_find_range_containing(g, row, s - 1) + 1;
  bool descend = false;
  uint64_t start = 0, end = 0;
  if (fri >= g.wst_nb[row]) {
    descend = true;
  } else {
    start = ws_wst_off(g, row, fri);
    end = ws_wst_off(g, row, fri + 1);
    if (end > e) descend = true;
  }
  if (descend) {
    row += 1;
    if (row >= g.wst_rows) return out;
    fri = (s == 0) ? 0 : ws_find_range_containing(g, row, s - 1) + 1;
    if (fri >= g.wst_nb[row]) return out;
    start = ws_wst_off(g, row, fri);
    end = ws_wst_off(g, row, fri + 1);
    // never true on the reference's own path (one row down the first whole bucket always
    // fits); guards the divergence path above so a cover never overruns the window
    if (end > e) return out;
  }
  uint64_t lri = fri + 1;
  while (lri < g.wst_nb[row]) {
    uint64_t next_end = ws_wst_off(g, row, lri + 1);
    if (next_end > e) break;
    lri++;
    end = next_end;
  }
  out.ok = true; out.row = row; out.first = fri; out.last = lri;
  out.cover_s = start; out.cover_e = end;
  return out;
}

// range_filter_tree.h:297-401 (everything but the searches themselves and the final sort)
WS_HD void ws_decompose_fenwick(const WsGeom& g, float lo, float hi, uint32_t flags,
                                WsEmitter& em) {
  if (ws_check_empty(g, lo, hi)) return;
  uint64_t s = ws_lower_bound(g.labels, g.n, lo);
  uint64_t e = ws_lower_bound(g.labels, g.n, hi);
  if (e <= s) return;  // empty (or inverted) window: nothing to search
  WsSeqBuckets c = ws_find_largest_ranges(g, s, e);
  if (!c.ok) {
    em.scan(s, e, lo, hi);
    return;
  }
  for (uint64_t b = c.first; b < c.last; b++) ws_emit_wst_node(g, c.row, b, lo, hi, flags, em);
  uint64_t cover_s = c.cover_s, cover_e = c.cover_e;
  uint64_t left = c.first, right = c.last - 1;
  for (uint32_t row = c.row + 1; row < g.wst_rows; row++) {
    left *= g.split;
    right = right * g.split + (g.split - 1);
    while (left > 0) {
      uint64_t nls = ws_wst_off(g, row, left - 1);
      if (nls < s) break;
      cover_s = nls;
      left -= 1;
      ws_emit_wst_node(g, row, left, lo, hi, flags, em);
    }
    while (right + 1 < g.wst_nb[row]) {
      uint64_t nre = ws_wst_off(g, row, right + 2);
      if (nre > e) break;
      cover_e = nre;
      right += 1;
      ws_emit_wst_node(g, row, right, lo, hi, flags, em);
    }
  }
  em.scan(s, cover_s, lo, hi);
  em.scan(cover_e, e, lo, hi);
}

// range_filter_tree.h:403-471
WS_HD void ws_decompose_opt_postfilter(const WsGeom& g, float lo, float hi,
                                       const WsDecompParams& p, WsEmitter& em) {
  if (ws_check_empty(g, lo, hi)) return;
  uint64_t s = ws_lower_bound(g.labels, g.n, lo);
  uint64_t e = ws_lower_bound(g.labels, g.n, hi);
  if (e < s) return;  // inverted window: the reference's search can only come back empty
  if (4 * (e - s) < (uint64_t)(int64_t)g.cutoff) {
    ws_decompose_fenwick(g, lo, hi, 0, em);
    return;
  }
  uint32_t row = 0;
  uint64_t idx = 0;
  while (row + 1 < g.wst_rows) {
    uint32_t next = row + 1;
    bool have = false;
    uint64_t working = 0;
    for (uint64_t cand = idx * g.split; cand < idx * g.split + g.split; cand++) {
      if (cand >= g.wst_nb[next]) break;
      uint64_t ns = ws_wst_off(g, next, cand), ne = ws_wst_off(g, next, cand + 1);
      if (s >= ns && e <= ne) { working = cand; have = true; }
    }
    if (!have) break;
    idx = working;
    row = next;
  }
  uint64_t bucket_size = ws_wst_off(g, row, idx + 1) - ws_wst_off(g, row, idx);
  float ratio = (float)bucket_size / (float)(e - s);
  if (p.has_ratio && ratio > p.min_ratio) {
    ws_decompose_fenwick(g, lo, hi, 0, em);
    return;
  }
  ws_emit_wst_node(g, row, idx, lo, hi, 0, em);
}

// range_filter_tree.h:473-540
WS_HD void ws_decompose_three_split(const WsGeom& g, float lo, float hi,
                                    const WsDecompParams& p, WsEmitter& em) {
  if (ws_check_empty(g, lo, hi)) return;
  uint64_t s = ws_lower_bound(g.labels, g.n, lo);
  uint64_t e = ws_lower_bound(g.labels, g.n, hi);
  if (e <= s) return;
  WsSeqBuckets c = ws_find_largest_ranges(g, s, e);
  if (!c.ok) {
    ws_decompose_fenwick(g, lo, hi, WS_TF_MULT1, em);
    return;
  }
  for (uint64_t b = c.first; b < c.last; b++)
    ws_emit_wst_node(g, c.row, b, lo, hi, WS_TF_MULT1, em);
  if (c.cover_s > s) ws_decompose_opt_postfilter(g, lo, g.labels[c.cover_s], p, em);
  if (e > c.cover_e) ws_decompose_opt_postfilter(g, g.labels[c.cover_e], hi, p, em);
}

// super_optimized_postfilter_tree.h:187-258
WS_HD void ws_decompose_super(const WsGeom& g, float lo, float hi, WsEmitter& em) {
  if (ws_check_empty(g, lo, hi)) return;
  uint64_t s = ws_lower_bound(g.labels, g.n, lo);
  uint64_t e = ws_lower_bound(g.labels, g.n, hi);
  if (e < s) return;  // inverted window: the reference's search can only come back empty
  int64_t row;
  uint64_t idx = 0;
  for (row = (int64_t)g.sup_rows - 1; row >= 0; row--) {
    if (row == 0) { idx = 0; break; }
    uint64_t size = g.sup_size[row];
    if (size < e - s) continue;
    uint64_t shift = g.sup_shift[row];
    uint64_t nb = g.sup_nb[row];
    uint64_t first = s / shift, last = (e - 1) / shift;  // unsigned wrap for e == 0, as the reference
    if (first > nb - 1) first = nb - 1;
    if (last > nb - 1) last = nb - 1;
    bool hit = false;
    for (uint64_t t = first; t <= last; t++) {
      uint64_t bs = t * shift;
      uint64_t be = bs + size < g.n ? bs + size : g.n;
      if (s >= bs && e <= be) { idx = t; hit = true; break; }
    }
    if (hit) break;
  }
  em.graph(g.sup_nodes[g.sup_node_ptr[row] + idx], lo, hi, 0);
}

