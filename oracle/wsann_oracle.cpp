// wsann_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A plain, scalar C++ restatement of the reference's window-search query path, used only
// as the checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The
// product (rangefilteredann_b200/) never links, imports or calls anything in oracle/.
//
// Parity of this restatement is PINNED: tests/test_oracle_golden.py checks it against
// tests/golden/tiny_ref_outputs.npz, which was produced by the unmodified reference
// compiled from /root/reference (oracle/Makefile `ref` target, tests/golden/make_golden.py).
//
// Each function cites the reference code it follows (paths relative to /root/reference):
//   distance (L2 / MIPS)         ParlayANN/algorithms/utils/NSGDist.h:31-70, mips_point.h:60-66
//   beam_search                  ParlayANN/algorithms/utils/beamSearch.h:51-184
//   hash64_2                     ParlayANN/parlaylib/include/parlay/utilities.h:145-150
//   postfilter query/raw_query   src/postfilter_vamana.h:141-188,223-254
//   prefilter query_knn          src/prefiltering.h:154-204
//   lower bound                  src/tree_utils.h:19-37
//   B-WST build geometry         src/range_filter_tree.h:129-189
//   find_largest_ranges…         src/range_filter_tree.h:213-295
//   fenwick / optimized / three  src/range_filter_tree.h:297-540
//   super tree geometry + query  src/super_optimized_postfilter_tree.h:118-171,187-270
//   graph file                   ParlayANN/algorithms/utils/graph.h:126-196
//
// dist_mode 0 = reference summation order (8 strided lanes, then lanes added in index
//               order; L2 multiply and add rounded separately, MIPS as compiled —
//               both pinned empirically against oracle/_ref, tests/test_oracle_golden.py)
// dist_mode 1 = the device kernels' summation order (ws_device.cuh: team of 8 lanes,
//               lane t owns float4 columns t, t+8, …; butterfly 4,2,1) — lets GPU results
//               be compared bit for bit.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <limits>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Graph {
  int32_t n = 0, max_degree = 0;
  std::vector<int32_t> deg, off, edges;
};

Graph load_graph(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("oracle: cannot open " + path);
  Graph g;
  in.read((char*)&g.n, 4);
  in.read((char*)&g.max_degree, 4);
  g.deg.resize(g.n);
  in.read((char*)g.deg.data(), 4ll * g.n);
  g.off.resize(g.n + 1);
  g.off[0] = 0;
  for (int i = 0; i < g.n; i++) g.off[i + 1] = g.off[i] + g.deg[i];
  g.edges.resize(g.off[g.n]);
  in.read((char*)g.edges.data(), 4ll * g.off[g.n]);
  if (!in) throw std::runtime_error("oracle: truncated " + path);
  return g;
}

std::string graph_name(const std::string& cache, long L, long R, double alpha, float mn, float mx, size_t n) {
  return cache + "vamana_" + std::to_string(L) + "_" + std::to_string(R) + "_" + std::to_string(alpha) + "_" +
         std::to_string(mn) + "_" + std::to_string(mx) + "_" + std::to_string(n) + ".bin";
}

struct Node {
  size_t start = 0, count = 0;
  Graph g;
};

struct Pid {
  int32_t id;
  float dist;
};
inline bool pid_less(const Pid& a, const Pid& b) { return a.dist < b.dist || (a.dist == b.dist && a.id < b.id); }

struct Params {
  long k, beam;
  long final_mult, max_beam;
  int has_ratio;
  float ratio;
  long limit, degree_limit;
};

struct Oracle {
  int metric = 0, dist_mode = 0;
  size_t n = 0, dim = 0;
  bool sorted = false;
  std::vector<float> vecs;      // arena order
  std::vector<float> labels;    // arena order
  std::vector<uint32_t> decode; // arena rank -> original id (identity when !sorted)
  std::vector<Node> nodes;
  // B-WST
  std::vector<std::vector<size_t>> offs;
  std::vector<std::vector<int>> wst_nodes;
  size_t split = 2;
  int32_t cutoff = 1000;
  // RangeFilterTreeIndex<T, Point, PrefilterIndex> (range_filter_tree.h:32, python_bindings.cpp:119-127):
  // bucket sub-indices are PrefilterIndex-es over the bucket's slice, not graphs
  bool prefilter_nodes = false;
  // super tree
  std::vector<size_t> sup_size, sup_shift;
  std::vector<std::vector<int>> sup_nodes;
  // when set, node_query / brute record (kind, start, end, mult1) instead of searching
  std::vector<int64_t>* trace = nullptr;

  const float* row(size_t i) const { return &vecs[i * dim]; }

  float distance(const float* a, const float* b) const {
    const size_t d = dim;
    if (dist_mode == 0) {
      if (metric == 0) {  // NSGDist.h:39-69 — 8 lane accumulators over dims padded to 8
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        // the remainder block (D % 16 == 8) is accumulated first, as the reference does
        size_t D = (d + 7) & ~size_t(7), DR = D % 16, DD = D - DR;
        auto lane_block = [&](size_t base) {
          for (size_t j = 0; j < 8; j++) {
            size_t i = base + j;
            float diff = (i < d) ? a[i] - b[i] : 0.f;  // d is a multiple of 8 in every config used
            acc[j] = acc[j] + diff * diff;  // not fused: measured against the compiled reference
          }
        };
        if (DR) lane_block(DD);
        for (size_t i = 0; i < DD; i += 16) { lane_block(i); lane_block(i + 8); }
        return acc[0] + acc[1] + acc[2] + acc[3] + acc[4] + acc[5] + acc[6] + acc[7];
      }
      float r = 0;  // mips_point.h:60-66 — strict sequential sum
      for (size_t i = 0; i < d; i++) r = r + b[i] * a[i];  // not fused (measured)
      return -r;
    }
    // device order (ws_device.cuh ws_team_dist)
    float lane[8];
    size_t dpad4 = (((d * 4 + 63) / 64) * 64 / 4) / 4;
    for (size_t t = 0; t < 8; t++) {
      float acc = 0.f;
      for (size_t c = t; c < dpad4; c += 8)
        for (size_t e = 0; e < 4; e++) {
          size_t i = c * 4 + e;
          float av = i < d ? a[i] : 0.f, bv = i < d ? b[i] : 0.f;
          if (metric == 0) { float df = av - bv; acc = std::fmaf(df, df, acc); }
          else acc = std::fmaf(av, bv, acc);
        }
      lane[t] = acc;
    }
    float s4[4], s2[2];
    for (int t = 0; t < 4; t++) s4[t] = lane[t] + lane[t + 4];
    for (int t = 0; t < 2; t++) s2[t] = s4[t] + s4[t + 2];
    float r = s2[0] + s2[1];
    return metric == 0 ? r : -r;
  }

  static uint64_t hash64_2(uint64_t x) {  // utilities.h:145-150
    x = (x ^ (x >> 30)) * UINT64_C(0xbf58476d1ce4e5b9);
    x = (x ^ (x >> 27)) * UINT64_C(0x94d049bb133111eb);
    return x ^ (x >> 31);
  }

  // beamSearch.h:51-184 with QP.k = QP.beamSize = beam; returns the final frontier
  std::vector<Pid> beam_search(const float* q, long qid, const Node& nd, long beam, const Params& P,
                               uint64_t* nvis, uint64_t* ncmp) const {
    int bits = std::max<int>(10, (int)std::ceil(std::log2((double)(beam * beam))) - 2);
    std::vector<int32_t> hash_filter((size_t)1 << bits, -1);
    auto seen = [&](int32_t a) {
      size_t loc = hash64_2((uint64_t)(int64_t)a) & (((size_t)1 << bits) - 1);
      if (hash_filter[loc] == a) return true;
      hash_filter[loc] = a;
      return false;
    };
    const size_t base = nd.start;
    std::vector<Pid> frontier, visited, unvisited((size_t)beam), new_frontier, candidates;
    frontier.reserve(beam);
    frontier.push_back({0, distance(row(base), q)});
    unvisited[0] = frontier[0];
    new_frontier.resize(beam + nd.g.max_degree);
    uint64_t dc = 1;
    long remain = 1, num_visited = 0;
    std::vector<int32_t> keep;
    while (remain > 0 && num_visited < P.limit) {
      Pid cur = unvisited[0];
      visited.insert(std::upper_bound(visited.begin(), visited.end(), cur, pid_less), cur);
      num_visited++;
      candidates.clear();
      keep.clear();
      long ne = std::min<long>(nd.g.deg[cur.id], P.degree_limit);
      for (long i = 0; i < ne; i++) {
        int32_t a = nd.g.edges[nd.g.off[cur.id] + i];
        if (a == qid || seen(a)) continue;
        keep.push_back(a);
      }
      float cutoff = ((long)frontier.size() < beam) ? (float)std::numeric_limits<int>::max() : frontier.back().dist;
      for (int32_t a : keep) {
        float dist = distance(row(base + a), q);
        dc++;
        if (dist >= cutoff) continue;
        candidates.push_back({a, dist});
      }
      std::sort(candidates.begin(), candidates.end(), pid_less);
      size_t nf = std::set_union(frontier.begin(), frontier.end(), candidates.begin(), candidates.end(),
                                 new_frontier.begin(), pid_less) - new_frontier.begin();
      nf = std::min<size_t>(beam, nf);
      // `cut` pruning (beamSearch.h:162-167) is dead here: it needs nf > QP.k = beam
      frontier.assign(new_frontier.begin(), new_frontier.begin() + nf);
      remain = std::set_difference(frontier.begin(), frontier.end(), visited.begin(), visited.end(),
                                   unvisited.begin(), pid_less) - unvisited.begin();
    }
    *nvis += num_visited;
    *ncmp += dc;
    return frontier;
  }

  // postfilter_vamana.h:223-254 — closed-interval predicate, ids mapped to arena ranks
  std::vector<Pid> raw_query(const float* q, long qid, const Node& nd, float lo, float hi, long beam,
                             const Params& P, uint64_t* nvis, uint64_t* ncmp, uint64_t* ns) const {
    std::vector<Pid> fr = beam_search(q, qid, nd, beam, P, nvis, ncmp);
    (*ns)++;
    std::vector<Pid> out;
    for (const Pid& p : fr) {
      float v = labels[nd.start + p.id];
      if (v >= lo && v <= hi) out.push_back({(int32_t)(nd.start + p.id), p.dist});
    }
    return out;
  }

  // postfilter_vamana.h:141-188
  std::vector<Pid> node_query(const float* q, long qid, const Node& nd, float lo, float hi, const Params& P,
                              long final_mult, uint64_t* nvis, uint64_t* ncmp, uint64_t* ns) const {
    if (prefilter_nodes) {  // PrefilterIndex::query -> query_knn over the bucket's own labels (prefiltering.h:148-204)
      const float* nl = &labels[nd.start];
      auto bound = [&](float v) {
        size_t l = 0, r = nd.count - 1;
        while (l < r) { size_t mid = (l + r) / 2; if (nl[mid] < v) l = mid + 1; else r = mid; }
        return l;
      };
      size_t a = nd.start + bound(lo), e = nd.start + bound(hi);
      std::vector<Pid> fr;
      if (e > a) brute(q, a, e, fr);
      sort_truncate(fr, P.k);
      return fr;
    }
    if (trace) {
      int64_t rec[4] = {0, (int64_t)nd.start, (int64_t)(nd.start + nd.count), final_mult != P.final_mult ? 1 : 0};
      trace->insert(trace->end(), rec, rec + 4);
      return {};
    }
    long beam = P.beam;
    std::vector<Pid> fr;
    while ((long)fr.size() < P.k && beam < P.max_beam) {
      fr = raw_query(q, qid, nd, lo, hi, beam, P, nvis, ncmp, ns);
      if ((long)fr.size() < P.k) beam *= 2;
    }
    long fin = std::min<long>(beam * final_mult, P.max_beam);
    if (fin > beam) fr = raw_query(q, qid, nd, lo, hi, fin, P, nvis, ncmp, ns);
    return fr;
  }

  size_t lower_bound(float v) const {  // tree_utils.h:19-37
    if (labels[0] >= v) return 0;
    size_t s = 0, e = n;
    while (s + 1 < e) {
      size_t mid = (s + e) / 2;
      if (labels[mid] >= v) e = mid; else s = mid;
    }
    return e;
  }
  bool check_empty(float lo, float hi) const { return hi < labels.front() || lo > labels.back(); }

  void brute(const float* q, size_t a, size_t b, std::vector<Pid>& out) const {
    if (trace) {
      int64_t rec[4] = {-1, (int64_t)a, (int64_t)b, 0};
      if (a < b) trace->insert(trace->end(), rec, rec + 4);
      return;
    }
    for (size_t i = a; i < b; i++) out.push_back({(int32_t)i, distance(row(i), q)});
  }
  static void sort_truncate(std::vector<Pid>& v, size_t k) {  // range_filter_tree.h:542-549 (ties by id here)
    std::sort(v.begin(), v.end(), pid_less);
    if (v.size() > k) v.resize(k);
  }

  size_t find_containing(size_t r, size_t index) const {  // range_filter_tree.h:213-232
    const auto& o = offs[r];
    size_t b = std::upper_bound(o.begin(), o.end(), index) - o.begin();
    return b - 1;
  }

  struct Seq { bool ok; size_t row, first, last, cs, ce; };
  // range_filter_tree.h:234-295; where the reference would index past the row and throw
  //  this follows the engine's documented divergence: descend / no cover
  Seq find_largest(size_t s, size_t e) const {
    Seq out{false, 0, 0, 0, 0, 0};
    size_t range = e - s, r = 0;
    bool found = false;
    for (size_t i = 0; i < offs.size(); i++) {
      size_t bs = offs[i][1] - offs[i][0] - 1;
      if (bs <= range) { r = i; found = true; break; }
    }
    if (!found) return out;
    auto first_after = [&](size_t rr) { return s == 0 ? 0 : find_containing(rr, s - 1) + 1; };
    size_t fri = first_after(r), nb = offs[r].size() - 1, start = 0, end = 0;
    bool descend = fri >= nb;
    if (!descend) { start = offs[r][fri]; end = offs[r][fri + 1]; descend = end > e; }
    if (descend) {
      r++;
      if (r >= offs.size()) return out;
      fri = first_after(r);
      nb = offs[r].size() - 1;
      if (fri >= nb) return out;
      start = offs[r][fri]; end = offs[r][fri + 1];
      if (end > e) return out;
    }
    size_t lri = fri + 1;
    while (lri < nb) {
      size_t ne = offs[r][lri + 1];
      if (ne > e) break;
      lri++; end = ne;
    }
    return Seq{true, r, fri, lri, start, end};
  }

  // range_filter_tree.h:297-401
  std::vector<Pid> fenwick(const float* q, long qid, float lo, float hi, const Params& P, long mult,
                           uint64_t* c3) const {
    std::vector<Pid> fr;
    if (check_empty(lo, hi)) return fr;
    size_t s = lower_bound(lo), e = lower_bound(hi);
    if (e <= s) return fr;
    Seq c = find_largest(s, e);
    std::vector<std::pair<size_t, size_t>> todo;
    bool have_cover = c.ok;
    size_t cs = 0, ce = 0;
    if (c.ok) {
      for (size_t b = c.first; b < c.last; b++) todo.push_back({c.row, b});
      cs = c.cs; ce = c.ce;
      size_t left = c.first, right = c.last - 1;
      for (size_t r = c.row + 1; r < offs.size(); r++) {
        left *= split; right = right * split + split - 1;
        while (left > 0) {
          size_t nls = offs[r][left - 1];
          if (nls < s) break;
          cs = nls; left--; todo.push_back({r, left});
        }
        while (right < offs[r].size() - 2) {
          size_t nre = offs[r][right + 2];
          if (nre > e) break;
          ce = nre; right++; todo.push_back({r, right});
        }
      }
    }
    for (auto& rb : todo) {
      auto part = node_query(q, qid, nodes[wst_nodes[rb.first][rb.second]], lo, hi, P, mult, c3, c3 + 1, c3 + 2);
      fr.insert(fr.end(), part.begin(), part.end());
    }
    if (have_cover) { brute(q, s, cs, fr); brute(q, ce, e, fr); }
    else brute(q, s, e, fr);
    sort_truncate(fr, P.k);
    return fr;
  }

  // range_filter_tree.h:403-471
  std::vector<Pid> optimized(const float* q, long qid, float lo, float hi, const Params& P, uint64_t* c3) const {
    if (check_empty(lo, hi)) return {};
    size_t s = lower_bound(lo), e = lower_bound(hi);
    if (e < s) return {};
    if (4 * (e - s) < (size_t)cutoff) return fenwick(q, qid, lo, hi, P, P.final_mult, c3);
    size_t r = 0, idx = 0;
    while (r + 1 < offs.size()) {
      size_t nr = r + 1;
      long found = -1;
      for (size_t cnd = idx * split; cnd < idx * split + split; cnd++) {
        if (cnd >= offs[nr].size() - 1) break;
        if (s >= offs[nr][cnd] && e <= offs[nr][cnd + 1]) found = (long)cnd;
      }
      if (found < 0) break;
      idx = (size_t)found; r = nr;
    }
    size_t bsize = offs[r][idx + 1] - offs[r][idx];
    float ratio = (float)bsize / (float)(e - s);
    if (P.has_ratio && ratio > P.ratio) return fenwick(q, qid, lo, hi, P, P.final_mult, c3);
    return node_query(q, qid, nodes[wst_nodes[r][idx]], lo, hi, P, P.final_mult, c3, c3 + 1, c3 + 2);
  }

  // range_filter_tree.h:473-540
  std::vector<Pid> three_split(const float* q, long qid, float lo, float hi, const Params& P, uint64_t* c3) const {
    if (check_empty(lo, hi)) return {};
    size_t s = lower_bound(lo), e = lower_bound(hi);
    if (e <= s) return {};
    Seq c = find_largest(s, e);
    if (!c.ok) return fenwick(q, qid, lo, hi, P, 1, c3);
    std::vector<Pid> fr;
    for (size_t b = c.first; b < c.last; b++) {
      auto part = node_query(q, qid, nodes[wst_nodes[c.row][b]], lo, hi, P, 1, c3, c3 + 1, c3 + 2);
      fr.insert(fr.end(), part.begin(), part.end());
    }
    if (c.cs > s) { auto part = optimized(q, qid, lo, labels[c.cs], P, c3); fr.insert(fr.end(), part.begin(), part.end()); }
    if (e > c.ce) { auto part = optimized(q, qid, labels[c.ce], hi, P, c3); fr.insert(fr.end(), part.begin(), part.end()); }
    sort_truncate(fr, P.k);
    return fr;
  }

  // super_optimized_postfilter_tree.h:187-270
  std::vector<Pid> super_query(const float* q, long qid, float lo, float hi, const Params& P, uint64_t* c3) const {
    if (check_empty(lo, hi)) return {};
    size_t s = lower_bound(lo), e = lower_bound(hi);
    if (e < s) return {};
    long r;
    size_t idx = 0;
    for (r = (long)sup_size.size() - 1; r >= 0; r--) {
      if (r == 0) { idx = 0; break; }
      size_t size = sup_size[r];
      if (size < e - s) continue;
      size_t shift = sup_shift[r], nb = sup_nodes[r].size();
      size_t first = std::min(s / shift, nb - 1), last = std::min((e - 1) / shift, nb - 1);
      bool hit = false;
      for (size_t t = first; t <= last; t++) {
        size_t bs = t * shift, be = std::min(bs + size, n);
        if (s >= bs && e <= be) { idx = t; hit = true; break; }
      }
      if (hit) break;
    }
    return node_query(q, qid, nodes[sup_nodes[r][idx]], lo, hi, P, P.final_mult, c3, c3 + 1, c3 + 2);
  }

  // prefiltering.h:154-204 (binary searches with r = n-1)
  std::vector<Pid> prefilter(const float* q, float lo, float hi, long k) const {
    auto bound = [&](float v) {
      size_t l = 0, r = n - 1;
      while (l < r) { size_t mid = (l + r) / 2; if (labels[mid] < v) l = mid + 1; else r = mid; }
      return l;
    };
    size_t s = bound(lo), e = bound(hi);
    std::vector<Pid> fr;
    if (e > s) brute(q, s, e, fr);
    sort_truncate(fr, k);
    return fr;
  }
};

int add_node(Oracle* o, const std::string& cache, long L, long R, double alpha, size_t start, size_t count) {
  float mn, mx;
  if (o->sorted) { mn = o->labels[start]; mx = o->labels[start + count - 1]; }
  else { mn = *std::min_element(o->labels.begin(), o->labels.end()); mx = *std::max_element(o->labels.begin(), o->labels.end()); }
  Node nd;
  nd.start = start; nd.count = count;
  if (cache == "<none>") {  // geometry-only oracle (decomposition traces): no graphs loaded
    o->nodes.push_back(std::move(nd));
    return (int)o->nodes.size() - 1;
  }
  nd.g = load_graph(graph_name(cache, L, R, alpha, mn, mx, count));
  if ((size_t)nd.g.n != count) throw std::runtime_error("oracle: graph size mismatch");
  o->nodes.push_back(std::move(nd));
  return (int)o->nodes.size() - 1;
}

thread_local std::string g_err;

}  // namespace

extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }

// kind: 0 prefilter (sorted arena, no graphs), 1 flat postfilter (unsorted, one graph), 4 B-WST over
// PrefilterIndex sub-indices (no graphs),
//       2 B-WST with Vamana nodes, 3 super-postfilter tree
void* oracle_create(int kind, int metric, int dist_mode, uint64_t n, uint32_t dim, const float* points,
                    const float* labels, int32_t cutoff, float split_factor, float shift_factor, long L, long R,
                    double alpha, const char* cache_path) {
  try {
    Oracle* o = new Oracle();
    o->metric = metric; o->dist_mode = dist_mode; o->n = n; o->dim = dim;
    o->sorted = kind != 1;
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    if (o->sorted)  // tree_utils.h:62-66 (ties: original id; the reference's order is unspecified)
      std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return labels[a] < labels[b]; });
    o->vecs.resize((size_t)n * dim);
    o->labels.resize(n);
    o->decode = order;
    for (size_t i = 0; i < n; i++) {
      std::memcpy(&o->vecs[i * dim], points + (size_t)order[i] * dim, dim * sizeof(float));
      o->labels[i] = labels[order[i]];
    }
    std::string cache = cache_path ? cache_path : "<none>";
    if (kind == 1) {
      add_node(o, cache, L, R, alpha, 0, n);
    } else if (kind == 2 || kind == 4) {  // range_filter_tree.h:129-189
      if (kind == 4) { o->prefilter_nodes = true; cache = "<none>"; }
      o->split = (size_t)split_factor; o->cutoff = cutoff;
      o->offs.push_back({0, (size_t)n});
      while ((long)o->offs.back()[1] > (long)cutoff) {
        const auto& last = o->offs.back();
        size_t lnb = last.size() - 1;
        std::vector<size_t> next(lnb * o->split + 1);
        next.back() = n;
        for (size_t b = 0; b < lnb; b++) {
          size_t ls = last[b], size = last[b + 1] - ls;
          size_t large = (size + o->split - 1) / o->split, small = large - 1, nl = size - small * o->split;
          for (size_t i = 0; i < o->split; i++)
            next[b * o->split + i] = i < nl ? ls + i * large : ls + nl * large + (i - nl) * small;
        }
        o->offs.push_back(next);
      }
      for (auto& row : o->offs) {
        std::vector<int> ids;
        for (size_t b = 0; b + 1 < row.size(); b++) ids.push_back(add_node(o, cache, L, R, alpha, row[b], row[b + 1] - row[b]));
        o->wst_nodes.push_back(ids);
      }
    } else if (kind == 3) {  // super_optimized_postfilter_tree.h:134-170
      o->cutoff = cutoff;
      o->sup_size.push_back(n); o->sup_shift.push_back(0);
      o->sup_nodes.push_back({add_node(o, cache, L, R, alpha, 0, n)});
      while ((long)o->sup_size.back() > (long)cutoff) {
        size_t last = o->sup_size.back();
        size_t bsize = (size_t)((last + split_factor - 1) / split_factor);
        size_t bshift = (size_t)std::ceil(bsize * shift_factor);
        o->sup_size.push_back(bsize); o->sup_shift.push_back(bshift);
        size_t nb = ((n - bsize) + bshift - 1) / bshift + 1;
        std::vector<int> ids;
        for (size_t b = 0; b < nb; b++) {
          size_t s = b * bshift, e = std::min<size_t>(s + bsize, n);
          ids.push_back(add_node(o, cache, L, R, alpha, s, e - s));
        }
        o->sup_nodes.push_back(ids);
      }
    }
    return o;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    return nullptr;
  }
}

void oracle_destroy(void* h) { delete (Oracle*)h; }

// method: 0 fenwick, 1 optimized_postfilter, 2 three_split, 3 super, 10 prefilter, 11 flat postfilter
// pad_id: value written into empty slots; stats3 (may be null): visited, dist_cmps, searches
int oracle_batch(void* h, int method, const float* queries, const float* windows, uint64_t nq, long k, long beam,
                 long final_mult, long max_beam, int has_ratio, float ratio, uint32_t pad_id, int threads,
                 uint32_t* ids, float* dists, uint64_t* stats3) {
  Oracle* o = (Oracle*)h;
  Params P{k, beam, final_mult, max_beam, has_ratio, ratio, 10000000L, 10000L};
  if (threads < 1) threads = 1;
  std::vector<std::vector<uint64_t>> counters(threads, std::vector<uint64_t>(3, 0));
  std::string err;
  auto work = [&](int t) {
    try {
      for (uint64_t i = t; i < nq; i += threads) {
        const float* q = queries + i * o->dim;
        float lo = windows[2 * i], hi = windows[2 * i + 1];
        uint64_t* c3 = counters[t].data();
        std::vector<Pid> r;
        switch (method) {
          case 0: r = o->fenwick(q, (long)i, lo, hi, P, P.final_mult, c3); break;
          case 1: r = o->optimized(q, (long)i, lo, hi, P, c3); break;
          case 2: r = o->three_split(q, (long)i, lo, hi, P, c3); break;
          case 3: r = o->super_query(q, (long)i, lo, hi, P, c3); break;
          case 10: r = o->prefilter(q, lo, hi, k); break;
          case 11: r = o->node_query(q, (long)i, o->nodes[0], lo, hi, P, P.final_mult, c3, c3 + 1, c3 + 2); break;
        }
        for (long j = 0; j < k; j++) {
          if (j < (long)r.size()) { ids[i * k + j] = o->decode[r[j].id]; dists[i * k + j] = r[j].dist; }
          else { ids[i * k + j] = pad_id; dists[i * k + j] = std::numeric_limits<float>::max(); }
        }
      }
    } catch (const std::exception& ex) { err = ex.what(); }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
  if (!err.empty()) { g_err = err; return -1; }
  if (stats3) for (int j = 0; j < 3; j++) { stats3[j] = 0; for (auto& c : counters) stats3[j] += c[j]; }
  return 0;
}

// Task decomposition of one window, for the CPU tests of ws_debug_decompose_host: records
// (kind, start, end, mult1) per sub-index query (kind 0) / brute-force slice (kind -1) in the
// order the reference would issue them.  Returns the number of records (<= cap) or -1.
int oracle_decompose(void* h, int method, float lo, float hi, long beam, long final_mult, int has_ratio,
                     float ratio, uint32_t cap, int64_t* out) {
  Oracle* o = (Oracle*)h;
  Params P{10, beam, final_mult, 10000, has_ratio, ratio, 10000000L, 10000L};
  std::vector<int64_t> tr;
  std::vector<float> q(o->dim, 0.f);
  uint64_t c3[3] = {0, 0, 0};
  o->trace = &tr;
  try {
    switch (method) {
      case 0: o->fenwick(q.data(), -1, lo, hi, P, P.final_mult, c3); break;
      case 1: o->optimized(q.data(), -1, lo, hi, P, c3); break;
      case 2: o->three_split(q.data(), -1, lo, hi, P, c3); break;
      case 3: o->super_query(q.data(), -1, lo, hi, P, c3); break;
      case 10: o->prefilter(q.data(), lo, hi, 10); break;
      default: o->trace = nullptr; return -1;
    }
  } catch (const std::exception& ex) { o->trace = nullptr; g_err = ex.what(); return -1; }
  o->trace = nullptr;
  size_t nrec = tr.size() / 4;
  if (nrec > cap) return -1;
  std::memcpy(out, tr.data(), tr.size() * sizeof(int64_t));
  return (int)nrec;
}

}  // extern "C"
