// window_index.hpp — host-side index classes of the drop-in boundary's upper face.
//
// Same class names, constructor arguments and batch_search signatures as the reference's
// header-only classes; all query work is delegated to libwsann_cuda.so through the C ABI
// (include/wsann.h).  What stays on the host is what the reference also does once at
// construction: validate inputs, argsort the labels, lay the points out label-sorted,
// derive the tree geometry and find each node's cached Vamana graph.
//
//   PrefilterIndex                      src/prefiltering.h:28-205
//   PostfilterVamanaIndex               src/postfilter_vamana.h:29-255
//   RangeFilterTreeIndex (Vamana nodes) src/range_filter_tree.h:30-550
//   SuperOptimizedPostfilterTree        src/super_optimized_postfilter_tree.h:29-271
//   sort_python_and_convert             src/tree_utils.h:39-98
//   Graph file format                   ParlayANN/algorithms/utils/graph.h:126-196
//   QueryParams / BuildParams           ParlayANN/algorithms/utils/types.h:77-140
//
// Graphs are loaded from the reference builder's cache files (postfilter_vamana.h:54-61,
// 126-132) whenever they exist.  Missing ones are built on the device with the same
// algorithm and saved in the same format (Arena::realize_graphs), so the reference can load
// them back and both implementations search identical graphs.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <limits>
#include <memory>
#include <numeric>
#include <optional>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/wsann.h"

namespace wsann {

struct BuildParams {  // types.h:77-112 (only the fields this path reads)
  long L = 0;
  long R = 0;
  double alpha = 0;
  std::string cache_path;
  BuildParams() {}
  BuildParams(long R, long L, double a, std::string cache_path)
      : L(L), R(R), alpha(a), cache_path(std::move(cache_path)) {}
};

struct QueryParams {  // types.h:115-140
  long k = 0;
  long beamSize = 0;
  double cut = 1.35;
  long limit = 0;
  long degree_limit = 0;
  long final_beam_multiply = 8;
  long postfiltering_max_beam = 10000;
  std::optional<float> min_query_to_bucket_ratio = std::nullopt;
  bool verbose = false;
  QueryParams() {}
  QueryParams(long k, long Q, double cut, long limit, long dg, long final_beam_multiply,
              long postfiltering_max_beam, std::optional<float> min_query_to_bucket_ratio, bool verbose)
      : k(k), beamSize(Q), cut(cut), limit(limit), degree_limit(dg),
        final_beam_multiply(final_beam_multiply), postfiltering_max_beam(postfiltering_max_beam),
        min_query_to_bucket_ratio(min_query_to_bucket_ratio), verbose(verbose) {}

  ws_query_params to_c() const {
    ws_query_params c;
    std::memset(&c, 0, sizeof(c));
    c.k = k; c.beam_size = beamSize; c.cut = cut; c.limit = limit; c.degree_limit = degree_limit;
    c.final_beam_multiply = final_beam_multiply; c.postfiltering_max_beam = postfiltering_max_beam;
    c.has_min_query_to_bucket_ratio = min_query_to_bucket_ratio.has_value() ? 1 : 0;
    c.min_query_to_bucket_ratio = min_query_to_bucket_ratio.value_or(0.f);
    c.verbose = verbose ? 1 : 0;
    return c;
  }
};

inline void check(int status, const char* what) {
  if (status != WS_OK)
    throw std::runtime_error(std::string(what) + ": " + ws_last_error() + " (status " + std::to_string(status) + ")");
}

inline int default_device() {
  if (const char* e = std::getenv("WSANN_DEVICE")) return std::atoi(e);
  return 0;
}

// Which GPUs an index uses.  The reference spreads a batch over the host's cores (PARLAY_NUM_THREADS,
// run_our_method.py:131); here the environment names the devices:
//   WSANN_DEVICE=3            one GPU (default: 0) — what a one-process-per-GPU launcher sets per rank
//   WSANN_DEVICES=0,1,2,3     several GPUs behind ONE batch_search call ("all" = every visible device)
//   WSANN_SHARD_MODE=label    with several devices: contiguous label ranges instead of replicas (data sets beyond one
//                             GPU's HBM; PrefilterIndex and the range-filter trees — the other classes replicate)
struct DevicePlan {
  std::vector<int> devices;
  bool label_sharded = false;
};

inline DevicePlan device_plan() {
  DevicePlan p;
  if (const char* e = std::getenv("WSANN_DEVICES")) {
    std::string v(e);
    if (v == "all") {
      int c = 0;
      ws_device_count(&c);
      for (int i = 0; i < c; i++) p.devices.push_back(i);
    } else {
      size_t pos = 0;
      while (pos < v.size()) {
        size_t comma = v.find(',', pos);
        if (comma == std::string::npos) comma = v.size();
        if (comma > pos) p.devices.push_back(std::atoi(v.substr(pos, comma - pos).c_str()));
        pos = comma + 1;
      }
    }
  }
  if (p.devices.empty()) p.devices.push_back(default_device());
  if (const char* m = std::getenv("WSANN_SHARD_MODE")) p.label_sharded = std::string(m) == "label" && p.devices.size() > 1;
  return p;
}

// contiguous, balanced [lo, hi) slice `i` of `n` items cut into `parts` (sizes differ by at most one)
inline std::pair<size_t, size_t> shard_bounds(size_t n, size_t i, size_t parts) {
  const size_t base = n / parts, rem = n % parts;
  const size_t lo = i * base + std::min(i, rem);
  return {lo, lo + base + (i < rem ? 1 : 0)};
}

// ---- reference graph cache -----------------------------------------------------------------
struct GraphFile {
  int32_t n = 0, max_degree = 0;
  std::vector<int32_t> degrees, edges;
};

// graph.h:126-172 layout: int32 n, int32 maxDeg, int32 degree[n], int32 edges[sum degree]
inline GraphFile read_graph_file(const std::string& path) {
  std::ifstream in(path, std::ios::binary);
  if (!in) throw std::runtime_error("cannot open graph file " + path);
  GraphFile g;
  in.read(reinterpret_cast<char*>(&g.n), 4);
  in.read(reinterpret_cast<char*>(&g.max_degree), 4);
  if (!in || g.n <= 0 || g.max_degree <= 0) throw std::runtime_error("bad graph header in " + path);
  g.degrees.resize(g.n);
  in.read(reinterpret_cast<char*>(g.degrees.data()), 4ll * g.n);
  if (!in) throw std::runtime_error("truncated degree table in " + path);
  int64_t total = 0;
  for (int32_t d : g.degrees) {
    if (d < 0 || d > g.max_degree) throw std::runtime_error("bad degree in " + path);
    total += d;
  }
  g.edges.resize(total);
  in.read(reinterpret_cast<char*>(g.edges.data()), 4ll * total);
  if (!in) throw std::runtime_error("truncated edge list in " + path);
  return g;
}

// postfilter_vamana.h:126-132 (std::to_string of long / double / float)
inline std::string graph_filename(const BuildParams& bp, float min_label, float max_label, size_t n) {
  return bp.cache_path + "vamana_" + std::to_string(bp.L) + "_" + std::to_string(bp.R) + "_" +
         std::to_string(bp.alpha) + "_" + std::to_string(min_label) + "_" + std::to_string(max_label) + "_" +
         std::to_string(n) + ".bin";
}

// ---- the HBM arena + its graphs --------------------------------------------------------------
class Arena {
 public:
  Arena() = default;
  Arena(const Arena&) = delete;
  Arena& operator=(const Arena&) = delete;
  ~Arena() {
    if (group_) ws_group_destroy(group_);
    for (ws_index* r : replicas_) ws_index_destroy(r);
    if (idx_) ws_index_destroy(idx_);
  }

  // tree_utils.h:39-98: argsort labels, physically permute, decode[sorted] = original.
  // Equal labels are ordered by original id (the reference's unstable parlay sort leaves
  // that order implementation-defined).
  void init_sorted(const float* points, const float* labels, size_t n, size_t dim, int metric, int device) {
    n_ = n; dim_ = dim;
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return labels[a] < labels[b]; });
    std::vector<float> sorted(n * dim);
    labels_.resize(n);
    for (size_t i = 0; i < n; i++) {
      std::memcpy(&sorted[i * dim], points + (size_t)order[i] * dim, dim * sizeof(float));
      labels_[i] = labels[order[i]];
    }
    check(ws_index_create(device, metric, n, (uint32_t)dim, sorted.data(), labels_.data(), order.data(), 1, &idx_),
          "ws_index_create");
  }

  // one label shard: points already label-sorted, decode = the shard's ORIGINAL ids
  void init_presorted(const float* sorted_points, const float* sorted_labels, const uint32_t* decode, size_t n, size_t dim,
                      int metric, int device) {
    n_ = n; dim_ = dim;
    labels_.assign(sorted_labels, sorted_labels + n);
    check(ws_index_create(device, metric, n, (uint32_t)dim, sorted_points, sorted_labels, decode, 1, &idx_), "ws_index_create");
  }

  // Upload the staged geometry; with several devices, clone the finished arena onto the others (device to device)
  // and put the replicas behind one group: a batch is then cut into one slice per GPU.
  void finalize(const std::vector<int>& devices = {}) {
    check(ws_index_finalize(idx_), "ws_index_finalize");
    replicate(devices);
  }

  // A finished arena read back from one file (ws_index_load): no sort, no tree derivation, no graph files.
  void load_snapshot(const std::string& path, const std::vector<int>& devices) {
    check(ws_index_load(path.c_str(), devices.empty() ? default_device() : devices[0], &idx_), "ws_index_load");
    uint64_t n = 0;
    uint32_t dim = 0;
    check(ws_index_shape(idx_, &n, &dim, nullptr, nullptr), "ws_index_shape");
    n_ = n; dim_ = dim;
    replicate(devices);
  }
  void save_snapshot(const std::string& path) const { check(ws_index_save(idx_, path.c_str()), "ws_index_save"); }

  void replicate(const std::vector<int>& devices) {
    int own = -1;
    ws_index_device(idx_, &own);
    if (own < 0 || devices.size() < 2) return;
    std::vector<ws_index*> members{idx_};
    for (size_t i = 1; i < devices.size(); i++) {  // devices[0] is this arena's own device
      const int d = devices[i];
      ws_index* r = nullptr;
      check(ws_index_replicate(idx_, d, &r), "ws_index_replicate");
      replicas_.push_back(r);
      members.push_back(r);
    }
    if (members.size() > 1) check(ws_group_create(members.data(), (int)members.size(), WS_GROUP_REPLICATED, &group_), "ws_group_create");
  }
  ws_group* group() const { return group_; }

  // PostfilterVamanaIndex over the points as given (postfilter_vamana.h:91-124)
  void init_unsorted(const float* points, const float* labels, size_t n, size_t dim, int metric, int device) {
    n_ = n; dim_ = dim;
    labels_.assign(labels, labels + n);
    check(ws_index_create(device, metric, n, (uint32_t)dim, points, labels, nullptr, 0, &idx_), "ws_index_create");
  }

  // Queue one node's graph.  Nothing is loaded or built until realize_graphs().
  size_t plan_graph(size_t start, size_t count, float min_label, float max_label) {
    plans_.push_back(Plan{start, count, min_label, max_label});
    return plans_.size() - 1;
  }

  // Resolve every planned graph to a node handle, in plan order.  Graphs whose cache file
  // exists (postfilter_vamana.h:54-61) are loaded; the missing ones are built together on
  // the device with the reference builder's algorithm (ws_build_graphs) and — when
  // cache_path is set — saved in the reference's format, so that the reference loads the
  // same graphs.  WSANN_GRAPH_BUILD=0 turns a cache miss into an error instead.
  std::vector<int32_t> realize_graphs(const BuildParams& bp) {
    std::vector<int32_t> handles(plans_.size(), -1);
    std::vector<size_t> missing;
    for (size_t i = 0; i < plans_.size(); i++) {
      const Plan& pl = plans_[i];
      std::string path = bp.cache_path.empty() ? std::string() : graph_filename(bp, pl.min_label, pl.max_label, pl.count);
      std::ifstream probe;
      if (!path.empty()) probe.open(path, std::ios::binary);
      if (path.empty() || !probe) { missing.push_back(i); continue; }
      probe.close();
      GraphFile g = read_graph_file(path);
      if ((size_t)g.n != pl.count)
        throw std::runtime_error("graph " + path + " has " + std::to_string(g.n) + " nodes, expected " + std::to_string(pl.count));
      check(ws_index_add_graph(idx_, pl.start, pl.count, (uint32_t)g.max_degree, g.degrees.data(), g.edges.data(), &handles[i]),
            "ws_index_add_graph");
    }
    if (!missing.empty()) {
      const char* env = std::getenv("WSANN_GRAPH_BUILD");
      if (env && std::string(env) == "0") {
        const Plan& pl = plans_[missing[0]];
        throw std::runtime_error("graph cache miss: " + graph_filename(bp, pl.min_label, pl.max_label, pl.count) +
                                 " (and " + std::to_string(missing.size() - 1) + " more) with WSANN_GRAPH_BUILD=0");
      }
      int device = -1;
      ws_index_device(idx_, &device);
      if (device < 0) {  // host-only geometry index: nodes without adjacency
        for (size_t i : missing)
          check(ws_index_add_graph(idx_, plans_[i].start, plans_[i].count, (uint32_t)bp.R, nullptr, nullptr, &handles[i]), "ws_index_add_graph");
      } else {
        std::vector<uint64_t> starts, counts;
        uint64_t rows = 0;
        for (size_t i : missing) { starts.push_back(plans_[i].start); counts.push_back(plans_[i].count); rows += plans_[i].count; }
        std::vector<int32_t> built(missing.size(), -1);
        std::fprintf(stderr, "[wsann] %zu graph(s) missing from cache '%s': building %llu node-rows on the GPU (R=%ld L=%ld alpha=%g)\n",
                     missing.size(), bp.cache_path.c_str(), (unsigned long long)rows, bp.R, bp.L, bp.alpha);
        check(ws_build_graphs(idx_, (uint32_t)missing.size(), starts.data(), counts.data(), (uint32_t)bp.R, (uint32_t)bp.L,
                              bp.alpha, 0x5eedull, built.data()),
              "ws_build_graphs");
        for (size_t m = 0; m < missing.size(); m++) {
          handles[missing[m]] = built[m];
          if (!bp.cache_path.empty()) save_graph(bp, plans_[missing[m]], built[m]);
        }
      }
    }
    plans_.clear();
    return handles;
  }

  ws_index* get() const { return idx_; }
  size_t size() const { return n_; }
  size_t dim() const { return dim_; }
  const std::vector<float>& labels() const { return labels_; }

 private:
  struct Plan {
    size_t start, count;
    float min_label, max_label;
  };

  // graph.h:174-196 layout, so the reference's Graph(char*) constructor reads it back
  void save_graph(const BuildParams& bp, const Plan& pl, int32_t node) {
    uint32_t R = 0;
    std::vector<int32_t> deg(pl.count);
    check(ws_index_get_graph(idx_, node, &R, nullptr, nullptr), "ws_index_get_graph");
    std::vector<int32_t> rows(pl.count * (size_t)R);
    check(ws_index_get_graph(idx_, node, &R, deg.data(), rows.data()), "ws_index_get_graph");
    std::string path = graph_filename(bp, pl.min_label, pl.max_label, pl.count);
    {
      std::error_code ec;
      auto dir = std::filesystem::path(path).parent_path();
      if (!dir.empty()) std::filesystem::create_directories(dir, ec);
    }
    std::string tmp = path + ".tmp";
    {
      std::ofstream out(tmp, std::ios::binary);
      if (!out) throw std::runtime_error("cannot write graph cache file " + tmp);
      int32_t hdr[2] = {(int32_t)pl.count, (int32_t)bp.R};
      out.write(reinterpret_cast<const char*>(hdr), 8);
      out.write(reinterpret_cast<const char*>(deg.data()), 4ll * (long long)pl.count);
      std::vector<int32_t> edges;
      edges.reserve(pl.count * 32);
      for (size_t i = 0; i < pl.count; i++)
        for (int32_t j = 0; j < deg[i]; j++) edges.push_back(rows[i * R + j]);
      out.write(reinterpret_cast<const char*>(edges.data()), 4ll * (long long)edges.size());
      if (!out) throw std::runtime_error("short write on " + tmp);
    }
    if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("cannot rename " + tmp);
  }

  ws_index* idx_ = nullptr;
  std::vector<ws_index*> replicas_;
  ws_group* group_ = nullptr;
  size_t n_ = 0, dim_ = 0;
  std::vector<float> labels_;
  std::vector<Plan> plans_;
};

// ---- label-range shards ------------------------------------------------------
// The label-sorted points are cut into one contiguous range per device; every shard is an Arena of its own (own
// tree, own graphs, decode table = original ids), built on its device by its own host thread.  One group call
// answers a batch on all shards and merges the partial rows (ws_group, WS_GROUP_LABEL_SHARDED).
class LabelShards {
 public:
  LabelShards() = default;
  LabelShards(const LabelShards&) = delete;
  LabelShards& operator=(const LabelShards&) = delete;
  ~LabelShards() { if (group_) ws_group_destroy(group_); }

  // per_shard(arena, shard, count) stages geometry / graphs on an arena whose points are already uploaded
  template <class F>
  void init(const float* points, const float* labels, size_t n, size_t dim, int metric, const std::vector<int>& devices,
            F&& per_shard) {
    const size_t G = devices.size();
    if (n < G) throw std::runtime_error("fewer points than label shards");
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return labels[a] < labels[b]; });
    arenas_.resize(G);
    std::vector<std::string> errors(G);
    std::vector<std::thread> threads;
    for (size_t s = 0; s < G; s++) {
      arenas_[s] = std::make_unique<Arena>();
      threads.emplace_back([&, s]() {
        try {
          auto [lo, hi] = shard_bounds(n, s, G);
          const size_t cnt = hi - lo;
          std::vector<float> pts(cnt * dim), lab(cnt);
          for (size_t i = 0; i < cnt; i++) {
            std::memcpy(&pts[i * dim], points + (size_t)order[lo + i] * dim, dim * sizeof(float));
            lab[i] = labels[order[lo + i]];
          }
          Arena& a = *arenas_[s];
          a.init_presorted(pts.data(), lab.data(), order.data() + lo, cnt, dim, metric, devices[s]);
          per_shard(a, s, cnt);
          a.finalize();
          // the reference's r = n-1 rule (prefiltering.h:159-184) excludes the DATA SET's last point, not a shard's
          if (s + 1 < G) check(ws_index_set_option(a.get(), "prefilter_open_tail", 1), "ws_index_set_option");
        } catch (const std::exception& e) {
          errors[s] = e.what();
          if (errors[s].empty()) errors[s] = "unknown error";
        }
      });
    }
    for (std::thread& t : threads) t.join();
    for (size_t s = 0; s < G; s++)
      if (!errors[s].empty()) throw std::runtime_error("label shard " + std::to_string(s) + ": " + errors[s]);
    std::vector<ws_index*> members;
    for (auto& a : arenas_) members.push_back(a->get());
    check(ws_group_create(members.data(), (int)members.size(), WS_GROUP_LABEL_SHARDED, &group_), "ws_group_create");
  }
  bool active() const { return group_ != nullptr; }
  ws_group* group() const { return group_; }
  Arena& shard(size_t i) { return *arenas_[i]; }
  size_t size() const { return arenas_.size(); }

 private:
  std::vector<std::unique_ptr<Arena>> arenas_;
  ws_group* group_ = nullptr;
};

// constructor tag: build the index from a file written by save_snapshot() of the same class
struct FromSnapshot {
  std::string path;
};

struct BatchResult {
  std::vector<uint32_t> ids;
  std::vector<float> dists;
};

// ---- PrefilterIndex (src/prefiltering.h) -------------------------------------------------------
class PrefilterIndex {
 public:
  PrefilterIndex(const float* points, const float* labels, size_t n, size_t dim, int metric, const BuildParams&,
                 const DevicePlan& plan = device_plan()) : dim_(dim) {
    if (plan.label_sharded) {
      shards_.init(points, labels, n, dim, metric, plan.devices, [](Arena&, size_t, size_t) {});
      return;
    }
    arena_.init_sorted(points, labels, n, dim, metric, plan.devices[0]);
    arena_.finalize(plan.devices);
  }
  PrefilterIndex(const FromSnapshot& snap, const DevicePlan& plan = device_plan()) {
    arena_.load_snapshot(snap.path, plan.devices);
    dim_ = arena_.dim();
  }
  void save_snapshot(const std::string& path) {
    if (shards_.active()) throw std::runtime_error("a label-sharded index is saved shard by shard");
    arena_.save_snapshot(path);
  }
  // prefiltering.h:124-146
  void batch_search(const float* queries, const float* filters, uint64_t nq, const QueryParams& qp, uint32_t* ids,
                    float* dists) {
    if (qp.k < 1) throw std::runtime_error("k must be >= 1");
    ws_group* g = shards_.active() ? shards_.group() : arena_.group();
    if (g) check(ws_group_prefilter_batch(g, queries, filters, nq, (uint32_t)qp.k, ids, dists), "ws_group_prefilter_batch");
    else check(ws_prefilter_batch(arena_.get(), queries, filters, nq, (uint32_t)qp.k, ids, dists, 0), "ws_prefilter_batch");
  }
  Arena& arena() { return shards_.active() ? shards_.shard(0) : arena_; }
  ws_group* group() { return shards_.active() ? shards_.group() : arena_.group(); }
  size_t dim() const { return dim_; }

 private:
  Arena arena_;
  LabelShards shards_;
  size_t dim_ = 0;
};

// ---- PostfilterVamanaIndex (src/postfilter_vamana.h) -------------------------------------------
class PostfilterVamanaIndex {
 public:
  PostfilterVamanaIndex(const float* points, const float* labels, size_t n, size_t dim, int metric,
                        const BuildParams& bp, const DevicePlan& plan = device_plan()) {
    arena_.init_unsorted(points, labels, n, dim, metric, plan.devices[0]);
    float mn = *std::min_element(labels, labels + n), mx = *std::max_element(labels, labels + n);
    arena_.plan_graph(0, n, mn, mx);
    node_ = arena_.realize_graphs(bp)[0];
    arena_.finalize(plan.devices);
  }
  PostfilterVamanaIndex(const FromSnapshot& snap, const DevicePlan& plan = device_plan()) {
    arena_.load_snapshot(snap.path, plan.devices);
    node_ = 0;
  }
  void save_snapshot(const std::string& path) { arena_.save_snapshot(path); }
  // postfilter_vamana.h:191-219 (missing slots: id 0xFFFFFFFF, FLT_MAX)
  void batch_search(const float* queries, const float* filters, uint64_t nq, const QueryParams& qp, uint32_t* ids,
                    float* dists) {
    ws_query_params c = qp.to_c();
    if (arena_.group())
      check(ws_group_postfilter_batch(arena_.group(), node_, queries, filters, nq, &c, WS_PAD_MINUS1, ids, dists),
            "ws_group_postfilter_batch");
    else
      check(ws_postfilter_batch(arena_.get(), node_, queries, filters, nq, &c, WS_PAD_MINUS1, ids, dists, 0),
            "ws_postfilter_batch");
  }
  ws_group* group() { return arena_.group(); }
  Arena& arena() { return arena_; }
  size_t dim() const { return arena_.dim(); }

 private:
  Arena arena_;
  int32_t node_ = -1;
};

// ---- RangeFilterTreeIndex (src/range_filter_tree.h) --------------------------------------------
// The reference instantiates the template twice (python_bindings.cpp:119-127,136-145): with
// PostfilterVamanaIndex sub-indices ("VamanaRangeFilterTreeIndex…") and with PrefilterIndex
// sub-indices ("RangeFilterTreeIndex…", the template's default, range_filter_tree.h:32).  Geometry,
// query methods and result conventions are the same; only what answers a bucket differs.
class RangeFilterTreeBase {
 protected:
  RangeFilterTreeBase(const float* points, const float* labels, size_t n, size_t dim, int metric,
                      int32_t cutoff, size_t split_factor, const BuildParams& bp, bool vamana_nodes,
                      const DevicePlan& plan) : dim_(dim) {
    if (split_factor < 2) throw std::runtime_error("split_factor must be at least 2");
    if (plan.label_sharded) {
      // every shard is a B-WST of its own over its label range (the reference's tree cut at the shard boundaries)
      std::vector<std::vector<std::vector<uint64_t>>> per(plan.devices.size());
      shards_.init(points, labels, n, dim, metric, plan.devices, [&](Arena& a, size_t s, size_t cnt) {
        per[s] = stage_tree(a, cnt, cutoff, split_factor, bp, vamana_nodes);
      });
      offsets_ = per[0];
      return;
    }
    arena_.init_sorted(points, labels, n, dim, metric, plan.devices[0]);
    offsets_ = stage_tree(arena_, n, cutoff, split_factor, bp, vamana_nodes);
    arena_.finalize(plan.devices);
  }

  explicit RangeFilterTreeBase(const FromSnapshot& snap, const DevicePlan& plan) {
    arena_.load_snapshot(snap.path, plan.devices);
    dim_ = arena_.dim();
  }

  // range_filter_tree.h:129-189 over the arena's n points: bucket offsets per row, one graph per bucket
  static std::vector<std::vector<uint64_t>> stage_tree(Arena& arena, size_t n, int32_t cutoff, size_t split_factor,
                                                       const BuildParams& bp, bool vamana_nodes) {
    std::vector<std::vector<uint64_t>> offsets;
    const std::vector<float>& sl = arena.labels();
    offsets.push_back({0, (uint64_t)n});
    while ((int64_t)offsets.back()[1] > (int64_t)cutoff) {
      const std::vector<uint64_t>& last = offsets.back();
      size_t last_nb = last.size() - 1;
      std::vector<uint64_t> next(last_nb * split_factor + 1);
      next.back() = n;
      for (size_t b = 0; b < last_nb; b++) {
        uint64_t ls = last[b], le = last[b + 1], size = le - ls;
        uint64_t large = (size + split_factor - 1) / split_factor;
        uint64_t small = large - 1;
        uint64_t n_large = size - small * split_factor;
        for (size_t i = 0; i < split_factor; i++) {
          uint64_t s = i < n_large ? ls + i * large : ls + n_large * large + (i - n_large) * small;
          next[b * split_factor + i] = s;
        }
      }
      for (size_t i = 0; i + 1 < next.size(); i++)
        if (next[i + 1] <= next[i]) throw std::runtime_error("cutoff/split_factor produce an empty bucket");
      offsets.push_back(std::move(next));
    }
    std::vector<uint32_t> row_nb;
    std::vector<uint64_t> off_flat;
    for (auto& row : offsets) {
      row_nb.push_back((uint32_t)row.size() - 1);
      off_flat.insert(off_flat.end(), row.begin(), row.end());
      if (vamana_nodes)
        for (size_t b = 0; b + 1 < row.size(); b++)
          arena.plan_graph(row[b], row[b + 1] - row[b], sl[row[b]], sl[row[b + 1] - 1]);
    }
    // no node handles = PrefilterIndex sub-indices: a bucket query is a scan of the bucket's slice
    std::vector<int32_t> nodes_flat;
    if (vamana_nodes) nodes_flat = arena.realize_graphs(bp);
    check(ws_index_set_wst(arena.get(), (uint32_t)offsets.size(), (uint32_t)split_factor, cutoff, row_nb.data(),
                           off_flat.data(), vamana_nodes ? nodes_flat.data() : nullptr),
          "ws_index_set_wst");
    return offsets;
  }

 public:

  // range_filter_tree.h:62-96: any method string other than the two named ones is "fenwick"
  static int method_from_string(const std::string& m) {
    if (m == "optimized_postfilter") return WS_METHOD_OPT_POSTFILTER;
    if (m == "three_split") return WS_METHOD_THREE_SPLIT;
    return WS_METHOD_FENWICK;
  }
  void batch_search(const float* queries, const float* filters, uint64_t nq, const std::string& query_method,
                    const QueryParams& qp, uint32_t* ids, float* dists) {
    ws_query_params c = qp.to_c();
    ws_group* g = group();
    if (g) check(ws_group_tree_batch(g, method_from_string(query_method), queries, filters, nq, &c, ids, dists), "ws_group_tree_batch");
    else check(ws_tree_batch(arena_.get(), method_from_string(query_method), queries, filters, nq, &c, ids, dists, 0), "ws_tree_batch");
  }
  void save_snapshot(const std::string& path) {
    if (shards_.active()) throw std::runtime_error("a label-sharded index is saved shard by shard");
    arena_.save_snapshot(path);
  }
  Arena& arena() { return shards_.active() ? shards_.shard(0) : arena_; }
  ws_group* group() { return shards_.active() ? shards_.group() : arena_.group(); }
  size_t dim() const { return dim_; }
  // bucket offsets of the tree (of shard 0's tree when label-sharded; empty for an index loaded from a snapshot)
  const std::vector<std::vector<uint64_t>>& bucket_offsets() const { return offsets_; }

 private:
  Arena arena_;
  LabelShards shards_;
  size_t dim_ = 0;
  std::vector<std::vector<uint64_t>> offsets_;
};

// RangeFilterTreeIndex<T, Point, PostfilterVamanaIndex<T, Point>> (python_bindings.cpp:136-145)
class VamanaRangeFilterTreeIndex : public RangeFilterTreeBase {
 public:
  VamanaRangeFilterTreeIndex(const float* points, const float* labels, size_t n, size_t dim, int metric,
                             int32_t cutoff, size_t split_factor, const BuildParams& bp,
                             const DevicePlan& plan = device_plan())
      : RangeFilterTreeBase(points, labels, n, dim, metric, cutoff, split_factor, bp, true, plan) {}
  VamanaRangeFilterTreeIndex(const FromSnapshot& snap, const DevicePlan& plan = device_plan()) : RangeFilterTreeBase(snap, plan) {}
};

// RangeFilterTreeIndex<T, Point> = PrefilterIndex sub-indices (python_bindings.cpp:119-127)
class RangeFilterTreeIndex : public RangeFilterTreeBase {
 public:
  RangeFilterTreeIndex(const float* points, const float* labels, size_t n, size_t dim, int metric,
                       int32_t cutoff, size_t split_factor, const BuildParams& bp, const DevicePlan& plan = device_plan())
      : RangeFilterTreeBase(points, labels, n, dim, metric, cutoff, split_factor, bp, false, plan) {}
  RangeFilterTreeIndex(const FromSnapshot& snap, const DevicePlan& plan = device_plan()) : RangeFilterTreeBase(snap, plan) {}
};

// ---- SuperOptimizedPostfilterTree (src/super_optimized_postfilter_tree.h) ------------------------
class SuperOptimizedPostfilterTree {
 public:
  SuperOptimizedPostfilterTree(const float* points, const float* labels, size_t n, size_t dim, int metric,
                               int32_t cutoff, float split_factor, float shift_factor, const BuildParams& bp,
                               const DevicePlan& plan = device_plan()) {
    // super_optimized_postfilter_tree.h:127-132
    if (split_factor <= 1) throw std::runtime_error("split_factor must be greater than 1");
    if (shift_factor >= 1 || shift_factor <= 0) throw std::runtime_error("shift_factor must be between 0 and 1");
    arena_.init_sorted(points, labels, n, dim, metric, plan.devices[0]);
    const std::vector<float>& sl = arena_.labels();
    std::vector<uint64_t> sizes{(uint64_t)n}, shifts{0};
    std::vector<uint32_t> row_nb{1};
    arena_.plan_graph(0, n, sl[0], sl[n - 1]);
    // :145-170 — bucket size is evaluated in float, as the reference does
    while ((int64_t)sizes.back() > (int64_t)cutoff) {
      size_t last = sizes.back();
      size_t bucket_size = (size_t)((last + split_factor - 1) / split_factor);
      size_t bucket_shift = (size_t)std::ceil(bucket_size * shift_factor);
      if (bucket_size == 0 || bucket_shift == 0 || bucket_size >= last) throw std::runtime_error("degenerate super tree row");
      sizes.push_back(bucket_size);
      shifts.push_back(bucket_shift);
      size_t nb = ((n - bucket_size) + bucket_shift - 1) / bucket_shift + 1;
      row_nb.push_back((uint32_t)nb);
      for (size_t b = 0; b < nb; b++) {
        size_t s = b * bucket_shift, e = std::min(s + bucket_size, n);
        arena_.plan_graph(s, e - s, sl[s], sl[e - 1]);
      }
    }
    std::vector<int32_t> nodes_flat = arena_.realize_graphs(bp);
    check(ws_index_set_super(arena_.get(), (uint32_t)sizes.size(), cutoff, sizes.data(), shifts.data(), row_nb.data(),
                             nodes_flat.data()),
          "ws_index_set_super");
    arena_.finalize(plan.devices);
    sizes_ = sizes; shifts_ = shifts;
  }
  SuperOptimizedPostfilterTree(const FromSnapshot& snap, const DevicePlan& plan = device_plan()) {
    arena_.load_snapshot(snap.path, plan.devices);
  }
  void save_snapshot(const std::string& path) { arena_.save_snapshot(path); }
  // super_optimized_postfilter_tree.h:60-87
  void batch_search(const float* queries, const float* filters, uint64_t nq, const QueryParams& qp, uint32_t* ids,
                    float* dists) {
    ws_query_params c = qp.to_c();
    if (arena_.group())
      check(ws_group_tree_batch(arena_.group(), WS_METHOD_SUPER_POSTFILTER, queries, filters, nq, &c, ids, dists), "ws_group_tree_batch");
    else
      check(ws_tree_batch(arena_.get(), WS_METHOD_SUPER_POSTFILTER, queries, filters, nq, &c, ids, dists, 0), "ws_tree_batch");
  }
  ws_group* group() { return arena_.group(); }
  Arena& arena() { return arena_; }
  size_t dim() const { return arena_.dim(); }
  const std::vector<uint64_t>& bucket_sizes() const { return sizes_; }
  const std::vector<uint64_t>& bucket_shifts() const { return shifts_; }

 private:
  Arena arena_;
  std::vector<uint64_t> sizes_, shifts_;
};

}  // namespace wsann
