// wsann.cu — libwsann_cuda.so: HBM arena management, batch orchestration and the C ABI
// declared in include/wsann.h.  No PyTorch, no CPU fallback: every *_batch call needs a
// CUDA device and returns WS_ERR_CUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/wsann.h"
#include "ws_launch.h"
#include "ws_gemm.h"

#include <cmath>
#include <random>

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

static int ws_fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define WS_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return ws_fail(_e == cudaErrorMemoryAllocation ? WS_ERR_OOM : WS_ERR_CUDA,            \
                     "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,      \
                     __LINE__);                                                             \
  } while (0)

#define WS_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != WS_OK) return _s; \
  } while (0)

// ------------------------------------------------------------------------------------------
// index object
// ------------------------------------------------------------------------------------------
// beam tiers: 0..4 = warp-per-task kernels (cap 64 / 128 / 256 / 512 / 1024, visited table in shared memory),
// 5 = CTA-per-task with a global visited bitmap (cap 12288)
static const uint32_t kBeamTierCaps[5] = {64, 128, 256, 512, 1024};
static const int kBeamTierCS[5] = {7, 7, 8, 9, 10};
#define WS_NUM_WARP_TIERS 5
#define WS_NUM_TIERS 6
static const uint32_t kBeamCapLarge = 12288;                // global-bitmap tier
static const uint32_t kMaxK = WS_TOPK_BUF / 2;
static const size_t kAdjSlabBytes = 256ull << 20;
#define WS_NUM_KERNEL_KINDS 13  // 0 decompose, 1-5 warp beam tiers 64/128/256/512/1024, 6 beam large, 7 scan, 8 merge,
                                // 9 tensor-core prefilter sweep, 10 its bounds+plan+pack, 11 its re-rank, 12 its seed thresholds

struct WsDevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct ws_index {
  // One batch at a time per index: the *_batch / build / merge entry points mutate per-index scratch
  // (task slots, queues, control words, one stream), so they serialise here.  The reference's Python
  // callers hold the GIL for a whole batch_search; this engine's bindings release it.
  std::recursive_mutex mu;
  int device = -1;  // -1: host-only geometry index (decomposition tests)
  int metric = 0;
  uint64_t n = 0;
  uint32_t dim = 0, dpad = 0;
  bool label_sorted = false;
  bool finalized = false;
  bool has_decode = false;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int num_sms = 0;
  size_t smem_optin = 0;

  // device arena
  float* d_vecs = nullptr;
  float* d_labels = nullptr;
  uint32_t* d_decode = nullptr;
  WsNode* d_nodes = nullptr;
  std::vector<void*> adj_slabs;
  std::vector<void*> build_allocs;   // adjacency/degree arrays produced by ws_build_graphs
  std::vector<int32_t*> node_deg;     // per node: device degree array (built graphs) or nullptr
  size_t slab_used = 0;
  uint64_t hbm_bytes = 0;

  // host mirrors
  std::vector<float> h_labels;
  std::vector<WsNode> h_nodes;
  uint32_t R = 0;
  uint32_t max_node_count = 0;

  // geometry (host copies own the storage the host WsGeom points to)
  std::vector<uint32_t> wst_nb, wst_off_ptr, wst_node_ptr;
  std::vector<uint64_t> wst_off;
  std::vector<int32_t> wst_nodes;
  std::vector<uint64_t> sup_size, sup_shift;
  std::vector<uint32_t> sup_nb, sup_node_ptr;
  std::vector<int32_t> sup_nodes;
  uint32_t wst_rows = 0, split = 0, sup_rows = 0;
  bool wst_prefilter_nodes = false;  // B-WST buckets answered by PrefilterIndex sub-indices (no graphs)
  int32_t cutoff = 0, sup_cutoff = 0;
  WsGeom hgeom{};
  WsGeom dgeom{};
  std::vector<void*> geom_dev_allocs;

  // options
  int64_t opt_expand = 1;
  int64_t opt_skip_query_id = 1;
  int64_t opt_scan_chunk = 8192;
  int64_t opt_hash_factor = 32;
  int64_t opt_warp_tiers = 1;    // use the warp-per-task kernels for beams <= 128
  int64_t opt_warp_hash = 2048;
  int64_t opt_warp_scan = 1;     // warp-per-task scan kernel for k <= 128
  int64_t opt_hash16 = 1;        // 16-bit visited tags in the warp tiers when node sizes allow
  int64_t opt_warp256 = 1;       // escalated beams 129..256: warp kernel when many tasks are queued, CTA tier otherwise
  int64_t opt_warp256_min = 384; // ... threshold on the number of queued tasks
  int64_t opt_fuse_scan = 0;     // let the first warp-tier beam launch drain the scan queue too (measured neutral)  // visited-table entries per warp in those kernels
  int64_t opt_build_expand = 1;  // nodes expanded per step while BUILDING graphs
  uint64_t build_stats[4] = {0, 0, 0, 0};  // inserts, visited, dist_cmps, overflow re-prunes  // smem visited-table entries per unit of beam capacity

  // scratch (grown on demand)
  WsDevBuf tasks, res_keys, res_cnt, counts, queues, ctrl, d_queries, d_windows, d_ids, d_dists,
      bitmap, flush;
  // multi-GPU (ws_group.inl): NCCL communicator of this arena's rank, gathered partial rows, and — for
  // device-pointer batches issued by a group — the host copy of the batch's windows (routing sample)
  void* comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  WsDevBuf x_all_ids, x_all_dists;
  // pinned staging of host-buffer batches (ws_stage_h2d)
  WsDevBuf stage_in, stage_out;
  std::vector<std::atomic<uint32_t>> stage_flags;
  int64_t opt_stage_copies = 1;
  const float* hint_windows = nullptr;
  uint32_t query_id_base = 0;  // position of this batch's first query in the caller's batch (ws_group slices)
  unsigned long long* d_stats = nullptr;
  // sticky error word, never cleared by a batch: bit 0 = task-slot capacity overflow (tasks dropped), bit 1 = a task
  // outgrew the last beam tier and was dropped.  Read (and cleared) by host-buffer batches before they return and by
  // ws_index_sync for device-pointer batches, which only enqueue work.
  uint32_t* d_sticky = nullptr;
  uint64_t launches = 0;

  // tensor-core prefilter (ws_gemm.cuh)
  int64_t opt_open_tail = 0;     // 1: prefilter windows may include the arena's last point (label shards, WsGeom::pf_n)
  int64_t opt_direct = 2;        // one-launch prefilter (K1d): 0 never, 1 always, 2 auto (host-sampled mean window <= scan_chunk)
  int64_t opt_gemm = 2;          // 0 never, 1 whenever eligible, 2 auto (host-sampled mean window >= opt_gemm_min_window)
  int64_t opt_gemm_min_window = 768;
  int64_t opt_gemm_dynamic = 1;  // sweep kernel draws work items from a device counter (0: static striping)
  int64_t opt_gemm_items = 0;    // target work items per plan (0: 2 per SM)
  int64_t opt_gemm_min_tiles = 8;
  int64_t opt_gemm_debug = 0;     // timing experiments (ws_gemm.h WsGemmArgs::dbg); results are invalid when set
  int64_t opt_gemm_chunk_mb = 8;  // largest slice of the label axis one work item sweeps
  bool gemm_ready = false;
  bool gemm_attr_set = false;
  bool gemm_unfit = false;        // the arena cannot be mirrored in fp16 (non-finite or absurdly scaled components)
  int32_t g_x_exp = 0;            // the fp16 mirror holds half(x * 2^g_x_exp)
  WsDevBuf g_norms, g_ctrl, g_perm, g_row_a, g_row_b, g_items, g_group_items, g_group_cnt, g_qpack, g_rscale, g_vecs16, g_slack, g_cand,
      g_cand_cnt, g_cand_thr, g_res_keys, g_res_cnt, g_thr0, g_qnorm, g_qa, g_qb;
  CUtensorMap g_tm_b{};

  // optional per-kernel CUDA-event timing (ws_index_kernel_times)
  int64_t opt_profile = 0;
  std::vector<cudaEvent_t> ev_pool;
  struct EvPair { int kind; size_t a, b; };
  std::vector<EvPair> ev_pairs;
  size_t ev_used = 0;
  double kernel_ms[WS_NUM_KERNEL_KINDS] = {0};
  uint64_t kernel_launches[WS_NUM_KERNEL_KINDS] = {0};
};

// brackets one kernel launch with events on the index stream when profiling is on
struct WsKernelScope {
  ws_index* idx;
  int kind;
  size_t a = 0;
  bool on;
  static size_t grab(ws_index* idx) {
    if (idx->ev_used == idx->ev_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      idx->ev_pool.push_back(e);
    }
    return idx->ev_used++;
  }
  WsKernelScope(ws_index* i, int k) : idx(i), kind(k), on(i->opt_profile != 0) {
    if (on) { a = grab(idx); cudaEventRecord(idx->ev_pool[a], idx->stream); }
  }
  ~WsKernelScope() {
    idx->launches++;
    if (on) {
      size_t b = grab(idx);
      cudaEventRecord(idx->ev_pool[b], idx->stream);
      idx->ev_pairs.push_back({kind, a, b});
    }
  }
};

static int ws_ensure(ws_index* idx, WsDevBuf& b, size_t bytes) {
  if (b.bytes >= bytes) return WS_OK;
  if (b.p) {
    WS_CUDA(cudaStreamSynchronize(idx->stream));
    WS_CUDA(cudaFree(b.p));
    idx->hbm_bytes -= b.bytes;
    b.p = nullptr;
    b.bytes = 0;
  }
  size_t want = bytes + bytes / 4 + 256;
  WS_CUDA(cudaMalloc(&b.p, want));
  b.bytes = want;
  idx->hbm_bytes += want;
  return WS_OK;
}

static uint32_t ws_dim_round_up(uint32_t dim) {
  // point_range.h:39-44 with sizeof(T) = 4: rows are a multiple of 64 bytes
  uint64_t bytes = (uint64_t)dim * 4;
  if (bytes % 64 == 0) return dim;
  return (uint32_t)(((bytes / 64) + 1) * 64 / 4);
}

// ------------------------------------------------------------------------------------------
extern "C" {

const char* ws_last_error(void) { return g_last_error.c_str(); }
int ws_abi_version(void) { return WSANN_ABI_VERSION; }

int ws_device_count(int* count) {
  if (!count) return ws_fail(WS_ERR_BADARG, "count is null");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    *count = 0;
    return ws_fail(WS_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
  }
  *count = c;
  return WS_OK;
}

int ws_index_create(int device, int metric, uint64_t n, uint32_t dim, const float* vectors,
                    const float* labels, const uint32_t* decode, int label_sorted,
                    ws_index** out) {
  if (!out) return ws_fail(WS_ERR_BADARG, "out is null");
  *out = nullptr;
  if (n == 0 || dim == 0) return ws_fail(WS_ERR_BADARG, "empty index (n=%llu dim=%u)", (unsigned long long)n, dim);
  if (n >= (1ull << 31)) return ws_fail(WS_ERR_BADARG, "n=%llu exceeds the 2^31 ranks one arena addresses", (unsigned long long)n);
  if (metric != WS_METRIC_L2 && metric != WS_METRIC_MIPS) return ws_fail(WS_ERR_BADARG, "unknown metric %d", metric);
  if (!labels) return ws_fail(WS_ERR_BADARG, "labels is null");
  if (label_sorted)
    for (uint64_t i = 1; i < n; i++)
      if (labels[i] < labels[i - 1]) return ws_fail(WS_ERR_BADARG, "labels not sorted at %llu", (unsigned long long)i);
  ws_index* idx = new ws_index();
  idx->device = device;
  idx->metric = metric;
  idx->n = n;
  idx->dim = dim;
  idx->dpad = ws_dim_round_up(dim);
  idx->label_sorted = label_sorted != 0;
  idx->h_labels.assign(labels, labels + n);
  if (device < 0) {  // host-only geometry index
    *out = idx;
    return WS_OK;
  }
  if (!vectors) { delete idx; return ws_fail(WS_ERR_BADARG, "vectors is null"); }
  if (idx->dpad > 1024) { delete idx; return ws_fail(WS_ERR_BADARG, "dim %u: padded rows above 1024 floats are not supported", dim); }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device >= ndev) {
    delete idx;
    return ws_fail(WS_ERR_CUDA, "no usable CUDA device %d (%s); this engine has no CPU fallback", device,
                   e != cudaSuccess ? cudaGetErrorString(e) : "index out of range");
  }
#define WS_CREATE_CUDA(expr)                                                                 \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      ws_index_destroy(idx);                                                                 \
      return ws_fail(_e == cudaErrorMemoryAllocation ? WS_ERR_OOM : WS_ERR_CUDA, "%s: %s", #expr, \
                     cudaGetErrorString(_e));                                                \
    }                                                                                        \
  } while (0)
  WS_CREATE_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  WS_CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
  idx->num_sms = prop.multiProcessorCount;
  idx->smem_optin = prop.sharedMemPerBlockOptin;
  WS_CREATE_CUDA(cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking));
  WS_CREATE_CUDA(cudaEventCreate(&idx->ev0));
  WS_CREATE_CUDA(cudaEventCreate(&idx->ev1));
  size_t vbytes = (size_t)n * idx->dpad * sizeof(float);
  WS_CREATE_CUDA(cudaMalloc(&idx->d_vecs, vbytes));
  idx->hbm_bytes += vbytes;
  if (idx->dpad == dim) {
    WS_CREATE_CUDA(cudaMemcpy(idx->d_vecs, vectors, vbytes, cudaMemcpyHostToDevice));
  } else {  // zero-filled pad (the reference leaves it uninitialised)
    WS_CREATE_CUDA(cudaMemset(idx->d_vecs, 0, vbytes));
    WS_CREATE_CUDA(cudaMemcpy2D(idx->d_vecs, (size_t)idx->dpad * 4, vectors, (size_t)dim * 4, (size_t)dim * 4, n,
                                cudaMemcpyHostToDevice));
  }
  WS_CREATE_CUDA(cudaMalloc(&idx->d_labels, n * sizeof(float)));
  WS_CREATE_CUDA(cudaMemcpy(idx->d_labels, labels, n * sizeof(float), cudaMemcpyHostToDevice));
  idx->hbm_bytes += n * sizeof(float);
  if (decode) {
    WS_CREATE_CUDA(cudaMalloc(&idx->d_decode, n * sizeof(uint32_t)));
    WS_CREATE_CUDA(cudaMemcpy(idx->d_decode, decode, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
    idx->hbm_bytes += n * sizeof(uint32_t);
    idx->has_decode = true;
  }
  WS_CREATE_CUDA(cudaMalloc(&idx->d_stats, 8 * sizeof(unsigned long long)));
  WS_CREATE_CUDA(cudaMemset(idx->d_stats, 0, 8 * sizeof(unsigned long long)));
  WS_CREATE_CUDA(cudaMalloc(&idx->d_sticky, 4 * sizeof(uint32_t)));
  WS_CREATE_CUDA(cudaMemset(idx->d_sticky, 0, 4 * sizeof(uint32_t)));
#undef WS_CREATE_CUDA
  *out = idx;
  return WS_OK;
}

void ws_index_destroy(ws_index* idx) {
  if (!idx) return;
  if (idx->device >= 0) {
    cudaSetDevice(idx->device);
    if (idx->stream) cudaStreamSynchronize(idx->stream);
    cudaFree(idx->d_vecs);
    cudaFree(idx->d_labels);
    cudaFree(idx->d_decode);
    cudaFree(idx->d_stats);
    cudaFree(idx->d_sticky);
    for (void* p : idx->adj_slabs) cudaFree(p);
    for (void* p : idx->build_allocs) cudaFree(p);
    for (void* p : idx->geom_dev_allocs) cudaFree(p);
    WsDevBuf* bufs[] = {&idx->tasks, &idx->res_keys, &idx->res_cnt, &idx->counts, &idx->queues, &idx->ctrl,
                        &idx->d_queries, &idx->d_windows, &idx->d_ids, &idx->d_dists, &idx->bitmap, &idx->flush,
                        &idx->x_all_ids, &idx->x_all_dists, &idx->g_norms, &idx->g_ctrl, &idx->g_perm, &idx->g_row_a, &idx->g_row_b, &idx->g_items,
                        &idx->g_group_items, &idx->g_group_cnt, &idx->g_qpack, &idx->g_rscale, &idx->g_vecs16, &idx->g_slack, &idx->g_cand,
                        &idx->g_cand_cnt, &idx->g_cand_thr, &idx->g_res_keys, &idx->g_res_cnt, &idx->g_thr0, &idx->g_qnorm, &idx->g_qa, &idx->g_qb};
    for (WsDevBuf* b : bufs) cudaFree(b->p);
    if (idx->stage_in.p) cudaFreeHost(idx->stage_in.p);
    if (idx->stage_out.p) cudaFreeHost(idx->stage_out.p);
    for (cudaEvent_t e : idx->ev_pool) cudaEventDestroy(e);
    if (idx->ev0) cudaEventDestroy(idx->ev0);
    if (idx->ev1) cudaEventDestroy(idx->ev1);
    if (idx->stream) cudaStreamDestroy(idx->stream);
  }
  delete idx;
}

int ws_index_add_graph(ws_index* idx, uint64_t start, uint64_t count, uint32_t max_degree,
                       const int32_t* degrees, const int32_t* edges, int32_t* node_out) {
  if (!idx || !node_out) return ws_fail(WS_ERR_BADARG, "null argument");
  if (idx->finalized) return ws_fail(WS_ERR_STATE, "index already finalized");
  if (count == 0 || start + count > idx->n) return ws_fail(WS_ERR_BADARG, "graph range [%llu,+%llu) outside the arena", (unsigned long long)start, (unsigned long long)count);
  if (max_degree == 0 || max_degree > 64) return ws_fail(WS_ERR_BADARG, "max_degree %u unsupported (1..64)", max_degree);
  uint32_t R = (max_degree + 3u) & ~3u;
  if (idx->R == 0) idx->R = R;
  if (idx->R != R) return ws_fail(WS_ERR_BADARG, "all graphs of one index must share max_degree (%u vs %u)", idx->R, R);
  WsNode node;
  node.adj = nullptr;
  node.start = (uint32_t)start;
  node.count = (uint32_t)count;
  if (idx->device >= 0) {
    if (!degrees || !edges) return ws_fail(WS_ERR_BADARG, "degrees/edges null");
    WS_CUDA(cudaSetDevice(idx->device));
    size_t bytes = (size_t)count * R * sizeof(int32_t);
    std::vector<int32_t> rows((size_t)count * R, -1);
    size_t off = 0;
    for (uint64_t i = 0; i < count; i++) {
      int32_t deg = degrees[i];
      if (deg < 0 || (uint32_t)deg > max_degree) return ws_fail(WS_ERR_BADARG, "degree %d of local node %llu out of range", deg, (unsigned long long)i);
      for (int32_t j = 0; j < deg; j++) {
        int32_t v = edges[off + j];
        if (v < 0 || (uint64_t)v >= count) return ws_fail(WS_ERR_BADARG, "edge %d of local node %llu outside the graph", v, (unsigned long long)i);
        rows[i * R + j] = v;
      }
      off += deg;
    }
    void* dst = nullptr;
    size_t aligned = (bytes + 255) & ~(size_t)255;
    if (aligned > kAdjSlabBytes) {
      WS_CUDA(cudaMalloc(&dst, aligned));
      idx->adj_slabs.insert(idx->adj_slabs.begin(), dst);  // keep the bump slab last
      idx->hbm_bytes += aligned;
      if (idx->adj_slabs.size() == 1) idx->slab_used = kAdjSlabBytes;  // no bump slab yet
    } else {
      if (idx->adj_slabs.empty() || idx->slab_used + aligned > kAdjSlabBytes) {
        void* slab = nullptr;
        WS_CUDA(cudaMalloc(&slab, kAdjSlabBytes));
        idx->adj_slabs.push_back(slab);
        idx->slab_used = 0;
        idx->hbm_bytes += kAdjSlabBytes;
      }
      dst = (char*)idx->adj_slabs.back() + idx->slab_used;
      idx->slab_used += aligned;
    }
    WS_CUDA(cudaMemcpy(dst, rows.data(), bytes, cudaMemcpyHostToDevice));
    node.adj = (const int32_t*)dst;
  }
  idx->h_nodes.push_back(node);
  idx->node_deg.push_back(nullptr);
  idx->max_node_count = std::max<uint32_t>(idx->max_node_count, (uint32_t)count);
  *node_out = (int32_t)idx->h_nodes.size() - 1;
  return WS_OK;
}

int ws_index_set_wst(ws_index* idx, uint32_t rows, uint32_t split_factor, int32_t cutoff,
                     const uint32_t* row_nb, const uint64_t* offsets_flat,
                     const int32_t* node_ids_flat) {
  if (!idx || !row_nb || !offsets_flat) return ws_fail(WS_ERR_BADARG, "null argument");
  if (idx->finalized) return ws_fail(WS_ERR_STATE, "index already finalized");
  if (!idx->label_sorted) return ws_fail(WS_ERR_BADARG, "tree geometry needs a label-sorted arena");
  idx->wst_prefilter_nodes = node_ids_flat == nullptr;
  if (rows == 0 || split_factor < 2) return ws_fail(WS_ERR_BADARG, "rows=%u split_factor=%u", rows, split_factor);
  idx->wst_rows = rows;
  idx->split = split_factor;
  idx->cutoff = cutoff;
  idx->wst_nb.assign(row_nb, row_nb + rows);
  idx->wst_off_ptr.resize(rows);
  idx->wst_node_ptr.resize(rows);
  size_t no = 0, nn = 0;
  for (uint32_t r = 0; r < rows; r++) {
    idx->wst_off_ptr[r] = (uint32_t)no;
    idx->wst_node_ptr[r] = (uint32_t)nn;
    no += (size_t)row_nb[r] + 1;
    nn += row_nb[r];
  }
  idx->wst_off.assign(offsets_flat, offsets_flat + no);
  if (node_ids_flat) idx->wst_nodes.assign(node_ids_flat, node_ids_flat + nn);
  else idx->wst_nodes.assign(nn, -1);
  for (uint32_t r = 0; r < rows; r++) {
    const uint64_t* o = &idx->wst_off[idx->wst_off_ptr[r]];
    if (o[0] != 0 || o[row_nb[r]] != idx->n) return ws_fail(WS_ERR_BADARG, "row %u offsets do not span [0,n)", r);
    for (uint32_t b = 0; b < row_nb[r]; b++) {
      if (o[b + 1] <= o[b]) return ws_fail(WS_ERR_BADARG, "row %u bucket %u empty", r, b);
      if (idx->wst_prefilter_nodes) continue;
      int32_t h = idx->wst_nodes[idx->wst_node_ptr[r] + b];
      if (h < 0 || (size_t)h >= idx->h_nodes.size()) return ws_fail(WS_ERR_BADARG, "row %u bucket %u: bad node handle %d", r, b, h);
      if (idx->h_nodes[h].start != o[b] || idx->h_nodes[h].count != o[b + 1] - o[b])
        return ws_fail(WS_ERR_BADARG, "row %u bucket %u: node range mismatch", r, b);
    }
  }
  return WS_OK;
}

int ws_index_set_super(ws_index* idx, uint32_t rows, int32_t cutoff, const uint64_t* bucket_sizes,
                       const uint64_t* bucket_shifts, const uint32_t* row_nb,
                       const int32_t* node_ids_flat) {
  if (!idx || !bucket_sizes || !bucket_shifts || !row_nb || !node_ids_flat) return ws_fail(WS_ERR_BADARG, "null argument");
  if (idx->finalized) return ws_fail(WS_ERR_STATE, "index already finalized");
  if (!idx->label_sorted) return ws_fail(WS_ERR_BADARG, "tree geometry needs a label-sorted arena");
  if (rows == 0) return ws_fail(WS_ERR_BADARG, "rows=0");
  idx->sup_rows = rows;
  idx->sup_cutoff = cutoff;
  idx->sup_size.assign(bucket_sizes, bucket_sizes + rows);
  idx->sup_shift.assign(bucket_shifts, bucket_shifts + rows);
  idx->sup_nb.assign(row_nb, row_nb + rows);
  idx->sup_node_ptr.resize(rows);
  size_t nn = 0;
  for (uint32_t r = 0; r < rows; r++) {
    idx->sup_node_ptr[r] = (uint32_t)nn;
    nn += row_nb[r];
    if (r > 0 && bucket_shifts[r] == 0) return ws_fail(WS_ERR_BADARG, "row %u shift is zero", r);
  }
  idx->sup_nodes.assign(node_ids_flat, node_ids_flat + nn);
  for (size_t i = 0; i < nn; i++)
    if (idx->sup_nodes[i] < 0 || (size_t)idx->sup_nodes[i] >= idx->h_nodes.size())
      return ws_fail(WS_ERR_BADARG, "bad node handle %d", idx->sup_nodes[i]);
  return WS_OK;
}

}  // extern "C"

template <typename T>
static int ws_upload(ws_index* idx, const std::vector<T>& v, const T** out) {
  *out = nullptr;
  if (v.empty()) return WS_OK;
  void* p = nullptr;
  WS_CUDA(cudaMalloc(&p, v.size() * sizeof(T)));
  WS_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  idx->geom_dev_allocs.push_back(p);
  idx->hbm_bytes += v.size() * sizeof(T);
  *out = (const T*)p;
  return WS_OK;
}

extern "C" {

int ws_index_finalize(ws_index* idx) {
  if (!idx) return ws_fail(WS_ERR_BADARG, "null index");
  if (idx->finalized) return ws_fail(WS_ERR_STATE, "index already finalized");
  WsGeom& h = idx->hgeom;
  h.labels = idx->h_labels.data();
  h.n = idx->n;
  h.pf_n = idx->n + (idx->opt_open_tail ? 1 : 0);
  h.wst_rows = idx->wst_rows; h.split = idx->split; h.cutoff = idx->cutoff;
  h.wst_nb = idx->wst_nb.data(); h.wst_off_ptr = idx->wst_off_ptr.data(); h.wst_off = idx->wst_off.data();
  h.wst_node_ptr = idx->wst_node_ptr.data(); h.wst_nodes = idx->wst_nodes.data();
  h.wst_prefilter_nodes = idx->wst_prefilter_nodes ? 1u : 0u;
  h.sup_rows = idx->sup_rows;
  h.sup_size = idx->sup_size.data(); h.sup_shift = idx->sup_shift.data(); h.sup_nb = idx->sup_nb.data();
  h.sup_node_ptr = idx->sup_node_ptr.data(); h.sup_nodes = idx->sup_nodes.data();
  if (idx->device >= 0) {
    WS_CUDA(cudaSetDevice(idx->device));
    WsGeom& d = idx->dgeom;
    d = h;
    d.labels = idx->d_labels;
    WS_TRY(ws_upload(idx, idx->wst_nb, &d.wst_nb));
    WS_TRY(ws_upload(idx, idx->wst_off_ptr, &d.wst_off_ptr));
    WS_TRY(ws_upload(idx, idx->wst_off, &d.wst_off));
    WS_TRY(ws_upload(idx, idx->wst_node_ptr, &d.wst_node_ptr));
    WS_TRY(ws_upload(idx, idx->wst_nodes, &d.wst_nodes));
    WS_TRY(ws_upload(idx, idx->sup_size, &d.sup_size));
    WS_TRY(ws_upload(idx, idx->sup_shift, &d.sup_shift));
    WS_TRY(ws_upload(idx, idx->sup_nb, &d.sup_nb));
    WS_TRY(ws_upload(idx, idx->sup_node_ptr, &d.sup_node_ptr));
    WS_TRY(ws_upload(idx, idx->sup_nodes, &d.sup_nodes));
    const WsNode* dn = nullptr;
    WS_TRY(ws_upload(idx, idx->h_nodes, &dn));
    idx->d_nodes = const_cast<WsNode*>(dn);  // storage owned through geom_dev_allocs
  }
  idx->finalized = true;
  return WS_OK;
}

// ------------------------------------------------------------------------------------------
// batch orchestration
// ------------------------------------------------------------------------------------------
}  // extern "C"

static uint32_t ws_task_capacity(const ws_index* idx, int mode) {
  auto scan_tasks = [&](uint64_t rows) { return (uint32_t)(rows / (uint64_t)idx->opt_scan_chunk + 2); };
  if (mode == WS_MODE_POSTFILTER || mode == WS_METHOD_SUPER_POSTFILTER) return 1;
  if (mode == WS_MODE_PREFILTER) return scan_tasks(idx->n);
  // B-WST methods
  uint64_t split = idx->split, rows = idx->wst_rows;
  uint64_t graph = split * split + split + 2 + 2 * (split - 1) * (rows ? rows - 1 : 0);
  uint64_t last_bucket = 1;
  if (rows) last_bucket = idx->wst_off[idx->wst_off_ptr[rows - 1] + 1];
  uint64_t first_bucket_prev = rows >= 2 ? idx->wst_off[idx->wst_off_ptr[rows - 2] + 1] : idx->n;
  // an uncovered window is shorter than ~2 buckets of the row above the last one
  uint64_t scans = 2ull * scan_tasks(std::max<uint64_t>(2 * first_bucket_prev, 4 * last_bucket));
  uint64_t fen = graph + scans;
  // PrefilterIndex sub-indices: every bucket of the cover is scanned, in scan_chunk pieces
  if (idx->wst_prefilter_nodes) fen += scan_tasks(idx->n);
  if (mode == WS_METHOD_THREE_SPLIT) return (uint32_t)(3 * fen);
  return (uint32_t)fen;
}

static int ws_pick_kq(uint32_t dpad) {
  uint32_t need = (dpad / 4 + WS_TEAM - 1) / WS_TEAM;
  const int opts[] = {1, 2, 3, 4, 8, 16, 32};
  for (int o : opts)
    if ((uint32_t)o >= need) return o;
  return 32;
}

// ------------------------------------------------------------------------------------------
// tensor-core prefilter (ws_gemm.cuh): host orchestration
// ------------------------------------------------------------------------------------------
typedef CUresult (*ws_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is not linked (the library must load on machines without a driver): the encoder
// comes from the runtime's driver entry point table.
static int ws_tmap_encoder(ws_tmap_encode_fn* out) {
  static ws_tmap_encode_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    WS_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
    if (!p || qr != cudaDriverEntryPointSuccess) return ws_fail(WS_ERR_CUDA, "cuTensorMapEncodeTiled not available from this driver");
    fn = (ws_tmap_encode_fn)p;
  }
  *out = fn;
  return WS_OK;
}

// [rows][dpad] fp16 row-major -> boxes of 128 rows x 64 columns (128 B), 128-byte swizzle; columns past dpad and rows
// past `rows` read as zeros
static int ws_make_tmap(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t dpad) {
  ws_tmap_encode_fn enc;
  WS_TRY(ws_tmap_encoder(&enc));
  cuuint64_t gdim[2] = {dpad, rows};
  cuuint64_t gstride[1] = {(cuuint64_t)dpad * sizeof(uint16_t)};
  cuuint32_t box[2] = {WSG_KBLK, WSG_TILE_N};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return ws_fail(WS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for %llu x %u", (int)r, (unsigned long long)rows, dpad);
  return WS_OK;
}

static bool ws_gemm_eligible(const ws_index* idx, uint32_t k) {
  return idx->label_sorted && !idx->gemm_unfit && k <= WSG_KTOP && idx->dpad <= WSG_MAX_KB * WSG_KBLK &&
         idx->n >= 2 * WSG_TILE_N && idx->n < 0x7FFFFE00ull;
}

// one-time per index: |x|^2 table, max norm, the fp16 mirror of the arena and its tensor map.  Sets gemm_unfit
// (and returns WS_OK) when the arena cannot be mirrored; the caller then answers with the scan kernels.
static int ws_gemm_prepare(ws_index* idx) {
  if (idx->gemm_ready || idx->gemm_unfit) return WS_OK;
  const uint64_t npad = idx->n + 2 * WSG_TILE_N;
  WS_TRY(ws_ensure(idx, idx->g_norms, npad * sizeof(float)));
  WS_TRY(ws_ensure(idx, idx->g_ctrl, 64 * sizeof(unsigned long long)));
  WS_CUDA(cudaMemsetAsync(idx->g_ctrl.p, 0, 64 * sizeof(unsigned long long), idx->stream));
  WsGemmNormArgs na;
  na.vecs = idx->d_vecs; na.n = idx->n; na.dpad = idx->dpad; na.npad = (uint32_t)npad; na.metric = idx->metric;
  na.norms = (float*)idx->g_norms.p; na.max_sq = (uint32_t*)idx->g_ctrl.p + 8; na.max_abs = (uint32_t*)idx->g_ctrl.p + 9;
  WS_CUDA(wsg_launch_norm(idx->num_sms * 8, idx->stream, na));
  idx->launches++;
  uint32_t h_max[2] = {0, 0};  // bits of max |x|^2 and max |x_i|
  WS_CUDA(cudaMemcpyAsync(h_max, (uint32_t*)idx->g_ctrl.p + 8, sizeof(h_max), cudaMemcpyDeviceToHost, idx->stream));
  WS_CUDA(cudaStreamSynchronize(idx->stream));
  float max_sq, max_abs;
  std::memcpy(&max_sq, &h_max[0], 4);
  std::memcpy(&max_abs, &h_max[1], 4);
  int ex = 0;
  if (std::isfinite(max_abs) && max_abs > 0.f) std::frexp(max_abs, &ex);  // max_abs = m * 2^ex, m in [0.5, 1)
  // the largest component lands in [2^13, 2^14): far from fp16's overflow (2^16) and subnormals (2^-14)
  if (!std::isfinite(max_abs) || !std::isfinite(max_sq) || !(max_abs > 0.f) || ex < -60 || ex > 60) {
    idx->gemm_unfit = true;
    return WS_OK;
  }
  idx->g_x_exp = 14 - ex;
  WS_TRY(ws_ensure(idx, idx->g_vecs16, (size_t)idx->n * idx->dpad * sizeof(uint16_t)));
  WsGemmCvtArgs ca;
  ca.vecs = idx->d_vecs; ca.count = idx->n * idx->dpad; ca.scale = std::ldexp(1.0f, idx->g_x_exp); ca.out = (uint16_t*)idx->g_vecs16.p;
  WS_CUDA(wsg_launch_cvt(idx->num_sms * 8, idx->stream, ca));
  idx->launches++;
  WS_TRY(ws_make_tmap(&idx->g_tm_b, idx->g_vecs16.p, idx->n, idx->dpad));
  idx->gemm_ready = true;
  return WS_OK;
}

// PrefilterIndex::batch_search (prefiltering.h:124-146) for a whole batch on the tensor cores.
// dq/dw/dids/ddists are device pointers; work is only enqueued on the index stream.
static int ws_run_prefilter_gemm(ws_index* idx, const float* dq, const float* dw, uint64_t nq, uint32_t k, uint32_t* dids,
                                 float* ddists, const uint32_t* decode, uint32_t pad_id, uint32_t* overflow_flag) {
  cudaStream_t st = idx->stream;
  // 64-column fp16 blocks per tile: 1-4, 6 or 8 (a pipeline stage holds at most 4 of them; the columns past dpad are
  // zero-filled by the TMA unit on the point side and by the pack kernel on the query side)
  uint32_t nkb = (idx->dpad + WSG_KBLK - 1) / WSG_KBLK;
  if (nkb == 5) nkb = 6;
  if (nkb == 7) nkb = 8;
  const uint32_t kbps = nkb <= 4 ? nkb : nkb / 2;
  const uint32_t kcols = nkb * WSG_KBLK;
  const uint32_t acc_col0 = nkb <= 4 ? 128u : 256u;  // TMEM columns [0, acc_col0) hold the queries (32 per block)
  const uint32_t nacc = nkb <= 4 ? 3u : 2u;
  const uint32_t max_rows = (uint32_t)std::min<uint64_t>(WSG_MAX_ROWS, (nq + 127) / 128 * 128);
  const uint32_t max_groups = max_rows / 128;
  const uint32_t target = (uint32_t)(idx->opt_gemm_items > 0 ? idx->opt_gemm_items : 2 * (int64_t)idx->num_sms);
  // slices of the label axis swept at the same time by different query groups must stay in L2
  const uint32_t max_tiles = (uint32_t)std::max<uint64_t>(32, ((uint64_t)idx->opt_gemm_chunk_mb << 20) / ((uint64_t)idx->dpad * 2 * WSG_TILE_N));
  const uint64_t tiles_bound = (uint64_t)max_groups * ((idx->n + WSG_TILE_N - 1) / WSG_TILE_N);
  const uint32_t max_items = (uint32_t)(target + tiles_bound / max_tiles + 2 * max_groups + 8);
  WS_TRY(ws_ensure(idx, idx->g_qa, max_rows * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->g_qb, max_rows * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->g_perm, max_rows * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->g_row_a, max_rows * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->g_row_b, max_rows * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->g_items, max_items * sizeof(WsGemmItem)));
  WS_TRY(ws_ensure(idx, idx->g_group_items, (size_t)max_groups * WSG_MAX_SPLITS * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->g_group_cnt, max_groups * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->g_qpack, (size_t)max_rows * kcols * sizeof(uint16_t)));
  WS_TRY(ws_ensure(idx, idx->g_rscale, max_rows * sizeof(float)));
  WS_TRY(ws_ensure(idx, idx->g_slack, max_rows * sizeof(float)));
  WS_TRY(ws_ensure(idx, idx->g_thr0, max_rows * sizeof(float)));
  WS_TRY(ws_ensure(idx, idx->g_qnorm, max_rows * sizeof(float)));
  WS_TRY(ws_ensure(idx, idx->g_cand, (size_t)max_items * WSG_CAND_CAP * WSG_TILE_M * sizeof(uint64_t)));
  WS_TRY(ws_ensure(idx, idx->g_cand_cnt, (size_t)max_items * WSG_TILE_M * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->g_cand_thr, (size_t)max_items * WSG_TILE_M * sizeof(float)));
  WS_TRY(ws_ensure(idx, idx->g_res_keys, (size_t)max_rows * k * sizeof(uint64_t)));
  WS_TRY(ws_ensure(idx, idx->g_res_cnt, max_rows * sizeof(uint32_t)));
  const size_t gemm_smem = wsg_topk_smem_bytes();
  if (gemm_smem > idx->smem_optin) return ws_fail(WS_ERR_CUDA, "tensor-core prefilter needs %zu B of shared memory (> %zu)", gemm_smem, idx->smem_optin);
  if (!idx->gemm_attr_set) {  // function attributes are per device: once per arena, not once per process
    WS_CUDA(wsg_init_attributes());
    idx->gemm_attr_set = true;
  }
  unsigned long long* gctrl = (unsigned long long*)idx->g_ctrl.p;  // [0] survivors, [1] fallbacks, u32 view: [8] max |x|^2, [10] nitems
  uint32_t* gctrl32 = (uint32_t*)idx->g_ctrl.p;
  const int kq = ws_pick_kq(idx->dpad);
  const bool exact_rows = (uint32_t)kq * WS_TEAM * 4 == idx->dpad;

  for (uint64_t q0 = 0; q0 < nq; q0 += WSG_MAX_ROWS) {
    const uint32_t sn = (uint32_t)std::min<uint64_t>(WSG_MAX_ROWS, nq - q0);
    const uint32_t rows_pad = (sn + 127) / 128 * 128;
    uint32_t nsort = 128;
    while (nsort < rows_pad) nsort <<= 1;
    WsGemmPlanArgs pa;
    pa.windows = dw + 2 * q0; pa.labels = idx->d_labels; pa.n = idx->n; pa.n_bound = idx->hgeom.pf_n; pa.nq = sn; pa.rows_pad = rows_pad;
    pa.qa = (uint32_t*)idx->g_qa.p; pa.qb = (uint32_t*)idx->g_qb.p; pa.max_tiles = max_tiles;
    pa.perm = (uint32_t*)idx->g_perm.p; pa.row_a = (uint32_t*)idx->g_row_a.p; pa.row_b = (uint32_t*)idx->g_row_b.p;
    pa.items = (WsGemmItem*)idx->g_items.p; pa.nitems = gctrl32 + 10; pa.sched_ctr = gctrl32 + 12; pa.max_items = max_items;
    pa.group_items = (uint32_t*)idx->g_group_items.p; pa.group_cnt = (uint32_t*)idx->g_group_cnt.p;
    pa.target_items = target; pa.min_tiles = (uint32_t)std::max<int64_t>(1, idx->opt_gemm_min_tiles);
    pa.overflow = overflow_flag;
    WsGemmPackArgs ka;
    ka.queries = dq + q0 * idx->dim; ka.dim = idx->dim; ka.dpad = idx->dpad; ka.rows_pad = rows_pad; ka.metric = idx->metric;
    ka.perm = pa.perm; ka.max_sq = gctrl32 + 8; ka.qpack = (uint16_t*)idx->g_qpack.p; ka.slack = (float*)idx->g_slack.p;
    ka.kcols = kcols; ka.x_exp = idx->g_x_exp; ka.rscale = (float*)idx->g_rscale.p;
    ka.qnorm = (float*)idx->g_qnorm.p;
    WsGemmSeedArgs sa;
    sa.vecs = idx->d_vecs; sa.queries = ka.queries; sa.dim = idx->dim; sa.dpad = idx->dpad; sa.rows_pad = rows_pad; sa.k = k;
    sa.perm = pa.perm; sa.row_a = pa.row_a; sa.row_b = pa.row_b; sa.slack = ka.slack; sa.qnorm = ka.qnorm; sa.thr0 = (uint32_t*)idx->g_thr0.p;
    {
      WsKernelScope ks(idx, 10);
      WS_CUDA(wsg_launch_plan(nsort, st, pa));
      WS_CUDA(wsg_launch_pack(st, ka));
      idx->launches += 2;
    }
    {
      WsKernelScope ks(idx, 12);
      WS_CUDA(wsg_launch_seed(kq, idx->metric, exact_rows, st, sa));
    }
    WsGemmArgs ga;
    ga.items = pa.items; ga.nitems = pa.nitems; ga.sched_ctr = pa.sched_ctr; ga.dyn = idx->opt_gemm_dynamic ? 1u : 0u; ga.row_a = pa.row_a; ga.row_b = pa.row_b; ga.slack = ka.slack; ga.gthr = sa.thr0;
    ga.norms = (const float*)idx->g_norms.p; ga.cand = (uint64_t*)idx->g_cand.p; ga.cand_cnt = (uint32_t*)idx->g_cand_cnt.p;
    ga.cand_thr = (float*)idx->g_cand_thr.p; ga.nkb = nkb; ga.kbps = kbps; ga.nacc = nacc; ga.acc_col0 = acc_col0; ga.k = k;
    ga.dbg = (uint32_t)idx->opt_gemm_debug; ga.qpack = ka.qpack; ga.rscale = ka.rscale; ga.kcols = kcols;
    {
      WsKernelScope ks(idx, 9);
      WS_CUDA(wsg_launch_topk(idx->num_sms, st, idx->g_tm_b, ga));
    }
    WsGemmRerankArgs ra;
    ra.vecs = idx->d_vecs; ra.queries = ka.queries; ra.dim = idx->dim; ra.dpad = idx->dpad; ra.rows_pad = rows_pad;
    ra.perm = pa.perm; ra.row_a = pa.row_a; ra.row_b = pa.row_b; ra.group_items = pa.group_items; ra.group_cnt = pa.group_cnt;
    ra.cand = ga.cand; ra.cand_cnt = ga.cand_cnt; ra.cand_thr = ga.cand_thr; ra.k = k;
    ra.out_ids = dids + q0 * k; ra.out_dists = ddists + q0 * k; ra.decode = decode; ra.pad_id = pad_id;
    ra.res_keys = (uint64_t*)idx->g_res_keys.p; ra.res_cnt = (uint32_t*)idx->g_res_cnt.p; ra.stats = idx->d_stats; ra.gstats = gctrl;
    {
      WsKernelScope ks(idx, 11);
      WS_CUDA(wsg_launch_rerank(kq, idx->metric, exact_rows, st, ra));
    }
  }
  return WS_OK;
}

// ---- host buffers -> HBM -------------------------------------------------------------------------
// The caller's arrays are ordinary pageable memory (numpy).  cudaMemcpyAsync from pageable memory is staged by
// the driver through one thread: 0.25 ms of a 0.43 ms call for 10 000 x 128 queries, where the DMA itself needs
// 0.1 ms (profiles/scripts/e2e_overhead_probe.py).  Here a few helper threads copy 512 KB chunks into a pinned
// staging buffer of the arena while the calling thread hands each finished chunk to the copy engine.
struct WsCopyJob {
  char* dst = nullptr;
  const char* src = nullptr;
  size_t bytes = 0, chunk = 0;
  uint32_t nchunks = 0;
  std::atomic<uint32_t> next{0};
  std::atomic<uint32_t>* done = nullptr;  // one flag per chunk
  std::atomic<int> active{0};             // helpers currently inside the job
  std::atomic<uint64_t> gen{0};           // the slot is reused: which job it holds
};

// Helper threads that spin for a short while after a job (batches of a sweep follow each other within
// microseconds) and sleep otherwise.  One job slot, owned by the pool, so that a helper that wakes up late only ever
// touches memory that is still there; the calling thread copies chunks too, so a sleeping pool costs nothing but
// the wake-up call.
class WsCopyPool {
 public:
  static WsCopyPool& get() {
    // leaked on purpose: helpers wait on its condition variable for the life of the process, and destroying a
    // condition variable with waiters blocks (pthread_cond_destroy) — a static instance hangs the process at exit
    static WsCopyPool* pool = new WsCopyPool();
    return *pool;
  }
  bool enabled() const { return nthreads_ > 0; }
  // nullptr: the pool is busy with another arena's batch (the caller copies alone)
  WsCopyJob* begin(char* dst, const char* src, size_t bytes, size_t chunk, std::atomic<uint32_t>* done) {
    if (nthreads_ == 0 || !busy_.try_lock()) return nullptr;
    WsCopyJob* j = &slot_;
    while (j->active.load(std::memory_order_acquire) != 0) cpu_relax();  // a straggler of the previous job
    j->dst = dst; j->src = src; j->bytes = bytes; j->chunk = chunk;
    j->nchunks = (uint32_t)((bytes + chunk - 1) / chunk);
    j->done = done;
    j->next.store(0, std::memory_order_relaxed);
    j->gen.fetch_add(1, std::memory_order_relaxed);
    cur_.store(j, std::memory_order_release);
    if (sleepers_.load(std::memory_order_acquire) > 0) {
      std::lock_guard<std::mutex> lk(mu_);
      cv_.notify_all();
    }
    return j;
  }
  void end(WsCopyJob* j) {  // every chunk has been seen done by the caller
    cur_.store(nullptr, std::memory_order_release);
    while (j->active.load(std::memory_order_acquire) != 0) cpu_relax();
    busy_.unlock();
  }
  // one chunk of the job on the calling thread; false when none is left
  static bool help(WsCopyJob* j) {
    const uint32_t i = j->next.fetch_add(1, std::memory_order_relaxed);
    if (i >= j->nchunks) return false;
    const size_t off = (size_t)i * j->chunk;
    std::memcpy(j->dst + off, j->src + off, std::min(j->chunk, j->bytes - off));
    j->done[i].store(1, std::memory_order_release);
    return true;
  }
  static void cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }

 private:
  WsCopyPool() {
    // Measured on the 16-thread B200 host: 10 000 x 128 queries per call, 0.42 ms (driver staging) -> 0.33 ms with 4
    // helpers.  Helpers spin while batches follow each other, so a box that runs one process per GPU must not be
    // oversubscribed: at most half of the hardware threads, shared between the visible GPUs.
    int n = 4;
    {
      const unsigned hw = std::thread::hardware_concurrency();
      int ngpu = 0;
      if (cudaGetDeviceCount(&ngpu) != cudaSuccess) { cudaGetLastError(); ngpu = 1; }
      if (hw > 0) n = std::max(1, std::min(4, (int)(hw / (2u * (unsigned)std::max(1, ngpu)))));
    }
    if (const char* e = std::getenv("WSANN_COPY_THREADS")) n = std::atoi(e);
    nthreads_ = std::max(0, std::min(n, 16));
    for (int i = 0; i < nthreads_; i++) std::thread([this] { loop(); }).detach();  // process-lifetime helpers
  }
  void loop() {
    auto last = std::chrono::steady_clock::now();
    for (;;) {
      WsCopyJob* j = cur_.load(std::memory_order_acquire);
      if (j != nullptr) {
        j->active.fetch_add(1, std::memory_order_acq_rel);
        uint64_t mine = 0;
        if (cur_.load(std::memory_order_acquire) == j) {
          mine = j->gen.load(std::memory_order_relaxed);
          while (help(j)) {}
          last = std::chrono::steady_clock::now();
        }
        j->active.fetch_sub(1, std::memory_order_acq_rel);
        // the job stays published until the caller retires it: do not hammer its counter meanwhile
        while (mine != 0 && cur_.load(std::memory_order_acquire) == j && j->gen.load(std::memory_order_relaxed) == mine) cpu_relax();
        continue;
      }
      if (std::chrono::steady_clock::now() - last < std::chrono::microseconds(400)) {
        cpu_relax();
        continue;
      }
      std::unique_lock<std::mutex> lk(mu_);
      sleepers_.fetch_add(1, std::memory_order_acq_rel);
      cv_.wait_for(lk, std::chrono::milliseconds(200), [&] { return cur_.load(std::memory_order_acquire) != nullptr; });
      sleepers_.fetch_sub(1, std::memory_order_acq_rel);
      last = std::chrono::steady_clock::now();
    }
  }
  int nthreads_ = 0;
  WsCopyJob slot_;
  std::atomic<WsCopyJob*> cur_{nullptr};
  std::atomic<int> sleepers_{0};
  std::mutex busy_, mu_;
  std::condition_variable cv_;
};

static const size_t kStageChunk = 256u << 10;

static int ws_ensure_pinned(ws_index* idx, WsDevBuf& b, size_t bytes) {
  if (b.bytes >= bytes) return WS_OK;
  if (b.p) {
    WS_CUDA(cudaStreamSynchronize(idx->stream));
    WS_CUDA(cudaFreeHost(b.p));
    b.p = nullptr; b.bytes = 0;
  }
  const size_t want = bytes + bytes / 4 + 4096;
  WS_CUDA(cudaMallocHost(&b.p, want));
  b.bytes = want;
  return WS_OK;
}

// dst (device) <- src (pageable host), enqueued on the index stream; small copies go straight to the driver
static int ws_stage_h2d(ws_index* idx, void* dst, const void* src, size_t bytes) {
  cudaStream_t st = idx->stream;
  WsCopyPool& pool = WsCopyPool::get();
  bool direct = bytes < 4 * kStageChunk || idx->opt_stage_copies == 0 || !pool.enabled();
  if (!direct) {  // a buffer the caller pinned itself goes to the copy engine as it is
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, src) == cudaSuccess) direct = attr.type != cudaMemoryTypeUnregistered;
    else cudaGetLastError();
  }
  if (direct) {
    WS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return WS_OK;
  }
  WS_TRY(ws_ensure_pinned(idx, idx->stage_in, bytes));
  const uint32_t nchunks = (uint32_t)((bytes + kStageChunk - 1) / kStageChunk);
  if (idx->stage_flags.size() < nchunks) idx->stage_flags = std::vector<std::atomic<uint32_t>>(nchunks);
  for (uint32_t i = 0; i < nchunks; i++) idx->stage_flags[i].store(0, std::memory_order_relaxed);
  char* stage = (char*)idx->stage_in.p;
  WsCopyJob* job = pool.begin(stage, (const char*)src, bytes, kStageChunk, idx->stage_flags.data());
  if (job == nullptr) {  // pool busy with another arena (the members of a group run in parallel anyway)
    WS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return WS_OK;
  }
  // copy chunks alongside the helpers; every run of finished chunks goes to the copy engine as one transfer
  int rc = WS_OK;
  uint32_t issued = 0;
  auto flush = [&](bool wait_all) {
    for (;;) {
      uint32_t upto = issued;
      while (upto < nchunks && idx->stage_flags[upto].load(std::memory_order_acquire) != 0) upto++;
      if (upto > issued) {
        const size_t off = (size_t)issued * kStageChunk, len = std::min(bytes, (size_t)upto * kStageChunk) - off;
        if (rc == WS_OK && cudaMemcpyAsync((char*)dst + off, stage + off, len, cudaMemcpyHostToDevice, st) != cudaSuccess)
          rc = ws_fail(WS_ERR_CUDA, "staged host-to-device copy: %s", cudaGetErrorString(cudaGetLastError()));
        issued = upto;
      }
      if (!wait_all || issued == nchunks) return;
      WsCopyPool::cpu_relax();
    }
  };
  while (WsCopyPool::help(job)) flush(false);
  flush(true);
  pool.end(job);
  return rc;
}

// host-only exercise of the helper pool (no CUDA): `reps` jobs of `bytes` bytes, results compared (CPU tests)
extern "C" int ws_debug_copy_pool_selftest(uint64_t bytes, uint32_t reps) {
  WsCopyPool& pool = WsCopyPool::get();
  std::vector<char> src(bytes), dst(bytes);
  for (uint64_t i = 0; i < bytes; i++) src[i] = (char)(i * 131u + (i >> 9));
  const uint32_t nchunks = (uint32_t)((bytes + kStageChunk - 1) / kStageChunk);
  std::vector<std::atomic<uint32_t>> flags(std::max(1u, nchunks));
  double copy_us = 0;
  for (uint32_t r = 0; r < reps; r++) {
    std::fill(dst.begin(), dst.end(), 0);
    for (auto& f : flags) f.store(0);
    src[r % std::max<uint64_t>(1, bytes)] ^= 0x5a;
    const auto t0 = std::chrono::steady_clock::now();
    WsCopyJob* job = pool.begin(dst.data(), src.data(), bytes, kStageChunk, flags.data());
    if (job == nullptr) {
      if (pool.enabled()) return ws_fail(WS_ERR_STATE, "copy pool busy");
      std::memcpy(dst.data(), src.data(), bytes);
    } else {
      while (WsCopyPool::help(job)) {}
      for (uint32_t i = 0; i < nchunks; i++)
        while (flags[i].load(std::memory_order_acquire) == 0) WsCopyPool::cpu_relax();
      pool.end(job);
    }
    copy_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    if (std::memcmp(dst.data(), src.data(), bytes) != 0) return ws_fail(WS_ERR_STATE, "copy pool: job %u copied wrong bytes", r);
    if ((r & 7) == 7) std::this_thread::sleep_for(std::chrono::milliseconds(1));  // let the helpers fall asleep now and then
  }
  if (std::getenv("WSANN_DEBUG_POOL")) std::fprintf(stderr, "[wsann] copy pool: %.1f us per %llu-byte job\n", copy_us / std::max(1u, reps), (unsigned long long)bytes);
  return WS_OK;
}

// turns the device-side sticky error word into a status (and clears it)
static int ws_report_sticky(ws_index* idx, uint32_t bits) {
  cudaMemsetAsync(idx->d_sticky, 0, sizeof(uint32_t), idx->stream);
  cudaStreamSynchronize(idx->stream);
  if (bits & 1u) return ws_fail(WS_ERR_STATE, "task slot capacity overflowed: tasks were dropped (internal bound too small)");
  return ws_fail(WS_ERR_STATE, "a graph task outgrew the last beam tier and was dropped (sticky bits 0x%x)", bits);
}

struct WsBatchPlan {
  int mode;
  int32_t node;
  uint32_t k;
  ws_query_params qp;
  uint32_t pad_id;
  bool use_decode;
};

static int ws_run_batch(ws_index* idx, const WsBatchPlan& plan, const float* queries, const float* windows,
                        uint64_t nq, uint32_t* ids, float* dists, uint32_t flags) {
  if (!idx) return ws_fail(WS_ERR_BADARG, "null index");
  std::lock_guard<std::recursive_mutex> lock(idx->mu);
  if (idx->device < 0) return ws_fail(WS_ERR_CUDA, "host-only geometry index: no CUDA device, and this engine has no CPU fallback");
  if (!idx->finalized) return ws_fail(WS_ERR_STATE, "ws_index_finalize has not been called");
  if (nq == 0) return WS_OK;
  if (!queries || !windows || !ids || !dists) return ws_fail(WS_ERR_BADARG, "null buffer");
  if (nq > (1ull << 24)) return ws_fail(WS_ERR_BADARG, "nq=%llu above 2^24 per batch", (unsigned long long)nq);
  const uint32_t k = plan.k;
  if (k == 0 || k > kMaxK) return ws_fail(WS_ERR_BADARG, "k=%u outside 1..%u", k, kMaxK);
  const bool needs_graph = plan.mode != WS_MODE_PREFILTER && !(plan.mode <= 2 && idx->wst_prefilter_nodes);
  const ws_query_params& qp = plan.qp;
  if (needs_graph) {
    if (idx->h_nodes.empty()) return ws_fail(WS_ERR_STATE, "index has no graphs");
    if (qp.beam_size < 1) return ws_fail(WS_ERR_BADARG, "beam_size=%lld", (long long)qp.beam_size);
    if (qp.postfiltering_max_beam > (int64_t)kBeamCapLarge)
      return ws_fail(WS_ERR_BADARG, "postfiltering_max_beam=%lld above the %u this build keeps in shared memory",
                     (long long)qp.postfiltering_max_beam, kBeamCapLarge);
    if (qp.final_beam_multiply < 1) return ws_fail(WS_ERR_BADARG, "final_beam_multiply=%lld", (long long)qp.final_beam_multiply);
  }
  if ((plan.mode == WS_MODE_PREFILTER || plan.mode <= 3) && !idx->label_sorted)
    return ws_fail(WS_ERR_STATE, "this query needs a label-sorted arena");
  if (plan.mode <= 2 && idx->wst_rows == 0) return ws_fail(WS_ERR_STATE, "no B-WST geometry set");
  if (plan.mode == 3 && idx->sup_rows == 0) return ws_fail(WS_ERR_STATE, "no super-postfilter geometry set");
  if (plan.mode == WS_MODE_POSTFILTER && (plan.node < 0 || (size_t)plan.node >= idx->h_nodes.size()))
    return ws_fail(WS_ERR_BADARG, "bad node handle %d", plan.node);

  WS_CUDA(cudaSetDevice(idx->device));
  cudaStream_t st = idx->stream;
  const uint32_t cap = ws_task_capacity(idx, plan.mode);
  const size_t slots = (size_t)nq * cap;
  if (slots >= (1ull << 32)) return ws_fail(WS_ERR_BADARG, "batch too large: %llu task slots", (unsigned long long)slots);

  // ---- scratch
  WS_TRY(ws_ensure(idx, idx->tasks, slots * sizeof(WsTask)));
  WS_TRY(ws_ensure(idx, idx->res_keys, slots * k * sizeof(uint64_t)));
  WS_TRY(ws_ensure(idx, idx->res_cnt, slots * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->counts, nq * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->queues, (WS_NUM_TIERS + 1) * slots * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->ctrl, 64 * sizeof(uint32_t)));
  const bool dev_ptrs = (flags & WS_FLAG_DEVICE_PTRS) != 0;
  const float* dq = queries;
  const float* dw = windows;
  uint32_t* dids = ids;
  float* ddists = dists;
  if (!dev_ptrs) {
    WS_TRY(ws_ensure(idx, idx->d_queries, nq * idx->dim * sizeof(float)));
    WS_TRY(ws_ensure(idx, idx->d_windows, nq * 2 * sizeof(float)));
    WS_TRY(ws_ensure(idx, idx->d_ids, nq * k * sizeof(uint32_t)));
    WS_TRY(ws_ensure(idx, idx->d_dists, nq * k * sizeof(float)));
    WS_CUDA(cudaMemcpyAsync(idx->d_windows.p, windows, nq * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    WS_TRY(ws_stage_h2d(idx, idx->d_queries.p, queries, nq * idx->dim * sizeof(float)));
    dq = (const float*)idx->d_queries.p;
    dw = (const float*)idx->d_windows.p;
    dids = (uint32_t*)idx->d_ids.p;
    ddists = (float*)idx->d_dists.p;
  }
  // ---- prefilter batches: how the batch is answered depends on the window sizes.  With host buffers a
  // sample of the batch's windows is measured here; device-pointer calls take the options as they are.
  double mean_window = -1.0;  // rows per window over a sample of 64 queries (host buffers only)
  const float* host_windows = dev_ptrs ? idx->hint_windows : windows;
  if (plan.mode == WS_MODE_PREFILTER && host_windows && nq >= 256) {
    const float* windows = host_windows;
    const std::vector<float>& L = idx->h_labels;
    const uint64_t step = std::max<uint64_t>(1, nq / 64);
    double sum = 0; uint64_t cnt = 0;
    for (uint64_t i = 0; i < nq; i += step, cnt++) {
      const uint64_t a = std::lower_bound(L.begin(), L.end(), windows[2 * i]) - L.begin();
      const uint64_t b = std::lower_bound(L.begin(), L.end(), windows[2 * i + 1]) - L.begin();
      sum += b > a ? (double)(b - a) : 0.0;
    }
    if (cnt > 0) mean_window = sum / (double)cnt;
  }
  // large windows: dense query x slice contraction on the tensor cores (K1g)
  bool use_gemm = false;
  if (plan.mode == WS_MODE_PREFILTER && idx->opt_gemm != 0 && ws_gemm_eligible(idx, k)) {
    if (idx->opt_gemm == 1) use_gemm = true;
    else use_gemm = mean_window >= (double)idx->opt_gemm_min_window;
    if (use_gemm) {
      WS_TRY(ws_gemm_prepare(idx));  // first use: norm table + fp16 mirror of the arena
      use_gemm = !idx->gemm_unfit;
    }
  }
  // small windows: the whole batch in one launch (K1d)
  bool use_direct = false;
  if (plan.mode == WS_MODE_PREFILTER && !use_gemm && k <= 128 && idx->opt_warp_scan && idx->opt_direct != 0) {
    if (idx->opt_direct == 1) use_direct = true;
    else use_direct = mean_window >= 0.0 && mean_window <= (double)idx->opt_scan_chunk;
  }
  // ctrl layout: [0..6] queue counts (beam tiers 0..5, scan = 6), [8..14] queue heads
  uint32_t* ctrl = (uint32_t*)idx->ctrl.p;
  if (!use_direct) WS_CUDA(cudaMemsetAsync(ctrl, 0, 64 * sizeof(uint32_t), st));
  uint32_t* queues = (uint32_t*)idx->queues.p;
  const int kq = ws_pick_kq(idx->dpad);
  const int metric = idx->metric;
  const bool exact_rows = (uint32_t)kq * WS_TEAM * 4 == idx->dpad;  // padded row = 8*KQ float4s: no column predicate
#define WS_LAUNCH(what, expr)                                                                               \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) return ws_fail(WS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(_e));             \
  } while (0)

  if (use_gemm) {
    WS_TRY(ws_run_prefilter_gemm(idx, dq, dw, nq, k, dids, ddists, plan.use_decode ? idx->d_decode : nullptr, plan.pad_id, idx->d_sticky));
  } else if (use_direct) {
    WsPrefilterDirectArgs pa;
    pa.s.vecs = idx->d_vecs; pa.s.queries = dq; pa.s.dim = idx->dim; pa.s.dpad = idx->dpad;
    pa.s.tasks = nullptr; pa.s.res_keys = (uint64_t*)idx->res_keys.p; pa.s.res_cnt = (uint32_t*)idx->res_cnt.p;
    pa.s.k = k; pa.s.q_in = nullptr; pa.s.q_in_count = nullptr; pa.s.q_head = nullptr; pa.s.stats = idx->d_stats;
    pa.s.out_ids = dids; pa.s.out_dists = ddists; pa.s.decode = plan.use_decode ? idx->d_decode : nullptr; pa.s.pad_id = plan.pad_id;
    pa.labels = idx->d_labels; pa.n = idx->hgeom.pf_n; pa.windows = dw; pa.nq = (uint32_t)nq;
    int occ = 0;
    WS_LAUNCH("occupancy query", wsl_prefilter_direct_occ(kq, metric, exact_rows, &occ));
    if (occ < 1) return ws_fail(WS_ERR_CUDA, "prefilter kernel does not fit on an SM");
    const int grid = (int)std::min<uint64_t>((nq + WS_WARPS_PER_CTA - 1) / WS_WARPS_PER_CTA, (uint64_t)idx->num_sms * occ);
    {
      WsKernelScope ks(idx, 7);
      WS_LAUNCH("prefilter kernel launch", wsl_prefilter_direct(kq, metric, exact_rows, grid, st, pa));
    }
  } else {
  // ---- tiers: which launch takes fresh graph tasks
  const int lowest_tier = idx->opt_warp_tiers ? 0 : WS_NUM_WARP_TIERS;
  int first_tier = WS_NUM_TIERS - 1;
  if (needs_graph) {
    for (int t = lowest_tier; t < WS_NUM_TIERS - 1; t++)
      if ((uint64_t)qp.beam_size <= kBeamTierCaps[t]) { first_tier = t; break; }
  }

  // ---- K3 decomposition
  WsDecompArgs da;
  da.g = idx->dgeom;
  da.p.beam = needs_graph ? (uint32_t)std::min<int64_t>(qp.beam_size, 0x7fffffff) : 0;
  da.p.has_ratio = qp.has_min_query_to_bucket_ratio;
  da.p.min_ratio = qp.min_query_to_bucket_ratio;
  da.p.scan_chunk = (uint32_t)idx->opt_scan_chunk;
  da.mode = plan.mode;
  da.node = plan.node;
  da.windows = dw;
  da.nq = (uint32_t)nq;
  da.cap = cap;
  da.tasks = (WsTask*)idx->tasks.p;
  da.counts = (uint32_t*)idx->counts.p;
  da.gq = queues + (size_t)first_tier * slots;
  da.gq_count = ctrl + first_tier;
  da.sq = queues + (size_t)WS_NUM_TIERS * slots;
  da.sq_count = ctrl + WS_NUM_TIERS;
  da.overflow = idx->d_sticky;
  da.stats = idx->d_stats;
  {
    WsKernelScope ks(idx, 0);
    WS_LAUNCH("decomposition kernel launch", wsl_decompose((int)((nq + 127) / 128), st, da));
  }

  // the first warp-tier launch can also take the batch's scan tasks (same warps, same smem)
  const bool has_scans = plan.mode != WS_MODE_POSTFILTER && plan.mode != WS_METHOD_SUPER_POSTFILTER;
  const bool fuse_scan = needs_graph && has_scans && idx->opt_fuse_scan && first_tier < 2 &&
                         k <= std::min<uint32_t>(128u, kBeamTierCaps[first_tier]);

  // ---- K2 beam search, one persistent launch per tier
  if (needs_graph) {
    const uint32_t E = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(8, idx->opt_expand));
    uint32_t cand_cap = 64;
    while (cand_cap < E * idx->R) cand_cap <<= 1;
    // active tiers, in order; a task that outgrows tier i re-queues itself for tier i+1 of this list
    int tiers[WS_NUM_TIERS];
    int ntiers = 0;
    for (int t = first_tier; t < WS_NUM_TIERS; t++) {
      if (ntiers > 0 && (int64_t)kBeamTierCaps[tiers[ntiers - 1]] >= qp.postfiltering_max_beam) break;  // nothing can need more
      tiers[ntiers++] = t;
    }
    for (int ti = 0; ti < ntiers; ti++) {
      const int t = tiers[ti];
      const int t_next = ti + 1 < ntiers ? tiers[ti + 1] : -1;
      const bool large = (t == WS_NUM_TIERS - 1);
      const bool warp_tier = !large;
      const uint32_t beam_cap = large ? kBeamCapLarge : kBeamTierCaps[t];
      const int cs = large ? 14 : kBeamTierCS[t];
      // Visited table of the warp tiers: 2048 entries up to beam 128, then 16 per unit of beam capacity, 8 at 1024
      // (a beam-B search sees ~7 B distinct nodes; the table only has to avoid most recomputation).
      // 16-bit tags need id < 2^(log2(entries) + 12); larger nodes use 32-bit entries.
      uint32_t hash_entries = 0, hash16 = 0;
      if (warp_tier) {
        hash_entries = (uint32_t)idx->opt_warp_hash;
        if (t == 2) hash_entries *= 2;
        if (t >= 3) hash_entries *= 4;
        uint32_t hb = 0;
        while ((1u << hb) < hash_entries) hb++;
        hash16 = (idx->opt_hash16 && (uint64_t)idx->max_node_count <= (1ull << (hb + 12))) ? 1u : 0u;
      }
      if (idx->R > 64) return ws_fail(WS_ERR_BADARG, "graphs with max_degree > 64 are not supported by the query kernels");
      size_t smem;
      if (warp_tier)
        smem = (size_t)WS_WARPS_PER_CTA * ws_warp_smem_bytes(beam_cap, hash_entries, hash16);
      else
        smem = (size_t)beam_cap * 8 + 64 * 8 * 2 + 64 * 4 * 2;
      if (smem > idx->smem_optin) return ws_fail(WS_ERR_BADARG, "beam tier %d needs %zu B of shared memory (> %zu)", t, smem, idx->smem_optin);
      int occ = 0;
      WS_LAUNCH("occupancy query", warp_tier ? wsl_beam_warp_occ(kq, metric, exact_rows, cs, smem, &occ)
                                            : wsl_beam_cta_occ(kq, metric, exact_rows, smem, &occ));
      if (occ < 1) return ws_fail(WS_ERR_CUDA, "beam kernel does not fit on an SM (smem %zu)", smem);
      int grid = idx->num_sms * occ;
      WsBeamArgs ba;
      ba.vecs = idx->d_vecs; ba.labels = idx->d_labels; ba.nodes = idx->d_nodes; ba.queries = dq;
      ba.dim = idx->dim; ba.dpad = idx->dpad; ba.R = idx->R;
      ba.tasks = (WsTask*)idx->tasks.p; ba.res_keys = (uint64_t*)idx->res_keys.p; ba.res_cnt = (uint32_t*)idx->res_cnt.p;
      ba.k = k;
      ba.q_in = queues + (size_t)t * slots; ba.q_in_count = ctrl + t; ba.q_head = ctrl + 8 + t;
      ba.q_out = t_next >= 0 ? queues + (size_t)t_next * slots : nullptr;
      ba.q_out_count = t_next >= 0 ? ctrl + t_next : nullptr;
      ba.beam_cap = beam_cap; ba.hash_mask = hash_entries ? hash_entries - 1 : 0; ba.cand_cap = cand_cap;
      ba.expand = E; ba.skip_query_id = (int32_t)idx->opt_skip_query_id; ba.query_id_base = idx->query_id_base;
      ba.max_beam = qp.postfiltering_max_beam; ba.final_mult = qp.final_beam_multiply;
      ba.limit = qp.limit > 0 ? qp.limit : (1ll << 62);
      ba.degree_limit = qp.degree_limit > 0 ? qp.degree_limit : (1ll << 62);
      ba.bitmap = nullptr; ba.bitmap_words = 0;
      if (large) {
        ba.bitmap_words = ((uint64_t)idx->max_node_count + 31) / 32;
        ba.bitmap_words = (ba.bitmap_words + 31) & ~31ull;
        WS_TRY(ws_ensure(idx, idx->bitmap, (size_t)grid * ba.bitmap_words * sizeof(uint32_t)));
        ba.bitmap = (uint32_t*)idx->bitmap.p;
      }
      ba.stats = idx->d_stats;
      ba.sticky = idx->d_sticky;
      ba.hash16 = hash16;
      ba.min_tasks = 0u;
      ba.sq_in = nullptr; ba.sq_count = nullptr; ba.sq_head = nullptr;
      if (fuse_scan && t == first_tier) {  // this launch also drains the scan queue
        ba.sq_in = queues + (size_t)WS_NUM_TIERS * slots; ba.sq_count = ctrl + WS_NUM_TIERS; ba.sq_head = ctrl + 8 + WS_NUM_TIERS;
      }
      ba.out_ids = dids; ba.out_dists = ddists; ba.decode = plan.use_decode ? idx->d_decode : nullptr; ba.pad_id = plan.pad_id;
      {
        WsKernelScope ks(idx, 1 + t);
        WS_LAUNCH("beam kernel launch", warp_tier ? wsl_beam_warp(kq, metric, exact_rows, cs, grid, smem, st, ba)
                                                  : wsl_beam_cta(kq, metric, exact_rows, grid, smem, st, ba));
      }
    }
  }

  // ---- K1 scans
  if (has_scans && !fuse_scan) {
    WsScanArgs sa;
    sa.vecs = idx->d_vecs; sa.queries = dq; sa.dim = idx->dim; sa.dpad = idx->dpad;
    sa.tasks = (const WsTask*)idx->tasks.p; sa.res_keys = (uint64_t*)idx->res_keys.p; sa.res_cnt = (uint32_t*)idx->res_cnt.p;
    sa.k = k;
    sa.q_in = queues + (size_t)WS_NUM_TIERS * slots; sa.q_in_count = ctrl + WS_NUM_TIERS; sa.q_head = ctrl + 8 + WS_NUM_TIERS;
    sa.stats = idx->d_stats;
    sa.out_ids = dids; sa.out_dists = ddists; sa.decode = plan.use_decode ? idx->d_decode : nullptr; sa.pad_id = plan.pad_id;
    if (k <= 128 && idx->opt_warp_scan) {  // warp-per-task streaming scan
      int occ = 0;
      WS_LAUNCH("occupancy query", wsl_scan_warp_occ(kq, metric, exact_rows, &occ));
      if (occ < 1) return ws_fail(WS_ERR_CUDA, "scan kernel does not fit on an SM");
      int grid = idx->num_sms * occ;
      WsKernelScope ks(idx, 7);
      WS_LAUNCH("scan kernel launch", wsl_scan_warp(kq, metric, exact_rows, grid, st, sa));
    } else {
      size_t smem = (size_t)WS_TOPK_BUF * 8 + (size_t)idx->dpad * 4;
      int grid = idx->num_sms * 8;
      WsKernelScope ks(idx, 7);
      WS_LAUNCH("scan kernel launch", wsl_scan(kq, metric, grid, smem, st, sa));
    }
  }

  // ---- K4 merge + decode
  {
    WsMergeArgs ma;
    ma.counts = (const uint32_t*)idx->counts.p; ma.cap = cap;
    ma.res_keys = (const uint64_t*)idx->res_keys.p; ma.res_cnt = (const uint32_t*)idx->res_cnt.p;
    ma.k = k; ma.decode = plan.use_decode ? idx->d_decode : nullptr; ma.pad_id = plan.pad_id;
    ma.nq = (uint32_t)nq; ma.ids = dids; ma.dists = ddists;
    int grid = (int)std::min<uint64_t>(nq, (uint64_t)idx->num_sms * 16);
    WsKernelScope ks(idx, 8);
    WS_LAUNCH("merge kernel launch", wsl_merge(grid, st, ma));
  }
  }  // !use_gemm

  if (!dev_ptrs) {
    uint32_t h_sticky = 0;
    const size_t rows_bytes = nq * k * sizeof(uint32_t);
    bool staged_out = rows_bytes >= (128u << 10) && idx->opt_stage_copies != 0;
    if (staged_out) {
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, ids) == cudaSuccess) staged_out = attr.type == cudaMemoryTypeUnregistered;
      else cudaGetLastError();
    }
    if (staged_out) {
      // results through a pinned buffer: the copy engine writes it at link speed, one memcpy hands it to the caller
      WS_TRY(ws_ensure_pinned(idx, idx->stage_out, 2 * rows_bytes + 64));
      char* so = (char*)idx->stage_out.p;
      WS_CUDA(cudaMemcpyAsync(so, dids, rows_bytes, cudaMemcpyDeviceToHost, st));
      WS_CUDA(cudaMemcpyAsync(so + rows_bytes, ddists, rows_bytes, cudaMemcpyDeviceToHost, st));
      WS_CUDA(cudaMemcpyAsync(so + 2 * rows_bytes, idx->d_sticky, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      WS_CUDA(cudaStreamSynchronize(st));
      std::memcpy(ids, so, rows_bytes);
      std::memcpy(dists, so + rows_bytes, rows_bytes);
      std::memcpy(&h_sticky, so + 2 * rows_bytes, sizeof(uint32_t));
    } else {
      WS_CUDA(cudaMemcpyAsync(ids, dids, rows_bytes, cudaMemcpyDeviceToHost, st));
      WS_CUDA(cudaMemcpyAsync(dists, ddists, rows_bytes, cudaMemcpyDeviceToHost, st));
      WS_CUDA(cudaMemcpyAsync(&h_sticky, idx->d_sticky, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      WS_CUDA(cudaStreamSynchronize(st));
    }
    if (h_sticky) return ws_report_sticky(idx, h_sticky);
  }
  return WS_OK;
}

extern "C" {

int ws_prefilter_batch(ws_index* idx, const float* queries, const float* windows, uint64_t nq,
                       uint32_t k, uint32_t* ids, float* dists, uint32_t flags) {
  WsBatchPlan p{};
  p.mode = WS_MODE_PREFILTER;
  p.node = -1;
  p.k = k;
  p.pad_id = 0xFFFFFFFFu;
  p.use_decode = idx && idx->has_decode;
  return ws_run_batch(idx, p, queries, windows, nq, ids, dists, flags);
}

int ws_postfilter_batch(ws_index* idx, int32_t node, const float* queries, const float* windows,
                        uint64_t nq, const ws_query_params* qp, int pad, uint32_t* ids,
                        float* dists, uint32_t flags) {
  if (!qp) return ws_fail(WS_ERR_BADARG, "null query params");
  if (qp->k < 1 || qp->k > kMaxK) return ws_fail(WS_ERR_BADARG, "k=%lld outside 1..%u", (long long)qp->k, kMaxK);
  WsBatchPlan p{};
  p.mode = WS_MODE_POSTFILTER;
  p.node = node;
  p.k = (uint32_t)qp->k;
  p.qp = *qp;
  p.pad_id = pad == WS_PAD_ZERO ? 0u : 0xFFFFFFFFu;
  p.use_decode = idx && idx->has_decode;
  return ws_run_batch(idx, p, queries, windows, nq, ids, dists, flags);
}

int ws_tree_batch(ws_index* idx, int method, const float* queries, const float* windows,
                  uint64_t nq, const ws_query_params* qp, uint32_t* ids, float* dists,
                  uint32_t flags) {
  if (!qp) return ws_fail(WS_ERR_BADARG, "null query params");
  if (method < 0 || method > 3) return ws_fail(WS_ERR_BADARG, "unknown method %d", method);
  if (qp->k < 1 || qp->k > kMaxK) return ws_fail(WS_ERR_BADARG, "k=%lld outside 1..%u", (long long)qp->k, kMaxK);
  WsBatchPlan p{};
  p.mode = method;
  p.node = -1;
  p.k = (uint32_t)qp->k;
  p.qp = *qp;
  p.pad_id = 0u;
  p.use_decode = idx && idx->has_decode;
  return ws_run_batch(idx, p, queries, windows, nq, ids, dists, flags);
}

// ---- plumbing ---------------------------------------------------------------------------
int ws_index_device(const ws_index* idx, int* device) {
  if (!idx || !device) return ws_fail(WS_ERR_BADARG, "null argument");
  *device = idx->device;
  return WS_OK;
}
#define WS_NEED_DEVICE(idx)                                                     \
  if (!(idx)) return ws_fail(WS_ERR_BADARG, "null index");                      \
  std::lock_guard<std::recursive_mutex> lock_((idx)->mu);                        \
  if ((idx)->device < 0) return ws_fail(WS_ERR_CUDA, "host-only geometry index"); \
  WS_CUDA(cudaSetDevice((idx)->device));

int ws_index_sync(ws_index* idx) {
  WS_NEED_DEVICE(idx);
  WS_CUDA(cudaStreamSynchronize(idx->stream));
  uint32_t h_sticky = 0;
  WS_CUDA(cudaMemcpy(&h_sticky, idx->d_sticky, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (h_sticky) return ws_report_sticky(idx, h_sticky);
  return WS_OK;
}
int ws_device_alloc(ws_index* idx, size_t bytes, void** dptr) {
  WS_NEED_DEVICE(idx);
  if (!dptr) return ws_fail(WS_ERR_BADARG, "null dptr");
  WS_CUDA(cudaMalloc(dptr, bytes));
  return WS_OK;
}
int ws_device_free(ws_index* idx, void* dptr) {
  WS_NEED_DEVICE(idx);
  WS_CUDA(cudaStreamSynchronize(idx->stream));
  WS_CUDA(cudaFree(dptr));
  return WS_OK;
}
int ws_host_alloc_pinned(size_t bytes, void** hptr) {
  if (!hptr) return ws_fail(WS_ERR_BADARG, "null hptr");
  WS_CUDA(cudaMallocHost(hptr, bytes));
  return WS_OK;
}
int ws_host_free_pinned(void* hptr) {
  WS_CUDA(cudaFreeHost(hptr));
  return WS_OK;
}
int ws_copy_h2d(ws_index* idx, void* dst, const void* src, size_t bytes) {
  WS_NEED_DEVICE(idx);
  WS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, idx->stream));
  WS_CUDA(cudaStreamSynchronize(idx->stream));
  return WS_OK;
}
int ws_copy_d2h(ws_index* idx, void* dst, const void* src, size_t bytes) {
  WS_NEED_DEVICE(idx);
  WS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, idx->stream));
  WS_CUDA(cudaStreamSynchronize(idx->stream));
  return WS_OK;
}
int ws_timer_start(ws_index* idx) {
  WS_NEED_DEVICE(idx);
  WS_CUDA(cudaEventRecord(idx->ev0, idx->stream));
  return WS_OK;
}
int ws_timer_stop(ws_index* idx, float* elapsed_ms) {
  WS_NEED_DEVICE(idx);
  if (!elapsed_ms) return ws_fail(WS_ERR_BADARG, "null elapsed_ms");
  WS_CUDA(cudaEventRecord(idx->ev1, idx->stream));
  WS_CUDA(cudaEventSynchronize(idx->ev1));
  WS_CUDA(cudaEventElapsedTime(elapsed_ms, idx->ev0, idx->ev1));
  return WS_OK;
}
int ws_flush_l2(ws_index* idx) {
  WS_NEED_DEVICE(idx);
  const size_t bytes = 256ull << 20;  // 2x the 126 MB L2
  WS_TRY(ws_ensure(idx, idx->flush, bytes));
  WS_CUDA(wsl_fill(idx->num_sms * 8, idx->stream, (uint4*)idx->flush.p, bytes / 16));
  return WS_OK;
}


// ------------------------------------------------------------------------------------------
// GPU graph construction (ws_build.cuh) — setup path, not the query hot path
// ------------------------------------------------------------------------------------------
}  // extern "C"

// the reference's batch schedule (vamana/index.h:225-268)
static void ws_build_schedule(size_t n, std::vector<std::pair<uint32_t, uint32_t>>& out) {
  out.clear();
  size_t m = n, inc = 0, count = 0;
  size_t max_batch = std::min(static_cast<size_t>(0.02 * static_cast<float>(n)), (size_t)1000000ul);
  if (max_batch == 0) max_batch = n;
  while (count < m) {
    size_t floor, ceiling;
    if (std::pow(2.0, (double)inc) <= (double)max_batch) {
      floor = static_cast<size_t>(std::pow(2.0, (double)inc)) - 1;
      ceiling = std::min(static_cast<size_t>(std::pow(2.0, (double)(inc + 1))), m) - 1;
      count = std::min(static_cast<size_t>(std::pow(2.0, (double)(inc + 1))), m) - 1;
    } else {
      floor = count;
      ceiling = std::min(count + max_batch, m);
      count += max_batch;
    }
    if (ceiling > floor) out.push_back({(uint32_t)floor, (uint32_t)ceiling});
    inc++;
  }
}

extern "C" {

int ws_build_graphs(ws_index* idx, uint32_t ngraphs, const uint64_t* starts, const uint64_t* counts,
                    uint32_t max_degree, uint32_t beam_l, double alpha, uint64_t seed, int32_t* nodes_out) {
  WS_NEED_DEVICE(idx);
  if (idx->finalized) return ws_fail(WS_ERR_STATE, "index already finalized");
  if (!starts || !counts || !nodes_out || ngraphs == 0) return ws_fail(WS_ERR_BADARG, "null/empty argument");
  if (max_degree == 0 || max_degree > 64) return ws_fail(WS_ERR_BADARG, "max_degree %u unsupported (1..64)", max_degree);
  if (beam_l < 1 || beam_l > 1024) return ws_fail(WS_ERR_BADARG, "build beam L=%u outside 1..1024", beam_l);
  const uint32_t R = (max_degree + 3u) & ~3u;
  if (idx->R == 0) idx->R = R;
  if (idx->R != R) return ws_fail(WS_ERR_BADARG, "all graphs of one index must share max_degree (%u vs %u)", idx->R, R);
  cudaStream_t st = idx->stream;

  // ---- layout + insertion orders + schedules
  std::vector<WsBuildGraph> g(ngraphs);
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> sched(ngraphs);
  uint64_t rows = 0;
  size_t max_rounds = 0;
  for (uint32_t i = 0; i < ngraphs; i++) {
    if (counts[i] == 0 || starts[i] + counts[i] > idx->n) return ws_fail(WS_ERR_BADARG, "graph %u range outside the arena", i);
    g[i].start = (uint32_t)starts[i];
    g[i].count = (uint32_t)counts[i];
    g[i].row_off = (uint32_t)rows;
    g[i].floor = g[i].ceil = g[i].task_off = 0;
    rows += counts[i];
    ws_build_schedule(counts[i], sched[i]);
    max_rounds = std::max(max_rounds, sched[i].size());
  }
  if (rows >= (1ull << 32)) return ws_fail(WS_ERR_BADARG, "too many node-rows for one build call");
  std::vector<int32_t> perm(rows);
  size_t max_tasks = 0;
  {
    std::vector<size_t> per_round(max_rounds, 0);
    for (uint32_t i = 0; i < ngraphs; i++) {
      std::mt19937_64 rng(seed * 0x9E3779B97F4A7C15ull + i);
      int32_t* p = perm.data() + g[i].row_off;
      for (uint32_t j = 0; j < g[i].count; j++) p[j] = (int32_t)j;
      for (uint32_t j = g[i].count; j > 1; j--) std::swap(p[j - 1], p[rng() % j]);
      for (size_t r = 0; r < sched[i].size(); r++) per_round[r] += sched[i][r].second - sched[i][r].first;
    }
    for (size_t v : per_round) max_tasks = std::max(max_tasks, v);
  }

  int32_t *d_adj = nullptr, *d_deg = nullptr, *d_perm = nullptr, *d_new_out = nullptr, *d_new_cnt = nullptr;
  WsBuildGraph* d_graphs = nullptr;
  uint64_t *d_pairs = nullptr, *d_pairs2 = nullptr;
  uint32_t *d_heads = nullptr, *d_ctrl = nullptr;
  void* d_temp = nullptr;
  size_t temp_bytes = 0;
  unsigned long long* d_bstats = nullptr;
  const size_t max_pairs = max_tasks * R;
  wsl_sort_keys(nullptr, &temp_bytes, d_pairs, d_pairs2, (int)max_pairs, st);
  auto cleanup = [&]() {
    cudaFree(d_perm); cudaFree(d_new_out); cudaFree(d_new_cnt); cudaFree(d_graphs); cudaFree(d_pairs);
    cudaFree(d_pairs2); cudaFree(d_heads); cudaFree(d_ctrl); cudaFree(d_temp); cudaFree(d_bstats);
  };
#define WS_B(expr)                                                                                   \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      cleanup();                                                                                     \
      cudaFree(d_adj); cudaFree(d_deg);                                                              \
      return ws_fail(_e == cudaErrorMemoryAllocation ? WS_ERR_OOM : WS_ERR_CUDA, "%s: %s (graph build)", #expr, cudaGetErrorString(_e)); \
    }                                                                                                \
  } while (0)
  WS_B(cudaMalloc(&d_adj, rows * R * sizeof(int32_t)));
  WS_B(cudaMalloc(&d_deg, rows * sizeof(int32_t)));
  WS_B(cudaMalloc(&d_perm, rows * sizeof(int32_t)));
  WS_B(cudaMalloc(&d_new_out, std::max<size_t>(1, max_tasks) * R * sizeof(int32_t)));
  WS_B(cudaMalloc(&d_new_cnt, std::max<size_t>(1, max_tasks) * sizeof(int32_t)));
  WS_B(cudaMalloc(&d_graphs, ngraphs * sizeof(WsBuildGraph)));
  WS_B(cudaMalloc(&d_pairs, std::max<size_t>(1, max_pairs) * sizeof(uint64_t)));
  WS_B(cudaMalloc(&d_pairs2, std::max<size_t>(1, max_pairs) * sizeof(uint64_t)));
  WS_B(cudaMalloc(&d_heads, std::max<size_t>(1, max_pairs) * sizeof(uint32_t)));
  WS_B(cudaMalloc(&d_ctrl, 16 * sizeof(uint32_t)));
  WS_B(cudaMalloc(&d_temp, std::max<size_t>(16, temp_bytes)));
  WS_B(cudaMalloc(&d_bstats, 8 * sizeof(unsigned long long)));
  WS_B(cudaMemsetAsync(d_bstats, 0, 8 * sizeof(unsigned long long), st));
  WS_B(cudaMemcpyAsync(d_perm, perm.data(), rows * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  WS_B(wsl_fill_i32(idx->num_sms * 8, st, d_adj, rows * R, -1));
  WS_B(cudaMemsetAsync(d_deg, 0, rows * sizeof(int32_t), st));

  const int kq = ws_pick_kq(idx->dpad);
  uint32_t beam_cap = 64;
  while (beam_cap < beam_l) beam_cap <<= 1;
  const uint32_t E = (uint32_t)std::max<int64_t>(1, std::min<int64_t>(8, idx->opt_build_expand));
  uint32_t cand_cap = 64;
  while (cand_cap < E * R) cand_cap <<= 1;
  const uint32_t hash_entries = 8192;
  const size_t smem = (size_t)2 * beam_cap * 8 + (size_t)cand_cap * 24 + (size_t)idx->dpad * 4 + (size_t)hash_entries * 4 +
                      (size_t)WS_BUILD_VCAP * 8;
  int occ = 0;
  WS_B(wsl_build_insert_occ(kq, idx->metric, smem, &occ));
  if (occ < 1) { cleanup(); cudaFree(d_adj); cudaFree(d_deg); return ws_fail(WS_ERR_CUDA, "build kernel does not fit (smem %zu)", smem); }

  for (size_t r = 0; r < max_rounds; r++) {
    uint32_t ntasks = 0;
    for (uint32_t i = 0; i < ngraphs; i++) {
      g[i].task_off = ntasks;
      if (r < sched[i].size()) { g[i].floor = sched[i][r].first; g[i].ceil = sched[i][r].second; }
      else { g[i].floor = g[i].ceil = 0; }
      ntasks += g[i].ceil - g[i].floor;
    }
    if (ntasks == 0) continue;
    WS_B(cudaMemcpyAsync(d_graphs, g.data(), ngraphs * sizeof(WsBuildGraph), cudaMemcpyHostToDevice, st));
    WS_B(cudaMemsetAsync(d_ctrl, 0, 16 * sizeof(uint32_t), st));
    WsBuildArgs a;
    a.vecs = idx->d_vecs; a.dim = idx->dim; a.dpad = idx->dpad; a.R = R;
    a.adj = d_adj; a.deg = d_deg; a.perm = d_perm; a.graphs = d_graphs; a.ngraphs = ngraphs; a.ntasks = ntasks;
    a.head = d_ctrl + 0; a.new_out = d_new_out; a.new_cnt = d_new_cnt; a.pairs = d_pairs; a.pair_count = d_ctrl + 1;
    a.L = beam_l; a.beam_cap = beam_cap; a.hash_mask = hash_entries - 1; a.cand_cap = cand_cap; a.expand = E;
    a.alpha = alpha; a.stats = d_bstats;
    const int grid = (int)std::min<uint64_t>(ntasks, (uint64_t)idx->num_sms * occ);
    WS_B(wsl_build_insert(kq, idx->metric, grid, smem, st, a));
    WS_B(wsl_build_apply((int)((ntasks + 7) / 8), st, a));
    uint32_t npairs = 0;
    WS_B(cudaMemcpyAsync(&npairs, d_ctrl + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    WS_B(cudaStreamSynchronize(st));
    if (npairs == 0) continue;
    WS_B(wsl_sort_keys(d_temp, &temp_bytes, d_pairs, d_pairs2, (int)npairs, st));
    WS_B(wsl_build_heads((int)((npairs + 255) / 256), st, d_pairs2, npairs, d_heads, d_ctrl + 2));
    uint32_t nheads = 0;
    WS_B(cudaMemcpyAsync(&nheads, d_ctrl + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    WS_B(cudaStreamSynchronize(st));
    WsBuildRevArgs ra;
    ra.vecs = idx->d_vecs; ra.dim = idx->dim; ra.dpad = idx->dpad; ra.R = R; ra.adj = d_adj; ra.deg = d_deg;
    ra.graphs = d_graphs; ra.ngraphs = ngraphs; ra.pairs = d_pairs2; ra.npairs = npairs; ra.heads = d_heads;
    ra.nheads = nheads; ra.head = d_ctrl + 3; ra.alpha = alpha; ra.stats = d_bstats;
    const int rgrid = (int)std::min<uint64_t>(nheads, (uint64_t)idx->num_sms * 8);
    WS_B(wsl_build_reverse(kq, idx->metric, rgrid, st, ra));
  }
  {
    WsBuildSortArgs sa;
    sa.vecs = idx->d_vecs; sa.dim = idx->dim; sa.dpad = idx->dpad; sa.R = R; sa.adj = d_adj; sa.deg = d_deg;
    sa.graphs = d_graphs; sa.ngraphs = ngraphs; sa.rows = (uint32_t)rows;
    const int sgrid = (int)std::min<uint64_t>(rows, (uint64_t)idx->num_sms * 16);
    WS_B(wsl_build_sort(kq, idx->metric, sgrid, st, sa));
  }
  unsigned long long hs[8];
  WS_B(cudaMemcpyAsync(hs, d_bstats, sizeof(hs), cudaMemcpyDeviceToHost, st));
  WS_B(cudaStreamSynchronize(st));
#undef WS_B
  cleanup();
  idx->build_stats[0] += hs[0]; idx->build_stats[1] += hs[1]; idx->build_stats[2] += hs[2]; idx->build_stats[3] += hs[3];
  idx->build_allocs.push_back(d_adj);
  idx->build_allocs.push_back(d_deg);
  idx->hbm_bytes += rows * R * sizeof(int32_t) + rows * sizeof(int32_t);
  for (uint32_t i = 0; i < ngraphs; i++) {
    WsNode node;
    node.adj = d_adj + (size_t)g[i].row_off * R;
    node.start = g[i].start;
    node.count = g[i].count;
    idx->h_nodes.push_back(node);
    idx->node_deg.push_back(d_deg + g[i].row_off);
    idx->max_node_count = std::max<uint32_t>(idx->max_node_count, g[i].count);
    nodes_out[i] = (int32_t)idx->h_nodes.size() - 1;
  }
  return WS_OK;
}

// degrees[count] and rows[count][R] (-1 padded) of a node, for saving in the reference's
// graph format (graph.h:174-196)
int ws_index_get_graph(ws_index* idx, int32_t node, uint32_t* max_degree_out, int32_t* degrees, int32_t* rows) {
  WS_NEED_DEVICE(idx);
  if (node < 0 || (size_t)node >= idx->h_nodes.size()) return ws_fail(WS_ERR_BADARG, "bad node handle %d", node);
  const WsNode& nd = idx->h_nodes[node];
  if (max_degree_out) *max_degree_out = idx->R;
  if (rows) WS_CUDA(cudaMemcpy(rows, nd.adj, (size_t)nd.count * idx->R * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (degrees) {
    if (idx->node_deg[node]) {
      WS_CUDA(cudaMemcpy(degrees, idx->node_deg[node], (size_t)nd.count * sizeof(int32_t), cudaMemcpyDeviceToHost));
    } else {
      if (!rows) return ws_fail(WS_ERR_BADARG, "degrees of a loaded graph need the rows buffer too");
      for (uint32_t i = 0; i < nd.count; i++) {
        int32_t d = 0;
        while (d < (int32_t)idx->R && rows[(size_t)i * idx->R + d] >= 0) d++;
        degrees[i] = d;
      }
    }
  }
  return WS_OK;
}

int ws_index_build_stats(ws_index* idx, uint64_t* out4) {
  if (!idx || !out4) return ws_fail(WS_ERR_BADARG, "null argument");
  for (int i = 0; i < 4; i++) out4[i] = idx->build_stats[i];
  return WS_OK;
}

int ws_merge_partial_topk(ws_index* idx, const uint32_t* ids, const float* dists, uint32_t parts, uint64_t nq,
                          uint32_t k, uint32_t pad_id, uint32_t* out_ids, float* out_dists) {
  WS_NEED_DEVICE(idx);
  if (!ids || !dists || !out_ids || !out_dists) return ws_fail(WS_ERR_BADARG, "null buffer");
  if (k == 0 || k > kMaxK || parts == 0) return ws_fail(WS_ERR_BADARG, "k=%u parts=%u", k, parts);
  if (nq == 0) return WS_OK;
  if (parts > WS_MAX_PARTS) return ws_fail(WS_ERR_BADARG, "parts=%u above %d", parts, WS_MAX_PARTS);
  WsMergePartsArgs a{};
  a.ids = ids; a.dists = dists; a.parts = parts; a.k = k; a.nq = (uint32_t)nq; a.nq_total = (uint32_t)nq; a.q0 = 0; a.pad_id = pad_id;
  a.out_ids = out_ids; a.out_dists = out_dists;
  int grid = (int)std::min<uint64_t>(nq, (uint64_t)idx->num_sms * 16);
  WS_CUDA(wsl_merge_parts(grid, idx->stream, a));
  idx->launches++;
  return WS_OK;
}

int ws_index_set_decode(ws_index* idx, const uint32_t* decode) {
  WS_NEED_DEVICE(idx);
  if (!decode) return ws_fail(WS_ERR_BADARG, "null decode table");
  WS_CUDA(cudaStreamSynchronize(idx->stream));
  if (!idx->d_decode) {
    WS_CUDA(cudaMalloc(&idx->d_decode, idx->n * sizeof(uint32_t)));
    idx->hbm_bytes += idx->n * sizeof(uint32_t);
    idx->has_decode = true;
  }
  WS_CUDA(cudaMemcpy(idx->d_decode, decode, idx->n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  return WS_OK;
}

int ws_index_get_stats(ws_index* idx, ws_stats* out) {
  WS_NEED_DEVICE(idx);
  if (!out) return ws_fail(WS_ERR_BADARG, "null out");
  unsigned long long h[8];
  WS_CUDA(cudaStreamSynchronize(idx->stream));
  WS_CUDA(cudaMemcpy(h, idx->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
  out->graph_searches = h[WS_ST_SEARCHES];
  out->visited = h[WS_ST_VISITED];
  out->dist_cmps = h[WS_ST_DISTCMPS];
  out->scan_points = h[WS_ST_SCANPTS];
  out->graph_tasks = h[WS_ST_GTASKS];
  out->scan_tasks = h[WS_ST_STASKS];
  out->escalated_tasks = h[WS_ST_ESCALATED];
  out->beam_sum = h[WS_ST_BEAMSUM];
  return WS_OK;
}
int ws_index_reset_stats(ws_index* idx) {
  WS_NEED_DEVICE(idx);
  WS_CUDA(cudaStreamSynchronize(idx->stream));
  WS_CUDA(cudaMemset(idx->d_stats, 0, 8 * sizeof(unsigned long long)));
  return WS_OK;
}
int ws_index_launch_count(const ws_index* idx, uint64_t* out) {
  if (!idx || !out) return ws_fail(WS_ERR_BADARG, "null argument");
  *out = idx->launches;
  return WS_OK;
}
int ws_index_set_option(ws_index* idx, const char* name, int64_t value) {
  if (!idx || !name) return ws_fail(WS_ERR_BADARG, "null argument");
  std::lock_guard<std::recursive_mutex> lock(idx->mu);
  std::string s(name);
  if (s == "expand_width") {
    if (value < 1 || value > 8) return ws_fail(WS_ERR_BADARG, "expand_width must be 1..8");
    idx->opt_expand = value;
  } else if (s == "emulate_query_id_skip") {
    idx->opt_skip_query_id = value != 0;
  } else if (s == "stage_copies") {
    idx->opt_stage_copies = value != 0;
  } else if (s == "prefilter_open_tail") {
    idx->opt_open_tail = value != 0;
    idx->hgeom.pf_n = idx->dgeom.pf_n = idx->n + (idx->opt_open_tail ? 1 : 0);
  } else if (s == "scan_chunk") {
    if (value < 256) return ws_fail(WS_ERR_BADARG, "scan_chunk must be >= 256");
    idx->opt_scan_chunk = value;
  } else if (s == "warp_tiers") {
    idx->opt_warp_tiers = value != 0;
  } else if (s == "hash16") {
    idx->opt_hash16 = value != 0;
  } else if (s == "warp256") {
    idx->opt_warp256 = value != 0;
  } else if (s == "warp256_min") {
    idx->opt_warp256_min = value < 0 ? 0 : value;
  } else if (s == "fuse_scan") {
    idx->opt_fuse_scan = value != 0;
  } else if (s == "warp_scan") {
    idx->opt_warp_scan = value != 0;
  } else if (s == "warp_hash") {
    if (value < 256 || value > 16384 || (value & (value - 1))) return ws_fail(WS_ERR_BADARG, "warp_hash must be a power of two in 256..16384");
    idx->opt_warp_hash = value;
  } else if (s == "build_expand_width") {
    if (value < 1 || value > 8) return ws_fail(WS_ERR_BADARG, "build_expand_width must be 1..8");
    idx->opt_build_expand = value;
  } else if (s == "prefilter_direct") {
    if (value < 0 || value > 2) return ws_fail(WS_ERR_BADARG, "prefilter_direct must be 0 (never), 1 (always) or 2 (auto)");
    idx->opt_direct = value;
  } else if (s == "gemm_prefilter") {
    if (value < 0 || value > 2) return ws_fail(WS_ERR_BADARG, "gemm_prefilter must be 0 (never), 1 (always when eligible) or 2 (auto)");
    idx->opt_gemm = value;
  } else if (s == "gemm_min_window") {
    idx->opt_gemm_min_window = value < 0 ? 0 : value;
  } else if (s == "gemm_dynamic") {
    idx->opt_gemm_dynamic = value != 0;
  } else if (s == "gemm_items") {
    if (value < 0 || value > 65536) return ws_fail(WS_ERR_BADARG, "gemm_items must be 0..65536");
    idx->opt_gemm_items = value;
  } else if (s == "gemm_debug") {
    idx->opt_gemm_debug = value;
  } else if (s == "gemm_chunk_mb") {
    if (value < 1 || value > 4096) return ws_fail(WS_ERR_BADARG, "gemm_chunk_mb must be 1..4096");
    idx->opt_gemm_chunk_mb = value;
  } else if (s == "gemm_min_tiles") {
    if (value < 1) return ws_fail(WS_ERR_BADARG, "gemm_min_tiles must be >= 1");
    idx->opt_gemm_min_tiles = value;
  } else if (s == "profile_kernels") {
    idx->opt_profile = value != 0;
  } else if (s == "hash_factor") {
    if (value < 1 || value > 64) return ws_fail(WS_ERR_BADARG, "hash_factor must be 1..64");
    idx->opt_hash_factor = value;
  } else {
    return ws_fail(WS_ERR_BADARG, "unknown option '%s'", name);
  }
  return WS_OK;
}
int ws_index_kernel_times(ws_index* idx, double* ms_out, uint64_t* launches_out, int reset) {
  WS_NEED_DEVICE(idx);
  WS_CUDA(cudaStreamSynchronize(idx->stream));
  for (const ws_index::EvPair& p : idx->ev_pairs) {
    float ms = 0.f;
    WS_CUDA(cudaEventElapsedTime(&ms, idx->ev_pool[p.a], idx->ev_pool[p.b]));
    idx->kernel_ms[p.kind] += ms;
    idx->kernel_launches[p.kind]++;
  }
  idx->ev_pairs.clear();
  idx->ev_used = 0;
  for (int i = 0; i < WS_NUM_KERNEL_KINDS; i++) {
    if (ms_out) ms_out[i] = idx->kernel_ms[i];
    if (launches_out) launches_out[i] = idx->kernel_launches[i];
    if (reset) { idx->kernel_ms[i] = 0; idx->kernel_launches[i] = 0; }
  }
  return WS_OK;
}

int ws_index_hbm_bytes(const ws_index* idx, uint64_t* out) {
  if (!idx || !out) return ws_fail(WS_ERR_BADARG, "null argument");
  *out = idx->hbm_bytes;
  return WS_OK;
}

int ws_index_task_capacity(ws_index* idx, int method, uint32_t* cap) {
  if (!idx || !cap) return ws_fail(WS_ERR_BADARG, "null argument");
  if (!idx->finalized) return ws_fail(WS_ERR_STATE, "ws_index_finalize has not been called");
  *cap = ws_task_capacity(idx, method);
  return WS_OK;
}

int ws_debug_decompose_host(ws_index* idx, int method, const float* windows, uint64_t nq,
                            const ws_query_params* qp, uint32_t cap, int64_t* out_tasks,
                            uint32_t* out_counts) {
  if (!idx || !windows || !qp || !out_tasks || !out_counts) return ws_fail(WS_ERR_BADARG, "null argument");
  if (!idx->finalized) return ws_fail(WS_ERR_STATE, "ws_index_finalize has not been called");
  if (method < 0 || (method > 3 && method != WS_MODE_PREFILTER)) return ws_fail(WS_ERR_BADARG, "unknown method %d", method);
  if (method <= 2 && idx->wst_rows == 0) return ws_fail(WS_ERR_STATE, "no B-WST geometry set");
  if (method == 3 && idx->sup_rows == 0) return ws_fail(WS_ERR_STATE, "no super-postfilter geometry set");
  WsDecompParams p;
  p.beam = (uint32_t)qp->beam_size;
  p.has_ratio = qp->has_min_query_to_bucket_ratio;
  p.min_ratio = qp->min_query_to_bucket_ratio;
  p.scan_chunk = (uint32_t)idx->opt_scan_chunk;
  std::vector<WsTask> slots(cap);
  for (uint64_t q = 0; q < nq; q++) {
    WsEmitter em;
    em.slots = slots.data(); em.cap = cap; em.count = 0; em.overflow = 0; em.query = (uint32_t)q;
    em.beam = p.beam; em.scan_chunk = p.scan_chunk;
    float lo = windows[2 * q], hi = windows[2 * q + 1];
    switch (method) {
      case 0: ws_decompose_fenwick(idx->hgeom, lo, hi, 0, em); break;
      case 1: ws_decompose_opt_postfilter(idx->hgeom, lo, hi, p, em); break;
      case 2: ws_decompose_three_split(idx->hgeom, lo, hi, p, em); break;
      case 3: ws_decompose_super(idx->hgeom, lo, hi, em); break;
      default: {
        uint64_t s = ws_prefilter_bound(idx->hgeom.labels, idx->hgeom.pf_n, lo);
        uint64_t e = ws_prefilter_bound(idx->hgeom.labels, idx->hgeom.pf_n, hi);
        em.scan(s, e, lo, hi);
      }
    }
    if (em.overflow) return ws_fail(WS_ERR_STATE, "query %llu overflowed capacity %u", (unsigned long long)q, cap);
    out_counts[q] = em.count;
    for (uint32_t i = 0; i < cap; i++) {
      int64_t* o = out_tasks + ((size_t)q * cap + i) * 4;
      if (i < em.count) {
        const WsTask& t = slots[i];
        o[0] = t.node;
        if (t.node >= 0) { o[1] = idx->h_nodes[t.node].start; o[2] = o[1] + idx->h_nodes[t.node].count; }
        else { o[1] = t.a; o[2] = t.b; }
        o[3] = t.flags;
      } else {
        o[0] = o[1] = o[2] = o[3] = -2;
      }
    }
  }
  return WS_OK;
}

}  // extern "C"

#include "ws_group.inl"
#include "ws_snapshot.inl"
