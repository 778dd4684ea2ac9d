// ws_group.inl — multi-GPU inside the engine (included at the end of wsann.cu; SURVEY.md §8b "ws_ctx", §8e).
//
// The reference runs one `parlay::parallel_for` over the queries of a batch on all host cores
// (range_filter_tree.h:70, prefiltering.h:131, postfilter_vamana.h:199).  Here one host call drives all the
// GPUs of the box:
//
//   WS_GROUP_REPLICATED      every member holds the whole arena (ws_index_replicate clones it device to device
//                            over NVLink); a batch is cut into one contiguous slice per member, each slice runs on
//                            its member's stream from its own host thread.  No data-path collective: queries are
//                            independent units.
//   WS_GROUP_LABEL_SHARDED   member g holds a contiguous range of the label-sorted points with its own tree; every
//                            member answers the WHOLE batch on its shard (windows that miss the shard come back as
//                            pads), then the [nq][k] partial rows are exchanged and merged per query
//                            (sort_and_truncate across shards, range_filter_tree.h:542-549):
//                              exchange 0  one kernel per member gathers its slice of the queries from every peer's
//                                          HBM with plain loads over NVLink while it merges (peer access), no
//                                          staging buffer and no collective launch
//                              exchange 1  ncclAllGather of the rows on every member's stream, then the merge kernel
//
// One process per GPU (the torchrun-style plumbing of bench.py / label_shard.py) uses the same pieces through
// ws_nccl_unique_id / ws_index_comm_init / ws_allgather_merge.  NCCL is loaded with dlopen on first use so that
// the library neither needs it at link time nor fights over which libnccl.so.2 a host process (e.g. one that also
// imports torch) has already loaded.

#include <dlfcn.h>
#include <nccl.h>

#include <condition_variable>
#include <functional>
#include <thread>

namespace {

struct WsNccl {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
};

static WsNccl g_nccl;
static std::mutex g_nccl_mu;

static int ws_nccl_load() {
  std::lock_guard<std::mutex> lock(g_nccl_mu);
  if (g_nccl.lib) return WS_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* nm : names) {
    lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) return ws_fail(WS_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
  WsNccl t;
  t.lib = lib;
#define WS_NCCL_SYM(field, name)                                                          \
  t.field = reinterpret_cast<decltype(t.field)>(dlsym(lib, name));                         \
  if (!t.field) return ws_fail(WS_ERR_NCCL, "libnccl.so.2 lacks the symbol %s", name);
  WS_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  WS_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  WS_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  WS_NCCL_SYM(AllGather, "ncclAllGather")
  WS_NCCL_SYM(GroupStart, "ncclGroupStart")
  WS_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  WS_NCCL_SYM(GetErrorString, "ncclGetErrorString")
  WS_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef WS_NCCL_SYM
  g_nccl = t;
  return WS_OK;
}

#define WS_NCCL(expr)                                                                                   \
  do {                                                                                                  \
    ncclResult_t _r = (expr);                                                                           \
    if (_r != ncclSuccess) return ws_fail(WS_ERR_NCCL, "%s: %s", #expr, g_nccl.GetErrorString(_r));     \
  } while (0)

// exchange scratch of one arena: gathered rows of every rank + this rank's send buffers
static int ws_comm_scratch(ws_index* idx, uint32_t parts, uint64_t nq, uint32_t k) {
  WS_TRY(ws_ensure(idx, idx->x_all_ids, (size_t)parts * nq * k * sizeof(uint32_t)));
  WS_TRY(ws_ensure(idx, idx->x_all_dists, (size_t)parts * nq * k * sizeof(float)));
  return WS_OK;
}

}  // namespace

extern "C" {

int ws_nccl_unique_id(void* id_out) {
  if (!id_out) return ws_fail(WS_ERR_BADARG, "null id_out");
  WS_TRY(ws_nccl_load());
  static_assert(sizeof(ncclUniqueId) == WS_NCCL_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  WS_NCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(id_out, &id, sizeof(id));
  return WS_OK;
}

int ws_nccl_version(int* version) {
  if (!version) return ws_fail(WS_ERR_BADARG, "null version");
  WS_TRY(ws_nccl_load());
  WS_NCCL(g_nccl.GetVersion(version));
  return WS_OK;
}

int ws_index_comm_init(ws_index* idx, int nranks, int rank, const void* id) {
  WS_NEED_DEVICE(idx);
  if (!id || nranks < 1 || rank < 0 || rank >= nranks) return ws_fail(WS_ERR_BADARG, "rank %d of %d", rank, nranks);
  if (nranks > WS_MAX_PARTS) return ws_fail(WS_ERR_BADARG, "at most %d ranks", WS_MAX_PARTS);
  if (idx->comm) return ws_fail(WS_ERR_STATE, "this arena already has a communicator");
  WS_TRY(ws_nccl_load());
  ncclUniqueId uid;
  std::memcpy(&uid, id, sizeof(uid));
  ncclComm_t comm = nullptr;
  WS_NCCL(g_nccl.CommInitRank(&comm, nranks, uid, rank));
  idx->comm = comm;
  idx->comm_rank = rank;
  idx->comm_size = nranks;
  return WS_OK;
}

int ws_index_comm_destroy(ws_index* idx) {
  WS_NEED_DEVICE(idx);
  if (idx->comm) {
    cudaStreamSynchronize(idx->stream);
    g_nccl.CommDestroy((ncclComm_t)idx->comm);
    idx->comm = nullptr;
  }
  return WS_OK;
}

// ncclAllGather of this rank's [nq][k] rows into [ranks][nq][k] (ids and dists), then K4b on the index stream.
// Everything is device memory; nothing is synchronised here.
int ws_allgather_merge(ws_index* idx, const uint32_t* ids, const float* dists, uint64_t nq, uint32_t k, uint32_t pad_id,
                       uint32_t* out_ids, float* out_dists) {
  WS_NEED_DEVICE(idx);
  if (!idx->comm) return ws_fail(WS_ERR_STATE, "ws_index_comm_init has not been called");
  if (!ids || !dists || !out_ids || !out_dists) return ws_fail(WS_ERR_BADARG, "null buffer");
  if (k == 0 || k > kMaxK) return ws_fail(WS_ERR_BADARG, "k=%u", k);
  if (nq == 0) return WS_OK;
  const uint32_t parts = (uint32_t)idx->comm_size;
  WS_TRY(ws_comm_scratch(idx, parts, nq, k));
  const size_t count = (size_t)nq * k;
  WS_NCCL(g_nccl.GroupStart());
  WS_NCCL(g_nccl.AllGather(ids, idx->x_all_ids.p, count, ncclUint32, (ncclComm_t)idx->comm, idx->stream));
  WS_NCCL(g_nccl.AllGather(dists, idx->x_all_dists.p, count, ncclFloat32, (ncclComm_t)idx->comm, idx->stream));
  WS_NCCL(g_nccl.GroupEnd());
  idx->launches += 2;
  WsMergePartsArgs a{};
  a.ids = (const uint32_t*)idx->x_all_ids.p; a.dists = (const float*)idx->x_all_dists.p;
  a.parts = parts; a.k = k; a.pad_id = pad_id; a.nq_total = (uint32_t)nq; a.q0 = 0; a.nq = (uint32_t)nq;
  a.out_ids = out_ids; a.out_dists = out_dists;
  const int grid = (int)std::min<uint64_t>(nq, (uint64_t)idx->num_sms * 16);
  WS_CUDA(wsl_merge_parts(grid, idx->stream, a));
  idx->launches++;
  return WS_OK;
}

// ---- replication ---------------------------------------------------------------------------------
// Clone of a finalized arena on another device: vectors, labels, decode table and every adjacency array travel
// device to device (cudaMemcpyPeerAsync: NVLink when peer access is possible, staged by the driver otherwise);
// geometry and node table are re-derived from the host mirrors.  Options are copied; scratch is not.
int ws_index_replicate(ws_index* src, int device, ws_index** out) {
  if (!src || !out) return ws_fail(WS_ERR_BADARG, "null argument");
  *out = nullptr;
  std::lock_guard<std::recursive_mutex> lock(src->mu);
  if (src->device < 0) return ws_fail(WS_ERR_CUDA, "host-only geometry index");
  if (!src->finalized) return ws_fail(WS_ERR_STATE, "only a finalized arena can be replicated");
  int ndev = 0;
  WS_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return ws_fail(WS_ERR_CUDA, "no CUDA device %d", device);
  ws_index* idx = new ws_index();
#define WS_REP_CUDA(expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ws_index_destroy(idx);                                                                       \
      return ws_fail(_e == cudaErrorMemoryAllocation ? WS_ERR_OOM : WS_ERR_CUDA, "%s: %s", #expr,  \
                     cudaGetErrorString(_e));                                                      \
    }                                                                                              \
  } while (0)
  idx->device = device; idx->metric = src->metric; idx->n = src->n; idx->dim = src->dim; idx->dpad = src->dpad;
  idx->label_sorted = src->label_sorted; idx->has_decode = src->has_decode;
  idx->h_labels = src->h_labels;
  idx->R = src->R; idx->max_node_count = src->max_node_count;
  idx->wst_nb = src->wst_nb; idx->wst_off_ptr = src->wst_off_ptr; idx->wst_node_ptr = src->wst_node_ptr;
  idx->wst_off = src->wst_off; idx->wst_nodes = src->wst_nodes;
  idx->sup_size = src->sup_size; idx->sup_shift = src->sup_shift; idx->sup_nb = src->sup_nb;
  idx->sup_node_ptr = src->sup_node_ptr; idx->sup_nodes = src->sup_nodes;
  idx->wst_rows = src->wst_rows; idx->split = src->split; idx->sup_rows = src->sup_rows;
  idx->wst_prefilter_nodes = src->wst_prefilter_nodes; idx->cutoff = src->cutoff; idx->sup_cutoff = src->sup_cutoff;
  idx->opt_expand = src->opt_expand; idx->opt_skip_query_id = src->opt_skip_query_id; idx->opt_scan_chunk = src->opt_scan_chunk;
  idx->opt_warp_tiers = src->opt_warp_tiers; idx->opt_warp_hash = src->opt_warp_hash; idx->opt_warp_scan = src->opt_warp_scan;
  idx->opt_hash16 = src->opt_hash16; idx->opt_direct = src->opt_direct; idx->opt_gemm = src->opt_gemm;
  idx->opt_gemm_min_window = src->opt_gemm_min_window; idx->opt_gemm_chunk_mb = src->opt_gemm_chunk_mb;
  idx->opt_open_tail = src->opt_open_tail;
  WS_REP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  WS_REP_CUDA(cudaGetDeviceProperties(&prop, device));
  idx->num_sms = prop.multiProcessorCount;
  idx->smem_optin = prop.sharedMemPerBlockOptin;
  WS_REP_CUDA(cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking));
  WS_REP_CUDA(cudaEventCreate(&idx->ev0));
  WS_REP_CUDA(cudaEventCreate(&idx->ev1));
  cudaStream_t st = idx->stream;
  const size_t vbytes = (size_t)idx->n * idx->dpad * sizeof(float);
  WS_REP_CUDA(cudaMalloc(&idx->d_vecs, vbytes));
  WS_REP_CUDA(cudaMemcpyPeerAsync(idx->d_vecs, device, src->d_vecs, src->device, vbytes, st));
  WS_REP_CUDA(cudaMalloc(&idx->d_labels, idx->n * sizeof(float)));
  WS_REP_CUDA(cudaMemcpyPeerAsync(idx->d_labels, device, src->d_labels, src->device, idx->n * sizeof(float), st));
  idx->hbm_bytes += vbytes + idx->n * sizeof(float);
  if (src->d_decode) {
    WS_REP_CUDA(cudaMalloc(&idx->d_decode, idx->n * sizeof(uint32_t)));
    WS_REP_CUDA(cudaMemcpyPeerAsync(idx->d_decode, device, src->d_decode, src->device, idx->n * sizeof(uint32_t), st));
    idx->hbm_bytes += idx->n * sizeof(uint32_t);
  }
  WS_REP_CUDA(cudaMalloc(&idx->d_stats, 8 * sizeof(unsigned long long)));
  WS_REP_CUDA(cudaMemsetAsync(idx->d_stats, 0, 8 * sizeof(unsigned long long), st));
  WS_REP_CUDA(cudaMalloc(&idx->d_sticky, 4 * sizeof(uint32_t)));
  WS_REP_CUDA(cudaMemsetAsync(idx->d_sticky, 0, 4 * sizeof(uint32_t), st));
  // adjacency: one bump allocation per node, as ws_index_add_graph lays them out
  idx->h_nodes = src->h_nodes;
  idx->node_deg.assign(src->h_nodes.size(), nullptr);
  for (size_t i = 0; i < src->h_nodes.size(); i++) {
    const WsNode& sn = src->h_nodes[i];
    const size_t bytes = (size_t)sn.count * idx->R * sizeof(int32_t);
    const size_t aligned = (bytes + 255) & ~(size_t)255;
    void* dst = nullptr;
    if (aligned > kAdjSlabBytes) {
      WS_REP_CUDA(cudaMalloc(&dst, aligned));
      idx->adj_slabs.insert(idx->adj_slabs.begin(), dst);
      idx->hbm_bytes += aligned;
      if (idx->adj_slabs.size() == 1) idx->slab_used = kAdjSlabBytes;
    } else {
      if (idx->adj_slabs.empty() || idx->slab_used + aligned > kAdjSlabBytes) {
        void* slab = nullptr;
        WS_REP_CUDA(cudaMalloc(&slab, kAdjSlabBytes));
        idx->adj_slabs.push_back(slab);
        idx->slab_used = 0;
        idx->hbm_bytes += kAdjSlabBytes;
      }
      dst = (char*)idx->adj_slabs.back() + idx->slab_used;
      idx->slab_used += aligned;
    }
    WS_REP_CUDA(cudaMemcpyPeerAsync(dst, device, sn.adj, src->device, bytes, st));
    idx->h_nodes[i].adj = (const int32_t*)dst;
  }
  WS_REP_CUDA(cudaStreamSynchronize(st));
#undef WS_REP_CUDA
  int rc = ws_index_finalize(idx);
  if (rc != WS_OK) { ws_index_destroy(idx); return rc; }
  *out = idx;
  return WS_OK;
}

}  // extern "C"

// ---- the group -------------------------------------------------------------------------------------
struct ws_group {
  int mode = WS_GROUP_REPLICATED;
  std::vector<ws_index*> m;
  std::mutex call_mu;  // one batch at a time per group
  // one persistent host thread per member (the CUDA runtime's current device is per thread)
  std::vector<std::thread> workers;
  std::mutex mu;
  std::condition_variable cv_job, cv_done;
  std::function<int(int)> job;
  uint64_t generation = 0;
  int pending = 0;
  bool stop = false;
  std::vector<int> status;
  std::vector<std::string> errors;
  // label-sharded exchange
  int64_t opt_exchange = 0;  // 0 peer loads, 1 ncclAllGather
  bool peer_ok = false;
  bool comms_ready = false;
  std::vector<WsDevBuf> part_ids, part_dists, out_ids, out_dists;  // per member
  std::vector<cudaEvent_t> ev_done, ev_t0, ev_t1, ev_t2;
  double last_ms[3] = {0, 0, 0};
};

namespace {

static void ws_group_worker(ws_group* g, int i) {
  uint64_t seen = 0;
  cudaSetDevice(g->m[i]->device);
  for (;;) {
    std::function<int(int)> job;
    {
      std::unique_lock<std::mutex> lk(g->mu);
      g->cv_job.wait(lk, [&] { return g->stop || g->generation != seen; });
      if (g->stop) return;
      seen = g->generation;
      job = g->job;
    }
    int rc = job(i);
    std::string err = rc != WS_OK ? std::string(ws_last_error()) : std::string();
    {
      std::lock_guard<std::mutex> lk(g->mu);
      g->status[i] = rc;
      g->errors[i] = std::move(err);
      if (--g->pending == 0) g->cv_done.notify_all();
    }
  }
}

// runs fn(member) on every member's thread; first failure wins
static int ws_group_run(ws_group* g, std::function<int(int)> fn) {
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->job = std::move(fn);
    g->pending = (int)g->m.size();
    g->generation++;
  }
  g->cv_job.notify_all();
  {
    std::unique_lock<std::mutex> lk(g->mu);
    g->cv_done.wait(lk, [&] { return g->pending == 0; });
  }
  for (size_t i = 0; i < g->m.size(); i++)
    if (g->status[i] != WS_OK) return ws_fail(g->status[i], "device %d: %s", g->m[i]->device, g->errors[i].c_str());
  return WS_OK;
}

static int ws_group_ensure_comms(ws_group* g) {
  if (g->comms_ready) return WS_OK;
  WS_TRY(ws_nccl_load());
  ncclUniqueId uid;
  WS_NCCL(g_nccl.GetUniqueId(&uid));
  const int n = (int)g->m.size();
  // ncclCommInitRank blocks until every rank has joined: one thread per member
  WS_TRY(ws_group_run(g, [g, n, uid](int i) { return ws_index_comm_init(g->m[i], n, i, &uid); }));
  g->comms_ready = true;
  return WS_OK;
}

struct WsGroupCall {
  int kind;  // 0 prefilter, 1 postfilter, 2 tree
  int method;
  int32_t node;
  int pad;
  uint32_t k;
  const ws_query_params* qp;
};

static int ws_group_member_call(ws_index* idx, const WsGroupCall& c, const float* q, const float* w, uint64_t nq, uint32_t* ids,
                                float* dists, uint32_t flags) {
  switch (c.kind) {
    case 0: return ws_prefilter_batch(idx, q, w, nq, c.k, ids, dists, flags);
    case 1: return ws_postfilter_batch(idx, c.node, q, w, nq, c.qp, c.pad, ids, dists, flags);
    default: return ws_tree_batch(idx, c.method, q, w, nq, c.qp, ids, dists, flags);
  }
}

static int ws_group_batch(ws_group* g, const WsGroupCall& c, const float* queries, const float* windows, uint64_t nq,
                          uint32_t* ids, float* dists) {
  if (!g) return ws_fail(WS_ERR_BADARG, "null group");
  std::lock_guard<std::mutex> call_lock(g->call_mu);
  if (nq == 0) return WS_OK;
  if (!queries || !windows || !ids || !dists) return ws_fail(WS_ERR_BADARG, "null buffer");
  const int G = (int)g->m.size();
  const uint32_t dim = g->m[0]->dim;
  const uint32_t k = c.k;
  if (g->mode == WS_GROUP_REPLICATED) {
    // contiguous slices, sizes differ by at most one query
    return ws_group_run(g, [=](int i) {
      const uint64_t base = nq / G, rem = nq % G;
      const uint64_t lo = (uint64_t)i * base + std::min<uint64_t>(i, rem);
      const uint64_t cnt = base + ((uint64_t)i < rem ? 1 : 0);
      if (cnt == 0) return (int)WS_OK;
      // beamSearch.h:128 compares neighbour ids with the query's position in the batch: keep the caller's numbering
      std::lock_guard<std::recursive_mutex> lock(g->m[i]->mu);
      g->m[i]->query_id_base = (uint32_t)lo;
      const int rc = ws_group_member_call(g->m[i], c, queries + lo * dim, windows + 2 * lo, cnt, ids + lo * k, dists + lo * k, 0);
      g->m[i]->query_id_base = 0;
      return rc;
    });
  }
  // ---- label-sharded: every member answers the whole batch on its shard
  if (k == 0 || k > kMaxK) return ws_fail(WS_ERR_BADARG, "k=%u outside 1..%u", k, kMaxK);
  if (nq > (1ull << 24)) return ws_fail(WS_ERR_BADARG, "nq=%llu above 2^24 per batch", (unsigned long long)nq);
  const bool use_nccl = g->opt_exchange == 1 || !g->peer_ok;
  if (use_nccl) WS_TRY(ws_group_ensure_comms(g));
  const size_t rows_bytes = (size_t)nq * k * sizeof(uint32_t);
  // 1. upload + local search (device-pointer batches: only enqueued on the member's stream)
  WS_TRY(ws_group_run(g, [=](int i) {
    ws_index* idx = g->m[i];
    std::lock_guard<std::recursive_mutex> lock(idx->mu);
    WS_CUDA(cudaSetDevice(idx->device));
    WS_TRY(ws_ensure(idx, g->part_ids[i], rows_bytes));
    WS_TRY(ws_ensure(idx, g->part_dists[i], rows_bytes));
    WS_TRY(ws_ensure(idx, g->out_ids[i], rows_bytes));
    WS_TRY(ws_ensure(idx, g->out_dists[i], rows_bytes));
    WS_TRY(ws_ensure(idx, idx->d_queries, nq * dim * sizeof(float)));
    WS_TRY(ws_ensure(idx, idx->d_windows, nq * 2 * sizeof(float)));
    cudaStream_t st = idx->stream;
    WS_CUDA(cudaEventRecord(g->ev_t0[i], st));
    WS_CUDA(cudaMemcpyAsync(idx->d_queries.p, queries, nq * dim * sizeof(float), cudaMemcpyHostToDevice, st));
    WS_CUDA(cudaMemcpyAsync(idx->d_windows.p, windows, nq * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    idx->hint_windows = windows;  // the host copy: prefilter batches are routed by a sample of their window sizes
    const int rc = ws_group_member_call(idx, c, (const float*)idx->d_queries.p, (const float*)idx->d_windows.p, nq,
                                        (uint32_t*)g->part_ids[i].p, (float*)g->part_dists[i].p, WS_FLAG_DEVICE_PTRS);
    idx->hint_windows = nullptr;
    if (rc != WS_OK) return rc;
    WS_CUDA(cudaEventRecord(g->ev_done[i], st));
    WS_CUDA(cudaEventRecord(g->ev_t1[i], st));
    return (int)WS_OK;
  }));
  // 2. exchange + merge.  The pad id of a shard's rows never reaches the result (pads carry FLT_MAX and are
  //    re-created by the merge), so the caller's convention is applied here.
  const uint32_t pad_id = c.kind == 2 ? 0u : (c.kind == 1 && c.pad == WS_PAD_ZERO ? 0u : 0xFFFFFFFFu);
  WS_TRY(ws_group_run(g, [=](int i) {
    ws_index* idx = g->m[i];
    std::lock_guard<std::recursive_mutex> lock(idx->mu);
    WS_CUDA(cudaSetDevice(idx->device));
    cudaStream_t st = idx->stream;
    const uint64_t base = nq / G, rem = nq % G;
    const uint64_t lo = (uint64_t)i * base + std::min<uint64_t>(i, rem);
    const uint64_t cnt = base + ((uint64_t)i < rem ? 1 : 0);
    if (use_nccl) {
      // every member receives every rank's rows and merges its own slice of the queries
      const uint32_t parts = (uint32_t)G;
      WS_TRY(ws_comm_scratch(idx, parts, nq, k));
      WS_NCCL(g_nccl.GroupStart());
      WS_NCCL(g_nccl.AllGather(g->part_ids[i].p, idx->x_all_ids.p, (size_t)nq * k, ncclUint32, (ncclComm_t)idx->comm, st));
      WS_NCCL(g_nccl.AllGather(g->part_dists[i].p, idx->x_all_dists.p, (size_t)nq * k, ncclFloat32, (ncclComm_t)idx->comm, st));
      WS_NCCL(g_nccl.GroupEnd());
      idx->launches += 2;
    } else {
      for (int p = 0; p < G; p++)
        if (p != i) WS_CUDA(cudaStreamWaitEvent(st, g->ev_done[p], 0));
    }
    if (cnt > 0) {
      WsMergePartsArgs a{};
      if (use_nccl) {
        a.ids = (const uint32_t*)idx->x_all_ids.p; a.dists = (const float*)idx->x_all_dists.p;
      } else {
        a.ids = nullptr; a.dists = nullptr;
        for (int p = 0; p < G; p++) { a.part_ids[p] = (const uint32_t*)g->part_ids[p].p; a.part_dists[p] = (const float*)g->part_dists[p].p; }
      }
      a.parts = (uint32_t)G; a.k = k; a.pad_id = pad_id; a.nq_total = (uint32_t)nq; a.q0 = (uint32_t)lo; a.nq = (uint32_t)cnt;
      a.out_ids = (uint32_t*)g->out_ids[i].p; a.out_dists = (float*)g->out_dists[i].p;
      const int grid = (int)std::min<uint64_t>(cnt, (uint64_t)idx->num_sms * 16);
      WS_CUDA(wsl_merge_parts(grid, st, a));
      idx->launches++;
      WS_CUDA(cudaMemcpyAsync(ids + lo * k, (uint32_t*)g->out_ids[i].p + lo * k, cnt * k * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
      WS_CUDA(cudaMemcpyAsync(dists + lo * k, (float*)g->out_dists[i].p + lo * k, cnt * k * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    WS_CUDA(cudaEventRecord(g->ev_t2[i], st));
    uint32_t h_sticky = 0;
    WS_CUDA(cudaMemcpyAsync(&h_sticky, idx->d_sticky, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    WS_CUDA(cudaStreamSynchronize(st));
    if (h_sticky) return ws_report_sticky(idx, h_sticky);
    return (int)WS_OK;
  }));
  // peers may still be READING this member's partial rows when its own stream is done: a member must not start
  // the next batch (which overwrites them) before every merge has finished — all streams were synchronised above.
  double search = 0, exch = 0, total = 0;
  for (int i = 0; i < G; i++) {
    float a = 0, b = 0;
    cudaSetDevice(g->m[i]->device);
    cudaEventElapsedTime(&a, g->ev_t0[i], g->ev_t1[i]);
    cudaEventElapsedTime(&b, g->ev_t1[i], g->ev_t2[i]);
    search = std::max(search, (double)a); exch = std::max(exch, (double)b); total = std::max(total, (double)a + b);
  }
  g->last_ms[0] = search; g->last_ms[1] = exch; g->last_ms[2] = total;
  return WS_OK;
}

}  // namespace

extern "C" {

int ws_group_create(ws_index* const* members, int count, int mode, ws_group** out) {
  if (!out) return ws_fail(WS_ERR_BADARG, "out is null");
  *out = nullptr;
  if (!members || count < 1 || count > WS_MAX_PARTS) return ws_fail(WS_ERR_BADARG, "a group has 1..%d members", WS_MAX_PARTS);
  if (mode != WS_GROUP_REPLICATED && mode != WS_GROUP_LABEL_SHARDED) return ws_fail(WS_ERR_BADARG, "unknown group mode %d", mode);
  for (int i = 0; i < count; i++) {
    ws_index* a = members[i];
    if (!a || a->device < 0 || !a->finalized) return ws_fail(WS_ERR_STATE, "member %d is not a finalized device arena", i);
    if (a->dim != members[0]->dim || a->metric != members[0]->metric) return ws_fail(WS_ERR_BADARG, "member %d: dim / metric differ", i);
    if (mode == WS_GROUP_REPLICATED && a->n != members[0]->n) return ws_fail(WS_ERR_BADARG, "member %d is not a replica (n differs)", i);
  }
  ws_group* g = new ws_group();
  g->mode = mode;
  g->m.assign(members, members + count);
  g->status.assign(count, WS_OK);
  g->errors.assign(count, std::string());
  g->part_ids.resize(count); g->part_dists.resize(count); g->out_ids.resize(count); g->out_dists.resize(count);
  g->ev_done.assign(count, nullptr); g->ev_t0.assign(count, nullptr); g->ev_t1.assign(count, nullptr); g->ev_t2.assign(count, nullptr);
  // peer access between every pair of member devices (label-sharded exchange by peer loads)
  g->peer_ok = true;
  for (int i = 0; i < count; i++) {
    cudaSetDevice(g->m[i]->device);
    cudaEventCreateWithFlags(&g->ev_done[i], cudaEventDisableTiming);
    cudaEventCreate(&g->ev_t0[i]); cudaEventCreate(&g->ev_t1[i]); cudaEventCreate(&g->ev_t2[i]);
    for (int j = 0; j < count; j++) {
      if (g->m[j]->device == g->m[i]->device) continue;
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, g->m[i]->device, g->m[j]->device) != cudaSuccess || !can) { g->peer_ok = false; continue; }
      cudaError_t e = cudaDeviceEnablePeerAccess(g->m[j]->device, 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
      else if (e != cudaSuccess) { cudaGetLastError(); g->peer_ok = false; }
    }
  }
  for (int i = 0; i < count; i++) g->workers.emplace_back(ws_group_worker, g, i);
  *out = g;
  return WS_OK;
}

void ws_group_destroy(ws_group* g) {
  if (!g) return;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->stop = true;
  }
  g->cv_job.notify_all();
  for (std::thread& t : g->workers) t.join();
  for (size_t i = 0; i < g->m.size(); i++) {
    ws_index* idx = g->m[i];
    cudaSetDevice(idx->device);
    cudaStreamSynchronize(idx->stream);
    if (g->comms_ready) ws_index_comm_destroy(idx);
    WsDevBuf* bufs[] = {&g->part_ids[i], &g->part_dists[i], &g->out_ids[i], &g->out_dists[i]};
    for (WsDevBuf* b : bufs) {
      if (b->p) { cudaFree(b->p); idx->hbm_bytes -= b->bytes; }
    }
    cudaEventDestroy(g->ev_done[i]); cudaEventDestroy(g->ev_t0[i]); cudaEventDestroy(g->ev_t1[i]); cudaEventDestroy(g->ev_t2[i]);
  }
  delete g;
}

int ws_group_size(const ws_group* g, int* count) {
  if (!g || !count) return ws_fail(WS_ERR_BADARG, "null argument");
  *count = (int)g->m.size();
  return WS_OK;
}

int ws_group_member(const ws_group* g, int i, ws_index** out) {
  if (!g || !out || i < 0 || (size_t)i >= g->m.size()) return ws_fail(WS_ERR_BADARG, "bad member %d", i);
  *out = g->m[i];
  return WS_OK;
}

int ws_group_set_option(ws_group* g, const char* name, int64_t value) {
  if (!g || !name) return ws_fail(WS_ERR_BADARG, "null argument");
  std::lock_guard<std::mutex> call_lock(g->call_mu);
  if (std::string(name) == "exchange") {
    if (value != 0 && value != 1) return ws_fail(WS_ERR_BADARG, "exchange: 0 = peer loads, 1 = ncclAllGather");
    g->opt_exchange = value;
    return WS_OK;
  }
  for (ws_index* idx : g->m) WS_TRY(ws_index_set_option(idx, name, value));  // forwarded to every member
  return WS_OK;
}

int ws_group_info(ws_group* g, int* peer_access, int* exchange_in_use, double* last_ms3) {
  if (!g) return ws_fail(WS_ERR_BADARG, "null group");
  if (peer_access) *peer_access = g->peer_ok ? 1 : 0;
  if (exchange_in_use) *exchange_in_use = (g->opt_exchange == 1 || !g->peer_ok) ? 1 : 0;
  if (last_ms3) { last_ms3[0] = g->last_ms[0]; last_ms3[1] = g->last_ms[1]; last_ms3[2] = g->last_ms[2]; }
  return WS_OK;
}

int ws_group_prefilter_batch(ws_group* g, const float* queries, const float* windows, uint64_t nq, uint32_t k, uint32_t* ids,
                             float* dists) {
  WsGroupCall c{0, 0, -1, WS_PAD_MINUS1, k, nullptr};
  return ws_group_batch(g, c, queries, windows, nq, ids, dists);
}

int ws_group_postfilter_batch(ws_group* g, int32_t node, const float* queries, const float* windows, uint64_t nq,
                              const ws_query_params* qp, int pad, uint32_t* ids, float* dists) {
  if (!qp) return ws_fail(WS_ERR_BADARG, "null query params");
  if (g && g->mode == WS_GROUP_LABEL_SHARDED) return ws_fail(WS_ERR_BADARG, "a flat postfilter index cannot be label-sharded");
  WsGroupCall c{1, 0, node, pad, (uint32_t)qp->k, qp};
  return ws_group_batch(g, c, queries, windows, nq, ids, dists);
}

int ws_group_tree_batch(ws_group* g, int method, const float* queries, const float* windows, uint64_t nq,
                        const ws_query_params* qp, uint32_t* ids, float* dists) {
  if (!qp) return ws_fail(WS_ERR_BADARG, "null query params");
  WsGroupCall c{2, method, -1, WS_PAD_ZERO, (uint32_t)qp->k, qp};
  return ws_group_batch(g, c, queries, windows, nq, ids, dists);
}

}  // extern "C"
