/*
 * wsann.h — C ABI of libwsann_cuda.so, the sm_100a window-search engine.
 *
 * This is the LOWER face of the drop-in boundary.  The reference
 * (JoshEngels/RangeFilteredANN) has no FFI layer of its own: its boundary is the
 * pybind11 module `window_ann` (python_bindings/python_bindings.cpp:160-238) whose
 * classes call header-only C++ (`batch_search` in the headers under src/).  Every entry point below
 * names the reference function(s) whose work it replaces.  The host-side C++ index
 * classes in rangefilteredann_b200/csrc/host/ (same class names / constructor /
 * batch_search signatures as the reference) are the only intended callers; the
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C types only; no C++/torch types cross this line
 *   - every call returns ws_status (0 = ok, <0 = error); ws_last_error() gives text
 *   - host buffers are caller-owned and only read/written during the call
 *   - there is NO CPU fallback: without a usable CUDA device every call that
 *     needs one returns WS_ERR_CUDA
 *   - "arena order" = the order in which vectors were handed to ws_index_create
 *     (label-sorted for the tree / prefilter classes, original order for the
 *     standalone postfilter class)
 */
#ifndef WSANN_H_
#define WSANN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSANN_ABI_VERSION 1

typedef enum ws_status {
  WS_OK = 0,
  WS_ERR_BADARG = -1,
  WS_ERR_CUDA = -2,
  WS_ERR_OOM = -3,
  WS_ERR_NCCL = -4,
  WS_ERR_STATE = -5
} ws_status;

/* Euclidian_Point<float>::distance (ParlayANN/algorithms/utils/euclidian_point.h:62-65,
 * NSGDist.h:31-70) -> squared L2;  Mips_Point<float>::distance (mips_point.h:60-66) -> -dot */
typedef enum ws_metric { WS_METRIC_L2 = 0, WS_METRIC_MIPS = 1 } ws_metric;

/* tree-level query methods (range_filter_tree.h:76-82, super_optimized_postfilter_tree.h:187) */
typedef enum ws_method {
  WS_METHOD_FENWICK = 0,          /* fenwick_tree_search            range_filter_tree.h:297-401 */
  WS_METHOD_OPT_POSTFILTER = 1,   /* optimized_postfiltering_search range_filter_tree.h:403-471 */
  WS_METHOD_THREE_SPLIT = 2,      /* three_split_search             range_filter_tree.h:473-540 */
  WS_METHOD_SUPER_POSTFILTER = 3  /* super_optimized_postfiltering_search  super_optimized_postfilter_tree.h:187-270 */
} ws_method;

/* what to write into result slots that have no neighbour */
typedef enum ws_pad {
  WS_PAD_ZERO = 0,     /* tree classes: id 0, FLT_MAX        range_filter_tree.h:89-92 */
  WS_PAD_MINUS1 = 1    /* postfilter / prefilter classes: id 0xFFFFFFFF, FLT_MAX   postfilter_vamana.h:211-214 */
} ws_pad;

/* flags for the *_batch calls */
#define WS_FLAG_DEVICE_PTRS 1u /* queries/windows/ids/dists are device pointers on the index's device;
                                  the call only enqueues work on the index stream (use ws_index_sync) */

/* POD mirror of QueryParams (ParlayANN/algorithms/utils/types.h:115-140) */
typedef struct ws_query_params {
  int64_t k;
  int64_t beam_size;
  double cut;                       /* dead on this path; carried for ABI parity */
  int64_t limit;                    /* idem */
  int64_t degree_limit;             /* idem */
  int64_t final_beam_multiply;
  int64_t postfiltering_max_beam;
  float min_query_to_bucket_ratio;  /* valid iff has_min_query_to_bucket_ratio */
  int32_t has_min_query_to_bucket_ratio;
  int32_t verbose;
} ws_query_params;

/* device-side counters, accumulated since the last ws_index_reset_stats
 * (counted the way beamSearch.h:94,117,141 count them) */
typedef struct ws_stats {
  uint64_t graph_searches;     /* beam_search invocations (one per doubling round)      */
  uint64_t visited;            /* expanded nodes  (num_visited, beamSearch.h:117)        */
  uint64_t dist_cmps;          /* distance evaluations in graph search (beamSearch.h:141)*/
  uint64_t scan_points;        /* brute-force distance evaluations (prefiltering.h:189-194, range_filter_tree.h:386-397) */
  uint64_t graph_tasks;        /* (query,node) sub-index queries dispatched              */
  uint64_t scan_tasks;         /* (query,slice) brute-force tasks dispatched             */
  uint64_t escalated_tasks;    /* tasks that left the first beam tier                    */
  uint64_t beam_sum;           /* sum of beamSize over graph_searches (frontier bytes) */
} ws_stats;

typedef struct ws_index ws_index; /* opaque: one HBM arena on one device (several of them form a ws_group) */

const char* ws_last_error(void);
int ws_abi_version(void);
int ws_device_count(int* count);

/* ---- arena construction -------------------------------------------------------------
 * Replaces PointRange / SubsetPointRange storage (point_range.h:49-202) and the per-class
 * copies of labels / decode tables (tree_utils.h:39-98).  `vectors` is [n][dim] fp32 in
 * arena order and is copied into HBM with the reference's 64-byte row rule
 * (point_range.h:39-44), pad zero-filled.  `labels` is [n] in arena order.  `decode`
 * ([n], arena rank -> original id) may be NULL (identity).  `label_sorted` != 0 promises
 * labels are non-decreasing (required by the prefilter / tree calls). */
int ws_index_create(int device, int metric, uint64_t n, uint32_t dim, const float* vectors,
                    const float* labels, const uint32_t* decode, int label_sorted,
                    ws_index** out);
void ws_index_destroy(ws_index* idx);
/* Replace the id table (arena rank -> the id written into result rows): a label shard built over a slice of a data
 * set re-labels its rows with data-set-wide ids, so that rows of different shards can be merged. */
int ws_index_set_decode(ws_index* idx, const uint32_t* decode);

/* One Vamana graph over arena ranks [start, start+count) — a PostfilterVamanaIndex's
 * Graph<int32> (postfilter_vamana.h:35,54-79; graph.h:115-206).  `degrees`[count] and
 * `edges` (concatenated rows, local ids) are exactly the payload of the reference's
 * .bin file (graph.h:174-196).  Returns the node handle in *node_out. */
int ws_index_add_graph(ws_index* idx, uint64_t start, uint64_t count, uint32_t max_degree,
                       const int32_t* degrees, const int32_t* edges, int32_t* node_out);

/* B-WST geometry (range_filter_tree.h:129-189): `rows` rows, row r has row_nb[r] buckets,
 * offsets_flat holds the concatenated per-row offset arrays (row_nb[r]+1 entries each),
 * node_ids_flat the node handle of every bucket (row-major).  node_ids_flat == NULL selects the
 * reference's other instantiation, RangeFilterTreeIndex<T, Point, PrefilterIndex>
 * (range_filter_tree.h:32, python_bindings.cpp:119-127): a bucket query is then PrefilterIndex::query_knn
 * (prefiltering.h:154-204) over the bucket's own slice, i.e. a brute-force scan, and no graph is needed. */
int ws_index_set_wst(ws_index* idx, uint32_t rows, uint32_t split_factor, int32_t cutoff,
                     const uint32_t* row_nb, const uint64_t* offsets_flat,
                     const int32_t* node_ids_flat);

/* Super-postfilter tree geometry (super_optimized_postfilter_tree.h:118-171). */
int ws_index_set_super(ws_index* idx, uint32_t rows, int32_t cutoff, const uint64_t* bucket_sizes,
                       const uint64_t* bucket_shifts, const uint32_t* row_nb,
                       const int32_t* node_ids_flat);

/* ---- graph construction on the device (setup path) ---------------------
 * Builds one Vamana graph per [starts[i], starts[i]+counts[i]) range with the reference
 * builder's algorithm (ParlayANN/algorithms/vamana/index.h:61-108,123-135,211-313: batch
 * insertion with prefix doubling, beam search L, robustPrune alpha, reverse edges, final
 * distance sort), all graphs in lock step, and registers them as nodes of the index.
 * Used when a node's cache file (postfilter_vamana.h:54-61) is missing; the host then
 * saves the result in the reference's .bin format so both implementations load the same
 * graph.  The insertion order comes from `seed`, so graphs are valid Vamana graphs but not
 * bit-identical to a reference-builder run. */
int ws_build_graphs(ws_index* idx, uint32_t ngraphs, const uint64_t* starts, const uint64_t* counts,
                    uint32_t max_degree, uint32_t beam_l, double alpha, uint64_t seed, int32_t* nodes_out);
/* degrees[count] and rows[count][R] (-1 padded, R = *max_degree_out) of one node */
int ws_index_get_graph(ws_index* idx, int32_t node, uint32_t* max_degree_out, int32_t* degrees, int32_t* rows);
/* builder counters: points inserted, nodes expanded, distance evaluations, overflow re-prunes */
int ws_index_build_stats(ws_index* idx, uint64_t* out4);

/* Upload everything staged so far; must be called once before any *_batch call. */
int ws_index_finalize(ws_index* idx);

/* ---- queries ------------------------------------------------------------------------
 * queries [nq][dim] fp32, windows [nq][2] fp32 (lo,hi), ids [nq][k] uint32, dists [nq][k]
 * fp32.  Rows come back sorted ascending by (distance, id). */

/* PrefilterIndex::batch_search (src/prefiltering.h:124-146, query_knn :154-204).
 * Window = [lb(lo), lb(hi)) with the reference's r = n-1 quirk. */
int ws_prefilter_batch(ws_index* idx, const float* queries, const float* windows, uint64_t nq,
                       uint32_t k, uint32_t* ids, float* dists, uint32_t flags);

/* PostfilterVamanaIndex::batch_search on one graph (src/postfilter_vamana.h:191-219;
 * query :141-188, raw_query :223-254, beam_search beamSearch.h:51-184).  Ids are decoded
 * through the arena's decode table; missing slots padded per `pad`. */
int ws_postfilter_batch(ws_index* idx, int32_t node, const float* queries, const float* windows,
                        uint64_t nq, const ws_query_params* qp, int pad, uint32_t* ids,
                        float* dists, uint32_t flags);

/* RangeFilterTreeIndex<…,PostfilterVamanaIndex>::batch_search (range_filter_tree.h:62-96)
 * and SuperOptimizedPostfilterTree::batch_search (super_optimized_postfilter_tree.h:60-87).
 * Window→node decomposition, the per-node searches, edge scans, merge (sort_and_truncate,
 * range_filter_tree.h:542-549) and the sorted→original decode all run on the device. */
int ws_tree_batch(ws_index* idx, int method, const float* queries, const float* windows,
                  uint64_t nq, const ws_query_params* qp, uint32_t* ids, float* dists,
                  uint32_t flags);

/* Label-range sharded mode (datasets larger than one GPU's HBM): every rank
 * owns a contiguous label range with its own sub-tree, answers the whole batch locally, the
 * per-rank [nq][k] rows are all-gathered (ws_allgather_merge / ws_group below do both steps; this entry point
 * takes rows the host plumbing gathered by other means) and this kernel merges the `parts` lists per query — sort_and_truncate (range_filter_tree.h:542-549)
 * across shards.  All pointers are DEVICE pointers; rows with dist == FLT_MAX are pads.
 * Enqueues on the index stream. */
int ws_merge_partial_topk(ws_index* idx, const uint32_t* ids, const float* dists, uint32_t parts, uint64_t nq,
                          uint32_t k, uint32_t pad_id, uint32_t* out_ids, float* out_dists);

/* ---- multi-GPU inside the engine ------------------------------------------------------------
 * The reference spreads one batch over the host's cores with `parlay::parallel_for` over the queries
 * (range_filter_tree.h:70, prefiltering.h:131, postfilter_vamana.h:199, super_optimized_postfilter_tree.h:66).
 * A ws_group spreads one batch over the GPUs of the box behind the same call (SURVEY.md §8b `ws_ctx`, §8e):
 *
 *   WS_GROUP_REPLICATED     every member holds the whole arena; the batch is cut into one contiguous slice per
 *                           member (queries are independent units: no data-path collective)
 *   WS_GROUP_LABEL_SHARDED  member g holds a contiguous range of the label-sorted points and its own tree; every
 *                           member answers the whole batch on its shard, the [nq][k] partial rows are exchanged
 *                           and merged per query (sort_and_truncate, range_filter_tree.h:542-549, across shards).
 *                           Option "exchange": 0 = each member's merge kernel reads its peers' rows straight from
 *                           their HBM over NVLink (peer access; default when available), 1 = ncclAllGather + merge.
 *
 * Host buffers only; the call returns when ids / dists are filled.  Members must be finalized and are not owned
 * by the group.  Any other option name given to ws_group_set_option is forwarded to every member. */
typedef struct ws_group ws_group;
typedef enum ws_group_mode { WS_GROUP_REPLICATED = 0, WS_GROUP_LABEL_SHARDED = 1 } ws_group_mode;

/* Clone of a finalized arena on another device (vectors, labels, decode, adjacency copied device to device). */
int ws_index_replicate(ws_index* src, int device, ws_index** out);

int ws_group_create(ws_index* const* members, int count, int mode, ws_group** out);
void ws_group_destroy(ws_group* g);
int ws_group_size(const ws_group* g, int* count);
int ws_group_member(const ws_group* g, int i, ws_index** out);
int ws_group_set_option(ws_group* g, const char* name, int64_t value);
/* peer_access: all member pairs can read each other's HBM; exchange_in_use: 0 peer loads / 1 NCCL;
 * last_ms3: device times of the last label-sharded batch, max over members: {search, exchange + merge, total} */
int ws_group_info(ws_group* g, int* peer_access, int* exchange_in_use, double* last_ms3);
/* PrefilterIndex::batch_search / PostfilterVamanaIndex::batch_search / RangeFilterTreeIndex::batch_search /
 * SuperOptimizedPostfilterTree::batch_search over all members (same arguments as the ws_*_batch calls) */
int ws_group_prefilter_batch(ws_group* g, const float* queries, const float* windows, uint64_t nq, uint32_t k,
                             uint32_t* ids, float* dists);
int ws_group_postfilter_batch(ws_group* g, int32_t node, const float* queries, const float* windows, uint64_t nq,
                              const ws_query_params* qp, int pad, uint32_t* ids, float* dists);
int ws_group_tree_batch(ws_group* g, int method, const float* queries, const float* windows, uint64_t nq,
                        const ws_query_params* qp, uint32_t* ids, float* dists);

/* One process per GPU (torchrun-style launch): the label-sharded exchange with NCCL inside this library.
 * Rank 0 makes an id (ncclGetUniqueId) and hands the 128 bytes to the other ranks by whatever means the host
 * program has; every rank then binds a communicator to its arena.  ws_allgather_merge all-gathers this rank's
 * [nq][k] rows (DEVICE pointers, global ids, pads carry FLT_MAX) on the index stream and merges the ranks' lists
 * per query into out_ids / out_dists (device pointers).  NCCL is loaded with dlopen("libnccl.so.2") on first use. */
#define WS_NCCL_ID_BYTES 128
int ws_nccl_unique_id(void* id_out);
int ws_nccl_version(int* version);
int ws_index_comm_init(ws_index* idx, int nranks, int rank, const void* id);
int ws_index_comm_destroy(ws_index* idx);
int ws_allgather_merge(ws_index* idx, const uint32_t* ids, const float* dists, uint64_t nq, uint32_t k, uint32_t pad_id,
                       uint32_t* out_ids, float* out_dists);

/* ---- arena snapshot ----------------------------------------------------------
 * The reference persists only graphs (postfilter_vamana.h:54-79: one .bin per node) and re-sorts, re-derives and
 * re-reads everything at every start.  ws_index_save writes the FINISHED arena (padded vectors, labels, id table,
 * node table, adjacency rows, tree geometry — not scratch, options or counters) into one file; ws_index_load
 * brings it back on any device, finalized and ready for the *_batch calls.  Layout: csrc/ws_snapshot.inl. */
int ws_index_save(ws_index* idx, const char* path);
int ws_index_load(const char* path, int device, ws_index** out);
/* n, dim, metric and the number of graphs of an arena (any of the out pointers may be NULL) */
int ws_index_shape(const ws_index* idx, uint64_t* n, uint32_t* dim, int* metric, uint64_t* nodes);

/* ---- device plumbing for callers that keep batches resident in HBM (bench `value`) --- */
int ws_index_device(const ws_index* idx, int* device);
int ws_index_sync(ws_index* idx);
int ws_device_alloc(ws_index* idx, size_t bytes, void** dptr);
int ws_device_free(ws_index* idx, void* dptr);
int ws_host_alloc_pinned(size_t bytes, void** hptr);
int ws_host_free_pinned(void* hptr);
int ws_copy_h2d(ws_index* idx, void* dst, const void* src, size_t bytes);
int ws_copy_d2h(ws_index* idx, void* dst, const void* src, size_t bytes);
/* CUDA-event timer on the index stream (the stream every kernel of this index runs on) */
int ws_timer_start(ws_index* idx);
int ws_timer_stop(ws_index* idx, float* elapsed_ms);
/* writes `bytes` of zeros into an internal scratch buffer (> L2) to evict L2 between timed steps */
int ws_flush_l2(ws_index* idx);

/* ---- introspection ------------------------------------------------------------------ */
int ws_index_get_stats(ws_index* idx, ws_stats* out);
int ws_index_reset_stats(ws_index* idx);
/* number of kernels launched by this index since creation (bench `gpu_launches`) */
int ws_index_launch_count(const ws_index* idx, uint64_t* out);
/* tuning knobs: "emulate_query_id_skip" (beamSearch.h:128 `a == p.id()`, default 1),
 * "scan_chunk" (rows per brute-force task), "profile_kernels", "warp_tiers", "warp_hash",
 * "prefilter_open_tail" (1: prefilter windows may include the arena's last point — every label shard but the last;
 * the reference's r = n-1 rule, SURVEY.md §A-2, then applies to the data set's last point only),
 * "prefilter_direct" (one-launch prefilter kernel for batches of small windows: 0 never, 1 always,
 * 2 auto = host-sampled mean window <= scan_chunk),
 * "gemm_prefilter" (tensor-core prefilter: 0 never, 1 whenever eligible, 2 auto = host-sampled mean
 * window >= "gemm_min_window"), "gemm_items", "gemm_min_tiles", "gemm_chunk_mb",
 * "warp_scan", "fuse_scan", "hash_factor", "build_expand_width" (nodes expanded per step by
 * the device-side graph BUILDER; the query kernels always expand one node per step, as the
 * reference does) */
int ws_index_set_option(ws_index* idx, const char* name, int64_t value);
int ws_index_hbm_bytes(const ws_index* idx, uint64_t* out);
/* With option "profile_kernels" = 1 every kernel launch is bracketed by CUDA events on the
 * index stream.  Returns accumulated milliseconds / launch counts per kernel kind:
 * 0 decompose, 1..5 beam search warp-per-task tiers (beam <= 64 / 128 / 256 / 512 / 1024), 6 beam search
 * CTA-per-task large tier (<= 12288, global visited bitmap), 7 scan, 8 merge, 9 tensor-core prefilter
 * sweep, 10 its bounds + plan + pack kernels, 11 its re-rank kernel, 12 its seed-threshold kernel.
 * ms_out / launches_out are [13]. */
int ws_index_kernel_times(ws_index* idx, double* ms_out, uint64_t* launches_out, int reset);

/* Host-side evaluation of the window→task decomposition, same code the device runs
 * (testing hook: lets CPU-only CI check the tree logic; performs no search).
 * out_tasks: [nq][cap][4] int64 = (node or -1, start, end, flags); out_counts [nq]. */
int ws_debug_decompose_host(ws_index* idx, int method, const float* windows, uint64_t nq,
                            const ws_query_params* qp, uint32_t cap, int64_t* out_tasks,
                            uint32_t* out_counts);
int ws_index_task_capacity(ws_index* idx, int method, uint32_t* cap);
/* Host-only exercise of the pinned-staging helper threads (no CUDA; testing hook): `reps` copies of `bytes` bytes
 * through the pool, each compared with its source. */
int ws_debug_copy_pool_selftest(uint64_t bytes, uint32_t reps);

#ifdef __cplusplus
}
#endif
#endif /* WSANN_H_ */
