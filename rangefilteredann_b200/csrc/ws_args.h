// ws_args.h — plain argument structs of the sm_100a kernels (ws_kernels.cuh, ws_build.cuh).  Kept apart from
// the kernels so that every translation unit can include the kernels with internal linkage (anonymous namespace)
// while the launchers declared in ws_launch.h share these types across translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ws_decompose.h"

// launch geometry shared by hosts and kernels
#define WS_TEAM 8             // lanes cooperating on one vector (128-bit loads, 128 B per team pass)
#define WS_CTA_THREADS 128    // CTA-per-task scan / merge / build kernels: 4 warps = 16 teams
#define WS_WARPS_PER_CTA 4    // warp-per-task kernels
#define WS_CTA2_THREADS 256   // CTA-per-task beam kernel of the large tier
#define WS_TOPK_BUF 2048      // streaming top-k buffer of the CTA scan / merge kernels (k <= WS_TOPK_BUF / 2)

// shared memory one warp of the warp-per-task beam kernel needs (bytes)
static inline __host__ __device__ size_t ws_warp_smem_bytes(uint32_t cap, uint32_t hash_entries, uint32_t hash16) {
  return (size_t)cap * 8 + 64 * 8 * 2 + 64 * 4 * 2 + (size_t)hash_entries * (hash16 ? 2 : 4);
}

struct WsNode {
  const int32_t* adj;  // [count][R] local neighbour ids, -1 padded
  uint32_t start;      // first arena rank of the node
  uint32_t count;
};

enum { WS_MODE_PREFILTER = 10, WS_MODE_POSTFILTER = 11 };

// stats slots (unsigned long long[8]) — order of ws_stats
enum { WS_ST_SEARCHES = 0, WS_ST_VISITED, WS_ST_DISTCMPS, WS_ST_SCANPTS, WS_ST_GTASKS, WS_ST_STASKS,
       WS_ST_ESCALATED, WS_ST_BEAMSUM };

struct WsDecompArgs {
  WsGeom g;
  WsDecompParams p;
  int mode;
  int32_t node;            // WS_MODE_POSTFILTER
  const float* windows;    // [nq][2]
  uint32_t nq;
  uint32_t cap;
  WsTask* tasks;           // [nq][cap]
  uint32_t* counts;        // [nq]
  uint32_t* gq;            // graph-task queue (slot indices)
  uint32_t* gq_count;
  uint32_t* sq;            // scan-task queue
  uint32_t* sq_count;
  uint32_t* overflow;
  unsigned long long* stats;
};

struct WsBeamArgs {
  const float* vecs;       // [n][dpad]
  const float* labels;     // [n]
  const WsNode* nodes;
  const float* queries;    // [nq][dim]
  uint32_t dim, dpad, R;
  WsTask* tasks;
  uint64_t* res_keys;      // [slots][k]
  uint32_t* res_cnt;       // [slots]
  uint32_t k;
  const uint32_t* q_in;
  const uint32_t* q_in_count;
  uint32_t* q_head;
  uint32_t* q_out;         // next tier (may be null on the last tier)
  uint32_t* q_out_count;
  uint32_t beam_cap;       // largest beam this launch can hold in shared memory
  uint32_t hash_mask;      // smem visited table entries - 1 (GLOBAL_SEEN == false)
  uint32_t cand_cap;       // power of two >= expand * R
  uint32_t expand;         // nodes expanded per step (1 = reference order)
  int32_t skip_query_id;   // emulate `a == p.id()` (beamSearch.h:128)
  uint32_t query_id_base;  // p.id() of this launch's first query: a query's position in the CALLER's batch (a group
                           // hands each member a slice of the batch)
  long long max_beam, final_mult, limit, degree_limit;
  uint32_t* bitmap;        // GLOBAL_SEEN: [gridDim.x][bitmap_words]
  uint64_t bitmap_words;
  unsigned long long* stats;
  uint32_t* sticky;        // index-wide error word (bit 1: a task outgrew the last tier and was dropped)
  uint32_t hash16;         // warp tiers: visited table holds 16-bit tags (ws_seen_warp2_h16)
  uint32_t min_tasks;      // warp tiers fed by escalation: below this many queued tasks, hand them all to q_out
  // optional: brute-force scan tasks of the same batch, drained by the same warps once the graph
  // queue is empty (warp tiers only; null otherwise)
  const uint32_t* sq_in;
  const uint32_t* sq_count;
  uint32_t* sq_head;
  // final result rows (written directly for WS_TF_SOLO tasks; K4 handles the rest)
  uint32_t* out_ids;
  float* out_dists;
  const uint32_t* decode;
  uint32_t pad_id;
};

struct WsScanArgs {
  const float* vecs;
  const float* queries;
  uint32_t dim, dpad;
  const WsTask* tasks;
  uint64_t* res_keys;
  uint32_t* res_cnt;
  uint32_t k;
  const uint32_t* q_in;
  const uint32_t* q_in_count;
  uint32_t* q_head;
  unsigned long long* stats;
  uint32_t* out_ids;
  float* out_dists;
  const uint32_t* decode;
  uint32_t pad_id;
};

struct WsMergeArgs {
  const uint32_t* counts;   // tasks per query
  uint32_t cap;
  const uint64_t* res_keys;
  const uint32_t* res_cnt;
  uint32_t k;
  const uint32_t* decode;   // may be null
  uint32_t pad_id;
  uint32_t nq;
  uint32_t* ids;            // [nq][k]
  float* dists;             // [nq][k]
};

struct WsPrefilterDirectArgs {
  WsScanArgs s;            // tasks / q_in* unused
  const float* labels;     // [n] sorted
  uint64_t n;
  const float* windows;    // [nq][2]
  uint32_t nq;
};

#define WS_MAX_PARTS 16
struct WsMergePartsArgs {
  // either one gathered buffer [parts][nq_total][k] (the all-gather layout) ...
  const uint32_t* ids;
  const float* dists;
  // ... or, with ids == nullptr, one [nq_total][k] buffer per part, each possibly in a PEER device's memory (the
  // kernel then gathers over NVLink with plain loads while it merges: no staging copy, no collective launch)
  const uint32_t* part_ids[WS_MAX_PARTS];
  const float* part_dists[WS_MAX_PARTS];
  uint32_t parts, k, pad_id;
  uint32_t nq_total;      // rows per part
  uint32_t q0, nq;        // this launch merges rows [q0, q0 + nq)
  uint32_t* out_ids;      // [nq_total][k]; rows [q0, q0 + nq) are written
  float* out_dists;
};

#define WS_BUILD_VCAP 2048  // candidate capacity of one prune (visited list + current neighbours)

struct WsBuildGraph {
  uint32_t start;     // first arena rank
  uint32_t count;     // points in the graph
  uint32_t row_off;   // first row in the build adjacency
  uint32_t floor;     // this round inserts perm[floor .. ceil)
  uint32_t ceil;
  uint32_t task_off;  // first task index of this graph in the round
};

struct WsBuildArgs {
  const float* vecs;
  uint32_t dim, dpad, R;
  int32_t* adj;            // [rows][R]
  int32_t* deg;            // [rows]
  const int32_t* perm;     // [rows] local ids in insertion order
  const WsBuildGraph* graphs;
  uint32_t ngraphs;
  uint32_t ntasks;
  uint32_t* head;          // task counter
  int32_t* new_out;        // [ntasks][R]
  int32_t* new_cnt;        // [ntasks]
  uint64_t* pairs;         // reverse edges: (target row << 32) | source local id
  uint32_t* pair_count;
  uint32_t L;              // build beam
  uint32_t beam_cap, hash_mask, cand_cap, expand;
  double alpha;
  unsigned long long* stats;
};

struct WsBuildRevArgs {
  const float* vecs;
  uint32_t dim, dpad, R;
  int32_t* adj;
  int32_t* deg;
  const WsBuildGraph* graphs;
  uint32_t ngraphs;
  const uint64_t* pairs;
  uint32_t npairs;
  const uint32_t* heads;
  uint32_t nheads;
  uint32_t* head;  // work counter
  double alpha;
  unsigned long long* stats;
};

struct WsBuildSortArgs {
  const float* vecs;
  uint32_t dim, dpad, R;
  int32_t* adj;
  const int32_t* deg;
  const WsBuildGraph* graphs;
  uint32_t ngraphs;
  uint32_t rows;
};
