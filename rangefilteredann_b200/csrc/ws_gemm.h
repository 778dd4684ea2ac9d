// ws_gemm.h — argument structs and launchers of the tensor-core prefilter (kernels: ws_gemm.cu).
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>

#define WSG_TILE_M 128       // queries per group = TMEM lanes
#define WSG_TILE_N 128       // points per tile   = TMEM columns per accumulator stage
#define WSG_KBLK 64          // fp16 columns per 128-byte swizzle block
#define WSG_KBLK_BYTES (WSG_TILE_N * 128)  // one operand block: 128 rows x 128 B = 16 KB
#define WSG_MAX_KB 8         // dpad <= 512
#define WSG_B_STAGES 12      // ring of point blocks (192 KB)
#define WSG_ACC_STAGES 3     // accumulator stages of 128 TMEM columns: 3 next to a query operand of <= 128 columns
                             // (dpad <= 256), 2 next to one of 256 columns (dpad <= 512)
#define WSG_KTOP 16          // k <= 16 on this path
#define WSG_CAND_CAP 512     // survivors kept per (item, query)
#define WSG_SCHED 64         // per-CTA ring of item ids (see WsGemmSmem::sched)
#define WSG_RING 16          // per-query ring of fresh scores waiting to be folded into the running top-k
#define WSG_EPI_WARPS 16     // 4 per TMEM lane quarter: each takes 32 of a tile's 128 columns
#define WSG_THREADS 704      // warp 0: TMA producer, warp 1: MMA issuer, warps 2-17: epilogue, 18-21: thresholds
#define WSG_PLAN_THREADS 1024
#define WSG_MAX_ROWS 16384   // queries per plan (one CTA sorts them in shared memory)
#define WSG_MAX_SPLITS 256   // chunks of the label axis
#define WSG_RR_LIST 1024     // survivors one re-rank warp can hold

struct WsGemmItem {
  uint32_t row0;    // first sorted row of the query group (multiple of 128)
  uint32_t p0;      // first point (arena rank) of the item's sweep
  uint32_t ntiles;  // 128-point tiles
  uint32_t pend;    // end of the item's slice of the label axis (the last tile may reach past it)
};

struct WsGemmNormArgs {
  const float* vecs;
  uint64_t n;
  uint32_t dpad;
  uint32_t npad;      // entries of norms (n rounded up + one tile); the tail is +inf
  int metric;
  float* norms;       // L2: |x|^2; MIPS: 0
  uint32_t* max_sq;   // float bits of max |x|^2 (non-negative floats order like uints)
  uint32_t* max_abs;  // float bits of max |x_i| (NaN / inf components make it >= 0x7f800000: not eligible)
};

// fp16 mirror of the arena for the sweep: half(x * scale), scale a power of two that puts the largest
// component in [2^13, 2^14).  Rows keep the arena's stride (dpad elements).
struct WsGemmCvtArgs {
  const float* vecs;
  uint64_t count;     // n * dpad
  float scale;
  uint16_t* out;      // __half bits
};

struct WsGemmPlanArgs {
  const float* windows;   // [nq][2] of this slice
  const float* labels;
  uint64_t n;
  uint64_t n_bound;       // n, or n + 1 for a label shard whose last point may be inside a window (WsGeom::pf_n)
  uint32_t nq;            // <= WSG_MAX_ROWS
  uint32_t rows_pad;      // nq rounded up to 128
  uint32_t* qa;           // [nq] scratch: window bounds per query (ws_gemm_bounds_kernel)
  uint32_t* qb;
  uint32_t* perm;         // [rows_pad] sorted row -> query of the slice (0xFFFFFFFF: padding row)
  uint32_t* row_a;        // [rows_pad]
  uint32_t* row_b;        // [rows_pad]
  WsGemmItem* items;
  uint32_t* nitems;       // out
  uint32_t* sched_ctr;    // out: zeroed here, the sweep kernel's dynamic item counter
  uint32_t max_items;
  uint32_t* group_items;  // [groups][WSG_MAX_SPLITS] item indices of each group
  uint32_t* group_cnt;    // [groups]
  uint32_t target_items;
  uint32_t min_tiles;
  uint32_t max_tiles;
  uint32_t* overflow;
};

struct WsGemmPackArgs {
  const float* queries;   // [nq][dim] of this slice
  uint32_t dim, dpad, rows_pad;
  int metric;
  const uint32_t* perm;
  const uint32_t* max_sq;
  uint32_t kcols;         // fp16 columns of a packed row (blocks per tile x 64; zero beyond dpad)
  int32_t x_exp;          // the arena mirror holds half(x * 2^x_exp)
  uint16_t* qpack;        // [rows_pad][kcols] half(scale * q * 2^q_exp(row))
  float* rscale;          // [rows_pad]  2^-(x_exp + q_exp(row)): accumulator -> score units
  float* slack;           // [rows_pad]  2E
  float* qnorm;           // [rows_pad]  |q|^2
};

// seed thresholds: exact distances of a small even sample of each query's window
struct WsGemmSeedArgs {
  const float* vecs;
  const float* queries;
  uint32_t dim, dpad, rows_pad, k;
  const uint32_t* perm;
  const uint32_t* row_a;
  const uint32_t* row_b;
  const float* slack;
  const float* qnorm;
  uint32_t* thr0;         // [rows_pad] seed threshold, order-preserving uint encoding (ws_ord)
};

struct WsGemmArgs {
  const WsGemmItem* items;
  const uint32_t* nitems;
  uint32_t* sched_ctr;  // items beyond the first gridDim.x are drawn from this counter (dyn != 0)
  uint32_t dyn;         // 0: item i of CTA b is b + i*gridDim.x (static striping)
  const uint32_t* row_a;
  const uint32_t* row_b;
  const float* slack;
  uint32_t* gthr;       // [rows_pad] per query: best threshold any of its work items has reached so far
                        // (ws_ord encoding, atomicMin); seeded by ws_gemm_seed_kernel
  const float* norms;
  uint64_t* cand;       // [max_items][WSG_CAND_CAP][128]  (score~, point) keys
  uint32_t* cand_cnt;   // [max_items][128]   0xFFFFFFFF: overflow
  float* cand_thr;      // [max_items][128]   final threshold
  const uint16_t* qpack;  // [rows_pad][kcols] packed fp16 queries (the A operand, copied into TMEM per item)
  const float* rscale;  // [rows_pad] accumulator -> score units
  uint32_t kcols;
  uint32_t nkb;         // 64-column fp16 blocks per tile (1, 2, 3, 4, 6 or 8)
  uint32_t kbps;        // blocks per pipeline stage (<= 4: a stage is at most 64 KB)
  uint32_t nacc;        // accumulator stages in use (3, or 2 when the query operand takes 256 columns)
  uint32_t acc_col0;    // first accumulator column (128 or 256)
  uint32_t k;
  uint32_t dbg;         // timing experiments only (results invalid): 1 = epilogue releases stages without
                        // reading them, 2 = producer signals point blocks without loading them, 4 = epilogue reads
                        // TMEM and releases, nothing else, 8 = reads, converts and filters but keeps no candidate
};

struct WsGemmRerankArgs {
  const float* vecs;
  const float* queries;    // [nq][dim] of this slice
  uint32_t dim, dpad;
  uint32_t rows_pad;
  const uint32_t* perm;
  const uint32_t* row_a;
  const uint32_t* row_b;
  const uint32_t* group_items;
  const uint32_t* group_cnt;
  const uint64_t* cand;
  const uint32_t* cand_cnt;
  const float* cand_thr;
  uint32_t k;
  uint32_t* out_ids;       // [nq][k] of this slice
  float* out_dists;
  const uint32_t* decode;
  uint32_t pad_id;
  uint64_t* res_keys;      // scratch [rows_pad][k] (ws_scan_task writes partial rows there)
  uint32_t* res_cnt;       // scratch [rows_pad]
  unsigned long long* stats;
  unsigned long long* gstats;  // [0] survivors re-ranked, [1] queries that fell back to the exact scan
};

// launchers (defined in ws_gemm.cu, its own translation unit)
size_t wsg_topk_smem_bytes();
cudaError_t wsg_init_attributes();
cudaError_t wsg_launch_norm(int grid, cudaStream_t st, const WsGemmNormArgs& a);
cudaError_t wsg_launch_cvt(int grid, cudaStream_t st, const WsGemmCvtArgs& a);
cudaError_t wsg_launch_plan(uint32_t nsort, cudaStream_t st, const WsGemmPlanArgs& a);
cudaError_t wsg_launch_pack(cudaStream_t st, const WsGemmPackArgs& a);
cudaError_t wsg_launch_seed(int kq, int metric, bool exact, cudaStream_t st, const WsGemmSeedArgs& a);
cudaError_t wsg_launch_topk(int grid, cudaStream_t st, const CUtensorMap& tm_b, const WsGemmArgs& a);
cudaError_t wsg_launch_rerank(int kq, int metric, bool exact, cudaStream_t st, const WsGemmRerankArgs& a);
