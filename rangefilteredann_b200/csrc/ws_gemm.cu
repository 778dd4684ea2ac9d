// ws_gemm.cu — K1g: tensor-core prefilter (BASELINE north_star item 1).
//
// PrefilterIndex::query_knn (src/prefiltering.h:154-204) scores every point of the query's
// label-sorted slice.  For a BATCH of queries whose windows overlap, that is a dense
// query x slice contraction: queries are sorted by window start, grouped 128 at a time, and
// each group sweeps the union of its windows in 128-point tiles:
//
//   plan    ws_gemm_plan_kernel     window -> [a,b) (prefiltering.h:159-184), sort by a, groups,
//                                   work items (group x chunk of the label axis)
//   pack    ws_gemm_pack_kernel     sorted query matrix, pre-scaled (-2q for L2, -q for MIPS), as fp16
//                                   with a per-query power-of-two scale; per-query error slack
//   GEMM    ws_gemm_topk_kernel     tcgen05.mma kind::f16 (fp16 x fp16 -> fp32), 128x128 accumulators in
//                                   TMEM, points staged by TMA (128B swizzle) from an fp16 MIRROR of the
//                                   arena (half(x * 2^e): same 11-bit significand as tf32, half the bytes,
//                                   twice the tensor rate; the fp32 arena stays the source of truth); the
//                                   epilogue reads the accumulators with tcgen05.ld, rescales, adds |x|^2,
//                                   and keeps per query (one TMEM lane = one query) every point whose
//                                   approximate score is below (k-th best approximate score + slack)
//   re-rank ws_gemm_rerank_kernel   exact fp32 distances of the survivors with the scan kernel's
//                                   arithmetic, top-k, decode, pad  (prefiltering.h:196-201)
//
// Exactness.  score~(q,x) = |x|^2 - 2 q.x (L2) or -q.x (MIPS) computed with fp16 operands
// differs from the fp32 value by at most E = c * |q| * max|x| + a (c: two roundings to 11-bit
// significands per product, Cauchy-Schwarz over the row, fp32 accumulation; a: absolute term for
// the norm table and tiny queries; see ws_gemm_pack_kernel).  A point is dropped only when
// score~ >= kth~ + 2E, which implies its true distance is >= the true distance of k points that
// were kept — so the re-ranked top-k is the exact fp32 top-k of the scan kernel, bit for bit,
// without a verification pass.  If a query's survivor buffer overflows (degenerate data: very
// many points inside the slack) the re-rank warp falls back to the exact streaming scan of the
// window (ws_scan_task), still on the device.
#include "ws_gemm.h"

#include <cuda_fp16.h>

// The shared device code (scan task, sorts, result writers) is header-only; this translation unit
// gets its own internal-linkage copy so the kernels it does not use never clash at link time.
namespace {
#include "ws_kernels.cuh"
}



// ---- raw PTX wrappers (sm_100a) ------------------------------------------------------------
__device__ __forceinline__ uint32_t wsg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void wsg_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(wsg_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void wsg_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(wsg_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wsg_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(wsg_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool wsg_mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(wsg_smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// A pipeline bug must surface as a launch error, not as a hung device: trap after ~2 s.
// The clock is only consulted every 4096 failed polls, so the spin loop itself stays three
// instructions long and does not crowd the issue slots of the other warps of its SM sub-core.
__device__ __forceinline__ void wsg_mbar_wait(uint64_t* bar, uint32_t parity) {
  if (wsg_mbar_try(bar, parity)) return;
  long long t0 = 0;
  uint32_t polls = 0;
  while (!wsg_mbar_try(bar, parity)) {
    if ((++polls & 4095u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}
// Same, for warps that can afford to be late (epilogue): back off between polls so that the
// MMA-issuing and TMA-issuing warps sharing the sub-core get the issue slots.
__device__ __forceinline__ void wsg_mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  if (wsg_mbar_try(bar, parity)) return;
  long long t0 = 0;
  uint32_t polls = 0;
  while (!wsg_mbar_try(bar, parity)) {
    __nanosleep(32);
    if ((++polls & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
  }
}
// one lane of a converged warp (elect.sync): the compiler then emits the single-thread tcgen05 /
// TMA instructions as plain predicated instructions instead of a per-lane serialisation loop
__device__ __forceinline__ bool wsg_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void wsg_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void wsg_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wsg_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wsg_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void wsg_tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          wsg_smem_u32(smem_dst)),
      "l"(tm), "r"(wsg_smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void wsg_prefetch_tmap(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

__device__ __forceinline__ void wsg_tmem_alloc(uint32_t* smem_slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(wsg_smem_u32(smem_slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void wsg_tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, fp16 operands, fp32 accumulate; one thread issues.  The A operand lives in
// TMEM (lane = row, one 32-bit column per PAIR of fp16 elements: K = 16 per instruction = 8 columns): only the point
// block is read from shared memory — an SS-mode 128x128 MMA needs 8 KB of shared-memory operands per 64 tensor
// cycles, twice what the operand path delivers (measured in round 1 with tf32: 140 cycles per MMA, 45 % tensor
// utilisation with nothing else running).
__device__ __forceinline__ void wsg_mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void wsg_tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void wsg_mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(wsg_smem_u32(bar)) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void wsg_tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}

// K-major operand block in shared memory: 128 rows x 128 B, 128-byte swizzle (what the TMA box
// {64 fp16, 128 rows} with CU_TENSOR_MAP_SWIZZLE_128B writes).  Descriptor fields
// (cute::UMMA::SmemDescriptor): start address >> 4 [0,14), leading byte offset >> 4 [16,30)
// (unused for swizzled K-major), stride byte offset >> 4 [32,46) = 1024 B between 8-row
// groups, version 1 [46,48), layout SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ uint64_t wsg_make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 [4,6), a/b format F16 = 0
// [7,10)/[10,13), both K-major, N >> 3 [17,23), M >> 4 [24,29)
__device__ __forceinline__ uint32_t wsg_make_idesc() {
  return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(WSG_TILE_N >> 3) << 17) | ((uint32_t)(WSG_TILE_M >> 4) << 24);
}

// ---- one-time per index: |x|^2 and max |x| -------------------------------------------------
__global__ void __launch_bounds__(256) ws_gemm_norm_kernel(WsGemmNormArgs A) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const uint64_t nwarps = (uint64_t)gridDim.x * (blockDim.x >> 5);
  float mx = 0.f;
  uint32_t ma = 0;  // bits of the largest |x_i| (as unsigned: NaN / inf order above every finite value)
  for (uint64_t r = warp; r < A.npad; r += nwarps) {
    if (r >= A.n) {
      if (lane == 0) A.norms[r] = __int_as_float(0x7f800000);
      continue;
    }
    const float* row = A.vecs + r * A.dpad;
    float acc = 0.f;
    for (uint32_t c = lane; c < A.dpad; c += 32) {
      const float v = row[c];
      acc = fmaf(v, v, acc);
      ma = max(ma, __float_as_uint(v) & 0x7FFFFFFFu);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) A.norms[r] = A.metric == 0 ? acc : 0.f;
    mx = fmaxf(mx, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ma = max(ma, __shfl_xor_sync(0xffffffffu, ma, o));
  if (lane == 0) {
    atomicMax(A.max_sq, __float_as_uint(mx));
    atomicMax(A.max_abs, ma);
  }
}

__global__ void __launch_bounds__(256) ws_gemm_cvt_kernel(WsGemmCvtArgs A) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
  for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < A.count; i += stride) {  // count % 16 == 0
    const float4 v = *reinterpret_cast<const float4*>(A.vecs + i);
    const __half2 a = __floats2half2_rn(v.x * A.scale, v.y * A.scale), b = __floats2half2_rn(v.z * A.scale, v.w * A.scale);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&a);
    o.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(A.out + i) = o;
  }
}

// ---- plan ------------------------------------------------------------------------------------

// prefiltering.h:159-184 (see ws_decompose.h: ws_prefilter_bound)
__device__ __forceinline__ uint32_t wsg_bound(const float* labels, uint64_t n, float v) {
  uint64_t l = 0, r = n - 1;
  while (l < r) {
    const uint64_t mid = (l + r) / 2;
    if (__ldg(labels + mid) < v) l = mid + 1; else r = mid;
  }
  return (uint32_t)l;
}

// window -> [a,b) for every query of the slice (one thread each; the plan CTA only sorts)
__global__ void __launch_bounds__(128) ws_gemm_bounds_kernel(WsGemmPlanArgs A) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.nq) return;
  uint32_t a = wsg_bound(A.labels, A.n_bound, A.windows[2 * (size_t)i]);
  uint32_t b = wsg_bound(A.labels, A.n_bound, A.windows[2 * (size_t)i + 1]);
  if (b <= a) { a = 0; b = 0; }
  A.qa[i] = a; A.qb[i] = b;
}

#define WSG_PLAN_BUCKETS 4096
__global__ void __launch_bounds__(WSG_PLAN_THREADS) ws_gemm_plan_kernel(WsGemmPlanArgs A) {
  extern __shared__ uint32_t s_sorted[];  // [rows_pad] query of every sorted row
  __shared__ uint32_t s_hist[WSG_PLAN_BUCKETS + 1];
  __shared__ uint32_t s_warp_sum[WSG_PLAN_THREADS / 32];
  __shared__ uint32_t s_ga[WSG_MAX_ROWS / 128], s_gb[WSG_MAX_ROWS / 128];
  __shared__ uint32_t s_chunk_cnt[WSG_MAX_SPLITS], s_chunk_off[WSG_MAX_SPLITS];
  __shared__ unsigned long long s_total;
  __shared__ uint32_t s_chunk_pts, s_nchunks;
  const int tid = threadIdx.x;
  // 1. order the queries by window start.  Only locality matters (rows of a group should start
  //    near each other), so a counting sort over 4096 buckets of the label axis replaces a full
  //    sort; empty windows go to a last bucket of their own.
  for (int i = tid; i <= WSG_PLAN_BUCKETS; i += WSG_PLAN_THREADS) s_hist[i] = 0;
  if (tid == 0) s_total = 0;
  if (tid < WSG_MAX_SPLITS) s_chunk_cnt[tid] = 0;
  __syncthreads();
  for (uint32_t i = tid; i < A.nq; i += WSG_PLAN_THREADS) {
    const uint32_t a = A.qa[i], b = A.qb[i];
    const uint32_t bucket = b > a ? (uint32_t)(((uint64_t)a * WSG_PLAN_BUCKETS) / A.n) : WSG_PLAN_BUCKETS;
    atomicAdd(&s_hist[bucket], 1u);
  }
  __syncthreads();
  {  // exclusive scan of the 4097 counters: 4 per thread (+ the last one), warp scan, block scan
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) { v[j] = s_hist[tid * 4 + j]; sum += v[j]; }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += y;
    }
    if ((tid & 31) == 31) s_warp_sum[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
      uint32_t w = s_warp_sum[tid], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, wi, o);
        if (tid >= o) wi += y;
      }
      s_warp_sum[tid] = wi - w;
    }
    __syncthreads();
    uint32_t run = s_warp_sum[tid >> 5] + incl - sum;
#pragma unroll
    for (int j = 0; j < 4; j++) { s_hist[tid * 4 + j] = run; run += v[j]; }
    if (tid == WSG_PLAN_THREADS - 1) s_hist[WSG_PLAN_BUCKETS] = run;  // empties start after every real bucket
  }
  __syncthreads();
  for (uint32_t i = tid; i < A.nq; i += WSG_PLAN_THREADS) {
    const uint32_t a = A.qa[i], b = A.qb[i];
    const uint32_t bucket = b > a ? (uint32_t)(((uint64_t)a * WSG_PLAN_BUCKETS) / A.n) : WSG_PLAN_BUCKETS;
    s_sorted[atomicAdd(&s_hist[bucket], 1u)] = i;
  }
  for (uint32_t r = A.nq + tid; r < A.rows_pad; r += WSG_PLAN_THREADS) s_sorted[r] = 0xFFFFFFFFu;
  __syncthreads();
  // 3. sorted rows, 4. group extents (one warp per group of 128 rows)
  const uint32_t groups = A.rows_pad / 128;
  for (uint32_t g = tid >> 5; g < groups; g += WSG_PLAN_THREADS / 32) {
    uint32_t ga = 0xFFFFFFFFu, gb = 0;
    for (uint32_t r = g * 128 + (tid & 31); r < (g + 1) * 128; r += 32) {
      const uint32_t q = s_sorted[r];
      uint32_t a = 0, b = 0;
      if (q != 0xFFFFFFFFu) { a = A.qa[q]; b = A.qb[q]; }
      A.perm[r] = q; A.row_a[r] = a; A.row_b[r] = b;
      if (b > a) { ga = min(ga, a); gb = max(gb, b); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ga = min(ga, __shfl_xor_sync(0xffffffffu, ga, o));
      gb = max(gb, __shfl_xor_sync(0xffffffffu, gb, o));
    }
    if ((tid & 31) == 0) {
      if (gb == 0) ga = 0;
      s_ga[g] = ga; s_gb[g] = gb;
      atomicAdd(&s_total, (unsigned long long)((gb - ga + WSG_TILE_N - 1) / WSG_TILE_N));
    }
  }
  __syncthreads();
  // 5. chunk the label axis so that about target_items items of equal size come out
  if (tid == 0) {
    unsigned long long per = (s_total + A.target_items - 1) / A.target_items;
    if (per < A.min_tiles) per = A.min_tiles;
    // the slices swept concurrently must stay L2-resident; when the whole batch is huge (every group
    // sweeps most of the axis) twice the chunk measured better: fewer items, same sharing
    const unsigned long long cap_tiles = s_total >= 300000ull ? 2ull * A.max_tiles : A.max_tiles;
    if (per > cap_tiles) per = cap_tiles;
    unsigned long long pts = per * WSG_TILE_N;
    const unsigned long long floor_pts = (A.n + WSG_MAX_SPLITS - 1) / WSG_MAX_SPLITS;
    if (pts < floor_pts) pts = floor_pts;
    pts = (pts + WSG_TILE_N - 1) / WSG_TILE_N * WSG_TILE_N;
    s_chunk_pts = (uint32_t)(pts > 0xFFFFFF80ull ? 0xFFFFFF80ull : pts);
    s_nchunks = (uint32_t)((A.n + s_chunk_pts - 1) / s_chunk_pts);
  }
  __syncthreads();
  const uint32_t cp = s_chunk_pts;
  // items are ordered chunk-major so that CTAs running at the same time sweep the same
  // points (for different query groups) and share them through L2
  uint32_t c0 = 0, c1 = 0;
  const bool live = (uint32_t)tid < groups && s_gb[tid] > s_ga[tid];
  if (live) {
    c0 = s_ga[tid] / cp; c1 = (s_gb[tid] - 1) / cp;
    for (uint32_t c = c0; c <= c1; c++) atomicAdd(&s_chunk_cnt[c], 1u);
  }
  __syncthreads();
  if (tid == 0) {
    uint32_t off = 0;
    for (uint32_t c = 0; c < s_nchunks; c++) { s_chunk_off[c] = off; off += s_chunk_cnt[c]; s_chunk_cnt[c] = 0; }
    *A.nitems = off <= A.max_items ? off : 0;
    *A.sched_ctr = 0;
    if (off > A.max_items) atomicExch(A.overflow, 1u);
  }
  __syncthreads();
  if ((uint32_t)tid < groups) {
    uint32_t cnt = 0;
    if (live && *A.nitems) {
      for (uint32_t c = c0; c <= c1; c++) {
        const uint32_t it = s_chunk_off[c] + atomicAdd(&s_chunk_cnt[c], 1u);
        const uint32_t lo = max(s_ga[tid], c * cp);
        const uint64_t hi64 = (uint64_t)(c + 1) * cp;
        const uint32_t hi = (uint32_t)min((uint64_t)s_gb[tid], hi64);
        WsGemmItem item;
        item.row0 = tid * 128; item.p0 = lo; item.ntiles = (hi - lo + WSG_TILE_N - 1) / WSG_TILE_N; item.pend = hi;
        A.items[it] = item;
        A.group_items[tid * WSG_MAX_SPLITS + cnt] = it;
        cnt++;
      }
    }
    A.group_cnt[tid] = cnt;
  }
}

// ---- pack ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ws_gemm_pack_kernel(WsGemmPackArgs A) {
  const int lane = threadIdx.x & 31;
  const uint32_t row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= A.rows_pad) return;
  const uint32_t q = A.perm[row];
  uint16_t* out = A.qpack + (size_t)row * A.kcols;
  const float scale = A.metric == 0 ? -2.f : -1.f;
  // pass 1: |q|^2 and the largest |scale * q_i|
  float acc = 0.f, mab = 0.f;
  if (q != 0xFFFFFFFFu)
    for (uint32_t c = lane; c < A.dim; c += 32) {
      const float v = A.queries[(size_t)q * A.dim + c];
      acc = fmaf(v, v, acc);
      mab = fmaxf(mab, fabsf(scale * v));
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    mab = fmaxf(mab, __shfl_xor_sync(0xffffffffu, mab, o));
  }
  // per-query power-of-two scale: the largest component lands in [2^13, 2^14) (fp16 overflows at 2^16)
  int qe = 0;
  const bool sane = acc < 3.0e38f && mab < 3.0e38f;  // false for NaN / inf
  if (sane && mab > 0.f) {
    int ex;
    frexpf(mab, &ex);  // mab = m * 2^ex, m in [0.5, 1)
    qe = 14 - ex;
  }
  const bool ok = sane && qe >= -60 && qe <= 60;
  if (!ok) qe = 0;
  const float qs = exp2f((float)qe);
  for (uint32_t c = lane; c < A.kcols; c += 32) {
    float v = 0.f;
    if (ok && q != 0xFFFFFFFFu && c < A.dim) v = scale * A.queries[(size_t)q * A.dim + c] * qs;
    out[c] = __half_as_ushort(__float2half_rn(v));
  }
  if (lane == 0) {
    // Per product (scale q_i)(x_i): both factors rounded to nearest fp16 (11-bit significands, 2^-11 each; the scaled
    // values sit far above the subnormal range)  ->  |err| <= (2^-10 + 2^-22) |scale q_i x_i|, summed with
    // Cauchy-Schwarz; the products are exact in fp32 and accumulated in fp32 over dpad terms (dpad * 2^-21 covers an
    // accumulator that truncates).  The absolute term covers the fp32 rounding of |x|^2 (summed in another order than
    // the exact distance), of the exact distance itself, and queries so small that the relative term vanishes.
    const float xmax2 = __uint_as_float(*A.max_sq);
    const float xmax = sqrtf(xmax2);
    const float rel = 0.0009765625f * 1.02f + (float)A.dpad * 4.76837158e-7f;
    const float e = fabsf(scale) * sqrtf(acc) * xmax * rel + (float)A.dpad * 2.38418579e-7f * (acc + xmax2);
    // a query this engine cannot scale (NaN, inf, absurd magnitude): infinite slack keeps every point, the survivor
    // list overflows and the re-rank warp answers with the exact streaming scan
    A.slack[row] = ok ? 2.f * e : __int_as_float(0x7f800000);
    A.qnorm[row] = ok ? acc : 0.f;
    A.rscale[row] = exp2f((float)(-(A.x_exp + qe)));
  }
}

// ---- seed ------------------------------------------------------------------------------------
// A finite threshold before the sweep starts: D = k-th smallest EXACT distance over an even
// sample of <= 64 points of the window is an upper bound of the true k-th distance, so every
// true top-k point has score~ <= D - |q|^2 + E.  Starting from thr0 = D - |q|^2 + E (instead of
// +inf) removes the warm-up of every work item, where otherwise every column would pass.
#define WSG_SEED_SAMPLES 64
template <int KQ, int METRIC, bool EXACT>
__global__ void __launch_bounds__(WS_WARPS_PER_CTA * 32) ws_gemm_seed_kernel(WsGemmSeedArgs A) {
  __shared__ float s_d[WS_WARPS_PER_CTA][WSG_SEED_SAMPLES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane & (WS_TEAM - 1), team = lane / WS_TEAM;
  const uint32_t row = blockIdx.x * WS_WARPS_PER_CTA + warp;
  if (row >= A.rows_pad) return;
  const uint32_t qi = A.perm[row];
  float out = __int_as_float(0x7f800000);
  const uint32_t a = A.row_a[row], b = A.row_b[row];
  const uint32_t w = b - a;
  if (qi != 0xFFFFFFFFu && w >= A.k) {
    const uint32_t ns = min(w, (uint32_t)WSG_SEED_SAMPLES);
    const int dpad4 = A.dpad >> 2;
    float4 q[KQ];
    ws_load_query_global<KQ, EXACT>(A.queries, A.dim, A.dpad, qi, tl, q);
    const float4* vtl = reinterpret_cast<const float4*>(A.vecs) + tl;
#pragma unroll 4
    for (int r0 = 0; r0 < WSG_SEED_SAMPLES; r0 += 4) {
      const uint32_t i = r0 + team;
      const bool valid = i < ns;
      const uint32_t p = a + (valid ? (uint32_t)(((uint64_t)(2 * i + 1) * w) / (2ull * ns)) : 0u);
      const float d = ws_team_dist_nv<KQ, METRIC, EXACT>(vtl + (size_t)p * dpad4, q, tl, dpad4);
      if (tl == 0) s_d[warp][i] = valid ? d : __int_as_float(0x7f800000);
    }
    __syncwarp();
    uint64_t k0 = ws_key(s_d[warp][lane], lane), k1 = ws_key(s_d[warp][lane + 32], lane + 32);
    ws_warp_sort64(k0, k1, lane);
    const uint32_t kk = A.k - 1;
    const uint64_t kth = ws_shfl_idx_u64(kk < 32 ? k0 : k1, kk & 31);
    const float D = ws_unord((uint32_t)(kth >> 32));
    const float e = 0.5f * A.slack[row];
    out = METRIC == 0 ? D - A.qnorm[row] + e : D + e;  // +inf when the slack is (unscalable query)
  }
  if (lane == 0) A.thr0[row] = ws_ord(out);
}


// ---- the tensor-core sweep -------------------------------------------------------------------

struct WsGemmSmem {
  uint64_t full[WSG_B_STAGES], empty[WSG_B_STAGES];
  uint64_t a_full, a_empty;
  uint64_t acc_full[WSG_ACC_STAGES], acc_empty[WSG_ACC_STAGES];
  uint32_t tmem_base;
  uint32_t pad_;
  alignas(16) float wnorm[WSG_EPI_WARPS][32];   // |x|^2 of the warp's 32 columns of the current tile
  float thr[WSG_TILE_M];                         // per query: k-th best score~ so far + slack
  uint32_t cnt[WSG_TILE_M];                      // survivors appended so far
  uint32_t head[WSG_TILE_M];                     // fresh scores pushed into the ring so far
  uint32_t warps_done;                           // epilogue warps that finished an item (monotonic)
  uint32_t pad2_[3];
  uint32_t ring[WSG_RING][WSG_TILE_M];           // fresh scores (float bits), consumed by the owner thread
  // The CTA's item sequence: entry i & (WSG_SCHED-1) = (i + 1) << 32 | item id (WSG_SENT: no more items),
  // written by the TMA warp — the role that runs ahead of all others — and polled by the other roles.
  // No hand-back is needed: the TMA warp is at most WSG_B_STAGES tiles, hence items, ahead of the MMA
  // warp, which is at most one item ahead of the epilogue and threshold warps.
  unsigned long long sched[WSG_SCHED];
};
#define WSG_SENT 0xFFFFFFFFu
static_assert(WSG_SCHED >= 2 * (WSG_B_STAGES + WSG_ACC_STAGES + 2), "item ring too short for the pipeline depth");

// item number `seq` of this CTA, as published by the TMA warp (warp-uniform)
__device__ __forceinline__ uint32_t wsg_take_item(WsGemmSmem* S, uint32_t seq) {
  volatile unsigned long long* p = &S->sched[seq & (WSG_SCHED - 1)];
  unsigned long long v = *p;
  long long t0 = 0;
  uint32_t polls = 0;
  while ((uint32_t)(v >> 32) != seq + 1) {
    __nanosleep(32);
    if ((++polls & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ll) __trap();
    }
    v = *p;
  }
  return (uint32_t)v;
}

// Tiles of an item may be swept in any order.  The groups that sweep the same slice of the label
// axis at the same time share its points through L2, so they must stay CLOSE to each other (a
// spread over the whole slice made the live footprint ~110 MB and the L2 hit rate 38 %), but not
// on the very same lines at the very same moment: each group starts a few tiles further in and
// wraps around.
__device__ __forceinline__ uint32_t wsg_rotation(const WsGemmItem& item) {
  return (((item.row0 >> 7) & 7u) * 3u) % item.ntiles;
}
__device__ __forceinline__ uint32_t wsg_tile(uint32_t t, uint32_t rot, uint32_t ntiles) {
  const uint32_t x = t + rot;
  return x >= ntiles ? x - ntiles : x;
}

// Threshold thread of a query: fold the fresh scores the epilogue warps pushed into the ring into
// the running top-KTOP (ascending, registers) and publish the new threshold.  Slots are taken with
// an atomic exchange, so a score is folded at most once; a slot whose value has not landed yet
// (or was overwritten on wrap-around) is simply skipped — that can only leave thr looser.
__device__ __forceinline__ bool wsg_flush(float (&tk)[WSG_KTOP], WsGemmSmem* S, int lrow, uint32_t& last_head, int k,
                                          float slack, float& thr, float thr0) {
  const uint32_t h = *(volatile uint32_t*)&S->head[lrow];
  const uint32_t nnew = min(h - last_head, (uint32_t)WSG_RING);
  if (!__any_sync(0xffffffffu, nnew != 0)) return false;
  // tk holds KTOP - k phantom entries of -inf in front, so the k-th best real score is always tk[KTOP-1]
  float kth = tk[WSG_KTOP - 1];
  bool changed = false;
#pragma unroll 1
  for (uint32_t i = 0; i < WSG_RING; i++) {
    if (!__any_sync(0xffffffffu, i < nnew)) break;
    uint32_t bits = WSG_SENT;
    if (i < nnew) bits = atomicExch(&S->ring[(h - 1 - i) & (WSG_RING - 1)][lrow], WSG_SENT);
    float v = bits == WSG_SENT ? __int_as_float(0x7f800000) : __uint_as_float(bits);
    // scores between the k-th best and the threshold pass the filter but cannot move the top-k
    if (!__any_sync(0xffffffffu, v < kth)) continue;
    changed = true;
#pragma unroll
    for (int x = 0; x < WSG_KTOP; x++) {
      const float lo = fminf(tk[x], v);
      v = fmaxf(tk[x], v);
      tk[x] = lo;
    }
    kth = tk[WSG_KTOP - 1];
  }
  last_head = h;
  if (changed) {
    thr = fminf(thr0, kth + slack);  // the seed threshold while fewer than k points have been seen
    *(volatile float*)&S->thr[lrow] = thr;
  }
  return true;
}

__global__ void __launch_bounds__(WSG_THREADS, 1)
ws_gemm_topk_kernel(const __grid_constant__ CUtensorMap tmB, WsGemmArgs A) {
  extern __shared__ unsigned char wsg_smem_raw[];
  // operand blocks need 1024-byte alignment (128B swizzle atoms)
  unsigned char* base = wsg_smem_raw + ((1024u - (wsg_smem_u32(wsg_smem_raw) & 1023u)) & 1023u);
  unsigned char* sB = base;                                  // ring of 16 KB point blocks
  WsGemmSmem* S = (WsGemmSmem*)(sB + WSG_B_STAGES * WSG_KBLK_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t nitems = *A.nitems;
  const uint32_t nkb = A.nkb;

  if (threadIdx.x == 0) {
    for (int i = 0; i < WSG_B_STAGES; i++) { wsg_mbar_init(&S->full[i], 1); wsg_mbar_init(&S->empty[i], 1); }
    wsg_mbar_init(&S->a_full, WSG_EPI_WARPS); wsg_mbar_init(&S->a_empty, 1);
    for (int i = 0; i < WSG_ACC_STAGES; i++) { wsg_mbar_init(&S->acc_full[i], 1); wsg_mbar_init(&S->acc_empty[i], WSG_EPI_WARPS); }
    wsg_fence_barrier_init();
    wsg_prefetch_tmap(&tmB);
  }
  if (warp == 1) wsg_tmem_alloc(&S->tmem_base, 512);
  if (threadIdx.x == 64) S->warps_done = 0;
  if (threadIdx.x >= 96 && threadIdx.x < 96 + WSG_SCHED) S->sched[threadIdx.x - 96] = 0ull;
  wsg_tc_fence_before();
  __syncthreads();
  wsg_tc_fence_after();
  const uint32_t tmem_base = S->tmem_base;

  // A pipeline stage holds kbps blocks of 16 KB — a whole tile of points up to 256 columns, half a tile above
  // (a stage is at most 64 KB): one barrier round trip and one tcgen05.commit per 4*kbps MMAs.  (With a stage
  // per block the tensor pipe drained at every commit: 1830 cycles per tile instead of the 1024 its MMAs needed.)
  const uint32_t kbps = A.kbps, spt = nkb / kbps;  // blocks per stage, stages per tile
  const uint32_t nstages = WSG_B_STAGES / kbps;
  if (warp == 0) {
    // ===== TMA producer (whole warp walks the loop, one elected lane issues) =====
    uint32_t stage = 0, phase = 0;
    uint32_t it = blockIdx.x;
    for (uint32_t iseq = 0;; iseq++) {
      // publish the CTA's next item to the other roles, then fetch its points
      if (it >= nitems) it = WSG_SENT;
      if (lane == 0) *(volatile unsigned long long*)&S->sched[iseq & (WSG_SCHED - 1)] = ((unsigned long long)(iseq + 1) << 32) | it;
      if (it == WSG_SENT) break;
      const WsGemmItem item = A.items[it];
      // items are ordered chunk-major: drawing them from one counter keeps the CTAs that run at the same
      // time on neighbouring items (same slice of the label axis) and ends every CTA within one item of
      // the others, whatever the item sizes
      uint32_t nx = it + gridDim.x;
      if (A.dyn) {
        if (lane == 0) nx = gridDim.x + atomicAdd(A.sched_ctr, 1u);
        nx = __shfl_sync(0xffffffffu, nx, 0);
      }
      it = nx;
      const uint32_t rot = wsg_rotation(item);
      for (uint32_t t = 0; t < item.ntiles; t++) {
        const int p = (int)(item.p0 + wsg_tile(t, rot, item.ntiles) * WSG_TILE_N);
        for (uint32_t sp = 0; sp < spt; sp++) {
          wsg_mbar_wait(&S->empty[stage], phase ^ 1);
          if (wsg_elect_one()) {
            if (A.dbg & 2) {
              wsg_mbar_arrive(&S->full[stage]);
            } else {
              // blocks past the row's last column (dpad rounded up to the block count) are zero-filled by the TMA unit
              wsg_mbar_expect_tx(&S->full[stage], kbps * WSG_KBLK_BYTES);
              for (uint32_t kb = 0; kb < kbps; kb++)
                wsg_tma_load_2d(sB + (stage * kbps + kb) * WSG_KBLK_BYTES, &tmB, &S->full[stage], (int)((sp * kbps + kb) * WSG_KBLK), p);
            }
          }
          __syncwarp();
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (whole warp walks the loop, one elected lane issues) =====
    // tcgen05.mma issue is in lock step with the tensor pipe (a short queue), so every instruction
    // between the last MMA of a tile and the first MMA of the next one is a bubble on the pipe
    // (measured: 330-640 idle cycles per 1024-cycle tile).  The barrier waits of tile t+1 are
    // therefore taken in the MIDDLE of tile t's MMAs, and descriptors are plain adds.
    const uint32_t idesc = wsg_make_idesc();
    const uint64_t desc0 = wsg_make_desc(wsg_smem_u32(sB));
    const uint32_t tmem_acc0 = tmem_base + A.acc_col0;  // the columns below hold the queries
    const uint32_t nacc = A.nacc;
    const uint32_t nmma = kbps * 4, half = nmma / 2;    // one MMA = 16 fp16 columns = 32 B of every row, 8 TMEM columns of A
    uint32_t stage = 0, phase = 0, a_phase = 0, acc = 0, acc_phase = 0;
    for (uint32_t iseq = 0;; iseq++) {
      const uint32_t it = wsg_take_item(S, iseq);
      if (it == WSG_SENT) break;
      const uint32_t ntiles = A.items[it].ntiles;
      wsg_mbar_wait(&S->a_full, a_phase);
      a_phase ^= 1;
      wsg_mbar_wait(&S->acc_empty[acc], acc_phase ^ 1);
      wsg_mbar_wait(&S->full[stage], phase);
      wsg_tc_fence_after();
      for (uint32_t t = 0; t < ntiles; t++) {
        const uint32_t d_tmem = tmem_acc0 + acc * WSG_TILE_N;
        for (uint32_t sp = 0; sp < spt; sp++) {
          const uint64_t db = desc0 + (uint64_t)(stage * kbps) * (WSG_KBLK_BYTES >> 4);
          const uint32_t ta = tmem_base + sp * kbps * 32;  // this stage's 32 query columns per block
          if (wsg_elect_one()) {
            for (uint32_t i = 0; i < half; i++)  // MMA i: block i/4 of the stage, 16-column step i%4
              wsg_mma_f16_ts(d_tmem, ta + i * 8, db + (uint64_t)((i >> 2) * (WSG_KBLK_BYTES >> 4) + (i & 3) * 2), idesc, (sp | i) != 0 ? 1u : 0u);
          }
          __syncwarp();
          const bool last_sp = sp + 1 == spt;
          uint32_t nstage = stage + 1, nphase = phase, nxacc = acc, nxacc_phase = acc_phase;
          if (nstage == nstages) { nstage = 0; nphase ^= 1; }
          if (last_sp && ++nxacc == nacc) { nxacc = 0; nxacc_phase ^= 1; }
          if (!(last_sp && t + 1 == ntiles)) {  // next stage's operands (and accumulator), while this stage's MMAs are queued
            if (last_sp) wsg_mbar_wait(&S->acc_empty[nxacc], nxacc_phase ^ 1);
            wsg_mbar_wait(&S->full[nstage], nphase);
            wsg_tc_fence_after();
          }
          if (wsg_elect_one()) {
            for (uint32_t i = half; i < nmma; i++)
              wsg_mma_f16_ts(d_tmem, ta + i * 8, db + (uint64_t)((i >> 2) * (WSG_KBLK_BYTES >> 4) + (i & 3) * 2), idesc, 1u);
            wsg_mma_commit(&S->empty[stage]);
            if (last_sp) wsg_mma_commit(&S->acc_full[acc]);
          }
          __syncwarp();
          stage = nstage; phase = nphase; acc = nxacc; acc_phase = nxacc_phase;
        }
      }
      if (wsg_elect_one()) wsg_mma_commit(&S->a_empty);
      __syncwarp();
    }
  } else if (warp < 2 + WSG_EPI_WARPS) {
    // ===== epilogue: 16 warps; a query is one TMEM lane, served by the four warps of its lane
    // quarter, each scanning 32 of the tile's 128 columns: filter against the query's published
    // threshold, append survivors, push their scores to the query's ring. =====
    const int e = warp - 2;
    const int quarter = warp & 3;               // the TMEM lane quarter this warp may read
    const int chunk = e >> 2;                   // its 32 columns of every tile
    const int lrow = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float INF = __int_as_float(0x7f800000);
    float* my_norm = S->wnorm[e];
    uint32_t acc = 0, acc_phase = 0, a_phase = 0;
    for (uint32_t iseq = 0;; iseq++) {
      const uint32_t it = wsg_take_item(S, iseq);
      if (it == WSG_SENT) break;
      const WsGemmItem item = A.items[it];
      const uint32_t row = item.row0 + lrow;
      uint64_t* cand = A.cand + (size_t)it * WSG_CAND_CAP * WSG_TILE_M + lrow;
      const uint32_t rot = wsg_rotation(item);
      // the query's window clipped to this item's slice of the label axis (ranks < 2^31)
      const int wlo = (int)max(A.row_a[row], item.p0), whi = (int)min(A.row_b[row], item.pend);
      const float rs = A.rscale[row];  // accumulator units -> score units (a power of two)
      {
        // the group's queries -> TMEM columns [32*chunk, 32*chunk+32) of lanes [32*quarter, +32), once the
        // previous item's MMAs have drained
        wsg_mbar_wait(&S->a_empty, a_phase ^ 1);
        a_phase ^= 1;
        wsg_tc_fence_after();
        // block c of the packed row = 64 fp16 = 32 TMEM columns of this lane (two elements per column)
        for (uint32_t c = (uint32_t)chunk; c < nkb; c += 4) {
          uint32_t qv[32];
          const uint4* qrow = reinterpret_cast<const uint4*>(A.qpack + (size_t)row * A.kcols) + c * 8;
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const uint4 x = __ldg(qrow + j);
            qv[4 * j] = x.x; qv[4 * j + 1] = x.y; qv[4 * j + 2] = x.z; qv[4 * j + 3] = x.w;
          }
          wsg_tmem_st32(tmem_base + lane_addr + c * 32, qv);
        }
        wsg_tc_fence_before();
        __syncwarp();
        if (lane == 0) wsg_mbar_arrive(&S->a_full);
      }
      asm volatile("bar.sync 1, 640;" ::: "memory");  // per-query state reset by the threshold warps
      float nrm_next = __ldg(A.norms + item.p0 + rot * WSG_TILE_N + chunk * 32 + lane);
      for (uint32_t t = 0; t < item.ntiles; t++) {
        const int cbase = (int)(item.p0 + wsg_tile(t, rot, item.ntiles) * WSG_TILE_N + chunk * 32);
        __syncwarp();
        my_norm[lane] = nrm_next;
        __syncwarp();
        if (t + 1 < item.ntiles)
          nrm_next = __ldg(A.norms + item.p0 + wsg_tile(t + 1, rot, item.ntiles) * WSG_TILE_N + chunk * 32 + lane);
        wsg_mbar_wait_relaxed(&S->acc_full[acc], acc_phase);
        wsg_tc_fence_after();
        const int lo = max(0, min(32, wlo - cbase));
        const int hi = max(0, min(32, whi - cbase));
        if (__any_sync(0xffffffffu, hi > lo) && !(A.dbg & 1)) {
          // Windows are contiguous and tiles walk the sorted points, so a query meets a PARTIALLY covered chunk only
          // where its window starts and where it ends — twice per item.  Everywhere else its 32 columns are either all
          // inside (compare as they are) or all outside (nothing may pass: threshold -inf).  The per-element masking
          // (4 instructions per score) is therefore taken only by the few warp-tiles that hold a window edge.
          const bool outside = hi <= lo;
          const bool edge = __any_sync(0xffffffffu, !outside && (lo != 0 || hi != 32));
          float s[32];
          wsg_tmem_ld32(tmem_base + lane_addr + A.acc_col0 + acc * WSG_TILE_N + chunk * 32, s);
          // the scores are in registers: hand the accumulator stage back before filtering them
          wsg_tc_fence_before();
          __syncwarp();
          if (lane == 0) wsg_mbar_arrive(&S->acc_empty[acc]);
          if (A.dbg & 4) { if (++acc == A.nacc) { acc = 0; acc_phase ^= 1; } continue; }  // timing: TMEM read only
          const float4* nr = reinterpret_cast<const float4*>(my_norm);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float4 nv = nr[j];
            s[4 * j + 0] = fmaf(s[4 * j + 0], rs, nv.x); s[4 * j + 1] = fmaf(s[4 * j + 1], rs, nv.y);
            s[4 * j + 2] = fmaf(s[4 * j + 2], rs, nv.z); s[4 * j + 3] = fmaf(s[4 * j + 3], rs, nv.w);
          }
          if (edge) {
#pragma unroll
            for (int j = 0; j < 32; j++) s[j] = (j >= lo && j < hi) ? s[j] : INF;
          }
          const float thr = outside ? -INF : *(volatile float*)&S->thr[lrow];
          float m4[8];
#pragma unroll
          for (int g = 0; g < 8; g++) m4[g] = fminf(fminf(s[4 * g], s[4 * g + 1]), fminf(s[4 * g + 2], s[4 * g + 3]));
          const float m = fminf(fminf(fminf(m4[0], m4[1]), fminf(m4[2], m4[3])), fminf(fminf(m4[4], m4[5]), fminf(m4[6], m4[7])));
          if (__any_sync(0xffffffffu, m < thr) && !(A.dbg & 8)) {  // dbg 8 (timing): filter, but keep nothing
#pragma unroll
            for (int g = 0; g < 8; g++) {
              if (!__any_sync(0xffffffffu, m4[g] < thr)) continue;
#pragma unroll
              for (int u = 0; u < 4; u++) {
                const float v = s[4 * g + u];
                if (v < thr) {
                  const uint32_t pos = atomicAdd(&S->cnt[lrow], 1u);
                  if (pos < WSG_CAND_CAP) cand[(size_t)pos * WSG_TILE_M] = ws_key(v, (uint32_t)(cbase + 4 * g + u));
                  const uint32_t slot = atomicAdd(&S->head[lrow], 1u) & (WSG_RING - 1);
                  S->ring[slot][lrow] = __float_as_uint(v);
                }
              }
            }
          }
        } else {
          wsg_tc_fence_before();
          __syncwarp();
          if (lane == 0) wsg_mbar_arrive(&S->acc_empty[acc]);
        }
        if (++acc == A.nacc) { acc = 0; acc_phase ^= 1; }
      }
      __syncwarp();
      if (lane == 0) atomicAdd(&S->warps_done, 1u);
      asm volatile("bar.sync 1, 640;" ::: "memory");  // every append of this item has been issued
    }
  } else {
    // ===== threshold warps (4): thread = query.  Hold the query's running top-k of approximate
    // scores in registers, fold what the epilogue warps push into the ring, publish
    // thr = min(seed, k-th best + slack).  Off the accumulator pipeline on purpose: a slow fold never
    // delays the release of a TMEM stage. =====
    const int lrow = (warp - 2 - WSG_EPI_WARPS) * 32 + lane;
    const float INF = __int_as_float(0x7f800000);
    const int k = (int)A.k;
    uint32_t seq = 0;
    for (uint32_t iseq = 0;; iseq++) {
      const uint32_t it = wsg_take_item(S, iseq);
      if (it == WSG_SENT) break;
      const uint32_t row = A.items[it].row0 + lrow;
      const float slack = A.slack[row];
      // A threshold reached by ANY work item of the query (k-th best of a subset + slack) bounds the
      // query's true k-th distance, so items share it through global memory: later slices of the label
      // axis start from the threshold earlier ones reached instead of the seed.
      uint32_t* gthr = A.gthr + row;
      float thr0 = ws_unord(*(volatile uint32_t*)gthr);
      float tk[WSG_KTOP];
#pragma unroll
      for (int x = 0; x < WSG_KTOP; x++) tk[x] = x < WSG_KTOP - k ? -INF : INF;
      float thr = thr0, pub = thr0;
      uint32_t last_head = 0, polls = 0;
      S->thr[lrow] = thr0; S->cnt[lrow] = 0; S->head[lrow] = 0;
#pragma unroll
      for (int i = 0; i < WSG_RING; i++) S->ring[i][lrow] = WSG_SENT;
      asm volatile("bar.sync 1, 640;" ::: "memory");
      seq++;
      while (*(volatile uint32_t*)&S->warps_done < seq * WSG_EPI_WARPS) {
        const bool got = wsg_flush(tk, S, lrow, last_head, k, slack, thr, thr0);
        if (thr < pub) { atomicMin(gthr, ws_ord(thr)); pub = thr; }
        if (!got) {
          if ((++polls & 15u) == 0) {
            const float g = ws_unord(*(volatile uint32_t*)gthr);
            if (g < thr) { thr0 = g; thr = g; pub = g; *(volatile float*)&S->thr[lrow] = g; }
          }
          __nanosleep(64);
        }
      }
      asm volatile("bar.sync 1, 640;" ::: "memory");
      wsg_flush(tk, S, lrow, last_head, k, slack, thr, thr0);
      if (thr < pub) atomicMin(gthr, ws_ord(thr));
      const uint32_t c = S->cnt[lrow];
      A.cand_cnt[(size_t)it * WSG_TILE_M + lrow] = c > WSG_CAND_CAP ? 0xFFFFFFFFu : c;
      A.cand_thr[(size_t)it * WSG_TILE_M + lrow] = thr;
    }
  }
  wsg_tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    wsg_tmem_dealloc(tmem_base, 512);
  }
}

// ---- re-rank ---------------------------------------------------------------------------------
// One warp per query: collect the survivors of all of the query's items whose approximate score
// is under T = min over items of the item's final threshold (every dropped point of every item
// has score~ >= T, so k kept points are provably at least as near), compute their exact fp32
// distances with the scan kernel's team-of-8 arithmetic (ws_team_dist_nv: bit-identical to
// ws_scan_task on the same rows), keep the top-k by (distance, rank) and write the final row.

// ws_scan_task over an explicit list of rows (same folding machinery, same arithmetic)
template <int KQ, int METRIC, bool EXACT>
__device__ __forceinline__ int ws_scan_list(const float* vecs, uint32_t dpad, const uint32_t* list, int nl, const float4 (&q)[KQ],
                                            int B, uint64_t* fr, uint64_t* sk, uint64_t* sk2, int* cpos) {
  const int lane = threadIdx.x & 31;
  const int tl = lane & (WS_TEAM - 1), team = lane / WS_TEAM;
  const int dpad4 = dpad >> 2;
  const unsigned lt = (1u << lane) - 1u;
  const bool leader = tl == 0;
  const float4* vtl = reinterpret_cast<const float4*>(vecs) + tl;
  int n = 0, s = 0;
  uint64_t cutoff = WS_KEY_MAX;
  for (int r0 = 0; r0 < nl; r0 += 16) {
    uint32_t r[4];
    bool ok[4];
    float d[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = r0 + 4 * u + team;
      ok[u] = i < nl;
      r[u] = list[ok[u] ? i : nl - 1];
      d[u] = ws_team_dist_nv<KQ, METRIC, EXACT>(vtl + (size_t)r[u] * dpad4, q, tl, dpad4);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint64_t key = ws_key(d[u], r[u]);
      const bool pass = leader && ok[u] && key < cutoff;
      const unsigned bal = __ballot_sync(0xffffffffu, pass);
      if (pass) sk[s + __popc(bal & lt)] = key;
      s += __popc(bal);
    }
    if (s <= 48 && r0 + 16 < nl) continue;
    if (s == 0) continue;
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
    const int mc2 = s;
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
  return n;
}

template <int KQ, int METRIC, bool EXACT>
__global__ void __launch_bounds__(WS_WARPS_PER_CTA * 32) ws_gemm_rerank_kernel(WsGemmRerankArgs A) {
  __shared__ uint64_t s_fr[WS_WARPS_PER_CTA][128];
  __shared__ uint64_t s_sk[WS_WARPS_PER_CTA][64];
  __shared__ uint64_t s_sk2[WS_WARPS_PER_CTA][64];
  __shared__ int s_cpos[WS_WARPS_PER_CTA][64];
  __shared__ uint32_t s_list[WS_WARPS_PER_CTA][WSG_RR_LIST];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tl = lane & (WS_TEAM - 1);
  const uint32_t row = blockIdx.x * WS_WARPS_PER_CTA + warp;
  if (row >= A.rows_pad) return;
  const uint32_t qi = A.perm[row];
  if (qi == 0xFFFFFFFFu) return;
  const uint32_t a = A.row_a[row], b = A.row_b[row];
  const uint32_t g = row / WSG_TILE_M, lr = row % WSG_TILE_M;
  const uint32_t ni = A.group_cnt[g];
  const int B = (int)A.k;
  uint64_t* fr = s_fr[warp];
  uint32_t* list = s_list[warp];
  const unsigned lt = (1u << lane) - 1u;

  float T = __int_as_float(0x7f800000);
  bool ovf = false;
  if (b > a) {
    for (uint32_t j = lane; j < ni; j += 32) {
      const uint32_t it = A.group_items[g * WSG_MAX_SPLITS + j];
      if (A.cand_cnt[(size_t)it * WSG_TILE_M + lr] == 0xFFFFFFFFu) ovf = true;
      T = fminf(T, A.cand_thr[(size_t)it * WSG_TILE_M + lr]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) T = fminf(T, __shfl_xor_sync(0xffffffffu, T, o));
    ovf = __any_sync(0xffffffffu, ovf);
  }
  int nl = 0;
  if (b > a && !ovf) {
    const uint32_t tord = ws_ord(T);
    for (uint32_t j = 0; j < ni; j++) {
      const uint32_t it = A.group_items[g * WSG_MAX_SPLITS + j];
      const uint32_t c = A.cand_cnt[(size_t)it * WSG_TILE_M + lr];
      const uint64_t* cb = A.cand + (size_t)it * WSG_CAND_CAP * WSG_TILE_M + lr;
      for (uint32_t e0 = 0; e0 < c; e0 += 32) {
        const uint32_t e = e0 + lane;
        const uint64_t key = e < c ? cb[(size_t)e * WSG_TILE_M] : WS_KEY_MAX;
        const bool pass = e < c && (uint32_t)(key >> 32) < tord;
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        const int pos = nl + __popc(bal & lt);
        if (pass && pos < WSG_RR_LIST) list[pos] = (uint32_t)key;
        nl += __popc(bal);
      }
    }
    if (nl > WSG_RR_LIST) ovf = true;
  }
  __syncwarp();
  float4 q[KQ];
  ws_load_query_global<KQ, EXACT>(A.queries, A.dim, A.dpad, qi, tl, q);
  if (ovf) {
    // degenerate data: exact streaming scan of the whole window by this warp
    WsScanOut so;
    so.vecs = A.vecs; so.dpad = A.dpad; so.res_keys = A.res_keys; so.res_cnt = A.res_cnt; so.stats = A.stats;
    so.out_ids = A.out_ids; so.out_dists = A.out_dists; so.decode = A.decode; so.pad_id = A.pad_id;
    WsTask task;
    task.query = qi; task.node = -1; task.a = a; task.b = b; task.lo = 0.f; task.hi = 0.f; task.beam = 0; task.flags = WS_TF_SOLO;
    ws_scan_task<KQ, METRIC, EXACT>(so, task, row, q, B, fr, s_sk[warp], s_sk2[warp], s_cpos[warp]);
    if (lane == 0) atomicAdd(A.gstats + 1, 1ull);
    return;
  }
  int n = 0;
  if (nl > 0) n = ws_scan_list<KQ, METRIC, EXACT>(A.vecs, A.dpad, list, nl, q, B, fr, s_sk[warp], s_sk2[warp], s_cpos[warp]);
  for (int j = lane; j < B; j += 32) {
    if (j < n) ws_write_result(A.out_ids, A.out_dists, A.decode, qi, B, j, fr[j]);
    else ws_write_pad(A.out_ids, A.out_dists, A.pad_id, qi, B, j);
  }
  if (lane == 0) {
    atomicAdd(A.stats + WS_ST_SCANPTS, (unsigned long long)(b - a));
    atomicAdd(A.gstats + 0, (unsigned long long)nl);
  }
}

// ---- launchers -------------------------------------------------------------------------------
size_t wsg_topk_smem_bytes() { return 1024 + (size_t)WSG_B_STAGES * WSG_KBLK_BYTES + sizeof(WsGemmSmem); }

cudaError_t wsg_init_attributes() {
  cudaError_t e = cudaFuncSetAttribute(ws_gemm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsg_topk_smem_bytes());
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(ws_gemm_plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WSG_MAX_ROWS * 8);
}
cudaError_t wsg_launch_norm(int grid, cudaStream_t st, const WsGemmNormArgs& a) {
  ws_gemm_norm_kernel<<<grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t wsg_launch_plan(uint32_t nsort, cudaStream_t st, const WsGemmPlanArgs& a) {
  ws_gemm_bounds_kernel<<<(a.nq + 127) / 128, 128, 0, st>>>(a);
  ws_gemm_plan_kernel<<<1, WSG_PLAN_THREADS, nsort * sizeof(uint32_t), st>>>(a);
  return cudaGetLastError();
}
cudaError_t wsg_launch_pack(cudaStream_t st, const WsGemmPackArgs& a) {
  ws_gemm_pack_kernel<<<(a.rows_pad + 7) / 8, 256, 0, st>>>(a);
  return cudaGetLastError();
}
cudaError_t wsg_launch_topk(int grid, cudaStream_t st, const CUtensorMap& tm_b, const WsGemmArgs& a) {
  ws_gemm_topk_kernel<<<grid, WSG_THREADS, wsg_topk_smem_bytes(), st>>>(tm_b, a);
  return cudaGetLastError();
}
cudaError_t wsg_launch_cvt(int grid, cudaStream_t st, const WsGemmCvtArgs& a) {
  ws_gemm_cvt_kernel<<<grid, 256, 0, st>>>(a);
  return cudaGetLastError();
}
template <int KQ, int METRIC>
static cudaError_t wsg_launch_seed_t(bool exact, int grid, cudaStream_t s, const WsGemmSeedArgs& a) {
  if (exact) ws_gemm_seed_kernel<KQ, METRIC, true><<<grid, WS_WARPS_PER_CTA * 32, 0, s>>>(a);
  else ws_gemm_seed_kernel<KQ, METRIC, false><<<grid, WS_WARPS_PER_CTA * 32, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t wsg_launch_seed(int kq, int metric, bool exact, cudaStream_t st, const WsGemmSeedArgs& a) {
  const int grid = (int)((a.rows_pad + WS_WARPS_PER_CTA - 1) / WS_WARPS_PER_CTA);
#define WSG_SD(KQ_)                                                                  \
  case KQ_:                                                                          \
    return metric == 0 ? wsg_launch_seed_t<KQ_, 0>(exact, grid, st, a) : wsg_launch_seed_t<KQ_, 1>(exact, grid, st, a);
  switch (kq) {
    WSG_SD(1) WSG_SD(2) WSG_SD(3) WSG_SD(4) WSG_SD(8) WSG_SD(16)
    default: return cudaErrorInvalidValue;  // dpad <= 512 on this path
  }
#undef WSG_SD
}
template <int KQ, int METRIC>
static cudaError_t wsg_launch_rerank_t(bool exact, int grid, cudaStream_t s, const WsGemmRerankArgs& a) {
  if (exact) ws_gemm_rerank_kernel<KQ, METRIC, true><<<grid, WS_WARPS_PER_CTA * 32, 0, s>>>(a);
  else ws_gemm_rerank_kernel<KQ, METRIC, false><<<grid, WS_WARPS_PER_CTA * 32, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t wsg_launch_rerank(int kq, int metric, bool exact, cudaStream_t st, const WsGemmRerankArgs& a) {
  const int grid = (int)((a.rows_pad + WS_WARPS_PER_CTA - 1) / WS_WARPS_PER_CTA);
#define WSG_RR(KQ_)                                                                  \
  case KQ_:                                                                          \
    return metric == 0 ? wsg_launch_rerank_t<KQ_, 0>(exact, grid, st, a) : wsg_launch_rerank_t<KQ_, 1>(exact, grid, st, a);
  switch (kq) {
    WSG_RR(1) WSG_RR(2) WSG_RR(3) WSG_RR(4) WSG_RR(8) WSG_RR(16)
    default: return cudaErrorInvalidValue;  // dpad <= 512 on this path
  }
#undef WSG_RR
}
