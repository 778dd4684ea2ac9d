// ws_launch.h — launchers of the sm_100a kernels.  The kernels are templates over (floats per lane, metric, …);
// each family is instantiated in its own translation unit (ws_k_*.cu) so that the library builds in parallel,
// and the host orchestration (wsann.cu) only sees these plain functions.
#pragma once
#include <mutex>

#include "ws_args.h"

// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to a (function, device) pair and only ever has to grow.  One of
// these per kernel instantiation keeps it monotone under a lock: host threads of a ws_group launch the same kernel
// with different shared-memory sizes at the same time, possibly on the same device.
struct WsSmemAttr {
  std::mutex mu;
  int cur[64] = {0};
  template <class F>
  cudaError_t ensure(F* fn, size_t smem) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    if ((int)smem <= cur[dev & 63]) return cudaSuccess;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) cur[dev & 63] = (int)smem;
    return e;
  }
};

// KQ = float4 columns per lane of a team of 8 (ws_device.cuh): the instantiated values
#define WS_KQ_SWITCH(KQV, CALL)  \
  switch (KQV) {                 \
    case 1: CALL(1); break;      \
    case 2: CALL(2); break;      \
    case 3: CALL(3); break;      \
    case 4: CALL(4); break;      \
    case 8: CALL(8); break;      \
    case 16: CALL(16); break;    \
    default: CALL(32); break;    \
  }

// K2 warp-per-task beam search; cs = log2(beam capacity) in {7, 8, 9, 10}
cudaError_t wsl_beam_warp(int kq, int metric, bool exact, int cs, int grid, size_t smem, cudaStream_t s, const WsBeamArgs& a);
cudaError_t wsl_beam_warp_occ(int kq, int metric, bool exact, int cs, size_t smem, int* blocks);
// K2 CTA-per-task beam search of the large tier (global visited bitmap)
cudaError_t wsl_beam_cta(int kq, int metric, bool exact, int grid, size_t smem, cudaStream_t s, const WsBeamArgs& a);
cudaError_t wsl_beam_cta_occ(int kq, int metric, bool exact, size_t smem, int* blocks);
// K1 scans
cudaError_t wsl_scan(int kq, int metric, int grid, size_t smem, cudaStream_t s, const WsScanArgs& a);
cudaError_t wsl_scan_warp(int kq, int metric, bool exact, int grid, cudaStream_t s, const WsScanArgs& a);
cudaError_t wsl_scan_warp_occ(int kq, int metric, bool exact, int* blocks);
cudaError_t wsl_prefilter_direct(int kq, int metric, bool exact, int grid, cudaStream_t s, const WsPrefilterDirectArgs& a);
cudaError_t wsl_prefilter_direct_occ(int kq, int metric, bool exact, int* blocks);
// K3 / K4 / plumbing
cudaError_t wsl_decompose(int grid, cudaStream_t s, const WsDecompArgs& a);
cudaError_t wsl_merge(int grid, cudaStream_t s, const WsMergeArgs& a);
cudaError_t wsl_merge_parts(int grid, cudaStream_t s, const WsMergePartsArgs& a);
cudaError_t wsl_fill(int grid, cudaStream_t s, uint4* p, size_t n16);
cudaError_t wsl_fill_i32(int grid, cudaStream_t s, int32_t* p, size_t n, int32_t v);
// device-side Vamana construction (setup path)
cudaError_t wsl_build_insert(int kq, int metric, int grid, size_t smem, cudaStream_t s, const WsBuildArgs& a);
cudaError_t wsl_build_insert_occ(int kq, int metric, size_t smem, int* blocks);
cudaError_t wsl_build_apply(int grid, cudaStream_t s, const WsBuildArgs& a);
cudaError_t wsl_build_heads(int grid, cudaStream_t s, const uint64_t* pairs, uint32_t n, uint32_t* heads, uint32_t* head_count);
cudaError_t wsl_build_reverse(int kq, int metric, int grid, cudaStream_t s, const WsBuildRevArgs& a);
cudaError_t wsl_build_sort(int kq, int metric, int grid, cudaStream_t s, const WsBuildSortArgs& a);
cudaError_t wsl_sort_keys(void* temp, size_t* temp_bytes, const uint64_t* in, uint64_t* out, int n, cudaStream_t s);
