// ws_k_misc.cu — the metric-independent kernels (K3 decomposition, K4 merges, L2 flush) and the metric dispatch of
// every launcher declared in ws_launch.h.
#include "ws_launch.h"
namespace {
#include "ws_kernels.cuh"
}

cudaError_t wsl_decompose(int grid, cudaStream_t s, const WsDecompArgs& a) {
  ws_decompose_kernel<<<grid, 128, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t wsl_merge(int grid, cudaStream_t s, const WsMergeArgs& a) {
  ws_merge_kernel<<<grid, WS_CTA_THREADS, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t wsl_merge_parts(int grid, cudaStream_t s, const WsMergePartsArgs& a) {
  ws_merge_parts_kernel<<<grid, WS_CTA_THREADS, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t wsl_fill(int grid, cudaStream_t s, uint4* p, size_t n16) {
  ws_fill_kernel<<<grid, 256, 0, s>>>(p, n16);
  return cudaGetLastError();
}

// ---- per-(metric, capacity) objects
#define WSK_DECL_BW(M, CS)                                                                                              \
  cudaError_t wsl_beam_warp_m##M##_##CS(int kq, bool exact, int grid, size_t smem, cudaStream_t s, const WsBeamArgs& a); \
  cudaError_t wsl_beam_warp_occ_m##M##_##CS(int kq, bool exact, size_t smem, int* blocks);
WSK_DECL_BW(0, 7) WSK_DECL_BW(0, 8) WSK_DECL_BW(0, 9) WSK_DECL_BW(0, 10)
WSK_DECL_BW(1, 7) WSK_DECL_BW(1, 8) WSK_DECL_BW(1, 9) WSK_DECL_BW(1, 10)
#define WSK_DECL_M(M)                                                                                                  \
  cudaError_t wsl_beam_cta_m##M(int kq, bool exact, int grid, size_t smem, cudaStream_t s, const WsBeamArgs& a);        \
  cudaError_t wsl_beam_cta_occ_m##M(int kq, bool exact, size_t smem, int* blocks);                                      \
  cudaError_t wsl_scan_m##M(int kq, int grid, size_t smem, cudaStream_t s, const WsScanArgs& a);                        \
  cudaError_t wsl_scan_warp_m##M(int kq, bool exact, int grid, cudaStream_t s, const WsScanArgs& a);                    \
  cudaError_t wsl_scan_warp_occ_m##M(int kq, bool exact, int* blocks);                                                  \
  cudaError_t wsl_prefilter_direct_m##M(int kq, bool exact, int grid, cudaStream_t s, const WsPrefilterDirectArgs& a);  \
  cudaError_t wsl_prefilter_direct_occ_m##M(int kq, bool exact, int* blocks);                                           \
  cudaError_t wsl_build_insert_m##M(int kq, int grid, size_t smem, cudaStream_t s, const WsBuildArgs& a);               \
  cudaError_t wsl_build_insert_occ_m##M(int kq, size_t smem, int* blocks);                                              \
  cudaError_t wsl_build_reverse_m##M(int kq, int grid, cudaStream_t s, const WsBuildRevArgs& a);                        \
  cudaError_t wsl_build_sort_m##M(int kq, int grid, cudaStream_t s, const WsBuildSortArgs& a);
WSK_DECL_M(0) WSK_DECL_M(1)

cudaError_t wsl_beam_warp(int kq, int metric, bool exact, int cs, int grid, size_t smem, cudaStream_t s, const WsBeamArgs& a) {
  switch (cs) {
    case 7: return metric == 0 ? wsl_beam_warp_m0_7(kq, exact, grid, smem, s, a) : wsl_beam_warp_m1_7(kq, exact, grid, smem, s, a);
    case 8: return metric == 0 ? wsl_beam_warp_m0_8(kq, exact, grid, smem, s, a) : wsl_beam_warp_m1_8(kq, exact, grid, smem, s, a);
    case 9: return metric == 0 ? wsl_beam_warp_m0_9(kq, exact, grid, smem, s, a) : wsl_beam_warp_m1_9(kq, exact, grid, smem, s, a);
    case 10: return metric == 0 ? wsl_beam_warp_m0_10(kq, exact, grid, smem, s, a) : wsl_beam_warp_m1_10(kq, exact, grid, smem, s, a);
  }
  return cudaErrorInvalidValue;
}
cudaError_t wsl_beam_warp_occ(int kq, int metric, bool exact, int cs, size_t smem, int* blocks) {
  switch (cs) {
    case 7: return metric == 0 ? wsl_beam_warp_occ_m0_7(kq, exact, smem, blocks) : wsl_beam_warp_occ_m1_7(kq, exact, smem, blocks);
    case 8: return metric == 0 ? wsl_beam_warp_occ_m0_8(kq, exact, smem, blocks) : wsl_beam_warp_occ_m1_8(kq, exact, smem, blocks);
    case 9: return metric == 0 ? wsl_beam_warp_occ_m0_9(kq, exact, smem, blocks) : wsl_beam_warp_occ_m1_9(kq, exact, smem, blocks);
    case 10: return metric == 0 ? wsl_beam_warp_occ_m0_10(kq, exact, smem, blocks) : wsl_beam_warp_occ_m1_10(kq, exact, smem, blocks);
  }
  return cudaErrorInvalidValue;
}
cudaError_t wsl_beam_cta(int kq, int metric, bool exact, int grid, size_t smem, cudaStream_t s, const WsBeamArgs& a) {
  return metric == 0 ? wsl_beam_cta_m0(kq, exact, grid, smem, s, a) : wsl_beam_cta_m1(kq, exact, grid, smem, s, a);
}
cudaError_t wsl_beam_cta_occ(int kq, int metric, bool exact, size_t smem, int* blocks) {
  return metric == 0 ? wsl_beam_cta_occ_m0(kq, exact, smem, blocks) : wsl_beam_cta_occ_m1(kq, exact, smem, blocks);
}
cudaError_t wsl_scan(int kq, int metric, int grid, size_t smem, cudaStream_t s, const WsScanArgs& a) {
  return metric == 0 ? wsl_scan_m0(kq, grid, smem, s, a) : wsl_scan_m1(kq, grid, smem, s, a);
}
cudaError_t wsl_scan_warp(int kq, int metric, bool exact, int grid, cudaStream_t s, const WsScanArgs& a) {
  return metric == 0 ? wsl_scan_warp_m0(kq, exact, grid, s, a) : wsl_scan_warp_m1(kq, exact, grid, s, a);
}
cudaError_t wsl_scan_warp_occ(int kq, int metric, bool exact, int* blocks) {
  return metric == 0 ? wsl_scan_warp_occ_m0(kq, exact, blocks) : wsl_scan_warp_occ_m1(kq, exact, blocks);
}
cudaError_t wsl_prefilter_direct(int kq, int metric, bool exact, int grid, cudaStream_t s, const WsPrefilterDirectArgs& a) {
  return metric == 0 ? wsl_prefilter_direct_m0(kq, exact, grid, s, a) : wsl_prefilter_direct_m1(kq, exact, grid, s, a);
}
cudaError_t wsl_prefilter_direct_occ(int kq, int metric, bool exact, int* blocks) {
  return metric == 0 ? wsl_prefilter_direct_occ_m0(kq, exact, blocks) : wsl_prefilter_direct_occ_m1(kq, exact, blocks);
}
cudaError_t wsl_build_insert(int kq, int metric, int grid, size_t smem, cudaStream_t s, const WsBuildArgs& a) {
  return metric == 0 ? wsl_build_insert_m0(kq, grid, smem, s, a) : wsl_build_insert_m1(kq, grid, smem, s, a);
}
cudaError_t wsl_build_insert_occ(int kq, int metric, size_t smem, int* blocks) {
  return metric == 0 ? wsl_build_insert_occ_m0(kq, smem, blocks) : wsl_build_insert_occ_m1(kq, smem, blocks);
}
cudaError_t wsl_build_reverse(int kq, int metric, int grid, cudaStream_t s, const WsBuildRevArgs& a) {
  return metric == 0 ? wsl_build_reverse_m0(kq, grid, s, a) : wsl_build_reverse_m1(kq, grid, s, a);
}
cudaError_t wsl_build_sort(int kq, int metric, int grid, cudaStream_t s, const WsBuildSortArgs& a) {
  return metric == 0 ? wsl_build_sort_m0(kq, grid, s, a) : wsl_build_sort_m1(kq, grid, s, a);
}
