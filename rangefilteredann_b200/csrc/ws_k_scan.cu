// ws_k_scan.cu — the brute-force kernels (K1 CTA scan, K1 warp scan, K1d one-launch prefilter) for one metric
// (-DWSK_METRIC=<0|1>).
#include "ws_launch.h"
namespace {
#include "ws_kernels.cuh"
}
#if !defined(WSK_METRIC)
#error "compile with -DWSK_METRIC=<0|1>"
#endif
#define WSK_CAT_(a, b) a##b
#define WSK_CAT(a, b) WSK_CAT_(a, b)

template <int KQ>
static WsSmemAttr& smem_attr() {
  static WsSmemAttr a;  // shared by the launcher and the occupancy query of this instantiation
  return a;
}
template <int KQ>
static cudaError_t scan_t(int grid, size_t smem, cudaStream_t s, const WsScanArgs& a) {
  cudaError_t e = smem_attr<KQ>().ensure(ws_scan_kernel<KQ, WSK_METRIC>, smem);
  if (e != cudaSuccess) return e;
  ws_scan_kernel<KQ, WSK_METRIC><<<grid, WS_CTA_THREADS, smem, s>>>(a);
  return cudaGetLastError();
}
template <int KQ, bool EXACT>
static cudaError_t scan_warp_t(int grid, cudaStream_t s, const WsScanArgs& a) {
  ws_scan_warp_kernel<KQ, WSK_METRIC, EXACT><<<grid, WS_WARPS_PER_CTA * 32, 0, s>>>(a);
  return cudaGetLastError();
}
template <int KQ, bool EXACT>
static cudaError_t direct_t(int grid, cudaStream_t s, const WsPrefilterDirectArgs& a) {
  ws_prefilter_direct_kernel<KQ, WSK_METRIC, EXACT><<<grid, WS_WARPS_PER_CTA * 32, 0, s>>>(a);
  return cudaGetLastError();
}
template <int KQ, bool EXACT>
static cudaError_t scan_warp_occ_t(int* blocks) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, ws_scan_warp_kernel<KQ, WSK_METRIC, EXACT>, WS_WARPS_PER_CTA * 32, 0);
}
template <int KQ, bool EXACT>
static cudaError_t direct_occ_t(int* blocks) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, ws_prefilter_direct_kernel<KQ, WSK_METRIC, EXACT>, WS_WARPS_PER_CTA * 32, 0);
}

cudaError_t WSK_CAT(wsl_scan_m, WSK_METRIC)(int kq, int grid, size_t smem, cudaStream_t s, const WsScanArgs& a) {
#define WSK_L(KQ_) return scan_t<KQ_>(grid, smem, s, a)
  WS_KQ_SWITCH(kq, WSK_L)
#undef WSK_L
  return cudaErrorInvalidValue;
}
cudaError_t WSK_CAT(wsl_scan_warp_m, WSK_METRIC)(int kq, bool exact, int grid, cudaStream_t s, const WsScanArgs& a) {
#define WSK_L(KQ_) return exact ? scan_warp_t<KQ_, true>(grid, s, a) : scan_warp_t<KQ_, false>(grid, s, a)
  WS_KQ_SWITCH(kq, WSK_L)
#undef WSK_L
  return cudaErrorInvalidValue;
}
cudaError_t WSK_CAT(wsl_scan_warp_occ_m, WSK_METRIC)(int kq, bool exact, int* blocks) {
#define WSK_L(KQ_) return exact ? scan_warp_occ_t<KQ_, true>(blocks) : scan_warp_occ_t<KQ_, false>(blocks)
  WS_KQ_SWITCH(kq, WSK_L)
#undef WSK_L
  return cudaErrorInvalidValue;
}
cudaError_t WSK_CAT(wsl_prefilter_direct_m, WSK_METRIC)(int kq, bool exact, int grid, cudaStream_t s, const WsPrefilterDirectArgs& a) {
#define WSK_L(KQ_) return exact ? direct_t<KQ_, true>(grid, s, a) : direct_t<KQ_, false>(grid, s, a)
  WS_KQ_SWITCH(kq, WSK_L)
#undef WSK_L
  return cudaErrorInvalidValue;
}
cudaError_t WSK_CAT(wsl_prefilter_direct_occ_m, WSK_METRIC)(int kq, bool exact, int* blocks) {
#define WSK_L(KQ_) return exact ? direct_occ_t<KQ_, true>(blocks) : direct_occ_t<KQ_, false>(blocks)
  WS_KQ_SWITCH(kq, WSK_L)
#undef WSK_L
  return cudaErrorInvalidValue;
}
