// ws_k_beam_warp.cu — instantiations of ws_beam_warp_kernel for ONE (metric, capacity) pair, selected with
// -DWSK_METRIC=<0|1> -DWSK_CS=<7|8|9|10>: the library compiles this file once per pair, in parallel.
#include "ws_launch.h"
namespace {
#include "ws_kernels.cuh"
}
#if !defined(WSK_METRIC) || !defined(WSK_CS)
#error "compile with -DWSK_METRIC=<0|1> -DWSK_CS=<7..10>"
#endif
#define WSK_CAT_(a, b, c) a##b##_##c
#define WSK_CAT(a, b, c) WSK_CAT_(a, b, c)

template <int KQ, bool EXACT>
static WsSmemAttr& smem_attr() {
  static WsSmemAttr a;  // shared by the launcher and the occupancy query of this instantiation
  return a;
}
template <int KQ, bool EXACT>
static cudaError_t launch_t(int grid, size_t smem, cudaStream_t s, const WsBeamArgs& a) {
  cudaError_t e = smem_attr<KQ, EXACT>().ensure(ws_beam_warp_kernel<KQ, WSK_METRIC, EXACT, WSK_CS>, smem);
  if (e != cudaSuccess) return e;
  ws_beam_warp_kernel<KQ, WSK_METRIC, EXACT, WSK_CS><<<grid, WS_WARPS_PER_CTA * 32, smem, s>>>(a);
  return cudaGetLastError();
}
template <int KQ, bool EXACT>
static cudaError_t occ_t(size_t smem, int* blocks) {
  cudaError_t e = smem_attr<KQ, EXACT>().ensure(ws_beam_warp_kernel<KQ, WSK_METRIC, EXACT, WSK_CS>, smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, ws_beam_warp_kernel<KQ, WSK_METRIC, EXACT, WSK_CS>, WS_WARPS_PER_CTA * 32, smem);
}

cudaError_t WSK_CAT(wsl_beam_warp_m, WSK_METRIC, WSK_CS)(int kq, bool exact, int grid, size_t smem, cudaStream_t s, const WsBeamArgs& a) {
#define WSK_L(KQ_) return exact ? launch_t<KQ_, true>(grid, smem, s, a) : launch_t<KQ_, false>(grid, smem, s, a)
  WS_KQ_SWITCH(kq, WSK_L)
#undef WSK_L
  return cudaErrorInvalidValue;
}
cudaError_t WSK_CAT(wsl_beam_warp_occ_m, WSK_METRIC, WSK_CS)(int kq, bool exact, size_t smem, int* blocks) {
#define WSK_O(KQ_) return exact ? occ_t<KQ_, true>(smem, blocks) : occ_t<KQ_, false>(smem, blocks)
  WS_KQ_SWITCH(kq, WSK_O)
#undef WSK_O
  return cudaErrorInvalidValue;
}
