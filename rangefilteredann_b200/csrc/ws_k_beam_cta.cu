// ws_k_beam_cta.cu — ws_beam_cta2_kernel, the CTA-per-task beam search of the large tier (beams 1025..12288, visited
// set = a bitmap in global memory), for one metric (-DWSK_METRIC=<0|1>).
#include "ws_launch.h"
namespace {
#include "ws_kernels.cuh"
}
#if !defined(WSK_METRIC)
#error "compile with -DWSK_METRIC=<0|1>"
#endif
#define WSK_CAT_(a, b) a##b
#define WSK_CAT(a, b) WSK_CAT_(a, b)

template <int KQ, bool EXACT>
static WsSmemAttr& smem_attr() {
  static WsSmemAttr a;  // shared by the launcher and the occupancy query of this instantiation
  return a;
}
template <int KQ, bool EXACT>
static cudaError_t launch_t(int grid, size_t smem, cudaStream_t s, const WsBeamArgs& a) {
  cudaError_t e = smem_attr<KQ, EXACT>().ensure(ws_beam_cta2_kernel<KQ, WSK_METRIC, EXACT, true, 14>, smem);
  if (e != cudaSuccess) return e;
  ws_beam_cta2_kernel<KQ, WSK_METRIC, EXACT, true, 14><<<grid, WS_CTA2_THREADS, smem, s>>>(a);
  return cudaGetLastError();
}
template <int KQ, bool EXACT>
static cudaError_t occ_t(size_t smem, int* blocks) {
  cudaError_t e = smem_attr<KQ, EXACT>().ensure(ws_beam_cta2_kernel<KQ, WSK_METRIC, EXACT, true, 14>, smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, ws_beam_cta2_kernel<KQ, WSK_METRIC, EXACT, true, 14>, WS_CTA2_THREADS, smem);
}
cudaError_t WSK_CAT(wsl_beam_cta_m, WSK_METRIC)(int kq, bool exact, int grid, size_t smem, cudaStream_t s, const WsBeamArgs& a) {
#define WSK_L(KQ_) return exact ? launch_t<KQ_, true>(grid, smem, s, a) : launch_t<KQ_, false>(grid, smem, s, a)
  WS_KQ_SWITCH(kq, WSK_L)
#undef WSK_L
  return cudaErrorInvalidValue;
}
cudaError_t WSK_CAT(wsl_beam_cta_occ_m, WSK_METRIC)(int kq, bool exact, size_t smem, int* blocks) {
#define WSK_O(KQ_) return exact ? occ_t<KQ_, true>(smem, blocks) : occ_t<KQ_, false>(smem, blocks)
  WS_KQ_SWITCH(kq, WSK_O)
#undef WSK_O
  return cudaErrorInvalidValue;
}
