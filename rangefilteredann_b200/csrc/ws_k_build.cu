// ws_k_build.cu — device-side Vamana construction kernels (ws_build.cuh; setup path) for one metric
// (-DWSK_METRIC=<0|1>); the metric-0 object also carries the metric-independent helpers.
#include "ws_launch.h"
#include <cub/device/device_radix_sort.cuh>
namespace {
#include "ws_kernels.cuh"
#include "ws_build.cuh"
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
static cudaError_t insert_t(int grid, size_t smem, cudaStream_t s, const WsBuildArgs& a) {
  cudaError_t e = smem_attr<KQ>().ensure(ws_build_insert_kernel<KQ, WSK_METRIC>, smem);
  if (e != cudaSuccess) return e;
  ws_build_insert_kernel<KQ, WSK_METRIC><<<grid, WS_CTA_THREADS, smem, s>>>(a);
  return cudaGetLastError();
}
template <int KQ>
static cudaError_t insert_occ_t(size_t smem, int* blocks) {
  cudaError_t e = smem_attr<KQ>().ensure(ws_build_insert_kernel<KQ, WSK_METRIC>, smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks, ws_build_insert_kernel<KQ, WSK_METRIC>, WS_CTA_THREADS, smem);
}
template <int KQ>
static cudaError_t reverse_t(int grid, cudaStream_t s, const WsBuildRevArgs& a) {
  ws_build_reverse_kernel<KQ, WSK_METRIC><<<grid, WS_CTA_THREADS, 0, s>>>(a);
  return cudaGetLastError();
}
template <int KQ>
static cudaError_t sort_t(int grid, cudaStream_t s, const WsBuildSortArgs& a) {
  ws_build_sortadj_kernel<KQ, WSK_METRIC><<<grid, WS_CTA_THREADS, 0, s>>>(a);
  return cudaGetLastError();
}

cudaError_t WSK_CAT(wsl_build_insert_m, WSK_METRIC)(int kq, int grid, size_t smem, cudaStream_t s, const WsBuildArgs& a) {
#define WSK_L(KQ_) return insert_t<KQ_>(grid, smem, s, a)
  WS_KQ_SWITCH(kq, WSK_L)
#undef WSK_L
  return cudaErrorInvalidValue;
}
cudaError_t WSK_CAT(wsl_build_insert_occ_m, WSK_METRIC)(int kq, size_t smem, int* blocks) {
#define WSK_L(KQ_) return insert_occ_t<KQ_>(smem, blocks)
  WS_KQ_SWITCH(kq, WSK_L)
#undef WSK_L
  return cudaErrorInvalidValue;
}
cudaError_t WSK_CAT(wsl_build_reverse_m, WSK_METRIC)(int kq, int grid, cudaStream_t s, const WsBuildRevArgs& a) {
#define WSK_L(KQ_) return reverse_t<KQ_>(grid, s, a)
  WS_KQ_SWITCH(kq, WSK_L)
#undef WSK_L
  return cudaErrorInvalidValue;
}
cudaError_t WSK_CAT(wsl_build_sort_m, WSK_METRIC)(int kq, int grid, cudaStream_t s, const WsBuildSortArgs& a) {
#define WSK_L(KQ_) return sort_t<KQ_>(grid, s, a)
  WS_KQ_SWITCH(kq, WSK_L)
#undef WSK_L
  return cudaErrorInvalidValue;
}

#if WSK_METRIC == 0
cudaError_t wsl_build_apply(int grid, cudaStream_t s, const WsBuildArgs& a) {
  ws_build_apply_kernel<<<grid, 256, 0, s>>>(a);
  return cudaGetLastError();
}
cudaError_t wsl_build_heads(int grid, cudaStream_t s, const uint64_t* pairs, uint32_t n, uint32_t* heads, uint32_t* head_count) {
  ws_build_heads_kernel<<<grid, 256, 0, s>>>(pairs, n, heads, head_count);
  return cudaGetLastError();
}
cudaError_t wsl_fill_i32(int grid, cudaStream_t s, int32_t* p, size_t n, int32_t v) {
  ws_fill_i32_kernel<<<grid, 256, 0, s>>>(p, n, v);
  return cudaGetLastError();
}
cudaError_t wsl_sort_keys(void* temp, size_t* temp_bytes, const uint64_t* in, uint64_t* out, int n, cudaStream_t s) {
  return cub::DeviceRadixSort::SortKeys(temp, *temp_bytes, in, out, n, 0, 64, s);
}
#endif
