// ws_snapshot.inl — an arena on disk (included at the end of wsann.cu).
//
// The reference persists only its graphs (postfilter_vamana.h:54-79, one .bin per node); every process start
// re-sorts the points, re-derives the tree and re-reads up to thousands of graph files.  A snapshot is the finished
// arena as it sits in HBM — padded vectors, labels, id table, node table, adjacency rows, tree geometry — in one
// file that loads with a handful of large sequential reads and host-to-device copies.
//
//   u8[8]  "WSANNSNP"        u32 version (1)        u32 header words that follow (16)
//   u32    metric, dim, dpad, R, label_sorted, has_decode, wst_rows, split, sup_rows, wst_prefilter_nodes, open_tail
//   i32    cutoff, sup_cutoff        u32 max_node_count        u64 n        u64 nodes
//   then sections, each `u64 bytes` + payload, in this order:
//     vectors [n][dpad] f32 | labels [n] f32 | decode [n] u32 (has_decode) | nodes [nodes] {u32 start, u32 count} |
//     adjacency of node 0 .. nodes-1, each [count][R] i32 | wst_nb | wst_off_ptr | wst_node_ptr | wst_off | wst_nodes |
//     sup_size | sup_shift | sup_nb | sup_node_ptr | sup_nodes
// Little-endian, no alignment padding.  Scratch, options and statistics are not part of a snapshot.

namespace {

struct WsFile {
  FILE* f = nullptr;
  ~WsFile() { if (f) std::fclose(f); }
};

static const size_t kSnapChunk = 64ull << 20;

static int ws_snap_write_dev(FILE* f, ws_index* idx, const void* dptr, uint64_t bytes, std::vector<char>& buf) {
  if (std::fwrite(&bytes, 8, 1, f) != 1) return ws_fail(WS_ERR_STATE, "snapshot: short write");
  for (uint64_t off = 0; off < bytes; off += kSnapChunk) {
    const size_t n = (size_t)std::min<uint64_t>(kSnapChunk, bytes - off);
    WS_CUDA(cudaMemcpyAsync(buf.data(), (const char*)dptr + off, n, cudaMemcpyDeviceToHost, idx->stream));
    WS_CUDA(cudaStreamSynchronize(idx->stream));
    if (std::fwrite(buf.data(), 1, n, f) != n) return ws_fail(WS_ERR_STATE, "snapshot: short write");
  }
  return WS_OK;
}

template <class T>
static int ws_snap_write_vec(FILE* f, const std::vector<T>& v) {
  const uint64_t bytes = v.size() * sizeof(T);
  if (std::fwrite(&bytes, 8, 1, f) != 1) return ws_fail(WS_ERR_STATE, "snapshot: short write");
  if (bytes && std::fwrite(v.data(), 1, bytes, f) != bytes) return ws_fail(WS_ERR_STATE, "snapshot: short write");
  return WS_OK;
}

static int ws_snap_read_len(FILE* f, uint64_t* bytes, uint64_t expect, const char* what) {
  if (std::fread(bytes, 8, 1, f) != 1) return ws_fail(WS_ERR_BADARG, "snapshot: truncated before %s", what);
  if (expect != ~0ull && *bytes != expect)
    return ws_fail(WS_ERR_BADARG, "snapshot: section %s has %llu bytes, expected %llu", what, (unsigned long long)*bytes, (unsigned long long)expect);
  return WS_OK;
}

static int ws_snap_read_dev(FILE* f, ws_index* idx, void* dptr, uint64_t bytes, const char* what, std::vector<char>& buf) {
  uint64_t got = 0;
  WS_TRY(ws_snap_read_len(f, &got, bytes, what));
  for (uint64_t off = 0; off < bytes; off += kSnapChunk) {
    const size_t n = (size_t)std::min<uint64_t>(kSnapChunk, bytes - off);
    if (std::fread(buf.data(), 1, n, f) != n) return ws_fail(WS_ERR_BADARG, "snapshot: truncated inside %s", what);
    WS_CUDA(cudaMemcpyAsync((char*)dptr + off, buf.data(), n, cudaMemcpyHostToDevice, idx->stream));
    WS_CUDA(cudaStreamSynchronize(idx->stream));
  }
  return WS_OK;
}

template <class T>
static int ws_snap_read_vec(FILE* f, std::vector<T>& v, const char* what, uint64_t max_bytes) {
  uint64_t bytes = 0;
  WS_TRY(ws_snap_read_len(f, &bytes, ~0ull, what));
  if (bytes % sizeof(T) != 0 || bytes > max_bytes) return ws_fail(WS_ERR_BADARG, "snapshot: bad length of %s", what);
  v.resize(bytes / sizeof(T));
  if (bytes && std::fread(v.data(), 1, bytes, f) != bytes) return ws_fail(WS_ERR_BADARG, "snapshot: truncated inside %s", what);
  return WS_OK;
}

}  // namespace

extern "C" {

int ws_index_save(ws_index* idx, const char* path) {
  WS_NEED_DEVICE(idx);
  if (!path) return ws_fail(WS_ERR_BADARG, "null path");
  if (!idx->finalized) return ws_fail(WS_ERR_STATE, "only a finalized arena can be saved");
  const std::string tmp = std::string(path) + ".tmp";
  {
    WsFile file;
    file.f = std::fopen(tmp.c_str(), "wb");
    if (!file.f) return ws_fail(WS_ERR_BADARG, "cannot write %s", tmp.c_str());
    FILE* f = file.f;
    const uint32_t hdr[16] = {1u, 16u, (uint32_t)idx->metric, idx->dim, idx->dpad, idx->R, idx->label_sorted ? 1u : 0u,
                              idx->d_decode ? 1u : 0u, idx->wst_rows, idx->split, idx->sup_rows, idx->wst_prefilter_nodes ? 1u : 0u,
                              idx->opt_open_tail ? 1u : 0u, (uint32_t)idx->cutoff, (uint32_t)idx->sup_cutoff, idx->max_node_count};
    const uint64_t n = idx->n, nodes = idx->h_nodes.size();
    if (std::fwrite("WSANNSNP", 1, 8, f) != 8 || std::fwrite(hdr, 4, 16, f) != 16 || std::fwrite(&n, 8, 1, f) != 1 ||
        std::fwrite(&nodes, 8, 1, f) != 1)
      return ws_fail(WS_ERR_STATE, "snapshot: short write");
    std::vector<char> buf(kSnapChunk);
    WS_TRY(ws_snap_write_dev(f, idx, idx->d_vecs, n * idx->dpad * sizeof(float), buf));
    WS_TRY(ws_snap_write_dev(f, idx, idx->d_labels, n * sizeof(float), buf));
    if (idx->d_decode) WS_TRY(ws_snap_write_dev(f, idx, idx->d_decode, n * sizeof(uint32_t), buf));
    std::vector<uint32_t> spans(2 * nodes);
    for (size_t i = 0; i < nodes; i++) { spans[2 * i] = idx->h_nodes[i].start; spans[2 * i + 1] = idx->h_nodes[i].count; }
    WS_TRY(ws_snap_write_vec(f, spans));
    for (size_t i = 0; i < nodes; i++) {
      const WsNode& nd = idx->h_nodes[i];
      if (!nd.adj) return ws_fail(WS_ERR_STATE, "node %zu has no adjacency", i);
      WS_TRY(ws_snap_write_dev(f, idx, nd.adj, (uint64_t)nd.count * idx->R * sizeof(int32_t), buf));
    }
    WS_TRY(ws_snap_write_vec(f, idx->wst_nb)); WS_TRY(ws_snap_write_vec(f, idx->wst_off_ptr)); WS_TRY(ws_snap_write_vec(f, idx->wst_node_ptr));
    WS_TRY(ws_snap_write_vec(f, idx->wst_off)); WS_TRY(ws_snap_write_vec(f, idx->wst_nodes));
    WS_TRY(ws_snap_write_vec(f, idx->sup_size)); WS_TRY(ws_snap_write_vec(f, idx->sup_shift)); WS_TRY(ws_snap_write_vec(f, idx->sup_nb));
    WS_TRY(ws_snap_write_vec(f, idx->sup_node_ptr)); WS_TRY(ws_snap_write_vec(f, idx->sup_nodes));
    if (std::fflush(f) != 0) return ws_fail(WS_ERR_STATE, "snapshot: flush failed");
  }
  if (std::rename(tmp.c_str(), path) != 0) return ws_fail(WS_ERR_STATE, "cannot rename %s", tmp.c_str());
  return WS_OK;
}

int ws_index_load(const char* path, int device, ws_index** out) {
  if (!path || !out) return ws_fail(WS_ERR_BADARG, "null argument");
  *out = nullptr;
  WsFile file;
  file.f = std::fopen(path, "rb");
  if (!file.f) return ws_fail(WS_ERR_BADARG, "cannot open %s", path);
  FILE* f = file.f;
  char magic[8];
  uint32_t hdr[16];
  uint64_t n = 0, nodes = 0;
  if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "WSANNSNP", 8) != 0) return ws_fail(WS_ERR_BADARG, "%s is not an arena snapshot", path);
  if (std::fread(hdr, 4, 16, f) != 16 || std::fread(&n, 8, 1, f) != 1 || std::fread(&nodes, 8, 1, f) != 1)
    return ws_fail(WS_ERR_BADARG, "snapshot: truncated header");
  if (hdr[0] != 1u || hdr[1] != 16u) return ws_fail(WS_ERR_BADARG, "snapshot version %u is not supported", hdr[0]);
  const uint32_t metric = hdr[2], dim = hdr[3], dpad = hdr[4], R = hdr[5];
  if (metric > 1 || dim == 0 || dpad != ws_dim_round_up(dim) || dpad > 1024 || n == 0 || n >= (1ull << 31) || R > 64 || (R & 3u) ||
      nodes > (1ull << 28))
    return ws_fail(WS_ERR_BADARG, "snapshot: implausible header (n=%llu dim=%u dpad=%u R=%u nodes=%llu)", (unsigned long long)n, dim, dpad, R,
                   (unsigned long long)nodes);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || device < 0 || device >= ndev)
    return ws_fail(WS_ERR_CUDA, "no usable CUDA device %d; this engine has no CPU fallback", device);
  ws_index* idx = new ws_index();
  struct Guard { ws_index*& p; bool armed = true; ~Guard() { if (armed && p) { ws_index_destroy(p); p = nullptr; } } } guard{idx};
  idx->device = device; idx->metric = (int)metric; idx->n = n; idx->dim = dim; idx->dpad = dpad; idx->R = R;
  idx->label_sorted = hdr[6] != 0; idx->has_decode = hdr[7] != 0; idx->wst_rows = hdr[8]; idx->split = hdr[9]; idx->sup_rows = hdr[10];
  idx->wst_prefilter_nodes = hdr[11] != 0; idx->opt_open_tail = hdr[12] != 0; idx->cutoff = (int32_t)hdr[13]; idx->sup_cutoff = (int32_t)hdr[14];
  idx->max_node_count = hdr[15];
  WS_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  WS_CUDA(cudaGetDeviceProperties(&prop, device));
  idx->num_sms = prop.multiProcessorCount;
  idx->smem_optin = prop.sharedMemPerBlockOptin;
  WS_CUDA(cudaStreamCreateWithFlags(&idx->stream, cudaStreamNonBlocking));
  WS_CUDA(cudaEventCreate(&idx->ev0));
  WS_CUDA(cudaEventCreate(&idx->ev1));
  std::vector<char> buf(kSnapChunk);
  const uint64_t vbytes = n * dpad * sizeof(float);
  WS_CUDA(cudaMalloc(&idx->d_vecs, vbytes));
  idx->hbm_bytes += vbytes;
  WS_TRY(ws_snap_read_dev(f, idx, idx->d_vecs, vbytes, "vectors", buf));
  WS_CUDA(cudaMalloc(&idx->d_labels, n * sizeof(float)));
  idx->hbm_bytes += n * sizeof(float);
  WS_TRY(ws_snap_read_dev(f, idx, idx->d_labels, n * sizeof(float), "labels", buf));
  idx->h_labels.resize(n);
  WS_CUDA(cudaMemcpy(idx->h_labels.data(), idx->d_labels, n * sizeof(float), cudaMemcpyDeviceToHost));
  if (idx->label_sorted)
    for (uint64_t i = 1; i < n; i++)
      if (idx->h_labels[i] < idx->h_labels[i - 1]) return ws_fail(WS_ERR_BADARG, "snapshot: labels not sorted at %llu", (unsigned long long)i);
  if (idx->has_decode) {
    WS_CUDA(cudaMalloc(&idx->d_decode, n * sizeof(uint32_t)));
    idx->hbm_bytes += n * sizeof(uint32_t);
    WS_TRY(ws_snap_read_dev(f, idx, idx->d_decode, n * sizeof(uint32_t), "decode", buf));
  }
  WS_CUDA(cudaMalloc(&idx->d_stats, 8 * sizeof(unsigned long long)));
  WS_CUDA(cudaMemset(idx->d_stats, 0, 8 * sizeof(unsigned long long)));
  WS_CUDA(cudaMalloc(&idx->d_sticky, 4 * sizeof(uint32_t)));
  WS_CUDA(cudaMemset(idx->d_sticky, 0, 4 * sizeof(uint32_t)));
  std::vector<uint32_t> spans;
  WS_TRY(ws_snap_read_vec(f, spans, "nodes", 1ull << 32));
  if (spans.size() != 2 * nodes) return ws_fail(WS_ERR_BADARG, "snapshot: node table length");
  for (uint64_t i = 0; i < nodes; i++) {
    const uint64_t start = spans[2 * i], count = spans[2 * i + 1];
    if (count == 0 || start + count > n || R == 0) return ws_fail(WS_ERR_BADARG, "snapshot: node %llu outside the arena", (unsigned long long)i);
    const size_t bytes = (size_t)count * R * sizeof(int32_t);
    const size_t aligned = (bytes + 255) & ~(size_t)255;
    void* dst = nullptr;
    if (aligned > kAdjSlabBytes) {
      WS_CUDA(cudaMalloc(&dst, aligned));
      idx->adj_slabs.insert(idx->adj_slabs.begin(), dst);
      idx->hbm_bytes += aligned;
      if (idx->adj_slabs.size() == 1) idx->slab_used = kAdjSlabBytes;
    } else {
      if (idx->adj_slabs.empty() || idx->slab_used + aligned > kAdjSlabBytes) {
        void* slab = nullptr;
        WS_CUDA(cudaMalloc(&slab, kAdjSlabBytes));
        idx->adj_slabs.push_back(slab);
        idx->slab_used = 0;
        idx->hbm_bytes += kAdjSlabBytes;
      }
      dst = (char*)idx->adj_slabs.back() + idx->slab_used;
      idx->slab_used += aligned;
    }
    WS_TRY(ws_snap_read_dev(f, idx, dst, bytes, "adjacency", buf));
    WsNode nd;
    nd.adj = (const int32_t*)dst; nd.start = (uint32_t)start; nd.count = (uint32_t)count;
    idx->h_nodes.push_back(nd);
    idx->node_deg.push_back(nullptr);
  }
  const uint64_t cap = 1ull << 34;
  WS_TRY(ws_snap_read_vec(f, idx->wst_nb, "wst_nb", cap)); WS_TRY(ws_snap_read_vec(f, idx->wst_off_ptr, "wst_off_ptr", cap));
  WS_TRY(ws_snap_read_vec(f, idx->wst_node_ptr, "wst_node_ptr", cap)); WS_TRY(ws_snap_read_vec(f, idx->wst_off, "wst_off", cap));
  WS_TRY(ws_snap_read_vec(f, idx->wst_nodes, "wst_nodes", cap));
  WS_TRY(ws_snap_read_vec(f, idx->sup_size, "sup_size", cap)); WS_TRY(ws_snap_read_vec(f, idx->sup_shift, "sup_shift", cap));
  WS_TRY(ws_snap_read_vec(f, idx->sup_nb, "sup_nb", cap)); WS_TRY(ws_snap_read_vec(f, idx->sup_node_ptr, "sup_node_ptr", cap));
  WS_TRY(ws_snap_read_vec(f, idx->sup_nodes, "sup_nodes", cap));
  if (idx->wst_nb.size() != idx->wst_rows || idx->wst_off_ptr.size() != idx->wst_rows || idx->sup_size.size() != idx->sup_rows ||
      idx->sup_nb.size() != idx->sup_rows)
    return ws_fail(WS_ERR_BADARG, "snapshot: geometry tables do not match the header");
  for (int32_t h : idx->wst_nodes)
    if (h >= (int64_t)nodes) return ws_fail(WS_ERR_BADARG, "snapshot: tree refers to node %d of %llu", h, (unsigned long long)nodes);
  for (int32_t h : idx->sup_nodes)
    if (h < 0 || h >= (int64_t)nodes) return ws_fail(WS_ERR_BADARG, "snapshot: tree refers to node %d of %llu", h, (unsigned long long)nodes);
  WS_TRY(ws_index_finalize(idx));
  guard.armed = false;
  *out = idx;
  return WS_OK;
}

int ws_index_shape(const ws_index* idx, uint64_t* n, uint32_t* dim, int* metric, uint64_t* nodes) {
  if (!idx) return ws_fail(WS_ERR_BADARG, "null index");
  if (n) *n = idx->n;
  if (dim) *dim = idx->dim;
  if (metric) *metric = idx->metric;
  if (nodes) *nodes = idx->h_nodes.size();
  return WS_OK;
}

}  // extern "C"
