// LightGCN.set_graph on the device: COO (+ self loops) -> (row, col)-sorted CSR + deg^-1/2.
// Replaces src/models/lightgcn.py:92-112,127-132 (and the sort torch_sparse's SparseTensor
// storage performs on the triplets it is given).  One-time per model load; the sort is
// cub::DeviceRadixSort over the 64-bit key row * N + col.
#include <cub/device/device_radix_sort.cuh>
#include "common.cuh"

namespace {

__global__ void make_keys(const int64_t* __restrict__ er, const int64_t* __restrict__ ec,
                          const float* __restrict__ ew, int64_t m, int64_t n,
                          uint64_t* __restrict__ keys, float* __restrict__ vals, int* __restrict__ bad) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int64_t nnz = m + n;
  if (e >= nnz) return;
  int64_t r, c;
  float w = 1.0f;
  if (e < m) {
    r = er[e]; c = ec[e];
    if (ew) w = ew[e];
    if (r < 0 || r >= n || c < 0 || c >= n) { atomicExch(bad, 1); r = 0; c = 0; }
  } else {
    r = c = e - m;                      // self loop, weight 1 (lightgcn.py:127-132)
  }
  keys[e] = (uint64_t)r * (uint64_t)n + (uint64_t)c;
  if (vals) vals[e] = w;
}

__global__ void split_keys(const uint64_t* __restrict__ keys, int64_t nnz, int64_t n,
                           int32_t* __restrict__ rowptr, int32_t* __restrict__ col) {
  int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= nnz) return;
  uint64_t k = keys[e];
  int64_t r = (int64_t)(k / (uint64_t)n);
  col[e] = (int32_t)(k - (uint64_t)r * (uint64_t)n);
  // every row holds at least its self loop, so each row id appears and starts exactly once
  if (e == 0 || (int64_t)(keys[e - 1] / (uint64_t)n) != r) rowptr[r] = (int32_t)e;
  if (e == nnz - 1) rowptr[n] = (int32_t)nnz;
}

// deg[i] = sum of the row's weights in CSR order (deterministic); dis = deg^-1/2, inf -> 0.
__global__ void degree_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ w, int64_t n,
                              float* __restrict__ dis, int32_t heavy_threshold,
                              int32_t* __restrict__ heavy_rows, int32_t* __restrict__ heavy_count) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int32_t b = rowptr[i], e = rowptr[i + 1];
  float deg;
  if (w) {
    deg = 0.f;
    for (int32_t p = b; p < e; ++p) deg += w[p];
  } else {
    deg = (float)(e - b);
  }
  float d = 1.0f / sqrtf(deg);          // deg.pow(-0.5), lightgcn.py:104
  if (isinf(d)) d = 0.f;                // :105
  dis[i] = d;
  if (e - b > heavy_threshold) {
    int32_t slot = atomicAdd(heavy_count, 1);
    heavy_rows[slot] = (int32_t)i;
  }
}

struct Layout {
  size_t keys_in, keys_out, vals_in, count, bad, cub, total, cub_bytes;
};

Layout plan(int64_t n, int64_t m, int weighted) {
  Layout L{};
  int64_t nnz = n + m;
  auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t off = 0;
  L.keys_in = off; off += up(nnz * sizeof(uint64_t));
  L.keys_out = off; off += up(nnz * sizeof(uint64_t));
  L.vals_in = off; off += weighted ? up(nnz * sizeof(float)) : 0;
  L.count = off; off += 256;
  L.bad = off; off += 256;
  size_t cb = 0;
  if (weighted)
    cub::DeviceRadixSort::SortPairs(nullptr, cb, (uint64_t*)nullptr, (uint64_t*)nullptr, (float*)nullptr,
                                    (float*)nullptr, (int)1, 0, 64);
  else
    cub::DeviceRadixSort::SortKeys(nullptr, cb, (uint64_t*)nullptr, (uint64_t*)nullptr, (int)1, 0, 64);
  // temp storage grows with the item count; query with the real size
  size_t cb2 = 0;
  if (weighted)
    cub::DeviceRadixSort::SortPairs(nullptr, cb2, (uint64_t*)nullptr, (uint64_t*)nullptr, (float*)nullptr,
                                    (float*)nullptr, nnz, 0, 64);
  else
    cub::DeviceRadixSort::SortKeys(nullptr, cb2, (uint64_t*)nullptr, (uint64_t*)nullptr, nnz, 0, 64);
  L.cub_bytes = cb2 > cb ? cb2 : cb;
  L.cub = off; off += up(L.cub_bytes);
  L.total = off;
  return L;
}

}  // namespace

extern "C" size_t hnm_graph_build_workspace_bytes(int64_t num_nodes, int64_t num_edges, int weighted) {
  if (num_nodes <= 0 || num_edges < 0) return 0;
  return plan(num_nodes, num_edges, weighted).total;
}

extern "C" int hnm_graph_build(const int64_t* edge_row, const int64_t* edge_col, const float* edge_w,
                               int64_t num_edges, int64_t num_nodes, int32_t* csr_rowptr, int32_t* csr_col,
                               float* csr_w, float* dis, int32_t heavy_threshold, int32_t* heavy_rows,
                               int32_t* num_heavy_host, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!csr_rowptr || !csr_col || !dis || !heavy_rows || !num_heavy_host || !workspace) return HNM_E_NULL;
  if (num_edges > 0 && (!edge_row || !edge_col)) return HNM_E_NULL;
  if ((edge_w != nullptr) != (csr_w != nullptr)) return HNM_E_NULL;
  if (num_nodes <= 0 || num_edges < 0) return HNM_E_RANGE;
  int64_t nnz = num_nodes + num_edges;
  if (nnz >= (int64_t)INT32_MAX || num_nodes >= (int64_t)INT32_MAX) return HNM_E_RANGE;
  const int weighted = edge_w != nullptr;
  Layout L = plan(num_nodes, num_edges, weighted);
  if (workspace_bytes < L.total) return HNM_E_WORKSPACE;
  char* ws = (char*)workspace;
  uint64_t* keys_in = (uint64_t*)(ws + L.keys_in);
  uint64_t* keys_out = (uint64_t*)(ws + L.keys_out);
  float* vals_in = weighted ? (float*)(ws + L.vals_in) : nullptr;
  int32_t* count = (int32_t*)(ws + L.count);
  int* bad = (int*)(ws + L.bad);
  HNM_CUDA_TRY(cudaMemsetAsync(count, 0, 4, stream));
  HNM_CUDA_TRY(cudaMemsetAsync(bad, 0, 4, stream));
  const int T = 256;
  unsigned grid = (unsigned)((nnz + T - 1) / T);
  make_keys<<<grid, T, 0, stream>>>(edge_row, edge_col, edge_w, num_edges, num_nodes, keys_in, vals_in, bad);
  HNM_LAUNCH_CHECK();
  int end_bit = 1;
  {
    // keys are < N*N
    unsigned __int128 lim = (unsigned __int128)num_nodes * (unsigned __int128)num_nodes;
    while (end_bit < 64 && ((unsigned __int128)1 << end_bit) < lim) ++end_bit;
  }
  size_t cb = L.cub_bytes;
  if (weighted) {
    HNM_CUDA_TRY(cub::DeviceRadixSort::SortPairs(ws + L.cub, cb, keys_in, keys_out, vals_in, csr_w, nnz, 0,
                                                 end_bit, stream));
  } else {
    HNM_CUDA_TRY(cub::DeviceRadixSort::SortKeys(ws + L.cub, cb, keys_in, keys_out, nnz, 0, end_bit, stream));
  }
  split_keys<<<grid, T, 0, stream>>>(keys_out, nnz, num_nodes, csr_rowptr, csr_col);
  HNM_LAUNCH_CHECK();
  degree_kernel<<<(unsigned)((num_nodes + T - 1) / T), T, 0, stream>>>(csr_rowptr, csr_w, num_nodes, dis,
                                                                      heavy_threshold, heavy_rows, count);
  HNM_LAUNCH_CHECK();
  int host_bad = 0;
  HNM_CUDA_TRY(cudaMemcpyAsync(num_heavy_host, count, 4, cudaMemcpyDeviceToHost, stream));
  HNM_CUDA_TRY(cudaMemcpyAsync(&host_bad, bad, 4, cudaMemcpyDeviceToHost, stream));
  HNM_CUDA_TRY(cudaStreamSynchronize(stream));
  if (host_bad) return HNM_E_RANGE;
  return HNM_OK;
}
