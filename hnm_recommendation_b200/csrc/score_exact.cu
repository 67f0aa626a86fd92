// Exact (SIMT) scoring entry points:
//   hnm_pair_scores      LightGCN.predict            src/models/lightgcn.py:180-184
//   hnm_score_all_items  LightGCN.predict_all_items  src/models/lightgcn.py:199-202
//   hnm_topk_exact       LightGCN.recommend          src/models/lightgcn.py:345-356 (fp64, canonical ties)
//   hnm_merge_topk       multi-GPU merge of per-shard top-k lists
// hnm_topk_exact is the correctness anchor of the tensor-core path and its fallback for
// users whose certificate fails; it never writes the [users, items] score matrix.
#include <math.h>
#include <algorithm>
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ pair scores
__global__ void pair_scores_kernel(const float* __restrict__ ue, const float* __restrict__ ie,
                                   const int64_t* __restrict__ uids, const int64_t* __restrict__ iids, int64_t batch,
                                   int dim, int64_t nu, int64_t ni, float* __restrict__ out, int* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= batch) return;
  int64_t u = uids[b], i = iids[b];
  if (u < 0) u += nu;   // torch indexing accepts negative indices
  if (i < 0) i += ni;
  if (u < 0 || u >= nu || i < 0 || i >= ni) { if (lane == 0) { if (bad) atomicExch(bad, 1); out[b] = nanf(""); } return; }
  const float* pu = ue + (size_t)u * dim;
  const float* pi = ie + (size_t)i * dim;
  float s = 0.f;
  for (int c = lane; c < dim; c += 32) s = fmaf(pu[c], pi[c], s);
#pragma unroll
  for (int off = 16; off; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  if (lane == 0) out[b] = s;
}

// ------------------------------------------------------------------ all-item scores (fp32 SGEMM, NT)
constexpr int GT = 64;   // output tile edge
constexpr int GK = 16;   // k chunk
__global__ void __launch_bounds__(256)
score_all_kernel(const float* __restrict__ ue, const float* __restrict__ ie, const int64_t* __restrict__ uids,
                 int64_t batch, int64_t ni, int dim, float* __restrict__ out) {
  __shared__ float sa[GK][GT + 1];
  __shared__ float sb[GK][GT + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t u0 = (int64_t)blockIdx.y * GT, i0 = (int64_t)blockIdx.x * GT;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < dim; k0 += GK) {
    // 64 rows x 16 k per operand = 1024 floats, 4 per thread
    for (int t = threadIdx.x; t < GT * GK; t += 256) {
      const int r = t / GK, k = t % GK;
      float a = 0.f, b = 0.f;
      if (k0 + k < dim) {
        if (u0 + r < batch) {
          const int64_t u = uids ? uids[u0 + r] : u0 + r;
          a = ue[(size_t)u * dim + k0 + k];
        }
        if (i0 + r < ni) b = ie[(size_t)(i0 + r) * dim + k0 + k];
      }
      sa[k][r] = a;
      sb[k][r] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) a[m] = sa[k][ty * 4 + m];
#pragma unroll
      for (int n = 0; n < 4; ++n) b[n] = sb[k][tx + 16 * n];
#pragma unroll
      for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int n = 0; n < 4; ++n) acc[m][n] = fmaf(a[m], b[n], acc[m][n]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    const int64_t u = u0 + ty * 4 + m;
    if (u >= batch) continue;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
      const int64_t i = i0 + tx + 16 * n;
      if (i < ni) out[(size_t)u * ni + i] = acc[m][n];
    }
  }
}

// ------------------------------------------------------------------ exact top-k
constexpr int XT = 256;        // threads per CTA
constexpr int XU = 8;          // users per CTA
constexpr int XCAP = 512;      // candidate slots per user; compaction when more than XCAP - XT are used
constexpr int XKMAX = XCAP - XT;

struct Cand {
  double s;
  int64_t id;
};

__device__ __forceinline__ bool excluded(const int64_t* __restrict__ items, int64_t lo, int64_t hi, int64_t id) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = items[mid];
    if (v == id) return true;
    if (v < id) lo = mid + 1; else hi = mid;
  }
  return false;
}

// In-place bitonic sort of n = XCAP entries by (score desc, id asc); all XT threads of the CTA.
__device__ void sort_cands(Cand* c) {
  for (int size = 2; size <= XCAP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < XCAP / 2; t += XT) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;   // first half of each bitonic block ascending in rank (best first)
        Cand a = c[lo], b = c[hi];
        const bool a_first = hnm_before(a.s, a.id, b.s, b.id);
        if (up != a_first) { c[lo] = b; c[hi] = a; }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(XT)
topk_exact_kernel(const float* __restrict__ ue, const float* __restrict__ ie, const int64_t* __restrict__ uids,
                  int64_t batch, int64_t item_begin, int64_t item_end, int dim, int k,
                  const int64_t* __restrict__ excl_ptr, const int64_t* __restrict__ excl_items,
                  int64_t* __restrict__ out_ids, double* __restrict__ out_scores) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Cand* cands = reinterpret_cast<Cand*>(smem_raw);                         // [XU][XCAP]
  double* ud = reinterpret_cast<double*>(cands + XU * XCAP);               // [dim][XU]
  __shared__ int count[XU];
  __shared__ double thr_s[XU];
  __shared__ int64_t thr_id[XU];
  __shared__ int need_compact;

  const int64_t ub = (int64_t)blockIdx.x * XU;
  const int nu = (batch - ub < XU) ? (int)(batch - ub) : XU;
  if (gridDim.y > 1) {     // item range split across blockIdx.y; partial lists are merged afterwards
    const int64_t chunk = (item_end - item_begin + gridDim.y - 1) / gridDim.y;
    const int64_t b0 = item_begin + chunk * blockIdx.y;
    ie += (size_t)(b0 - item_begin) * dim;
    item_end = (b0 + chunk < item_end) ? b0 + chunk : item_end;
    item_begin = b0;
    out_ids += (size_t)blockIdx.y * batch * k;
    out_scores += (size_t)blockIdx.y * batch * k;
  }
  for (int t = threadIdx.x; t < dim * XU; t += XT) {
    const int kk = t / XU, u = t % XU;
    double v = 0.0;
    if (u < nu) {
      const int64_t uid = uids ? uids[ub + u] : ub + u;
      v = (double)ue[(size_t)uid * dim + kk];
    }
    ud[kk * XU + u] = v;
  }
  if (threadIdx.x < XU) {
    count[threadIdx.x] = 0;
    thr_s[threadIdx.x] = -INFINITY;
    thr_id[threadIdx.x] = INT64_MAX;
  }
  if (threadIdx.x == 0) need_compact = 0;
  __syncthreads();

  const bool vec = (dim % 4 == 0);
  for (int64_t base = item_begin; base < item_end; base += XT) {
    const int64_t j = base + threadIdx.x;
    if (j < item_end) {
      double acc[XU];
#pragma unroll
      for (int u = 0; u < XU; ++u) acc[u] = 0.0;
      const float* row = ie + (size_t)(j - item_begin) * dim;
      if (vec) {
        for (int k0 = 0; k0 < dim; k0 += 4) {
          const float4 f = ldg_f4(row + k0);
          const double v[4] = {(double)f.x, (double)f.y, (double)f.z, (double)f.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double* up = ud + (k0 + q) * XU;
#pragma unroll
            for (int u = 0; u < XU; ++u) acc[u] = fma(up[u], v[q], acc[u]);
          }
        }
      } else {
        for (int kk = 0; kk < dim; ++kk) {
          const double v = (double)__ldg(row + kk);
          const double* up = ud + kk * XU;
#pragma unroll
          for (int u = 0; u < XU; ++u) acc[u] = fma(up[u], v, acc[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < XU; ++u) {
        if (u >= nu) break;
        double s = acc[u];
        if (excl_ptr) {
          const int64_t lo = excl_ptr[ub + u], hi = excl_ptr[ub + u + 1];
          if (lo < hi && excluded(excl_items, lo, hi, j)) s = -INFINITY;   // lightgcn.py:349-353
        }
        if (hnm_before(s, j, thr_s[u], thr_id[u])) {
          const int slot = atomicAdd(&count[u], 1);
          cands[u * XCAP + slot] = Cand{s, j};
          if (slot + 1 > XCAP - XT) need_compact = 1;
        }
      }
    }
    __syncthreads();
    if (need_compact) {
      for (int u = 0; u < nu; ++u) {
        const int n = count[u];
        if (n > XCAP - XT) {
          for (int t = n + threadIdx.x; t < XCAP; t += XT) cands[u * XCAP + t] = Cand{-INFINITY, INT64_MAX};
          sort_cands(cands + u * XCAP);
          if (threadIdx.x == 0) {
            count[u] = k;
            thr_s[u] = cands[u * XCAP + k - 1].s;
            thr_id[u] = cands[u * XCAP + k - 1].id;
          }
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) need_compact = 0;
      __syncthreads();
    }
  }
  for (int u = 0; u < nu; ++u) {
    const int n = count[u];
    for (int t = n + threadIdx.x; t < XCAP; t += XT) cands[u * XCAP + t] = Cand{-INFINITY, INT64_MAX};
    sort_cands(cands + u * XCAP);
    for (int t = threadIdx.x; t < k; t += XT) {
      out_ids[(size_t)(ub + u) * k + t] = cands[u * XCAP + t].id;
      out_scores[(size_t)(ub + u) * k + t] = cands[u * XCAP + t].s;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ top-k of materialised score rows
// NeuralCF.recommend (src/models/neural_cf.py:300-326): the logits of a pair are an MLP, not a dot product,
// so the [batch, items] fp32 matrix exists (predict_all_items); this kernel is the `scores[i, filter] = -inf`
// + `torch.topk` tail with the canonical (score desc, id asc) order.  One CTA per row, the same candidate
// buffer + block-wide bitonic compaction as the exact kernel above.
__global__ void __launch_bounds__(XT)
topk_dense_kernel(const float* __restrict__ scores, int64_t num_items, const int64_t* __restrict__ excl_ptr,
                  const int64_t* __restrict__ excl_items, int k, int64_t* __restrict__ out_ids,
                  float* __restrict__ out_scores) {
  __shared__ Cand cands[XCAP];
  __shared__ int count;
  __shared__ double thr_s;
  __shared__ int64_t thr_id;
  const int64_t b = blockIdx.x;
  const float* row = scores + (size_t)b * num_items;
  int64_t ex_lo = 0, ex_hi = 0;
  if (excl_ptr) { ex_lo = excl_ptr[b]; ex_hi = excl_ptr[b + 1]; }
  if (threadIdx.x == 0) { count = 0; thr_s = -INFINITY; thr_id = INT64_MAX; }
  __syncthreads();
  for (int64_t base = 0; base < num_items; base += XT) {
    const int64_t j = base + threadIdx.x;
    if (j < num_items) {
      double s = (double)__ldg(row + j);
      if (s != s) s = -INFINITY;                                  // NaN logits rank last (torch.topk ranks them first)
      if (ex_lo < ex_hi && excluded(excl_items, ex_lo, ex_hi, j)) s = -INFINITY;
      if (hnm_before(s, j, thr_s, thr_id)) cands[atomicAdd(&count, 1)] = Cand{s, j};
    }
    __syncthreads();
    if (count > XCAP - XT) {                                      // block-uniform
      for (int t = count + threadIdx.x; t < XCAP; t += XT) cands[t] = Cand{-INFINITY, INT64_MAX};
      sort_cands(cands);
      if (threadIdx.x == 0) { count = k; thr_s = cands[k - 1].s; thr_id = cands[k - 1].id; }
      __syncthreads();
    }
  }
  for (int t = count + threadIdx.x; t < XCAP; t += XT) cands[t] = Cand{-INFINITY, INT64_MAX};
  sort_cands(cands);
  for (int t = threadIdx.x; t < k; t += XT) {
    out_ids[(size_t)b * k + t] = cands[t].id;
    if (out_scores) out_scores[(size_t)b * k + t] = (float)cands[t].s;
  }
}

// ------------------------------------------------------------------ merge of sorted per-shard lists
constexpr int kMaxMergeShards = 64;   // a list per GPU of one node, or per item split of hnm_topk_exact (<= 64)
__global__ void merge_topk_kernel(const int64_t* __restrict__ in_ids, const double* __restrict__ in_s, int shards,
                                  int64_t batch, int k, int64_t* __restrict__ out_ids,
                                  double* __restrict__ out_s) {
  const int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (b >= batch) return;
  int head[kMaxMergeShards];
  for (int g = 0; g < shards; ++g) head[g] = 0;
  for (int t = 0; t < k; ++t) {
    int best = -1;
    double bs = 0.0;
    int64_t bi = 0;
    for (int g = 0; g < shards; ++g) {
      if (head[g] >= k) continue;
      const size_t off = ((size_t)g * batch + b) * k + head[g];
      const double s = in_s[off];
      const int64_t id = in_ids[off];
      if (best < 0 || hnm_before(s, id, bs, bi)) { best = g; bs = s; bi = id; }
    }
    head[best]++;
    out_ids[(size_t)b * k + t] = bi;
    out_s[(size_t)b * k + t] = bs;
  }
}

}  // namespace

extern "C" int hnm_pair_scores(const float* user_emb, const float* item_emb, const int64_t* user_ids,
                               const int64_t* item_ids, int64_t batch, int32_t dim, int64_t num_users,
                               int64_t num_items, float* out, int32_t* out_of_range, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (batch == 0) return HNM_OK;
  if (!user_emb || !item_emb || !user_ids || !item_ids || !out) return HNM_E_NULL;
  if (batch < 0 || dim <= 0) return HNM_E_RANGE;
  // Asynchronous and capturable: the indices are validated by the caller (engine.pair_scores, like every other
  // entry point); an out-of-range pair yields NaN in its slot (and `out_of_range`, a caller-provided device flag
  // that may be NULL, is raised) instead of a device-side malloc + read-back + stream synchronisation per call.
  const int wpc = 8;
  pair_scores_kernel<<<(unsigned)((batch + wpc - 1) / wpc), wpc * 32, 0, stream>>>(
      user_emb, item_emb, user_ids, item_ids, batch, dim, num_users, num_items, out, out_of_range);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_score_all_items(const float* user_emb, const float* item_emb, const int64_t* user_ids,
                                   int64_t batch, int64_t num_items, int32_t dim, float* scores, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (batch == 0 || num_items == 0) return HNM_OK;
  if (!user_emb || !item_emb || !scores) return HNM_E_NULL;
  if (batch < 0 || num_items < 0 || dim <= 0) return HNM_E_RANGE;
  const int64_t gy = (batch + GT - 1) / GT, gx = (num_items + GT - 1) / GT;
  if (gy > 65535) return HNM_E_RANGE;
  score_all_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, stream>>>(user_emb, item_emb, user_ids, batch,
                                                                       num_items, dim, scores);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_topk_exact(const float* user_emb, const float* item_emb, const int64_t* user_ids, int64_t batch,
                              int64_t item_begin, int64_t item_end, int32_t dim, int32_t k,
                              const int64_t* excl_ptr, const int64_t* excl_items, int32_t item_splits,
                              int64_t* out_ids, double* out_scores, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (batch == 0) return HNM_OK;
  if (!user_emb || !item_emb || !out_ids || !out_scores) return HNM_E_NULL;
  if ((excl_ptr != nullptr) != (excl_items != nullptr)) return HNM_E_NULL;
  if (batch < 0 || dim <= 0 || item_begin < 0 || item_end <= item_begin) return HNM_E_RANGE;
  if (k <= 0 || k > XKMAX || k > item_end - item_begin) return HNM_E_RANGE;
  const size_t smem = sizeof(Cand) * XU * XCAP + sizeof(double) * XU * (size_t)dim;
  if (smem > 200 * 1024) return HNM_E_DIM;
  HNM_CUDA_TRY(hnm_allow_smem(topk_exact_kernel, 200 * 1024));
  const unsigned grid = (unsigned)((batch + XU - 1) / XU);
  const int64_t items = item_end - item_begin;
  if (item_splits < 1 || item_splits > 64 || (item_splits > 1 && items / item_splits < std::max<int64_t>(k, 1)))
    return HNM_E_RANGE;
  topk_exact_kernel<<<dim3(grid, item_splits), XT, smem, stream>>>(user_emb, item_emb, user_ids, batch, item_begin,
                                                                   item_end, dim, k, excl_ptr, excl_items, out_ids,
                                                                   out_scores);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_merge_topk(const int64_t* in_ids, const double* in_scores, int32_t num_shards, int64_t batch,
                              int32_t k, int64_t* out_ids, double* out_scores, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (batch == 0) return HNM_OK;
  if (!in_ids || !in_scores || !out_ids || !out_scores) return HNM_E_NULL;
  if (num_shards <= 0 || num_shards > kMaxMergeShards || batch < 0 || k <= 0) return HNM_E_RANGE;
  merge_topk_kernel<<<(unsigned)((batch + 127) / 128), 128, 0, stream>>>(in_ids, in_scores, num_shards, batch, k,
                                                                       out_ids, out_scores);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_topk_dense(const float* scores, int64_t batch, int64_t num_items, const int64_t* excl_ptr,
                              const int64_t* excl_items, int32_t k, int64_t* out_ids, float* out_scores,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (batch == 0) return HNM_OK;
  if (!scores || !out_ids) return HNM_E_NULL;
  if ((excl_ptr != nullptr) != (excl_items != nullptr)) return HNM_E_NULL;
  if (batch < 0 || batch > INT32_MAX || num_items < 1 || k < 1 || k > XKMAX || k > num_items) return HNM_E_RANGE;
  topk_dense_kernel<<<(unsigned)batch, XT, 0, stream>>>(scores, num_items, excl_ptr, excl_items, k, out_ids, out_scores);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}
