// LightGCN propagation: E <- A_hat E with the degree normalisation and the layer sum fused.
// Replaces src/models/lightgcn.py:151-158 (L x torch_sparse SpMM + 2(L+1) elementwise passes).
//
// Formulation (DESIGN.md "propagate"): with xs = dis (.) E kept between layers,
//     (A_hat E)_i = dis_i * sum_{j in row i} w_ij * xs[col_j]
// so the gather loop needs neither a per-edge value array (w_ij = 1 when the reference is
// given edge_weight=None) nor a per-edge dis[col] gather.  HBM-bound integer/gather work:
// one warp per row, D/4 lanes x 128-bit loads per embedding row, several rows in flight per
// lane, column indices read coalesced and broadcast by shuffle.  Rows longer than
// `heavy_threshold` are summed by a cluster of 8 CTAs in a fixed order (deterministic).
#include <algorithm>
#include <stdlib.h>
#include <cooperative_groups.h>
#include "common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kRowsPerWarp = 4;
constexpr int kHeavyThreads = 512;
constexpr int kHugeThreads = 256;
constexpr int kHeavyCluster = 8;

// Where a row's entries live: [b[row - off], e[row - off]).  The whole-row form is {rowptr, rowptr + 1, 0};
// the user-sharded propagation passes per-rank sub-ranges of the item rows (hnm_lightgcn_partial).
struct Seg {
  const int32_t* b;
  const int32_t* e;
  int64_t off;
  // Peer scatter of the partial sums (hnm_lightgcn_partial_peer): local row lr belongs to rank lr / rpo, and
  // this rank's partial for it goes straight into the owner's staging buffer [world][rpo][dim] over NVLink.
  float* const* peer;      // device array of the ranks' staging buffers (mapped peer memory), or nullptr
  int rpo;                 // rows per owner
  int rank;
  __device__ __forceinline__ int beg(int64_t row) const { return __ldg(b + (row - off)); }
  __device__ __forceinline__ int end(int64_t row) const { return __ldg(e + (row - off)); }
  // where row `row` of the output goes (partial mode: a row of the partial-sum table)
  __device__ __forceinline__ float* out_row(float* xs_out, int64_t row, int dim, int partial) const {
    if (!(partial & 1)) return xs_out ? xs_out + (size_t)row * dim : nullptr;
    const int64_t lr = row - off;
    if (peer == nullptr) return xs_out + (size_t)lr * dim;
    const int64_t o = lr / rpo;
    return peer[o] + ((size_t)rank * rpo + (size_t)(lr - o * rpo)) * dim;
  }
};

template <int D>
struct Shape {
  static constexpr int LPR = (D / 4 < 32) ? D / 4 : 32;  // lanes per embedding row
  static constexpr int VEC = D / (4 * LPR);              // float4 per lane per row
  static constexpr int G = 32 / LPR;                     // rows gathered side by side in a warp
  static constexpr int UNROLL_ = (VEC >= 2) ? 4 : 8;     // independent loads in flight per lane = UNROLL*VEC
  static constexpr int UNROLL = (G * UNROLL_ > 32) ? 32 / G : UNROLL_;  // one step never spans more than 32 entries
};

__device__ __forceinline__ void add4(float4& a, const float4& b) {
  a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
}
__device__ __forceinline__ void fma4(float4& a, float w, const float4& b) {
  a.x = fmaf(w, b.x, a.x); a.y = fmaf(w, b.y, a.y); a.z = fmaf(w, b.z, a.z); a.w = fmaf(w, b.w, a.w);
}

// Sum w * xs[col] over CSR entries [beg, end) with one warp.  On return every lane of
// group 0 (lane < LPR) holds the full sum for its float4 slots.
template <int D, bool WEIGHTED>
__device__ __forceinline__ void warp_gather(const int32_t* __restrict__ col, const float* __restrict__ w,
                                            const float* __restrict__ xs, int beg, int end, int lane,
                                            float4 (&acc)[Shape<D>::VEC]) {
  using S = Shape<D>;
  const int g = lane / S::LPR;
  const int sub = lane % S::LPR;
#pragma unroll
  for (int t = 0; t < S::VEC; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = beg; base < end; base += 32) {
    const int idx = base + lane;
    const int c = idx < end ? __ldg(col + idx) : 0;
    float wv = 1.f;
    if (WEIGHTED) wv = idx < end ? __ldg(w + idx) : 0.f;
    const int cnt = min(32, end - base);
    for (int j0 = 0; j0 < cnt; j0 += S::G * S::UNROLL) {
      float4 v[S::UNROLL][S::VEC];
      float wj[S::UNROLL];
#pragma unroll
      for (int u = 0; u < S::UNROLL; ++u) {
        const int j = j0 + u * S::G + g;
        const int cj = __shfl_sync(0xffffffffu, c, j & 31);
        if (WEIGHTED) wj[u] = __shfl_sync(0xffffffffu, wv, j & 31);
        const bool ok = j < cnt;
        const float* p = xs + (size_t)cj * D + sub * 4;
#pragma unroll
        for (int t = 0; t < S::VEC; ++t)
          v[u][t] = ok ? ldg_f4(p + t * S::LPR * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < S::UNROLL; ++u) {
#pragma unroll
        for (int t = 0; t < S::VEC; ++t) {
          if (WEIGHTED) fma4(acc[t], wj[u], v[u][t]);
          else add4(acc[t], v[u][t]);
        }
      }
    }
  }
#pragma unroll
  for (int off = S::LPR; off < 32; off <<= 1) {
#pragma unroll
    for (int t = 0; t < S::VEC; ++t) {
      acc[t].x += __shfl_xor_sync(0xffffffffu, acc[t].x, off);
      acc[t].y += __shfl_xor_sync(0xffffffffu, acc[t].y, off);
      acc[t].z += __shfl_xor_sync(0xffffffffu, acc[t].z, off);
      acc[t].w += __shfl_xor_sync(0xffffffffu, acc[t].w, off);
    }
  }
}

// e = dis_i * sum;  xs_out = dis_i * e;  acc += alpha * e   (lightgcn.py:152,158)
__device__ __forceinline__ void row_epilogue(float4 s, float di, float alpha, float* __restrict__ xs_out_p,
                                             float* __restrict__ acc_p, int partial = 0) {
  if (partial & 1) {      // raw neighbour sum of a sub-range: normalisation happens in hnm_lightgcn_finish
    if (partial & 4)      // ... added to what earlier sub-ranges of the same rows left there
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(xs_out_p), "f"(s.x), "f"(s.y), "f"(s.z), "f"(s.w)
                   : "memory");
    else
      *reinterpret_cast<float4*>(xs_out_p) = s;
    return;
  }
  float4 e = make_float4(__fmul_rn(di, s.x), __fmul_rn(di, s.y), __fmul_rn(di, s.z), __fmul_rn(di, s.w));
  if (xs_out_p) {
    float4 x = make_float4(__fmul_rn(di, e.x), __fmul_rn(di, e.y), __fmul_rn(di, e.z), __fmul_rn(di, e.w));
    *reinterpret_cast<float4*>(xs_out_p) = x;
  }
  if (partial & 2) {
    // acc += alpha * e as a fire-and-forget vector reduction in L2: the warp does not wait for the old
    // value to come back (one writer per element, so the result is the same round-to-nearest add).
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(acc_p), "f"(__fmul_rn(alpha, e.x)),
                 "f"(__fmul_rn(alpha, e.y)), "f"(__fmul_rn(alpha, e.z)), "f"(__fmul_rn(alpha, e.w))
                 : "memory");
    return;
  }
  float4 a = *reinterpret_cast<const float4*>(acc_p);
  a.x = __fadd_rn(a.x, __fmul_rn(alpha, e.x));
  a.y = __fadd_rn(a.y, __fmul_rn(alpha, e.y));
  a.z = __fadd_rn(a.z, __fmul_rn(alpha, e.z));
  a.w = __fadd_rn(a.w, __fmul_rn(alpha, e.w));
  *reinterpret_cast<float4*>(acc_p) = a;
}

template <int D, bool WEIGHTED, int WPC, int RPW>
__global__ void __launch_bounds__(WPC * 32)
spmm_rows_kernel(Seg seg, int partial, const int32_t* __restrict__ col, const float* __restrict__ w,
                 const float* __restrict__ dis, const float* __restrict__ xs_in, float* __restrict__ xs_out,
                 float* __restrict__ accbuf, float alpha, int64_t row_begin, int64_t row_end,
                 int32_t heavy_threshold) {
  using S = Shape<D>;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t first = row_begin + ((int64_t)blockIdx.x * WPC + warp) * RPW;
#pragma unroll 1
  for (int r = 0; r < RPW; ++r) {
    const int64_t row = first + r;
    if (row >= row_end) return;
    const int beg = seg.beg(row), end = seg.end(row);
    if (end - beg > heavy_threshold) continue;  // summed by spmm_heavy_kernel
    float4 acc[S::VEC];
    warp_gather<D, WEIGHTED>(col, w, xs_in, beg, end, lane, acc);
    if (lane < S::LPR) {
      const float di = (partial & 1) ? 1.f : __ldg(dis + row);
      float* orow = seg.out_row(xs_out, row, D, partial);
#pragma unroll
      for (int t = 0; t < S::VEC; ++t) {
        const size_t c = (size_t)(t * S::LPR + lane) * 4;
        row_epilogue(acc[t], di, alpha, orow ? orow + c : nullptr, accbuf + (size_t)row * D + c, partial);
      }
    }
  }
}

// ---- short rows (the user rows: 24 entries on average at the H&M shape) --------------------------------------
// With one row at a time a warp pays two dependent round trips (row pointer -> column indices) before it can
// issue the ~1.5 batches of gathers of an average user row: the user side ran at 10 TB/s where the same gathers
// stream at 19-21 TB/s when the indices are simply there (tools/bench_l2_gather.cu, uniform and Zipf).  Here a
// warp takes 8 consecutive rows: one coalesced load brings their 9 row pointers (and 8 dis values), one sweep
// brings ALL their column indices (contiguous in the CSR) into shared memory, and the rows are then gathered
// back to back with the indices read as shared-memory broadcasts.  Groups with more than kStageCap entries
// (or a long row inside) fall back to the per-row path.
constexpr int kStageRows = 8;
constexpr int kStageCap = 512;

template <int D, bool WEIGHTED>
__device__ __forceinline__ void staged_gather(const int* __restrict__ s_col, const float* __restrict__ s_w,
                                              const float* __restrict__ xs, int cnt, int lane,
                                              float4 (&acc)[Shape<D>::VEC]) {
  using S = Shape<D>;
  const int g = lane / S::LPR;
  const int sub = lane % S::LPR;
#pragma unroll
  for (int t = 0; t < S::VEC; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = 0; j0 < cnt; j0 += S::G * S::UNROLL) {
    float4 v[S::UNROLL][S::VEC];
    float wj[S::UNROLL];
#pragma unroll
    for (int u = 0; u < S::UNROLL; ++u) {
      const int j = j0 + u * S::G + g;
      const bool ok = j < cnt;
      const int cj = s_col[ok ? j : 0];
      if (WEIGHTED) wj[u] = ok ? s_w[j] : 0.f;
      const float* p = xs + (size_t)cj * D + sub * 4;
#pragma unroll
      for (int t = 0; t < S::VEC; ++t)
        v[u][t] = ok ? ldg_f4(p + t * S::LPR * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < S::UNROLL; ++u) {
#pragma unroll
      for (int t = 0; t < S::VEC; ++t) {
        if (WEIGHTED) fma4(acc[t], wj[u], v[u][t]);
        else add4(acc[t], v[u][t]);
      }
    }
  }
#pragma unroll
  for (int off = S::LPR; off < 32; off <<= 1) {
#pragma unroll
    for (int t = 0; t < S::VEC; ++t) {
      acc[t].x += __shfl_xor_sync(0xffffffffu, acc[t].x, off);
      acc[t].y += __shfl_xor_sync(0xffffffffu, acc[t].y, off);
      acc[t].z += __shfl_xor_sync(0xffffffffu, acc[t].z, off);
      acc[t].w += __shfl_xor_sync(0xffffffffu, acc[t].w, off);
    }
  }
}

template <int D, bool WEIGHTED, int WPC>
__global__ void __launch_bounds__(WPC * 32)
spmm_rows_staged_kernel(const int32_t* __restrict__ rowptr, int partial, const int32_t* __restrict__ col,
                        const float* __restrict__ w, const float* __restrict__ dis, const float* __restrict__ xs_in,
                        float* __restrict__ xs_out, float* __restrict__ accbuf, float alpha, int64_t row_begin,
                        int64_t row_end, int32_t heavy_threshold) {
  using S = Shape<D>;
  __shared__ int s_col[WPC][kStageCap];
  __shared__ float s_w[WEIGHTED ? WPC : 1][WEIGHTED ? kStageCap : 1];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int64_t first = row_begin + ((int64_t)blockIdx.x * WPC + warp) * kStageRows;
  if (first >= row_end) return;
  const int nrows = (int)min((int64_t)kStageRows, row_end - first);
  const int rp = lane <= nrows ? __ldg(rowptr + first + lane) : 0;
  const float dv = lane < nrows ? __ldg(dis + first + lane) : 0.f;
  const int beg = __shfl_sync(0xffffffffu, rp, 0);
  const int total = __shfl_sync(0xffffffffu, rp, nrows) - beg;
  const bool staged = total <= kStageCap;
  if (staged) {
    for (int i = lane; i < total; i += 32) {
      s_col[warp][i] = __ldg(col + beg + i);
      if (WEIGHTED) s_w[warp][i] = __ldg(w + beg + i);
    }
    __syncwarp();
  }
  const Seg seg{rowptr, rowptr + 1, 0, nullptr, 1, 0};
  if (staged) {
    // G = 32 / LPR rows side by side, one per lane group (d = 64: one row per half warp): every group walks its
    // own row with UNROLL gathers in flight per lane, so a 24-entry row is three full batches of 8 instead of
    // one and a half batches of 16, and no partial sums cross the groups.
    const int g = lane / S::LPR;
    const int sub = lane % S::LPR;
#pragma unroll 1
    for (int r0 = 0; r0 < nrows; r0 += S::G) {
      const int r = min(r0 + g, nrows - 1);                    // (a group past the last row repeats it, unused)
      const int b = __shfl_sync(0xffffffffu, rp, r), e = __shfl_sync(0xffffffffu, rp, r + 1);
      const bool mine = r0 + g < nrows && e - b <= heavy_threshold;     // long rows: summed by the long-row kernels
      const int cnt = mine ? e - b : 0;
      int maxcnt = cnt;
#pragma unroll
      for (int off = S::LPR; off < 32; off <<= 1) maxcnt = max(maxcnt, __shfl_xor_sync(0xffffffffu, maxcnt, off));
      const int* sc = s_col[warp] + (b - beg);
      float4 acc[S::VEC];
#pragma unroll
      for (int t = 0; t < S::VEC; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j0 = 0; j0 < maxcnt; j0 += S::UNROLL) {
        float4 v[S::UNROLL][S::VEC];
        float wj[S::UNROLL];
#pragma unroll
        for (int u = 0; u < S::UNROLL; ++u) {
          const int j = j0 + u;
          const bool ok = j < cnt;
          const int cj = sc[ok ? j : 0];
          if (WEIGHTED) wj[u] = ok ? s_w[warp][(b - beg) + j] : 0.f;
          const float* p = xs_in + (size_t)cj * D + sub * 4;
#pragma unroll
          for (int t = 0; t < S::VEC; ++t)
            v[u][t] = ok ? ldg_f4(p + t * S::LPR * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < S::UNROLL; ++u) {
#pragma unroll
          for (int t = 0; t < S::VEC; ++t) {
            if (WEIGHTED) fma4(acc[t], wj[u], v[u][t]);
            else add4(acc[t], v[u][t]);
          }
        }
      }
      const float di = __shfl_sync(0xffffffffu, dv, r);
      if (mine) {
        const int64_t row = first + r;
        float* orow = seg.out_row(xs_out, row, D, partial);
#pragma unroll
        for (int t = 0; t < S::VEC; ++t) {
          const size_t c = (size_t)(t * S::LPR + sub) * 4;
          row_epilogue(acc[t], di, alpha, orow ? orow + c : nullptr, accbuf + (size_t)row * D + c, partial);
        }
      }
    }
    return;
  }
#pragma unroll 1
  for (int r = 0; r < nrows; ++r) {
    const int b = __shfl_sync(0xffffffffu, rp, r), e = __shfl_sync(0xffffffffu, rp, r + 1);
    if (e - b > heavy_threshold) continue;                    // summed by the long-row kernels
    const int64_t row = first + r;
    float4 acc[S::VEC];
    warp_gather<D, WEIGHTED>(col, w, xs_in, b, e, lane, acc);
    const float di = __shfl_sync(0xffffffffu, dv, r);
    if (lane < S::LPR) {
      float* orow = seg.out_row(xs_out, row, D, partial);
#pragma unroll
      for (int t = 0; t < S::VEC; ++t) {
        const size_t c = (size_t)(t * S::LPR + lane) * 4;
        row_epilogue(acc[t], di, alpha, orow ? orow + c : nullptr, accbuf + (size_t)row * D + c, partial);
      }
    }
  }
}

// Long rows (more than heavy_threshold entries): one 512-thread CTA per row, warps take contiguous
// 32-aligned chunks (coalesced col[] reads) and the partial sums are added in warp order.
template <int D, bool WEIGHTED>
__global__ void __launch_bounds__(kHeavyThreads)
spmm_heavy_kernel(Seg seg, int partial, const int32_t* __restrict__ col, const float* __restrict__ w,
                  const float* __restrict__ dis, const float* __restrict__ xs_in, float* __restrict__ xs_out,
                  float* __restrict__ accbuf, float alpha, const int32_t* __restrict__ heavy_rows,
                  int64_t row_begin, int64_t row_end) {
  using S = Shape<D>;
  constexpr int NW = kHeavyThreads / 32;
  __shared__ float4 part[NW][D / 4];
  const int64_t row = heavy_rows[blockIdx.x];
  if (row < row_begin || row >= row_end) return;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int beg = seg.beg(row), end = seg.end(row);
  const int per = ((end - beg + NW - 1) / NW + 31) & ~31;
  const int b = min(end, beg + warp * per), e = min(end, b + per);
  float4 acc[S::VEC];
  warp_gather<D, WEIGHTED>(col, w, xs_in, b, e, lane, acc);
  if (lane < S::LPR) {
#pragma unroll
    for (int t = 0; t < S::VEC; ++t) part[warp][t * S::LPR + lane] = acc[t];
  }
  __syncthreads();
  if (threadIdx.x < D / 4) {
    float4 s = part[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < NW; ++k) add4(s, part[k][threadIdx.x]);
    float* orow = seg.out_row(xs_out, row, D, partial);
    const size_t c = (size_t)threadIdx.x * 4;
    row_epilogue(s, (partial & 1) ? 1.f : __ldg(dis + row), alpha, orow ? orow + c : nullptr,
                 accbuf + (size_t)row * D + c, partial);
  }
}

// Very long rows (more than HNM_HUGE_ROW entries; popular items reach ~45 k at the H&M shape): one
// thread-block CLUSTER of kHeavyCluster CTAs per row.  Every CTA sums a contiguous slice, CTA 0 then adds
// the slices in rank order through distributed shared memory (fixed order: deterministic).  Without it
// the single longest row is the critical path of a layer once the work is split over 8 GPUs.
template <int D, bool WEIGHTED>
__global__ void __cluster_dims__(kHeavyCluster, 1, 1) __launch_bounds__(kHugeThreads)
spmm_huge_kernel(Seg seg, int partial, const int32_t* __restrict__ col, const float* __restrict__ w,
                  const float* __restrict__ dis, const float* __restrict__ xs_in, float* __restrict__ xs_out,
                  float* __restrict__ accbuf, float alpha, const int32_t* __restrict__ heavy_rows,
                  int64_t row_begin, int64_t row_end) {
  using S = Shape<D>;
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int NW = kHugeThreads / 32;
  __shared__ float4 part[NW][D / 4];
  __shared__ float4 cta_sum[D / 4];
  const unsigned crank = cluster.block_rank();
  const int64_t row = heavy_rows[blockIdx.x / kHeavyCluster];
  if (row < row_begin || row >= row_end) return;          // uniform over the whole cluster
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int beg = seg.beg(row), end = seg.end(row);
  constexpr int NWT = NW * kHeavyCluster;                  // warps working on this row
  const int per = ((end - beg + NWT - 1) / NWT + 31) & ~31;
  const int gw = (int)crank * NW + warp;
  const int b = min(end, beg + gw * per), e = min(end, b + per);
  float4 acc[S::VEC];
  warp_gather<D, WEIGHTED>(col, w, xs_in, b, e, lane, acc);
  if (lane < S::LPR) {
#pragma unroll
    for (int t = 0; t < S::VEC; ++t) part[warp][t * S::LPR + lane] = acc[t];
  }
  __syncthreads();
  if (threadIdx.x < D / 4) {
    float4 s = part[0][threadIdx.x];
#pragma unroll
    for (int k = 1; k < NW; ++k) add4(s, part[k][threadIdx.x]);
    cta_sum[threadIdx.x] = s;
  }
  cluster.sync();
  if (crank == 0 && threadIdx.x < D / 4) {
    float4 s = cta_sum[threadIdx.x];
    for (unsigned r = 1; r < kHeavyCluster; ++r) {
      const float4* remote = cluster.map_shared_rank(cta_sum, r);
      add4(s, remote[threadIdx.x]);
    }
    float* orow = seg.out_row(xs_out, row, D, partial);
    const size_t c = (size_t)threadIdx.x * 4;
    row_epilogue(s, (partial & 1) ? 1.f : __ldg(dis + row), alpha, orow ? orow + c : nullptr,
                 accbuf + (size_t)row * D + c, partial);
  }
  cluster.sync();                                          // keep every CTA's shared memory alive until read
}

// Any dimension: one warp per row, lanes stride over the columns, edges in order.
template <bool WEIGHTED>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
spmm_generic_kernel(Seg seg, int partial, const int32_t* __restrict__ col, const float* __restrict__ w,
                    const float* __restrict__ dis, const float* __restrict__ xs_in, float* __restrict__ xs_out,
                    float* __restrict__ accbuf, float alpha, int dim, int64_t row_begin, int64_t row_end) {
  const int lane = threadIdx.x & 31;
  const int64_t row = row_begin + (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= row_end) return;
  const int beg = seg.beg(row), end = seg.end(row);
  const float di = partial ? 1.f : dis[row];
  for (int c0 = 0; c0 < dim; c0 += 32 * 8) {
    float s[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) s[t] = 0.f;
    for (int p = beg; p < end; ++p) {
      const float* x = xs_in + (size_t)col[p] * dim;
      const float wv = WEIGHTED ? w[p] : 1.f;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int c = c0 + t * 32 + lane;
        if (c < dim) s[t] = WEIGHTED ? fmaf(wv, __ldg(x + c), s[t]) : s[t] + __ldg(x + c);
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int c = c0 + t * 32 + lane;
      if (c < dim) {
        if (partial) {
          float* o = xs_out + (size_t)(row - seg.off) * dim + c;
          *o = (partial & 4) ? *o + s[t] : s[t];
          continue;
        }
        const size_t off = (size_t)row * dim + c;
        const float e = __fmul_rn(di, s[t]);
        if (xs_out) xs_out[off] = __fmul_rn(di, e);
        accbuf[off] = __fadd_rn(accbuf[off], __fmul_rn(alpha, e));
      }
    }
  }
}

__global__ void prescale_kernel(const float* __restrict__ e0, const float* __restrict__ dis, float alpha0,
                                float* __restrict__ xs, float* __restrict__ acc, int64_t total, int dim) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = e0[i];
    const float di = dis[i / dim];
    xs[i] = __fmul_rn(di, v);
    acc[i] = __fmul_rn(alpha0, v);   // 0 + alpha0 * E0 == alpha0 * E0 exactly (lightgcn.py:156-158)
  }
}

__global__ void prescale_kernel_v4(const float4* __restrict__ e0, const float* __restrict__ dis, float alpha0,
                                   float4* __restrict__ xs, float4* __restrict__ acc, int64_t total4, int dim4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = e0[i];
    const float di = __ldg(dis + i / dim4);
    xs[i] = make_float4(__fmul_rn(di, v.x), __fmul_rn(di, v.y), __fmul_rn(di, v.z), __fmul_rn(di, v.w));
    acc[i] = make_float4(__fmul_rn(alpha0, v.x), __fmul_rn(alpha0, v.y), __fmul_rn(alpha0, v.z),
                         __fmul_rn(alpha0, v.w));
  }
}

struct SideStream {
  cudaStream_t stream;
  cudaEvent_t fork, join;
};

// one side stream + event pair per device, created on first use
SideStream* side_stream() {
  static SideStream per_device[16];
  static bool ready[16] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
  if (!ready[dev]) {
    SideStream s{};
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    per_device[dev] = s;
    ready[dev] = true;
  }
  return &per_device[dev];
}

template <int D, bool WEIGHTED>
int launch_layer(Seg seg, int partial, const int32_t* col, const float* w, const float* dis, const float* xs_in,
                 float* xs_out, float* acc, float alpha, int64_t row_begin, int64_t row_end,
                 const int32_t* heavy_rows, int32_t num_heavy, int32_t num_huge, int32_t heavy_threshold,
                 cudaStream_t stream, int short_rows = 0) {
  // The long rows run on a side stream next to the warp-per-row kernel (they touch disjoint output
  // rows); fork/join with events so the caller still sees one stream-ordered operation.
  cudaStream_t main_stream = stream;
  SideStream* side = num_heavy > 0 ? side_stream() : nullptr;
  if (side) {
    HNM_CUDA_TRY(cudaEventRecord(side->fork, main_stream));
    HNM_CUDA_TRY(cudaStreamWaitEvent(side->stream, side->fork, 0));
    stream = side->stream;
  }
  if (num_huge > 0) {
    spmm_huge_kernel<D, WEIGHTED><<<num_huge * kHeavyCluster, kHugeThreads, 0, stream>>>(seg, partial, col, w, dis, xs_in, xs_out,
                                                                                      acc, alpha, heavy_rows, row_begin,
                                                                                      row_end);
    HNM_LAUNCH_CHECK();
  }
  if (num_heavy > num_huge) {
    spmm_heavy_kernel<D, WEIGHTED><<<num_heavy - num_huge, kHeavyThreads, 0, stream>>>(
        seg, partial, col, w, dis, xs_in, xs_out, acc, alpha, heavy_rows + num_huge, row_begin, row_end);
    HNM_LAUNCH_CHECK();
  }
  if (side) {
    HNM_CUDA_TRY(cudaEventRecord(side->join, side->stream));
    stream = main_stream;
  }
  const int64_t rows = row_end - row_begin;
  static const int variant = getenv("HNM_SPMM_VARIANT") ? atoi(getenv("HNM_SPMM_VARIANT")) : 0;
  const int32_t thr = num_heavy > 0 ? heavy_threshold : INT32_MAX;
#define HNM_ROWS(WPC, RPW)                                                                                   \
  {                                                                                                          \
    const int64_t per_cta = (int64_t)(WPC) * (RPW);                                                          \
    const unsigned grid = (unsigned)((rows + per_cta - 1) / per_cta);                                        \
    if (grid > 0)                                                                                            \
      spmm_rows_kernel<D, WEIGHTED, WPC, RPW><<<grid, (WPC)*32, 0, stream>>>(seg, partial, col, w, dis, xs_in, xs_out, acc, \
                                                                             alpha, row_begin, row_end, thr); \
  }
  if (short_rows && !(partial & 1) && seg.e == seg.b + 1 && seg.off == 0 && variant == 0) {
    constexpr int WPC = 4;
    const int64_t per_cta = (int64_t)WPC * kStageRows;
    const unsigned grid = (unsigned)((rows + per_cta - 1) / per_cta);
    if (grid > 0)
      spmm_rows_staged_kernel<D, WEIGHTED, WPC><<<grid, WPC * 32, 0, stream>>>(seg.b, partial, col, w, dis, xs_in, xs_out,
                                                                               acc, alpha, row_begin, row_end, thr);
    HNM_LAUNCH_CHECK();
    if (side) HNM_CUDA_TRY(cudaStreamWaitEvent(main_stream, side->join, 0));
    return HNM_OK;
  }
  switch (variant) {
    // small CTAs retire (and free their warp slots) at a finer grain: 2 warps x 4 rows measured best
    // at the H&M shape (6.28 ms / 3 layers vs 7.32 ms for 8 x 4; profiles/r1_spmm_notes.md)
    case 1: HNM_ROWS(4, 4); break;
    case 2: HNM_ROWS(8, 4); break;
    case 3: HNM_ROWS(4, 1); break;
    default: HNM_ROWS(2, 4); break;
  }
#undef HNM_ROWS
  HNM_LAUNCH_CHECK();
  if (side) HNM_CUDA_TRY(cudaStreamWaitEvent(main_stream, side->join, 0));
  return HNM_OK;
}

template <bool WEIGHTED>
int dispatch_layer(int dim, Seg seg, int partial, const int32_t* col, const float* w, const float* dis,
                   const float* xs_in, float* xs_out, float* acc, float alpha, int64_t row_begin, int64_t row_end,
                   const int32_t* heavy_rows, int32_t num_heavy, int32_t num_huge, int32_t heavy_threshold,
                   cudaStream_t stream, int short_rows = 0) {
#define HNM_CASE(DD)                                                                                          \
  case DD:                                                                                                    \
    return launch_layer<DD, WEIGHTED>(seg, partial, col, w, dis, xs_in, xs_out, acc, alpha, row_begin, row_end,   \
                                      heavy_rows, num_heavy, num_huge, heavy_threshold, stream, short_rows)
  switch (dim) {
    HNM_CASE(8);
    HNM_CASE(16);
    HNM_CASE(32);
    HNM_CASE(64);
    HNM_CASE(128);
    HNM_CASE(256);
    default: break;
  }
#undef HNM_CASE
  if (dim > 256 * 8) return HNM_E_DIM;
  const int64_t rows = row_end - row_begin;
  const unsigned grid = (unsigned)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
  if (grid > 0) {
    spmm_generic_kernel<WEIGHTED><<<grid, kWarpsPerCta * 32, 0, stream>>>(seg, partial & 5, col, w, dis, xs_in, xs_out, acc,
                                                                       alpha, dim, row_begin, row_end);
    HNM_LAUNCH_CHECK();
  }
  return HNM_OK;
}

}  // namespace

extern "C" int hnm_lightgcn_prescale(const float* e0, const float* dis, float alpha0, float* xs, float* acc,
                                     int64_t num_rows, int32_t dim, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!e0 || !dis || !xs || !acc) return HNM_E_NULL;
  if (num_rows <= 0 || dim <= 0) return HNM_E_RANGE;
  const int64_t total = num_rows * dim;
  const int T = 256;
  const int max_grid = hnm_num_sms() * 16;
  if (dim % 4 == 0 && hnm_aligned16(e0) && hnm_aligned16(xs) && hnm_aligned16(acc)) {
    const int64_t t4 = total / 4;
    const unsigned grid = (unsigned)std::min<int64_t>((t4 + T - 1) / T, max_grid);
    prescale_kernel_v4<<<grid, T, 0, stream>>>((const float4*)e0, dis, alpha0, (float4*)xs, (float4*)acc, t4, dim / 4);
  } else {
    const unsigned grid = (unsigned)std::min<int64_t>((total + T - 1) / T, max_grid);
    prescale_kernel<<<grid, T, 0, stream>>>(e0, dis, alpha0, xs, acc, total, dim);
  }
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_lightgcn_layer(const int32_t* csr_rowptr, const int32_t* csr_col, const float* csr_w,
                                  const float* dis, const float* xs_in, float* xs_out, float* acc, float alpha,
                                  int64_t num_nodes, int32_t dim, int64_t row_begin, int64_t row_end,
                                  const int32_t* heavy_rows, int32_t num_heavy, int32_t num_huge,
                                  int32_t heavy_threshold, int32_t short_rows, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!csr_rowptr || !csr_col || !dis || !xs_in || !acc) return HNM_E_NULL;
  if (num_heavy > 0 && !heavy_rows) return HNM_E_NULL;
  if (num_huge < 0 || num_huge > num_heavy) return HNM_E_RANGE;
  if (num_nodes <= 0 || dim <= 0 || row_begin < 0 || row_end > num_nodes || row_begin > row_end) return HNM_E_RANGE;
  if (xs_in == xs_out) return HNM_E_RANGE;  // rows are gathered while others are written
  if (dim % 4 == 0 && !(hnm_aligned16(xs_in) && hnm_aligned16(acc) && (!xs_out || hnm_aligned16(xs_out))))
    return HNM_E_ALIGN;
  if (row_begin == row_end) return HNM_OK;
  // bit 1 of the mode word: layer sum by red.global.add instead of load + add + store (A/B switch)
  static const int red_mode = (getenv("HNM_SPMM_RED") ? atoi(getenv("HNM_SPMM_RED")) : 1) ? 2 : 0;
  if (csr_w)
    return dispatch_layer<true>(dim, Seg{csr_rowptr, csr_rowptr + 1, 0, nullptr, 1, 0}, red_mode, csr_col, csr_w, dis, xs_in, xs_out, acc, alpha, row_begin, row_end,
                                heavy_rows, num_heavy, num_huge, heavy_threshold, stream, short_rows);
  return dispatch_layer<false>(dim, Seg{csr_rowptr, csr_rowptr + 1, 0, nullptr, 1, 0}, red_mode, csr_col, csr_w, dis, xs_in, xs_out, acc, alpha, row_begin, row_end,
                               heavy_rows, num_heavy, num_huge, heavy_threshold, stream, short_rows);
}


namespace {
__global__ void finish_kernel(const float4* __restrict__ partial, const float4* __restrict__ xs_in,
                              const float* __restrict__ dis, float alpha, float4* __restrict__ xs_out,
                              float4* __restrict__ acc, int64_t row_begin, int64_t total4, int dim4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t g = row_begin * dim4 + i;               // position in the full [N, dim/4] tables
    const float di = __ldg(dis + row_begin + i / dim4);
    float4 s = partial[i];
    const float4 self = xs_in[g];                          // the self loop (weight 1, lightgcn.py:127-132)
    s.x += self.x; s.y += self.y; s.z += self.z; s.w += self.w;
    float4 e = make_float4(__fmul_rn(di, s.x), __fmul_rn(di, s.y), __fmul_rn(di, s.z), __fmul_rn(di, s.w));
    if (xs_out) xs_out[g] = make_float4(__fmul_rn(di, e.x), __fmul_rn(di, e.y), __fmul_rn(di, e.z), __fmul_rn(di, e.w));
    float4 a = acc[g];
    a.x = __fadd_rn(a.x, __fmul_rn(alpha, e.x));
    a.y = __fadd_rn(a.y, __fmul_rn(alpha, e.y));
    a.z = __fadd_rn(a.z, __fmul_rn(alpha, e.z));
    a.w = __fadd_rn(a.w, __fmul_rn(alpha, e.w));
    acc[g] = a;
  }
}
}  // namespace

extern "C" int hnm_lightgcn_partial(const int32_t* seg_begin, const int32_t* seg_end, const int32_t* csr_col,
                                    const float* csr_w, const float* xs_in, float* partial, int32_t dim,
                                    int64_t row_begin, int64_t row_end, const int32_t* heavy_rows, int32_t num_heavy,
                                    int32_t num_huge, int32_t heavy_threshold, int32_t accumulate, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!seg_begin || !seg_end || !csr_col || !xs_in || !partial) return HNM_E_NULL;
  const int mode = accumulate ? 5 : 1;
  if (num_heavy > 0 && !heavy_rows) return HNM_E_NULL;
  if (num_huge < 0 || num_huge > num_heavy) return HNM_E_RANGE;
  if (dim <= 0 || row_begin < 0 || row_begin > row_end) return HNM_E_RANGE;
  if (dim % 4 == 0 && !(hnm_aligned16(xs_in) && hnm_aligned16(partial))) return HNM_E_ALIGN;
  if (row_begin == row_end) return HNM_OK;
  const Seg seg{seg_begin, seg_end, row_begin, nullptr, 1, 0};
  if (csr_w)
    return dispatch_layer<true>(dim, seg, mode, csr_col, csr_w, nullptr, xs_in, partial, nullptr, 0.f, row_begin, row_end,
                                heavy_rows, num_heavy, num_huge, heavy_threshold, stream);
  return dispatch_layer<false>(dim, seg, mode, csr_col, csr_w, nullptr, xs_in, partial, nullptr, 0.f, row_begin, row_end,
                               heavy_rows, num_heavy, num_huge, heavy_threshold, stream);
}

extern "C" int hnm_lightgcn_finish(const float* partial, const float* xs_in, const float* dis, float alpha,
                                   float* xs_out, float* acc, int64_t row_begin, int64_t num_rows, int32_t dim,
                                   void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!partial || !xs_in || !dis || !acc) return HNM_E_NULL;
  if (num_rows < 0 || row_begin < 0 || dim <= 0 || dim % 4 != 0) return HNM_E_RANGE;
  if (!(hnm_aligned16(partial) && hnm_aligned16(xs_in) && hnm_aligned16(acc) && (!xs_out || hnm_aligned16(xs_out))))
    return HNM_E_ALIGN;
  if (num_rows == 0) return HNM_OK;
  const int64_t total4 = num_rows * (dim / 4);
  const int T = 256;
  const unsigned grid = (unsigned)std::min<int64_t>((total4 + T - 1) / T, (int64_t)hnm_num_sms() * 16);
  finish_kernel<<<grid, T, 0, stream>>>((const float4*)partial, (const float4*)xs_in, dis, alpha, (float4*)xs_out,
                                        (float4*)acc, row_begin, total4, dim / 4);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}


// ----------------------------------------------------------------------------- peer-memory exchange (multi-GPU)
namespace {
// Owner side of the exchange: rank `rank` owns item rows [rank * rpo, min(I, (rank + 1) * rpo)).  Its staging
// buffer holds one partial sum per rank for each of them ([world][rpo][dim], written by the peers' gather
// kernels over NVLink).  The partials are added in rank order (deterministic), the self loop, the degree
// normalisation and the layer sum are applied exactly as hnm_lightgcn_finish does, and the new row of the
// pre-scaled table goes to EVERY rank's copy (peer stores) -- one kernel instead of an all-reduce of the item
// block plus a finish pass replicated on every rank.
__global__ void finish_peer_kernel(const float4* __restrict__ stage, int world, int rpo, int rank,
                                   const float4* __restrict__ xs_in, const float* __restrict__ dis, float alpha,
                                   float4* const* __restrict__ peer_xs_out, float4* __restrict__ acc,
                                   float4* const* __restrict__ peer_acc, int64_t item_row_begin, int64_t own_rows,
                                   int dim4) {
  const int64_t total4 = own_rows * dim4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lr = i / dim4;
    const int64_t row = item_row_begin + (int64_t)rank * rpo + lr;            // node id
    const int64_t g = row * dim4 + (i - lr * dim4);
    float4 s = stage[i];
    for (int r = 1; r < world; ++r) {
      const float4 p = stage[(int64_t)r * rpo * dim4 + i];
      s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
    }
    const float di = __ldg(dis + row);
    const float4 self = xs_in[g];                          // the self loop (weight 1, lightgcn.py:127-132)
    s.x += self.x; s.y += self.y; s.z += self.z; s.w += self.w;
    const float4 e = make_float4(__fmul_rn(di, s.x), __fmul_rn(di, s.y), __fmul_rn(di, s.z), __fmul_rn(di, s.w));
    if (peer_xs_out) {
      const float4 x = make_float4(__fmul_rn(di, e.x), __fmul_rn(di, e.y), __fmul_rn(di, e.z), __fmul_rn(di, e.w));
      for (int r = 0; r < world; ++r) peer_xs_out[r][g] = x;
    }
    float4 a = acc[g];
    a.x = __fadd_rn(a.x, __fmul_rn(alpha, e.x));
    a.y = __fadd_rn(a.y, __fmul_rn(alpha, e.y));
    a.z = __fadd_rn(a.z, __fmul_rn(alpha, e.z));
    a.w = __fadd_rn(a.w, __fmul_rn(alpha, e.w));
    acc[g] = a;
    if (peer_acc) {                                        // last layer: the finished layer sum goes to every rank
      for (int r = 0; r < world; ++r)
        if (r != rank) peer_acc[r][g] = a;
    }
  }
}
}  // namespace

extern "C" int hnm_lightgcn_partial_peer(const int32_t* seg_begin, const int32_t* seg_end, const int32_t* csr_col,
                                         const float* csr_w, const float* xs_in, void* const* peer_stage,
                                         int32_t rows_per_owner, int32_t rank, int32_t dim, int64_t row_begin,
                                         int64_t row_end, const int32_t* heavy_rows, int32_t num_heavy,
                                         int32_t num_huge, int32_t heavy_threshold, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!seg_begin || !seg_end || !csr_col || !xs_in || !peer_stage) return HNM_E_NULL;
  if (num_heavy > 0 && !heavy_rows) return HNM_E_NULL;
  if (num_huge < 0 || num_huge > num_heavy) return HNM_E_RANGE;
  if (dim <= 0 || dim % 4 != 0 || row_begin < 0 || row_begin > row_end || rows_per_owner < 1 || rank < 0)
    return HNM_E_RANGE;
  if (!hnm_aligned16(xs_in)) return HNM_E_ALIGN;
  if (dim != 8 && dim != 16 && dim != 32 && dim != 64 && dim != 128 && dim != 256) return HNM_E_DIM;
  if (row_begin == row_end) return HNM_OK;
  const Seg seg{seg_begin, seg_end, row_begin, reinterpret_cast<float* const*>(peer_stage), rows_per_owner, rank};
  float* dummy = reinterpret_cast<float*>(16);            // never dereferenced: out_row() goes through seg.peer
  if (csr_w)
    return dispatch_layer<true>(dim, seg, 1, csr_col, csr_w, nullptr, xs_in, dummy, nullptr, 0.f, row_begin, row_end,
                                heavy_rows, num_heavy, num_huge, heavy_threshold, stream);
  return dispatch_layer<false>(dim, seg, 1, csr_col, csr_w, nullptr, xs_in, dummy, nullptr, 0.f, row_begin, row_end,
                               heavy_rows, num_heavy, num_huge, heavy_threshold, stream);
}

extern "C" int hnm_lightgcn_finish_peer(const float* stage, int32_t world, int32_t rows_per_owner, int32_t rank,
                                        const float* xs_in, const float* dis, float alpha, void* const* peer_xs_out,
                                        float* acc, void* const* peer_acc, int64_t item_row_begin, int64_t num_items,
                                        int32_t dim, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!stage || !xs_in || !dis || !acc) return HNM_E_NULL;
  if (world < 1 || rows_per_owner < 1 || rank < 0 || rank >= world || dim <= 0 || dim % 4 != 0 || num_items < 0 ||
      item_row_begin < 0)
    return HNM_E_RANGE;
  if (!(hnm_aligned16(stage) && hnm_aligned16(xs_in) && hnm_aligned16(acc))) return HNM_E_ALIGN;
  const int64_t own = std::max<int64_t>(0, std::min<int64_t>(rows_per_owner, num_items - (int64_t)rank * rows_per_owner));
  if (own == 0) return HNM_OK;
  const int64_t total4 = own * (dim / 4);
  const int T = 256;
  const unsigned grid = (unsigned)std::min<int64_t>((total4 + T - 1) / T, (int64_t)hnm_num_sms() * 8);
  finish_peer_kernel<<<grid, T, 0, stream>>>((const float4*)stage, world, rows_per_owner, rank, (const float4*)xs_in,
                                             dis, alpha, reinterpret_cast<float4* const*>(peer_xs_out), (float4*)acc,
                                             reinterpret_cast<float4* const*>(peer_acc), item_row_begin, own, dim / 4);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}
