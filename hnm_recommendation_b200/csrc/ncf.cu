// NeuralCF scoring: fused embedding gather + GMF + MLP + prediction layer.
// Replaces src/models/neural_cf.py:125-139 (forward) and :167-206 (predict_all_items), eval mode.
//
// Layer 1 of the MLP is separable over the concatenated input,
//     W1 [mu ; mi] + b1 = (W1[:, :h] mu) + (W1[:, h:] mi + b1) = P[u] + Q[i],
// so hnm_ncf_precompute builds P (users) and Q (items, bias folded in) once per weight
// update and a pair costs one 2*h1-float gather plus the small tail of the MLP.
// The tail weights sit in shared memory and are read as warp-wide broadcasts; see the two kernels
// for the thread mapping.
#include <algorithm>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int kMaxLayers = 8;
constexpr int kMaxWidth = 256;

struct NcfDims {
  int num_layers;              // Linear layers in the MLP, first one included
  int width[kMaxLayers];       // output width of each layer
  int tail_floats;             // floats in the packed tail (W2, b2, W3, b3, ...)
};

// out[r, o] = sum_c w1[o, col_offset + c] * emb[r, c] (+ bias[o]); one warp per row.
__global__ void __launch_bounds__(256)
ncf_precompute_kernel(const float* __restrict__ emb, int64_t rows, int h, const float* __restrict__ w1, int h1,
                      int ld, int col_offset, const float* __restrict__ bias, float* __restrict__ out) {
  extern __shared__ float wt[];   // [h][h1 + 1] transposed slice of W1
  for (int t = threadIdx.x; t < h * h1; t += blockDim.x) {
    const int o = t / h, c = t % h;
    wt[c * (h1 + 1) + o] = w1[(size_t)o * ld + col_offset + c];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int wpc = blockDim.x >> 5;
  for (int64_t r = (int64_t)blockIdx.x * wpc + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpc) {
    const float* x = emb + (size_t)r * h;
    for (int o = lane; o < h1; o += 32) {
      float s = 0.f;
      for (int c = 0; c < h; ++c) s = fmaf(wt[c * (h1 + 1) + o], __ldg(x + c), s);
      if (bias) s += bias[o];
      out[(size_t)r * h1 + o] = s;
    }
  }
}

// Fast path: mf_dim 64, MLP [128 -> 64 -> 32] (the reference default, configs/model/neural_cf.yaml).
// A warp takes 32 consecutive pairs.  Gather phase: half a warp per pair, 16 lanes x 128-bit, so every
// 256-byte table row is one coalesced request (a thread-per-pair gather issues 32-sector requests and
// was LSU bound); h1 = relu(P[u] + Q[i]) goes to a padded shared-memory tile and the GMF dot product is
// reduced across the 16 lanes.  Compute phase: lane = pair, h1 from shared memory (conflict free),
// W2 / b2 / wp broadcast from shared memory.
constexpr int kNcfWarps = 8;
constexpr int kNcfStride = 65;       // floats per h1 row in shared memory (64 + 1: bank = (pair + c) % 32)

__global__ void __launch_bounds__(kNcfWarps * 32)
ncf_score_default_kernel(const float* __restrict__ gu, const float* __restrict__ gi, const float* __restrict__ pu,
                         const float* __restrict__ qi, const float* __restrict__ tail, const float* __restrict__ wp,
                         float bp, const int64_t* __restrict__ user_ids, const int64_t* __restrict__ item_ids,
                         const int32_t* __restrict__ cand_items, int cand_per_user, int64_t total,
                         float* __restrict__ out) {
  constexpr int MF = 64, H1 = 64, H2 = 32;
  extern __shared__ __align__(16) float smem_ncf[];
  float* s_w2 = smem_ncf;                       // [H2][H1]
  float* s_b2 = s_w2 + H2 * H1;                 // [H2]
  float* s_wp = s_b2 + H2;                      // [MF + H2]
  float* s_h1 = s_wp + MF + H2 + (threadIdx.x >> 5) * (32 * kNcfStride + 32);   // per warp: [32][65] + gmf[32]
  float* s_g = s_h1 + 32 * kNcfStride;
  for (int t = threadIdx.x; t < H2 * H1; t += blockDim.x) s_w2[t] = tail[t];
  for (int t = threadIdx.x; t < H2; t += blockDim.x) s_b2[t] = tail[H2 * H1 + t];
  for (int t = threadIdx.x; t < MF + H2; t += blockDim.x) s_wp[t] = wp[t];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int half = lane >> 4, sub = lane & 15;
  const int64_t warp_global = (int64_t)blockIdx.x * kNcfWarps + (threadIdx.x >> 5);
  const int64_t warp_count = (int64_t)gridDim.x * kNcfWarps;
  const float4 wg = *reinterpret_cast<const float4*>(s_wp + sub * 4);
  for (int64_t t0 = warp_global * 32; t0 < total; t0 += warp_count * 32) {
    const int64_t t = t0 + lane;
    int64_t u = 0, i = 0;
    if (t < total) {
      if (cand_per_user > 0) {
        const int64_t r = t / cand_per_user;
        u = user_ids ? user_ids[r] : r;
        i = cand_items ? (int64_t)cand_items[t] : t - r * cand_per_user;
      } else {
        u = user_ids[t];
        i = item_ids[t];
      }
    }
    // ---- gather: pair l = 2 * step + half, this lane's 4 columns are [4 sub, 4 sub + 4)
#pragma unroll 4
    for (int step = 0; step < 16; ++step) {
      const int l = 2 * step + half;
      const int64_t ul = __shfl_sync(0xffffffffu, u, l);
      const int64_t il = __shfl_sync(0xffffffffu, i, l);
      const float4 a = ldg_f4(pu + (size_t)ul * H1 + sub * 4), b = ldg_f4(qi + (size_t)il * H1 + sub * 4);
      const float4 x = ldg_f4(gu + (size_t)ul * MF + sub * 4), z = ldg_f4(gi + (size_t)il * MF + sub * 4);
      float* h = s_h1 + l * kNcfStride + sub * 4;
      h[0] = fmaxf(a.x + b.x, 0.f);
      h[1] = fmaxf(a.y + b.y, 0.f);
      h[2] = fmaxf(a.z + b.z, 0.f);
      h[3] = fmaxf(a.w + b.w, 0.f);
      // GMF term: (gu * gi) . wp[:MF]   (neural_cf.py:127,136-139)
      float g = wg.x * __fmul_rn(x.x, z.x);
      g = fmaf(wg.y, __fmul_rn(x.y, z.y), g);
      g = fmaf(wg.z, __fmul_rn(x.z, z.z), g);
      g = fmaf(wg.w, __fmul_rn(x.w, z.w), g);
#pragma unroll
      for (int off = 8; off; off >>= 1) g += __shfl_xor_sync(0xffffffffu, g, off);
      if (sub == 0) s_g[l] = g;
    }
    __syncwarp();
    // ---- compute: lane = pair
    float h1[H1];
#pragma unroll
    for (int c = 0; c < H1; ++c) h1[c] = s_h1[lane * kNcfStride + c];
    float y = s_g[lane];
#pragma unroll 4
    for (int j = 0; j < H2; ++j) {
      float s = 0.f;
      const float4* w = reinterpret_cast<const float4*>(s_w2 + j * H1);
#pragma unroll
      for (int c = 0; c < H1 / 4; ++c) {
        const float4 ww = w[c];
        s = fmaf(ww.x, h1[4 * c], s);
        s = fmaf(ww.y, h1[4 * c + 1], s);
        s = fmaf(ww.z, h1[4 * c + 2], s);
        s = fmaf(ww.w, h1[4 * c + 3], s);
      }
      s = fmaxf(s + s_b2[j], 0.f);
      y = fmaf(s_wp[MF + j], s, y);
    }
    if (t < total) out[t] = y + bp;
    __syncwarp();
  }
}

// ----------------------------------------------------------------------------- tensor-core fast path
// Same default architecture, layer 2 ([pairs, 64] x [64, 32]) on the tensor cores with a 3xTF32 split
//     h1 W2^T  ~=  hi(h1) hi(W2)^T + hi(h1) lo(W2)^T + lo(h1) hi(W2)^T        (error ~2^-21 per product,
// inside the 1e-5 tolerance of P4; a plain TF32 product would be 2^-11), warp-level mma.sync m16n8k8.
// A warp takes 32 pairs = two 16-row MMA tiles.  The MMA's A fragment gives the 4 lanes of a quad the same
// two rows and lane c of the quad 2 of every 8 K-columns; since the contraction does not care about the
// order of K, lane c is given the 16-byte chunks [16 v + 4 c, 16 v + 4 c + 4), v = 0..3, of its rows (so a
// quad reads 64 contiguous bytes per request: full sectors) and K step j = 2 v + p pairs slot c with column
// 16 v + 4 c + 2 p and slot c + 4 with the next one.  W2 is laid out to match, once per CTA, in shared memory
// as {hi(b0), hi(b1), lo(b0), lo(b1)} per (K step, N tile, lane): one conflict-free LDS.128 per 3 MMAs.
// h1 = relu(P[u] + Q[i]) never leaves registers (round 1 staged it through shared memory and ran layer 2 as
// 2 048 FFMA per pair: 285 ms at configs[3], 3.4x the FMA floor).
// CANDS = true (candidate / all-items form): a warp owns a contiguous range of 32-candidate tiles, so the
// user's slices of P[u] and gw[u] = gu[u] * wp[:64] stay in registers for the ~31 tiles of a user;
// CANDS = false (pair form): every row brings its own user.
constexpr int kTcWarps = 4;           // 3 CTAs of 4 warps per SM at <= 168 registers (one CTA of 8 warps at 182: 12 % of the warp slots)

__device__ __forceinline__ uint32_t tf32_hi(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool CANDS>
__global__ void __launch_bounds__(kTcWarps * 32, 3)
ncf_score_tc_kernel(const float* __restrict__ gu, const float* __restrict__ gi, const float* __restrict__ pu,
                    const float* __restrict__ qi, const float* __restrict__ tail, const float* __restrict__ wp,
                    float bp, const int64_t* __restrict__ user_ids, const int64_t* __restrict__ item_ids,
                    const int32_t* __restrict__ cand_items, int cand_per_user, int64_t num_rows, int64_t total,
                    float* __restrict__ out) {
  constexpr int MF = 64, H1 = 64, H2 = 32;
  __shared__ __align__(16) float4 s_b[8][4][32];       // [K step][N tile][lane] = {hi b0, hi b1, lo b0, lo b1}
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  for (int t = threadIdx.x; t < 8 * 4 * 32; t += blockDim.x) {
    const int l = t & 31, nt = (t >> 5) & 3, j = t >> 7;
    const int v = j >> 1, p = j & 1, lc = l & 3, lg = l >> 2;
    const int col = 16 * v + 4 * lc + 2 * p, n = 8 * nt + lg;
    const float b0 = tail[n * H1 + col], b1 = tail[n * H1 + col + 1];
    const uint32_t h0 = tf32_hi(b0), h1 = tf32_hi(b1);
    s_b[j][nt][l] = make_float4(__uint_as_float(h0), __uint_as_float(h1), b0 - __uint_as_float(h0),
                                b1 - __uint_as_float(h1));
  }
  __syncthreads();
  // layer-2 bias and prediction weights of the 32 outputs, read in the tail as {b2, wp} pairs
  __shared__ float2 s_tail2[32];
  if (threadIdx.x < H2) s_tail2[threadIdx.x] = make_float2(tail[H2 * H1 + threadIdx.x], wp[MF + threadIdx.x]);
  __syncthreads();
  const int64_t warp_global = (int64_t)blockIdx.x * kTcWarps + (threadIdx.x >> 5);
  const int64_t warp_count = (int64_t)gridDim.x * kTcWarps;
  const int tiles_per_row = CANDS ? (cand_per_user + 31) / 32 : 1;
  const int64_t num_tiles = CANDS ? num_rows * tiles_per_row : (total + 31) / 32;
  // contiguous tile ranges: consecutive tiles of a warp belong to the same user
  const int64_t t_begin = num_tiles * warp_global / warp_count, t_end = num_tiles * (warp_global + 1) / warp_count;
  int64_t cur_row = -1;
  float4 pr[4], gw[4];                                  // CANDS: this lane's slices of P[u] and gu[u] * wp
  for (int64_t tile = t_begin; tile < t_end; ++tile) {
    // ---- ids of the warp's 32 pairs (lane = pair), then of this lane's four rows g, g + 8, g + 16, g + 24
    int64_t my_u = 0, my_i = 0, my_out = -1;
    if (CANDS) {
      const int64_t r = tile / tiles_per_row;
      const int cpos = (int)(tile - r * tiles_per_row) * 32 + lane;
      if (r != cur_row) {
        cur_row = r;
        const int64_t u = user_ids ? user_ids[r] : r;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          pr[v] = ldg_f4(pu + (size_t)u * H1 + 16 * v + 4 * c);
          const float4 x = ldg_f4(gu + (size_t)u * MF + 16 * v + 4 * c);
          const float4 wgv = ldg_f4(wp + 16 * v + 4 * c);            // wp[:64] on this lane's GMF columns
          gw[v] = make_float4(wgv.x * x.x, wgv.y * x.y, wgv.z * x.z, wgv.w * x.w);
        }
      }
      if (cpos < cand_per_user) {
        my_out = r * cand_per_user + cpos;
        my_i = cand_items ? (int64_t)cand_items[my_out] : cpos;
      }
    } else {
      const int64_t t = tile * 32 + lane;
      if (t < total) { my_out = t; my_u = user_ids[t]; my_i = item_ids[t]; }
    }
    int64_t it[4], ut[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      it[r] = __shfl_sync(0xffffffffu, my_i, g + 8 * r);
      if (!CANDS) ut[r] = __shfl_sync(0xffffffffu, my_u, g + 8 * r);
    }
    // ---- gather: 16 columns of Q[i] (and P[u]) per row -> h1 = relu(P + Q); 16 columns of the GMF product
    float h1[4][16];
    float gsum[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      float4 qv[4], zv[4];
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        qv[v] = ldg_f4(qi + (size_t)it[r] * H1 + 16 * v + 4 * c);
        zv[v] = ldg_f4(gi + (size_t)it[r] * MF + 16 * v + 4 * c);
      }
      float gs = 0.f;
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float4 pv = CANDS ? pr[v] : ldg_f4(pu + (size_t)ut[r] * H1 + 16 * v + 4 * c);
        float4 wv;
        if (CANDS) {
          wv = gw[v];
        } else {
          const float4 x = ldg_f4(gu + (size_t)ut[r] * MF + 16 * v + 4 * c);
          const float4 wgv = ldg_f4(wp + 16 * v + 4 * c);
          wv = make_float4(wgv.x * x.x, wgv.y * x.y, wgv.z * x.z, wgv.w * x.w);
        }
        h1[r][4 * v + 0] = fmaxf(pv.x + qv[v].x, 0.f);
        h1[r][4 * v + 1] = fmaxf(pv.y + qv[v].y, 0.f);
        h1[r][4 * v + 2] = fmaxf(pv.z + qv[v].z, 0.f);
        h1[r][4 * v + 3] = fmaxf(pv.w + qv[v].w, 0.f);
        gs = fmaf(wv.x, zv[v].x, gs);
        gs = fmaf(wv.y, zv[v].y, gs);
        gs = fmaf(wv.z, zv[v].z, gs);
        gs = fmaf(wv.w, zv[v].w, gs);
      }
      gsum[r] = gs;
    }
    // ---- layer 2 on the tensor cores: two M tiles (rows {g, g+8} and {g+16, g+24}) x four N tiles
    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k0 = 4 * (j >> 1) + 2 * (j & 1);        // index of slot c's column inside this lane's 16
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float a0 = h1[2 * mt][k0], a1 = h1[2 * mt + 1][k0], a2 = h1[2 * mt][k0 + 1], a3 = h1[2 * mt + 1][k0 + 1];
        ah[mt][0] = tf32_hi(a0); ah[mt][1] = tf32_hi(a1); ah[mt][2] = tf32_hi(a2); ah[mt][3] = tf32_hi(a3);
        al[mt][0] = __float_as_uint(a0 - __uint_as_float(ah[mt][0]));
        al[mt][1] = __float_as_uint(a1 - __uint_as_float(ah[mt][1]));
        al[mt][2] = __float_as_uint(a2 - __uint_as_float(ah[mt][2]));
        al[mt][3] = __float_as_uint(a3 - __uint_as_float(ah[mt][3]));
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float4 b = s_b[j][nt][lane];
        const uint32_t bh0 = __float_as_uint(b.x), bh1 = __float_as_uint(b.y);
        const uint32_t bl0 = __float_as_uint(b.z), bl1 = __float_as_uint(b.w);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_tf32(acc[mt][nt], al[mt], bh0, bh1);      // small terms first
          mma_tf32(acc[mt][nt], ah[mt], bl0, bl1);
          mma_tf32(acc[mt][nt], ah[mt], bh0, bh1);
        }
      }
    }
    // ---- tail: relu(+ b2), prediction weights, sum over the 32 outputs (8 here, the rest across the quad)
    float y[4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float s_lo = 0.f, s_hi = 0.f;                      // rows g + 16 mt and g + 16 mt + 8
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float2 t0 = s_tail2[8 * nt + 2 * c], t1 = s_tail2[8 * nt + 2 * c + 1];      // {b2, wp} of two columns
        s_lo = fmaf(t0.y, fmaxf(acc[mt][nt][0] + t0.x, 0.f), s_lo);
        s_lo = fmaf(t1.y, fmaxf(acc[mt][nt][1] + t1.x, 0.f), s_lo);
        s_hi = fmaf(t0.y, fmaxf(acc[mt][nt][2] + t0.x, 0.f), s_hi);
        s_hi = fmaf(t1.y, fmaxf(acc[mt][nt][3] + t1.x, 0.f), s_hi);
      }
      y[2 * mt] = s_lo + gsum[2 * mt];
      y[2 * mt + 1] = s_hi + gsum[2 * mt + 1];
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      y[r] += __shfl_xor_sync(0xffffffffu, y[r], 1);
      y[r] += __shfl_xor_sync(0xffffffffu, y[r], 2);
    }
    // row g + 8 r lives in every lane of quad g; hand it to lane (g + 8 r) so that the store is coalesced
    float mine = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float v = __shfl_sync(0xffffffffu, y[r], 4 * (lane & 7));
      if ((lane >> 3) == r) mine = v;
    }
    if (my_out >= 0) out[my_out] = mine + bp;
  }
}

// Any depth / width up to kMaxWidth: activations ping-pong through local memory.
__global__ void __launch_bounds__(128)
ncf_score_generic_kernel(const float* __restrict__ gu, const float* __restrict__ gi, const float* __restrict__ pu,
                         const float* __restrict__ qi, const float* __restrict__ tail, NcfDims dims,
                         const float* __restrict__ wp, float bp, int mf, const int64_t* __restrict__ user_ids,
                         const int64_t* __restrict__ item_ids, const int32_t* __restrict__ cand_items,
                         int cand_per_user, int64_t total, float* __restrict__ out) {
  extern __shared__ float s_tail[];
  for (int t = threadIdx.x; t < dims.tail_floats; t += blockDim.x) s_tail[t] = tail[t];
  __syncthreads();
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    int64_t u, i;
    if (cand_per_user > 0) {
      const int64_t r = t / cand_per_user;
      u = user_ids ? user_ids[r] : r;
      i = cand_items ? (int64_t)cand_items[t] : t - r * cand_per_user;
    } else {
      u = user_ids[t];
      i = item_ids[t];
    }
    float xa[kMaxWidth], xb[kMaxWidth];
    const int h1 = dims.width[0];
    for (int c = 0; c < h1; ++c) xa[c] = fmaxf(pu[(size_t)u * h1 + c] + qi[(size_t)i * h1 + c], 0.f);
    float* cur = xa;
    float* nxt = xb;
    int in = h1, off = 0;
    for (int l = 1; l < dims.num_layers; ++l) {
      const int o = dims.width[l];
      const float* w = s_tail + off;
      const float* bb = w + o * in;
      for (int j = 0; j < o; ++j) {
        float s = 0.f;
        for (int c = 0; c < in; ++c) s = fmaf(w[j * in + c], cur[c], s);
        nxt[j] = fmaxf(s + bb[j], 0.f);
      }
      off += o * in + o;
      in = o;
      float* tmp = cur; cur = nxt; nxt = tmp;
    }
    float y = 0.f;
    for (int c = 0; c < mf; ++c) y = fmaf(wp[c], __fmul_rn(gu[(size_t)u * mf + c], gi[(size_t)i * mf + c]), y);
    for (int j = 0; j < in; ++j) y = fmaf(wp[mf + j], cur[j], y);
    out[t] = y + bp;
  }
}

int fill_dims(NcfDims& d, const int32_t* widths_host, int num_layers) {
  if (!widths_host) return HNM_E_NULL;
  if (num_layers < 1 || num_layers > kMaxLayers) return HNM_E_DIM;
  d.num_layers = num_layers;
  d.tail_floats = 0;
  for (int l = 0; l < num_layers; ++l) {
    d.width[l] = widths_host[l];
    if (d.width[l] < 1 || d.width[l] > kMaxWidth) return HNM_E_DIM;
    if (l > 0) d.tail_floats += d.width[l] * d.width[l - 1] + d.width[l];
  }
  return HNM_OK;
}

int launch_score(const float* gu, const float* gi, const float* pu, const float* qi, const float* tail,
                 const int32_t* widths_host, int num_layers, const float* wp, float bp, int mf,
                 const int64_t* user_ids, const int64_t* item_ids, const int32_t* cand_items, int cand_per_user,
                 int64_t total, float* out, cudaStream_t stream) {
  if (total == 0) return HNM_OK;
  if (!gu || !gi || !pu || !qi || !wp || !out) return HNM_E_NULL;
  NcfDims d{};
  int rc = fill_dims(d, widths_host, num_layers);
  if (rc != HNM_OK) return rc;
  if (num_layers > 1 && !tail) return HNM_E_NULL;
  if (mf < 1 || total < 0) return HNM_E_RANGE;
  const bool is_default = num_layers == 2 && mf == 64 && d.width[0] == 64 && d.width[1] == 32 &&
                          hnm_aligned16(gu) && hnm_aligned16(gi) && hnm_aligned16(pu) && hnm_aligned16(qi);
  static const bool simt = getenv("HNM_NCF_SIMT") != nullptr;          // A/B switch: round 1's FFMA kernel
  if (is_default && !simt && hnm_aligned16(wp)) {
    const int T = kTcWarps * 32;
    const int64_t rows = cand_per_user > 0 ? total / cand_per_user : 0;
    const int64_t tiles = cand_per_user > 0 ? rows * ((cand_per_user + 31) / 32) : (total + 31) / 32;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((tiles + kTcWarps - 1) / kTcWarps,
                                                                           (int64_t)hnm_num_sms() * 3));
    if (cand_per_user > 0)
      ncf_score_tc_kernel<true><<<grid, T, 0, stream>>>(gu, gi, pu, qi, tail, wp, bp, user_ids, item_ids, cand_items,
                                                        cand_per_user, rows, total, out);
    else
      ncf_score_tc_kernel<false><<<grid, T, 0, stream>>>(gu, gi, pu, qi, tail, wp, bp, user_ids, item_ids, cand_items,
                                                         0, 0, total, out);
  } else if (is_default) {
    const int T = kNcfWarps * 32;
    const size_t smem = sizeof(float) * (64 * 32 + 32 + 64 + 32 + kNcfWarps * (32 * kNcfStride + 32));
    HNM_CUDA_TRY(hnm_allow_smem(ncf_score_default_kernel, (int)smem));
    const unsigned grid = (unsigned)std::min<int64_t>((total + T - 1) / T, (int64_t)hnm_num_sms() * 6);
    ncf_score_default_kernel<<<grid, T, smem, stream>>>(gu, gi, pu, qi, tail, wp, bp, user_ids, item_ids, cand_items,
                                                       cand_per_user, total, out);
  } else {
    const size_t smem = sizeof(float) * (size_t)d.tail_floats;
    if (smem > 200 * 1024) return HNM_E_DIM;
    HNM_CUDA_TRY(hnm_allow_smem(ncf_score_generic_kernel, 200 * 1024));
    const int T = 128;
    const unsigned grid = (unsigned)std::min<int64_t>((total + T - 1) / T, (int64_t)hnm_num_sms() * 8);
    ncf_score_generic_kernel<<<grid, T, smem, stream>>>(gu, gi, pu, qi, tail, d, wp, bp, mf, user_ids, item_ids,
                                                       cand_items, cand_per_user, total, out);
  }
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

}  // namespace

extern "C" int hnm_ncf_precompute(const float* mlp_emb, int64_t rows, int32_t h, const float* w1, int32_t h1,
                                  int32_t w1_cols, int32_t col_offset, const float* bias, float* out,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (rows == 0) return HNM_OK;
  if (!mlp_emb || !w1 || !out) return HNM_E_NULL;
  if (rows < 0 || h < 1 || h1 < 1 || col_offset < 0 || col_offset + h > w1_cols) return HNM_E_RANGE;
  const size_t smem = sizeof(float) * (size_t)h * (h1 + 1);
  if (smem > 200 * 1024) return HNM_E_DIM;
  HNM_CUDA_TRY(hnm_allow_smem(ncf_precompute_kernel, 200 * 1024));
  const unsigned grid = (unsigned)std::min<int64_t>((rows + 7) / 8, (int64_t)hnm_num_sms() * 8);
  ncf_precompute_kernel<<<grid, 256, smem, stream>>>(mlp_emb, rows, h, w1, h1, w1_cols, col_offset, bias, out);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_ncf_score_pairs(const float* gmf_user, const float* gmf_item, const float* pu, const float* qi,
                                   const float* mlp_tail, const int32_t* widths_host, int32_t num_layers,
                                   const float* wp, float bp, const int64_t* user_ids, const int64_t* item_ids,
                                   int64_t num_pairs, int32_t mf_dim, float* out, void* stream_) {
  if (num_pairs > 0 && (!user_ids || !item_ids)) return HNM_E_NULL;
  return launch_score(gmf_user, gmf_item, pu, qi, mlp_tail, widths_host, num_layers, wp, bp, mf_dim, user_ids,
                      item_ids, nullptr, 0, num_pairs, out, (cudaStream_t)stream_);
}

extern "C" int hnm_ncf_score_candidates(const float* gmf_user, const float* gmf_item, const float* pu,
                                        const float* qi, const float* mlp_tail, const int32_t* widths_host,
                                        int32_t num_layers, const float* wp, float bp, const int64_t* user_ids,
                                        int64_t num_rows, const int32_t* cand_items, int32_t cand_per_user,
                                        int32_t mf_dim, float* out, void* stream_) {
  if (cand_per_user < 1 || num_rows < 0) return HNM_E_RANGE;
  return launch_score(gmf_user, gmf_item, pu, qi, mlp_tail, widths_host, num_layers, wp, bp, mf_dim, user_ids,
                      nullptr, cand_items, cand_per_user, num_rows * cand_per_user, out, (cudaStream_t)stream_);
}
