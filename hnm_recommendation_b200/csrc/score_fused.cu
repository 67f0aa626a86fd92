// Fused full-catalog score + candidate select on the 5th-gen tensor cores (sm_100a).
// Replaces the `torch.matmul(user_embeds, item_embeddings.t())` + `torch.topk` pair of
// src/models/lightgcn.py:202,356 without ever writing the [users, items] score matrix.
//
//   hnm_absmax / hnm_score_pack : fp32 table -> power-of-two scaled fp16, zero padded
//   hnm_score_topk_fused        : TMA -> smem (128B swizzle) -> tcgen05.mma (fp16 x fp16 -> fp32
//                                 accumulators in TMEM) -> tcgen05.ld epilogue that keeps, per user,
//                                 every item whose approximate score beats a running threshold
//   hnm_rescore_topk            : exact fp64 scores of the survivors, canonical top-k, certificate
//
// Kernel shape (DESIGN.md "score_topk"): one persistent CTA per SM, 20 warps:
//   warp 0   TMA producer      A: 4 user tiles x [128 x 64] fp16 per "super tile" (double buffered),
//                              B: item tiles [128 x 64] fp16 through a 4-stage ring
//   warp 1   MMA issuer        for every item tile, 4 accumulators (one per user tile), each
//                              4 x tcgen05.mma M128 N128 K16; accumulator m lives in TMEM columns
//                              [128 m, 128 m + 128)
//   warp 2   TMEM allocator
//   warps 4..19  epilogue, 4 warpgroups; warpgroup m drains accumulator m, thread = one user row.
// Holding 4 user tiles per CTA makes every 16 KB item tile feed 4 MMAs: the L2 -> smem stream
// drops to ~30 GB/s per SM, which the L2 can supply to all 148 SMs at tensor-core speed.
//
// Select (per user row, all in registers): 32 bucket maxima (bucket = position of the 8-column
// group inside a pair of item tiles); tau = the `kth_sel`-th largest bucket maximum is a lower
// bound on the kth_sel-th best score seen so far, because bucket maxima belong to distinct items.
// A column group is inspected element-wise only if its maximum exceeds tau; survivors are appended
// to the user's candidate list in global memory.  tau is refreshed on a geometric schedule.
// The first two item tiles are run twice: once to seed the buckets, once to collect.
#include <algorithm>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <math.h>
#include "common.cuh"

namespace {

constexpr int kDim = HNM_FUSED_DIM;         // 64 fp16 = one 128-byte swizzle row
constexpr int kUserTile = 128;              // UMMA M
constexpr int kMU = 4;                      // user tiles per CTA (accumulators in flight)
constexpr int kSuper = kUserTile * kMU;     // 512 users per CTA pass
constexpr int kItemTile = 128;              // UMMA N
constexpr int kStagesB = 4;
constexpr int kBootTiles = 2;               // 2 x 128 items = 32 buckets x 8 columns
constexpr int kEpiWarps = 4 * kMU;
constexpr int kThreads = (4 + kEpiWarps) * 32;   // 640
constexpr int kTileBytes = kItemTile * kDim * 2; // 16384 (A tile and B tile have the same shape)
constexpr int kNumBuckets = 32;

static_assert(kUserTile == HNM_FUSED_USER_TILE && kItemTile == HNM_FUSED_ITEM_TILE && kSuper == HNM_FUSED_USER_BLOCK, "header mismatch");

struct __align__(8) Barriers {
  uint64_t a_full[2], a_empty[2];
  uint64_t b_full[kStagesB], b_empty[kStagesB];
  uint64_t t_full[kMU], t_empty[kMU];
  uint32_t tmem_base;
};
constexpr size_t kSmemBytes = 1024 /*align slack*/ + 2 * kMU * kTileBytes + kStagesB * kTileBytes + sizeof(Barriers);

// ----------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row atoms of 1024 B (SBO), LBO unused.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;                 // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                 // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// kind::f16, A = B = fp16 (K-major), D = fp32, M = 128, N = 128
constexpr uint32_t kInstrDesc = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kItemTile >> 3) << 17) |
                                ((uint32_t)(kUserTile >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(kInstrDesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// The registers written by tcgen05.ld are only valid after tcgen05.wait::ld; passing them through the
// wait as in/out operands keeps the compiler from scheduling their uses above it.
__device__ __forceinline__ void tmem_ld_wait(float (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),
                 "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
               :
               : "memory");
}

// ----------------------------------------------------------------------------- select state
struct RowState {
  float tau;                  // collect threshold (+inf while seeding)
  int cnt;                    // candidates appended so far (may exceed the capacity: overflow)
  float bm[kNumBuckets];      // bucket maxima
};

__device__ __forceinline__ float kth_largest(const float (&bm)[kNumBuckets], int kth) {
  float prev = INFINITY, cur = -INFINITY;
  for (int r = 0; r < kth; ++r) {
    cur = -INFINITY;
#pragma unroll
    for (int i = 0; i < kNumBuckets; ++i) cur = (bm[i] < prev) ? fmaxf(cur, bm[i]) : cur;
    prev = cur;
  }
  return cur;
}

// 16 accumulator columns of one user row: 2 groups of 8.  PAR selects the bucket half.
template <int PAR>
__device__ __forceinline__ void select16(const float (&v)[16], int chunk, int col0, RowState& st,
                                         uint2* __restrict__ cand, int cap) {
#pragma unroll
  for (int g = 0; g < 2; ++g) {
    const float* x = v + 8 * g;
    const float m8 = fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7])));
    float& b = st.bm[PAR * 16 + chunk * 2 + g];
    b = fmaxf(b, m8);
    if (m8 > st.tau) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (x[j] > st.tau) {
          if (st.cnt < cap) cand[st.cnt] = make_uint2(__float_as_uint(x[j]), (uint32_t)(col0 + 8 * g + j));
          ++st.cnt;
        }
      }
    }
  }
}

template <int PAR>
__device__ __forceinline__ void drain_tile(uint32_t taddr, int item0, int valid_cols, RowState& st,
                                           uint2* __restrict__ cand, int cap, uint64_t* t_empty, int lane) {
  float va[16], vb[16];
  tmem_ld16(taddr, va);
  tmem_ld_wait(va);
#pragma unroll
  for (int c = 0; c < 8; c += 2) {
    tmem_ld16(taddr + (c + 1) * 16, vb);
    if (valid_cols < kItemTile) {
#pragma unroll
      for (int j = 0; j < 16; ++j) va[j] = (c * 16 + j < valid_cols) ? va[j] : -INFINITY;
    }
    select16<PAR>(va, c, item0 + c * 16, st, cand, cap);
    tmem_ld_wait(vb);
    if (c + 2 < 8) {
      tmem_ld16(taddr + (c + 2) * 16, va);
    } else {
      // every column of this accumulator is in registers: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty);
    }
    if (valid_cols < kItemTile) {
#pragma unroll
      for (int j = 0; j < 16; ++j) vb[j] = ((c + 1) * 16 + j < valid_cols) ? vb[j] : -INFINITY;
    }
    select16<PAR>(vb, c + 1, item0 + (c + 1) * 16, st, cand, cap);
    if (c + 2 < 8) tmem_ld_wait(va);
  }
}

// ----------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kThreads, 1)
score_topk_fused_kernel(const __grid_constant__ CUtensorMap map_users, const __grid_constant__ CUtensorMap map_items,
                        int num_users, int num_super, int num_items, int num_item_tiles, int kth_sel,
                        uint2* __restrict__ cand, int cap, int32_t* __restrict__ cand_count,
                        float* __restrict__ cand_thresh) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                   // [2][kMU][kTileBytes]
  uint8_t* smem_b = smem + 2 * kMU * kTileBytes;            // [kStagesB][kTileBytes]
  Barriers* bars = reinterpret_cast<Barriers*>(smem_b + kStagesB * kTileBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int boot = num_item_tiles < kBootTiles ? num_item_tiles : kBootTiles;
  const int num_iters = num_item_tiles + boot;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_users);
    tma_prefetch_desc(&map_items);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->a_full[i], 1); mbar_init(&bars->a_empty[i], 1); }
    for (int i = 0; i < kStagesB; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_empty[i], 1); }
    for (int i = 0; i < kMU; ++i) { mbar_init(&bars->t_full[i], 1); mbar_init(&bars->t_empty[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t g = 0;
      int n = 0;
      for (int st = blockIdx.x; st < num_super; st += gridDim.x, ++n) {
        const int abuf = n & 1;
        mbar_wait(&bars->a_empty[abuf], ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(&bars->a_full[abuf], kMU * kTileBytes);
        for (int m = 0; m < kMU; ++m)
          tma_load_2d(smem_a + (abuf * kMU + m) * kTileBytes, &map_users, &bars->a_full[abuf], 0,
                      st * kSuper + m * kUserTile);
        for (int it = 0; it < num_iters; ++it, ++g) {
          const int tile = it < boot ? it : it - boot;
          const int stage = g % kStagesB;
          mbar_wait(&bars->b_empty[stage], ((g / kStagesB) & 1) ^ 1);
          mbar_expect_tx(&bars->b_full[stage], kTileBytes);
          tma_load_2d(smem_b + stage * kTileBytes, &map_items, &bars->b_full[stage], 0, tile * kItemTile);
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    uint32_t g = 0;
    int n = 0;
    for (int st = blockIdx.x; st < num_super; st += gridDim.x, ++n) {
      const int abuf = n & 1;
      mbar_wait(&bars->a_full[abuf], (n >> 1) & 1);
      for (int it = 0; it < num_iters; ++it, ++g) {
        const int stage = g % kStagesB;
        mbar_wait(&bars->b_full[stage], (g / kStagesB) & 1);
        const uint32_t b_addr = smem_u32(smem_b + stage * kTileBytes);
#pragma unroll 1
        for (int m = 0; m < kMU; ++m) {
          mbar_wait(&bars->t_empty[m], (g & 1) ^ 1);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_addr = smem_u32(smem_a + (abuf * kMU + m) * kTileBytes);
#pragma unroll
            for (int k = 0; k < kDim / 16; ++k) {
              umma_f16(tmem_base + m * kItemTile, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32),
                       k > 0 ? 1u : 0u);
            }
            umma_commit(&bars->t_full[m]);
          }
          __syncwarp();
        }
        if (lane == 0) umma_commit(&bars->b_empty[stage]);
        __syncwarp();
      }
      if (lane == 0) umma_commit(&bars->a_empty[abuf]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ===================================================== epilogue: warpgroup m drains accumulator m
    const int m = (warp - 4) >> 2;
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + m * kItemTile;
    uint32_t g = 0;
    for (int st = blockIdx.x; st < num_super; st += gridDim.x) {
      const int row = st * kSuper + m * kUserTile + q * 32 + lane;
      uint2* my_cand = cand + (size_t)(row < num_users ? row : 0) * cap;
      const int my_cap = row < num_users ? cap : 0;      // padded rows count but never store
      RowState rs;
      rs.tau = INFINITY;
      rs.cnt = 0;
#pragma unroll
      for (int i = 0; i < kNumBuckets; ++i) rs.bm[i] = -INFINITY;
      int next_refresh = boot + 2;
      for (int it = 0; it < num_iters; ++it, ++g) {
        const int tile = it < boot ? it : it - boot;
        if (it == boot) rs.tau = kth_largest(rs.bm, kth_sel);          // buckets seeded: start collecting
        if (it == next_refresh) {
          rs.tau = kth_largest(rs.bm, kth_sel);
          next_refresh = boot + 2 * (it - boot);
        }
        const int valid = min(kItemTile, num_items - tile * kItemTile);
        mbar_wait(&bars->t_full[m], g & 1);
        tc_fence_after();
        if (tile & 1) drain_tile<1>(taddr, tile * kItemTile, valid, rs, my_cand, my_cap, &bars->t_empty[m], lane);
        else drain_tile<0>(taddr, tile * kItemTile, valid, rs, my_cand, my_cap, &bars->t_empty[m], lane);
      }
      if (row < num_users) {
        cand_count[row] = rs.cnt;
        cand_thresh[row] = kth_largest(rs.bm, kth_sel);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ----------------------------------------------------------------------------- pack / absmax
__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(x[i]));
#pragma unroll
  for (int off = 16; off; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // m >= 0
}

// one 8-lane group per row of 64: lane handles 8 consecutive floats -> one 16-byte store
__global__ void pack_kernel(const float* __restrict__ emb, const int64_t* __restrict__ row_ids, int64_t num_rows,
                            int64_t rows_padded, float scale, __half* __restrict__ out, float* __restrict__ sumsq) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t r = t >> 3;
  const int sub = (int)(t & 7);
  if (r >= rows_padded) return;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (r < num_rows) {
    const int64_t src = row_ids ? row_ids[r] : r;
    const float* p = emb + (size_t)src * kDim + sub * 8;
    a = ldg_f4(p);
    b = ldg_f4(p + 4);
  }
  __half2 h[4];
  h[0] = __floats2half2_rn(a.x * scale, a.y * scale);
  h[1] = __floats2half2_rn(a.z * scale, a.w * scale);
  h[2] = __floats2half2_rn(b.x * scale, b.y * scale);
  h[3] = __floats2half2_rn(b.z * scale, b.w * scale);
  *reinterpret_cast<uint4*>(out + (size_t)r * kDim + sub * 8) = *reinterpret_cast<uint4*>(h);
  if (sumsq) {
    float s = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (sub == 0 && r < num_rows) sumsq[r] = s;
  }
}

// ----------------------------------------------------------------------------- rescoring
constexpr int kMaxPerLane = 4;   // candidate capacity handled = 32 * kMaxPerLane

__device__ __forceinline__ bool in_sorted(const int64_t* __restrict__ a, int64_t lo, int64_t hi, int64_t x) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = a[mid];
    if (v == x) return true;
    if (v < x) lo = mid + 1; else hi = mid;
  }
  return false;
}

__global__ void __launch_bounds__(256)
rescore_kernel(const float* __restrict__ user_emb, const float* __restrict__ item_emb,
               const int64_t* __restrict__ user_ids, int64_t batch, int dim, int64_t item_begin,
               const uint2* __restrict__ cand, int cap, const int32_t* __restrict__ cand_count,
               const float* __restrict__ cand_thresh, double inv_scale, double max_item_norm,
               const int64_t* __restrict__ excl_ptr, const int64_t* __restrict__ excl_items, int k,
               int64_t* __restrict__ out_ids, double* __restrict__ out_scores, int32_t* __restrict__ certified) {
  const int lane = threadIdx.x & 31;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= batch) return;
  const int64_t uid = user_ids ? user_ids[b] : b;
  const float* urow = user_emb + (size_t)uid * dim;
  const int raw = cand_count[b];
  const int n = min(raw, cap);
  const float thr = cand_thresh[b];
  const uint2* mine = cand + (size_t)b * cap;
  int64_t ex_lo = 0, ex_hi = 0;
  if (excl_ptr) { ex_lo = excl_ptr[b]; ex_hi = excl_ptr[b + 1]; }

  double s[kMaxPerLane];
  int64_t id[kMaxPerLane];
  int kept = 0;
#pragma unroll
  for (int e = 0; e < kMaxPerLane; ++e) {
    s[e] = -INFINITY;
    id[e] = INT64_MAX;
    const int idx = lane + 32 * e;
    if (idx < n) {
      const uint2 c = mine[idx];
      // entries at or below the final threshold are covered by the certificate bound
      if (__uint_as_float(c.x) > thr) {
        const int64_t gid = item_begin + (int64_t)c.y;
        if (!(ex_lo < ex_hi && in_sorted(excl_items, ex_lo, ex_hi, gid))) {
          const float* irow = item_emb + (size_t)c.y * dim;
          double acc = 0.0;
          for (int kk = 0; kk < dim; kk += 4) {
            const float4 u = ldg_f4(urow + kk), v = ldg_f4(irow + kk);
            acc = fma((double)u.x, (double)v.x, acc);
            acc = fma((double)u.y, (double)v.y, acc);
            acc = fma((double)u.z, (double)v.z, acc);
            acc = fma((double)u.w, (double)v.w, acc);
          }
          s[e] = acc;
          id[e] = gid;
          ++kept;
        }
      }
    }
  }
  // ||u||_2 for the error bound
  double un = 0.0;
  for (int kk = lane; kk < dim; kk += 32) { const double u = (double)urow[kk]; un = fma(u, u, un); }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    un += __shfl_xor_sync(0xffffffffu, un, off);
    kept += __shfl_xor_sync(0xffffffffu, kept, off);
  }
  un = sqrt(un);

  double kth_score = -INFINITY;
  for (int t = 0; t < k; ++t) {
    // lane-local best, then warp argmax by (score desc, id asc)
    double bs = s[0];
    int64_t bi = id[0];
    int be = 0;
#pragma unroll
    for (int e = 1; e < kMaxPerLane; ++e)
      if (hnm_before(s[e], id[e], bs, bi)) { bs = s[e]; bi = id[e]; be = e; }
    double ws = bs;
    int64_t wi = bi;
#pragma unroll
    for (int off = 16; off; off >>= 1) {
      const double os = __shfl_xor_sync(0xffffffffu, ws, off);
      const int64_t oi = __shfl_xor_sync(0xffffffffu, wi, off);
      if (hnm_before(os, oi, ws, wi)) { ws = os; wi = oi; }
    }
    if (wi == bi && ws == bs && bi != INT64_MAX) {   // ids are unique, so exactly one lane matches
#pragma unroll
      for (int e = 0; e < kMaxPerLane; ++e)
        if (e == be) { s[e] = -INFINITY; id[e] = INT64_MAX; }
    }
    if (lane == 0) {
      out_ids[(size_t)b * k + t] = wi;
      out_scores[(size_t)b * k + t] = ws;
    }
    kth_score = ws;
  }
  if (lane == 0) {
    // |approx - exact| <= eps for every pair of this user (DESIGN.md "certificate")
    const double eps = 1.1 * 0.0009765625 * un * max_item_norm + (double)dim * 0.00390625 * inv_scale;
    const bool ok = raw <= cap && kept >= k && kth_score > (double)thr * inv_scale + eps;
    certified[b] = ok ? 1 : 0;
  }
}

int make_map(CUtensorMap* map, const void* base, int64_t rows) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return HNM_E_DRIVER;
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)kDim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)(kDim * 2)};
  cuuint32_t box[2] = {(cuuint32_t)kDim, (cuuint32_t)kItemTile};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HNM_OK : HNM_E_DRIVER;
}

}  // namespace

extern "C" int hnm_absmax(const float* emb, int64_t count, float* out_absmax, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!emb || !out_absmax) return HNM_E_NULL;
  if (count <= 0) return HNM_E_RANGE;
  const int T = 256;
  const unsigned grid = (unsigned)std::min<int64_t>((count + T - 1) / T, (int64_t)hnm_num_sms() * 8);
  absmax_kernel<<<grid, T, 0, stream>>>(emb, count, out_absmax);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_score_pack(const float* emb, const int64_t* row_ids, int64_t num_rows, int64_t rows_padded,
                              int32_t dim, float scale, void* out_f16, float* out_sumsq, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!emb || !out_f16) return HNM_E_NULL;
  if (dim != kDim) return HNM_E_DIM;
  if (num_rows < 0 || rows_padded < num_rows || rows_padded <= 0) return HNM_E_RANGE;
  if (!hnm_aligned16(emb) || !hnm_aligned16(out_f16)) return HNM_E_ALIGN;
  const int T = 256;
  const int64_t threads = rows_padded * 8;
  pack_kernel<<<(unsigned)((threads + T - 1) / T), T, 0, stream>>>(emb, row_ids, num_rows, rows_padded, scale,
                                                                  (__half*)out_f16, out_sumsq);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_score_topk_fused(const void* users_f16, int64_t num_users, int64_t users_padded,
                                    const void* items_f16, int64_t num_items, int64_t items_padded, int32_t kth_sel,
                                    void* cand, int32_t cand_cap, int32_t* cand_count, float* cand_thresh,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!users_f16 || !items_f16 || !cand || !cand_count || !cand_thresh) return HNM_E_NULL;
  if (num_users <= 0 || num_items <= 0 || users_padded < num_users || items_padded < num_items) return HNM_E_RANGE;
  if (users_padded % kSuper != 0 || items_padded % kItemTile != 0) return HNM_E_RANGE;
  if (users_padded > INT32_MAX || items_padded > INT32_MAX) return HNM_E_RANGE;
  if (kth_sel < 1 || kth_sel > kNumBuckets || cand_cap < 1 || cand_cap > 32 * kMaxPerLane) return HNM_E_RANGE;
  if ((reinterpret_cast<uintptr_t>(users_f16) & 127) || (reinterpret_cast<uintptr_t>(items_f16) & 127)) return HNM_E_ALIGN;
  int rc = hnm_check_device();
  if (rc != HNM_OK) return rc;
  CUtensorMap map_u, map_i;
  if ((rc = make_map(&map_u, users_f16, users_padded)) != HNM_OK) return rc;
  if ((rc = make_map(&map_i, items_f16, items_padded)) != HNM_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    HNM_CUDA_TRY(cudaFuncSetAttribute(score_topk_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kSmemBytes));
    attr_set = true;
  }
  const int num_super = (int)(users_padded / kSuper);
  const int num_tiles = (int)(items_padded / kItemTile);
  const int grid = std::min(num_super, hnm_num_sms());
  score_topk_fused_kernel<<<grid, kThreads, kSmemBytes, stream>>>(map_u, map_i, (int)num_users, num_super,
                                                                  (int)num_items, num_tiles, kth_sel, (uint2*)cand,
                                                                  cand_cap, cand_count, cand_thresh);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_rescore_topk(const float* user_emb, const float* item_emb, const int64_t* user_ids, int64_t batch,
                                int32_t dim, int64_t item_begin, const void* cand, int32_t cand_cap,
                                const int32_t* cand_count, const float* cand_thresh, double inv_scale_product,
                                double max_item_norm, const int64_t* excl_ptr, const int64_t* excl_items, int32_t k,
                                int64_t* out_ids, double* out_scores, int32_t* out_certified, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (batch == 0) return HNM_OK;
  if (!user_emb || !item_emb || !cand || !cand_count || !cand_thresh || !out_ids || !out_scores || !out_certified)
    return HNM_E_NULL;
  if ((excl_ptr != nullptr) != (excl_items != nullptr)) return HNM_E_NULL;
  if (batch < 0 || dim <= 0 || dim % 4 != 0 || k < 1 || k > 32 || cand_cap < 1 || cand_cap > 32 * kMaxPerLane)
    return HNM_E_RANGE;
  const int wpc = 8;
  rescore_kernel<<<(unsigned)((batch + wpc - 1) / wpc), wpc * 32, 0, stream>>>(
      user_emb, item_emb, user_ids, batch, dim, item_begin, (const uint2*)cand, cand_cap, cand_count, cand_thresh,
      inv_scale_product, max_item_norm, excl_ptr, excl_items, k, out_ids, out_scores, out_certified);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}
