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
// Kernel shape (DESIGN.md "score_topk"): one persistent CTA per SM, 16 warps:
//   warp 0   TMA producer      A: 3 user tiles x [128 x 64] fp16 per "super tile" (double buffered),
//                              B: item tiles [128 x 64] fp16 through a 6-stage ring
//   warps 1-3 MMA issuers      warp 1+m issues user tile m; work item w = (item tile, user tile m),
//                              accumulator slot = w mod 4: 4 x tcgen05.mma M128 N128 K16 into TMEM
//                              columns [128 slot, 128 slot + 128); warp 2 also owns the TMEM allocation
//   warps 4..15  epilogue, 3 warpgroups; warpgroup m drains every accumulator of user tile m,
//                thread = one user row (TMEM lane).
// Three user tiles share each 16 KB item tile (L2 -> smem stream ~40 GB/s per SM at full rate).  Each
// user tile owns one accumulator, its own issuer thread and its own epilogue warpgroup: while
// warpgroup m drains its accumulator the tensor pipe works for the other two user tiles.
//
// Select (per user row, all in registers): 32 bucket maxima (bucket = position of the 4-column
// group inside an item tile).  Bucket maxima belong to distinct items, so the kth_sel-th
// largest of them, tau, is a lower bound on the kth_sel-th best score seen so far.  A column group
// is inspected element-wise only if its maximum exceeds tau; survivors are appended to the user's
// candidate list in global memory.  tau is refreshed (in-place sorting network over the bucket
// registers) every time the number of item tiles seen has grown by 1/4.  The first kBootTiles item
// tiles are run twice: once to seed the buckets, once to collect.
#include <algorithm>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int kDim = HNM_FUSED_DIM;         // 64 fp16 = one 128-byte swizzle row
constexpr int kUserTile = 128;              // UMMA M
constexpr int kMU = 3;                      // user tiles per CTA
constexpr int kSlots = kMU;                 // one TMEM accumulator (128 columns) per user tile
constexpr int kSuper = kUserTile * kMU;     // 384 users per CTA pass
constexpr int kItemTile = 128;              // UMMA N
constexpr int kStagesB = 6;
constexpr int kBootTiles = 32;              // item tiles used to seed the bucket maxima (run twice)
constexpr int kEpiWarps = 4 * kMU;
constexpr int kThreads = (4 + kEpiWarps) * 32;   // 512
constexpr int kTileBytes = kItemTile * kDim * 2; // 16384 (A tile and B tile have the same shape)
constexpr int kNumBuckets = 32;

static_assert(kUserTile == HNM_FUSED_USER_TILE && kItemTile == HNM_FUSED_ITEM_TILE, "header mismatch");

struct __align__(8) Barriers {
  uint64_t a_full[2], a_empty[2];
  uint64_t b_full[kStagesB], b_empty[kStagesB];
  uint64_t t_full[kSlots], t_empty[kSlots];
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// tcgen05.wait::ld: the registers written by earlier tcgen05.ld become readable.  The loaded registers
// are threaded through the statement as in/out operands so that the compiler cannot schedule a
// consumer above it.
#define HNM_F8(v, o) "+f"(v[o]), "+f"(v[o + 1]), "+f"(v[o + 2]), "+f"(v[o + 3]), "+f"(v[o + 4]), "+f"(v[o + 5]), "+f"(v[o + 6]), "+f"(v[o + 7])
__device__ __forceinline__ void tmem_ld_wait(float (&a)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : HNM_F8(a, 0), HNM_F8(a, 8), HNM_F8(a, 16), HNM_F8(a, 24)
               :
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait(float (&a)[32], float (&b)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : HNM_F8(a, 0), HNM_F8(a, 8), HNM_F8(a, 16), HNM_F8(a, 24), HNM_F8(b, 0), HNM_F8(b, 8), HNM_F8(b, 16),
                 HNM_F8(b, 24)
               :
               : "memory");
}
#undef HNM_F8

// ----------------------------------------------------------------------------- select state
struct RowState {
  float tau;                  // collect threshold (+inf while seeding)
  int cnt;                    // candidates appended so far (may exceed the capacity: overflow)
  float bm[kNumBuckets];      // bucket maxima (an unordered multiset: see refresh_tau)
};

__device__ __forceinline__ void cmpx(float& a, float& b) {   // a <- max, b <- min
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  a = hi;
  b = lo;
}

// Bitonic sorting network (descending, 240 comparators) over the 32 bucket registers, then
// tau = bm[kth-1].  Sorting in place is legal: after any permutation register r still holds the
// maximum of some item set S_r, the S_r stay pairwise disjoint, and later updates add each new item
// to exactly one S_r.
__device__ __forceinline__ float refresh_tau(float (&bm)[kNumBuckets], int kth) {
#pragma unroll
  for (int k = 2; k <= kNumBuckets; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < kNumBuckets; ++i) {
        const int l = i ^ j;
        if (l > i) {
          if ((i & k) == 0) cmpx(bm[i], bm[l]);
          else cmpx(bm[l], bm[i]);
        }
      }
    }
  }
  float t = bm[0];
#pragma unroll
  for (int i = 1; i < kNumBuckets; ++i) t = (i < kth) ? bm[i] : t;
  return t;
}

// 32 accumulator columns of one user row = 8 groups of 4 columns.  A group is both a bucket of the
// threshold estimator and the unit that gets nominated: when its maximum beats tau the pair
// {maximum, first column} is appended and hnm_rescore_topk rescores its four items exactly.
// The common case (nothing in the chunk beats tau) is ~30 straight-line instructions and one
// branch; the rare case is 8 predicated stores.  Earlier versions branched per group and inlined a
// per-element scan: the kernel then spent most of its time stalled on instruction fetch
// (profiles/r1_fused_notes.md).
// The four group maxima pairs of a 32-column chunk; after this the 32 accumulator values are dead,
// so their registers can take the next tcgen05.ld while the rest of the chunk is processed.
__device__ __forceinline__ void group_max(const float (&v)[32], float (&q)[8]) {
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    const float* x = v + 4 * h;
    q[h] = fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3]));
  }
}

// UPDATE = false on the second visit of the seed tiles: their items already sit in the buckets and an
// item must never be counted in two buckets (tau would stop being a lower bound).
template <bool UPDATE>
__device__ __forceinline__ void finish32(const float (&q)[8], int chunk, int col0, RowState& st,
                                         uint2* __restrict__ cand, int cap) {
  if (UPDATE) {
#pragma unroll
    for (int h = 0; h < 8; ++h) st.bm[chunk * 8 + h] = fmaxf(st.bm[chunk * 8 + h], q[h]);
  }
  const float m32 = fmaxf(fmaxf(fmaxf(q[0], q[1]), fmaxf(q[2], q[3])), fmaxf(fmaxf(q[4], q[5]), fmaxf(q[6], q[7])));
  if (m32 > st.tau) {
    if (st.cnt > cap - 8) {              // no room for a full chunk: stop collecting, flag the row
      st.tau = INFINITY;
      st.cnt = cap + 1;
    } else {
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        if (q[h] > st.tau) cand[st.cnt++] = make_uint2(__float_as_uint(q[h]), (uint32_t)(col0 + 4 * h));
      }
    }
  }
}

// One 128-column accumulator of one user row.  Two 32-column loads are kept in flight: chunks 2 and 3
// are fetched while chunks 0 and 1 are processed, so one TMEM round trip per tile is exposed, not four.
template <bool UPDATE>
__device__ __forceinline__ void drain_tile(uint32_t taddr, int item0, RowState& st, uint2* __restrict__ cand,
                                           int cap, uint64_t* t_empty, int lane) {
  float va[32], vb[32], q0[8], q1[8];
  tmem_ld32(taddr, va);
  tmem_ld32(taddr + 32, vb);
  tmem_ld_wait(va, vb);
  group_max(va, q0);
  tmem_ld32(taddr + 64, va);
  group_max(vb, q1);
  tmem_ld32(taddr + 96, vb);
  finish32<UPDATE>(q0, 0, item0, st, cand, cap);
  finish32<UPDATE>(q1, 1, item0 + 32, st, cand, cap);
  tmem_ld_wait(va, vb);
  // every column of this accumulator is in registers: hand it back to its MMA issuer
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(t_empty);
  group_max(va, q0);
  group_max(vb, q1);
  finish32<UPDATE>(q0, 2, item0 + 64, st, cand, cap);
  finish32<UPDATE>(q1, 3, item0 + 96, st, cand, cap);
}

// ----------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(kThreads, 1)
score_topk_fused_kernel(const __grid_constant__ CUtensorMap map_users, const __grid_constant__ CUtensorMap map_items,
                        int num_users, int num_user_tiles, int num_item_tiles, int kth_sel,
                        uint2* __restrict__ cand, int cap, int32_t* __restrict__ cand_count,
                        float* __restrict__ cand_thresh, int mode, int boot_tiles, int refresh_div) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                   // [2][kMU][kTileBytes]
  uint8_t* smem_b = smem + 2 * kMU * kTileBytes;            // [kStagesB][kTileBytes]
  Barriers* bars = reinterpret_cast<Barriers*>(smem_b + kStagesB * kTileBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int boot = num_item_tiles < boot_tiles ? num_item_tiles : boot_tiles;
  const int num_iters = num_item_tiles + boot;
  // user tiles are dealt out evenly: CTA b owns [t_begin, t_end) and walks it in passes of up to kMU
  // tiles, so CTAs differ by at most one tile (a third of a pass), not by a whole pass
  const int t_base = num_user_tiles / (int)gridDim.x, t_rem = num_user_tiles % (int)gridDim.x;
  const int t_begin = (int)blockIdx.x * t_base + min((int)blockIdx.x, t_rem);
  const int t_end = t_begin + t_base + ((int)blockIdx.x < t_rem ? 1 : 0);
  // every CTA sweeps the catalog from a different starting tile: otherwise all 148 SMs ask the L2 for
  // the same 16 KB item tile at the same moment
  const int rot = (int)(((long long)blockIdx.x * num_item_tiles) / (int)gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_users);
    tma_prefetch_desc(&map_items);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->a_full[i], 1); mbar_init(&bars->a_empty[i], kMU); }
    for (int i = 0; i < kStagesB; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_empty[i], kMU); }
    for (int i = 0; i < kSlots; ++i) { mbar_init(&bars->t_full[i], 1); mbar_init(&bars->t_empty[i], 4); }
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

  if (warp < 4) {
  // The 4 control warps give registers back so the 12 epilogue warps can hold a whole 64-column
  // double buffer plus the 32 bucket maxima without spilling (40 * 128 + 152 * 384 <= 64 K).
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t g = 0;
      int n = 0;
      for (int t0 = t_begin; t0 < t_end; t0 += kMU, ++n) {
        const int mc = min(kMU, t_end - t0);
        const int abuf = n & 1;
        mbar_wait(&bars->a_empty[abuf], ((n >> 1) & 1) ^ 1);
        mbar_expect_tx(&bars->a_full[abuf], mc * kTileBytes);
        for (int m = 0; m < mc; ++m)
          tma_load_2d(smem_a + (abuf * kMU + m) * kTileBytes, &map_users, &bars->a_full[abuf], 0,
                      (t0 + m) * kUserTile);
        for (int it = 0; it < num_iters; ++it, ++g) {
          int tile = (it < boot ? it : it - boot) + rot;
          if (tile >= num_item_tiles) tile -= num_item_tiles;
          const int stage = g % kStagesB;
          mbar_wait(&bars->b_empty[stage], ((g / kStagesB) & 1) ^ 1);
          mbar_expect_tx(&bars->b_full[stage], kTileBytes);
          tma_load_2d(smem_b + stage * kTileBytes, &map_items, &bars->b_full[stage], 0, tile * kItemTile);
        }
      }
    }
  } else {
    // ===================================================== MMA issuers: warp 1 + m serves user tile m
    // tcgen05.mma holds its issuing thread for about the duration of the MMA (tools/bench_mma.cu:
    // 73 cycles per M128 N128 K16), so with a single issuer every mbarrier wait / fence / commit adds
    // to the tensor pipe's critical path (measured: 562 cycles per accumulator instead of 292).  Three
    // issuing threads, one per user tile, overlap each other's bookkeeping with MMA issue.
    const int m = warp - 1;
    if (lane == 0) {
      uint32_t g = 0, uses = 0;
      int n = 0;
      const uint64_t desc_hi = umma_desc_sw128(0) & ~uint64_t(0x3FFF);
      for (int t0 = t_begin; t0 < t_end; t0 += kMU, ++n) {
        const int mc = min(kMU, t_end - t0);
        const bool active = m < mc;
        const int abuf = n & 1;
        mbar_wait(&bars->a_full[abuf], (n >> 1) & 1);
        tc_fence_after();
        const uint64_t a_desc = desc_hi | (uint64_t)((smem_u32(smem_a + (abuf * kMU + m) * kTileBytes) >> 4) & 0x3FFF);
        for (int it = 0; it < num_iters; ++it, ++g) {
          const int stage = g % kStagesB;
          mbar_wait(&bars->b_full[stage], (g / kStagesB) & 1);
          if (active) {
            // accumulator m belongs to user tile m alone (its use counter is `uses`), so no issuer ever
            // has to order itself against another one
            const uint64_t b_desc = desc_hi | (uint64_t)((smem_u32(smem_b + stage * kTileBytes) >> 4) & 0x3FFF);
            mbar_wait(&bars->t_empty[m], (uses & 1) ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + m * kItemTile;
#pragma unroll
            for (int k = 0; k < kDim / 16; ++k)      // +32 bytes along K = +2 in the 16-byte address field
              umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, k > 0 ? 1u : 0u);
            umma_commit(&bars->t_full[m]);
            umma_commit(&bars->b_empty[stage]);
            ++uses;
          } else {
            mbar_arrive(&bars->b_empty[stage]);       // keep the stage's arrival count at kMU
          }
        }
        if (active) umma_commit(&bars->a_empty[abuf]);
        else mbar_arrive(&bars->a_empty[abuf]);
      }
    }
  }
  } else {
    // ===================================================== epilogue: warpgroup m drains user tile m
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int m = (warp - 4) >> 2;
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t uses = 0;
    const uint32_t taddr = lane_base + m * kItemTile;
    uint64_t* t_empty = &bars->t_empty[m];
    for (int t0 = t_begin; t0 < t_end; t0 += kMU) {
      const int mc = min(kMU, t_end - t0);
      if (m >= mc) continue;                                                // short last pass: this warpgroup rests
      const int row = (t0 + m) * kUserTile + q * 32 + lane;
      uint2* my_cand = cand + (size_t)(row < num_users ? row : 0) * cap;
      const int my_cap = row < num_users ? cap : 0;      // padded rows count but never store
      RowState rs;
      rs.tau = INFINITY;
      rs.cnt = 0;
#pragma unroll
      for (int i = 0; i < kNumBuckets; ++i) rs.bm[i] = -INFINITY;
      int next_refresh = boot;
      for (int it = 0; it < num_iters; ++it) {
        int tile = (it < boot ? it : it - boot) + rot;
        if (tile >= num_item_tiles) tile -= num_item_tiles;
        if (it == next_refresh && mode == 0) {
          // it == boot: the seed pass is over, collecting starts (again from tile 0)
          rs.tau = refresh_tau(rs.bm, kth_sel);
          const int seen = max(boot, it - boot);         // item tiles behind the current bucket maxima
          next_refresh = it + max(2, seen / refresh_div);
        }
        mbar_wait(&bars->t_full[m], uses & 1);
        ++uses;
        tc_fence_after();
        if (mode == 1 || mode == 3 || mode == 4) {   // debug: drain only / handshake only / one load
          float va[32];
          float acc = 0.f;
          const int nld = mode == 1 ? 4 : (mode == 4 ? 1 : 0);
          for (int c = 0; c < nld; ++c) { tmem_ld32(taddr + c * 32, va); tmem_ld_wait(va); acc += va[c]; }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty);
          rs.bm[0] += acc;
        } else if (it >= boot && it < 2 * boot) {      // second visit of a seed tile: collect only
          drain_tile<false>(taddr, tile * kItemTile, rs, my_cand, my_cap, t_empty, lane);
        } else {
          drain_tile<true>(taddr, tile * kItemTile, rs, my_cand, my_cap, t_empty, lane);
        }
      }
      if (mode == 0) rs.tau = refresh_tau(rs.bm, kth_sel);
      if (row < num_users) {
        cand_count[row] = rs.cnt;
        cand_thresh[row] = rs.tau;
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
__global__ void absmax_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ center, int dim,
                              float* __restrict__ out) {
  float m = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(center ? __fsub_rn(x[i], center[i % dim]) : x[i]));
#pragma unroll
  for (int off = 16; off; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // m >= 0
}

// one 8-lane group per row of 64: lane handles 8 consecutive floats -> one 16-byte store
__global__ void pack_kernel(const float* __restrict__ emb, const int64_t* __restrict__ row_ids, int64_t num_rows,
                            int64_t rows_padded, const float* __restrict__ center, float scale,
                            __half* __restrict__ out, float* __restrict__ sumsq) {
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
    if (center) {     // w = fl(x - c): ranking of u.x and u.(x - c) is the same for a fixed user
      const float4 ca = ldg_f4(center + sub * 8), cb = ldg_f4(center + sub * 8 + 4);
      a = make_float4(__fsub_rn(a.x, ca.x), __fsub_rn(a.y, ca.y), __fsub_rn(a.z, ca.z), __fsub_rn(a.w, ca.w));
      b = make_float4(__fsub_rn(b.x, cb.x), __fsub_rn(b.y, cb.y), __fsub_rn(b.z, cb.z), __fsub_rn(b.w, cb.w));
    }
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
constexpr int kMaxPerLane = 8;   // candidate capacity handled = 32 * kMaxPerLane
constexpr int kGroup = 4;        // items per nominated group (select32)
constexpr int kMaxGroups = 64;   // surviving groups hnm_rescore_topk can take per user
constexpr int kMaxContenders = 128;  // rescored items above the cut it can rank per user
constexpr int kRescoreWarps = 4;     // one user per warp
constexpr int kTileStride = 65;      // floats per staged item row (64 + 1: conflict-free column walks)

__device__ __forceinline__ bool in_sorted(const int64_t* __restrict__ a, int64_t lo, int64_t hi, int64_t x) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const int64_t v = a[mid];
    if (v == x) return true;
    if (v < x) lo = mid + 1; else hi = mid;
  }
  return false;
}

// (score desc, id asc) with 32-bit ids; used by the warp-wide sort below
__device__ __forceinline__ bool before32(double sa, int ia, double sb, int ib) {
  return sa > sb || (sa == sb && ia < ib);
}

__global__ void __launch_bounds__(kRescoreWarps * 32)
rescore_kernel(const float* __restrict__ user_emb, const float* __restrict__ item_emb,
               const int64_t* __restrict__ user_ids, int64_t batch, int dim, int64_t item_begin, int num_items_local,
               const uint2* __restrict__ cand, int cap, const int32_t* __restrict__ cand_count,
               const float* __restrict__ cand_thresh, double inv_scale, double max_item_norm,
               const float* __restrict__ center, const int64_t* __restrict__ excl_ptr,
               const int64_t* __restrict__ excl_items, int k, int64_t* __restrict__ out_ids,
               double* __restrict__ out_scores, int32_t* __restrict__ certified) {
  __shared__ uint32_t s_col[kRescoreWarps][kMaxGroups];
  __shared__ float s_tile[kRescoreWarps][32 * kTileStride];
  __shared__ double s_sc[kRescoreWarps][kMaxContenders];
  __shared__ int s_id[kRescoreWarps][kMaxContenders];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (b >= batch) return;
  const int64_t uid = user_ids ? user_ids[b] : b;
  const float* urow = user_emb + (size_t)uid * dim;
  const int raw = cand_count[b];
  const int n = min(raw, cap);
  const float thr = cand_thresh[b];
  const uint2* mine = cand + (size_t)b * cap;
  int64_t ex_lo = 0, ex_hi = 0;
  if (excl_ptr) { ex_lo = excl_ptr[b]; ex_hi = excl_ptr[b + 1]; }

  // 1. keep the groups whose maximum ended above the final threshold (the others are covered by
  //    the certificate bound) and compact them: lane j takes kept group j
  int groups = 0;
#pragma unroll
  for (int e = 0; e < kMaxPerLane; ++e) {
    const int idx = lane + 32 * e;
    uint2 c = make_uint2(0u, 0u);
    bool keep = false;
    if (idx < n) {
      c = mine[idx];
      keep = __uint_as_float(c.x) > thr;
    }
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    const int pos = groups + __popc(mask & ((1u << lane) - 1u));
    if (keep && pos < kMaxGroups) s_col[wib][pos] = c.y;
    groups += __popc(mask);
  }
  __syncwarp();
  const bool too_many = groups > kMaxGroups;
  groups = min(groups, kMaxGroups);
  int total = 0;                                            // contenders found so far
  // ||u||_2, u.c and sum |u_k c_k| for the bound, each k on one lane
  double un = 0.0, uc = 0.0, uc_abs = 0.0;
  for (int kk = 4 * lane; kk < dim; kk += 128) {
    const float4 uf = ldg_f4(urow + kk);
    const double u0 = (double)uf.x, u1 = (double)uf.y, u2 = (double)uf.z, u3 = (double)uf.w;
    un = fma(u0, u0, un); un = fma(u1, u1, un); un = fma(u2, u2, un); un = fma(u3, u3, un);
    if (center) {
      const float4 cf = ldg_f4(center + kk);
      const double c0 = (double)cf.x, c1 = (double)cf.y, c2 = (double)cf.z, c3 = (double)cf.w;
      uc = fma(u0, c0, uc); uc = fma(u1, c1, uc); uc = fma(u2, c2, uc); uc = fma(u3, c3, uc);
      uc_abs += fabs(u0 * c0) + fabs(u1 * c1) + fabs(u2 * c2) + fabs(u3 * c3);
    }
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    un += __shfl_xor_sync(0xffffffffu, un, off);
    uc += __shfl_xor_sync(0xffffffffu, uc, off);
    uc_abs += __shfl_xor_sync(0xffffffffu, uc_abs, off);
  }
  un = sqrt(un);
  // for any item j outside the kept groups:  u.x_j = u.(x_j - c) + u.c <= thr/(su*si) + eps + u.c =: cut
  const double eps = 1.1 * 0.0009765625 * un * max_item_norm + (double)dim * 0.00390625 * inv_scale + 1e-12 * uc_abs;
  const double cut = (double)thr * inv_scale + eps + uc;

  // 2. exact fp64 scores (k = 0..dim-1 fma chain per item) of the items of the kept groups, 32 items
  //    per round: the rows are fetched by half warps (one coalesced 256-byte request per row) into a
  //    padded shared-memory tile, then lane j runs the chain of item j out of the tile.
  // 3. contenders = rescored items strictly above the cut, appended to the warp's list.
  float* tile = s_tile[wib];
  const int half = lane >> 4, sub = lane & 15;
  const int num_cand_items = groups * kGroup;
  for (int base = 0; base < num_cand_items || base == 0; base += 32) {
    const int it = base + lane;
    int item = -1;
    if (it < num_cand_items) item = (int)s_col[wib][it / kGroup] + (it % kGroup);
    bool live = item >= 0 && item < num_items_local;        // columns past the catalog are zero padding
    if (live && ex_lo < ex_hi && in_sorted(excl_items, ex_lo, ex_hi, item_begin + (int64_t)item)) live = false;
    if (dim == kDim) {
#pragma unroll 4
      for (int step = 0; step < 16; ++step) {
        const int l = 2 * step + half;
        const int il = __shfl_sync(0xffffffffu, live ? item : -1, l);
        if (il >= 0) {
          const float4 v = ldg_f4(item_emb + (size_t)il * kDim + sub * 4);
          float* t = tile + l * kTileStride + sub * 4;
          t[0] = v.x; t[1] = v.y; t[2] = v.z; t[3] = v.w;
        }
      }
      __syncwarp();
    }
    double acc = 0.0;
    if (live) {
      if (dim == kDim) {
        const float* t = tile + lane * kTileStride;
#pragma unroll 4
        for (int kk = 0; kk < kDim; kk += 4) {
          const float4 uf = ldg_f4(urow + kk);
          acc = fma((double)uf.x, (double)t[kk], acc);
          acc = fma((double)uf.y, (double)t[kk + 1], acc);
          acc = fma((double)uf.z, (double)t[kk + 2], acc);
          acc = fma((double)uf.w, (double)t[kk + 3], acc);
        }
      } else {
        const float* irow = item_emb + (size_t)item * dim;
        for (int kk = 0; kk < dim; kk += 4) {
          const float4 uf = ldg_f4(urow + kk), v = ldg_f4(irow + kk);
          acc = fma((double)uf.x, (double)v.x, acc);
          acc = fma((double)uf.y, (double)v.y, acc);
          acc = fma((double)uf.z, (double)v.z, acc);
          acc = fma((double)uf.w, (double)v.w, acc);
        }
      }
    }
    const bool c = live && acc > cut;
    const unsigned mask = __ballot_sync(0xffffffffu, c);
    const int pos = total + __popc(mask & ((1u << lane) - 1u));
    if (c && pos < kMaxContenders) { s_sc[wib][pos] = acc; s_id[wib][pos] = item; }
    total += __popc(mask);
    __syncwarp();
  }
  __syncwarp();
  if (total <= 32) {
    // the rule: one contender per lane, one warp-wide bitonic sort
    double ms = lane < total ? s_sc[wib][lane] : -INFINITY;
    int mi = lane < total ? s_id[wib][lane] : INT32_MAX;
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, ms, stride);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, stride);
        const bool lower = (lane & stride) == 0;                 // lower lane of the pair
        const bool desc = (lane & size) == 0;                    // this block sorts best-first
        const bool other_first = before32(os, oi, ms, mi);
        // the lower lane keeps the better entry in a best-first block, the worse one otherwise
        const bool take = (lower == desc) ? other_first : !other_first && !(os == ms && oi == mi);
        if (take) { ms = os; mi = oi; }
      }
    }
    if (lane < k) {
      out_ids[(size_t)b * k + lane] = mi == INT32_MAX ? INT64_MAX : item_begin + (int64_t)mi;
      out_scores[(size_t)b * k + lane] = ms;
    }
  } else if (total <= kMaxContenders) {
    // the exception (tau ended far below the k-th score): k rounds of warp argmax over the list
    for (int t = 0; t < k; ++t) {
      double bs = -INFINITY;
      int bi = INT32_MAX, bp = -1;
      for (int p = lane; p < total; p += 32) {
        const double x = s_sc[wib][p];
        const int xi = s_id[wib][p];
        if (before32(x, xi, bs, bi)) { bs = x; bi = xi; bp = p; }
      }
      double ws = bs;
      int wi = bi;
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, ws, off);
        const int oi = __shfl_xor_sync(0xffffffffu, wi, off);
        if (before32(os, oi, ws, wi)) { ws = os; wi = oi; }
      }
      if (bp >= 0 && wi == bi && ws == bs) { s_sc[wib][bp] = -INFINITY; s_id[wib][bp] = INT32_MAX; }
      __syncwarp();
      if (lane == 0) {
        out_ids[(size_t)b * k + t] = item_begin + (int64_t)wi;
        out_scores[(size_t)b * k + t] = ws;
      }
    }
  }
  if (lane == 0) {
    // bit 0: provably exact; bits 1.. say why not (list overflow, > 32 groups, < k contenders, > 32 contenders)
    const int why = (raw > cap ? 2 : 0) | (too_many ? 4 : 0) | (total < k ? 8 : 0) | (total > kMaxContenders ? 16 : 0);
    certified[b] = why == 0 ? 1 : why;
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

extern "C" int hnm_absmax(const float* emb, int64_t count, const float* center, int32_t dim, float* out_absmax,
                          void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!emb || !out_absmax) return HNM_E_NULL;
  if (count <= 0 || (center && dim <= 0)) return HNM_E_RANGE;
  const int T = 256;
  const unsigned grid = (unsigned)std::min<int64_t>((count + T - 1) / T, (int64_t)hnm_num_sms() * 8);
  absmax_kernel<<<grid, T, 0, stream>>>(emb, count, center, dim, out_absmax);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_score_pack(const float* emb, const int64_t* row_ids, int64_t num_rows, int64_t rows_padded,
                              int32_t dim, const float* center, float scale, void* out_f16, float* out_sumsq,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!emb || !out_f16) return HNM_E_NULL;
  if (dim != kDim) return HNM_E_DIM;
  if (num_rows < 0 || rows_padded < num_rows || rows_padded <= 0) return HNM_E_RANGE;
  if (!hnm_aligned16(emb) || !hnm_aligned16(out_f16) || (center && !hnm_aligned16(center))) return HNM_E_ALIGN;
  const int T = 256;
  const int64_t threads = rows_padded * 8;
  pack_kernel<<<(unsigned)((threads + T - 1) / T), T, 0, stream>>>(emb, row_ids, num_rows, rows_padded, center, scale,
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
  if (users_padded % kUserTile != 0 || items_padded % kItemTile != 0) return HNM_E_RANGE;
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
  static const int debug_mode = getenv("HNM_FUSED_DEBUG") ? atoi(getenv("HNM_FUSED_DEBUG")) : 0;
  static const int boot_tiles = getenv("HNM_FUSED_BOOT") ? std::max(1, atoi(getenv("HNM_FUSED_BOOT"))) : kBootTiles;
  static const int refresh_div = getenv("HNM_FUSED_REFRESH") ? std::max(1, atoi(getenv("HNM_FUSED_REFRESH"))) : 4;
  const int num_user_tiles = (int)(users_padded / kUserTile);
  const int num_tiles = (int)(items_padded / kItemTile);
  const int grid = std::min((num_user_tiles + kMU - 1) / kMU, hnm_num_sms());
  score_topk_fused_kernel<<<grid, kThreads, kSmemBytes, stream>>>(map_u, map_i, (int)num_users, num_user_tiles,
                                                                  num_tiles, kth_sel, (uint2*)cand,
                                                                  cand_cap, cand_count, cand_thresh, debug_mode, boot_tiles, refresh_div);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_rescore_topk(const float* user_emb, const float* item_emb, const int64_t* user_ids, int64_t batch,
                                int32_t dim, int64_t item_begin, int64_t num_items_local, const void* cand,
                                int32_t cand_cap,
                                const int32_t* cand_count, const float* cand_thresh, double inv_scale_product,
                                double max_item_norm, const float* center, const int64_t* excl_ptr,
                                const int64_t* excl_items, int32_t k, int64_t* out_ids, double* out_scores,
                                int32_t* out_certified, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (batch == 0) return HNM_OK;
  if (!user_emb || !item_emb || !cand || !cand_count || !cand_thresh || !out_ids || !out_scores || !out_certified)
    return HNM_E_NULL;
  if ((excl_ptr != nullptr) != (excl_items != nullptr)) return HNM_E_NULL;
  if (batch < 0 || dim <= 0 || dim % 4 != 0 || k < 1 || k > 32 || cand_cap < 1 || cand_cap > 32 * kMaxPerLane ||
      num_items_local < 1 || num_items_local > INT32_MAX)
    return HNM_E_RANGE;
  const int wpc = kRescoreWarps;
  rescore_kernel<<<(unsigned)((batch + wpc - 1) / wpc), wpc * 32, 0, stream>>>(
      user_emb, item_emb, user_ids, batch, dim, item_begin, (int)num_items_local, (const uint2*)cand, cand_cap,
      cand_count, cand_thresh,
      inv_scale_product, max_item_norm, center, excl_ptr, excl_items, k, out_ids, out_scores, out_certified);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}
