// Fused full-catalog score + candidate select on the 5th-gen tensor cores (sm_100a).
// Replaces the `torch.matmul(user_embeds, item_embeddings.t())` + `torch.topk` pair of
// src/models/lightgcn.py:202,356 without ever writing the [users, items] score matrix.
//
//   hnm_absmax / hnm_score_pack_items / hnm_score_pack_users : fp32 table -> power-of-two scaled fp16, zero
//                                 padded (one scale per item shard, one per user row; scales stay on the device)
//   hnm_score_topk_fused        : TMA -> smem (128B swizzle) -> tcgen05.mma (fp16 x fp16 -> fp32
//                                 accumulators in TMEM) -> tcgen05.ld epilogue that keeps, per user,
//                                 every item whose approximate score beats a running threshold
//   merge_split_kernel          : (inside hnm_score_topk_fused) one list + one threshold per user out of the
//                                 per-slice lists of the user tiles whose item range was sliced over the CTAs
//   hnm_rescore_topk            : exact fp64 scores of the survivors, canonical top-k, certificate
//
// Kernel shape: one persistent CTA per SM, 20 warps = 640 threads:
//   warp 0   TMA producer      A: 2 user tiles x KC chunks of [128 x 64] fp16 per pass (double buffered up to
//                              d = 128), B: (item tile, K chunk) stages [128 x 64] fp16 through a ring
//   warps 1-2  MMA issuers     warp 1+m, one elected thread, issues user tile m: per item tile 4 x KC
//                              tcgen05.mma M128 N128 K16 into one of the tile's two accumulators
//                              (2 x 2 x 128 = all 512 TMEM columns); warp 2 also owns the TMEM allocation
//   warp 3     idle; the four control warps hand registers to the epilogue (setmaxnreg)
//   warps 4..19  epilogue: 2 threads per user row (TMEM lane), each draining 64 of an accumulator's 128
//                columns; the two threads of a row share the row's threshold through shared memory and a
//                64-thread named barrier.
// While the epilogue drains one accumulator of a user tile the tensor pipe fills the other one.
//
// Select (per user row): 32 bucket maxima in shared memory (bucket = position of a 4-column group inside an
// item tile), 16 per thread of the row.  Bucket maxima belong to distinct items, so the kth_sel-th largest of
// them, tau, is a lower bound on the kth_sel-th best score seen so far.  The hot path of a 32-column chunk is
// 8 group maxima, their maximum and one compare; only a chunk that beats tau updates the buckets and appends
// ONE entry to the user's candidate list in global memory (the 8 group maxima cut to bf16 + the first column).
// tau is refreshed (32-wide sorting network) every time the number of item tiles seen has grown by 1/4.  The
// first seed tiles (kBootTiles, more for large catalogs) are run twice: once to seed the buckets, once to collect.
//
// Work distribution (SplitPlan below): whole passes of 2 user tiles x the whole catalog per CTA, and the
// left-over tiles in pairs whose item range is cut into slices, one (pair, slice) per CTA; merge_split_kernel
// then builds one list per sliced user.
#include <algorithm>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include "common.cuh"

namespace {

constexpr int kDim = HNM_FUSED_DIM;         // K chunk: 64 fp16 = one 128-byte swizzle row; embedding dim = KC chunks
constexpr int kUserTile = 128;              // UMMA M
constexpr int kItemTile = 128;              // UMMA N
constexpr int kBootTiles = 16;              // item tiles used to seed the bucket maxima (run twice)
constexpr int kTileBytes = kItemTile * kDim * 2; // 16384 (A tile and B tile have the same shape)
constexpr int kNumBuckets = 32;
constexpr uint32_t kWaitHintNs = 0;         // suspend-time hint of the control warps' mbarrier waits (0 = none)

// Shape of the persistent CTA (round 2):
//   2 user tiles per pass, 2 accumulators (128 TMEM columns) each = all 512 columns: the tensor pipe fills one
//   buffer of a user tile while the other one is drained (round 1: 3 tiles x 1 accumulator, where the chain
//   issue 4 MMAs -> complete -> wake -> 2 TMEM round trips -> arrive -> wake set the step time);
//   2 threads per user row, each draining 64 of an accumulator's 128 columns: 16 epilogue warps = 4 per
//   scheduler.  ncu of the one-thread-per-row form: ALU pipe 62 % busy, half of the epilogue's stall samples
//   fixed-latency waits with 2 warps per scheduler to hide them behind.
constexpr int kMU = 2;                      // user tiles per CTA pass
constexpr int kBUF = 2;                     // accumulators per user tile
constexpr int kHalves = 2;                  // threads per user row (column halves of an accumulator)
constexpr int kSlots = kMU * kBUF;
constexpr int kStagesB = 6;                 // B ring (d = 64); the wider shapes take what the A tiles leave
constexpr int kEpiWarps = 4 * kMU * kHalves;      // 16
constexpr int kThreads = (4 + kEpiWarps) * 32;    // 640
constexpr int kHalfBuckets = kNumBuckets / kHalves;
constexpr int kCtlRegs = 64, kEpiRegs = 104;      // setmaxnreg: the CTA is launched with 96 registers x 640 threads = 61 440, and
                                                  // 128 * 64 + 512 * 104 must not exceed THAT (not 64 K): a warpgroup whose
                                                  // setmaxnreg.inc cannot be served waits forever
static_assert(128 * kCtlRegs + 512 * kEpiRegs <= 96 * kThreads, "register pool");

static_assert(kUserTile == HNM_FUSED_USER_TILE && kItemTile == HNM_FUSED_ITEM_TILE, "header mismatch");
static_assert(kSlots * kItemTile <= 512, "TMEM columns");

struct __align__(8) Barriers {
  uint64_t a_full[2], a_empty[2];
  uint64_t b_full[kStagesB], b_empty[kStagesB];
  uint64_t t_full[kSlots], t_empty[kSlots];
  uint32_t tmem_base;
};
// Embedding dimension d = 64 KC (KC K-chunks of one 128-byte swizzle row each; BASELINE.json configs[4] is d = 256).
// The user tiles of a pass stay in shared memory for the whole catalog sweep: KC x 16 KB per tile.  d = 64 and
// 128 double-buffer them across passes; at d = 256 (128 KB for the two tiles) there is one buffer, and the few
// microseconds the producer waits at a pass boundary are nothing against the ~2 ms of a pass.  The B ring holds
// (item tile, chunk) stages.
template <int KC> struct KShape {
  static_assert(KC == 1 || KC == 2 || KC == 4, "d = 64, 128 or 256");
  static constexpr int kAbufs = KC == 4 ? 1 : 2;
  static constexpr int kStages = KC == 1 ? 6 : 4;    // 128 KB of user tiles + 33 KB of buckets leave room for four
  static_assert(kStages <= kStagesB, "barrier arrays");
};
// The 32 bucket maxima of a user row live in shared memory, 16 per thread of the row: they are touched only
// when a chunk beats tau (and by the threshold refresh), and in registers they cost the epilogue the room
// it needs to keep its list pointers out of the constant bank.  [user tile][quarter][half][bucket][lane]
struct PairBuckets {
  float bm[kMU][4][kHalves][kHalfBuckets][32];
  float tau[kMU][4][32];
};
template <int KC> constexpr size_t smem_bytes() {
  return 1024 /*align slack*/ + (size_t)KShape<KC>::kAbufs * kMU * KC * kTileBytes + KShape<KC>::kStages * kTileBytes +
         sizeof(PairBuckets) + sizeof(Barriers);
}
static_assert(smem_bytes<4>() <= 227 * 1024, "shared memory");

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
// The same wait with a suspend-time hint: the thread may stay descheduled for up to `ns` nanoseconds (it is
// woken as soon as the phase completes).  The four control warps sit in these waits most of the time; with the
// default (short) time limit their polling loops were 74 of the 170 warp instructions issued per 32-column
// chunk (profiles/r1_fused_notes.md), competing with the epilogue warps of the same scheduler for issue slots.
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
  if (ns == 0) { mbar_wait(bar, parity); return; }
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(ns)
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
// One of the two threads of a user row: it sees columns [64 h, 64 h + 64) of every item tile.
struct RowState {
  float tau;                  // collect threshold (+inf while seeding), common to both threads of the row
  int cnt;                    // chunks this thread appended so far (may exceed its capacity: overflow)
  uint32_t bm;                // shared-memory address of this thread's 16 bucket maxima: bucket b at bm + 128 b
  const uint32_t* sig;        // the row's exclusion signature (kSigWords words) or nullptr
};
// Purchased-item filter (src/models/lightgcn.py:349-353, the serving default scripts/serve.py:350-352): the
// excluded items are exactly a user's best-scoring ones, so a threshold that tracks the (k+3)-th best score of
// ALL items leaves fewer than k contenders after the filter.  Each row therefore carries a 1 024-bit signature
// of the 32-column chunks that hold an excluded item (bit = chunk index mod 1 024); such chunks are still
// nominated, but they never feed the bucket maxima, so tau is a lower bound on the kth best score among items
// of unflagged chunks -- all of them allowed.  A false positive (two chunks sharing a bit) only costs a little
// tightness.  hnm_rescore_topk applies the filter exactly.
constexpr int kSigWords = HNM_FUSED_SIG_WORDS;
__device__ __forceinline__ bool chunk_flagged(const uint32_t* __restrict__ sig, int col0) {
  const uint32_t chunk = (uint32_t)col0 >> 5;
  return (__ldg(sig + ((chunk >> 5) & (kSigWords - 1))) >> (chunk & 31u)) & 1u;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v)); }
// bucket <- max(bucket, q) for the 8 groups of chunk `chunk` (all loads first: one shared-memory round trip)
__device__ __forceinline__ void bucket_update(uint32_t bm, int chunk, const float (&q)[8]) {
  float b[8];
#pragma unroll
  for (int h = 0; h < 8; ++h) b[h] = lds_f32(bm + (chunk * 8 + h) * 128);
#pragma unroll
  for (int h = 0; h < 8; ++h) sts_f32(bm + (chunk * 8 + h) * 128, fmaxf(b[h], q[h]));
}

__device__ __forceinline__ void cmpx(float& a, float& b) {   // a <- max, b <- min
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  a = hi;
  b = lo;
}

// kth largest of 32 values: bitonic sorting network (descending, 240 comparators) on a scratch copy.
__device__ __forceinline__ float kth_largest32(float (&t)[kNumBuckets], int kth) {
#pragma unroll
  for (int k = 2; k <= kNumBuckets; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < kNumBuckets; ++i) {
        const int l = i ^ j;
        if (l > i) {
          if ((i & k) == 0) cmpx(t[i], t[l]);
          else cmpx(t[l], t[i]);
        }
      }
    }
  }
  float r = t[0];
#pragma unroll
  for (int i = 1; i < kNumBuckets; ++i) r = (i < kth) ? t[i] : r;
  return r;
}

__device__ __forceinline__ void pair_bar(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

// tau = the kth largest of the row's 32 bucket maxima (bucket = position of a 4-column group inside an item
// tile).  The buckets hold disjoint item sets, so tau is a lower bound on the kth best score seen.  The
// two warps of a row meet at a 64-thread named barrier (the schedule of refreshes is the same function of
// the pass for both); thread h = 0 sorts a register copy of all 32 and hands tau back through shared memory.
// The buckets themselves are never permuted, so revisiting an item (the seed tiles are swept twice) is
// idempotent.
__device__ __forceinline__ float refresh_tau(uint32_t row_bm /* smem: [2][16][32] of this (tile, quarter), + lane */,
                                             int h, uint32_t xt /* smem: tau slot of this lane */, int bar_id, int kth) {
  pair_bar(bar_id);                                   // both threads' bucket updates are visible
  if (h == 0) {
    float t[kNumBuckets];
#pragma unroll
    for (int i = 0; i < kNumBuckets; ++i) t[i] = lds_f32(row_bm + i * 128);
    sts_f32(xt, kth_largest32(t, kth));
  }
  pair_bar(bar_id);
  return lds_f32(xt);
}

// One nominated 32-column chunk of one user row: the chunk's eight group maxima (groups of 4 adjacent
// items), each cut to its upper 16 bits (bf16, truncated toward zero; q.x = {low half: group 0, high half:
// group 1}, ...), and the LOCAL index of the chunk's first item.  The two parts live in two arrays so that
// an entry costs 20 bytes of DRAM traffic and both are read back coalesced.
struct CandList {
  uint4* q;           // [rows * cap]
  uint32_t* col;      // [rows * cap]
  __host__ __device__ CandList at(size_t off) const { return CandList{q + off, col + off}; }
};
static_assert(HNM_FUSED_CAND_BYTES == sizeof(uint4) + sizeof(uint32_t), "header mismatch");
static inline CandList cand_list(void* base, size_t rows, int cap) {
  return CandList{reinterpret_cast<uint4*>(base),
                  reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(base) + rows * (size_t)cap * sizeof(uint4))};
}

// A stored 16-bit pattern t stands for the interval [lo, hi] that contains the group maximum q:
// truncation toward zero gives t <= q < t + ulp for q >= 0 and t - ulp < q <= t for q < 0.
__device__ __forceinline__ void cand_bounds(uint32_t h16, float& lo, float& hi) {
  const float t = __uint_as_float(h16 << 16);
  const float far = __uint_as_float((h16 + 1u) << 16);      // one bf16 step away from zero (may be +-inf)
  const bool neg = (h16 & 0x8000u) != 0u;
  lo = neg ? far : t;
  hi = neg ? t : far;
}

// 32 accumulator columns of one user row = 8 groups of 4 columns.  The hot path is the 8 group maxima, the
// chunk maximum and ONE compare-and-branch: 21 ALU-pipe instructions (FMNMX / FMNMX3 issue every other
// cycle per scheduler).  Everything else happens only for the rows whose chunk beats their threshold tau:
//   * the chunk's group maxima go into the 32 bucket maxima (bucket = position of the group inside an item
//     tile).  Skipping the chunks at or below tau changes nothing: tau is the kth-largest bucket maximum,
//     kth buckets already hold values >= tau, and a value <= tau cannot move the kth-largest upwards;
//   * ONE 32-byte entry {8 truncated group maxima, first column} is appended to the row's candidate list.
//     hnm_rescore_topk decides per group from the interval the 16 bits stand for.
// Round 1 stored one {max, column} pair per group above tau with eight predicated store sequences and
// updated the buckets on the hot path: ~96 warp instructions per chunk of which the hot path was a third
// (tools/bench_select_epilogue.cu: 247 of ~780 cycles per accumulator quarter and scheduler).
// The four group maxima pairs of a 32-column chunk; after this the 32 accumulator values are dead,
// so their registers can take the next tcgen05.ld while the rest of the chunk is processed.
__device__ __forceinline__ void group_max(const float (&v)[32], float (&q)[8]) {
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    const float* x = v + 4 * h;
    q[h] = fmaxf(fmaxf(fmaxf(x[0], x[1]), x[2]), x[3]);
  }
}

enum { kSeed = 0, kCollect = 1 };
// kSeed: first visit of the seed tiles -- buckets only (tau is +inf).  kCollect: everything else.
template <int MODE>
__device__ __forceinline__ void finish32(const float (&q)[8], int chunk, int col0, RowState& st,
                                         const CandList cand, int cap, int xp = 0) {
  if (MODE == kSeed) {
    if (st.sig == nullptr || !chunk_flagged(st.sig, col0)) bucket_update(st.bm, chunk, q);
    return;
  }
  const float m32 = fmaxf(fmaxf(fmaxf(fmaxf(q[0], q[1]), q[2]), fmaxf(fmaxf(q[3], q[4]), q[5])), fmaxf(q[6], q[7]));
  if (m32 > st.tau) {
    if (!(xp & 16) && (st.sig == nullptr || !chunk_flagged(st.sig, col0))) bucket_update(st.bm, chunk, q);
    if (xp & 8) { ++st.cnt; return; }    // experiment: count, do not store
    if (st.cnt >= cap) {                 // no room: stop collecting, flag the row
      st.tau = INFINITY;
      st.cnt = cap + 1;
    } else {
      const uint32_t p0 = __byte_perm(__float_as_uint(q[0]), __float_as_uint(q[1]), 0x7632);
      const uint32_t p1 = __byte_perm(__float_as_uint(q[2]), __float_as_uint(q[3]), 0x7632);
      const uint32_t p2 = __byte_perm(__float_as_uint(q[4]), __float_as_uint(q[5]), 0x7632);
      const uint32_t p3 = __byte_perm(__float_as_uint(q[6]), __float_as_uint(q[7]), 0x7632);
      // the list is write-only here
      asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(cand.q + st.cnt), "r"(p0), "r"(p1), "r"(p2), "r"(p3) : "memory");
      asm volatile("st.global.b32 [%0], %1;" ::"l"(cand.col + st.cnt), "r"(col0) : "memory");
      ++st.cnt;
    }
  }
}

// This thread's 64 columns of one accumulator: both 32-column loads are issued together, and the
// accumulator goes back to its MMA issuer as soon as they have landed (the 8 warps of the user tile arrive).
template <int MODE>
__device__ __forceinline__ void drain_half(uint32_t taddr, int col0, RowState& st, const CandList cand, int cap,
                                           uint64_t* t_empty, int lane, int xp = 0) {
  float va[32], vb[32], q[8];
  tmem_ld32(taddr, va);
  tmem_ld32(taddr + 32, vb);
  tmem_ld_wait(va, vb);
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(t_empty);
  group_max(va, q);
  finish32<MODE>(q, 0, col0, st, cand, cap, xp);
  group_max(vb, q);
  finish32<MODE>(q, 1, col0 + 32, st, cand, cap, xp);
}

// ----------------------------------------------------------------------------- work distribution
// A pass = up to kMU user tiles against a range of item tiles, and it costs the same for one tile as for
// three (the three tiles run concurrently).  Dealing whole tiles to the CTAs therefore quantises badly:
// 1 340 tiles over 148 CTAs (one rank of an 8-GPU run) is 9.05 tiles per CTA = 4 passes instead of 3.02.
// So every CTA runs `full_passes` passes of exactly kMU tiles over the whole catalog, and the left-over
// tiles (fewer than kMU per CTA) are grouped in triples whose ITEM range is cut into `slices` slices, one
// (triple, slice) per CTA.  A sliced user gets one candidate list and one tau per slice (split_* buffers);
// merge_split_kernel then keeps the groups above the largest of the slice thresholds -- each slice's tau
// is a lower bound on the kth_sel-th best score of a subset of the catalog, hence also of the catalog.
struct SplitPlan {
  int mu;               // user tiles per pass (Shape::MU)
  int full_passes;      // passes of mu tiles over the whole catalog, per CTA
  int tile0;            // first left-over user tile (= gridDim.x * mu * full_passes)
  int triples;          // ceil(left-over tiles / mu)
  int slices;           // item-range slices per triple; unit (triple, slice) u runs on CTA u % gridDim.x
  CandList cand;        // [left-over users][slices][cap]
  int cap;
  int32_t* count;       // [left-over users][slices]
  float* thresh;        // [left-over users][slices]
};

struct PassDesc {
  int t0, mc;           // user tiles [t0, t0 + mc)
  int i0, ni;           // item tiles [i0, i0 + ni)
  int boot;             // seed tiles (visited twice)
  int rot;              // rotation of the sweep inside the range
  int slice;            // >= 0: outputs go to the split buffers
};

__device__ __forceinline__ bool pass_desc(int n, int num_user_tiles, int num_item_tiles, int boot_tiles,
                                          const SplitPlan& sp, PassDesc& p) {
  const int b = (int)blockIdx.x;
  if (n < sp.full_passes) {
    p.t0 = (b * sp.full_passes + n) * sp.mu;
    p.mc = sp.mu;
    p.i0 = 0;
    p.ni = num_item_tiles;
    // every CTA sweeps the catalog from a different starting tile: otherwise all 148 SMs ask the L2 for
    // the same 16 KB item tile at the same moment
    p.rot = (int)(((long long)b * num_item_tiles) / (int)gridDim.x);
    p.slice = -1;
    p.boot = min(num_item_tiles, boot_tiles);
    return true;
  }
  // left-over units (pair of user tiles, item slice): unit u = b, b + grid, ... (SplitPlan: several rounds when
  // that balances better than one unsliced round on half of the CTAs)
  const int unit = b + (n - sp.full_passes) * (int)gridDim.x;
  if (unit >= sp.triples * sp.slices) return false;
  const int j = unit / sp.slices, sl = unit - j * sp.slices;
  p.t0 = sp.tile0 + j * sp.mu;
  p.mc = min(sp.mu, num_user_tiles - p.t0);
  p.i0 = (int)(((long long)sl * num_item_tiles) / sp.slices);
  p.ni = (int)(((long long)(sl + 1) * num_item_tiles) / sp.slices) - p.i0;
  p.rot = 0;
  if (sp.slices > 1) {
    p.slice = sl;
    p.boot = max(1, min(boot_tiles, p.ni / 4));
  } else {
    p.slice = -1;
    p.boot = min(p.ni, boot_tiles);
  }
  return true;
}

// index of a row's {lists, counts, tau}: the row itself, or (left-over row, slice)
__device__ __forceinline__ uint32_t out_slot(const PassDesc& p, const SplitPlan& sp, int m, int q, int lane) {
  const int row = (p.t0 + m) * kUserTile + q * 32 + lane;
  return p.slice < 0 ? (uint32_t)row : (uint32_t)((row - sp.tile0 * kUserTile) * sp.slices + p.slice);
}

__device__ __forceinline__ int pass_tile(const PassDesc& p, int it) {
  int t = (it < p.boot ? it : it - p.boot) + p.rot;
  if (t >= p.ni) t -= p.ni;
  return p.i0 + t;
}

// ----------------------------------------------------------------------------- the kernel
// Candidate storage of a row: `cap` entries, the first cap/2 for thread h = 0 (columns 0..63 of every item
// tile), the rest for h = 1; cand_count[2 row + h] entries are valid in each half.
template <int KC>
__global__ void __launch_bounds__(kThreads, 1)
score_topk_fused_kernel(const __grid_constant__ CUtensorMap map_users, const __grid_constant__ CUtensorMap map_items,
                        int num_users, int num_user_tiles, int num_item_tiles, int kth_sel,
                        const CandList cand, int cap, int32_t* __restrict__ cand_count,
                        float* __restrict__ cand_thresh, const uint32_t* __restrict__ excl_sig, int mode,
                        int boot_tiles, int refresh_div, uint32_t wait_hint_ns, const SplitPlan sp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kAbufs = KShape<KC>::kAbufs, kStagesB = KShape<KC>::kStages;
  uint8_t* smem_a = smem;                                   // [kAbufs][kMU][KC][kTileBytes]
  uint8_t* smem_b = smem + kAbufs * kMU * KC * kTileBytes;  // [kStagesB][kTileBytes]
  PairBuckets* buckets = reinterpret_cast<PairBuckets*>(smem_b + kStagesB * kTileBytes);
  Barriers* bars = reinterpret_cast<Barriers*>(buckets + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  PassDesc p;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_users);
    tma_prefetch_desc(&map_items);
    for (int i = 0; i < 2; ++i) { mbar_init(&bars->a_full[i], 1); mbar_init(&bars->a_empty[i], kMU); }
    for (int i = 0; i < kStagesB; ++i) { mbar_init(&bars->b_full[i], 1); mbar_init(&bars->b_empty[i], kMU); }
    for (int i = 0; i < kSlots; ++i) { mbar_init(&bars->t_full[i], 1); mbar_init(&bars->t_empty[i], 4 * kHalves); }
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
  // The control warpgroup gives registers back so that the 16 epilogue warps can hold two 32-column loads,
  // their 16 bucket maxima and the 32-value scratch of the threshold refresh without spilling.
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtlRegs));
  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      uint32_t g = 0;
      for (int n = 0; pass_desc(n, num_user_tiles, num_item_tiles, boot_tiles, sp, p); ++n) {
        const int mc = p.mc;
        const int abuf = n % kAbufs;
        mbar_wait_hint(&bars->a_empty[abuf], ((n / kAbufs) & 1) ^ 1, wait_hint_ns);
        mbar_expect_tx(&bars->a_full[abuf], mc * KC * kTileBytes);
        for (int m = 0; m < mc; ++m)
          for (int kc = 0; kc < KC; ++kc)
            tma_load_2d(smem_a + ((abuf * kMU + m) * KC + kc) * kTileBytes, &map_users, &bars->a_full[abuf], kc * kDim,
                        (p.t0 + m) * kUserTile);
        const int num_iters = p.ni + p.boot;
        for (int it = 0; it < num_iters; ++it) {
          const int tile = pass_tile(p, it);
          for (int kc = 0; kc < KC; ++kc, ++g) {
            const int stage = g % kStagesB;
            mbar_wait_hint(&bars->b_empty[stage], ((g / kStagesB) & 1) ^ 1, wait_hint_ns);
            mbar_expect_tx(&bars->b_full[stage], kTileBytes);
            tma_load_2d(smem_b + stage * kTileBytes, &map_items, &bars->b_full[stage], kc * kDim, tile * kItemTile);
          }
        }
      }
    }
  } else if (warp <= kMU) {
    // ===================================================== MMA issuers: warp 1 + m serves user tile m
    // tcgen05.mma holds its issuing thread for about the duration of the MMA (tools/bench_mma.cu:
    // 73 cycles per M128 N128 K16), so with a single issuer every mbarrier wait / fence / commit adds
    // to the tensor pipe's critical path.  One issuing thread per user tile.
    const int m = warp - 1;
    if (lane == 0) {
      uint32_t g = 0, uses = 0;
      const uint64_t desc_hi = umma_desc_sw128(0) & ~uint64_t(0x3FFF);
      for (int n = 0; pass_desc(n, num_user_tiles, num_item_tiles, boot_tiles, sp, p); ++n) {
        const int num_iters = p.ni + p.boot;
        const bool active = m < p.mc;
        const int abuf = n % kAbufs;
        mbar_wait_hint(&bars->a_full[abuf], (n / kAbufs) & 1, wait_hint_ns);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + (abuf * kMU + m) * KC * kTileBytes);
        for (int it = 0; it < num_iters; ++it) {
          // the accumulators of user tile m belong to it alone (use counter `uses`), so no issuer ever
          // has to order itself against another one
          const int slot = m * kBUF + (int)(uses % kBUF);
          const uint32_t d_tmem = tmem_base + slot * kItemTile;
          if (active) {
            mbar_wait_hint(&bars->t_empty[slot], ((uses / kBUF) & 1) ^ 1, wait_hint_ns);
            tc_fence_after();
          }
#pragma unroll
          for (int kc = 0; kc < KC; ++kc, ++g) {
            const int stage = g % kStagesB;
            mbar_wait_hint(&bars->b_full[stage], (g / kStagesB) & 1, wait_hint_ns);
            if (active) {
              tc_fence_after();
              const uint64_t a_desc = desc_hi | (uint64_t)(((a_addr + kc * kTileBytes) >> 4) & 0x3FFF);
              const uint64_t b_desc = desc_hi | (uint64_t)((smem_u32(smem_b + stage * kTileBytes) >> 4) & 0x3FFF);
#pragma unroll
              for (int k = 0; k < kDim / 16; ++k)      // +32 bytes along K = +2 in the 16-byte address field
                umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, (kc | k) > 0 ? 1u : 0u);
              umma_commit(&bars->b_empty[stage]);
            } else {
              mbar_arrive(&bars->b_empty[stage]);       // keep the stage's arrival count at kMU
            }
          }
          if (active) {
            umma_commit(&bars->t_full[slot]);
            ++uses;
          }
        }
        if (active) umma_commit(&bars->a_empty[abuf]);
        else mbar_arrive(&bars->a_empty[abuf]);
      }
    }
  }
  } else {
    // ===================================================== epilogue: 2 warpgroups (column halves) per user tile
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    const int e = warp - 4;
    const int m = e >> 3;                          // user tile
    const int h = (e >> 2) & 1;                    // column half
    const int q = warp & 3;                        // TMEM lane quarter this warp may read
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + h * (kItemTile / kHalves);
    const uint32_t row_bm = smem_u32(&buckets->bm[m][q][0][0][lane]);     // both threads' buckets: 32 x stride 128 B
    const uint32_t xt = smem_u32(&buckets->tau[m][q][lane]);
    const int bar_id = 1 + m * 4 + q;              // barrier 0 is __syncthreads
    uint32_t uses = 0;
    for (int n = 0; pass_desc(n, num_user_tiles, num_item_tiles, boot_tiles, sp, p); ++n) {
      if (m >= p.mc) continue;                                              // short pass: these warpgroups rest
      const int boot = p.boot;
      const int num_iters = p.ni + boot;
      // where this thread's candidates go: its half of the row's list, or of the row's list for this item slice
      const uint32_t slot = out_slot(p, sp, m, q, lane);
      const bool real = (p.t0 + m) * kUserTile + q * 32 + lane < num_users;
      const int out_cap = p.slice < 0 ? cap : sp.cap;
      const int half_cap = out_cap / kHalves;
      CandList my_cand = (p.slice < 0 ? cand : sp.cand).at((size_t)(real ? slot : 0) * out_cap + h * half_cap);
      // opaque to the compiler, so that it keeps the two pointers in registers instead of re-deriving them from
      // the kernel parameters (six constant-bank loads with their latency) inside the hit path
      asm volatile("" : "+l"(my_cand.q), "+l"(my_cand.col));
      const int my_cap = real ? half_cap : 0;            // padded rows count but never store
      int32_t* out_count = (p.slice < 0 ? cand_count : sp.count) + (size_t)slot * kHalves + h;
      float* out_thresh = (p.slice < 0 ? cand_thresh : sp.thresh) + slot;
      RowState rs;
      rs.tau = INFINITY;
      rs.cnt = 0;
      rs.bm = smem_u32(&buckets->bm[m][q][h][0][lane]);
      rs.sig = (excl_sig != nullptr && real)
                   ? excl_sig + (size_t)((p.t0 + m) * kUserTile + q * 32 + lane) * kSigWords : nullptr;
#pragma unroll
      for (int i = 0; i < kHalfBuckets; ++i) sts_f32(rs.bm + i * 128, -INFINITY);
      // item tile of this iteration.  Only whole-catalog passes are rotated, so the sweep wraps at the end of
      // the catalog in both kinds of pass (a slice never gets there before its last iteration).
      const int restart = p.i0 + p.rot;
      int cur = restart;
      int it = 0;
      // one accumulator: wait for it, drain this thread's half
#define HNM_TILE_STEP(MODE)                                                                        \
      {                                                                                            \
        const int col0 = cur * kItemTile + h * (kItemTile / kHalves);                              \
        if (++cur == num_item_tiles) cur = 0;                                                      \
        const int slot_t = m * kBUF + (int)(uses & 1u);                                            \
        mbar_wait(&bars->t_full[slot_t], (uses >> 1) & 1u);                                        \
        ++uses;                                                                                    \
        tc_fence_after();                                                                          \
        drain_half<MODE>(lane_base + slot_t * kItemTile, col0, rs, my_cand, my_cap,                \
                         &bars->t_empty[slot_t], lane, mode);                                      \
      }
      if ((mode & 7) != 0) {
        // debug shapes of the pipeline: 3 = handshakes only, 4 = one load, 1 = both loads, 2 = + the hot path
        for (; it < num_iters; ++it) {
          const int slot_t = m * kBUF + (int)(uses & 1u);
          if ((mode & 7) == 2) { HNM_TILE_STEP(kSeed) continue; }
          mbar_wait(&bars->t_full[slot_t], (uses >> 1) & 1u);
          ++uses;
          tc_fence_after();
          float va[32];
          float acc = 0.f;
          const int nld = (mode & 7) == 1 ? 2 : ((mode & 7) == 4 ? 1 : 0);
          for (int c = 0; c < nld; ++c) { tmem_ld32(lane_base + slot_t * kItemTile + c * 32, va); tmem_ld_wait(va); acc += va[c]; }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->t_empty[slot_t]);
          sts_f32(rs.bm, acc);
        }
      } else {
        for (; it < boot; ++it) HNM_TILE_STEP(kSeed)     // the seed tiles: buckets only
        cur = restart;                                    // ... and they are swept a second time, collecting
        while (it < num_iters) {
          if (!(mode & 32) || it == boot) rs.tau = refresh_tau(row_bm, h, xt, bar_id, kth_sel);
          const int seen = max(boot, it - boot);         // item tiles behind the current bucket maxima
          const int stop = min(num_iters, it + max(2, seen / refresh_div));
          for (; it < stop; ++it) HNM_TILE_STEP(kCollect)
        }
        rs.tau = refresh_tau(row_bm, h, xt, bar_id, kth_sel);
      }
#undef HNM_TILE_STEP
      if (real) {
        *out_count = rs.cnt;
        if (h == 0) *out_thresh = rs.tau;
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

// One warp per sliced user.  thr0 = max of the slice thresholds is a valid threshold (every group that a
// slice did not store has a maximum <= that slice's tau <= thr0), but a loose one: a slice's tau is the
// kth_sel-th best of a fraction of the catalog.  The stored group maxima belong to distinct items, so the
// kth_sel-th largest of them over all slices is again a lower bound on the kth_sel-th best score, and every
// group above it is stored.  thr = max(thr0, that); the groups above thr are gathered into the user's
// ordinary list.  What hnm_rescore_topk's certificate assumes holds: every item outside the kept groups
// scored <= thr.
__device__ __forceinline__ void cmpx_lane(float& v, int lane, int stride, bool desc) {
  const float o = __shfl_xor_sync(0xffffffffu, v, stride);
  const bool lower = (lane & stride) == 0;
  v = (lower == desc) ? fmaxf(v, o) : fminf(v, o);    // best-first block: the lower lane keeps the larger value
}

// Order-preserving key of a 16-bit float pattern (sign + exponent + 7 mantissa bits): larger key <=> larger value.
__device__ __forceinline__ uint32_t order_key16(uint32_t t) { return t ^ ((t & 0x8000u) ? 0xFFFFu : 0x8000u); }

__device__ __forceinline__ uint32_t cand_half(const uint4& q, int h) {     // 16-bit pattern of group h
  const uint32_t w = h < 2 ? q.x : (h < 4 ? q.y : (h < 6 ? q.z : q.w));
  return (h & 1) ? (w >> 16) : (w & 0xFFFFu);
}

// The sub-lists of a user are short (a slice sees a fraction of the catalog), so walking them one by one leaves
// most lanes idle and chains four dependent global loads per sub-list.  Instead the counts are loaded once,
// lane-parallel, into exclusive prefix sums in shared memory, and both sweeps run over the FLAT entry index
// (32 entries per step whatever sub-list they belong to; a lane finds its sub-list by bisection in shared
// memory), two steps' loads in flight at a time.
constexpr int kMergeBuf = 512;      // survivors buffered per warp before the sorting network runs (>= 2 x 256)
struct FlatLists {
  const int32_t* pre;    // [lists + 1] exclusive prefix sums of the counts (shared memory)
  int lists;
  __device__ __forceinline__ void locate(int f, int& l, int& i) const {      // largest l with pre[l] <= f
    int lo = 0, hi = lists;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (pre[mid] <= f) lo = mid; else hi = mid;
    }
    l = lo;
    i = f - pre[lo];
  }
};

__global__ void __launch_bounds__(128)
merge_split_kernel(SplitPlan sp, int num_users, int kth_sel, const CandList cand, int cap,
                   int32_t* __restrict__ cand_count, float* __restrict__ cand_thresh) {
  extern __shared__ int32_t merge_pre[];
  __shared__ float merge_buf[4][kMergeBuf];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int lu = (int)blockIdx.x * 4 + wib;
  const int row = sp.tile0 * kUserTile + lu;
  if (row >= num_users) return;                      // warps are independent: no block-wide barrier below
  const size_t base = (size_t)lu * sp.slices;
  const int half_cap = sp.cap / kHalves;
  const int lists = sp.slices * kHalves;             // sub-list l = (slice l / 2, column half l % 2)
  int32_t* pre = merge_pre + wib * (lists + 1);
  float thr = -INFINITY;
  bool over = false;
  for (int s = lane; s < sp.slices; s += 32) thr = fmaxf(thr, sp.thresh[base + s]);
  int entries = 0;
  for (int l0 = 0; l0 < lists; l0 += 32) {
    const int l = l0 + lane;
    const int c = l < lists ? sp.count[base * kHalves + l] : 0;
    over |= c > half_cap;
    int incl = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    if (l < lists) pre[l] = entries + incl - c;
    entries += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) pre[lists] = entries;
#pragma unroll
  for (int off = 16; off; off >>= 1) thr = fmaxf(thr, __shfl_xor_sync(0xffffffffu, thr, off));
  over = __any_sync(0xffffffffu, over);
  if (over) {                      // a slice ran out of room: the user is not certifiable from these lists
    if (lane == 0) {
      cand_count[(size_t)row * kHalves] = cap + 1;
      cand_count[(size_t)row * kHalves + 1] = 0;
      cand_thresh[row] = INFINITY;
    }
    return;
  }
  __syncwarp();
  const FlatLists fl{pre, lists};
  auto entry_of = [&](int f, uint4& q, uint32_t& col, bool want_col) {
    int l, i;
    fl.locate(f, l, i);
    const CandList in = sp.cand.at((base + (l >> 1)) * sp.cap + (l & 1) * half_cap);
    q = in.q[i];
    if (want_col) col = in.col[i];
  };
  // kth_sel-th largest stored group maximum (its lower bound: the stored 16 bits are a truncation): running
  // top 32 over all sub-lists, lane i = (i+1)-th largest.  Only values above `floor` = max(thr0, 32nd largest so
  // far) can matter, and they are rare (a lane or two per 32 entries), so they are first compacted into a small
  // shared-memory buffer and the 32-wide sorting network runs once per 32 SURVIVORS, not once per group column
  // of every step that holds one.
  float top = -INFINITY;
  float floor_v = thr;
  float* buf = merge_buf[wib];
  int nbuf = 0;
  auto flush = [&]() {
    for (int c = 0; c < nbuf; c += 32) {
      float v = c + lane < nbuf ? buf[c + lane] : -INFINITY;
#pragma unroll
      for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) cmpx_lane(v, lane, stride, (lane & size) == 0);
      }
      top = fmaxf(top, __shfl_sync(0xffffffffu, v, 31 - lane));
#pragma unroll
      for (int stride = 16; stride > 0; stride >>= 1) cmpx_lane(top, lane, stride, true);
    }
    nbuf = 0;
    floor_v = fmaxf(thr, __shfl_sync(0xffffffffu, top, 31));
    __syncwarp();
  };
  for (int f0 = 0; f0 < entries; f0 += 64) {
    uint4 qs[2];
    bool have[2];
    uint32_t unused = 0u;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int f = f0 + 32 * j + lane;
      have[j] = f < entries;
      qs[j] = make_uint4(0u, 0u, 0u, 0u);
      if (have[j]) entry_of(f, qs[j], unused, false);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      uint32_t kmask = 0u;
      if (have[j]) {
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          float lo, hi;
          cand_bounds(cand_half(qs[j], h), lo, hi);
          kmask |= lo > floor_v ? (1u << h) : 0u;
        }
      }
      if (!__any_sync(0xffffffffu, kmask != 0u)) continue;
      const int mine_cnt = __popc(kmask);
      int incl = mine_cnt;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, off);
        if (lane >= off) incl += o;
      }
      const int found = __shfl_sync(0xffffffffu, incl, 31);         // <= 256
      if (nbuf + found > kMergeBuf) flush();                         // (values at or below the new floor are harmless)
      int pos = nbuf + incl - mine_cnt;
      while (kmask) {
        const int h = __ffs(kmask) - 1;
        kmask &= kmask - 1;
        float lo, hi;
        cand_bounds(cand_half(qs[j], h), lo, hi);
        buf[pos++] = lo;
      }
      nbuf += found;
      __syncwarp();
    }
  }
  flush();
  // An unsliced row's tau is the kth_sel-th largest of 32 BUCKET maxima, which sits near the (kth_sel + 4)-th
  // best score because good items share buckets.  Taking the exact kth_sel-th best group here would leave less
  // room between the k-th score and the threshold, and 30x more sliced users failed their certificate.
  thr = fmaxf(thr, __shfl_sync(0xffffffffu, top, min(kth_sel + 4, 32) - 1));
  // gather the chunks that still hold a group above thr (judged by the upper end of its interval) into the
  // row's ordinary storage as ONE list: all `cap` entries belong to the first half, the second stays empty
  const CandList out = cand.at((size_t)row * cap);
  int total = 0;
  for (int f0 = 0; f0 < entries; f0 += 64) {
    uint4 qs[2];
    uint32_t cols[2];
    bool have[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int f = f0 + 32 * j + lane;
      have[j] = f < entries;
      qs[j] = make_uint4(0u, 0u, 0u, 0u);
      cols[j] = 0u;
      if (have[j]) entry_of(f, qs[j], cols[j], true);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      bool keep = false;
      if (have[j]) {
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          float lo, hi;
          cand_bounds(cand_half(qs[j], h), lo, hi);
          keep |= hi > thr;
        }
      }
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      const int pos = total + __popc(mask & ((1u << lane) - 1u));
      if (keep && pos < cap) {
        out.q[pos] = qs[j];
        out.col[pos] = cols[j];
      }
      total += __popc(mask);
    }
  }
  if (lane == 0) {
    cand_count[(size_t)row * kHalves] = total;       // > cap: overflow, flagged by hnm_rescore_topk
    cand_count[(size_t)row * kHalves + 1] = 0;
    cand_thresh[row] = total > cap ? INFINITY : thr;
  }
}

// ----------------------------------------------------------------------------- exclusion signatures
// One warp per user: word w of the signature is built by lane w.
__global__ void __launch_bounds__(128)
excl_signature_kernel(const int64_t* __restrict__ excl_ptr, const int64_t* __restrict__ excl_items, int64_t batch,
                      int64_t item_begin, int64_t num_items_local, uint32_t* __restrict__ sig) {
  static_assert(kSigWords == 32, "one word per lane");
  const int lane = threadIdx.x & 31;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= batch) return;
  uint32_t mine = 0u;
  const int64_t lo = excl_ptr[b], hi = excl_ptr[b + 1];
  for (int64_t base = lo; base < hi; base += 32) {
    int word = -1;
    uint32_t bit = 0u;
    if (base + lane < hi) {
      const int64_t local = excl_items[base + lane] - item_begin;
      if (local >= 0 && local < num_items_local) {
        const uint32_t chunk = (uint32_t)(local >> 5);
        word = (int)((chunk >> 5) & (kSigWords - 1));
        bit = 1u << (chunk & 31u);
      }
    }
#pragma unroll 8
    for (int j = 0; j < 32; ++j) {
      const int wj = __shfl_sync(0xffffffffu, word, j);
      const uint32_t bj = __shfl_sync(0xffffffffu, bit, j);
      if (wj == lane) mine |= bj;
    }
  }
  sig[(size_t)b * kSigWords + lane] = mine;
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

// Power of two s with absmax * s in [2^14, 2^15): fp16 keeps 11 significant bits and cannot overflow.
// Built from the exponent field, clamped to 2^+-100; 1 for a zero / non-finite absmax.
__device__ __forceinline__ float pow2_scale(float absmax) {
  const int e = (int)((__float_as_uint(absmax) >> 23) & 0xFFu);        // biased exponent; absmax >= 0
  if (e == 0 || e == 255) return 1.f;
  const int se = min(227, max(27, 268 - e));                            // 2^(14 - (e - 127)), biased
  return __uint_as_float((uint32_t)se << 23);
}

// D / 8 lanes per row (8 for d = 64, a whole warp for d = 256): a lane handles 8 consecutive floats -> one
// 16-byte store.
// PER_ROW = false (item shard): one scale for the table, derived from the device-resident absmax
//   (params[0]); the kernel also publishes params[1] = scale and params[2] = max_j ||x_j - c||^2.
// PER_ROW = true (users): every row gets its own power of two -- the ranking of the items for a fixed
//   user does not depend on that user's scale, so a table with a heavy-tailed norm distribution keeps
//   11 significant bits in every row; 1/scale goes to row_inv_scale[r] for hnm_rescore_topk.
template <bool PER_ROW, int D>
__global__ void pack_kernel(const float* __restrict__ emb, const int64_t* __restrict__ row_ids, int64_t num_rows,
                            int64_t rows_padded, const float* __restrict__ center, float* __restrict__ params,
                            __half* __restrict__ out, float* __restrict__ row_inv_scale) {
  constexpr int TPR = D / 8;                                             // lanes per row: 8, 16 or 32
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t r = t / TPR;
  const int sub = (int)(t % TPR);
  if (r >= rows_padded) return;                                          // whole lane groups leave together
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
  if (r < num_rows) {
    const int64_t src = row_ids ? row_ids[r] : r;
    const float* p = emb + (size_t)src * D + sub * 8;
    a = ldg_f4(p);
    b = ldg_f4(p + 4);
    if (center) {     // w = fl(x - c): ranking of u.x and u.(x - c) is the same for a fixed user
      const float4 ca = ldg_f4(center + sub * 8), cb = ldg_f4(center + sub * 8 + 4);
      a = make_float4(__fsub_rn(a.x, ca.x), __fsub_rn(a.y, ca.y), __fsub_rn(a.z, ca.z), __fsub_rn(a.w, ca.w));
      b = make_float4(__fsub_rn(b.x, cb.x), __fsub_rn(b.y, cb.y), __fsub_rn(b.z, cb.z), __fsub_rn(b.w, cb.w));
    }
  }
  // the lanes of a row are TPR consecutive lanes of one warp: xor-shuffles below TPR stay inside the row
  const unsigned gmask = TPR == 32 ? 0xffffffffu : (((1u << TPR) - 1u) << ((threadIdx.x & 31) & ~(TPR - 1)));
  float scale;
  if (PER_ROW) {
    float m = fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))),
                    fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w))));
#pragma unroll
    for (int off = 1; off < TPR; off <<= 1) m = fmaxf(m, __shfl_xor_sync(gmask, m, off));
    scale = pow2_scale(m);
    if (sub == 0 && r < num_rows) row_inv_scale[r] = 1.f / scale;       // exact: a power of two
  } else {
    scale = pow2_scale(params[0]);
    if (t == 0) params[1] = scale;
  }
  __half2 h[4];
  h[0] = __floats2half2_rn(a.x * scale, a.y * scale);
  h[1] = __floats2half2_rn(a.z * scale, a.w * scale);
  h[2] = __floats2half2_rn(b.x * scale, b.y * scale);
  h[3] = __floats2half2_rn(b.z * scale, b.w * scale);
  *reinterpret_cast<uint4*>(out + (size_t)r * D + sub * 8) = *reinterpret_cast<uint4*>(h);
  if (!PER_ROW) {
    float sq = a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w + b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
#pragma unroll
    for (int off = 1; off < TPR; off <<= 1) sq += __shfl_xor_sync(gmask, sq, off);
    if (sub == 0 && r < num_rows) atomicMax(reinterpret_cast<int*>(params + 2), __float_as_int(sq));   // sq >= 0
  }
}

// ----------------------------------------------------------------------------- rescoring
constexpr int kMaxPerLane = 32;  // candidate capacity handled = 32 * kMaxPerLane (HNM_FUSED_CAND_MAX)
constexpr int kGroup = 4;        // items per nominated group (select32)
constexpr int kMaxGroups = 64;   // surviving groups hnm_rescore_topk can take per user
constexpr int kMaxContenders = 128;  // rescored items above the cut it can rank per user
constexpr int kRescoreWarps = 4;     // one user per warp

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

// ----------------------------------------------------------------------------- rescoring, dim = 64
// The first version of this kernel ran the fp64 chain for every item of every kept group (~90 per
// user) and spent 57 % of its stall samples on F2F.F64.F32 (two fp32->fp64 conversions per DFMA at
// 16 conversions/clk/SM; profiles/r1_bench.md).  Now the work shrinks in three steps:
//   groups  g_i = the nominated groups' maxima (tensor-core scores of DISTINCT items, each within eps of the
//           exact score -- the bound the certificate rests on).  With g_(k) the k-th largest of them, k items
//           have an exact score >= T0 = g_(k)/(su si) - eps + u.c, so a group with g_i < g_(k) - 2 eps su si
//           cannot hold a top-k item and is dropped before anything is fetched.
//   pass 1  fp32: a_j = fl(u.x_j), rows staged by half warps into shared memory.
//           |a_j - u.x_j| <= gamma_64 * sum_k |u_k x_jk| <= gamma_64 ||u|| ||x_j||, so with
//           rad = 2^-17 ||u|| (max_j ||x_j - c|| + ||c||)   (2^-17 = 2 * 64 * 2^-24)
//           an item with a_j + rad < T0 has k items strictly above it; one with a_j + rad <= cut is not a
//           contender.  Both are dropped.
//   pass 2  fp64, survivors only (k plus the near ties, 16 per round): the half warp that stages a row
//           writes the PRODUCTS u_k * x_jk as doubles (exact: 24 + 24 significant bits), lane j then adds
//           them in the order k = 0..63.  round(p + acc) is what fma(u, x, acc) returns when p is exact,
//           so the scores are bit-identical to the fp64 fma chain of the oracle and of hnm_topk_exact.
// The certificate is unchanged: every item with an exact score above the cut among the top k survives,
// so "k contenders above the cut" holds for the same users as before.
constexpr int kRows2 = 16;                 // items per pass-2 round
constexpr int kStride1 = 68;               // floats per staged fp32 row (16-byte aligned, LDS.128 conflict free)
constexpr int kStride2 = 66;               // doubles per staged product row (16-byte aligned, LDS.128 conflict free)
constexpr int kMaxSel = 128;               // pass-1 survivors a user may have (more: not certified)
constexpr int kRescorePrefetch = 4096;     // users ahead whose inputs are pulled into the L2 (> users resident on the GPU)
static_assert(32 * kStride1 * 4 <= (kRows2 * kStride2 + 32) * 8, "tile union");

template <int D>
struct __align__(16) RescoreSmem {
  double tile[kRows2 * kStride2 + 32];     // pass 1: float [32][kStride1]; pass 2: double [16][kStride2] (one 64-wide slab)
  double sc[kMaxContenders];               // contender scores
  float uf[D];
  int id[kMaxContenders];
  uint32_t col[kMaxGroups];
  float glo[kMaxGroups];                   // interval of the group's tensor-core maximum
  float ghi[kMaxGroups];
  uint8_t sel[kMaxSel];
};

__device__ __forceinline__ void cmpx_desc(float& v, int lane, int stride, bool desc) {
  const float o = __shfl_xor_sync(0xffffffffu, v, stride);
  const bool lower = (lane & stride) == 0;
  // in a best-first block the lower lane keeps the larger value
  v = (lower == desc) ? fmaxf(v, o) : fminf(v, o);
}

template <int D>
__global__ void __launch_bounds__(kRescoreWarps * 32, 5)
rescore_dim_kernel(const float* __restrict__ user_emb, const float* __restrict__ item_emb,
                 const int64_t* __restrict__ user_ids, int64_t batch, int64_t item_begin, int num_items_local,
                 const CandList cand, int cap, const int32_t* __restrict__ cand_count,
                 const float* __restrict__ cand_thresh, const float* __restrict__ user_inv_scale,
                 const float* __restrict__ item_params,
                 const float* __restrict__ center, const int64_t* __restrict__ excl_ptr,
                 const int64_t* __restrict__ excl_items, int k, int64_t* __restrict__ out_ids,
                 double* __restrict__ out_scores, int32_t* __restrict__ certified) {
  constexpr int SL = D / kDim;              // 64-wide slabs of the embedding dimension
  __shared__ RescoreSmem<D> smem[kRescoreWarps];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int64_t b = (int64_t)blockIdx.x * kRescoreWarps + wib;
  if (b >= batch) return;
  RescoreSmem<D>& sm = smem[wib];
  float* tile1 = reinterpret_cast<float*>(sm.tile);
  const int64_t uid = user_ids ? user_ids[b] : b;
  const float* urow = user_emb + (size_t)uid * D;
  // the row's storage: cap entries; list 0 starts at entry 0, list 1 at cap / 2 (the two threads of the row in
  // the fused kernel); a lone list 0 (merge_split_kernel) may use all of it
  const int raw0 = cand_count[2 * b], raw1 = cand_count[2 * b + 1];
  const int cap0 = raw1 > 0 ? cap / 2 : cap;
  const bool list_overflow = raw0 > cap0 || raw1 > cap / 2;
  const int n0 = min(raw0, cap0), n1 = min(raw1, cap / 2);
  const float thr = cand_thresh[b];
  const CandList mine = cand.at((size_t)b * cap);
  int64_t ex_lo = 0, ex_hi = 0;
  if (excl_ptr) { ex_lo = excl_ptr[b]; ex_hi = excl_ptr[b + 1]; }
  const int half = lane >> 4, sub = lane & 15;
  // scales are powers of two (exact in any format): 1 / (su * si) with su per user, si per item shard;
  // item_params = {absmax, scale, max_j ||x_j - c||^2} as left on the device by hnm_absmax / hnm_score_pack_items
  const double inv_scale = (double)user_inv_scale[b] / (double)item_params[1];
  const double max_item_norm = sqrt((double)item_params[2]) * (1.0 + 1e-6);   // a hair above the fp32 sum
  {
    // The kernel is a chain of dependent loads per user (count -> list -> item rows), so warm the L2 for a user
    // that a later CTA will take: its list head (lane i: 32 bytes), count, threshold and embedding row.
    const int64_t pb = b + kRescorePrefetch;
    if (pb < batch) {
      if (4 * lane + 3 < cap) {      // 64 entries of each list: 64 bytes of maxima per lane, 256 bytes of columns
        const int pe = (lane < 16 ? 0 : cap / 2 - 64) + 4 * lane;
        const char* pa = reinterpret_cast<const char*>(cand.q + (size_t)pb * cap + pe);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + 32));
        if ((lane & 7) == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(cand.col + (size_t)pb * cap + pe)));
      }
      const char* pm = nullptr;
      if (lane == 0) pm = reinterpret_cast<const char*>(cand_count + 2 * pb);
      else if (lane == 1) pm = reinterpret_cast<const char*>(cand_thresh + pb);
      else if (lane < 10 && !user_ids) pm = reinterpret_cast<const char*>(user_emb + (size_t)pb * D) + (lane - 2) * 32;
      if (pm) asm volatile("prefetch.global.L2 [%0];" ::"l"(pm));
    }
  }

  // the user's row: fp32 copy in shared memory (pass 1 reads it as broadcasts, pass 2 turns it into doubles)
  // u.c in fp64 (it enters the cut); the three quantities that only feed error bounds as fp32 upper bounds
  double uc = 0.0;
  float un2 = 0.f, uc_absf = 0.f, cn2 = 0.f;
#pragma unroll
  for (int sl = 0; sl < SL; ++sl) {
    const float4 uf = ldg_f4(urow + sl * kDim + sub * 4);
    if (half == 0) {
      *reinterpret_cast<float4*>(sm.uf + sl * kDim + sub * 4) = uf;
      un2 += fmaf(uf.x, uf.x, fmaf(uf.y, uf.y, fmaf(uf.z, uf.z, uf.w * uf.w)));
      if (center) {
        const float4 cf = ldg_f4(center + sl * kDim + sub * 4);
        uc = fma((double)uf.x, (double)cf.x, uc); uc = fma((double)uf.y, (double)cf.y, uc);
        uc = fma((double)uf.z, (double)cf.z, uc); uc = fma((double)uf.w, (double)cf.w, uc);
        uc_absf += fabsf(uf.x * cf.x) + fabsf(uf.y * cf.y) + fabsf(uf.z * cf.z) + fabsf(uf.w * cf.w);
        cn2 += fmaf(cf.x, cf.x, fmaf(cf.y, cf.y, fmaf(cf.z, cf.z, cf.w * cf.w)));
      }
    }
  }

  // 1. keep the groups that may have ended above the final threshold (upper end of the interval the stored
  //    16 bits stand for), compacted: slot j = kept group j with the interval of its maximum
  int groups = 0;
  const uint32_t thr_key = order_key16(__float_as_uint(thr) >> 16);
  // one sweep over both lists: entry e < n0 comes from list 0, the others from list 1
  for (int e0 = 0; e0 < n0 + n1; e0 += 32) {
    const int e = e0 + lane;
    const int idx = e < n0 ? e : cap / 2 + (e - n0);
    const bool have = e < n0 + n1;
    uint4 q = make_uint4(0u, 0u, 0u, 0u);
    uint32_t col0 = 0u;
    if (have) {
      q = __ldg(mine.q + idx);
      col0 = __ldg(mine.col + idx);
    }
    // A group can exceed thr only if its 16-bit pattern is not below thr's own (truncation toward zero keeps
    // the order of the buckets), so the test is an integer compare on order keys; the float interval of a
    // group is decoded only for the ~20 of ~800 groups that pass.
    uint32_t kmask = 0u;
    if (have) {
#pragma unroll
      for (int h = 0; h < 8; ++h) kmask |= (order_key16(cand_half(q, h)) >= thr_key) ? (1u << h) : 0u;
    }
    // exclusive prefix sum of the per-lane counts
    const int mine_cnt = __popc(kmask);
    int incl = mine_cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off) incl += o;
    }
    int pos = groups + incl - mine_cnt;
    while (kmask) {
      const int h = __ffs(kmask) - 1;
      kmask &= kmask - 1;
      if (pos < kMaxGroups) {
        float lo, hi;
        cand_bounds(cand_half(q, h), lo, hi);
        sm.col[pos] = col0 + 4u * h;
        sm.glo[pos] = lo;
        sm.ghi[pos] = hi;
      }
      ++pos;
    }
    groups += __shfl_sync(0xffffffffu, incl, 31);
  }
  const bool too_many = groups > kMaxGroups;
  groups = min(groups, kMaxGroups);

#pragma unroll
  for (int off = 16; off; off >>= 1) {         // lanes 16..31 contribute zeros
    un2 += __shfl_xor_sync(0xffffffffu, un2, off);
    uc += __shfl_xor_sync(0xffffffffu, uc, off);
    uc_absf += __shfl_xor_sync(0xffffffffu, uc_absf, off);
    cn2 += __shfl_xor_sync(0xffffffffu, cn2, off);
  }
  // fp32 sums of D non-negative terms are within D * 2^-24 of the truth: 1.00002 makes them upper bounds
  const double un = (double)(sqrtf(un2) * 1.00002f);
  const double cn_norm = (double)(sqrtf(cn2) * 1.00002f);
  const double uc_abs = (double)(uc_absf * 1.00002f);
  // for any item j outside the kept groups:  u.x_j = u.(x_j - c) + u.c <= thr/(su*si) + eps + u.c =: cut
  const double eps = 1.1 * 0.0009765625 * un * max_item_norm + (double)D * 0.00390625 * inv_scale + 1e-12 * uc_abs;
  const double cut = (double)thr * inv_scale + eps + uc;
  const float cutf = __double2float_rd(cut);                // a_j + rad <= cutf  =>  u.x_j <= cut
  // fp32 dot product radius: 2 gamma_D ||u|| ||x_j|| (2 D 2^-24), ||x_j|| <= max ||x - c|| + ||c||; the absolute
  // term covers products that underflow in fp32
  const float rad = __double2float_ru(7.62939453125e-6 * SL * un * (max_item_norm + cn_norm) * 1.0001 + 1e-36);
  __syncwarp();

  // 2. g_(k) over (at most the first 32 of) the kept groups whose four items all exist and are not excluded
  float gv = -INFINITY;
  if (lane < groups) {
    const int c0 = (int)sm.col[lane];
    bool ok = c0 + kGroup <= num_items_local;               // a maximum that came from zero padding is no item
    if (ok && ex_lo < ex_hi) {
      for (int j = 0; j < kGroup; ++j) ok = ok && !in_sorted(excl_items, ex_lo, ex_hi, item_begin + (int64_t)(c0 + j));
    }
    if (ok) gv = sm.glo[lane];
  }
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) cmpx_desc(gv, lane, stride, (lane & size) == 0);
  }
  const float gk = __shfl_sync(0xffffffffu, gv, k - 1);      // -inf when fewer than k such groups
  // k items score >= T0 exactly; groups entirely below it are dropped (scaled units, rounded down)
  const double t0 = (double)gk * inv_scale - eps + uc;
  const float t0f = gk > -INFINITY ? __double2float_rd(t0) : -INFINITY;
  const float gmin = gk > -INFINITY ? __double2float_rd((double)gk - 2.0 * eps / inv_scale) : -INFINITY;
  {
    int kept = 0;
    for (int g0 = 0; g0 < groups; g0 += 32) {
      const int gi = g0 + lane;
      uint32_t c = 0u;
      bool keep = false;
      if (gi < groups) {
        c = sm.col[gi];
        keep = sm.ghi[gi] >= gmin;
      }
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      __syncwarp();
      if (keep) sm.col[kept + __popc(mask & ((1u << lane) - 1u))] = c;      // write position <= read position
      kept += __popc(mask);
      __syncwarp();
    }
    groups = kept;
  }

  // 3. pass 1: fp32 scores; survivors are compacted into sm.sel as they are found
  const int num_cand_items = groups * kGroup;
  int nsel = 0;
  for (int base = 0; base < num_cand_items; base += 32) {
    const int it = base + lane;
    int item = -1;
    if (it < num_cand_items) item = (int)sm.col[it / kGroup] + (it % kGroup);
    bool live = item >= 0 && item < num_items_local;        // columns past the catalog are zero padding
    if (live && ex_lo < ex_hi && in_sorted(excl_items, ex_lo, ex_hi, item_begin + (int64_t)item)) live = false;
    const int litem = live ? item : 0;                      // dead slots fetch row 0; their result is not used
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;          // four partial sums: any order satisfies the bound
#pragma unroll 1
    for (int sl = 0; sl < SL; ++sl) {                       // one 64-wide slab of the rows at a time
#pragma unroll
      for (int h0 = 0; h0 < 16; h0 += 8) {
        float4 v[8];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
          const int il = __shfl_sync(0xffffffffu, litem, 2 * (h0 + s) + half);
          v[s] = ldg_f4(item_emb + (size_t)il * D + sl * kDim + sub * 4);
        }
#pragma unroll
        for (int s = 0; s < 8; ++s)
          *reinterpret_cast<float4*>(tile1 + (2 * (h0 + s) + half) * kStride1 + sub * 4) = v[s];
      }
      __syncwarp();
      const float* t = tile1 + lane * kStride1;
#pragma unroll
      for (int kk = 0; kk < kDim; kk += 4) {
        const float4 u4 = *reinterpret_cast<const float4*>(sm.uf + sl * kDim + kk);
        const float4 x4 = *reinterpret_cast<const float4*>(t + kk);
        a0 = fmaf(u4.x, x4.x, a0);
        a1 = fmaf(u4.y, x4.y, a1);
        a2 = fmaf(u4.z, x4.z, a2);
        a3 = fmaf(u4.w, x4.w, a3);
      }
      if (SL > 1) __syncwarp();                             // the tile is rewritten by the next slab
    }
    const float up = ((a0 + a1) + (a2 + a3)) + rad;
    const bool s = live && up >= t0f && up > cutf;
    const unsigned mask = __ballot_sync(0xffffffffu, s);
    const int pos = nsel + __popc(mask & ((1u << lane) - 1u));
    if (s && pos < kMaxSel) sm.sel[pos] = (uint8_t)it;
    nsel += __popc(mask);
    __syncwarp();
  }
  const bool sel_overflow = nsel > kMaxSel;
  if (sel_overflow) nsel = 0;
  __syncwarp();

  // 4. pass 2: exact scores of the survivors; contenders = strictly above the cut
  int total = 0;
  for (int base = 0; base < nsel; base += kRows2) {
    double acc = 0.0;
    int item = -1;
    const bool mine2 = lane < kRows2 && base + lane < nsel;
    if (mine2) {
      const int it = sm.sel[base + lane];
      item = (int)sm.col[it / kGroup] + (it % kGroup);
    }
#pragma unroll 1
    for (int sl = 0; sl < SL; ++sl) {                          // slabs in order: the chain runs k = 0 .. D-1
      {
        // all eight row loads of this half warp are issued before the first one is used
        // (slots past the last survivor repeat it: no predicates, and nobody reads those rows)
        float4 v[kRows2 / 2];
#pragma unroll
        for (int s = 0; s < kRows2 / 2; ++s) {
          const int it = sm.sel[min(base + 2 * s + half, nsel - 1)];
          const int itm = (int)sm.col[it / kGroup] + (it % kGroup);
          v[s] = ldg_f4(item_emb + (size_t)itm * D + sl * kDim + sub * 4);
        }
        const float4 uf = *reinterpret_cast<const float4*>(sm.uf + sl * kDim + sub * 4);
        const double ud0 = (double)uf.x, ud1 = (double)uf.y, ud2 = (double)uf.z, ud3 = (double)uf.w;
#pragma unroll
        for (int s = 0; s < kRows2 / 2; ++s) {
          double2* t = reinterpret_cast<double2*>(sm.tile + (2 * s + half) * kStride2 + sub * 4);
          t[0] = make_double2(ud0 * (double)v[s].x, ud1 * (double)v[s].y);
          t[1] = make_double2(ud2 * (double)v[s].z, ud3 * (double)v[s].w);
        }
      }
      __syncwarp();
      if (mine2) {
        const double2* t = reinterpret_cast<const double2*>(sm.tile + lane * kStride2);
#pragma unroll 8
        for (int kk = 0; kk < kDim / 2; ++kk) {        // == fma(u_k, x_k, acc) in the order of k: the products are exact
          const double2 pr = t[kk];
          acc = __dadd_rn(pr.x, acc);
          acc = __dadd_rn(pr.y, acc);
        }
      }
      if (SL > 1) __syncwarp();                                // the tile is rewritten by the next slab
    }
    const bool c = mine2 && acc > cut;
    const unsigned mask = __ballot_sync(0xffffffffu, c);
    const int pos = total + __popc(mask & ((1u << lane) - 1u));
    if (c) { sm.sc[pos] = acc; sm.id[pos] = item; }           // pos < nsel <= kMaxSel = kMaxContenders
    total += __popc(mask);
    __syncwarp();
  }
  if (total <= 32) {
    // the rule: one contender per lane; its rank = how many contenders come before it (the (score, id) pairs
    // are distinct), `total` broadcast rounds instead of a 15-step sorting network on doubles
    const double ms = lane < total ? sm.sc[lane] : -INFINITY;
    const int mi = lane < total ? sm.id[lane] : INT32_MAX;
    int rank = 0;
    for (int r = 0; r < total; ++r) {
      const double os = __shfl_sync(0xffffffffu, ms, r);
      const int oi = __shfl_sync(0xffffffffu, mi, r);
      rank += before32(os, oi, ms, mi) ? 1 : 0;
    }
    if (lane < total) {
      if (rank < k) {
        out_ids[(size_t)b * k + rank] = item_begin + (int64_t)mi;
        out_scores[(size_t)b * k + rank] = ms;
      }
    } else if (lane < k) {                                    // fewer than k contenders: the user is not certified
      out_ids[(size_t)b * k + lane] = INT64_MAX;
      out_scores[(size_t)b * k + lane] = -INFINITY;
    }
  } else {
    // the exception (many near ties around the k-th score): k rounds of warp argmax over the list
    for (int t = 0; t < k; ++t) {
      double bs = -INFINITY;
      int bi = INT32_MAX, bp = -1;
      for (int p = lane; p < total; p += 32) {
        const double x = sm.sc[p];
        const int xi = sm.id[p];
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
      if (bp >= 0 && wi == bi && ws == bs) { sm.sc[bp] = -INFINITY; sm.id[bp] = INT32_MAX; }
      __syncwarp();
      if (lane == 0) {
        out_ids[(size_t)b * k + t] = item_begin + (int64_t)wi;
        out_scores[(size_t)b * k + t] = ws;
      }
    }
  }
  if (lane == 0) {
    // bit 0: provably exact; bits 1.. say why not (list overflow, > 64 groups, < k contenders, > 128 survivors)
    const int why = (list_overflow ? 2 : 0) | (too_many ? 4 : 0) | (total < k ? 8 : 0) | (sel_overflow ? 16 : 0);
    certified[b] = why == 0 ? 1 : why;
  }
}

int make_map(CUtensorMap* map, const void* base, int64_t rows, int dim) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) return HNM_E_DRIVER;
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};             // box = one 64-wide K chunk of 128 rows
  cuuint64_t gstride[1] = {(cuuint64_t)(dim * 2)};
  cuuint32_t box[2] = {(cuuint32_t)kDim, (cuuint32_t)kItemTile};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? HNM_OK : HNM_E_DRIVER;
}


// Column means of a [rows, D] table in two deterministic steps (fp64 sums: block partials over contiguous row
// ranges, then the partials in block order) -- the centre of the item shard for hnm_score_pack_items.
template <int D>
__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, int64_t rows, double* __restrict__ partial) {
  constexpr int C4 = D / 4, RL = 256 / C4;
  __shared__ double sm[RL][D];
  const int c = threadIdx.x % C4, r = threadIdx.x / C4;
  const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t b = (int64_t)blockIdx.x * per, e = min(rows, b + per);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (int64_t i = b + r; i < e; i += RL) {
    const float4 v = ldg_f4(x + (size_t)i * D + c * 4);
    a0 += (double)v.x; a1 += (double)v.y; a2 += (double)v.z; a3 += (double)v.w;
  }
  sm[r][c * 4 + 0] = a0; sm[r][c * 4 + 1] = a1; sm[r][c * 4 + 2] = a2; sm[r][c * 4 + 3] = a3;
  __syncthreads();
  for (int t = threadIdx.x; t < D; t += 256) {
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < RL; ++j) acc += sm[j][t];
    partial[(size_t)blockIdx.x * D + t] = acc;
  }
}

// One CTA of 1024 threads: thread (j, t) adds the partials of the blocks b = j, j + J, ... of column t, then
// the J sub-sums are added in the order of j -- a fixed order, so the centre is identical from run to run.
__global__ void __launch_bounds__(1024)
colmean_finish_kernel(const double* __restrict__ partial, int blocks, int dim, int64_t rows,
                      float* __restrict__ out) {
  __shared__ double sm[1024];
  const int t = threadIdx.x % dim, j = threadIdx.x / dim, J = 1024 / dim;
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int b = j;
  for (; b + 3 * J < blocks; b += 4 * J) {            // four loads in flight
    const double v0 = partial[(size_t)b * dim + t], v1 = partial[(size_t)(b + J) * dim + t];
    const double v2 = partial[(size_t)(b + 2 * J) * dim + t], v3 = partial[(size_t)(b + 3 * J) * dim + t];
    a0 += v0; a1 += v1; a2 += v2; a3 += v3;
  }
  for (; b < blocks; b += J) a0 += partial[(size_t)b * dim + t];
  sm[threadIdx.x] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (j == 0) {
    double acc = 0.0;
    for (int jj = 0; jj < J; ++jj) acc += sm[jj * dim + t];
    out[t] = (float)(acc / (double)rows);
  }
}
}  // namespace

extern "C" int64_t hnm_column_mean_workspace_bytes(int32_t dim) {
  if (dim <= 0) return -1;
  return (int64_t)hnm_num_sms() * 4 * dim * (int64_t)sizeof(double);
}

extern "C" int hnm_column_mean(const float* emb, int64_t rows, int32_t dim, float* out_mean, void* workspace,
                               int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!emb || !out_mean || !workspace) return HNM_E_NULL;
  if (dim != 64 && dim != 128 && dim != 256) return HNM_E_DIM;
  if (rows <= 0) return HNM_E_RANGE;
  if (!hnm_aligned16(emb) || (reinterpret_cast<uintptr_t>(workspace) & 7)) return HNM_E_ALIGN;
  if (workspace_bytes < hnm_column_mean_workspace_bytes(dim)) return HNM_E_WORKSPACE;
  const int blocks = (int)std::min<int64_t>((int64_t)hnm_num_sms() * 4, (rows + 63) / 64);
  double* partial = reinterpret_cast<double*>(workspace);
  if (dim == 64) colsum_kernel<64><<<blocks, 256, 0, stream>>>(emb, rows, partial);
  else if (dim == 128) colsum_kernel<128><<<blocks, 256, 0, stream>>>(emb, rows, partial);
  else colsum_kernel<256><<<blocks, 256, 0, stream>>>(emb, rows, partial);
  HNM_LAUNCH_CHECK();
  colmean_finish_kernel<<<1, 1024, 0, stream>>>(partial, blocks, dim, rows, out_mean);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

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

extern "C" int hnm_score_pack_items(const float* emb, int64_t num_rows, int64_t rows_padded, int32_t dim,
                                    const float* center, float* params, void* out_f16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!emb || !out_f16 || !params) return HNM_E_NULL;
  if (dim != 64 && dim != 128 && dim != 256) return HNM_E_DIM;
  if (num_rows < 0 || rows_padded < num_rows || rows_padded <= 0) return HNM_E_RANGE;
  if (!hnm_aligned16(emb) || !hnm_aligned16(out_f16) || (center && !hnm_aligned16(center))) return HNM_E_ALIGN;
  const int T = 256;
  const int64_t threads = rows_padded * (dim / 8);
  const unsigned grid = (unsigned)((threads + T - 1) / T);
  if (dim == 64) pack_kernel<false, 64><<<grid, T, 0, stream>>>(emb, nullptr, num_rows, rows_padded, center, params, (__half*)out_f16, nullptr);
  else if (dim == 128) pack_kernel<false, 128><<<grid, T, 0, stream>>>(emb, nullptr, num_rows, rows_padded, center, params, (__half*)out_f16, nullptr);
  else pack_kernel<false, 256><<<grid, T, 0, stream>>>(emb, nullptr, num_rows, rows_padded, center, params, (__half*)out_f16, nullptr);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_score_pack_users(const float* emb, const int64_t* row_ids, int64_t num_rows, int64_t rows_padded,
                                    int32_t dim, void* out_f16, float* out_inv_scale, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!emb || !out_f16 || !out_inv_scale) return HNM_E_NULL;
  if (dim != 64 && dim != 128 && dim != 256) return HNM_E_DIM;
  if (num_rows < 0 || rows_padded < num_rows || rows_padded <= 0) return HNM_E_RANGE;
  if (!hnm_aligned16(emb) || !hnm_aligned16(out_f16)) return HNM_E_ALIGN;
  const int T = 256;
  const int64_t threads = rows_padded * (dim / 8);
  const unsigned grid = (unsigned)((threads + T - 1) / T);
  if (dim == 64) pack_kernel<true, 64><<<grid, T, 0, stream>>>(emb, row_ids, num_rows, rows_padded, nullptr, nullptr, (__half*)out_f16, out_inv_scale);
  else if (dim == 128) pack_kernel<true, 128><<<grid, T, 0, stream>>>(emb, row_ids, num_rows, rows_padded, nullptr, nullptr, (__half*)out_f16, out_inv_scale);
  else pack_kernel<true, 256><<<grid, T, 0, stream>>>(emb, row_ids, num_rows, rows_padded, nullptr, nullptr, (__half*)out_f16, out_inv_scale);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

namespace {
constexpr int kSplitCap = 256;      // candidate entries (32-column chunks) per (sliced user, item slice)
constexpr int kMinSliceTiles = 32;  // an item slice is at least this many item tiles

int fused_mu() { return kMU; }

// Seed tiles: the number of chunks a row nominates grows like (k + 3) ln(#item tiles / #seed tiles), so the seed
// scales with the catalog: 16 tiles up to ~1 000 item tiles (the H&M catalog), then 1/48 of the catalog in steps
// of 16 (160 at 1 M items, where 16 seed tiles left 3 % of the users with an overflowing list:
// profiles/r2_sweep_n8_first.json); 2 % more MMA work.
int seed_tiles(int num_item_tiles) {
  static const int boot_env = getenv("HNM_FUSED_BOOT") ? std::max(1, atoi(getenv("HNM_FUSED_BOOT"))) : 0;
  return boot_env ? boot_env : std::max(kBootTiles, num_item_tiles / 48 / 16 * 16);
}

// The work distribution of one launch (see SplitPlan); pointers are filled in by the caller.
SplitPlan make_plan(int num_user_tiles, int num_item_tiles, int grid) {
  const int kMU = fused_mu();
  SplitPlan sp{};
  sp.mu = kMU;
  static const bool no_split = getenv("HNM_FUSED_NOSPLIT") != nullptr;      // A/B switch for profiling
  sp.full_passes = num_user_tiles / (grid * kMU);
  sp.tile0 = grid * kMU * sp.full_passes;
  const int left = num_user_tiles - sp.tile0;
  sp.triples = (left + kMU - 1) / kMU;
  sp.slices = 0;
  if (sp.triples > 0) {
    // slices per left-over pair: the count that minimises the length of the tail, rounds x (tiles + seed tiles of
    // a slice), over 1 .. min(grid, tiles / 32).  171 498 users (the H&M shape on 8 GPUs) leave 78 pairs for 148
    // CTAs: unsliced that is a whole extra pass on half of the GPU, three slices each run in two rounds of a
    // fraction of a pass.
    const int boot = seed_tiles(num_item_tiles);
    const int max_slices = std::max(1, std::min(grid, num_item_tiles / kMinSliceTiles));
    long long best = -1;
    for (int sl = 1; sl <= max_slices; ++sl) {
      const int ni = (num_item_tiles + sl - 1) / sl;
      const int b = sl > 1 ? std::max(1, std::min(boot, ni / 4)) : std::min(ni, boot);
      const long long rounds = ((long long)sp.triples * sl + grid - 1) / grid;
      const long long cost = rounds * (ni + b) * (100 + sl);       // 1 % per slice: near-ties go to fewer slices (less to merge)
      if (best < 0 || cost < best) { best = cost; sp.slices = sl; }
    }
    if (no_split) sp.slices = 1;
  }
  sp.cap = kSplitCap;
  return sp;
}

int fused_grid(int num_user_tiles) {
  const int kMU = fused_mu();
  static const bool no_split = getenv("HNM_FUSED_NOSPLIT") != nullptr;
  if (no_split) return std::min((num_user_tiles + kMU - 1) / kMU, hnm_num_sms());
  return hnm_num_sms();
}

size_t split_bytes(const SplitPlan& sp, size_t* off_count, size_t* off_thresh) {
  if (sp.slices <= 1) { *off_count = *off_thresh = 0; return 0; }
  const size_t slots = (size_t)sp.triples * sp.mu * kUserTile * sp.slices;
  *off_count = slots * kSplitCap * (sizeof(uint4) + sizeof(uint32_t));
  *off_thresh = *off_count + slots * kHalves * sizeof(int32_t);
  return *off_thresh + slots * sizeof(float);
}
}  // namespace

extern "C" int64_t hnm_score_topk_fused_workspace_bytes(int64_t users_padded, int64_t items_padded) {
  if (users_padded <= 0 || items_padded <= 0 || users_padded % kUserTile || items_padded % kItemTile ||
      users_padded > INT32_MAX || items_padded > INT32_MAX)
    return -1;
  const int tiles = (int)(users_padded / kUserTile);
  size_t a, b;
  return (int64_t)split_bytes(make_plan(tiles, (int)(items_padded / kItemTile), fused_grid(tiles)), &a, &b) + 256;
}

extern "C" int hnm_score_topk_fused_plan(int64_t users_padded, int64_t items_padded, int32_t* out6) {
  if (!out6) return HNM_E_NULL;
  if (users_padded <= 0 || items_padded <= 0 || users_padded % kUserTile || items_padded % kItemTile ||
      users_padded > INT32_MAX || items_padded > INT32_MAX)
    return HNM_E_RANGE;
  const int tiles = (int)(users_padded / kUserTile);
  const int grid = fused_grid(tiles);
  const SplitPlan sp = make_plan(tiles, (int)(items_padded / kItemTile), grid);
  out6[0] = grid; out6[1] = sp.full_passes; out6[2] = sp.tile0; out6[3] = sp.triples; out6[4] = sp.slices;
  out6[5] = sp.mu;
  return HNM_OK;
}

extern "C" int hnm_score_topk_fused(const void* users_f16, int64_t num_users, int64_t users_padded,
                                    const void* items_f16, int64_t num_items, int64_t items_padded, int32_t dim,
                                    int32_t kth_sel,
                                    void* cand, int32_t cand_cap, int32_t* cand_count, float* cand_thresh,
                                    const uint32_t* excl_sig, void* workspace, int64_t workspace_bytes,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!users_f16 || !items_f16 || !cand || !cand_count || !cand_thresh) return HNM_E_NULL;
  if (num_users <= 0 || num_items <= 0 || users_padded < num_users || items_padded < num_items) return HNM_E_RANGE;
  if (users_padded % kUserTile != 0 || items_padded % kItemTile != 0) return HNM_E_RANGE;
  if (users_padded > INT32_MAX || items_padded > INT32_MAX) return HNM_E_RANGE;
  if (kth_sel < 1 || kth_sel > kNumBuckets || cand_cap < 2 || cand_cap % 2 || cand_cap > 32 * kMaxPerLane) return HNM_E_RANGE;
  if ((reinterpret_cast<uintptr_t>(users_f16) & 127) || (reinterpret_cast<uintptr_t>(items_f16) & 127)) return HNM_E_ALIGN;
  int rc = hnm_check_device();
  if (rc != HNM_OK) return rc;
  CUtensorMap map_u, map_i;
  if (dim != 64 && dim != 128 && dim != 256) return HNM_E_DIM;
  if ((rc = make_map(&map_u, users_f16, users_padded, dim)) != HNM_OK) return rc;
  if ((rc = make_map(&map_i, items_f16, items_padded, dim)) != HNM_OK) return rc;
  static const int debug_mode = getenv("HNM_FUSED_DEBUG") ? atoi(getenv("HNM_FUSED_DEBUG")) : 0;
  static const int refresh_div = getenv("HNM_FUSED_REFRESH") ? std::max(1, atoi(getenv("HNM_FUSED_REFRESH"))) : 4;
  static const uint32_t wait_hint = getenv("HNM_FUSED_WAIT_NS") ? (uint32_t)atoi(getenv("HNM_FUSED_WAIT_NS")) : kWaitHintNs;
  const int num_user_tiles = (int)(users_padded / kUserTile);
  const int num_tiles = (int)(items_padded / kItemTile);
  const int boot_tiles = seed_tiles(num_tiles);
  const int grid = fused_grid(num_user_tiles);
  SplitPlan sp = make_plan(num_user_tiles, num_tiles, grid);
  size_t off_count = 0, off_thresh = 0;
  const size_t need = split_bytes(sp, &off_count, &off_thresh);
  if (need > 0) {
    if (!workspace) return HNM_E_NULL;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
    if ((int64_t)(need + (size_t)(ws - reinterpret_cast<uint8_t*>(workspace))) > workspace_bytes) return HNM_E_WORKSPACE;
    sp.cand = cand_list(ws, (size_t)sp.triples * sp.mu * kUserTile * sp.slices, kSplitCap);
    sp.count = reinterpret_cast<int32_t*>(ws + off_count);
    sp.thresh = reinterpret_cast<float*>(ws + off_thresh);
  }
#define HNM_FUSED_LAUNCH(KC)                                                                                       \
  {                                                                                                                \
    HNM_CUDA_TRY(hnm_allow_smem(score_topk_fused_kernel<KC>, (int)smem_bytes<KC>()));                              \
    score_topk_fused_kernel<KC><<<grid, kThreads, smem_bytes<KC>(), stream>>>(                                     \
        map_u, map_i, (int)num_users, num_user_tiles, num_tiles, kth_sel, cand_list(cand, (size_t)num_users, cand_cap), \
        cand_cap, cand_count, cand_thresh, excl_sig, debug_mode, boot_tiles, refresh_div, wait_hint, sp);          \
  }
  if (dim == 64) HNM_FUSED_LAUNCH(1)
  else if (dim == 128) HNM_FUSED_LAUNCH(2)
  else HNM_FUSED_LAUNCH(4)
#undef HNM_FUSED_LAUNCH
  HNM_LAUNCH_CHECK();
  if (need > 0) {
    const int64_t split_users = std::min<int64_t>((int64_t)sp.triples * sp.mu * kUserTile,
                                                  num_users - (int64_t)sp.tile0 * kUserTile);
    if (split_users > 0) {
      merge_split_kernel<<<(unsigned)((split_users + 3) / 4), 128, 4 * (sp.slices * kHalves + 1) * sizeof(int32_t), stream>>>(sp, (int)num_users, kth_sel, cand_list(cand, (size_t)num_users, cand_cap), cand_cap,
                                                                                cand_count, cand_thresh);
      HNM_LAUNCH_CHECK();
    }
  }
  return HNM_OK;
}

extern "C" int hnm_exclusion_signature(const int64_t* excl_ptr, const int64_t* excl_items, int64_t batch,
                                       int64_t item_begin, int64_t num_items_local, uint32_t* out_sig,
                                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (batch == 0) return HNM_OK;
  if (!excl_ptr || !excl_items || !out_sig) return HNM_E_NULL;
  if (batch < 0 || num_items_local < 1) return HNM_E_RANGE;
  excl_signature_kernel<<<(unsigned)((batch + 3) / 4), 128, 0, stream>>>(excl_ptr, excl_items, batch, item_begin,
                                                                        num_items_local, out_sig);
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}

extern "C" int hnm_rescore_topk(const float* user_emb, const float* item_emb, const int64_t* user_ids, int64_t batch,
                                int32_t dim, int64_t item_begin, int64_t num_items_local, const void* cand,
                                int32_t cand_cap, const int32_t* cand_count, const float* cand_thresh,
                                const float* user_inv_scale, const float* item_params, const float* center,
                                const int64_t* excl_ptr, const int64_t* excl_items, int32_t k, int64_t* out_ids,
                                double* out_scores, int32_t* out_certified, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (batch == 0) return HNM_OK;
  if (!user_emb || !item_emb || !cand || !cand_count || !cand_thresh || !user_inv_scale || !item_params || !out_ids ||
      !out_scores || !out_certified)
    return HNM_E_NULL;
  if ((excl_ptr != nullptr) != (excl_items != nullptr)) return HNM_E_NULL;
  if (dim != 64 && dim != 128 && dim != 256) return HNM_E_DIM;
  if (batch < 0 || k < 1 || k > 32 || cand_cap < 2 || cand_cap % 2 || cand_cap > 32 * kMaxPerLane || num_items_local < 1 ||
      num_items_local > INT32_MAX)
    return HNM_E_RANGE;
  if (!hnm_aligned16(user_emb) || !hnm_aligned16(item_emb) || !hnm_aligned16(cand) || (center && !hnm_aligned16(center)))
    return HNM_E_ALIGN;
  const int wpc = kRescoreWarps;
  const unsigned grid = (unsigned)((batch + wpc - 1) / wpc);
  const CandList cl = cand_list(const_cast<void*>(cand), (size_t)batch, cand_cap);
#define HNM_RESCORE(DD)                                                                                          \
  rescore_dim_kernel<DD><<<grid, wpc * 32, 0, stream>>>(user_emb, item_emb, user_ids, batch, item_begin,         \
                                                        (int)num_items_local, cl, cand_cap, cand_count, cand_thresh, \
                                                        user_inv_scale, item_params, center, excl_ptr, excl_items, k, \
                                                        out_ids, out_scores, out_certified)
  if (dim == 64) HNM_RESCORE(64);
  else if (dim == 128) HNM_RESCORE(128);
  else HNM_RESCORE(256);
#undef HNM_RESCORE
  HNM_LAUNCH_CHECK();
  return HNM_OK;
}
