// Micro-benchmark: the ceiling of random 256-byte row gathers on this GPU (what one LightGCN layer is
// made of: 65 M gathers of a 64-float embedding row per layer at the H&M shape, csrc/spmm.cu).
//
//   table 27 MB  (105 542 rows: the item block, L2 resident)   -> L2 -> SM ceiling
//   table 351 MB (1 371 980 rows: the user block, 2.8x the L2) -> DRAM-miss ceiling
//
// Two access paths, both at full occupancy with as much memory-level parallelism as the path allows:
//   ldg      half a warp per row (16 lanes x ld.global.nc.v4), 8 rows in flight per lane  (the SpMM's path)
//   gather4  cp.async.bulk.tensor.2d...tile::gather4 (4 rows per instruction by index) into a shared-memory
//            ring of 32 KB stages, one producer warp + consumer warps that sum the rows out of shared memory
// Reported: TB/s of gathered rows (rows x 256 B / time).  Indices are uniform random (every row equally
// likely), generated on the device; the sum of everything gathered is written out so nothing is elided.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_l2_gather bench_l2_gather.cu -lcuda
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kDim = 64;

__global__ void fill_table(float* t, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    t[i] = (float)(i & 1023) * 1e-3f;
}
// Zipf-like popularity as in hnm_recommendation_b200/synth.py: P(rank r) ~ 1 / (r + 100); the ranks are scattered
// over the table by a multiplicative hash so that popular rows are not neighbours.  Inverse CDF of the continuous
// approximation: r = (R + 100)^u * 100^(1-u) - 100.
__global__ void fill_idx_zipf(int* idx, size_t n, uint32_t rows, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const double u = (double)(z >> 11) * (1.0 / 9007199254740992.0);
    double r = pow((double)rows + 100.0, u) * pow(100.0, 1.0 - u) - 100.0;
    uint64_t rank = (uint64_t)fmin(fmax(r, 0.0), (double)rows - 1.0);
    idx[i] = (int)((rank * 2654435761ull) % rows);
  }
}

__global__ void fill_idx(int* idx, size_t n, uint32_t rows, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint64_t z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    idx[i] = (int)(z % rows);
  }
}

// ------------------------------------------------------------------------------------------- ldg path
template <int INFLIGHT>
__global__ void __launch_bounds__(256) gather_ldg(const float* __restrict__ table, const int* __restrict__ idx,
                                                  size_t n, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, half = lane >> 4, sub = lane & 15;
  const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  // a warp step = 32 indices (one coalesced load), consumed 2 rows (one per half warp) x INFLIGHT at a time
  for (size_t base = warp * 32; base + 32 <= n; base += warps * 32) {
    const int mine = idx[base + lane];
#pragma unroll
    for (int r0 = 0; r0 < 32; r0 += 2 * INFLIGHT) {
      float4 v[INFLIGHT];
#pragma unroll
      for (int j = 0; j < INFLIGHT; ++j) {
        const int row = __shfl_sync(0xffffffffu, mine, r0 + 2 * j + half);
        v[j] = __ldg(reinterpret_cast<const float4*>(table + (size_t)row * kDim) + sub);
      }
#pragma unroll
      for (int j = 0; j < INFLIGHT; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    }
  }
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

// ------------------------------------------------------------------------------------------- gather4 path
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
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, uint64_t* bar, int4 rows) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(0), "r"(rows.x), "r"(rows.y), "r"(rows.z), "r"(rows.w), "r"(smem_u32(bar))
      : "memory");
}

constexpr int kStageRows = 128;                    // 32 gather4 per stage, one per lane of the producer warp
constexpr int kStageBytes = kStageRows * kDim * 4; // 32 KB
template <int STAGES, int CONSUMERS>
__global__ void __launch_bounds__(32 * (1 + CONSUMERS), 1)
gather_tma(const __grid_constant__ CUtensorMap map, const int* __restrict__ idx, size_t n, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t stages_total = n / kStageRows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (warp == 0) {
    uint32_t g = 0;
    for (size_t st = blockIdx.x; st < stages_total; st += gridDim.x, ++g) {
      const int s = g % STAGES;
      mbar_wait(&empty[s], ((g / STAGES) & 1) ^ 1);
      const int4 rows = __ldg(reinterpret_cast<const int4*>(idx + st * kStageRows) + lane);
      if (lane == 0) mbar_expect_tx(&full[s], kStageBytes);
      __syncwarp();
      tma_gather4(smem + s * kStageBytes + lane * 1024, &map, &full[s], rows);
    }
  } else {
    uint32_t g = 0;
    const int c = warp - 1;
    for (size_t st = blockIdx.x; st < stages_total; st += gridDim.x, ++g) {
      const int s = g % STAGES;
      mbar_wait(&full[s], (g / STAGES) & 1);
      const float4* rows = reinterpret_cast<const float4*>(smem + s * kStageBytes);
      // consumer c sums rows c, c + CONSUMERS, ...: half a warp per row
#pragma unroll 4
      for (int r = 2 * c + (lane >> 4); r < kStageRows; r += 2 * CONSUMERS) {
        const float4 v = rows[r * 16 + (lane & 15)];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
  }
  out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

static int make_map(CUtensorMap* map, const float* base, uint64_t rows) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return 1;
  auto encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  cuuint64_t gdim[2] = {(cuuint64_t)kDim, rows};
  cuuint64_t gstride[1] = {(cuuint64_t)(kDim * 4)};
  cuuint32_t box[2] = {(cuuint32_t)kDim, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 2;
}

template <typename F>
static float time_ms(F&& launch, int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  launch();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  for (int i = 0; i < reps; ++i) launch();
  CK(cudaEventRecord(b));
  CK(cudaDeviceSynchronize());
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms / reps;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, L2 %.0f MB\n", prop.name, sms, prop.l2CacheSize / 1048576.0);
  const size_t n = (size_t)1 << 25;                 // 33.5 M gathers = 8.6 GB of rows per launch
  int* idx;
  float* out;
  CK(cudaMalloc(&idx, n * sizeof(int)));
  CK(cudaMalloc(&out, (size_t)sms * 64 * 1024 * sizeof(float)));
  for (uint32_t rows : {105542u, 1371980u}) {
    float* table;
    CK(cudaMalloc(&table, (size_t)rows * kDim * 4));
    fill_table<<<sms * 8, 256>>>(table, (size_t)rows * kDim);
    fill_idx<<<sms * 8, 256>>>(idx, n, rows, 12345u);
    CK(cudaDeviceSynchronize());
    const double gb = (double)n * kDim * 4 / 1e9;
    printf("table %u rows = %.0f MB, %zu gathers of 256 B = %.2f GB per launch\n", rows, rows * 256.0 / 1e6, n, gb);
    for (int ctas_per_sm : {4, 8}) {
      float ms = time_ms([&] { gather_ldg<8><<<sms * ctas_per_sm, 256>>>(table, idx, n, out); }, 5);
      printf("  ldg      8 rows in flight per lane, %d CTAs x 256 thr per SM : %.3f ms  %.2f TB/s\n", ctas_per_sm, ms, gb / ms);
      ms = time_ms([&] { gather_ldg<16><<<sms * ctas_per_sm, 256>>>(table, idx, n, out); }, 5);
      printf("  ldg     16 rows in flight per lane, %d CTAs x 256 thr per SM : %.3f ms  %.2f TB/s\n", ctas_per_sm, ms, gb / ms);
    }
    CUtensorMap map;
    const int rc = make_map(&map, table, rows);
    if (rc) {
      printf("  gather4: tensor map failed (%d)\n", rc);
    } else {
      {
        constexpr int S = 6, C = 8;
        const int smem = S * kStageBytes + 2 * S * 8;
        CK(cudaFuncSetAttribute(gather_tma<S, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        float ms = time_ms([&] { gather_tma<S, C><<<sms, 32 * (1 + C), smem>>>(map, idx, n, out); }, 5);
        cudaError_t e = cudaGetLastError();
        printf("  gather4  6 stages x 32 KB, 8 consumer warps, 1 CTA per SM   : %.3f ms  %.2f TB/s  (%s)\n", ms, gb / ms,
               cudaGetErrorString(e));
      }
      {
        constexpr int S = 3, C = 4;
        const int smem = S * kStageBytes + 2 * S * 8;
        CK(cudaFuncSetAttribute(gather_tma<S, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        float ms = time_ms([&] { gather_tma<S, C><<<sms * 2, 32 * (1 + C), smem>>>(map, idx, n, out); }, 5);
        cudaError_t e = cudaGetLastError();
        printf("  gather4  3 stages x 32 KB, 4 consumer warps, 2 CTAs per SM  : %.3f ms  %.2f TB/s  (%s)\n", ms, gb / ms,
               cudaGetErrorString(e));
      }
    }
    if (rows == 105542u) {
      // the same table read with the item popularity of the synthetic H&M graph (what the user rows of a layer do)
      fill_idx_zipf<<<sms * 8, 256>>>(idx, n, rows, 777u);
      CK(cudaDeviceSynchronize());
      for (int ctas_per_sm : {4, 8}) {
        float ms = time_ms([&] { gather_ldg<8><<<sms * ctas_per_sm, 256>>>(table, idx, n, out); }, 5);
        printf("  ldg  ZIPF 8 rows in flight per lane, %d CTAs x 256 thr per SM : %.3f ms  %.2f TB/s\n", ctas_per_sm, ms, gb / ms);
      }
    }
    CK(cudaFree(table));
  }
  printf("for comparison, one LightGCN layer at the H&M shape gathers 16.65 GB (65.05 M rows): 1.76 ms = 9.5 TB/s\n");
  return 0;
}
