// Micro-benchmark: TMEM -> register bandwidth of tcgen05.ld.32x32b.{x16,x32,x64} with 4..16 warps per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_tmem bench_tmem.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int X>
__device__ __forceinline__ float ld_sum(uint32_t taddr);

template <>
__device__ __forceinline__ float ld_sum<32>(uint32_t taddr) {
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
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s ^= r[i];
  return __uint_as_float(s);
}
template <>
__device__ __forceinline__ float ld_sum<16>(uint32_t taddr) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s ^= r[i];
  return __uint_as_float(s);
}

// two x32 loads in flight before one wait
__device__ __forceinline__ float ld_sum_2x32(uint32_t taddr) {
  uint32_t r[64];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t* q = r + 32 * h;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
          "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
          "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
          "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
        : "r"(taddr + 32 * h)
        : "memory");
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 64; ++i) s ^= r[i];
  return __uint_as_float(s);
}

template <int MODE>
__global__ void k(float* out, long long* cycles, int iters) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(&tmem_base_s)),
                 "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint32_t col = ((it + warp) & 3) * 128 + c * 32;
      if (MODE == 32) acc += ld_sum<32>(base + col);
      else if (MODE == 16) { acc += ld_sum<16>(base + col); acc += ld_sum<16>(base + col + 16); }
      else if (c % 2 == 0) acc += ld_sum_2x32(base + col);
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512));
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2000;
  for (int mode : {16, 32, 64}) {
    for (int warps : {4, 8, 12, 16}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 16) k<16><<<148, warps * 32>>>(out, cyc, iters);
        else if (mode == 32) k<32><<<148, warps * 32>>>(out, cyc, iters);
        else k<64><<<148, warps * 32>>>(out, cyc, iters);
      }
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double bytes = (double)iters * 4 * 32 * 32 * 4 * warps;   // per SM
      printf("mode x%d warps %2d: %lld cycles, %.1f B/cycle/SM, %.1f cycles per 4KB warp-load  (%s)\n", mode, warps,
             h[0], bytes / h[0], (double)h[0] / (iters * 4.0), cudaGetErrorString(e));
    }
  }
  return 0;
}
