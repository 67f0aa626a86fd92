// Micro-benchmark: issue rate of tcgen05.mma.cta_group::1.kind::f16 (SS operands, SWIZZLE_128B K-major)
// for M=128 and N in {64,128,256}; one CTA per SM, one issuing thread, no epilogue.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* cycles, int iters, int a_tiles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (3 * 16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 32) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 3 * 16384);
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      const uint64_t ad = desc_sw128(a0 + (it % a_tiles) * 16384), bd = desc_sw128(b0);
      const uint32_t d = tbase + ((it & 1) * 256 % 512);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) mma(d, ad + 2 * kk, bd + 2 * kk, idesc, kk > 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile(
        "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(
            smem_u32(&bar))
        : "memory");
    const long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 20000;
  const size_t smem = 1024 + 3 * 16384 + 32768;
  cudaFuncSetAttribute(k<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int n : {64, 128, 256}) {
    for (int grid : {1, 148}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (n == 64) k<64><<<grid, 128, smem>>>(cyc, iters, 3);
        else if (n == 128) k<128><<<grid, 128, smem>>>(cyc, iters, 3);
        else k<256><<<grid, 128, smem>>>(cyc, iters, 3);
      }
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
      const double per = (double)h[0] / iters;   // cycles per 128xNx64 tile (4 MMAs)
      printf("N=%3d grid=%3d: %.1f cycles per 128x%dx64 tile -> %.0f FLOP/cycle/SM (ideal %d cycles)  %s\n", n, grid, per, n,
             2.0 * 128 * n * 64 / per, n * 2, cudaGetErrorString(e));
    }
  }
  return 0;
}
