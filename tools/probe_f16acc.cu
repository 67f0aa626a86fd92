// Probe: TMEM layout of an fp16 accumulator (tcgen05.mma kind::f16, D format F16), M=128 N=128 K=16.
// A[r][0] = 1, B[n][0] = n  ->  D[r][n] = n.  Prints what lane 0 / lane 5 read from the first 32 columns.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
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
__device__ void put(uint8_t* tile, int r, int k, float v) {   // 128-byte rows, SWIZZLE_128B
  const int chunk = (k * 2) / 16, within = (k * 2) % 16;
  const int off = (r / 8) * 1024 + (r % 8) * 128 + ((chunk ^ (r % 8)) * 16) + within;
  *reinterpret_cast<__half*>(tile + off) = __float2half(v);
}
__global__ void k(uint32_t* out, int cfmt) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  __syncthreads();
  for (int r = threadIdx.x; r < 128; r += blockDim.x) { put(smem, r, 0, 1.0f); put(smem + 16384, r, 0, (float)r); }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tbase)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = ((uint32_t)cfmt << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  if (threadIdx.x == 0) {
    const uint64_t ad = desc_sw128(smem_u32(smem)), bd = desc_sw128(smem_u32(smem + 16384));
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tbase),
                 "l"(ad), "l"(bd), "r"(idesc), "r"(0)
                 : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 32) {
    uint32_t r[32];
    for (int chunk = 0; chunk < 4; ++chunk) {
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
          "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(tbase + chunk * 32)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int i = 0; i < 32; ++i) out[(threadIdx.x * 4 + chunk) * 32 + i] = r[i];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256));
}
int main() {
  uint32_t* d; cudaMalloc(&d, 32 * 4 * 32 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int cfmt : {1, 0}) {
    cudaMemset(d, 0xff, 32 * 4 * 32 * 4);
    k<<<1, 128, 40000>>>(d, cfmt);
    cudaError_t e = cudaDeviceSynchronize();
    uint32_t h[32 * 4 * 32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("c_format=%d (%s): %s\n", cfmt, cfmt ? "f32" : "f16", cudaGetErrorString(e));
    for (int lane : {0, 5}) for (int chunk : {0, 1, 3}) {
      printf(" lane %d cols %3d..: ", lane, chunk * 32);
      for (int i = 0; i < 8; ++i) {
        uint32_t v = h[(lane * 4 + chunk) * 32 + i];
        if (cfmt) printf("%g ", *(float*)&v);
        else printf("[%g,%g] ", __half2float(*(__half*)&v), __half2float(*((__half*)&v + 1)));
      }
      printf("\n");
    }
  }
  return 0;
}
