// Micro-benchmark for the NEXT select epilogue of score_topk_fused_kernel:
// cycles per 32-row x 128-column accumulator quarter for
//   A  the present hot path: tcgen05.ld.32x32b.x32 (fp32 accumulators), 8 group maxima of 4 columns,
//      8 bucket updates, one chunk maximum and one (never taken) branch per 32 columns;
//   B  the candidate: fp16 accumulators read with tcgen05.ld ... .pack::16b (two columns per register) and
//      half2 maxima -- 8 pair maxima, 8 bucket updates on 16 registers of 64 half buckets, one chunk maximum.
// TMEM is read uninitialised (timing only).  12 warps per CTA = 3 per scheduler, as in the kernel; 4 and 8 for
// comparison.  Written at the end of round 1 without GPU time left: compiled, not yet run.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bench_select_epilogue bench_select_epilogue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
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
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 columns of fp16 accumulators packed two per register
__device__ __forceinline__ void ld32_packed(uint32_t taddr, __half2 (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = *reinterpret_cast<__half2*>(&r[i]);
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cycles, int iters, float tau_in) {
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
  const uint32_t base = tmem_base_s + ((uint32_t)((warp & 3) * 32) << 16) + ((warp >> 2) % 3) * 128;
  float sink = 0.f;
  int hits = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (MODE == 0) {
    float bm[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) bm[i] = -INFINITY;
    const float tau = tau_in;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float v[32], q[8];
        ld32(base + c * 32, v);
#pragma unroll
        for (int h = 0; h < 8; ++h) q[h] = fmaxf(fmaxf(v[4 * h], v[4 * h + 1]), fmaxf(v[4 * h + 2], v[4 * h + 3]));
#pragma unroll
        for (int h = 0; h < 8; ++h) bm[c * 8 + h] = fmaxf(bm[c * 8 + h], q[h]);
        const float m32 = fmaxf(fmaxf(fmaxf(q[0], q[1]), fmaxf(q[2], q[3])), fmaxf(fmaxf(q[4], q[5]), fmaxf(q[6], q[7])));
        if (m32 > tau) ++hits;
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) sink += bm[i];
  } else {
    __half2 bm[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) bm[i] = __float2half2_rn(-60000.f);
    const __half2 tau2 = __float2half2_rn(tau_in);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        __half2 v[16], q[8];
        ld32_packed(base + c * 32, v);
        // register i holds columns (2i, 2i+1): q[h] = {max(c0, c2), max(c1, c3)} of group h
#pragma unroll
        for (int h = 0; h < 8; ++h) q[h] = __hmax2(v[2 * h], v[2 * h + 1]);
        // 64 half buckets in 16 registers: chunks 0,1 feed bm[0..7], chunks 2,3 feed bm[8..15]
#pragma unroll
        for (int h = 0; h < 8; ++h) bm[(c >> 1) * 8 + h] = __hmax2(bm[(c >> 1) * 8 + h], q[h]);
        const __half2 m = __hmax2(__hmax2(__hmax2(q[0], q[1]), __hmax2(q[2], q[3])),
                                  __hmax2(__hmax2(q[4], q[5]), __hmax2(q[6], q[7])));
        if (__hbgt2(m, tau2) || __low2float(m) > __high2float(tau2)) ++hits;      // either half above tau
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) sink += __low2float(bm[i]) + __high2float(bm[i]);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink + (float)hits;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(512));
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 4000;
  for (int mode : {0, 1}) {
    for (int warps : {4, 8, 12}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, cyc, iters, 3.0e38f);
        else k<1><<<148, warps * 32>>>(out, cyc, iters, 60000.f);
      }
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      // per scheduler: warps / 4 warps, each doing `iters` quarters
      printf("%s  warps %2d: %.0f cycles per quarter and warp, %.0f cycles per quarter and scheduler  (%s)\n",
             mode == 0 ? "A fp32  x32 loads, FMNMX    " : "B fp16  packed loads, HMNMX2", warps, (double)h[0] / iters,
             (double)h[0] / iters / (warps / 4), cudaGetErrorString(e));
    }
  }
  printf("the kernel today: ~830 cycles per quarter and scheduler (3 warps), profiles/r1_fused_notes.md\n");
  return 0;
}
