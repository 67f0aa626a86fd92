// Shared helpers for libhnm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "hnm_b200.h"

#define HNM_CUDA_TRY(expr)                         \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)

#define HNM_LAUNCH_CHECK() HNM_CUDA_TRY(cudaGetLastError())

static inline bool hnm_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// SM count of the CURRENT device (cached per device ordinal: one process may drive several GPUs).
static inline int hnm_num_sms() {
  static int sms[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (sms[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    sms[dev] = n > 0 ? n : 148;
  }
  return sms[dev];
}

// Function attributes are per device and a process may drive several GPUs, so the dynamic shared-memory
// limit is (re)set on every launch; the call is a host-side table update (well under a microsecond).
template <typename K>
static inline cudaError_t hnm_allow_smem(K kernel, int bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// (score desc, id asc): does a come strictly before b?
__device__ __forceinline__ bool hnm_before(double sa, int64_t ia, double sb, int64_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}
