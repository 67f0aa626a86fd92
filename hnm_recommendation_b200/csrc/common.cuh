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

static inline int hnm_num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

__device__ __forceinline__ float4 ldg_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// (score desc, id asc): does a come strictly before b?
__device__ __forceinline__ bool hnm_before(double sa, int64_t ia, double sb, int64_t ib) {
  return sa > sb || (sa == sb && ia < ib);
}
