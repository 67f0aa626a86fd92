// Version / error strings / device check.
#include "common.cuh"

extern "C" int hnm_abi_version(void) { return HNM_ABI_VERSION; }

extern "C" const char* hnm_strerror(int code) {
  switch (code) {
    case HNM_OK: return "ok";
    case HNM_E_NULL: return "hnm: required pointer is NULL";
    case HNM_E_RANGE: return "hnm: size or index argument out of range";
    case HNM_E_DIM: return "hnm: unsupported embedding dimension";
    case HNM_E_WORKSPACE: return "hnm: workspace too small";
    case HNM_E_ALIGN: return "hnm: pointer not 16-byte aligned";
    case HNM_E_ARCH: return "hnm: device is not sm_100 (B200)";
    case HNM_E_DRIVER: return "hnm: CUDA driver entry point unavailable";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "hnm: unknown error";
}

extern "C" int hnm_check_device(void) {
  int dev = 0;
  HNM_CUDA_TRY(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  HNM_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  HNM_CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return (major == 10 && minor == 0) ? HNM_OK : HNM_E_ARCH;      // only sm_100a code is in the library
}
