// runtime.cu — error state, launch accounting, device discovery.
#include "common.h"

namespace omb {

std::string& last_error_ref() {
  static thread_local std::string msg;
  return msg;
}

int fail(int status, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return status;
}

std::atomic<uint64_t>& launch_count() {
  static std::atomic<uint64_t> c{0};
  return c;
}

int current_device(DeviceInfo* out) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return fail(OMB_ERR_CUDA, "no usable CUDA device (%s); libomb200 has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  DeviceInfo d;
  OMB_CUDA_TRY(cudaGetDevice(&d.device));
  OMB_CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, d.device));
  OMB_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, d.device));
  OMB_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, d.device));
  OMB_CUDA_TRY(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, d.device));
  *out = d;
  return OMB_OK;
}

}  // namespace omb
