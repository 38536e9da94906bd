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

// FP32 peak probe: 16 independent FFMA chains per thread, 1024-thread CTAs, two per SM.
__global__ void __launch_bounds__(1024) k_fp32_probe(float* sink, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  if (s == 123.456f) sink[0] = s;  // never true in practice; keeps the chains alive
}

int probe_fp32_tflops(double* out) {
  if (!out) return fail(OMB_ERR_INVALID, "null argument");
#ifdef OMB_EMU
  return fail(OMB_ERR_UNSUPPORTED, "the FP32 probe measures a device; not available under the emulator");
#endif
  DeviceInfo dev;
  OMB_TRY(current_device(&dev));
  DeviceBuffer<float> sink;  // RAII: released on every return path
  OMB_TRY(sink.reserve(1));
  struct Events {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~Events() {
      if (e0) cudaEventDestroy(e0);
      if (e1) cudaEventDestroy(e1);
    }
  } ev;
  OMB_CUDA_TRY(cudaEventCreate(&ev.e0));
  OMB_CUDA_TRY(cudaEventCreate(&ev.e1));
  const int iters = 1 << 14;
  const unsigned grid = (unsigned)std::max(dev.sm_count, 1) * 2u;
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {  // first repetition warms up
    OMB_CUDA_TRY(cudaEventRecord(ev.e0, nullptr));
    OMB_LAUNCH(k_fp32_probe, dim3(grid), dim3(1024), 0, nullptr, sink.ptr, iters, 0.999f, 0.001f);
    OMB_CHECK_LAUNCH();
    OMB_CUDA_TRY(cudaEventRecord(ev.e1, nullptr));
    OMB_CUDA_TRY(cudaEventSynchronize(ev.e1));
    float ms = 0.0f;
    OMB_CUDA_TRY(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
    const double flops = 2.0 * 16.0 * (double)iters * 1024.0 * (double)grid;
    if (rep > 0 && ms > 0.0f) best = std::max(best, flops / ((double)ms * 1e-3) / 1e12);
  }
  *out = best;
  return OMB_OK;
}

}  // namespace omb
