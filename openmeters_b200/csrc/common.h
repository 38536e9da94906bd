// common.h — shared plumbing of libomb200 (error state, launch accounting, device buffers).
#pragma once

#ifdef OMB_EMU
#include "cuda_emu.h"  // tests/emu only; see that header. The nvcc build never defines OMB_EMU.
#else
#include <cuda_runtime.h>
#endif

// Packed FP32x2 arithmetic (FADD2 / FMUL2 / FFMA2, see fft16.cuh) in the device helpers.  Each class of primitive has its own
// switch because they were measured separately on B200 (profiles/r02b_packed_ab.md):
//   OMB_F32X2_ADD   complex add / subtract as one FADD2                          default 1  (round 1: +8 % on the cfg2 kernel)
//   OMB_F32X2_ROT   a -+ j b as one FADD2 with lane-swap / negate modifiers      default 0
//   OMB_F32X2_MUL   complex products of the radix-16 engine as FMUL2 + FFMA2     default 0  (-7 % cfg2, -9 % N = 2048, -12 % N = 1024:
//                   the engine is bound by the FMA pipe, which a packed instruction holds for two cycles, not by issue slots)
//   OMB_F32X2_CMUL  device_math.cuh cmul / cmul_conj (Stockham, generic tiers)   default 1  (+1 ... +16 % on those tiers)
// The CPU emulator and -DOMB_NO_F32X2 force scalar code everywhere.
#if defined(OMB_EMU) || defined(OMB_NO_F32X2)
#undef OMB_F32X2_ADD
#undef OMB_F32X2_ROT
#undef OMB_F32X2_MUL
#undef OMB_F32X2_CMUL
#define OMB_F32X2_ADD 0
#define OMB_F32X2_ROT 0
#define OMB_F32X2_MUL 0
#define OMB_F32X2_CMUL 0
#endif
#ifndef OMB_F32X2_ADD
#define OMB_F32X2_ADD 1
#endif
#ifndef OMB_F32X2_ROT
#define OMB_F32X2_ROT 0
#endif
#ifndef OMB_F32X2_MUL
#define OMB_F32X2_MUL 0
#endif
#ifndef OMB_F32X2_CMUL
#define OMB_F32X2_CMUL 1
#endif

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/omb200.h"

namespace omb {

// ---- error state (thread-local message behind omb_last_error)
std::string& last_error_ref();
int fail(int status, const char* fmt, ...);

#define OMB_CUDA_TRY(expr)                                                                      \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::omb::fail(OMB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                                   \
  } while (0)

#define OMB_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s < 0) return _s;       \
  } while (0)

// ---- launch accounting (omb_kernel_launch_count)
std::atomic<uint64_t>& launch_count();

#ifdef OMB_EMU
#define OMB_LAUNCH(kern, grid, block, smem, stream, ...)                                      \
  do {                                                                                        \
    ::omb::launch_count()++;                                                                  \
    auto _omb_args = std::make_tuple(__VA_ARGS__);                                            \
    ::omb_emu::launch((grid), (block), (smem), [&]() { std::apply(kern, _omb_args); });       \
  } while (0)
#else
#define OMB_LAUNCH(kern, grid, block, smem, stream, ...)                                      \
  do {                                                                                        \
    ::omb::launch_count()++;                                                                  \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                 \
  } while (0)
#endif

#define OMB_CHECK_LAUNCH()                                                                    \
  do {                                                                                        \
    cudaError_t _e = cudaGetLastError();                                                      \
    if (_e != cudaSuccess)                                                                    \
      return ::omb::fail(OMB_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// Dynamic shared memory, typed.
#ifdef OMB_EMU
#define OMB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(::omb_emu::dyn_smem())
#else
#define OMB_DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(1024) unsigned char name##_raw_[];   \
  type* name = reinterpret_cast<type*>(name##_raw_)
#endif

// ---- device info
struct DeviceInfo {
  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  int max_smem_optin = 0;
};
// Fails (OMB_ERR_CUDA) when no usable device exists: there is no CPU fallback.
int current_device(DeviceInfo* out);
int probe_fp32_tflops(double* out_tflops);  // runtime.cu: measured FFMA peak of the current device

// ---- RAII device / pinned-host buffers
template <class T>
struct DeviceBuffer {
  T* ptr = nullptr;
  size_t cap = 0;  // elements
  DeviceBuffer() = default;
  DeviceBuffer(const DeviceBuffer&) = delete;
  DeviceBuffer& operator=(const DeviceBuffer&) = delete;
  ~DeviceBuffer() { release(); }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
  // Grows (never shrinks); contents are NOT preserved.
  int reserve(size_t n) {
    if (n <= cap) return OMB_OK;
    release();
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return fail(OMB_ERR_NOMEM, "cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
    ptr = static_cast<T*>(p);
    cap = n;
    return OMB_OK;
  }
  int upload(const T* host, size_t n, cudaStream_t s) {
    OMB_TRY(reserve(n));
    if (n) OMB_CUDA_TRY(cudaMemcpyAsync(ptr, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
    return OMB_OK;
  }
  int upload(const std::vector<T>& v, cudaStream_t s) { return upload(v.data(), v.size(), s); }
};

template <class T>
struct PinnedBuffer {
  T* ptr = nullptr;
  size_t cap = 0;
  PinnedBuffer() = default;
  PinnedBuffer(const PinnedBuffer&) = delete;
  PinnedBuffer& operator=(const PinnedBuffer&) = delete;
  ~PinnedBuffer() { release(); }
  void release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
  }
  int reserve(size_t n) {
    if (n <= cap) return OMB_OK;
    release();
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e != cudaSuccess) return fail(OMB_ERR_NOMEM, "cudaMallocHost(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
    ptr = static_cast<T*>(p);
    cap = n;
    return OMB_OK;
  }
};

static inline bool is_pow2(uint64_t v) { return v && !(v & (v - 1)); }
static inline int ilog2(uint64_t v) { int l = 0; while ((uint64_t(1) << l) < v) ++l; return l; }
static inline uint64_t next_pow2(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return p; }

}  // namespace omb
