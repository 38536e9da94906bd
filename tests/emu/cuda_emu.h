// cuda_emu.h — DEVELOPMENT/TEST AID ONLY (never part of the product build).
//
// A minimal CUDA execution-model emulator so the kernel *sources* under
// openmeters_b200/csrc can be compiled with g++ (-DOMB_EMU) and executed
// thread-for-thread on the CPU: every CUDA thread of a block is a ucontext
// fiber; __syncthreads / warp collectives are cooperative barriers.  It exists
// because the dev container has no GPU: it catches indexing / barrier / layout
// bugs before GPU minutes are spent.  The product (libomb200.so, nvcc) contains
// none of this and has no CPU path; tests/emu builds a separate
// libomb200_emu.so that only tests/ loads.
//
// Not emulated: inline PTX (TMA bulk copies, mbarrier) — those code paths have a
// plain-C++ equivalent under #ifdef OMB_EMU next to them and are validated on
// the GPU only.
#pragma once
#ifndef OMB_EMU
#error "cuda_emu.h is only for -DOMB_EMU builds"
#endif

#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__ __restrict
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static thread_local
#define __constant__ static

struct dim3 {
  unsigned x = 1, y = 1, z = 1;
  dim3() = default;
  dim3(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct ushort2 { unsigned short x, y; };
struct ushort4 { unsigned short x, y, z, w; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return {x, y}; }
static inline ushort2 make_ushort2(unsigned short x, unsigned short y) { return {x, y}; }
static inline ushort4 make_ushort4(unsigned short x, unsigned short y, unsigned short z, unsigned short w) { return {x, y, z, w}; }

namespace omb_emu {

struct Fiber {
  ucontext_t ctx;
  std::vector<char> stack;
  uint3 tidx{0, 0, 0};
  unsigned linear = 0;
  bool done = false;
};

struct WarpState {
  uint32_t slot[32];
  int arrived = 0;
  int gen = 0;
};

struct NamedBar { int arrived = 0; int gen = 0; };

struct BlockState {
  NamedBar named[16];
  dim3 grid, block;
  uint3 bidx{0, 0, 0};
  unsigned nthreads = 0;
  int arrived = 0;
  int gen = 0;
  std::vector<Fiber> fibers;
  std::vector<WarpState> warps;
  ucontext_t sched;
  Fiber* cur = nullptr;
  const std::function<void()>* body = nullptr;
  std::vector<char> dyn_smem;
  std::vector<uint32_t> tmem;  // tensor memory of the CTA: [128 lanes][512 columns] (tmem_park.cuh), allocated on first use
};

inline BlockState*& tls_block() {
  static thread_local BlockState* b = nullptr;
  return b;
}
inline BlockState& blk() { return *tls_block(); }
inline Fiber& cur() { return *blk().cur; }

inline void yield() {
  BlockState& b = blk();
  swapcontext(&b.cur->ctx, &b.sched);
}

inline void syncthreads() {
  BlockState& b = blk();
  const int gen = b.gen;
  if (++b.arrived == (int)b.nthreads) {
    b.arrived = 0;
    b.gen++;
  } else {
    while (b.gen == gen) yield();
  }
}

// bar.sync id, count
inline void named_sync(int id, int count) {
  NamedBar& nb = blk().named[id & 15];
  const int gen = nb.gen;
  if (++nb.arrived == count) {
    nb.arrived = 0;
    nb.gen++;
  } else {
    while (nb.gen == gen) yield();
  }
}

// bar.arrive id, count: counts like bar.sync but does not wait
inline void named_arrive(int id, int count) {
  NamedBar& nb = blk().named[id & 15];
  if (++nb.arrived == count) {
    nb.arrived = 0;
    nb.gen++;
  }
}

inline unsigned warp_lanes(const BlockState& b, unsigned warp) {
  const unsigned first = warp * 32;
  return std::min(32u, b.nthreads - first);
}

inline void warp_sync() {
  BlockState& b = blk();
  const unsigned w = cur().linear / 32;
  WarpState& ws = b.warps[w];
  const int gen = ws.gen;
  if (++ws.arrived == (int)warp_lanes(b, w)) {
    ws.arrived = 0;
    ws.gen++;
  } else {
    while (ws.gen == gen) yield();
  }
}

inline uint32_t warp_exchange(uint32_t v, int src_lane) {
  BlockState& b = blk();
  WarpState& ws = b.warps[cur().linear / 32];
  const unsigned lane = cur().linear % 32;
  ws.slot[lane] = v;
  warp_sync();
  const uint32_t r = (src_lane >= 0 && src_lane < (int)warp_lanes(b, cur().linear / 32)) ? ws.slot[src_lane] : v;
  warp_sync();
  return r;
}

inline unsigned ballot(int pred) {
  BlockState& b = blk();
  WarpState& ws = b.warps[cur().linear / 32];
  const unsigned lane = cur().linear % 32;
  ws.slot[lane] = pred ? 1u : 0u;
  warp_sync();
  unsigned r = 0;
  const unsigned n = warp_lanes(b, cur().linear / 32);
  for (unsigned l = 0; l < n; ++l) if (ws.slot[l]) r |= 1u << l;
  warp_sync();
  return r;
}

inline void fiber_entry() {
  BlockState& b = blk();
  (*b.body)();
  b.cur->done = true;
  // a finished thread counts as permanently arrived at later barriers is NOT modelled:
  // kernels must not return before their last barrier (same rule as real CUDA).
  swapcontext(&b.cur->ctx, &b.sched);
}

inline void run_block(BlockState& b) {
  tls_block() = &b;
  b.arrived = 0;
  b.gen = 0;
  for (auto& w : b.warps) { w.arrived = 0; w.gen = 0; }
  for (auto& n : b.named) { n.arrived = 0; n.gen = 0; }
  const size_t stack_bytes = 192 * 1024;
  unsigned lin = 0;
  for (unsigned z = 0; z < b.block.z; ++z)
    for (unsigned y = 0; y < b.block.y; ++y)
      for (unsigned x = 0; x < b.block.x; ++x, ++lin) {
        Fiber& f = b.fibers[lin];
        f.tidx = {x, y, z};
        f.linear = lin;
        f.done = false;
        if (f.stack.size() != stack_bytes) f.stack.resize(stack_bytes);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.data();
        f.ctx.uc_stack.ss_size = f.stack.size();
        f.ctx.uc_link = &b.sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
      }
  unsigned remaining = b.nthreads;
  while (remaining) {
    unsigned progressed = 0;
    for (auto& f : b.fibers) {
      if (f.done) continue;
      b.cur = &f;
      swapcontext(&b.sched, &f.ctx);
      if (f.done) { --remaining; }
      ++progressed;
    }
    if (!progressed) break;
  }
  tls_block() = nullptr;
}

inline std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> c{0};
  return c;
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  launch_counter()++;
  const unsigned nblocks = grid.x * grid.y * grid.z;
  const unsigned nthreads = block.x * block.y * block.z;
  unsigned workers = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), nblocks);
  if (const char* e = getenv("OMB_EMU_THREADS")) workers = std::max(1, atoi(e));
  std::atomic<unsigned> next{0};
  auto worker = [&]() {
    BlockState b;
    b.grid = grid;
    b.block = block;
    b.nthreads = nthreads;
    b.fibers.resize(nthreads);
    b.warps.resize((nthreads + 31) / 32);
    b.body = &body;
    b.dyn_smem.resize(smem + 64);
    for (unsigned i; (i = next.fetch_add(1)) < nblocks;) {
      b.bidx = {i % grid.x, (i / grid.x) % grid.y, i / (grid.x * grid.y)};
      run_block(b);
    }
  };
  if (workers <= 1) { worker(); return; }
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < workers; ++t) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
}

// Tensor memory (tcgen05.ld / tcgen05.st in tmem_park.cuh): a plain per-CTA array, [lane][column].
inline uint32_t* tmem() {
  auto& v = blk().tmem;
  if (v.empty()) v.assign(128 * 512, 0u);
  return v.data();
}

inline void* dyn_smem() {
  auto& v = blk().dyn_smem;
  uintptr_t p = (uintptr_t)v.data();
  p = (p + 63) & ~uintptr_t(63);
  return (void*)p;
}

}  // namespace omb_emu

#define threadIdx (omb_emu::cur().tidx)
#define blockIdx (omb_emu::blk().bidx)
#define blockDim (omb_emu::blk().block)
#define gridDim (omb_emu::blk().grid)
#define warpSize 32

static inline void __syncthreads() { omb_emu::syncthreads(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { omb_emu::warp_sync(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

static inline unsigned __ballot_sync(unsigned, int pred) { return omb_emu::ballot(pred); }
static inline int __any_sync(unsigned, int pred) { return omb_emu::ballot(pred) != 0; }
static inline int __all_sync(unsigned, int pred) { return omb_emu::ballot(!pred) == 0; }

template <class T> static inline T omb_emu_shfl(T v, int src) {
  static_assert(sizeof(T) == 4 || sizeof(T) == 8, "shfl size");
  if constexpr (sizeof(T) == 4) {
    uint32_t u; std::memcpy(&u, &v, 4);
    u = omb_emu::warp_exchange(u, src);
    T r; std::memcpy(&r, &u, 4); return r;
  } else {
    uint32_t u[2]; std::memcpy(u, &v, 8);
    u[0] = omb_emu::warp_exchange(u[0], src);
    u[1] = omb_emu::warp_exchange(u[1], src);
    T r; std::memcpy(&r, u, 8); return r;
  }
}
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  const int lane = omb_emu::cur().linear % 32;
  return omb_emu_shfl(v, (lane / width) * width + (src % width));
}
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
  const int lane = omb_emu::cur().linear % 32;
  return omb_emu_shfl(v, lane ^ m);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = omb_emu::cur().linear % 32;
  const int src = lane + (int)d;
  return omb_emu_shfl(v, (src / width == lane / width) ? src : lane);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
  const int lane = omb_emu::cur().linear % 32;
  const int src = lane - (int)d;
  return omb_emu_shfl(v, (src >= 0 && src / width == lane / width) ? src : lane);
}

static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) {  // REDUX.MAX.U32
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned other = __shfl_xor_sync(mask, v, o);
    v = other > v ? other : v;
  }
  return v;
}

using std::isfinite;
using std::isinf;
using std::isnan;

// ---- math / bit intrinsics
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline unsigned __float2uint_rd(float a) {  // saturating, NaN -> 0, like cvt.rmi.u32.f32
  if (!(a > 0.0f)) return 0u;
  const float f = std::floor(a);
  return f >= 4294967296.0f ? 0xffffffffu : (unsigned)f;
}
static inline float __fdividef(float a, float b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline void sincospif(float x, float* s, float* c) { *s = (float)std::sin(3.14159265358979323846 * (double)x); *c = (float)std::cos(3.14159265358979323846 * (double)x); }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline unsigned __brev(unsigned x) {
  x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
  x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
  x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
  x = ((x >> 8) & 0x00ff00ffu) | ((x & 0x00ff00ffu) << 8);
  return (x >> 16) | (x << 16);
}
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline float fminf_(float a, float b) { return std::fmin(a, b); }

template <class T> static inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline float atomicAdd(float* p, float v) {  // blocks run on worker threads: CAS loop on the bit pattern
  unsigned* u = reinterpret_cast<unsigned*>(p);
  unsigned old = __atomic_load_n(u, __ATOMIC_RELAXED), want;
  float f;
  do {
    std::memcpy(&f, &old, 4);
    f += v;
    std::memcpy(&want, &f, 4);
  } while (!__atomic_compare_exchange_n(u, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  std::memcpy(&f, &old, 4);
  return f;
}
static inline unsigned atomicMax(unsigned* p, unsigned v) {
  unsigned o = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long o = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}
static inline int atomicMax(int* p, int v) {
  int o = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (o < v && !__atomic_compare_exchange_n(p, &o, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}

// ---- runtime API subset
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef struct omb_emu_event { double t; }* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrComputeCapabilityMajor = 75, cudaDevAttrComputeCapabilityMinor = 76,
       cudaDevAttrMaxSharedMemoryPerBlockOptin = 97 };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) {
  switch (attr) {
    case cudaDevAttrMultiProcessorCount: *v = 4; break;
    case cudaDevAttrComputeCapabilityMajor: *v = 10; break;
    case cudaDevAttrComputeCapabilityMinor: *v = 0; break;
    case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = 227 * 1024; break;
    default: *v = 0;
  }
  return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
// CUDA IPC under the emulator: one process, so a handle simply carries the pointer
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
static inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof(*h)); std::memcpy(h->reserved, &p, sizeof(p)); return cudaSuccess; }
static inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, h.reserved, sizeof(*p)); return cudaSuccess; }
static inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind,
                                            cudaStream_t = nullptr) {
  for (size_t r = 0; r < height; ++r) std::memmove(static_cast<char*>(d) + r * dpitch, static_cast<const char*>(s) + r * spitch, width);
  return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { if (n) std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new omb_emu_event{0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
