// Microbenchmark: the three-pass radix-16 4096-point engine (fft4096.cuh), 5 transforms per "frame":
//   A: 2 groups x 256 threads x 16 values (the shipped mapping)     B: 4 groups x 128 threads x 2 x 16 values (two virtual threads each)
#include <cstdio>
#include <vector>
#define OMB_F32X2_CMUL 0
#include "../../openmeters_b200/csrc/fft4096.cuh"
using namespace omb;
using namespace omb::f4k;
constexpr int kWSize = f16::phys_size(4096);
template <int NT>
__device__ __forceinline__ void gsync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "n"(NT) : "memory"); }

template <int kGroups, int kVT>  // kVT virtual threads per thread; group = 256 / kVT threads
__global__ void __launch_bounds__(kGroups * 256 / kVT, 1) k_core16(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ twg, int nframes) {
  constexpr int NT = 256 / kVT;
  extern __shared__ __align__(16) unsigned char smem[];
  float2* tw1 = reinterpret_cast<float2*>(smem);             // [4][256]
  float2* tw2 = tw1 + 4 * 256;                               // [15][16]
  float2* Wall = tw2 + 15 * 16;
  const int tid = threadIdx.x, tl = tid % NT;
  const int g = __shfl_sync(0xffffffffu, tid / NT, 0);
  float2* W = Wall + g * kWSize;
  for (int i = tid; i < 4 * 256 + 15 * 16; i += kGroups * NT) tw1[i] = twg[i];
  __syncthreads();
  Addr ad[kVT];
  const float2* tw1t[kVT];
  const float2* tw2o[kVT];
#pragma unroll
  for (int u = 0; u < kVT; ++u) {
    const int t = tl + NT * u;
    ad[u].pA = t + (t >> 4);
    ad[u].pB = 273 * (t >> 4) + (t & 15);
    ad[u].pC = 273 * (t & 15) + 17 * (t >> 4);
    tw1t[u] = tw1 + t;
    tw2o[u] = tw2 + (t & 15);
  }
  for (int f = blockIdx.x * kGroups + g; f < nframes; f += gridDim.x * kGroups) {
    float2 v[kVT][16];
#pragma unroll
    for (int u = 0; u < kVT; ++u)
#pragma unroll
      for (int j = 0; j < 16; ++j) v[u][j] = in[(size_t)(f & 1023) * 4096 + tl + NT * u + 256 * j];
#pragma unroll 1
    for (int tr = 0; tr < 5; ++tr) {
#pragma unroll
      for (int u = 0; u < kVT; ++u) {
        f16::dft16<false>(v[u]);
        twiddle15<false, 1, true>(v[u], tw1t[u], 256);
        float2* wa = W + ad[u].pA;
#pragma unroll
        for (int q = 0; q < 16; ++q) wa[273 * q] = v[u][q];
      }
      gsync<NT>(g);
#pragma unroll
      for (int u = 0; u < kVT; ++u) {
        float2* wb = W + ad[u].pB;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[u][j] = wb[17 * j];
        f16::dft16<false>(v[u]);
        twiddle15<false, 0, false>(v[u], tw2o[u], 16);
#pragma unroll
        for (int q = 0; q < 16; ++q) wb[17 * q] = v[u][q];
      }
      gsync<NT>(g);
#pragma unroll
      for (int u = 0; u < kVT; ++u) {
        const float2* wc = W + ad[u].pC;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[u][j] = wc[j];
        f16::dft16<false>(v[u]);
      }
      gsync<NT>(g);
    }
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < kVT; ++u)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc += v[u][j].x + v[u][j].y;
    if (acc == 123.456f) out[f & 1023] = v[0][0];
  }
}
template <int kGroups, int kVT>
void run(const char* name, const float2* in, float2* out, const float2* tw) {
  const int smem = (4 * 256 + 15 * 16 + kGroups * kWSize) * 8;
  cudaFuncSetAttribute(k_core16<kGroups, kVT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int nframes = 148 * 4 * 110;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) k_core16<kGroups, kVT><<<148, kGroups * 256 / kVT, smem>>>(in, out, tw, nframes);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) k_core16<kGroups, kVT><<<148, kGroups * 256 / kVT, smem>>>(in, out, tw, nframes);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_core16<kGroups, kVT>);
  printf("%-34s regs %3d local %4zu  %.3f ms per launch  %.3e frames/s (5 transforms each)  %s\n", name, fa.numRegs, fa.localSizeBytes, ms / 5, nframes / (ms / 5 * 1e-3), cudaGetErrorString(e));
}
int main() {
  float2 *in, *out, *tw;
  cudaMalloc(&in, 1024 * 4096 * 8); cudaMalloc(&out, 1024 * 4096 * 8); cudaMalloc(&tw, (4 * 256 + 240) * 8);
  std::vector<float2> h(1024 * 4096);
  for (size_t i = 0; i < h.size(); ++i) h[i] = make_float2((float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f - 0.5f, (float)((i * 40503u) & 0xffff) / 65536.f - 0.5f);
  cudaMemcpy(in, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  std::vector<float2> t(4 * 256 + 240, make_float2(0.8f, 0.6f));
  cudaMemcpy(tw, t.data(), t.size() * 8, cudaMemcpyHostToDevice);
  run<2, 1>("2 groups x 256 thr x 16 values", in, out, tw);
  run<4, 2>("4 groups x 128 thr x 2x16 values", in, out, tw);
  run<4, 1>("4 groups x 256 thr x 16 (64 regs)", in, out, tw);
  run<3, 1>("3 groups x 256 thr x 16 (85 regs)", in, out, tw);
  run<4, 4>("4 groups x 64 thr x 4x16 values", in, out, tw);
  run<6, 2>("6 groups x 128 thr x 2x16 values", in, out, tw);
  return 0;
}
