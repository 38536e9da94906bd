// Microbenchmark: throughput of the two-pass radix-64 transform core alone (5 transforms per "frame"), by team count / pass-B form.
#include <cstdio>
#include <vector>
#include "../../openmeters_b200/csrc/fft64.cuh"
using namespace omb;
constexpr int RS = 65;
__device__ __forceinline__ void team_sync(int team) { asm volatile("bar.sync %0, 64;" ::"r"(1 + team) : "memory"); }
__device__ __forceinline__ float2 cmul_s(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void twiddle63(float2 (&v)[64], const float2* tab) {
  float2 lo[8], hi[8];
#pragma unroll
  for (int i = 1; i < 8; ++i) { lo[i] = tab[(i - 1) * 64]; hi[i] = tab[(6 + i) * 64]; }
#pragma unroll
  for (int q = 1; q < 64; ++q) {
    const int a = q & 7, b = q >> 3;
    const float2 w = b == 0 ? lo[a] : (a == 0 ? hi[b] : cmul_s(lo[a], hi[b]));
    v[q] = f16::mul_tw<false>(v[q], w);
  }
}
template <int kTeams, int kDit>
__global__ void __launch_bounds__(64 * kTeams, 1) k_core(const float2* __restrict__ in, float2* __restrict__ out, const float2* __restrict__ twg, int nframes) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int tid = threadIdx.x, t = tid & 63;
  const int team = __shfl_sync(0xffffffffu, tid >> 6, 0);
  float2* W = reinterpret_cast<float2*>(smem) + team * (64 * RS);
  float2* tw = reinterpret_cast<float2*>(smem) + kTeams * 64 * RS;
  for (int i = tid; i < 14 * 64; i += 64 * kTeams) tw[i] = twg[i];
  __syncthreads();
  for (int f = blockIdx.x * kTeams + team; f < nframes; f += gridDim.x * kTeams) {
    float2 v[64];
    const float2* x = in + (size_t)(f & 1023) * 4096 + t;
#pragma unroll
    for (int j = 0; j < 64; ++j) v[j] = x[64 * j];
#pragma unroll 1
    for (int tr = 0; tr < 5; ++tr) {
      f64pt::dft64<false>(v);
      twiddle63(v, tw + t);
      team_sync(team);
#pragma unroll
      for (int q = 0; q < 64; ++q) W[q * RS + t] = v[q];
      team_sync(team);
#pragma unroll
      for (int s = 0; s < 64; ++s) v[s] = W[t * RS + s];
      if (kDit) f64pt::dft64_dit<false>(v); else f64pt::dft64<false>(v);
    }
    float2* y = out + (size_t)(f & 1023) * 4096 + t;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 64; ++j) acc += v[j].x + v[j].y;
    if (acc == 123.456f) y[0] = v[0];
  }
}
template <int kTeams, int kDit>
void run(const char* name, const float2* in, float2* out, const float2* tw) {
  const int smem = kTeams * 64 * RS * 8 + 14 * 64 * 8;
  cudaFuncSetAttribute(k_core<kTeams, kDit>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int nframes = 148 * kTeams * 110;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int i = 0; i < 2; ++i) k_core<kTeams, kDit><<<148, 64 * kTeams, smem>>>(in, out, tw, nframes);
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) k_core<kTeams, kDit><<<148, 64 * kTeams, smem>>>(in, out, tw, nframes);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  cudaError_t e = cudaGetLastError();
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_core<kTeams, kDit>);
  printf("%-22s regs %3d local %4zu  %.3f ms per launch  %.3e frames/s (5 transforms each)  %s\n", name, fa.numRegs, fa.localSizeBytes, ms / 5, nframes / (ms / 5 * 1e-3), cudaGetErrorString(e));
}
int main() {
  float2 *in, *out, *tw;
  cudaMalloc(&in, 1024 * 4096 * 8); cudaMalloc(&out, 1024 * 4096 * 8); cudaMalloc(&tw, 14 * 64 * 8);
  std::vector<float2> h(1024 * 4096);
  for (size_t i = 0; i < h.size(); ++i) h[i] = make_float2((float)((i * 2654435761u) >> 8 & 0xffff) / 65536.f - 0.5f, (float)((i * 40503u) & 0xffff) / 65536.f - 0.5f);
  cudaMemcpy(in, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  std::vector<float2> t(14 * 64, make_float2(0.8f, 0.6f));
  cudaMemcpy(tw, t.data(), t.size() * 8, cudaMemcpyHostToDevice);
  run<4, 0>("4 teams DIF-DIF", in, out, tw);
  run<4, 1>("4 teams DIF-DIT", in, out, tw);
  run<5, 0>("5 teams DIF-DIF", in, out, tw);
  run<5, 1>("5 teams DIF-DIT", in, out, tw);
  run<6, 0>("6 teams DIF-DIF", in, out, tw);
  run<6, 1>("6 teams DIF-DIT", in, out, tw);
  run<3, 0>("3 teams DIF-DIF", in, out, tw);
  run<2, 0>("2 teams DIF-DIF", in, out, tw);
  return 0;
}
