// fft_stockham.cuh — block-cooperative Stockham autosort FFT in shared memory (runtime power-of-two length).
//
// Used by the "smem" kernel family (stft_smem.cu, spectrum.cu): any transform length 4 <= M <= 8192 complex
// points ping-pongs between two shared buffers; output is in natural order, so the callers' pair steps /
// per-bin epilogues index frequencies directly.  Radix-4 stages (one radix-2 stage first when log2 M is odd),
// reads at stride M/4 (coalesced, conflict-free), one twiddle load per butterfly (w1; w2 = w1^2, w3 = w1*w2).
#pragma once
#include "device_math.cuh"

namespace omb {

// tw[i * tw_stride] = W_M^i = exp(-2 pi i / M) for i < M/4 (global, read-only; a table of W_{M*tw_stride} works).
// Forward transform, unnormalised.
// Returns the buffer holding the result (a or b). All threads of the block must call it.
__device__ __forceinline__ float2* stockham_fft(float2* a, float2* b, int M, int logM, const float2* __restrict__ tw,
                                                 int tw_stride = 1) {
  const int tid = threadIdx.x, nt = blockDim.x;
  float2* src = a;
  float2* dst = b;
  int p = 1;  // length of the sub-transforms already completed
  if (logM & 1) {  // radix-2 stage
    const int half = M >> 1;
    for (int i = tid; i < half; i += nt) {
      // p == 1: no twiddle
      const float2 u0 = src[i], u1 = src[i + half];
      dst[2 * i] = cadd(u0, u1);
      dst[2 * i + 1] = csub(u0, u1);
    }
    __syncthreads();
    float2* t = src; src = dst; dst = t;
    p = 2;
  }
  const int quarter = M >> 2;
  for (; p < M; p <<= 2) {
    for (int i = tid; i < quarter; i += nt) {
      const int k = i & (p - 1);
      const int j = ((i - k) << 2) + k;
      // alpha = -2 pi k / (4p)  ->  table index k * M / (4p)
      float2 w1 = make_float2(1.0f, 0.0f);
      if (p > 1) w1 = __ldg(&tw[k * (M / (4 * p)) * tw_stride]);
      const float2 w2 = cmul(w1, w1);
      const float2 w3 = cmul(w2, w1);
      const float2 u0 = src[i];
      const float2 u1 = cmul(src[i + quarter], w1);
      const float2 u2 = cmul(src[i + 2 * quarter], w2);
      const float2 u3 = cmul(src[i + 3 * quarter], w3);
      const float2 s0 = cadd(u0, u2), s1 = csub(u0, u2), s2 = cadd(u1, u3);
      const float2 d = csub(u1, u3);
      const float2 s3 = make_float2(d.y, -d.x);  // -j * d
      dst[j] = cadd(s0, s2);
      dst[j + p] = cadd(s1, s3);
      dst[j + 2 * p] = csub(s0, s2);
      dst[j + 3 * p] = csub(s1, s3);
    }
    __syncthreads();
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

}  // namespace omb
