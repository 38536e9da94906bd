// fft4096.cuh — the 4096-point complex FFT engine shared by the specialised kernels (stft_fast2.cu, spectrum_fast.cu).
//
// One 256-thread group transforms 4096 points in place in a padded shared buffer with three radix-16 register passes
// (fft16.cuh); several groups of one CTA run independent transforms and synchronise with their own named barrier.
#pragma once
#include "device_math.cuh"
#include "fft16.cuh"
#include "tmem_park.cuh"

namespace omb {
namespace f4k {

constexpr int kT = 256;  // threads per group

// bar.sync on the group's named barrier (ids 1.. ; 0 is __syncthreads)
__device__ __forceinline__ void group_sync(int g) {
#ifdef OMB_EMU
  omb_emu::named_sync(1 + g, kT);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(1 + g), "r"(kT) : "memory");
#endif
}


// kTw = 0: 15 table loads (shared memory).  kTw = 1: 4 loads (q = 1, 2, 4, 8) + 11 products — trades 11
// shared-memory loads for 44 flops (each derived twiddle is at most two multiplications from a table entry).
template <bool INV, int kTw, bool kCompact>
__device__ __forceinline__ void twiddle15(float2 (&v)[16], const float2* tab, int stride) {
  if (kTw == 0) {
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = f16::mul_tw<INV>(v[q], tab[(q - 1) * stride]);
  } else {
    float2 w[16];
    w[1] = tab[0 * stride];
    w[2] = tab[(kCompact ? 1 : 1) * stride];
    w[4] = tab[(kCompact ? 2 : 3) * stride];
    w[8] = tab[(kCompact ? 3 : 7) * stride];
    w[3] = cmul(w[1], w[2]);
    w[5] = cmul(w[1], w[4]);
    w[6] = cmul(w[2], w[4]);
    w[7] = cmul(w[3], w[4]);
#pragma unroll
    for (int q = 9; q < 16; ++q) w[q] = cmul(w[q - 8], w[8]);
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = f16::mul_tw<INV>(v[q], w[q]);
  }
}

struct Addr {
  int pA, pB, pC;
};

// DIF forward: v holds elements t + 256 j (access A). On return v[j] = X[t + 256 j] (only the kPrune subset).
// kLocal3: pass 3 stays inside the half-warp that ran pass 2 (ad.pC = 273 (t >> 4) + 17 (t & 15)) and thread t ends up with
// bins tf + 256 j, tf = (t >> 4) + 16 (t & 15), which lets the INVERSE that follows replace its first group barrier by a
// __syncwarp (stft_fast2.cu).  The barrier between passes 2 and 3 here stays a group barrier: a warp-level sync is
// sufficient (the block is written and read by one half-warp; racecheck clean) but was measured neutral to slightly slower
// on cfg2 (1.9437e7 -> 1.9393e7 frames/s) and -6 % in stft_fast2k.cu, where the extra per-thread constants spill.
template <int kPrune, int kTw2, bool kLocal3 = false>
__device__ __forceinline__ void fft_forward(float2 (&v)[16], float2* W, const float2* tw1t, const float2* tw2o, const Addr& ad, int g) {
  f16::dft16<false>(v);
  twiddle15<false, 1, true>(v, tw1t, kT);
  float2* wa = W + ad.pA;
#pragma unroll
  for (int q = 0; q < 16; ++q) wa[273 * q] = v[q];
  group_sync(g);
  float2* wb = W + ad.pB;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = wb[17 * j];
  f16::dft16<false>(v);
  twiddle15<false, kTw2, false>(v, tw2o, 16);
#pragma unroll
  for (int q = 0; q < 16; ++q) wb[17 * q] = v[q];
  group_sync(g);
  const float2* wc = W + ad.pC;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = wc[j];
  f16::dft16p<false, kPrune>(v);
}


// ---- twiddles from tensor memory.  The 15 + 15 twiddles a thread uses in passes 1 and 2 depend only on its index, never on the
// data: parked once per kernel in the thread's own TMEM columns (tmem_park.cuh) they cost one tcgen05.ld per pass instead of 4 + 15
// shared-memory loads and the 11 complex products that rebuild the pass-1 set (twiddle15<kTw = 1>) — shared-memory wavefronts are
// the first co-limiter of the radix-16 kernels (profiles/r02f_ncu_full_fast2.json).
struct TwTmem {
  uint32_t t1, t2;  // TMEM addresses: 30 columns each (15 float2, q = 1..15), padded to 32
};
template <bool INV>
__device__ __forceinline__ void twiddle15_tmem(float2 (&v)[16], uint32_t taddr) {
  float w[32];
  tmem_ld<32>(taddr, w);
#pragma unroll
  for (int q = 1; q < 16; ++q) v[q] = f16::mul_tw<INV>(v[q], make_float2(w[2 * (q - 1)], w[2 * (q - 1) + 1]));
}
template <int kPrune, bool kLocal3 = false>
__device__ __forceinline__ void fft_forward_tmem(float2 (&v)[16], float2* W, const TwTmem& tw, const Addr& ad, int g) {
  f16::dft16<false>(v);
  twiddle15_tmem<false>(v, tw.t1);
  float2* wa = W + ad.pA;
#pragma unroll
  for (int q = 0; q < 16; ++q) wa[273 * q] = v[q];
  group_sync(g);
  float2* wb = W + ad.pB;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = wb[17 * j];
  f16::dft16<false>(v);
  twiddle15_tmem<false>(v, tw.t2);
#pragma unroll
  for (int q = 0; q < 16; ++q) wb[17 * q] = v[q];
  group_sync(g);
  const float2* wc = W + ad.pC;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = wc[j];
  f16::dft16p<false, kPrune>(v);
}

}  // namespace f4k
}  // namespace omb
