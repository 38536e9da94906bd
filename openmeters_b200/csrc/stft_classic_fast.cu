// stft_classic_fast.cu — warp-per-frame classic spectrogram column for N = 1024, zero padding 1
// (BASELINE configs[0]; spectrogram/processor.rs:349-381, :103-108; util/audio/window.rs:66-88).
//
// One warp owns one frame from the first load to the last u16 store; nothing but __syncwarp separates its
// phases, so 32 frames per SM are in flight independently and the memory phases of some overlap the arithmetic
// of others.  The real 1024-point transform is the complex 512-point transform of z[n] = r[2n] + j r[2n+1]:
//
//   load     : half-warp s (16 lanes, r = lane & 15) reads z[n], z[n+256] for n = r + 16 j and forms the first
//              radix-2 DIF stage on the fly: s = 0 keeps z[n] + z[n+256] (even bins), s = 1 keeps
//              (z[n] - z[n+256]) * W_512^n (odd bins).  DC removal and the window are applied while loading.
//   256-point: each half-warp runs its own 256-point FFT as two in-register radix-16 butterflies (fft16.cuh) with a
//              16 x 16 transpose through a warp-private shared-memory tile in between.
//   split    : bins k and 1024/2 - k share the pair (Z[k], Z[512-k]): X[k] = E + W_1024^k O and
//              |X[512-k]| = |E - W_1024^k O|, so every lane converts 16 bins to dB / u16 from 8 pair loads.
//
// Algorithmic bytes per frame: hop * 4 read + 513 * 2 written (3074 B at hop 512).
// Packed FP32x2 switches of this translation unit (common.h; measured in profiles/r02b_packed_ab.md): everything packed
// (adds, rotations, products) — a small gain here (+0.5 ... +1.6 %).
#ifndef OMB_F32X2_MUL
#define OMB_F32X2_MUL 1
#endif
#ifndef OMB_F32X2_ROT
#define OMB_F32X2_ROT 1
#endif
#include "device_math.cuh"
#include "fft16.cuh"
#include "stft.h"

#include <cstdlib>

namespace omb {

namespace {

using namespace f16;

constexpr int kN = 1024, kM = 512, kWarps = 8, kTile = 16 * 17 * 2 + 8;  // two padded 16 x 16 tiles >= 513 entries

constexpr int kTile1 = 16 * 17 + 8;  // offset of the second half-transform's tile: 8 float2 past a bank period (see below)

struct SmemC {
  float2 w1024[256];      // W_1024^k, k < 256 (the split handles bins k and 512 - k together)
  float2 w512[256];       // first-stage twiddle W_512^n of the odd half-transform
  float2 tw2[15 * 16];    // W_256^{r q}, [(q - 1) * 16 + r]
  float2 win_lo[256];     // (h[2n], h[2n+1]), n < 256
  float2 win_hi[256];     // (h[2n+512], h[2n+513])
  float2 work[kWarps][kTile];
};

// Second version, after the first ncu capture (issue-bound, 1840 instructions per frame) and the second (L1 data-pipe
// wavefronts at 86 % of peak, 14 % of the shared-memory wavefronts bank conflicts):
//   * the half-transform a lane works for (s: even / odd bins) is data, not control flow: the radix-2 sign enters as
//     (xb - m) * (+-1) = fma(xb, sgn, -sgn m), exact, and only the odd half multiplies by W_512^n — no selects, no
//     duplicated predicated code, and the window tables are shared by both halves (broadcast loads);
//   * lanes are interleaved (s = lane & 1, r = lane >> 1) and the second tile sits 8 float2 past a bank period, so the
//     16 lanes of each 64-bit shared-memory phase (8 values of r, both s) always touch 16 distinct bank pairs, and
//     the spectrum store work[lane + 32 p] is contiguous (it was a 2-way conflict with s = lane >> 4);
//   * the factors 1/2 of the real-FFT split are folded into the bin normalisation (exact: powers of two);
//   * Z[0] is mirrored at work[512], so the partner index 512 - k needs no wrap;
//   * dB + u16 packing through classic_code_dev (device_math.cuh): 25 instructions per bin instead of 45.
__global__ void __launch_bounds__(kWarps * 32, 4) k_classic_1024(StftKernelArgs a) {
  OMB_DYN_SMEM(unsigned char, raw);
  SmemC& sm = *reinterpret_cast<SmemC*>(raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 256; i += kWarps * 32) {
    sm.w1024[i] = __ldg(&a.tw_fft[i]);
    sm.w512[i] = __ldg(&a.tw_fft[2 * i]);
    sm.win_lo[i] = make_float2(__ldg(&a.win[2 * i]), __ldg(&a.win[2 * i + 1]));
    sm.win_hi[i] = make_float2(__ldg(&a.win[2 * i + 512]), __ldg(&a.win[2 * i + 513]));
  }
  for (int i = tid; i < 240; i += kWarps * 32) {
    const int idx = 4 * (i / 16 + 1) * (i % 16);  // W_256^{rq} = W_1024^{4rq}; the table holds the upper half circle only
    float2 w = __ldg(&a.tw_fft[idx & 511]);
    if (idx >= 512) w = make_float2(-w.x, -w.y);
    sm.tw2[i] = w;
  }
  __syncthreads();

  const int s = lane & 1, r = lane >> 1;
  float2* work = sm.work[warp];
  float2* tile = work + s * kTile1;
  const float2* wlo = sm.win_lo + r;
  const float2* whi = sm.win_hi + r;
  const float2* wst = sm.w512 + r;
  const float sgn = s ? -1.0f : 1.0f;
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total = per_lane * a.n_lanes;
  const uint64_t stride_items = (uint64_t)gridDim.x * kWarps;
  // |X|^2 = |2X|^2 / 4: the split below works on 2X
  const float norm_edge = 0.25f * __ldg(&a.bin_norm[0]), norm_mid = 0.25f * __ldg(&a.bin_norm[1]);

  for (uint64_t item = (uint64_t)blockIdx.x * kWarps + warp; item < total; item += stride_items) {
    const uint64_t lane_idx = item / per_lane, frame = a.first_frame + item % per_lane;
    const float2* x2 = reinterpret_cast<const float2*>(a.lanes + lane_idx * a.lane_stride + frame * a.hop);

    // mean of the frame (window.rs:76-79)
    float part = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 v = __ldg(&x2[lane + 32 * i]);
      part += v.x + v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    const float mean = part * (1.0f / (float)kN);

    // load + DC removal + window + first radix-2 stage: d = (xa - m) h_lo +- (xb - m) h_hi; odd half: times W_512^n
    float2 v[16];
    const float2* xr = x2 + r;
    const float msgn = -sgn * mean;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float2 xa = __ldg(&xr[16 * j]), xb = __ldg(&xr[16 * j + 256]);
      const float2 wa = wlo[16 * j], wb = whi[16 * j];
      const float2 d = make_float2(fmaf(fmaf(xb.x, sgn, msgn), wb.x, (xa.x - mean) * wa.x),
                                   fmaf(fmaf(xb.y, sgn, msgn), wb.y, (xa.y - mean) * wa.y));
      v[j] = d;
      if (s) v[j] = mul_tw<false>(d, wst[16 * j]);
    }
    dft16<false>(v);  // over j: A[r][q]
    tile[r] = v[0];
#pragma unroll
    for (int q = 1; q < 16; ++q) tile[q * 17 + r] = mul_tw<false>(v[q], sm.tw2[(q - 1) * 16 + r]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = tile[r * 17 + i];  // lane q = r collects A'[0..15][q]
    dft16<false>(v);                                         // over r: Z_s[q + 16 p]
    __syncwarp();
#pragma unroll
    for (int p = 0; p < 16; ++p) work[lane + 32 * p] = v[p];  // Z[2 k' + s], k' = r + 16 p
    if (lane == 0) work[kM] = v[0];                                   // Z[512] := Z[0]
    __syncwarp();

    uint16_t* out = a.out_classic + (lane_idx * a.frames_per_lane + frame) * (uint64_t)(kM + 1);
    const float2* zlo = work + lane;
    const float2* zhi = work + kM - lane;
    const float2* wk = sm.w1024 + lane;
    uint16_t* olo = out + lane;
    uint16_t* ohi = out + kM - lane;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // k = lane + 32 i (0..255), partner bin 512 - k
      const float2 zk = zlo[32 * i], zm = zhi[-32 * i];
      const float2 E = make_float2(zk.x + zm.x, zk.y - zm.y);      // 2 E
      const float2 O = make_float2(zk.y + zm.y, zm.x - zk.x);      // 2 O = (zk - conj zm) / j
      const float2 T = cmul(wk[32 * i], O);
      const float2 Xk = cadd2(E, T), Xm = csub2(E, T);             // |X[512-k]| = |conj(E - T)|
      const float nk = (i == 0 && lane == 0) ? norm_edge : norm_mid;
      olo[32 * i] = classic_code_dev((Xk.x * Xk.x + Xk.y * Xk.y) * nk);
      ohi[-32 * i] = classic_code_dev((Xm.x * Xm.x + Xm.y * Xm.y) * nk);
    }
    if (lane == 0) {  // bin 256 pairs with itself: X = conj(Z[256]); here |Z|^2 is not doubled
      const float2 z = work[256];
      out[256] = classic_code_dev((z.x * z.x + z.y * z.y) * (4.0f * norm_mid));
    }
    __syncwarp();
  }
}

}  // namespace

bool stft_classic_fast_supported(const StftConfig& cfg, const DeviceInfo& dev) {
  if (cfg.reassign || cfg.window != (uint64_t)kN || cfg.zero_pad != 1 || (cfg.hop & 1)) return false;
  if (getenv("OMB_NO_CLASSIC_FAST")) return false;
  return dev.max_smem_optin == 0 || sizeof(SmemC) <= (size_t)dev.max_smem_optin;
}

int stft_classic_fast_prepare(StftPlan&) {
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_classic_1024, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemC)));
  return OMB_OK;
}

int launch_stft_classic_fast(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s) {
  const uint64_t total = (a.frames_per_lane - a.first_frame) * a.n_lanes;
  if (!total) return OMB_OK;
  const uint64_t want = (total + kWarps - 1) / kWarps;
  const int grid = (int)std::min<uint64_t>(want, (uint64_t)std::max(plan.dev.sm_count, 1) * 4);
  OMB_LAUNCH(k_classic_1024, dim3(grid), dim3(kWarps * 32), sizeof(SmemC), s, a);
  return OMB_OK;
}

}  // namespace omb
