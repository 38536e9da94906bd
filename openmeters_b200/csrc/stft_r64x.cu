// stft_r64x.cu — the reassigned STFT at N = 16384 (the largest size the settings UI offers, ui/settings.rs:146-147) on chip.
//
// Round 1 left this size to the global-scratch generic tier (2.5e4 frames/s: no faster than the CPU): a frame is 32768 samples
// and every one of its five transforms has 16384 complex points = 128 KB, so neither the ping-pong shared-memory tier nor the
// radix-16 engine (two frames of 4096 points in flight) can hold it.  The machinery of stft_r64.cu can: 64 complex values per
// thread in registers, ONE exchange buffer of 130 KB, everything that must survive a transform parked in tensor memory.
//
//   * One CTA of 256 threads per SM works on one frame; 16384 = 64 x 64 x 4: pass A = a radix-64 butterfly (fft64.cuh) over the
//     stride-256 subsequence a thread holds, pass B = a radix-64 butterfly inside each 256-point sub-problem (four threads per
//     sub-problem, stride 4), pass C = sixteen radix-4 butterflies across those four threads.  The exchange between B and C stays
//     inside a quad of one warp (__syncwarp), so a transform has two CTA barriers.
//   * The frame arrives by ONE 128 KB bulk copy of the TMA engine in the exchange buffer; the centre samples, the analysis input
//     c[n] (built once per frame: Im c crosses shared memory through the idle buffer), the h-window spectrum S and the cross term
//     nd wait in the thread's own TMEM columns (tmem_park.cuh) — the same column map as stft_r64.cu, because a thread owns 64
//     values and 33 bins there too.
//   * After pass C thread (k1, c) = (tid >> 2, tid & 3) owns bins tau + 256 j, tau = k1 + 64 c: not the lane order the ordered
//     column needs, so the epilogue stages the points by bin in the (then idle) buffer and a second pass in natural thread order
//     compacts them (ballot ranks, 33 x 8 warp counts, one scan).
//   * Window tables (2 x 64 KB) stay in global memory (L2-resident, coalesced); twiddles: 14 per thread and pass from shared
//     memory, the other 49 are products of two (as in stft_r64.cu).
// Same mathematics as the other reassigned kernels: packed real forward transform, fused Hilbert pair step, one
// inverse (run as the forward transform on conjugated data), three windowed transforms, Auger-Flandrin offsets
// (spectrogram/processor.rs:439-608).  Any hop that is a multiple of 4.
#ifndef OMB_F32X2_CMUL
#define OMB_F32X2_CMUL 0
#endif
#include <cstdlib>

#include "async_copy.cuh"
#include "device_math.cuh"
#include "fft64.cuh"
#include "stft.h"
#include "tmem_park.cuh"

namespace omb {

namespace {

constexpr int kM = 16384;                // complex points per transform = window length N
constexpr int kH = 2 * kM;               // samples per frame (Hilbert block)
constexpr int kOff = (kH - kM) / 2;      // first centre sample
constexpr int kT = 256;                  // threads
constexpr int kRS = 260;                 // row stride of exchange 1 (float2): [k1][t], read back at stride 4
constexpr int kRSb = 65;                 // exchange 2: [k1][a][k2] at k1 * 260 + a * 65 + k2;  pair rows: tau * 65 + j + 4 (tau >> 6)
constexpr int kWSize = 64 * kRS + 16;    // float2 elements (pair rows reach 65 * 255 + 63 + 12)
constexpr int kGroups = 33;              // bins tau + 256 j, j < 32, and bin 8192 (tau = 0, j = 32)
constexpr int kWarps = kT / 32;
constexpr unsigned kFrameBytes = kH * sizeof(float);
constexpr int kColC = 0, kColS = 128, kColNd = 192, kColSn = 224, kColsPerWarp = 256;  // as in stft_r64.cu

struct R64xArgs {
  StftKernelArgs a;
  const float2* tw1;    // global: [14][256]: rows 0..6 = W_16384^{t q}, q = 1..7; rows 7..13 = W_16384^{8 t q}
  const float2* tw2;    // global: [14][4]:   the same rows of W_256^{a q}, a < 4
  float norm_ac, norm_dc;
};

struct Smem {
  alignas(16) float2 W[kWSize];
  float2 tw1[14 * kT];
  float2 tw2[14 * 4];
  int cnt[(kGroups + 1) * kWarps];
  int offs[(kGroups + 1) * kWarps + 1];
  float x0_xm[2];
  alignas(8) uint64_t mbar;
  uint32_t tmem_base;
  uint32_t pad_[1];
};
static_assert(sizeof(float2) * kWSize >= kFrameBytes, "the exchange buffer must hold a whole frame");
static_assert(sizeof(float2) * kWSize >= 12 * (kM / 2 + 1), "the exchange buffer must hold a whole staged column");

__device__ __forceinline__ float2 cmul_s(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// v[q] *= w^q, q = 1..63, from 14 table entries (q = 1..7 and 8, 16, ..., 56) + 49 products of two.
__device__ __forceinline__ void twiddle63(float2 (&v)[64], const float2* tab, int stride) {
  float2 lo[8], hi[8];
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    lo[i] = tab[(i - 1) * stride];
    hi[i] = tab[(6 + i) * stride];
  }
#pragma unroll
  for (int q = 1; q < 64; ++q) {
    const int a = q & 7, b = q >> 3;
    const float2 w = b == 0 ? lo[a] : (a == 0 ? hi[b] : cmul_s(lo[a], hi[b]));
    v[q] = f16::mul_tw<false>(v[q], w);
  }
}

// Passes A and B of the forward transform and the loads of pass C.  On entry the thread holds elements ta + 256 j (ta: the residue
// class it owns, any bijection thread -> class); on return v[d + 16 a] holds the pass-C input (a < 4, d < 16) of the radix-4
// butterfly whose outputs k3 = 0..3 are the bins tau + 256 (d + 16 k3), tau = (tid >> 2) + 64 (tid & 3).
__device__ __forceinline__ void transform_ab(float2 (&v)[64], Smem& sm, int ta, int tid) {
  f64pt::dft64<false>(v);
  twiddle63(v, sm.tw1 + ta, kT);
  __syncthreads();  // every earlier read of W (frame samples, partner rows, staging, the previous transform) is done
  float2* wr = sm.W + ta;
#pragma unroll
  for (int q = 0; q < 64; ++q) wr[q * kRS] = v[q];
  __syncthreads();
  const int k1 = tid >> 2, a = tid & 3;
  const float2* rd = sm.W + k1 * kRS + a;
#pragma unroll
  for (int b = 0; b < 64; ++b) v[b] = rd[4 * b];
  f64pt::dft64<false>(v);
  twiddle63(v, sm.tw2 + a, 4);
  __syncthreads();  // all of exchange 1 has been read: exchange 2 reuses the buffer
  float2* wq = sm.W + k1 * kRS + a * kRSb;
#pragma unroll
  for (int k2 = 0; k2 < 64; ++k2) wq[k2] = v[k2];
  __syncwarp();  // exchange 2 stays inside the quad (k1, 0..3) of one warp
  const float2* rq = sm.W + k1 * kRS + a;  // this thread's c = a
#pragma unroll
  for (int d = 0; d < 16; ++d)
#pragma unroll
    for (int s = 0; s < 4; ++s) v[d + 16 * s] = rq[s * kRSb + 4 * d];
}
template <int kPrune>
__device__ __forceinline__ void transform_c(float2 (&v)[64]) {
  f64pt::dit_final<false, kPrune>(v);  // sixteen radix-4 butterflies over (v[d], v[d + 16], v[d + 32], v[d + 48]); pruned like stft_r64.cu
}

template <int N>
__device__ __forceinline__ void park_st(uint32_t tcol, int col, const float (&r)[N]) {
  tmem_st<N>(tcol + col, r);
  tmem_wait_st();
}
template <int N>
__device__ __forceinline__ void park_ld(uint32_t tcol, int col, float (&r)[N]) {
  tmem_ld<N>(tcol + col, r);
}

__global__ void __launch_bounds__(kT, 1) k_reassigned_r64x(R64xArgs ra) {
  OMB_DYN_SMEM(unsigned char, smem_raw);
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const StftKernelArgs& a = ra.a;
  const int tid = threadIdx.x, lane_id = tid & 31, warp = tid >> 5;
  const ReassignConsts rc{a.bin_hz, a.max_hz, a.inv_2pi, a.inv_hop, a.latency_hops};

  for (int i = tid; i < 14 * kT; i += kT) sm.tw1[i] = __ldg(&ra.tw1[i]);
  if (tid < 14 * 4) sm.tw2[tid] = __ldg(&ra.tw2[tid]);
  if (tid < 32) tmem_alloc(&sm.tmem_base);
  if (tid == 0) mbar_init(&sm.mbar, 1);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tcol = tmem_addr(sm.tmem_base, (uint32_t)(tid >> 7) * kColsPerWarp);

  // per-thread constants
  const int tau = (tid >> 2) + 64 * (tid & 3);  // bins tau + 256 j after a transform
  float cos_t, sin_t;                            // th_tau = 2 pi tau / H
  sincospif((float)tau / (float)kM, &sin_t, &cos_t);
  const float sign = (tid & 1) ? -1.0f : 1.0f;   // (-1)^n of the analysis point n = tid + 256 j
  const float ramp0 = (float)tid - (float)(kM - 1) * 0.5f;
  const int ptau = (kT - tau) & (kT - 1);        // owner class of the mirror bins M - k
  auto row_of = [](int r) { return r * kRSb + 4 * (r >> 6); };  // pair rows: skewed so that rows r and r + 64 use different banks

  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total = per_lane * a.n_lanes;
  auto frame_src = [&](uint64_t gi) {
    const uint64_t l = gi / per_lane, f = a.first_frame + gi % per_lane;
    return a.lanes + l * a.lane_stride + f * (uint64_t)a.hop;
  };
  uint64_t g = blockIdx.x;
  unsigned phase = 0;
  if (g < total && tid == 0) {
    mbar_expect_tx(&sm.mbar, kFrameBytes);
    bulk_g2s(sm.W, frame_src(g), kFrameBytes, &sm.mbar);
  }

  for (; g < total; g += gridDim.x) {
    const uint64_t lane = g / per_lane, f = a.first_frame + g % per_lane;
    mbar_wait(&sm.mbar, phase);
    phase ^= 1u;
    float2 v[64];
    {
      const float* wf = reinterpret_cast<const float*>(sm.W);
      float xc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) xc[j] = wf[kOff + tid + kT * j];
      park_st<64>(tcol, kColC, xc);
      const float2* wz = sm.W + tid;  // F input: z[n] = x[2n] + j x[2n+1], n = tid + 256 j
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = wz[kT * j];
    }
#pragma unroll 1
    for (int tr = 0; tr < 5; ++tr) {
      int ta = tid;
      if (tr == 1) {
        // ---- X: conj(Q[k]), Q[k] = cos(th_k) conj(Z[M-k]) + j sin(th_k) Z[k], k = tau + 256 j, th_k = th_tau + 2 pi j / 128
        __syncthreads();
        float2* row = sm.W + row_of(tau);
#pragma unroll
        for (int j = 0; j < 64; ++j) row[j] = v[j];
        if (tau == 0) {
          sm.x0_xm[0] = v[0].x + v[0].y;
          sm.x0_xm[1] = v[0].x - v[0].y;
        }
        __syncthreads();
        // partner of k = tau + 256 j: row 256 - tau, column 63 - j; for tau = 0 row 0, column 64 - j (j = 0: don't-care, DC is zeroed)
        const float2* prow = tau == 0 ? sm.W + 1 : sm.W + row_of(ptau);
        constexpr int kPc = 8;
        float2 zp[2][kPc];
#pragma unroll
        for (int i = 0; i < kPc; ++i) zp[0][i] = prow[63 - i];
#pragma unroll
        for (int c = 0; c < 64 / kPc; ++c) {
          if (c + 1 < 64 / kPc) {
#pragma unroll
            for (int i = 0; i < kPc; ++i) zp[(c + 1) & 1][i] = prow[63 - kPc * (c + 1) - i];
          }
#pragma unroll
          for (int i = 0; i < kPc; ++i) {
            const int j = kPc * c + i;
            const float cj = f64pt::kCos128[j], sj = f64pt::kSin128[j];
            const float ck = cos_t * cj - sin_t * sj;
            const float sk = sin_t * cj + cos_t * sj;
            const float2 z = v[j], p = zp[c & 1][i];
            v[j] = make_float2(ck * p.x - sk * z.y, ck * p.y - sk * z.x);
          }
        }
        if (tau == 0) v[0] = make_float2(0.0f, 0.0f);
        ta = tau;
      } else if (tr >= 2) {
        // ---- G input: c[n] w[n], n = tid + 256 j, w = h, dh, t*h (processor.rs:601-608, formed on the fly, bit-identical)
        float c[128];
        park_ld<128>(tcol, kColC, c);
        const float* tab = (tr == 3 ? a.dwin : a.win) + tid;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          float wv = __ldg(&tab[kT * j]);
          if (tr == 4) wv *= ramp0 + (float)(kT * j);
          v[j] = f16::cscale2(make_float2(c[2 * j], c[2 * j + 1]), wv);
        }
      }
      transform_ab(v, sm, ta, tid);
      if (tr == 0) {
        transform_c<f16::kAll>(v);
      } else if (tr == 1) {
        transform_c<f16::kMid8>(v);    // outputs 16..47: the centre half
      } else {
        transform_c<f16::kFirst9>(v);  // outputs 0..32: bins <= Nyquist
      }
      if (tr == 1) {
        // ---- centre half of the inverse: y[m] = conj(v), m = tau + 256 j, j = 16..47 -> staging as float2[m - 4096]; then every
        //      thread collects Im c of its analysis points n = tid + 256 j and parks c[n] = (M x[off + n] + bias) + j Im c[n]
        __syncthreads();
        float2* y2 = sm.W + tau;
#pragma unroll
        for (int q = 16; q < 48; ++q) y2[kT * (q - 16)] = make_float2(v[q].x, -v[q].y);
        __syncthreads();
        const float bias = sign * 0.5f * sm.x0_xm[1] - 0.5f * sm.x0_xm[0];
        const float* yf = reinterpret_cast<const float*>(sm.W) + tid;
        float c[128];
        {
          float xc[64];
          park_ld<64>(tcol, kColC, xc);
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            c[2 * j] = fmaf((float)kM, xc[j], bias);
            c[2 * j + 1] = yf[kT * j];
          }
        }
        park_st<128>(tcol, kColC, c);
      } else if (tr == 2) {
        float s[64];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s[2 * j] = v[j].x;
          s[2 * j + 1] = v[j].y;
        }
        park_st<64>(tcol, kColS, s);
        const float sn[2] = {v[32].x, v[32].y};
        park_st<2>(tcol, kColSn, sn);
      } else if (tr == 3) {
        float s[64], sn[4], nd[32], ndn[1];
        park_ld<64>(tcol, kColS, s);
        park_ld<4>(tcol, kColSn, sn);
#pragma unroll
        for (int j = 0; j < 32; ++j) nd[j] = v[j].y * s[2 * j] - v[j].x * s[2 * j + 1];
        ndn[0] = v[32].y * sn[0] - v[32].x * sn[1];
        park_st<32>(tcol, kColNd, nd);
        park_st<1>(tcol, kColSn + 2, ndn);
      }
    }
    // ---- R: reassigned points of this thread's bins, staged by bin in W (power < 0 marks a dropped bin)
    __syncthreads();  // every quad has finished reading exchange 2
    {
      float* stage = reinterpret_cast<float*>(sm.W);
      float nd[32], sn[4];
      park_ld<32>(tcol, kColNd, nd);
      park_ld<4>(tcol, kColSn, sn);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float s[32];
        if (c < 2) {
          park_ld<32>(tcol, kColS + 32 * c, s);
        } else {
          s[0] = sn[0];
          s[1] = sn[1];
        }
#pragma unroll
        for (int i = 0; i < (c < 2 ? 16 : 1); ++i) {
          const int j = 16 * c + i;
          const int bin = tau + kT * j;
          const float norm = (bin == 0 || j == 32) ? ra.norm_dc : ra.norm_ac;
          const float ndj = j < 32 ? nd[j & 31] : sn[2];
          omb_spectrogram_point pt;
          const bool k = reassign_bin_nd(make_float2(s[2 * i], s[2 * i + 1]), ndj, v[j], norm, bin, rc, &pt);
          if (j < 32 || tau == 0) {
            float* o = stage + 3 * bin;
            o[0] = pt.time_offset;
            o[1] = pt.freq_hz;
            o[2] = k ? pt.power : -1.0f;
          }
        }
      }
    }
    __syncthreads();
    // ---- ordered compaction in natural order: thread tid, group q -> bin tid + 256 q (thread 0 also bin 8192)
    {
      const float* stage = reinterpret_cast<const float*>(sm.W);
      unsigned keep_lo = 0, keep_hi = 0;
#pragma unroll
      for (int q = 0; q < kGroups; ++q) {
        const int bin = tid + kT * q;
        const bool k = (q < 32 || tid == 0) && stage[3 * (q < 32 ? bin : kM / 2) + 2] >= 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (lane_id == 0) sm.cnt[q * kWarps + warp] = __popc(m);
        if (k) {
          if (q < 32) keep_lo |= 1u << q; else keep_hi = 1u;
        }
      }
      __syncthreads();
      if (warp == 0) {  // exclusive scan of the 33 x 8 warp counts (order: group, warp), 9 entries per lane
        constexpr int kN = kGroups * kWarps, kPer = (kN + 31) / 32;
        int c[kPer], tot = 0;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int idx = lane_id * kPer + i;
          c[i] = idx < kN ? sm.cnt[idx] : 0;
          tot += c[i];
        }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int n = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane_id >= o) incl += n;
        }
        int run = incl - tot;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int idx = lane_id * kPer + i;
          if (idx < kN) sm.offs[idx] = run;
          run += c[i];
        }
        if (lane_id == 31) sm.offs[kN] = incl;
      }
      __syncthreads();
      const uint64_t slot = lane * a.frames_per_lane + f;
      float* out = reinterpret_cast<float*>(a.out_points + slot * a.point_stride);
      const unsigned lt_mask = (1u << lane_id) - 1u;
#pragma unroll
      for (int q = 0; q < kGroups; ++q) {
        const bool k = q < 32 ? ((keep_lo >> q) & 1u) != 0 : keep_hi != 0;
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (k) {
          const float* src = stage + 3 * (q < 32 ? tid + kT * q : kM / 2);
          float* o = out + 3 * (sm.offs[q * kWarps + warp] + __popc(m & lt_mask));
          o[0] = src[0];
          o[1] = src[1];
          o[2] = src[2];
        }
      }
      if (tid == 0) a.out_counts[slot] = (uint32_t)sm.offs[kGroups * kWarps];
    }
    // ---- the staged column has been read: the next frame may land
    __syncthreads();
    if (tid == 0 && g + gridDim.x < total) {
      fence_async_smem();
      mbar_expect_tx(&sm.mbar, kFrameBytes);
      bulk_g2s(sm.W, frame_src(g + gridDim.x), kFrameBytes, &sm.mbar);
    }
  }
  __syncthreads();
  if (tid < 32) tmem_free(sm.tmem_base);
}

size_t smem_bytes() { return sizeof(Smem); }

}  // namespace

bool stft_r64x_supported(const StftConfig& cfg, const DeviceInfo& dev) {
  if (!cfg.reassign || cfg.window != (uint64_t)kM || cfg.zero_pad != 1) return false;
  if (cfg.hop < 4 || (cfg.hop % 4) != 0) return false;  // bulk copies start at f * hop floats: 16-byte aligned
  if (dev.cc_major != 0 && dev.cc_major < 10) return false;  // tcgen05 / TMEM
  return dev.max_smem_optin == 0 || smem_bytes() <= (size_t)dev.max_smem_optin;
}

int stft_r64x_prepare(StftPlan& plan) {
  std::vector<float2> tab(14 * kT + 14 * 4);
  const double tau = 6.28318530717958647692;
  for (int i = 0; i < 14; ++i) {
    const int q = i < 7 ? i + 1 : 8 * (i - 6);
    for (int t = 0; t < kT; ++t) {
      const double ang = -tau * (double)((t * q) % kM) / (double)kM;
      tab[i * kT + t] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    for (int a = 0; a < 4; ++a) {
      const double ang = -tau * (double)((a * q) % 256) / 256.0;
      tab[14 * kT + i * 4 + a] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  }
  OMB_TRY(plan.d_r64_tables.upload(tab, plan.stream));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_r64x, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
  return OMB_OK;
}

int launch_stft_r64x(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s) {
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  if (per_lane == 0 || a.n_lanes == 0) return OMB_OK;
  if ((reinterpret_cast<uintptr_t>(a.lanes) & 15u) != 0 || (a.lane_stride % 4) != 0)
    return fail(OMB_ERR_INVALID, "specialised STFT kernel needs 16-byte aligned lanes (pointer and lane_stride % 4 == 0)");
  R64xArgs ra{};
  ra.a = a;
  ra.tw1 = plan.d_r64_tables.ptr;
  ra.tw2 = ra.tw1 + 14 * kT;
  ra.norm_ac = plan.h_norm.size() > 1 ? plan.h_norm[1] : plan.h_norm[0];
  ra.norm_dc = plan.h_norm[0];
  const uint64_t total = per_lane * a.n_lanes;
  const unsigned grid = (unsigned)std::min<uint64_t>(total, (uint64_t)std::max(plan.dev.sm_count, 1));
  OMB_LAUNCH(k_reassigned_r64x, dim3(grid), dim3(kT), smem_bytes(), s, ra);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
