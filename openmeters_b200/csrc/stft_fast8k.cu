// stft_fast8k.cu — specialised kernel for the reassigned STFT at N = 8192 (BASELINE configs[4]: 96 kHz, hop 2048).
//
// Same mathematics as stft_fast2.cu (packed real forward FFT, fused Hilbert pair step, one inverse, three windowed
// forward FFTs), one size up: every transform is 8192 complex points.  Each of them is split
// radix-2 into TWO independent 4096-point transforms, which the two 256-thread groups of a CTA run side by side on
// the shared radix-16 engine (fft4096.cuh), each in its own padded buffer:
//
//   forward  (DIF): u[n] = z[n] + z[n+4096] -> Z[2a]   (group 0),   v[n] = (z[n] - z[n+4096]) W_8192^n -> Z[2a+1] (group 1)
//   pair step     : Q[k] = cos(th_k) conj(Z[M-k]) + j sin(th_k) Z[k], th_k = 2 pi k / 16384.  The partner of an even
//                   bin is even, of an odd bin odd, so the step never leaves the group (even: a <-> 4096-a, odd:
//                   a <-> 4095-a).
//   inverse  (DIT): E = IFFT(Q[2a]) (group 0), O = IFFT(Q[2a+1]) (group 1); q[m] = E[m] + W^-m O[m],
//                   q[m+4096] = E[m] - W^-m O[m].  Only the centre half q[2048..6144) is used, i.e. each m exactly
//                   once: the groups swap the halves they need through the Y buffer (one CTA barrier).
//   analysis (DIF): the windowed centre samples c[n] w[n] are folded the same way; group 0 ends with the even bins,
//                   group 1 with the odd bins, and the ordered compaction interleaves the two groups' ballots.
//
// One CTA = one SM walks a run of consecutive frames of one lane; the H = 16384 samples of a frame live in a
// shared-memory ring of H + hop floats, and the next hop is fetched by 16-byte async copies into the one slot the
// current frame does not read while the current frame is computed.  The lower halves of the window h and of its derivative
// dh (the upper halves are their mirror images), the twiddle rows and Im(c) stay in shared memory.
// Rows a8-a10 of SURVEY.md §8; spectrogram/processor.rs:313-347,439-488,546-567.
// Packed FP32x2 switches of this translation unit (common.h; measured in profiles/r02b_packed_ab.md): packed complex adds
// only — packed products cost this FMA-pipe-bound kernel 2-13 %.
#ifndef OMB_F32X2_CMUL
#define OMB_F32X2_CMUL 0
#endif
#include <algorithm>
#include <cmath>

#include <cstdlib>

#include "device_math.cuh"
#include "fft16.cuh"
#include "fft4096.cuh"
#include "stft.h"

namespace omb {

namespace {

using namespace f4k;

constexpr int kSub = 4096;               // sub-transform length (per group)
constexpr int kN8 = 8192;                // window N = packed complex length M = H / 2
constexpr int kThreads = 2 * kT;
constexpr int kWarps = kT / 32;          // warps per group
constexpr int kWSize = f16::phys_size(kSub);
constexpr int kBinGroups = 9;            // a = t + 256 j, j < 8, and a = 2048 (group 0, t = 0, j = 8)
constexpr int kCntN = kBinGroups * kWarps;

struct Fast8kArgs {
  StftKernelArgs a;
  const float2* tw1;   // global: [15][256] W_4096^{b q}
  const float2* tw2;   // global: [15][16]  W_256^{o q}
  uint32_t frames_per_run, runs_per_lane, ring_len;
  uint64_t chunk;      // > 0: one contiguous range of `chunk` frames of the linearised (lane, frame) sequence per CTA (stft_fast2.cu)
  float norm_ac, norm_dc;  // bin_norm[k] for 0 < k < N/2 and for k in {0, N/2} (window.rs:100-108)
};

struct Smem8k {
  float2 W[2][kWSize];
  float2 tw1[4 * kT];   // rows q = 1, 2, 4, 8 of W_4096^{b q}
  float2 tw2[15 * 16];
  // The periodic cosine-sum windows are symmetric about N/2 (h[N - n] = h[n]) and their spectral derivative antisymmetric
  // (dh[N - n] = -dh[n]), so HALF of each table (n <= N/2) serves the whole window: both fit in shared memory (round 1 kept
  // all of h and 2816 entries of dh here and read the rest of dh through L1: 16 % long-scoreboard stalls).  The f32 tables of
  // the reference are symmetric only up to the rounding of cosf (~6e-8 absolute): mirroring changes the upper half of the
  // window by that much, 1e-9 of a spectral peak — far below the f32 transform's own rounding (checked against float64 in
  // tests/test_gpu_exact.py).
  float hh[kSub + 4];   // h[0 .. N/2]
  float dhh[kSub + 4];  // dh[0 .. N/2]
  float Y[kN8];         // Y[n] = Im c[n]; also the exchange buffer of the inverse radix-2 combine
  unsigned ball[2][kCntN];
  int offs[kCntN + 1];
  float x0_xm[2];
  int pad_[1];
  // float ring[ring_len] follows
};
static_assert(sizeof(Smem8k) % 16 == 0, "ring must stay 16-byte aligned");

__device__ __forceinline__ void async_copy16(float* dst_smem, const float* src_gmem) {
#ifdef OMB_EMU
  for (int i = 0; i < 4; ++i) dst_smem[i] = src_gmem[i];
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src_gmem));
#endif
}
__device__ __forceinline__ void async_commit() {
#ifndef OMB_EMU
  asm volatile("cp.async.commit_group;\n" ::);
#endif
}
__device__ __forceinline__ void async_wait_all() {
#ifndef OMB_EMU
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
#endif
}

// Copies `count` samples (multiple of 4) starting at x into the ring at position p0 (multiple of 4), wrapping at L.
__device__ __forceinline__ void ring_fetch(float* ring, int L, int p0, const float* x, int count) {
  for (int i = 4 * (int)threadIdx.x; i < count; i += 4 * kThreads) {
    int p = p0 + i;
    p -= (p >= L) ? L : 0;
    async_copy16(ring + p, x + i);
  }
}

__device__ __forceinline__ int wrap(int p, int L) { return p - ((p >= L) ? L : 0); }

template <int kTw2, bool kAnyHop>
__global__ void __launch_bounds__(kThreads, 1) k_reassigned_8k(Fast8kArgs fa) {
  OMB_DYN_SMEM(unsigned char, smem_raw);
  Smem8k& sm = *reinterpret_cast<Smem8k*>(smem_raw);
  float* ring = reinterpret_cast<float*>(smem_raw + sizeof(Smem8k));
  const StftKernelArgs& a = fa.a;
  const int tid = threadIdx.x, t = tid & (kT - 1), lane_id = t & 31, warp = t >> 5;
  const int g = __shfl_sync(0xffffffffu, tid >> 8, 0);  // warp-uniform by construction; tells the compiler so
  float2* W = sm.W[g];
  const int hop = (int)a.hop, H = 2 * kN8, L = (int)fa.ring_len;
  const int off = (H - kN8) / 2;
  const uint64_t total_runs = (uint64_t)fa.runs_per_lane * a.n_lanes;
  const ReassignConsts rc{a.bin_hz, a.max_hz, a.inv_2pi, a.inv_hop, a.latency_hops};

  // ---- one-off: tables into shared memory
  for (int i = tid; i < 4 * kT; i += kThreads) {
    const int row = (1 << (i >> 8)) - 1;  // q - 1 for q = 1, 2, 4, 8
    sm.tw1[i] = __ldg(&fa.tw1[row * kT + (i & (kT - 1))]);
  }
  for (int i = tid; i < 15 * 16; i += kThreads) sm.tw2[i] = __ldg(&fa.tw2[i]);
  for (int i = tid; i <= kSub; i += kThreads) {
    sm.hh[i] = __ldg(&a.win[i]);
    sm.dhh[i] = __ldg(&a.dwin[i]);
  }
  Addr ad;
  ad.pA = t + (t >> 4);
  ad.pB = 273 * (t >> 4) + (t & 15);
  ad.pC = 273 * (t & 15) + 17 * (t >> 4);
  const float2* tw1t = sm.tw1 + t;
  const float2* tw2o = sm.tw2 + (t & 15);
  float c8, s8;     // W_8192^t = c8 - j s8 (radix-2 twiddle of the odd group at j = 0)
  sincospif((float)t / (float)kSub, &s8, &c8);
  float cos_t, sin_t;  // th = 2 pi (2 t + g) / 16384: pair-step angle of this thread's bin at j = 0
  sincospif((float)(2 * t + g) / (float)kN8, &sin_t, &cos_t);
  const float sign = (t & 1) ? -1.0f : 1.0f;
  const float ramp0 = (float)t - (float)(kN8 - 1) * 0.5f;  // n - (N-1)/2 at j = 0
  // pair-step partner: even bins a <-> (4096 - a) & 4095, odd bins a <-> 4095 - a
  const int pt = g ? (kT - 1 - t) : ((kT - t) & (kT - 1));
  const int pPartner = 273 * (pt & 15) + 17 * (pt >> 4);
  const bool wrap_j = (g == 0 && t == 0);  // partner element index (16 - j) & 15 instead of 15 - j
  const float* hh_lo = sm.hh + t;            // h[t + n]
  const float* hh_hi = sm.hh + (kSub - t);   // h[t + n + N/2] = h[N/2 - (t + n)] at hh_hi[-n]
  const float* dh_lo = sm.dhh + t;
  const float* dh_hi = sm.dhh + (kSub - t);  // dh[t + n + N/2] = -dh[N/2 - (t + n)]
  const int ts2 = kAnyHop ? 2 * t : 0, ts1 = kAnyHop ? t : 0;
  __syncthreads();

  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  uint64_t gpos = (uint64_t)blockIdx.x * fa.chunk;
  const uint64_t g_end = gpos + fa.chunk < per_lane * a.n_lanes ? gpos + fa.chunk : per_lane * a.n_lanes;
  for (uint64_t run = blockIdx.x; fa.chunk ? gpos < g_end : run < total_runs; run += gridDim.x) {
    uint64_t lane, f_begin, f_end;
    if (fa.chunk) {
      lane = gpos / per_lane;
      const uint64_t fb = gpos % per_lane;
      const uint64_t n = per_lane - fb < g_end - gpos ? per_lane - fb : g_end - gpos;
      f_begin = a.first_frame + fb;
      f_end = f_begin + n;
      gpos += n;
    } else {
      lane = run / fa.runs_per_lane;
      f_begin = a.first_frame + (run % fa.runs_per_lane) * (uint64_t)fa.frames_per_run;
      f_end = (f_begin + fa.frames_per_run < a.frames_per_lane) ? f_begin + fa.frames_per_run : a.frames_per_lane;
    }
    const float* x = a.lanes + lane * a.lane_stride;
    int r0 = 0;  // ring position of the current frame's first sample
    ring_fetch(ring, L, 0, x + f_begin * (uint64_t)hop, H);
    async_commit();
    for (uint64_t f = f_begin; f < f_end; ++f) {
      async_wait_all();
      __syncthreads();  // ring holds frame f; everybody is done with frame f - 1
      if (f + 1 < f_end) ring_fetch(ring, L, wrap(r0 + H, L), x + f * (uint64_t)hop + (uint64_t)H, hop);  // the one free slot
      async_commit();

      // The frame occupies ring positions [r0, r0 + H) mod L: offsets below `split` are reached from lo, the rest
      // from hi = lo - L (for hops that are multiples of 512 split is one too, and a 512-sample block never straddles the wrap).
      const float* lo = ring + r0;
      const float* hi = lo - L;
      const int split = L - r0;
      // hops that are not multiples of 512 (the UI's N/32 ... N/128; always multiples of 4): a 512-sample block may
      // straddle the wrap, so the side is chosen per thread (kAnyHop; ts2 / ts1 are compile-time 0 for the aligned hops)
      float2 v[16];
      // ---- F: z[n] = x[2n] + j x[2n+1]; group 0: z[n] + z[n+4096], group 1: (z[n] - z[n+4096]) W_8192^n, n = t + 256 j
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float2 za = *reinterpret_cast<const float2*>((512 * j + ts2 < split ? lo : hi) + 512 * j + 2 * t);
        const float2 zb = *reinterpret_cast<const float2*>((512 * j + kN8 + ts2 < split ? lo : hi) + 512 * j + kN8 + 2 * t);
        if (g == 0) {
          v[j] = f16::cadd2(za, zb);
        } else {
          const float cj = f16::kCos32[j], sj = f16::kSin32[j];
          v[j] = f16::mul_cs<false>(f16::csub2(za, zb), c8 * cj - s8 * sj, s8 * cj + c8 * sj);
        }
      }
      fft_forward<f16::kAll, kTw2>(v, W, tw1t, tw2o, ad, g);
#pragma unroll
      for (int q = 0; q < 16; ++q) W[ad.pC + q] = v[q];
      if (tid == 0) {
        sm.x0_xm[0] = v[0].x + v[0].y;   // X[0]
        sm.x0_xm[1] = v[0].x - v[0].y;   // X[H/2]
      }
      group_sync(g);
      // ---- X: Q[k] = cos(th_k) conj(Z[M-k]) + j sin(th_k) Z[k], k = 2 (t + 256 j) + g, th_k = th_t + 2 pi j / 32
      {
        const float2* wp = W + pPartner;
        float2 zp[16];
        if (wrap_j) {
#pragma unroll
          for (int j = 0; j < 16; ++j) zp[j] = wp[(16 - j) & 15];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) zp[j] = wp[15 - j];
        }
        group_sync(g);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float cj = f16::kCos32[j], sj = f16::kSin32[j];
          const float ck = cos_t * cj - sin_t * sj;
          const float sk = sin_t * cj + cos_t * sj;
          const float2 z = v[j];
          v[j] = make_float2(ck * zp[j].x - sk * z.y, sk * z.x - ck * zp[j].y);
        }
        if (tid == 0) v[0] = make_float2(0.0f, 0.0f);
      }
      // ---- I: inverse (DIT) of this group's half; then the radix-2 combine through Y
      {
        f16::dft16<true>(v);
        float2* wc = W + ad.pC;
#pragma unroll
        for (int q = 0; q < 16; ++q) wc[q] = v[q];
        group_sync(g);
        float2* wb = W + ad.pB;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = wb[17 * j];
        twiddle15<true, kTw2, false>(v, tw2o, 16);
        f16::dft16<true>(v);
#pragma unroll
        for (int q = 0; q < 16; ++q) wb[17 * q] = v[q];
        group_sync(g);
        const float2* wa = W + ad.pA;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = wa[273 * j];
        twiddle15<true, 1, true>(v, tw1t, kT);
        f16::dft16<true>(v);
        // v[j] = E[m] (group 0) / O[m] (group 1), m = t + 256 j.  y2[i] = q[2048 + i]:
        //   group 0 produces q[m + 4096] = E[m] - W^-m O[m] for m <  2048 -> y2[m + 2048]
        //   group 1 produces q[m]        = E[m] + W^-m O[m] for m >= 2048 -> y2[m - 2048]
        float2* y2 = reinterpret_cast<float2*>(sm.Y) + t;
        if (g == 0) {
#pragma unroll
          for (int j = 8; j < 16; ++j) y2[kT * (j - 8)] = v[j];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float cj = f16::kCos32[j], sj = f16::kSin32[j];
            v[j] = f16::mul_cs<true>(v[j], c8 * cj - s8 * sj, s8 * cj + c8 * sj);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) y2[kT * (j + 8)] = v[j];
        }
        __syncthreads();
        if (g == 0) {
#pragma unroll
          for (int j = 0; j < 8; ++j) y2[kT * (j + 8)] = f16::csub2(v[j], y2[kT * (j + 8)]);
        } else {
#pragma unroll
          for (int j = 8; j < 16; ++j) y2[kT * (j - 8)] = f16::cadd2(y2[kT * (j - 8)], v[j]);
        }
        __syncthreads();
      }
      // ---- G: three windowed transforms of c[n] = (M x[off+n] + bias) + j Y[n], folded radix-2 like F
      const float bias = sign * 0.5f * sm.x0_xm[1] - 0.5f * sm.x0_xm[0];
      float2 S[kBinGroups];
      float nd[kBinGroups];
#pragma unroll 1
      for (int wsel = 0; wsel < 3; ++wsel) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = kT * j;  // + t
          float wa_, wb_;
          if (wsel == 1) {
            wa_ = dh_lo[n];
            wb_ = -dh_hi[-n];
          } else {
            wa_ = hh_lo[n];
            wb_ = hh_hi[-n];
          }
          if (wsel == 2) {  // t*h window, processor.rs:601-608
            wa_ *= ramp0 + (float)n;
            wb_ *= ramp0 + (float)(n + kSub);
          }
          const float xa = fmaf((float)kN8, ((off + n + ts1 < split ? lo : hi) + off + n)[t], bias);
          const float xb = fmaf((float)kN8, ((off + n + kSub + ts1 < split ? lo : hi) + off + n + kSub)[t], bias);
          const float2 ca = make_float2(xa * wa_, sm.Y[t + n] * wa_);
          const float2 cb = make_float2(xb * wb_, sm.Y[t + n + kSub] * wb_);
          if (g == 0) {
            v[j] = f16::cadd2(ca, cb);
          } else {
            const float cj = f16::kCos32[j], sj = f16::kSin32[j];
            v[j] = f16::mul_cs<false>(f16::csub2(ca, cb), c8 * cj - s8 * sj, s8 * cj + c8 * sj);
          }
        }
        fft_forward<f16::kFirst9, kTw2>(v, W, tw1t, tw2o, ad, g);
        if (wsel == 0) {
#pragma unroll
          for (int j = 0; j < kBinGroups; ++j) S[j] = v[j];
        } else if (wsel == 1) {
#pragma unroll
          for (int j = 0; j < kBinGroups; ++j) nd[j] = v[j].y * S[j].x - v[j].x * S[j].y;
        }
        group_sync(g);  // pass-3 reads done before the next pass-1 stores
      }
      // ---- R: reassignment + ordered compaction; bin = 2 (t + 256 j) + g, order (j, t, g)
      omb_spectrogram_point pts[kBinGroups];
      int rank[kBinGroups];
      unsigned keep = 0;
      const unsigned lt_mask = (1u << lane_id) - 1u;
#pragma unroll
      for (int j = 0; j < kBinGroups; ++j) {
        const int bin = 2 * (t + kT * j) + g;
        const float norm = (bin == 0 || bin == kN8 / 2) ? fa.norm_dc : fa.norm_ac;
        const bool k = reassign_bin_nd(S[j], nd[j], v[j], norm, bin, rc, &pts[j]) & (j < 8 || wrap_j);
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (lane_id == 0) sm.ball[g][j * kWarps + warp] = m;
        if (k) keep |= 1u << j;
        rank[j] = __popc(m & lt_mask);
      }
      __syncthreads();
      {
        const unsigned other_mask = g ? (lt_mask | (1u << lane_id)) : lt_mask;  // an odd bin follows its even twin
#pragma unroll
        for (int j = 0; j < kBinGroups; ++j) rank[j] += __popc(sm.ball[g ^ 1][j * kWarps + warp] & other_mask);
      }
      if (tid < 32) {
        int c[3], tot = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int idx = lane_id * 3 + i;
          c[i] = idx < kCntN ? __popc(sm.ball[0][idx]) + __popc(sm.ball[1][idx]) : 0;
          tot += c[i];
        }
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int n = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane_id >= o) incl += n;
        }
        int run_off = incl - tot;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
          const int idx = lane_id * 3 + i;
          if (idx < kCntN) sm.offs[idx] = run_off;
          run_off += c[i];
        }
        if (lane_id == 31) sm.offs[kCntN] = incl;
      }
      __syncthreads();
      {
        const uint64_t slot = lane * a.frames_per_lane + f;
        float* out = reinterpret_cast<float*>(a.out_points + slot * a.point_stride);
#pragma unroll
        for (int j = 0; j < kBinGroups; ++j)
          if (keep & (1u << j)) {
            float* o = out + 3 * (sm.offs[j * kWarps + warp] + rank[j]);
            o[0] = pts[j].time_offset;
            o[1] = pts[j].freq_hz;
            o[2] = pts[j].power;
          }
        if (tid == 0) a.out_counts[slot] = (uint32_t)sm.offs[kCntN];
      }
      r0 = wrap(r0 + hop, L);
    }
    async_wait_all();
    __syncthreads();
  }
}

uint32_t ring_len_for(uint64_t hop) { return (uint32_t)(2 * (uint64_t)kN8 + hop); }
size_t smem_bytes(uint64_t hop) { return sizeof(Smem8k) + (size_t)ring_len_for(hop) * sizeof(float); }

}  // namespace

bool stft_fast8k_supported(const StftConfig& cfg, const DeviceInfo& dev) {
  if (!cfg.reassign || cfg.window != (uint64_t)kN8 || cfg.zero_pad != 1) return false;
  // multiples of 512 up to 2048, or any multiple of 4 below 512 (per-thread ring wrap; the async copies need hop % 4 == 0)
  if (cfg.hop > 2048 || cfg.hop < 4 || (cfg.hop % 4) != 0 || (cfg.hop >= 512 && (cfg.hop % 512) != 0)) return false;
  return dev.max_smem_optin == 0 || smem_bytes(cfg.hop) <= (size_t)dev.max_smem_optin;
}

int stft_fast8k_prepare(StftPlan& plan) {
  // twiddle tables of the 4096-point engine: [15*256] W_4096^{b q} | [15*16] W_256^{o q}
  std::vector<float2> tab(15 * kT + 15 * 16);
  const double tau = 6.28318530717958647692;
  for (int q = 1; q < 16; ++q)
    for (int b = 0; b < kT; ++b) {
      const double ang = -tau * (double)((b * q) % kSub) / (double)kSub;
      tab[(q - 1) * kT + b] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  for (int q = 1; q < 16; ++q)
    for (int o = 0; o < 16; ++o) {
      const double ang = -tau * (double)((o * q) % 256) / 256.0;
      tab[15 * kT + (q - 1) * 16 + o] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  OMB_TRY(plan.d_fast_tables.upload(tab, plan.stream));
  const int smem = (int)smem_bytes(plan.cfg.hop);
  auto k_aligned = k_reassigned_8k<0, false>;
  auto k_any = k_reassigned_8k<0, true>;
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_aligned, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_any, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  return OMB_OK;
}

int launch_stft_fast8k(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s) {
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  if (per_lane == 0 || a.n_lanes == 0) return OMB_OK;
  if ((reinterpret_cast<uintptr_t>(a.lanes) & 15u) != 0 || (a.lane_stride % 4) != 0 || (a.first_frame * a.hop) % 4 != 0)
    return fail(OMB_ERR_INVALID, "specialised STFT kernel needs 16-byte aligned lanes (pointer and lane_stride % 4 == 0)");
  Fast8kArgs fa{};
  fa.a = a;
  fa.tw1 = plan.d_fast_tables.ptr;
  fa.tw2 = fa.tw1 + 15 * kT;
  fa.ring_len = ring_len_for(a.hop);
  fa.norm_ac = plan.h_norm.size() > 1 ? plan.h_norm[1] : plan.h_norm[0];
  fa.norm_dc = plan.h_norm[0];
  const uint64_t ctas = (uint64_t)std::max(plan.dev.sm_count, 1);
  uint64_t run = 64;  // long enough to amortise the ring prime (H samples vs hop per frame)
  if (a.hop < 512) run *= 512 / a.hop;  // small hops: the prime is worth 16384 / hop frames of new samples
  while (run > 8 && ((per_lane + run - 1) / run) * a.n_lanes < ctas * 6) run >>= 1;
  fa.frames_per_run = (uint32_t)std::min<uint64_t>(run, per_lane);
  fa.runs_per_lane = (uint32_t)((per_lane + fa.frames_per_run - 1) / fa.frames_per_run);
  const uint64_t total_runs = (uint64_t)fa.runs_per_lane * a.n_lanes;
  unsigned grid = (unsigned)std::min<uint64_t>(total_runs, ctas);
  // Large batches: one contiguous range of frames per CTA, primed once per lane touched (stft_fast2.cu; OMB_FAST8K_CONTIG=0: round-robin runs)
  static const bool contig_env = [] { const char* e = getenv("OMB_FAST8K_CONTIG"); return !(e && e[0] == '0'); }();
  const uint64_t total_frames = per_lane * a.n_lanes;
  fa.chunk = 0;
  if (contig_env && total_frames >= ctas * 48) {
    fa.chunk = ((total_frames + ctas - 1) / ctas + 0) / 1 * 1;
    grid = (unsigned)((total_frames + fa.chunk - 1) / fa.chunk);
  }
  auto k_aligned = k_reassigned_8k<0, false>;
  auto k_any = k_reassigned_8k<0, true>;
  if ((a.hop % 512) == 0) {
    OMB_LAUNCH(k_aligned, dim3(grid), dim3(kThreads), smem_bytes(a.hop), s, fa);
  } else {
    OMB_LAUNCH(k_any, dim3(grid), dim3(kThreads), smem_bytes(a.hop), s, fa);
  }
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
