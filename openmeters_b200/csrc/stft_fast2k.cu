// stft_fast2k.cu — specialised kernel for the reassigned STFT at N = 2048, the product's default analysis size
// (spectrogram/processor.rs:47-59: fft_size 2048, hop 64, Hann, reassignment on).
//
// Same mathematics and the same 4096-point FFT engine as stft_fast2.cu; every 4096-point transform carries TWO
// consecutive frames of one lane, interleaved in time:  a[2n] = A[n], a[2n + 1] = B[n]  gives
//     X[k] = FA[k] + W^k FB[k],   X[k + 2048] = FA[k] - W^k FB[k]          (W = e^{-2 pi i / 4096}, k < 2048)
// and thread t holds X[t + 256 j] for all j, so bins k and k + 2048 sit in the same thread: the two 2048-point spectra
// are separated (and, before the inverse, combined) in registers with 8 complex add/sub pairs and 8 twiddle
// multiplications per thread.  Thread t therefore owns bins t + 256 j, j < 8, of BOTH frames in the frequency domain
// and the samples (t >> 1) + 128 j of frame (t & 1) in the time domain; engine, padded layout, pass structure and the
// pruned inverse (outputs 4..11 = the centre half) are unchanged.  Power-of-two factors of the separation (2 FA) are
// folded into the pair-step coefficients and the bin normalisation (exact).
// One 512-thread CTA = two groups = four consecutive frames per iteration, sharing one staging ring.
// Packed FP32x2 switches of this translation unit (common.h; measured in profiles/r02b_packed_ab.md): packed complex adds
// only — packed products cost this FMA-pipe-bound kernel 2-13 %.
#ifndef OMB_F32X2_CMUL
#define OMB_F32X2_CMUL 0
#endif
#include <cstdlib>

#include "async_copy.cuh"
#include "device_math.cuh"
#include "fft16.cuh"
#include "fft4096.cuh"
#include "stft.h"

namespace omb {

namespace {

using namespace f4k;

constexpr int kN2 = 2048;                // window = complex points per frame
constexpr int kM = 4096;                 // engine length
constexpr int kGroupsPerCta = 2;
constexpr int kFramesPerGroup = 2;
constexpr int kFramesPerIter = kGroupsPerCta * kFramesPerGroup;
constexpr int kThreads = kT * kGroupsPerCta;
constexpr int kWarps = kT / 32;          // warps per group
constexpr int kWSize = f16::phys_size(kM);
constexpr int kBinGroups = 5;            // bins t + 256 j, j < 4, and bin 1024 (t = 0, j = 4)

struct Fast2kArgs {
  StftKernelArgs a;
  const float2* tw1;   // global: [15][256] W_4096^{b q}
  const float2* tw2;   // global: [15][16]  W_256^{o q}
  uint32_t frames_per_run, runs_per_lane, ring_len;
  uint32_t reps;       // loop iterations' worth of frames per CTA barrier / ring prefetch (stft_fast2.cu: `pairs`): as many as the ring holds
  uint64_t chunk;      // > 0: one contiguous range of `chunk` frames of the linearised (lane, frame) sequence per CTA (stft_fast2.cu)
  float norm_ac, norm_dc;  // bin_norm / 4 (the separated spectra are doubled)
};

struct GroupSmem {
  float2 W[kWSize];
  float Y[kFramesPerGroup][kN2 + 16];    // Y[fr][n] = Im c_fr[n]; +16: the two frames of a warp access (same n) hit disjoint banks
  int warp_cnt[kFramesPerGroup][kBinGroups * kWarps];
  int offs[kFramesPerGroup][kBinGroups * kWarps + 1];
  float x0_xm[kFramesPerGroup][2];
};

struct Smem2k {
  float2 tw1[4 * kT];   // rows q = 1, 2, 4, 8 of W_4096^{b q}
  float2 tw2[15 * 16];
  float h[kN2];
  float dh[kN2];
  GroupSmem g[kGroupsPerCta];
  // float ring[ring_len] follows
};

__device__ __forceinline__ void k2_ring_fetch(float* ring, int ring_mask, const float* x, uint64_t s0, uint64_t s1) {
  ring_fetch_pow2(ring, ring_mask, x, s0, s1, kThreads);
}

// After a forward transform of two interleaved frames: v[j] <- 2 FA[t + 256 j], v[8 + j] <- 2 FB[t + 256 j], j < 8.
// (c0, s0) = (cos, sin)(2 pi t / 4096); the bin's angle adds 2 pi j / 16.
__device__ __forceinline__ void separate2(float2 (&v)[16], float c0, float s0) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float cj = f16::kCos32[2 * j], sj = f16::kSin32[2 * j];
    const float c = c0 * cj - s0 * sj, s = s0 * cj + c0 * sj;
    const float2 a = v[j], b = v[j + 8];
    const float2 d = f16::csub2(a, b);
    v[j] = f16::cadd2(a, b);
    v[j + 8] = make_float2(d.x * c - d.y * s, d.y * c + d.x * s);  // d * conj(W^k) = d * (c + j s)
  }
}

__global__ void __launch_bounds__(kThreads, 1) k_reassigned_fast2k(Fast2kArgs fa) {
  constexpr int kTw2 = 0;
  OMB_DYN_SMEM(unsigned char, smem_raw);
  Smem2k& sm = *reinterpret_cast<Smem2k*>(smem_raw);
  float* ring = reinterpret_cast<float*>(smem_raw + sizeof(Smem2k));
  const StftKernelArgs& a = fa.a;
  const int tid = threadIdx.x, t = tid & (kT - 1), lane_id = t & 31, warp = t >> 5;
  const int g = __shfl_sync(0xffffffffu, tid >> 8, 0);
  GroupSmem& gs = sm.g[g];
  const int hop = (int)a.hop, H = 2 * kN2, ring_mask = (int)fa.ring_len - 1;  // ring_len is a power of two
  const int off = (H - kN2) / 2;
  const uint64_t total_runs = (uint64_t)fa.runs_per_lane * a.n_lanes;
  const ReassignConsts rc{a.bin_hz, a.max_hz, a.inv_2pi, a.inv_hop, a.latency_hops};

  for (int i = tid; i < 4 * kT; i += kThreads) {
    const int row = (1 << (i >> 8)) - 1;  // q - 1 for q = 1, 2, 4, 8
    sm.tw1[i] = __ldg(&fa.tw1[row * kT + (i & (kT - 1))]);
  }
  for (int i = tid; i < 15 * 16; i += kThreads) sm.tw2[i] = __ldg(&fa.tw2[i]);
  for (int i = tid; i < kN2; i += kThreads) {
    sm.h[i] = __ldg(&a.win[i]);
    sm.dh[i] = __ldg(&a.dwin[i]);
  }
  Addr ad;
  ad.pA = t + (t >> 4);
  ad.pB = 273 * (t >> 4) + (t & 15);
  ad.pC = 273 * (t & 15) + 17 * (t >> 4);
  const float2* tw1t = sm.tw1 + t;
  const float2* tw2o = sm.tw2 + (t & 15);
  float cos_t, sin_t;  // angle of bin t: 2 pi t / 4096 — the Hilbert pair-step angle (H = 4096) and the separation twiddle alike
  sincospif((float)t / (float)kN2, &sin_t, &cos_t);
  const int par = t & 1, hh = t >> 1;  // time domain: this thread's samples are hh + 128 j of frame `par`
  const float sign = (hh & 1) ? -1.0f : 1.0f;
  const float ramp0 = (float)hh - (float)(kN2 - 1) * 0.5f;  // n - (N-1)/2 at j = 0
  const int pt = (kT - t) & (kT - 1);  // owner of the partner bins 2048 - (t + 256 j)
  const int pPartner = 273 * (pt & 15) + 17 * (pt >> 4);
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
    const uint64_t s_end = (f_end - 1) * (uint64_t)hop + (uint64_t)H;  // one past the last sample this run reads
    const uint64_t iter_frames = (uint64_t)kFramesPerIter * fa.reps;
    {  // prime: everything the first four frames read
      const uint64_t s0 = f_begin * (uint64_t)hop;
      const uint64_t want = s0 + (uint64_t)H + (iter_frames - 1) * hop;
      k2_ring_fetch(ring, ring_mask, x, s0, want < s_end ? want : s_end);
      async_commit();
    }
    for (uint64_t fa0 = f_begin; fa0 < f_end; fa0 += iter_frames) {
      async_wait_all();
      __syncthreads();  // ring holds frames fa0 .. fa0+3; both groups are done with the previous four
      {                 // prefetch what the next four frames add: four hops
        const uint64_t s0 = fa0 * (uint64_t)hop + (uint64_t)H + (iter_frames - 1) * hop;
        const uint64_t want = s0 + iter_frames * hop;
        const uint64_t s1 = want < s_end ? want : s_end;
        if (s0 < s1) k2_ring_fetch(ring, ring_mask, x, s0, s1);
        async_commit();
      }
#pragma unroll 1
      for (uint32_t sp = 0; sp < fa.reps; ++sp) {
      const uint64_t fg = fa0 + (uint64_t)kFramesPerIter * sp + (uint64_t)kFramesPerGroup * g;  // this group's frames: fg, fg + 1
      if (fg < f_end) {
        // a missing second frame (odd tail of a run) is fed as zeros: the two frames share every transform, so stale ring
        // contents (possibly NaN bit patterns) would leak into the first frame through X = FA + W FB
        const bool fvalid = fg + (uint64_t)par < f_end;
        const int r0 = (int)((fg + par) * (uint64_t)hop) & ring_mask;  // ring origin of this thread's time-domain frame
        float2 v[16];
        // ---- F: z_par[n] = x[2n] + j x[2n+1], n = hh + 128 j  (engine input a[t + 256 j])
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          v[j] = *reinterpret_cast<const float2*>(ring + ((r0 + 2 * hh + 256 * j) & ring_mask));
          if (!fvalid) v[j] = make_float2(0.0f, 0.0f);
        }
        fft_forward<f16::kAll, kTw2, false>(v, gs.W, tw1t, tw2o, ad, g);
        separate2(v, cos_t, sin_t);  // v[j] = 2 ZA[t + 256 j], v[8 + j] = 2 ZB[t + 256 j]
#pragma unroll
        for (int q = 0; q < 16; ++q) gs.W[ad.pC + q] = v[q];
        if (t == 0) {  // X[0] and X[H/2] of each frame's real spectrum
          gs.x0_xm[0][0] = 0.5f * (v[0].x + v[0].y);
          gs.x0_xm[0][1] = 0.5f * (v[0].x - v[0].y);
          gs.x0_xm[1][0] = 0.5f * (v[8].x + v[8].y);
          gs.x0_xm[1][1] = 0.5f * (v[8].x - v[8].y);
        }
        group_sync(g);
        // ---- X: Q[k] = cos(th_k) conj(Z[2048-k]) + j sin(th_k) Z[k], th_k = 2 pi k / 4096, per frame; then the two
        //         inverse inputs are combined: X[k] = QA + W^k QB, X[k + 2048] = QA - W^k QB  (factors 1/2 folded: 0.25)
        {
          const float2* wp = gs.W + pPartner;
          float2 zp[16];
          if (t == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              zp[j] = wp[(8 - j) & 7];
              zp[8 + j] = wp[8 + ((8 - j) & 7)];
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              zp[j] = wp[7 - j];
              zp[8 + j] = wp[15 - j];
            }
          }
          group_sync(g);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float cj = f16::kCos32[2 * j], sj = f16::kSin32[2 * j];
            const float cu = cos_t * cj - sin_t * sj, su = sin_t * cj + cos_t * sj;
            const float ck = 0.25f * cu, sk = 0.25f * su;
            const float2 za = v[j], zb = v[8 + j];
            float2 qa = make_float2(ck * zp[j].x - sk * za.y, sk * za.x - ck * zp[j].y);
            float2 qb = make_float2(ck * zp[8 + j].x - sk * zb.y, sk * zb.x - ck * zp[8 + j].y);
            if (t == 0 && j == 0) qa = qb = make_float2(0.0f, 0.0f);
            const float2 wb = make_float2(qb.x * cu + qb.y * su, qb.y * cu - qb.x * su);  // qb * W^k = qb * (cu - j su)
            v[j] = f16::cadd2(qa, wb);
            v[8 + j] = f16::csub2(qa, wb);
          }
        }
        // ---- I: inverse, DIT; output a[t + 256 q]: sample hh + 128 q of frame par's packed analytic signal; centre half
        {
          f16::dft16<true>(v);
          float2* wc = gs.W + ad.pC;
#pragma unroll
          for (int q = 0; q < 16; ++q) wc[q] = v[q];
          group_sync(g);
          float2* wb = gs.W + ad.pB;
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = wb[17 * j];
          twiddle15<true, kTw2, false>(v, tw2o, 16);
          f16::dft16<true>(v);
#pragma unroll
          for (int q = 0; q < 16; ++q) wb[17 * q] = v[q];
          group_sync(g);
          const float2* wa = gs.W + ad.pA;
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = wa[273 * j];
          twiddle15<true, 1, true>(v, tw1t, kT);
          f16::dft16p<true, f16::kMid8>(v);
          // packed index m = hh + 128 q, q = 4..11 -> centre samples: Y[par] as float2[m - 512]
          float2* y2 = reinterpret_cast<float2*>(gs.Y[par]) + hh;
#pragma unroll
          for (int q = 4; q < 12; ++q) y2[128 * (q - 4)] = v[q];
          group_sync(g);
        }
        // ---- G: three windowed transforms of c_par[n] = (N x[off+n] + bias) + j Y[n], n = hh + 128 j
        const float bias = sign * 0.5f * gs.x0_xm[par][1] - 0.5f * gs.x0_xm[par][0];
        float2 S[kFramesPerGroup][kBinGroups];
        float nd[kFramesPerGroup][kBinGroups];
#pragma unroll 1
        for (int wsel = 0; wsel < 3; ++wsel) {
          const float* win = (wsel == 1 ? sm.dh : sm.h) + hh;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float wv = win[128 * j];
            if (wsel == 2) wv *= ramp0 + (float)(128 * j);      // t*h window, processor.rs:601-608
            const float cx = fmaf((float)kN2, ring[(r0 + off + hh + 128 * j) & ring_mask], bias);
            v[j] = fvalid ? make_float2(cx * wv, gs.Y[par][hh + 128 * j] * wv) : make_float2(0.0f, 0.0f);
          }
          fft_forward<f16::kAll, kTw2, false>(v, gs.W, tw1t, tw2o, ad, g);
          separate2(v, cos_t, sin_t);  // v[j] = 2 S_A[t + 256 j], v[8 + j] = 2 S_B[t + 256 j]
          if (wsel == 0) {
#pragma unroll
            for (int j = 0; j < kBinGroups; ++j) {
              S[0][j] = v[j];
              S[1][j] = v[8 + j];
            }
          } else if (wsel == 1) {
#pragma unroll
            for (int j = 0; j < kBinGroups; ++j) {
              nd[0][j] = v[j].y * S[0][j].x - v[j].x * S[0][j].y;
              nd[1][j] = v[8 + j].y * S[1][j].x - v[8 + j].x * S[1][j].y;
            }
          }
          group_sync(g);  // pass-3 reads done before the next pass-1 stores (and before warp_cnt reuse)
        }
        // ---- R: reassignment + ordered compaction, per frame (bin order (j, t) as in stft_fast2.cu)
        omb_spectrogram_point pts[kFramesPerGroup][kBinGroups];
        int rank[kFramesPerGroup][kBinGroups];
        unsigned keep = 0;
#pragma unroll
        for (int fr = 0; fr < kFramesPerGroup; ++fr)
#pragma unroll
          for (int j = 0; j < kBinGroups; ++j) {
            const int bin = t + kT * j;
            const float norm = (bin == 0 || j == 4) ? fa.norm_dc : fa.norm_ac;
            const bool k = reassign_bin_nd(S[fr][j], nd[fr][j], v[8 * fr + j], norm, bin, rc, &pts[fr][j]) & (j < 4 || t == 0);
            const unsigned m = __ballot_sync(0xffffffffu, k);
            if (lane_id == 0) gs.warp_cnt[fr][j * kWarps + warp] = __popc(m);
            if (k) keep |= 1u << (fr * kBinGroups + j);
            rank[fr][j] = __popc(m & ((1u << lane_id) - 1u));
          }
        group_sync(g);
        if (warp < kFramesPerGroup) {  // warp fr scans frame fr's 40 (bin group, warp) counts, two per lane
          const int i0 = lane_id * 2;
          const int n_cnt = kBinGroups * kWarps;
          const int c0 = i0 < n_cnt ? gs.warp_cnt[warp][i0] : 0;
          const int c1 = i0 + 1 < n_cnt ? gs.warp_cnt[warp][i0 + 1] : 0;
          const int tot = c0 + c1;
          int incl = tot;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane_id >= o) incl += n;
          }
          const int excl = incl - tot;
          if (i0 < n_cnt) gs.offs[warp][i0] = excl;
          if (i0 + 1 < n_cnt) gs.offs[warp][i0 + 1] = excl + c0;
          if (lane_id == 31) gs.offs[warp][n_cnt] = incl;
        }
        group_sync(g);
#pragma unroll
        for (int fr = 0; fr < kFramesPerGroup; ++fr) {
          const uint64_t f = fg + fr;
          if (f < f_end) {
            const uint64_t slot = lane * a.frames_per_lane + f;
            float* out = reinterpret_cast<float*>(a.out_points + slot * a.point_stride);
#pragma unroll
            for (int j = 0; j < kBinGroups; ++j)
              if (keep & (1u << (fr * kBinGroups + j))) {
                float* o = out + 3 * (gs.offs[fr][j * kWarps + warp] + rank[fr][j]);
                o[0] = pts[fr][j].time_offset;
                o[1] = pts[fr][j].freq_hz;
                o[2] = pts[fr][j].power;
              }
            if (t == 0) a.out_counts[slot] = (uint32_t)gs.offs[fr][kBinGroups * kWarps];
          }
        }
      }
      }  // sp
    }
    async_wait_all();
    __syncthreads();
  }
}

uint32_t ring_len_for(uint64_t hop) { return (uint32_t)next_pow2(2 * (uint64_t)kN2 + (2 * kFramesPerIter - 1) * hop); }
size_t smem_bytes(uint64_t hop) { return sizeof(Smem2k) + (size_t)ring_len_for(hop) * sizeof(float); }

}  // namespace

bool stft_fast2k_supported(const StftConfig& cfg, const DeviceInfo& dev) {
  if (!cfg.reassign || cfg.window != (uint64_t)kN2 || cfg.zero_pad != 1) return false;
  // any multiple of 4 (16-byte async copies) up to N/2: of the UI's N/4 ... N/128 only N/6 = 341 is excluded
  if (cfg.hop < 4 || (cfg.hop % 4) != 0 || cfg.hop > 1024) return false;
  return dev.max_smem_optin == 0 || smem_bytes(cfg.hop) <= (size_t)dev.max_smem_optin;
}

int stft_fast2k_prepare(StftPlan& plan) {
  // twiddle tables of the 4096-point engine: [15*256] W_4096^{b q} | [15*16] W_256^{o q}
  std::vector<float2> tab(15 * kT + 15 * 16);
  const double tau = 6.28318530717958647692;
  for (int q = 1; q < 16; ++q)
    for (int b = 0; b < kT; ++b) {
      const double ang = -tau * (double)((b * q) % kM) / (double)kM;
      tab[(q - 1) * kT + b] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  for (int q = 1; q < 16; ++q)
    for (int o = 0; o < 16; ++o) {
      const double ang = -tau * (double)((o * q) % 256) / 256.0;
      tab[15 * kT + (q - 1) * 16 + o] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  OMB_TRY(plan.d_fast_tables.upload(tab, plan.stream));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_fast2k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(plan.cfg.hop)));
  return OMB_OK;
}

int launch_stft_fast2k(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s) {
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  if (per_lane == 0 || a.n_lanes == 0) return OMB_OK;
  if ((reinterpret_cast<uintptr_t>(a.lanes) & 15u) != 0 || (a.lane_stride % 4) != 0)
    return fail(OMB_ERR_INVALID, "specialised STFT kernel needs 16-byte aligned lanes (pointer and lane_stride % 4 == 0)");
  Fast2kArgs fa{};
  fa.a = a;
  fa.tw1 = plan.d_fast_tables.ptr;
  fa.tw2 = fa.tw1 + 15 * kT;
  fa.ring_len = ring_len_for(a.hop);
  fa.norm_ac = 0.25f * (plan.h_norm.size() > 1 ? plan.h_norm[1] : plan.h_norm[0]);
  fa.norm_dc = 0.25f * plan.h_norm[0];
  const uint64_t ctas = (uint64_t)std::max(plan.dev.sm_count, 1);
  uint64_t run = 256;  // multiple of four, long enough to amortise the ring prime (4096 samples vs hop per frame)
  if (a.hop < 512) run *= 512 / a.hop;
  while (run > 16 && ((per_lane + run - 1) / run) * a.n_lanes < ctas * 6) run >>= 1;
  fa.frames_per_run = (uint32_t)std::min<uint64_t>(run, per_lane);
  fa.runs_per_lane = (uint32_t)((per_lane + fa.frames_per_run - 1) / fa.frames_per_run);
  const uint64_t total_runs = (uint64_t)fa.runs_per_lane * a.n_lanes;
  unsigned grid = (unsigned)std::min<uint64_t>(total_runs, ctas);
  // iterations per CTA barrier: as many as fit the ring next to the following prefetch, H + (2 F reps - 1) hop <= ring (OMB_FAST2K_REPS pins)
  static const int reps_env = [] { const char* e = getenv("OMB_FAST2K_REPS"); return e ? atoi(e) : 0; }();
  uint32_t reps = (uint32_t)std::min<uint64_t>(8, ((uint64_t)fa.ring_len - 2 * (uint64_t)kN2 + a.hop) / (2ull * kFramesPerIter * a.hop));
  if (reps_env > 0) reps = std::min<uint32_t>(reps, (uint32_t)reps_env);
  fa.reps = std::max<uint32_t>(1, reps);
  // Large batches: one contiguous range of frames per CTA, primed once per lane touched (stft_fast2.cu; OMB_FAST2K_CONTIG=0: round-robin runs)
  static const bool contig_env = [] { const char* e = getenv("OMB_FAST2K_CONTIG"); return !(e && e[0] == '0'); }();
  const uint64_t total_frames = per_lane * a.n_lanes;
  fa.chunk = 0;
  if (contig_env && total_frames >= ctas * 192) {
    fa.chunk = ((total_frames + ctas - 1) / ctas + 3) / 4 * 4;
    grid = (unsigned)((total_frames + fa.chunk - 1) / fa.chunk);
  }
  OMB_LAUNCH(k_reassigned_fast2k, dim3(grid), dim3(kThreads), smem_bytes(a.hop), s, fa);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
