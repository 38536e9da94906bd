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
//
// The kernel is a template on the team size kT: kT = 256 is the above (N = 16384 = 64 x 64 x 4, one team per CTA); kT = 128 is
// N = 8192 = 64 x 64 x 2 (pass C = radix-2 across a PAIR of threads), two independent 128-thread teams per CTA on two frames, the
// window tables in shared memory — the cfg5 size, next to stft_fast8k.cu (OMB_R64X_8K=0|1 picks; profiles/r02_notes.md).
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

template <int kT>
struct Geo {
  static constexpr int kM = 64 * kT;             // complex points per transform = window length N
  static constexpr int kH = 2 * kM;              // samples per frame (Hilbert block)
  static constexpr int kOff = (kH - kM) / 2;     // first centre sample
  static constexpr int kA = kT / 64;             // radix of pass C = threads per sub-problem (4 or 2)
  static constexpr int kD = 64 / kA;             // pass-C butterflies per thread
  static constexpr int kTeams = 256 / kT;        // teams (frames in flight) per CTA
  static constexpr int kRS = kT + kA;            // row stride of exchange 1 (float2): [k1][t], read back at stride kA — conflict-free
  static constexpr int kSkew = 16 / kA;          // pair rows: tau * 65 + j + kSkew (tau >> 6): rows tau, tau + 64 on different banks
  static constexpr int kWSize = 64 * kRS + 16;   // float2 elements of a team's buffer
  static constexpr int kWarps = kT / 32;
  static constexpr unsigned kFrameBytes = kH * sizeof(float);
  static constexpr bool kWinSmem = kT == 128;    // window tables in shared memory (2 x 32 KB) / global memory (2 x 64 KB)
};
constexpr int kRSb = 65;                 // exchange 2: [k1][a][k2] at k1 * kRS + a * 65 + k2
constexpr int kGroups = 33;              // bins tau + kT j, j < 32, and the Nyquist bin (tau = 0, j = 32)
constexpr int kColC = 0, kColS = 128, kColNd = 192, kColSn = 224, kColsPerWarp = 256;  // as in stft_r64.cu

struct R64xArgs {
  StftKernelArgs a;
  const float2* tw1;    // global: [14][kT]: rows 0..6 = W_M^{t q}, q = 1..7; rows 7..13 = W_M^{8 t q}
  const float2* tw2;    // global: [14][kA]: the same rows of W_kT^{a q}, a < kA
  float norm_ac, norm_dc;
};

template <int kT>
struct TeamSmem {
  alignas(16) float2 W[Geo<kT>::kWSize];
  int cnt[(kGroups + 1) * Geo<kT>::kWarps];
  int offs[(kGroups + 1) * Geo<kT>::kWarps + 2];
  float x0_xm[2];
  alignas(8) uint64_t mbar;
  uint64_t pad_;
};
template <int kT>
struct Smem {
  TeamSmem<kT> team[Geo<kT>::kTeams];
  float2 tw1[14 * kT];
  float2 tw2[14 * 4];
  uint32_t tmem_base;
  uint32_t pad_[3];
  // kWinSmem: float h[kM], dh[kM] follow
};

template <int kT>
__device__ __forceinline__ void team_sync(int team) {
#ifdef OMB_EMU
  omb_emu::named_sync(1 + team, kT);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(kT) : "memory");
#endif
}

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

// Passes A and B of the forward transform and the loads of pass C.  On entry the thread holds elements ta + kT j (ta: the residue
// class it owns, any bijection thread -> class); on return v[d + kD a] holds the pass-C input (a < kA, d < kD) of the radix-kA
// butterfly whose outputs k3 are the bins tau + kT (d + kD k3), tau = (t / kA) + 64 (t % kA).
template <int kT>
__device__ __forceinline__ void transform_ab(float2 (&v)[64], float2* W, const float2* tw1, const float2* tw2, int ta, int t, int team) {
  using G = Geo<kT>;
  f64pt::dft64<false>(v);
  twiddle63(v, tw1 + ta, kT);
  team_sync<kT>(team);  // every earlier read of W (frame samples, partner rows, staging, the previous transform) is done
  float2* wr = W + ta;
#pragma unroll
  for (int q = 0; q < 64; ++q) wr[q * G::kRS] = v[q];
  team_sync<kT>(team);
  const int k1 = t / G::kA, a = t % G::kA;
  const float2* rd = W + k1 * G::kRS + a;
#pragma unroll
  for (int b = 0; b < 64; ++b) v[b] = rd[G::kA * b];
  f64pt::dft64<false>(v);
  twiddle63(v, tw2 + a, 4);
  team_sync<kT>(team);  // all of exchange 1 has been read: exchange 2 reuses the buffer
  float2* wq = W + k1 * G::kRS + a * kRSb;
#pragma unroll
  for (int k2 = 0; k2 < 64; ++k2) wq[k2] = v[k2];
  __syncwarp();  // exchange 2 stays inside the kA threads (k1, 0..kA-1) of one warp
  const float2* rq = W + k1 * G::kRS + a;  // this thread's c = a
#pragma unroll
  for (int d = 0; d < G::kD; ++d)
#pragma unroll
    for (int s2 = 0; s2 < G::kA; ++s2) v[d + G::kD * s2] = rq[s2 * kRSb + G::kA * d];
}
// Pass C: radix-kA butterflies over (v[d], v[d + kD], ...); kPrune = kFirst9: only outputs j <= 32 (bins <= Nyquist), kMid8: only
// 16 <= j < 48 (the centre half of the inverse), kAll.
template <int kT, int kPrune>
__device__ __forceinline__ void transform_c(float2 (&v)[64]) {
  if (Geo<kT>::kA == 4) {
    f64pt::dit_final<false, kPrune>(v);
  } else {
#pragma unroll
    for (int d = 0; d < 32; ++d) {
      const float2 a0 = v[d], a1 = v[d + 32];
      if (kPrune == f16::kAll) {
        v[d] = f16::cadd2(a0, a1);
        v[d + 32] = f16::csub2(a0, a1);
      } else if (kPrune == f16::kFirst9) {
        v[d] = f16::cadd2(a0, a1);
        if (d == 0) v[32] = f16::csub2(a0, a1);
      } else {  // kMid8: j = d (k3 = 0) for d >= 16, j = d + 32 (k3 = 1) for d < 16
        if (d >= 16) v[d] = f16::cadd2(a0, a1);
        else v[d + 32] = f16::csub2(a0, a1);
      }
    }
  }
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

template <int kT>
__global__ void __launch_bounds__(256, 1) k_reassigned_r64x(R64xArgs ra) {
  using G = Geo<kT>;
  constexpr int kM = G::kM;
  OMB_DYN_SMEM(unsigned char, smem_raw);
  Smem<kT>& sm = *reinterpret_cast<Smem<kT>*>(smem_raw);
  float* sm_h = reinterpret_cast<float*>(smem_raw + sizeof(Smem<kT>));  // kWinSmem only
  float* sm_dh = sm_h + kM;
  const StftKernelArgs& a = ra.a;
  const int tid = threadIdx.x, t = tid % kT, lane_id = tid & 31, warp = t >> 5;
  const int team = __shfl_sync(0xffffffffu, tid / kT, 0);  // warp-uniform by construction; tells the compiler so
  TeamSmem<kT>& ts = sm.team[team];
  float2* W = ts.W;
  const ReassignConsts rc{a.bin_hz, a.max_hz, a.inv_2pi, a.inv_hop, a.latency_hops};

  for (int i = tid; i < 14 * kT; i += 256) sm.tw1[i] = __ldg(&ra.tw1[i]);
  if (tid < 14 * 4) sm.tw2[tid] = __ldg(&ra.tw2[tid]);
  if (G::kWinSmem) {
    for (int i = tid; i < kM; i += 256) {
      sm_h[i] = __ldg(&a.win[i]);
      sm_dh[i] = __ldg(&a.dwin[i]);
    }
  }
  if (tid < 32) tmem_alloc(&sm.tmem_base);
  if (t == 0) mbar_init(&ts.mbar, 1);
  tmem_fence_before_sync();
  __syncthreads();
  tmem_fence_after_sync();
  const uint32_t tcol = tmem_addr(sm.tmem_base, (uint32_t)(tid >> 7) * kColsPerWarp);

  // per-thread constants
  const int tau = (t / G::kA) + 64 * (t % G::kA);  // bins tau + kT j after a transform
  float cos_t, sin_t;                              // th_tau = 2 pi tau / H
  sincospif((float)tau / (float)kM, &sin_t, &cos_t);
  const float sign = (t & 1) ? -1.0f : 1.0f;       // (-1)^n of the analysis point n = t + kT j
  const float ramp0 = (float)t - (float)(kM - 1) * 0.5f;
  const int ptau = (kT - tau) & (kT - 1);          // owner class of the mirror bins M - k
  auto row_of = [](int r) { return r * kRSb + G::kSkew * (r >> 6); };

  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total = per_lane * a.n_lanes;
  const uint64_t stride = (uint64_t)gridDim.x * G::kTeams;
  auto frame_src = [&](uint64_t gi) {
    const uint64_t l = gi / per_lane, f = a.first_frame + gi % per_lane;
    return a.lanes + l * a.lane_stride + f * (uint64_t)a.hop;
  };
  uint64_t g = (uint64_t)blockIdx.x * G::kTeams + team;
  unsigned phase = 0;
  if (g < total && t == 0) {
    mbar_expect_tx(&ts.mbar, G::kFrameBytes);
    bulk_g2s(W, frame_src(g), G::kFrameBytes, &ts.mbar);
  }

  for (; g < total; g += stride) {
    const uint64_t lane = g / per_lane, f = a.first_frame + g % per_lane;
    mbar_wait(&ts.mbar, phase);
    phase ^= 1u;
    float2 v[64];
    {
      const float* wf = reinterpret_cast<const float*>(W);
      float xc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) xc[j] = wf[G::kOff + t + kT * j];
      park_st<64>(tcol, kColC, xc);
      const float2* wz = W + t;  // F input: z[n] = x[2n] + j x[2n+1], n = t + kT j
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = wz[kT * j];
    }
#pragma unroll 1
    for (int tr = 0; tr < 5; ++tr) {
      int ta = t;
      if (tr == 1) {
        // ---- X: conj(Q[k]), Q[k] = cos(th_k) conj(Z[M-k]) + j sin(th_k) Z[k], k = tau + kT j, th_k = th_tau + 2 pi j / 128
        team_sync<kT>(team);
        float2* row = W + row_of(tau);
#pragma unroll
        for (int j = 0; j < 64; ++j) row[j] = v[j];
        if (tau == 0) {
          ts.x0_xm[0] = v[0].x + v[0].y;
          ts.x0_xm[1] = v[0].x - v[0].y;
        }
        team_sync<kT>(team);
        // partner of k = tau + kT j: row kT - tau, column 63 - j; for tau = 0 row 0, column 64 - j (j = 0: don't-care, DC is zeroed)
        const float2* prow = tau == 0 ? W + 1 : W + row_of(ptau);
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
        // ---- G input: c[n] w[n], n = t + kT j, w = h, dh, t*h (processor.rs:601-608, formed on the fly, bit-identical)
        float c[128];
        park_ld<128>(tcol, kColC, c);
        const float* tab = (G::kWinSmem ? (tr == 3 ? sm_dh : sm_h) : (tr == 3 ? a.dwin : a.win)) + t;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
          float wv = G::kWinSmem ? tab[kT * j] : __ldg(&tab[kT * j]);
          if (tr == 4) wv *= ramp0 + (float)(kT * j);
          v[j] = f16::cscale2(make_float2(c[2 * j], c[2 * j + 1]), wv);
        }
      }
      transform_ab<kT>(v, W, sm.tw1, sm.tw2, ta, t, team);
      if (tr == 0) {
        transform_c<kT, f16::kAll>(v);
      } else if (tr == 1) {
        transform_c<kT, f16::kMid8>(v);    // outputs 16..47: the centre half
      } else {
        transform_c<kT, f16::kFirst9>(v);  // outputs 0..32: bins <= Nyquist
      }
      if (tr == 1) {
        // ---- centre half of the inverse: y[m] = conj(v), m = tau + kT j, j = 16..47 -> staging as float2[m - M/4]; then every
        //      thread collects Im c of its analysis points n = t + kT j and parks c[n] = (M x[off + n] + bias) + j Im c[n]
        team_sync<kT>(team);
        float2* y2 = W + tau;
#pragma unroll
        for (int q = 16; q < 48; ++q) y2[kT * (q - 16)] = make_float2(v[q].x, -v[q].y);
        team_sync<kT>(team);
        const float bias = sign * 0.5f * ts.x0_xm[1] - 0.5f * ts.x0_xm[0];
        const float* yf = reinterpret_cast<const float*>(W) + t;
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
    team_sync<kT>(team);  // every thread group has finished reading exchange 2
    {
      float* stage = reinterpret_cast<float*>(W);
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
    team_sync<kT>(team);
    // ---- ordered compaction in natural order: thread t, group q -> bin t + kT q (thread 0 also the Nyquist bin)
    {
      const float* stage = reinterpret_cast<const float*>(W);
      unsigned keep_lo = 0, keep_hi = 0;
#pragma unroll
      for (int q = 0; q < kGroups; ++q) {
        const int bin = t + kT * q;
        const bool k = (q < 32 || t == 0) && stage[3 * (q < 32 ? bin : kM / 2) + 2] >= 0.0f;
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (lane_id == 0) ts.cnt[q * G::kWarps + warp] = __popc(m);
        if (k) {
          if (q < 32) keep_lo |= 1u << q; else keep_hi = 1u;
        }
      }
      team_sync<kT>(team);
      if (warp == 0) {  // exclusive scan of the 33 x kWarps warp counts (order: group, warp)
        constexpr int kN = kGroups * G::kWarps, kPer = (kN + 31) / 32;
        int c[kPer], tot = 0;
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          const int idx = lane_id * kPer + i;
          c[i] = idx < kN ? ts.cnt[idx] : 0;
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
          if (idx < kN) ts.offs[idx] = run;
          run += c[i];
        }
        if (lane_id == 31) ts.offs[kN] = incl;
      }
      team_sync<kT>(team);
      const uint64_t slot = lane * a.frames_per_lane + f;
      float* out = reinterpret_cast<float*>(a.out_points + slot * a.point_stride);
      const unsigned lt_mask = (1u << lane_id) - 1u;
#pragma unroll
      for (int q = 0; q < kGroups; ++q) {
        const bool k = q < 32 ? ((keep_lo >> q) & 1u) != 0 : keep_hi != 0;
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (k) {
          const float* src = stage + 3 * (q < 32 ? t + kT * q : kM / 2);
          float* o = out + 3 * (ts.offs[q * G::kWarps + warp] + __popc(m & lt_mask));
          o[0] = src[0];
          o[1] = src[1];
          o[2] = src[2];
        }
      }
      if (t == 0) a.out_counts[slot] = (uint32_t)ts.offs[kGroups * G::kWarps];
    }
    // ---- the staged column has been read: the next frame may land
    team_sync<kT>(team);
    if (t == 0 && g + stride < total) {
      fence_async_smem();
      mbar_expect_tx(&ts.mbar, G::kFrameBytes);
      bulk_g2s(W, frame_src(g + stride), G::kFrameBytes, &ts.mbar);
    }
  }
  __syncthreads();
  if (tid < 32) tmem_free(sm.tmem_base);
}

template <int kT>
size_t smem_bytes() { return sizeof(Smem<kT>) + (Geo<kT>::kWinSmem ? 2 * sizeof(float) * Geo<kT>::kM : 0); }

// N = 8192: OMB_R64X_8K=1 / 0 pins this kernel / stft_fast8k.cu; unpinned, stft_fast8k.cu keeps the hops its ring handles with
// warp-uniform rows (multiples of 512 up to 2048: the two are within 2 % of each other there, profiles/r02_notes.md) and this
// kernel takes every other multiple of 4.
bool r64x_for_8k(uint64_t hop) {
  static const int v = [] { const char* e = getenv("OMB_R64X_8K"); return e ? atoi(e) : -1; }();
  if (v >= 0) return v != 0;
  return (hop % 512) != 0 || hop > 2048;
}

}  // namespace

bool stft_r64x_supported(const StftConfig& cfg, const DeviceInfo& dev) {
  if (!cfg.reassign || cfg.zero_pad != 1) return false;
  if (cfg.window != 16384u && !(cfg.window == 8192u && r64x_for_8k(cfg.hop))) return false;
  if (cfg.hop < 4 || (cfg.hop % 4) != 0) return false;  // bulk copies start at f * hop floats: 16-byte aligned
  if (dev.cc_major != 0 && dev.cc_major < 10) return false;  // tcgen05 / TMEM
  const size_t need = cfg.window == 16384u ? smem_bytes<256>() : smem_bytes<128>();
  return dev.max_smem_optin == 0 || need <= (size_t)dev.max_smem_optin;
}

int stft_r64x_prepare(StftPlan& plan) {
  const int kT = (int)(plan.cfg.window / 64), kA = kT / 64, M = 64 * kT;
  std::vector<float2> tab(14 * kT + 14 * 4);
  const double tau = 6.28318530717958647692;
  for (int i = 0; i < 14; ++i) {
    const int q = i < 7 ? i + 1 : 8 * (i - 6);
    for (int t = 0; t < kT; ++t) {
      const double ang = -tau * (double)((t * q) % M) / (double)M;
      tab[i * kT + t] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
    for (int a = 0; a < 4; ++a) {
      const double ang = a < kA ? -tau * (double)((a * q) % kT) / (double)kT : 0.0;
      tab[14 * kT + i * 4 + a] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  }
  OMB_TRY(plan.d_r64_tables.upload(tab, plan.stream));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_r64x<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<256>()));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_r64x<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<128>()));
  return OMB_OK;
}

int launch_stft_r64x(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s) {
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  if (per_lane == 0 || a.n_lanes == 0) return OMB_OK;
  if ((reinterpret_cast<uintptr_t>(a.lanes) & 15u) != 0 || (a.lane_stride % 4) != 0)
    return fail(OMB_ERR_INVALID, "specialised STFT kernel needs 16-byte aligned lanes (pointer and lane_stride % 4 == 0)");
  const int kT = (int)(a.window / 64);
  R64xArgs ra{};
  ra.a = a;
  ra.tw1 = plan.d_r64_tables.ptr;
  ra.tw2 = ra.tw1 + 14 * kT;
  ra.norm_ac = plan.h_norm.size() > 1 ? plan.h_norm[1] : plan.h_norm[0];
  ra.norm_dc = plan.h_norm[0];
  const uint64_t total = per_lane * a.n_lanes;
  const uint64_t teams = kT == 256 ? 1 : 2;
  const unsigned grid = (unsigned)std::min<uint64_t>((total + teams - 1) / teams, (uint64_t)std::max(plan.dev.sm_count, 1));
  if (kT == 256) {
    OMB_LAUNCH(k_reassigned_r64x<256>, dim3(grid), dim3(256), smem_bytes<256>(), s, ra);
  } else {
    OMB_LAUNCH(k_reassigned_r64x<128>, dim3(grid), dim3(256), smem_bytes<128>(), s, ra);
  }
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
