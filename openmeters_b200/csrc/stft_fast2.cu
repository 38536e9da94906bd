// stft_fast2.cu — second-generation specialised kernel for the reassigned STFT at N = 4096 (BASELINE configs[1]).
//
// Same mathematics as stft_fast.cu (5 complex 4096-point FFTs per frame: packed real forward FFT, fused
// Hilbert pair step, one inverse, three windowed forward FFTs; see that file and DESIGN.md §4.1).  What changed
// is the mapping onto the SM, driven by the round-1 ncu captures of the first kernel (issue-slot bound at 56 %
// issue utilisation: L1 misses on the twiddle/window tables, register spills at 2 CTAs/SM):
//
//   * ONE 512-thread CTA per SM = two 256-thread groups, each running its own frame with its own named
//     barrier (bar.sync 1+g, 256) — 16 warps/SM like 2 CTAs, but
//   * the two groups work on CONSECUTIVE frames of the same lane and share one staging ring (H + 3*hop floats:
//     frames f, f+1 in flight, two hops prefetched by 16-byte async copies one iteration ahead),
//   * every table the inner loops read (W_4096 twiddles 30 KB, W_256 twiddles, window h, derivative window dh)
//     lives in shared memory, loaded once per CTA: no global loads in the frame loop except the ring prefetch,
//   * Im(c) of the analytic signal is parked in a 16 KB buffer Y straight from the pruned last inverse pass
//     (only its centre half is needed); Re(c) is rebuilt from the ring where it is used, so the 32-register
//     c[] array of the first kernel is gone,
//   * the pair step uses Q[k] = cos(th_k) conj(Z[M-k]) + j sin(th_k) Z[k], th_k = 2 pi k / H (8 flops per bin),
//   * last passes are pruned to the outputs that are used (9 of 16 bins per thread <= Nyquist; centre 8 of 16
//     of the inverse).
// Round 2 (profiles/r02_notes.md): the ring is filled by the TMA engine (bulk copies + mbarrier) and the compacted column leaves by
// one bulk store; and — the part that moved the metric — the SCHEDULE: every CTA walks one contiguous range of the linearised
// (lane, frame) sequence (one ring prime per lane touched), an iteration covers as many frame pairs as the ring holds next to the
// following prefetch (2 at hop 1024, 4 at hop <= 512), and the barrier between the two groups at the top of an iteration is a
// lagged one (bar.arrive by every warp, bar.sync only by the prefetching warp): 2.02e7 -> 2.18e7 frames/s without touching the
// arithmetic.
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

constexpr int kM = 4096;
constexpr int kGroupsPerCta = 2;
constexpr int kThreads = kT * kGroupsPerCta;
constexpr int kWarps = kT / 32;          // warps per group
constexpr int kWSize = f16::phys_size(kM);
constexpr int kBinGroups = 9;            // bins t + 256 j, j < 8, and bin 2048 (t = 0, j = 8)
constexpr int kDefaultBulk = 3;          // OMB_FAST2_BULK default (see launch_stft_fast2); with it, the lagged barrier (variant bit 6)

struct Fast2Args {
  StftKernelArgs a;
  const float2* tw1;   // global: [15][256] W_4096^{b q}
  const float2* tw2;   // global: [15][16]  W_256^{o q}
  uint32_t frames_per_run, runs_per_lane, ring_len;
  uint32_t pairs;      // frame pairs per loop iteration (one CTA barrier and one ring prefetch per iteration): as many as the ring holds
  uint64_t chunk;      // > 0: every CTA walks ONE contiguous range of `chunk` frames of the linearised (lane, frame) sequence (the ring is
                       // primed once per lane it touches) instead of runs dealt round-robin (a prime per run)
  float norm_ac, norm_dc;  // bin_norm[k] for 0 < k < N/2 and for k in {0, N/2} (window.rs:100-108)
};

struct GroupSmem {
  alignas(16) float2 W[kWSize];           // transform buffer; after the last transform of a frame: staging of the output column
  float Y[kM];                           // y[off + n] = Im c[n]
  int warp_cnt[kBinGroups * kWarps];
  int offs[kBinGroups * kWarps + 1];
  float x0_xm[2];
};

struct Smem2 {
  float2 tw1[4 * kT];   // rows q = 1, 2, 4, 8 of W_4096^{b q}; the other 11 are products (twiddle15<kTw = 1>)
  float2 tw2[15 * 16];
  float h[kM];
  float dh[kM];
  GroupSmem g[kGroupsPerCta];
  alignas(16) uint64_t mbar[2];          // [0]: completion barrier of the ring's bulk copies (kBulkIn)
  uint32_t tmem_base;                    // kTmemTab
  uint32_t pad_[3];
  // float ring[ring_len] follows
};
static_assert(sizeof(Smem2) % 16 == 0 && sizeof(GroupSmem) % 16 == 0, "bulk copies need 16-byte aligned shared addresses");
static_assert(sizeof(float2) * kWSize >= 12 * (kM / 2 + 1) + 16, "the transform buffer must hold a whole output column");

// Copies lane samples [s0, s1) (multiples of 4) into the ring at (sample index mod ring_len); all CTA threads.
__device__ __forceinline__ void ring_fetch(float* ring, int ring_mask, const float* x, uint64_t s0, uint64_t s1) {
  ring_fetch_pow2(ring, ring_mask, x, s0, s1, kThreads);
}

// kVariant bit 0: pass-2 twiddles computed from 4 loads (kTw2 = 1) instead of 15 table loads.
// kVariant bit 1: the F and I transforms keep their 256-point sub-transforms inside one half-warp (no group barrier
// between passes 2 and 3 / 1 and 2): frequency-domain ownership becomes tf + 256 j, tf = (t >> 4) + 16 (t & 15).
// (Pass-1 twiddles are always computed from the 4 rows kept in shared memory: measured fastest, and the 22 KB
// saved pay for the power-of-two ring.)
// kVariant bit 2: hop is a multiple of 4 but not of 512 (the settings UI's N/16 ... N/128): a frame's ring origin is then
// not 512-aligned, so the wrap is applied per thread (one more LOP3 per ring read) instead of per warp-uniform row.
// kVariant bit 3 (kBulkIn): the hop-overlapped staging ring is filled by the TMA engine — ONE elected thread issues one or two
// bulk copies (cp.async.bulk.shared.global, 2 hops = 8 KB contiguous per frame pair) completing on an mbarrier, instead of
// 512 per-thread 16-byte LDGSTS with their 64-bit loop and address arithmetic.
// kVariant bit 4 (kBulkOut): the compacted column is staged in the (then idle) transform buffer and leaves the SM as ONE bulk
// copy (cp.async.bulk.global.shared) instead of 27 predicated scalar STG per thread; the staging offset reproduces the slot's
// misalignment (slots are 12-byte multiples) so the 16-byte aligned body is one copy and <= 3 head / tail floats are stored
// by three threads each.
// kVariant bit 5 (kTmemTab): everything a thread reads from a TABLE — its 15 + 15 twiddles and its 16 + 16 window values — sits in the
// thread's own tensor-memory columns (tcgen05.st once, one tcgen05.ld per pass / window): 95 shared-memory loads, 48 window loads
// and 220 twiddle-product instructions per thread-frame become 13 TMEM loads.
// kVariant bit 6 (kLagSync, needs kBulkIn): the CTA barrier at the top of an iteration only exists so that the ring prefetch does not
// overwrite samples the slower group still reads.  It becomes a LAGGED barrier: every warp ARRIVES (bar.arrive, non-blocking) when it
// has finished the previous iteration, and only the warp that issues the prefetch WAITS (bar.sync) — the groups drift up to one
// iteration apart instead of being re-aligned every few frames.  Two named barriers and two mbarriers alternate by iteration
// parity, so an early arrival / a completed phase of iteration i + 1 cannot be taken for iteration i.
template <int kVariant>
__global__ void __launch_bounds__(kThreads, 1) k_reassigned_fast2(Fast2Args fa) {
  constexpr int kTw2 = kVariant & 1;
  constexpr bool kLocal = (kVariant & 2) != 0;
  constexpr bool kAnyHop = (kVariant & 4) != 0;
  constexpr bool kBulkIn = (kVariant & 8) != 0;
  constexpr bool kBulkOut = (kVariant & 16) != 0;
  constexpr bool kTmemTab = (kVariant & 32) != 0;
  constexpr bool kLagSync = (kVariant & 64) != 0 && kBulkIn;
  OMB_DYN_SMEM(unsigned char, smem_raw);
  Smem2& sm = *reinterpret_cast<Smem2*>(smem_raw);
  float* ring = reinterpret_cast<float*>(smem_raw + sizeof(Smem2));
  const StftKernelArgs& a = fa.a;
  const int tid = threadIdx.x, t = tid & (kT - 1), lane_id = t & 31, warp = t >> 5;
  const int g = __shfl_sync(0xffffffffu, tid >> 8, 0);  // warp-uniform by construction; tells the compiler so
  GroupSmem& gs = sm.g[g];
  const int hop = (int)a.hop, H = 2 * kM, ring_mask = (int)fa.ring_len - 1;  // ring_len is a power of two
  const int off = (H - kM) / 2;
  const uint64_t total_runs = (uint64_t)fa.runs_per_lane * a.n_lanes;
  const ReassignConsts rc{a.bin_hz, a.max_hz, a.inv_2pi, a.inv_hop, a.latency_hops};

  // ---- one-off: tables into shared memory
  for (int i = tid; i < 4 * kT; i += kThreads) {
    const int row = (1 << (i >> 8)) - 1;  // q - 1 for q = 1, 2, 4, 8
    sm.tw1[i] = __ldg(&fa.tw1[row * kT + (i & (kT - 1))]);
  }
  for (int i = tid; i < 15 * 16; i += kThreads) sm.tw2[i] = __ldg(&fa.tw2[i]);
  for (int i = tid; i < kM; i += kThreads) {
    sm.h[i] = __ldg(&a.win[i]);
    sm.dh[i] = __ldg(&a.dwin[i]);
  }
  // per-thread constants
  Addr ad;
  ad.pA = t + (t >> 4);
  ad.pB = 273 * (t >> 4) + (t & 15);
  ad.pC = 273 * (t & 15) + 17 * (t >> 4);
  const float2* tw1t = sm.tw1 + t;
  const float2* tw2o = sm.tw2 + (t & 15);
  const int tf = kLocal ? (t >> 4) + 16 * (t & 15) : t;  // this thread's bins between F and I: tf + 256 j
  Addr adf = ad;
  adf.pC = 273 * (tf & 15) + 17 * (tf >> 4);
  float cos_t, sin_t;  // th_tf = 2 pi tf / H
  {
    sincospif((float)tf / (float)kM, &sin_t, &cos_t);  // 2 pi tf / 8192 = pi * (tf / 4096); one-off per thread
  }
  const float sign = (t & 1) ? -1.0f : 1.0f;
  const float ramp0 = (float)t - (float)(kM - 1) * 0.5f;  // n - (N-1)/2 at j = 0
  const int pt = (kT - tf) & (kT - 1);
  const int pPartner = 273 * (pt & 15) + 17 * (pt >> 4);
  if (kBulkIn && tid == 0) {
    mbar_init(&sm.mbar[0], 1);
    mbar_init(&sm.mbar[1], 1);
  }
  unsigned ring_phase = 0;  // parity of the mbarrier phase the next wait is for (!kLagSync: one mbarrier)
  unsigned it = 0;          // kLagSync: iteration counter; iteration `it` uses mbarrier it & 1 (phase parity (it >> 1) & 1) and named barrier 3 + (it & 1)
  if (kTmemTab && tid < 32) tmem_alloc(&sm.tmem_base);
  if (kTmemTab) tmem_fence_before_sync();
  __syncthreads();
  TwTmem twm{0u, 0u};
  uint32_t tm_h = 0u, tm_dh = 0u;
  if (kTmemTab) {
    tmem_fence_after_sync();
    // a warp owns 128 columns of its lane quadrant: [0,32) pass-1 twiddles, [32,64) pass-2 twiddles, [64,80) h, [80,96) dh
    const uint32_t tb = tmem_addr(sm.tmem_base, (uint32_t)(tid >> 7) * 128u);
    twm.t1 = tb;
    twm.t2 = tb + 32u;
    tm_h = tb + 64u;
    tm_dh = tb + 80u;
    float w1[32], w2[32], hv[16], dv[16];
#pragma unroll
    for (int q = 1; q < 16; ++q) {
      const float2 a1 = __ldg(&fa.tw1[(q - 1) * kT + t]);
      const float2 a2 = __ldg(&fa.tw2[(q - 1) * 16 + (t & 15)]);
      w1[2 * (q - 1)] = a1.x;
      w1[2 * (q - 1) + 1] = a1.y;
      w2[2 * (q - 1)] = a2.x;
      w2[2 * (q - 1) + 1] = a2.y;
    }
    w1[30] = w1[31] = w2[30] = w2[31] = 0.0f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      hv[j] = __ldg(&a.win[t + kT * j]);
      dv[j] = __ldg(&a.dwin[t + kT * j]);
    }
    tmem_st<32>(twm.t1, w1);
    tmem_st<32>(twm.t2, w2);
    tmem_st<16>(tm_h, hv);
    tmem_st<16>(tm_dh, dv);
    tmem_wait_st();
  }

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
    // prime: everything the frames of the first iteration read (2 * pairs frames)
    const uint64_t iter_frames = 2ull * fa.pairs, iter_span = (iter_frames - 1) * (uint64_t)hop;
    {
      const uint64_t s0 = f_begin * (uint64_t)hop;
      const uint64_t s1 = s0 + (uint64_t)H + iter_span < s_end ? s0 + (uint64_t)H + iter_span : s_end;
      if (kBulkIn) {
        if (tid == 0) {
          uint64_t* mb = &sm.mbar[kLagSync ? (it & 1u) : 0u];
          mbar_expect_tx(mb, ring_bytes(s0, s1));
          ring_fetch_bulk(ring, ring_mask, x, s0, s1, mb);
        }
      } else {
        ring_fetch(ring, ring_mask, x, s0, s1);
        async_commit();
      }
    }
    for (uint64_t fa0 = f_begin; fa0 < f_end; fa0 += iter_frames) {
      if (kLagSync) {
        mbar_wait(&sm.mbar[it & 1u], (it >> 1) & 1u);  // this iteration's samples have landed
      } else if (kBulkIn) {
        mbar_wait(&sm.mbar[0], ring_phase);
        ring_phase ^= 1u;
      } else {
        async_wait_all();
      }
      if (kBulkOut && t == 0) bulk_wait_read();  // the previous column has left the transform buffer
      if (kLagSync) {
        if (kBulkOut) group_sync(g);  // ... for every thread of the group
        // everybody has finished iteration it - 1; only the prefetching warp needs to know
#ifdef OMB_EMU
        if ((tid >> 5) == 0) omb_emu::named_sync(3 + (int)(it & 1u), kThreads); else omb_emu::named_arrive(3 + (int)(it & 1u), kThreads);
#else
        if ((tid >> 5) == 0) asm volatile("bar.sync %0, %1;" ::"r"(3 + (int)(it & 1u)), "n"(kThreads) : "memory");
        else asm volatile("bar.arrive %0, %1;" ::"r"(3 + (int)(it & 1u)), "n"(kThreads) : "memory");
#endif
      } else {
        __syncthreads();  // ring holds the frames of this iteration; both groups are done with the previous one
      }
      {                 // prefetch what the next iteration adds: 2 * pairs hops
        const uint64_t s0 = fa0 * (uint64_t)hop + (uint64_t)H + iter_span;
        const uint64_t s1 = s0 + iter_frames * hop < s_end ? s0 + iter_frames * hop : s_end;
        if (kBulkIn) {
          if (tid == 0) {  // every phase is armed (with 0 bytes at the end of a run) so the parity bookkeeping stays uniform
            uint64_t* mb = &sm.mbar[kLagSync ? ((it + 1u) & 1u) : 0u];
            mbar_expect_tx(mb, ring_bytes(s0, s1));
            ring_fetch_bulk(ring, ring_mask, x, s0, s1, mb);
          }
        } else {
          if (s0 < s1) ring_fetch(ring, ring_mask, x, s0, s1);
          async_commit();
        }
      }
#pragma unroll 1
      for (uint32_t sp = 0; sp < fa.pairs; ++sp) {
      const uint64_t f = fa0 + 2ull * sp + g;
      if (kBulkOut && sp > 0) {  // the previous column of this group must have left the transform buffer (bulk store) before it is reused
        if (t == 0) bulk_wait_read();
        group_sync(g);
      }
      if (f < f_end) {
        const int r0 = (int)(f * (uint64_t)hop) & ring_mask;  // multiple of 512 unless kAnyHop (then of 4)
        float2 v[16];
        // ---- F: z[n] = x[2n] + j x[2n+1], n = t + 256 j
        // ring position of sample s is s & mask; (r0 + 512 j) & mask is warp-uniform, 2t < 512 never carries
#pragma unroll
        for (int j = 0; j < 16; ++j)
          v[j] = kAnyHop ? *reinterpret_cast<const float2*>(ring + ((r0 + 2 * kT * j + 2 * t) & ring_mask))
                         : *reinterpret_cast<const float2*>(ring + ((r0 + 2 * kT * j) & ring_mask) + 2 * t);
        if (kTmemTab) fft_forward_tmem<f16::kAll, kLocal>(v, gs.W, twm, adf, g);
        else fft_forward<f16::kAll, kTw2, kLocal>(v, gs.W, tw1t, tw2o, adf, g);
#pragma unroll
        for (int q = 0; q < 16; ++q) gs.W[adf.pC + q] = v[q];
        if (t == 0) {
          gs.x0_xm[0] = v[0].x + v[0].y;
          gs.x0_xm[1] = v[0].x - v[0].y;
        }
        group_sync(g);
        // ---- X: Q[k] = cos(th_k) conj(Z[M-k]) + j sin(th_k) Z[k], k = t + 256 j, th_k = th_t + 2 pi j / 32
        {
          const float2* wp = gs.W + pPartner;
          float2 zp[16];
          if (t == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) zp[j] = wp[(16 - j) & 15];
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) zp[j] = wp[15 - j];
          }
          group_sync(g);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            // (ck, sk) = e^{j th_t} e^{j 2 pi j / 32} (immediate operands);  Q = ck conj(zp) + sk (j z): FMUL2 + FFMA2
            const float cj = f16::kCos32[j], sj = f16::kSin32[j];
            const float ck = cos_t * cj - sin_t * sj;
            const float sk = sin_t * cj + cos_t * sj;
            const float2 z = v[j];
            v[j] = make_float2(ck * zp[j].x - sk * z.y, sk * z.x - ck * zp[j].y);
          }
          if (t == 0) v[0] = make_float2(0.0f, 0.0f);
        }
        // ---- I: inverse, DIT, thread t holds Q[t + 256 j]; result y -> gs.Y
        {
          f16::dft16<true>(v);
          float2* wc = gs.W + adf.pC;
#pragma unroll
          for (int q = 0; q < 16; ++q) wc[q] = v[q];
          if (kLocal) __syncwarp(); else group_sync(g);
          float2* wb = gs.W + ad.pB;
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = wb[17 * j];
          if (kTmemTab) twiddle15_tmem<true>(v, twm.t2);
          else twiddle15<true, kTw2, false>(v, tw2o, 16);
          f16::dft16<true>(v);
#pragma unroll
          for (int q = 0; q < 16; ++q) wb[17 * q] = v[q];
          group_sync(g);
          const float2* wa = gs.W + ad.pA;
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = wa[273 * j];
          if (kTmemTab) twiddle15_tmem<true>(v, twm.t1);
          else twiddle15<true, 1, true>(v, tw1t, kT);
          f16::dft16p<true, f16::kMid8>(v);
          // q[m], m = t + 256 m2, m2 = 4..11 -> Y as float2[m - 1024]
          float2* y2 = reinterpret_cast<float2*>(gs.Y) + t;
#pragma unroll
          for (int q = 4; q < 12; ++q) y2[kT * (q - 4)] = v[q];
          group_sync(g);
        }
        // ---- G: three windowed transforms of c[n] = (M x[off+n] + bias) + j Y[n]
        const float bias = sign * 0.5f * gs.x0_xm[1] - 0.5f * gs.x0_xm[0];
        float2 S[kBinGroups];
        float nd[kBinGroups];
#pragma unroll 1
        for (int wsel = 0; wsel < 3; ++wsel) {
          const float* win = (wsel == 1 ? sm.dh : sm.h) + t;
          float wreg[16];
          if (kTmemTab) tmem_ld<16>(wsel == 1 ? tm_dh : tm_h, wreg);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float wv = kTmemTab ? wreg[j] : win[kT * j];
            if (wsel == 2) wv *= ramp0 + (float)(kT * j);       // t*h window, processor.rs:601-608
            const float xs = kAnyHop ? ring[(r0 + off + kT * j + t) & ring_mask] : ring[((r0 + off + kT * j) & ring_mask) + t];
            const float cx = fmaf((float)kM, xs, bias);
            v[j] = f16::cscale2(make_float2(cx, gs.Y[t + kT * j]), wv);
          }
          if (kTmemTab) fft_forward_tmem<f16::kFirst9>(v, gs.W, twm, ad, g);
          else fft_forward<f16::kFirst9, kTw2>(v, gs.W, tw1t, tw2o, ad, g);
          if (wsel == 0) {
#pragma unroll
            for (int j = 0; j < kBinGroups; ++j) S[j] = v[j];
          } else if (wsel == 1) {
#pragma unroll
            for (int j = 0; j < kBinGroups; ++j) nd[j] = v[j].y * S[j].x - v[j].x * S[j].y;
          }
          group_sync(g);  // pass-3 reads done before the next pass-1 stores (and before warp_cnt reuse)
        }
        // ---- R: reassignment + ordered compaction
        omb_spectrogram_point pts[kBinGroups];
        int rank[kBinGroups];
        unsigned keep = 0;
#pragma unroll
        for (int j = 0; j < kBinGroups; ++j) {
          const int bin = t + kT * j;
          const float norm = (bin == 0 || j == 8) ? fa.norm_dc : fa.norm_ac;
          const bool k = reassign_bin_nd(S[j], nd[j], v[j], norm, bin, rc, &pts[j]) & (j < 8 || t == 0);  // branch-free
          const unsigned m = __ballot_sync(0xffffffffu, k);
          if (lane_id == 0) gs.warp_cnt[j * kWarps + warp] = __popc(m);
          if (k) keep |= 1u << j;
          rank[j] = __popc(m & ((1u << lane_id) - 1u));
        }
        group_sync(g);
        if (warp == 0) {
          int c0 = 0, c1 = 0, c2 = 0;
          const int i0 = lane_id * 3;
          if (i0 + 0 < kBinGroups * kWarps) c0 = gs.warp_cnt[i0 + 0];
          if (i0 + 1 < kBinGroups * kWarps) c1 = gs.warp_cnt[i0 + 1];
          if (i0 + 2 < kBinGroups * kWarps) c2 = gs.warp_cnt[i0 + 2];
          const int tot = c0 + c1 + c2;
          int incl = tot;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane_id >= o) incl += n;
          }
          const int excl = incl - tot;
          if (i0 + 0 < kBinGroups * kWarps) gs.offs[i0 + 0] = excl;
          if (i0 + 1 < kBinGroups * kWarps) gs.offs[i0 + 1] = excl + c0;
          if (i0 + 2 < kBinGroups * kWarps) gs.offs[i0 + 2] = excl + c0 + c1;
          if (lane_id == 31) gs.offs[kBinGroups * kWarps] = incl;
        }
        group_sync(g);
        {
          const uint64_t slot = lane * a.frames_per_lane + f;
          float* out = reinterpret_cast<float*>(a.out_points + slot * a.point_stride);
          // kBulkOut: same stores, into the idle transform buffer at the slot's own offset from a 16-byte boundary
          const int mis = kBulkOut ? (int)((reinterpret_cast<uintptr_t>(out) >> 2) & 3u) : 0;
          float* dst = kBulkOut ? reinterpret_cast<float*>(gs.W) + mis : out;
#pragma unroll
          for (int j = 0; j < kBinGroups; ++j)
            if (keep & (1u << j)) {
              float* o = dst + 3 * (gs.offs[j * kWarps + warp] + rank[j]);
              o[0] = pts[j].time_offset;
              o[1] = pts[j].freq_hz;
              o[2] = pts[j].power;
            }
          const int count = gs.offs[kBinGroups * kWarps];
          if (t == 0) a.out_counts[slot] = (uint32_t)count;
          if (kBulkOut) {
            fence_async_smem();  // this thread's staging stores -> visible to the copy engine
            group_sync(g);
            const int n = 3 * count;                       // floats of the column
            const int head = ((4 - mis) & 3) < n ? ((4 - mis) & 3) : n;       // floats before the first 16-byte boundary of the slot
            const int body = (n - head) & ~3;
            const int tail = n - head - body;
            if (t == 0 && body > 0) {
              bulk_s2g(out + head, dst + head, (unsigned)body * 4u);
              bulk_commit();
            }
            if (t >= 32 && t < 35 && t - 32 < head) out[t - 32] = dst[t - 32];
            if (t >= 64 && t < 67 && t - 64 < tail) out[head + body + t - 64] = dst[head + body + t - 64];
          }
        }
      }
      }  // sp
      ++it;
    }
    if (kLagSync) {
      mbar_wait(&sm.mbar[it & 1u], (it >> 1) & 1u);  // the last (empty) phase armed inside the loop
      ++it;
    } else if (kBulkIn) {
      mbar_wait(&sm.mbar[0], ring_phase);  // the last (empty) phase armed inside the loop
      ring_phase ^= 1u;
    } else {
      async_wait_all();
    }
    __syncthreads();
  }
  if (kBulkOut && t == 0) bulk_wait_read();  // shared memory must outlive the copies that read it
  if (kTmemTab) {
    __syncthreads();
    if (tid < 32) tmem_free(sm.tmem_base);
  }
}

uint32_t ring_len_for(uint64_t hop) { return (uint32_t)next_pow2(2 * (uint64_t)kM + 3 * hop); }
size_t smem_bytes(uint64_t hop) { return sizeof(Smem2) + (size_t)ring_len_for(hop) * sizeof(float); }

}  // namespace

bool stft_fast2_supported(const StftConfig& cfg, const DeviceInfo& dev) {
  if (!cfg.reassign || cfg.window != (uint64_t)kM || cfg.zero_pad != 1) return false;
  // hop: multiples of 512 up to 2048 (warp-uniform ring rows), or any multiple of 4 below 512 (per-thread wrap; the
  // 16-byte async copies need hop % 4 == 0) — of the UI's N/4 ... N/128 only N/6 = 682 falls to the shared-memory tier
  if (cfg.hop > 2048 || cfg.hop < 4 || (cfg.hop % 4) != 0) return false;
  if (cfg.hop >= 512 && (cfg.hop % 512) != 0) return false;
  return dev.max_smem_optin == 0 || smem_bytes(cfg.hop) <= (size_t)dev.max_smem_optin;
}

int stft_fast2_prepare(StftPlan& plan) {
  const int smem = (int)smem_bytes(plan.cfg.hop);
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_fast2<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_fast2<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_fast2<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_fast2<14>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_fast2<26>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_fast2<30>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_fast2<90>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_fast2<94>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_fast2<58>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_fast2<62>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  return OMB_OK;
}

int launch_stft_fast2(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s) {
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  if (per_lane == 0 || a.n_lanes == 0) return OMB_OK;
  if ((reinterpret_cast<uintptr_t>(a.lanes) & 15u) != 0 || (a.lane_stride % 4) != 0)
    return fail(OMB_ERR_INVALID, "specialised STFT kernel needs 16-byte aligned lanes (pointer and lane_stride % 4 == 0)");
  Fast2Args fa{};
  fa.a = a;
  fa.tw1 = plan.d_fast_tables.ptr;
  fa.tw2 = fa.tw1 + 15 * kT;
  fa.ring_len = ring_len_for(a.hop);
  fa.norm_ac = plan.h_norm.size() > 1 ? plan.h_norm[1] : plan.h_norm[0];
  fa.norm_dc = plan.h_norm[0];
  const uint64_t ctas = (uint64_t)std::max(plan.dev.sm_count, 1);
  uint64_t run = 128;  // even, so both groups stay busy; long enough to amortise the ring prime
  if (a.hop < 512) run *= 512 / a.hop;  // small hops: the prime is worth 8192 / hop frames of new samples
  while (run > 16 && ((per_lane + run - 1) / run) * a.n_lanes < ctas * 6) run >>= 1;
  fa.frames_per_run = (uint32_t)std::min<uint64_t>(run, per_lane);
  fa.runs_per_lane = (uint32_t)((per_lane + fa.frames_per_run - 1) / fa.frames_per_run);
  const uint64_t total_runs = (uint64_t)fa.runs_per_lane * a.n_lanes;
  unsigned grid = (unsigned)std::min<uint64_t>(total_runs, ctas);
  // frame pairs per iteration: as many as fit the ring next to the prefetch of the following iteration, H + (4 pairs - 1) hop <= ring
  // (hop 1024: 2, hop <= 512: 4, hop 2048: 1); OMB_FAST2_PAIRS pins it (1 = one CTA barrier per pair, round 1)
  static const int pairs_env = [] { const char* e = getenv("OMB_FAST2_PAIRS"); return e ? atoi(e) : 0; }();
  uint32_t pairs = (uint32_t)std::min<uint64_t>(4, ((uint64_t)fa.ring_len - 2 * kM + a.hop) / (4 * (uint64_t)a.hop));
  if (pairs_env > 0) pairs = std::min<uint32_t>(pairs, (uint32_t)pairs_env);
  fa.pairs = std::max<uint32_t>(1, pairs);
  // Large batches: one contiguous range of frames per CTA (even length: the two groups work on frame pairs), balanced to a pair and
  // primed once per lane touched; OMB_FAST2_CONTIG=0 keeps the round-robin runs (measured: profiles/r02_notes.md)
  static const bool contig_env = [] { const char* e = getenv("OMB_FAST2_CONTIG"); return !(e && e[0] == '0'); }();
  const uint64_t total_frames = per_lane * a.n_lanes;
  fa.chunk = 0;
  if (contig_env && total_frames >= ctas * 96) {
    fa.chunk = ((total_frames + ctas - 1) / ctas + 1) & ~1ull;
    grid = (unsigned)((total_frames + fa.chunk - 1) / fa.chunk);
  }
  // OMB_FAST2_BULK: 0 = per-thread LDGSTS ring + scalar column stores (round 1), 1 = bulk ring fill, 3 = bulk ring fill and bulk
  // column store (default: kDefaultBulk)
  static const int bulk = [] { const char* e = getenv("OMB_FAST2_BULK"); return e ? atoi(e) & 7 : kDefaultBulk; }();
  const size_t smem = smem_bytes(a.hop);
  const bool any_hop = (a.hop % 512) != 0;
  // lagged cross-group barrier: measured +2.9 % on cfg2 (profiles/r02_notes.md); OMB_FAST2_LAG=0 restores the CTA barrier
  static const bool lag_sync = [] { const char* e = getenv("OMB_FAST2_LAG"); return !(e && e[0] == '0'); }();
#define OMB_FAST2_LAUNCH(V) OMB_LAUNCH(k_reassigned_fast2<V>, dim3(grid), dim3(kThreads), smem, s, fa)
  if (bulk == 0) {
    if (any_hop) OMB_FAST2_LAUNCH(6); else OMB_FAST2_LAUNCH(2);
  } else if (bulk == 1) {
    if (any_hop) OMB_FAST2_LAUNCH(14); else OMB_FAST2_LAUNCH(10);
  } else if (bulk == 3 && lag_sync) {  // + lagged cross-group barrier
    if (any_hop) OMB_FAST2_LAUNCH(94); else OMB_FAST2_LAUNCH(90);
  } else if (bulk == 7) {  // + tables in tensor memory
    if (any_hop) OMB_FAST2_LAUNCH(62); else OMB_FAST2_LAUNCH(58);
  } else {
    if (any_hop) OMB_FAST2_LAUNCH(30); else OMB_FAST2_LAUNCH(26);
  }
#undef OMB_FAST2_LAUNCH
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
