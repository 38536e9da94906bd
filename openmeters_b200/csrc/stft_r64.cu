// stft_r64.cu — third-generation kernel for the reassigned STFT at N = 4096 (BASELINE configs[1], the metric path;
// spectrogram/processor.rs:439-608): the same mathematics as stft_fast2.cu (packed real forward FFT, fused Hilbert pair
// step, one inverse, three windowed forward transforms = five complex 4096-point transforms per frame), re-mapped onto
// the SM after the round-2 ncu capture of that kernel (profiles/r02f_ncu_full_fast2.json): FMA pipe 52 %, issue slots
// 62 %, shared-memory wavefronts 67 % — three co-limiters, the largest being shared memory, because a radix-16 engine
// with 16 values per thread needs TWO exchanges per transform plus two twiddle passes.
//
//   * 4096 = 64 x 64: a transform is two in-register radix-64 butterflies (fft64.cuh) with ONE exchange through shared
//     memory; a transform belongs to a TEAM of 64 threads (two warps), each thread holding 64 complex values.
//   * Four teams per CTA (256 threads, one CTA per SM, up to 255 registers per thread) work on four different frames and
//     never synchronise with each other: team barriers only (bar.sync id, 64), so the LSU phases of one team overlap the
//     FMA phases of the others at a much finer grain than two 8-warp groups could.
//   * No staging ring: the frame's 8192 samples are fetched by ONE bulk copy of the TMA engine (cp.async.bulk, 32 KB,
//     mbarrier completion) straight into the team's exchange buffer while the team is busy with the epilogue of its
//     previous frame; consecutive frames overlap in L2, so HBM still sees every sample once.
//   * Everything a thread must carry across a transform lives in TENSOR MEMORY (tmem_park.cuh): the centre samples
//     x[2048 + t + 64 j] (read once from the landed frame, used by the three windowed transforms), the spectrum S of
//     the h-windowed transform and the cross term nd — all of them come back to the thread that parked them, which is
//     what tcgen05.st / tcgen05.ld .32x32b do.  TMEM is otherwise idle in this kernel (no MMA).  Im c, which the inverse
//     transform leaves in the wrong threads, crosses shared memory once (through the idle exchange buffer) and is then
//     parked as the complex analysis input c[n] — 128 columns that each windowed transform reloads with ONE tcgen05.ld.
//     Shared memory: 4 x 33 KB exchange + 32 KB windows + 7 KB twiddles = 172 KB.
//   * The inverse transform is the forward one on conjugated data, so ONE copy of the transform code serves all five
//     transforms of a frame (loop over the transform index; keeps the loop body inside the instruction cache).
// Any hop that is a multiple of 4 (16-byte aligned bulk copies); there is no ring and hence no hop-specific variant.
#ifndef OMB_F32X2_CMUL
#define OMB_F32X2_CMUL 0
#endif
#ifndef OMB_R64_PC
#define OMB_R64_PC 8      // partner values of the pair step in flight per chunk (A/B: profiles/r02_notes.md)
#endif
#ifndef OMB_R64_PRUNE
#define OMB_R64_PRUNE 0   // 1: last radix-4 step of pass B pruned to the outputs used (three code copies; measured slower)
#endif
#include <cstdlib>

#include "async_copy.cuh"
#include "device_math.cuh"
#include "fft64.cuh"
#include "stft.h"
#include "tmem_park.cuh"

namespace omb {

namespace {

constexpr int kM = 4096;                 // complex points per transform = window length N
constexpr int kH = 2 * kM;               // samples per frame (Hilbert block)
constexpr int kOff = (kH - kM) / 2;      // first centre sample
constexpr int kTeam = 64;                // threads per team
constexpr int kTeams = 4;
constexpr int kThreads = kTeam * kTeams;
constexpr int kRS = 66;                  // row stride of the exchange buffer (float2): column writes (64-bit) and row reads (128-bit) conflict-free
constexpr int kGroups = 33;              // bins t + 64 j, j < 32, and bin 2048 (t = 0, j = 32)
constexpr unsigned kFrameBytes = kH * sizeof(float);
// TMEM / scratch columns of one warp (a warp owns 256 of the 512 columns of its lane quadrant)
constexpr int kColC = 0;      // 64 centre samples x[off + n], later 64 complex c[n] (128 columns), n = t + 64 j
constexpr int kColS = 128;    // S[t + 64 j], j < 32 (64 columns)
constexpr int kColNd = 192;   // nd, j < 32
constexpr int kColSn = 224;   // S[2048] (2 columns), then nd[2048] at kColSn + 2
constexpr int kColsPerWarp = 256;

struct R64Args {
  StftKernelArgs a;
  const float2* tw;     // global: [14][64]: rows 0..6 = W_4096^{t q}, q = 1..7; rows 7..13 = W_4096^{8 t q}, q = 1..7
  float* scratch;       // global park (kTmem = false): [CTA][256 columns][256 threads]
  float norm_ac, norm_dc;
};

struct TeamSmem {
  alignas(16) float2 W[kTeam * kRS];  // landing zone of the frame (8192 floats, linear) / exchange buffer [row][66] / Im c staging
  int2 cnt[kGroups + 1];              // kept points of (warp 0, warp 1) per bin group
  float x0_xm[2];
  alignas(8) uint64_t mbar;
};

struct Smem {
  TeamSmem team[kTeams];
  float h[kM];
  float dh[kM];
  float2 tw[14 * kTeam];
  uint32_t tmem_base;
  uint32_t pad_[3];
};
static_assert(sizeof(float2) * kTeam * kRS >= kFrameBytes, "the exchange buffer must hold a whole frame");
static_assert(sizeof(TeamSmem) % 16 == 0 && sizeof(Smem) % 16 == 0, "bulk copies need 16-byte aligned shared addresses");

__device__ __forceinline__ void team_sync(int team) {
#ifdef OMB_EMU
  omb_emu::named_sync(1 + team, kTeam);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "n"(kTeam) : "memory");
#endif
}
__device__ __forceinline__ void sched_fence() {  // keeps ptxas from hoisting the next chunk's loads above this point (register pressure)
#ifndef OMB_EMU
  asm volatile("" ::: "memory");
#endif
}

__device__ __forceinline__ float2 cmul_s(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// v[q] *= W_4096^{t q}, q = 1..63: 14 table loads (q = 1..7 and 8, 16, ..., 56) + 49 products of two table entries.
__device__ __forceinline__ void twiddle63(float2 (&v)[64], const float2* tab) {
  float2 lo[8], hi[8];
#pragma unroll
  for (int i = 1; i < 8; ++i) {
    lo[i] = tab[(i - 1) * kTeam];
    hi[i] = tab[(6 + i) * kTeam];
  }
#pragma unroll
  for (int q = 1; q < 64; ++q) {
    const int a = q & 7, b = q >> 3;
    const float2 w = b == 0 ? lo[a] : (a == 0 ? hi[b] : cmul_s(lo[a], hi[b]));
    v[q] = f16::mul_tw<false>(v[q], w);
  }
}

// One 4096-point forward transform of the team: thread t enters with elements t + 64 j and leaves with bins t + 64 k.
// Pass A is decimation in frequency (its last step is four independent 16-point transforms whose outputs can leave as
// they complete), pass B decimation in time (its first step is four 16-point transforms that can start as their inputs
// arrive): measured +1.7 % over DIF / DIF on the transform alone.  The barrier before the stores orders them after every
// thread's previous reads of W (frame samples, partner rows, the previous exchange).
__device__ __forceinline__ void transform(float2 (&v)[64], float2* W, const float2* twt, int t, int team) {
  f64pt::dft64<false>(v);
  twiddle63(v, twt);
  team_sync(team);
  float2* wr = W + t;
#pragma unroll
  for (int q = 0; q < 64; ++q) wr[q * kRS] = v[q];
  team_sync(team);
  const float4* rd = reinterpret_cast<const float4*>(W + t * kRS);
#pragma unroll
  for (int s = 0; s < 32; ++s) {
    const float4 p = rd[s];
    v[2 * s] = make_float2(p.x, p.y);
    v[2 * s + 1] = make_float2(p.z, p.w);
  }
  f64pt::dit_front<false>(v);  // the caller finishes with dit_final<kPrune>: the last radix-4 step, pruned to the outputs it uses
}

template <bool kTmem, int N>
__device__ __forceinline__ void park_st(uint32_t tcol, float* scratch, int col, const float (&r)[N]) {
  if (kTmem) {
    tmem_st<N>(tcol + col, r);
    tmem_wait_st();
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) scratch[(size_t)(col + i) * kThreads] = r[i];
  }
}
template <bool kTmem, int N>
__device__ __forceinline__ void park_ld(uint32_t tcol, const float* scratch, int col, float (&r)[N]) {
  if (kTmem) {
    tmem_ld<N>(tcol + col, r);  // includes tcgen05.wait::ld
  } else {
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = scratch[(size_t)(col + i) * kThreads];
  }
}

template <bool kTmem>
__global__ void __launch_bounds__(kThreads, 1) k_reassigned_r64(R64Args ra) {
  constexpr int kCtaThreads = kThreads;
  OMB_DYN_SMEM(unsigned char, smem_raw);
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  const StftKernelArgs& a = ra.a;
  const int tid = threadIdx.x, t = tid & (kTeam - 1), lane_id = tid & 31;
  const int team = __shfl_sync(0xffffffffu, tid >> 6, 0);   // warp-uniform by construction; tells the compiler so
  const int wt = __shfl_sync(0xffffffffu, (tid >> 5) & 1, 0);  // warp within the team
  TeamSmem& ts = sm.team[team];
  const ReassignConsts rc{a.bin_hz, a.max_hz, a.inv_2pi, a.inv_hop, a.latency_hops};

  // ---- one-off: tables, TMEM, barriers
  for (int i = tid; i < kM; i += kCtaThreads) {
    sm.h[i] = __ldg(&a.win[i]);
    sm.dh[i] = __ldg(&a.dwin[i]);
  }
  for (int i = tid; i < 14 * kTeam; i += kCtaThreads) sm.tw[i] = __ldg(&ra.tw[i]);
  if (kTmem && tid < 32) tmem_alloc(&sm.tmem_base);
  if (t == 0) mbar_init(&ts.mbar, 1);
  if (kTmem) tmem_fence_before_sync();
  __syncthreads();
  if (kTmem) tmem_fence_after_sync();
  const uint32_t tbase = kTmem ? sm.tmem_base : 0u;
  const uint32_t tcol = kTmem ? tmem_addr(tbase, (uint32_t)(tid >> 7) * kColsPerWarp) : 0u;
  float* scratch = kTmem ? nullptr : ra.scratch + (size_t)blockIdx.x * kColsPerWarp * kThreads + tid;

  // per-thread constants
  const float2* twt = sm.tw + t;
  float cos_t, sin_t;  // th_t = 2 pi t / H
  sincospif((float)t / (float)kM, &sin_t, &cos_t);
  const float sign = (t & 1) ? -1.0f : 1.0f;
  const float ramp0 = (float)t - (float)(kM - 1) * 0.5f;  // n - (N-1)/2 at j = 0
  const int pt = (kTeam - t) & (kTeam - 1);               // the thread that holds the mirror bins M - k
  const unsigned lt_mask = (1u << lane_id) - 1u;

  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total = per_lane * a.n_lanes;
  const uint64_t stride = (uint64_t)gridDim.x * kTeams;
  uint64_t g = (uint64_t)blockIdx.x * kTeams + team;  // frames are dealt round-robin: CTAs stay within a few hundred frames of each other (L2)
  unsigned phase = 0;
  auto frame_src = [&](uint64_t gi) {
    const uint64_t l = gi / per_lane, f = a.first_frame + gi % per_lane;
    return a.lanes + l * a.lane_stride + f * (uint64_t)a.hop;
  };
  if (g < total && t == 0) {
    mbar_expect_tx(&ts.mbar, kFrameBytes);
    bulk_g2s(ts.W, frame_src(g), kFrameBytes, &ts.mbar);
  }

  for (; g < total; g += stride) {
    const uint64_t lane = g / per_lane, f = a.first_frame + g % per_lane;
    mbar_wait(&ts.mbar, phase);  // the frame has landed in W
    phase ^= 1u;
    float2 v[64];
    {
      const float* wf = reinterpret_cast<const float*>(ts.W);
      // centre samples of this thread's analysis points n = t + 64 j -> park (they become Re c after the forward transform)
      float xc[64];
#pragma unroll
      for (int j = 0; j < 64; ++j) xc[j] = wf[kOff + t + kTeam * j];
      park_st<kTmem, 64>(tcol, scratch, kColC, xc);
      // F input: z[n] = x[2n] + j x[2n+1], n = t + 64 j
      const float2* wz = ts.W + t;
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = wz[kTeam * j];
    }
#pragma unroll 1
    for (int tr = 0; tr < 5; ++tr) {
      if (tr == 1) {
        // ---- X: Q[k] = cos(th_k) conj(Z[M-k]) + j sin(th_k) Z[k], k = t + 64 j, th_k = th_t + 2 pi j / 128; the inverse
        //      transform runs as conj(forward(conj Q)), so conj(Q) is what enters the transform
        team_sync(team);
        float4* row = reinterpret_cast<float4*>(ts.W + t * kRS);
#pragma unroll
        for (int j = 0; j < 32; ++j) row[j] = make_float4(v[2 * j].x, v[2 * j].y, v[2 * j + 1].x, v[2 * j + 1].y);
        if (t == 0) {
          ts.x0_xm[0] = v[0].x + v[0].y;
          ts.x0_xm[1] = v[0].x - v[0].y;
        }
        team_sync(team);
        // partner of k = t + 64 j: row 64 - t, column 63 - j; for t = 0 row 0, column 64 - j (j = 0 reads a don't-care: DC is zeroed)
        const float2* prow = t == 0 ? ts.W + 1 : ts.W + pt * kRS;
        constexpr int kPc = OMB_R64_PC;  // partner values in flight: two chunks of 8 (register budget: v holds 128)
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
          sched_fence();
        }
        if (t == 0) v[0] = make_float2(0.0f, 0.0f);
      } else if (tr >= 2) {
        // ---- G input: c[n] w[n], w = h, dh, t*h (processor.rs:601-608: (n - (N-1)/2) * h[n], formed on the fly, bit-identical)
        float c[128];
        park_ld<kTmem, 128>(tcol, scratch, kColC, c);
        if (tr == 4) {
          const float* tab = sm.h + t;
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const float wv = (ramp0 + (float)(kTeam * j)) * tab[kTeam * j];
            v[j] = f16::cscale2(make_float2(c[2 * j], c[2 * j + 1]), wv);
          }
        } else {
          const float* tab = (tr == 3 ? sm.dh : sm.h) + t;
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = f16::cscale2(make_float2(c[2 * j], c[2 * j + 1]), tab[kTeam * j]);
        }
      }
      transform(v, ts.W, twt, t, team);
      if (tr == 0 || !OMB_R64_PRUNE) {
        f64pt::dit_final<false, f16::kAll>(v);
      } else if (tr == 1) {
        f64pt::dit_final<false, f16::kMid8>(v);    // only the centre half, outputs 16..47
      } else {
        f64pt::dit_final<false, f16::kFirst9>(v);  // only bins <= Nyquist, outputs 0..32
      }
      if (tr == 1) {
        // ---- centre half of the inverse: y[m] = conj(v), m = t + 64 n1, n1 = 16..47, carries Im c[2(m - 1024)], Im c[2(m - 1024) + 1]:
        //      redistributed through W (free between two exchanges) to the threads that own n = t + 64 j, combined with the parked
        //      centre samples into c[n] = (M x[off + n] + bias) + j Im c[n], and parked for the three windowed transforms
        team_sync(team);
        float2* y2 = ts.W + t;
#pragma unroll
        for (int q = 16; q < 48; ++q) y2[kTeam * (q - 16)] = make_float2(v[q].x, -v[q].y);
        team_sync(team);
        const float bias = sign * 0.5f * ts.x0_xm[1] - 0.5f * ts.x0_xm[0];
        const float* yf = reinterpret_cast<const float*>(ts.W) + t;
        float c[128];
        {
          float xc[64];
          park_ld<kTmem, 64>(tcol, scratch, kColC, xc);
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            c[2 * j] = fmaf((float)kM, xc[j], bias);
            c[2 * j + 1] = yf[kTeam * j];
          }
        }
        park_st<kTmem, 128>(tcol, scratch, kColC, c);
      } else if (tr == 2) {
        float s[64];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          s[2 * j] = v[j].x;
          s[2 * j + 1] = v[j].y;
        }
        park_st<kTmem, 64>(tcol, scratch, kColS, s);
        const float sn[2] = {v[32].x, v[32].y};
        park_st<kTmem, 2>(tcol, scratch, kColSn, sn);
      } else if (tr == 3) {
        float s[64], sn[4], nd[32], ndn[1];
        park_ld<kTmem, 64>(tcol, scratch, kColS, s);
        park_ld<kTmem, 4>(tcol, scratch, kColSn, sn);
#pragma unroll
        for (int j = 0; j < 32; ++j) nd[j] = v[j].y * s[2 * j] - v[j].x * s[2 * j + 1];
        ndn[0] = v[32].y * sn[0] - v[32].x * sn[1];
        park_st<kTmem, 32>(tcol, scratch, kColNd, nd);
        park_st<kTmem, 1>(tcol, scratch, kColSn + 2, ndn);
      }
    }
    // ---- every thread has read its row of the last exchange: the next frame may land
    team_sync(team);
    if (t == 0 && g + stride < total) {
      fence_async_smem();
      mbar_expect_tx(&ts.mbar, kFrameBytes);
      bulk_g2s(ts.W, frame_src(g + stride), kFrameBytes, &ts.mbar);
    }
    // ---- R: reassignment + ordered compaction (order: bin group j, then thread t)
    omb_spectrogram_point pts[kGroups];
    unsigned keep_lo = 0, keep_hi = 0;
    {
      float nd[32], sn[4];
      park_ld<kTmem, 32>(tcol, scratch, kColNd, nd);
      park_ld<kTmem, 4>(tcol, scratch, kColSn, sn);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float s[32];
        if (c < 2) {
          park_ld<kTmem, 32>(tcol, scratch, kColS + 32 * c, s);
        } else {
          s[0] = sn[0];
          s[1] = sn[1];
        }
#pragma unroll
        for (int i = 0; i < (c < 2 ? 16 : 1); ++i) {
          const int j = 16 * c + i;
          const int bin = t + kTeam * j;
          const float norm = (bin == 0 || j == 32) ? ra.norm_dc : ra.norm_ac;
          const float ndj = j < 32 ? nd[j & 31] : sn[2];
          const bool k = reassign_bin_nd(make_float2(s[2 * i], s[2 * i + 1]), ndj, v[j], norm, bin, rc, &pts[j]) & (j < 32 || t == 0);
          const unsigned m = __ballot_sync(0xffffffffu, k);
          if (lane_id == 0) {
            if (wt == 0) ts.cnt[j].x = __popc(m); else ts.cnt[j].y = __popc(m);
          }
          if (k) {
            if (j < 32) keep_lo |= 1u << j; else keep_hi = 1u;
          }
        }
      }
    }
    team_sync(team);
    {
      const uint64_t slot = lane * a.frames_per_lane + f;
      float* out = reinterpret_cast<float*>(a.out_points + slot * a.point_stride);
      int run = 0;
#pragma unroll
      for (int j = 0; j < kGroups; ++j) {
        const bool k = j < 32 ? ((keep_lo >> j) & 1u) != 0 : keep_hi != 0;
        const unsigned m = __ballot_sync(0xffffffffu, k);
        const int2 c = ts.cnt[j];
        if (k) {
          float* o = out + 3 * (run + (wt ? c.x : 0) + __popc(m & lt_mask));
          o[0] = pts[j].time_offset;
          o[1] = pts[j].freq_hz;
          o[2] = pts[j].power;
        }
        run += c.x + c.y;
      }
      if (t == 0) a.out_counts[slot] = (uint32_t)run;
    }
  }
  __syncthreads();
  if (kTmem && tid < 32) tmem_free(tbase);
}

size_t smem_bytes() { return sizeof(Smem); }

bool park_in_tmem() {  // OMB_R64_PARK=global parks in an L2-resident scratch instead (A/B and fallback)
  static const bool v = [] {
    const char* e = getenv("OMB_R64_PARK");
    return !(e && e[0] == 'g');
  }();
  return v;
}

}  // namespace

bool stft_r64_supported(const StftConfig& cfg, const DeviceInfo& dev) {
  if (!cfg.reassign || cfg.window != (uint64_t)kM || cfg.zero_pad != 1) return false;
  if (cfg.hop < 4 || (cfg.hop % 4) != 0) return false;  // bulk copies start at f * hop floats: 16-byte aligned
  if (dev.cc_major != 0 && dev.cc_major < 10) return false;  // tcgen05 / TMEM
  return dev.max_smem_optin == 0 || smem_bytes() <= (size_t)dev.max_smem_optin;
}

int stft_r64_prepare(StftPlan& plan) {
  std::vector<float2> tab(14 * kTeam);
  const double tau = 6.28318530717958647692;
  for (int i = 0; i < 14; ++i)
    for (int t = 0; t < kTeam; ++t) {
      const int e = i < 7 ? t * (i + 1) : 8 * t * (i - 6);
      const double ang = -tau * (double)(e % kM) / (double)kM;
      tab[i * kTeam + t] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  OMB_TRY(plan.d_r64_tables.upload(tab, plan.stream));
  if (!park_in_tmem())
    OMB_TRY(plan.d_r64_scratch.reserve((size_t)std::max(plan.dev.sm_count, 1) * kColsPerWarp * kThreads));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_r64<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_r64<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes()));
  return OMB_OK;
}

int launch_stft_r64(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s) {
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  if (per_lane == 0 || a.n_lanes == 0) return OMB_OK;
  if ((reinterpret_cast<uintptr_t>(a.lanes) & 15u) != 0 || (a.lane_stride % 4) != 0)
    return fail(OMB_ERR_INVALID, "specialised STFT kernel needs 16-byte aligned lanes (pointer and lane_stride % 4 == 0)");
  R64Args ra{};
  ra.a = a;
  ra.tw = plan.d_r64_tables.ptr;
  ra.scratch = plan.d_r64_scratch.ptr;
  ra.norm_ac = plan.h_norm.size() > 1 ? plan.h_norm[1] : plan.h_norm[0];
  ra.norm_dc = plan.h_norm[0];
  const uint64_t total = per_lane * a.n_lanes;
  const uint64_t ctas = (uint64_t)std::max(plan.dev.sm_count, 1);
  const unsigned grid = (unsigned)std::min<uint64_t>((total + kTeams - 1) / kTeams, ctas);
  if (park_in_tmem()) {
    OMB_LAUNCH(k_reassigned_r64<true>, dim3(grid), dim3(kThreads), smem_bytes(), s, ra);
  } else {
    OMB_LAUNCH(k_reassigned_r64<false>, dim3(grid), dim3(kThreads), smem_bytes(), s, ra);
  }
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
