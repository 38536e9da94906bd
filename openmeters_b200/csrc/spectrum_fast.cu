// spectrum_fast.cu — specialised power-spectrum kernel for the spectrum analyzer's default size N = 16384 (BASELINE cfg4).
//
// The real 16384-point transform is one complex 8192-point FFT of (r[2n], r[2n+1]); that FFT is split radix-2
// (decimation in time) into TWO independent 4096-point FFTs which the two 256-thread groups of a CTA run side by
// side with the shared radix-16 engine (fft4096.cuh).  The radix-2 combine, the real-FFT split and |X|^2 * norm are
// fused into one epilogue that turns each quadruple (E[a], O[a], E[4096-a], O[4096-a]) into the four bins
// a, 4096-a, 4096+a, 8192-a:
//     Z[a] = E + w8 O, Z[a+4096] = E - w8 O                       (w8 = W_8192^a)
//     X[k] = A + W_16384^k B,  X[8192-k] = conj(A - W_16384^k B),  A = (P + conj Q)/2, B = (P - conj Q)/(2j)
// DC removal uses per-hop f64 block sums from a small pre-kernel (mean = sum of N/hop block sums / N), so the PCM is
// read once by the FFT kernel.  Rows a4, a12 of SURVEY.md §8; spectrum/processor.rs:215-244.
// Packed FP32x2 switches of this translation unit (common.h; measured in profiles/r02b_packed_ab.md): everything packed
// (adds, rotations, products) — a small gain here (+0.5 ... +1.6 %).
#ifndef OMB_F32X2_MUL
#define OMB_F32X2_MUL 1
#endif
#ifndef OMB_F32X2_ROT
#define OMB_F32X2_ROT 1
#endif
#include <algorithm>
#include <cmath>

#include "fft4096.cuh"
#include "spectrum.h"

namespace omb {

namespace {

using namespace f4k;

constexpr int kN = 16384;
constexpr int kM = 4096;                  // sub-transform length
constexpr int kThreads = 2 * kT;
constexpr int kWSize = f16::phys_size(kM);

struct SpecFastArgs {
  SpectrumPowerArgs a;
  const float2* tw1;     // [15][256] W_4096^{b q} (rows 0,1,3,7 used)
  const float2* tw2;     // [15][16]  W_256^{o q}
  const float2* w16;     // [2049]    W_16384^a
  const double* bsum;    // [lane][n_blocks] f64 sums of hop-sized blocks
  uint64_t n_blocks;
  uint32_t blocks_per_frame;  // N / hop
};

struct SmemS {
  float2 tw1[4 * kT];
  float2 tw2[15 * 16];
  float2 W[2][kWSize];
  float mean;
};

// one warp per (lane, block): f64 sum of `hop` samples
__global__ void __launch_bounds__(256) k_block_sums(const float* lanes, uint64_t lane_stride, uint32_t n_lanes, uint64_t n_blocks,
                                                    uint32_t hop, double* out) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane_id = threadIdx.x & 31;
  const uint64_t total = n_blocks * n_lanes;
  double acc = 0.0;
  if (warp < total) {
    const uint64_t l = warp / n_blocks, b = warp % n_blocks;
    const float* x = lanes + l * lane_stride + b * hop;
    for (uint32_t i = lane_id; i < hop; i += 32) acc += (double)__ldg(&x[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (warp < total && lane_id == 0) out[warp] = acc;
}

__device__ __forceinline__ int posC(int k) { return 273 * (k & 15) + 17 * ((k >> 4) & 15) + (k >> 8); }

__global__ void __launch_bounds__(kThreads, 1) k_spectrum_power_16k(SpecFastArgs fa) {
  OMB_DYN_SMEM(unsigned char, smem_raw);
  SmemS& sm = *reinterpret_cast<SmemS*>(smem_raw);
  const SpectrumPowerArgs& a = fa.a;
  const int tid = threadIdx.x, t = tid & (kT - 1);
  const int g = __shfl_sync(0xffffffffu, tid >> 8, 0);
  for (int i = tid; i < 4 * kT; i += kThreads) {
    const int row = (1 << (i >> 8)) - 1;
    sm.tw1[i] = __ldg(&fa.tw1[row * kT + (i & (kT - 1))]);
  }
  for (int i = tid; i < 15 * 16; i += kThreads) sm.tw2[i] = __ldg(&fa.tw2[i]);
  Addr ad;
  ad.pA = t + (t >> 4);
  ad.pB = 273 * (t >> 4) + (t & 15);
  ad.pC = 273 * (t & 15) + 17 * (t >> 4);
  const float2* tw1t = sm.tw1 + t;
  const float2* tw2o = sm.tw2 + (t & 15);
  __syncthreads();

  const uint64_t total = a.hops * a.n_lanes;
  for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const uint64_t lane = item / a.hops, h = item % a.hops;
    const float* x = a.lanes + lane * a.lane_stride + h * a.hop;
    if (tid == 0) {
      const double* bs = fa.bsum + lane * fa.n_blocks + h;
      double s = 0.0;
      for (uint32_t i = 0; i < fa.blocks_per_frame; ++i) s += bs[i];
      sm.mean = (float)(s / (double)kN);
    }
    __syncthreads();
    const float mean = sm.mean;
    // group g transforms z[2m + g], z[n] = (r[2n], r[2n+1]); thread t owns m = t + 256 j -> samples 4m + 2g, +1
    float2 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int i0 = 4 * (t + kT * j) + 2 * g;
      const float2 xv = __ldg(reinterpret_cast<const float2*>(x + i0));
      const float2 wv = __ldg(reinterpret_cast<const float2*>(a.win + i0));
      v[j] = make_float2((xv.x - mean) * wv.x, (xv.y - mean) * wv.y);
    }
    fft_forward<f16::kAll, 0>(v, sm.W[g], tw1t, tw2o, ad, g);
    float2* wc = sm.W[g] + ad.pC;
#pragma unroll
    for (int q = 0; q < 16; ++q) wc[q] = v[q];
    __syncthreads();
    // epilogue over a = 0..2048
    float* out = a.power + item * a.bins;
    for (int aa = tid; aa <= kM / 2; aa += kThreads) {
      const int bb = kM - aa;
      const float2 Ea = sm.W[0][posC(aa)], Oa = sm.W[1][posC(aa)];
      const float2 Eb = sm.W[0][posC(bb & (kM - 1))], Ob = sm.W[1][posC(bb & (kM - 1))];
      const float2 w = __ldg(&fa.w16[aa]);        // W_16384^a
      const float2 w8 = cmul(w, w);               // W_8192^a
      const float2 ta = cmul(w8, Oa);
      const float2 tb = cmul(make_float2(-w8.x, w8.y), Ob);  // W_8192^{4096-a} = -conj(w8)
      const float2 Za = cadd(Ea, ta), Za2 = csub(Ea, ta);
      const float2 Zb = cadd(Eb, tb), Zb2 = csub(Eb, tb);
      // pair 1: P = Z[a], Q = Z[8192-a] = Zb2, twiddle W_16384^a -> bins a and 8192-a
      {
        const float2 A = make_float2(0.5f * (Za.x + Zb2.x), 0.5f * (Za.y - Zb2.y));
        const float2 d = make_float2(Za.x - Zb2.x, Za.y + Zb2.y);          // P - conj Q
        const float2 B = make_float2(0.5f * d.y, -0.5f * d.x);             // d / (2j)
        const float2 T = cmul(w, B);
        const float2 X0 = cadd(A, T), X1 = csub(A, T);
        out[aa] = (X0.x * X0.x + X0.y * X0.y) * __ldg(&a.bin_norm[aa]);
        out[2 * kM - aa] = (X1.x * X1.x + X1.y * X1.y) * __ldg(&a.bin_norm[2 * kM - aa]);
      }
      // pair 2: P = Z[b], Q = Z[4096+a] = Za2, twiddle W_16384^{4096-a} = -j conj(w) -> bins 4096-a and 4096+a
      {
        const float2 A = make_float2(0.5f * (Zb.x + Za2.x), 0.5f * (Zb.y - Za2.y));
        const float2 d = make_float2(Zb.x - Za2.x, Zb.y + Za2.y);
        const float2 B = make_float2(0.5f * d.y, -0.5f * d.x);
        const float2 wk = make_float2(-w.y, -w.x);                          // -j * conj(w), w = (c, -s) -> (s... ) see below
        const float2 T = cmul(wk, B);
        const float2 X0 = cadd(A, T), X1 = csub(A, T);
        out[bb] = (X0.x * X0.x + X0.y * X0.y) * __ldg(&a.bin_norm[bb]);
        out[kM + aa] = (X1.x * X1.x + X1.y * X1.y) * __ldg(&a.bin_norm[kM + aa]);
      }
    }
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Fused batch kernel: one CTA walks ALL hops of one lane.  The N samples of a frame live in a shared-memory ring
// of N + hop floats (the next hop arrives by 16-byte async copies while the current one is transformed), the window
// is shared-memory resident, and the smoothing recurrence (spectrum/processor.rs:349-402) runs in the FFT epilogue
// with its state in registers: every thread owns the same <= 18 bins for the whole lane, so the power spectrum
// never leaves the SM and the per-hop outputs [weighted, raw] are written exactly once (69 640 algorithmic bytes
// per lane-hop, no scratch round trip, no second kernel).  The arg-max (state.rs:321-325, last maximum wins) is a
// CTA reduction instead of global atomics.  Parallelism is one CTA per lane, so the plan uses this kernel only when
// there are enough lanes to fill the GPU; the two-kernel path above remains for few lanes and for streaming.
constexpr int kSlots = 4;                      // aa = tid + 512 i, i < 4 -> bins aa, 8192 - aa, 4096 - aa, 4096 + aa

struct SpecFusedArgs {
  SpectrumPowerArgs a;                         // lanes / lane_stride / n_lanes / hops / hop / win
  const float2* tw1;
  const float2* tw2;
  const float* means;                          // [lane][hops] frame means (f64 block sums -> f32), see k_frame_means
  const float* a_db;                           // [8193]
  float* out_weighted;                         // [(lane * hops + h) * 8193 + k]
  float* out_raw;
  int32_t* peak_bin;                           // [lane * hops + h] or null
  int peak_raw, peak_lo, peak_hi;              // peak spec (spectrum.h): trace selector and inclusive candidate bin range
  int mode;
  float alpha, decay, state_floor, floor_db;
  float norm_ac, norm_dc;
  uint32_t ring_len;
  // work items: (lane, hop segment).  The smoothing modes walk whole lanes (segs = 1: their state lives in registers across hops);
  // kPowerOnly splits every lane into `segs` segments of `seg_len` hops — the power spectrum has no state.
  uint32_t segs;
  uint64_t seg_len;
  uint64_t means_stride;                       // hops of the whole batch (the means of lane l start at l * means_stride)
};
constexpr int kPowerOnly = 3;                  // kMode beyond OMB_AVG_*: store the normalised power spectrum (two-kernel path)

struct SmemF {
  float2 W[2][kWSize];
  float2 tw1[4 * kT];
  float2 tw2[15 * 16];
  float win[kN];
  float adb_lo[kM + 4];                        // A-weights of bins 0..4096 (the upper half sits in registers)
  unsigned long long wkey[2][kThreads / 32];   // per-warp arg-max keys, double-buffered by hop parity
  // float ring[ring_len] follows
};
static_assert(sizeof(SmemF) % 16 == 0, "ring must stay 16-byte aligned");

__global__ void k_frame_means(const double* bsum, uint64_t n_blocks, uint64_t hops, uint32_t n_lanes, uint32_t blocks_per_frame,
                              float* means) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hops * n_lanes) return;
  const uint64_t lane = i / hops, h = i % hops;
  const double* bs = bsum + lane * n_blocks + h;
  double s = 0.0;
  for (uint32_t b = 0; b < blocks_per_frame; ++b) s += bs[b];
  means[i] = (float)(s / (double)kN);
}

__device__ __forceinline__ void f_async_copy16(float* dst_smem, const float* src_gmem) {
#ifdef OMB_EMU
  for (int i = 0; i < 4; ++i) dst_smem[i] = src_gmem[i];
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src_gmem));
#endif
}
__device__ __forceinline__ void f_async_commit_wait(bool wait) {
#ifndef OMB_EMU
  if (wait) asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  else asm volatile("cp.async.commit_group;\n" ::);
#endif
}
__device__ __forceinline__ void f_async_copy8(float2* dst_smem, const float* src_gmem) {
#ifdef OMB_EMU
  *dst_smem = make_float2(src_gmem[0], src_gmem[1]);
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(src_gmem));
#endif
}
// Planar ring (default since round 2; OMB_SPECTRUM_PLANAR=0 turns it off): the even and odd float2 of every 16-byte quad live in two planes of
// L / 4 float2 each, so that the group that transforms z[2m + g] reads plane g at unit stride (the interleaved ring is read
// at a 16-byte stride: a 2-way bank conflict on every 64-bit load, 22 % of the kernel's shared-memory wavefronts).
__device__ __forceinline__ void f_ring_fetch_planar(float2* plane_e, float2* plane_o, int L, int p0, const float* x, int count) {
  for (int i = 4 * (int)threadIdx.x; i < count; i += 4 * kThreads) {
    int p = p0 + i;
    p -= (p >= L) ? L : 0;
    f_async_copy8(plane_e + (p >> 2), x + i);
    f_async_copy8(plane_o + (p >> 2), x + i + 2);
  }
}
__device__ __forceinline__ void f_ring_fetch(float* ring, int L, int p0, const float* x, int count) {
  for (int i = 4 * (int)threadIdx.x; i < count; i += 4 * kThreads) {
    int p = p0 + i;
    p -= (p >= L) ? L : 0;
    f_async_copy16(ring + p, x + i);
  }
}

// log2(v) from the special-function unit (MUFU.LG2, <= 3 ulp like CUDA's __logf, of which this is the core) — the epilogue
// is issue-bound and the ~30-instruction accurate logf was its largest single item.  3 ulp of a dB value is <= 6e-6
// relative in linear power (the parity budget is 1e-5); never used for the integer-coded classic columns.  The
// arguments here are 0 or >= state_floor >= FLT_MIN, so __logf's subnormal rescaling (4 more instructions) is dead code.
__device__ __forceinline__ float fast_ln(float v) {
#ifdef OMB_EMU
  return logf(v);
#else
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return __fmul_rn(r, 0.69314718246459960938f);
#endif
}

struct FusedConsts {  // kernel arguments the epilogue reads, copied to registers once
  float alpha, one_minus_alpha, decay, state_floor, floor_db;
  bool peak_raw;
};

// One smoothed bin: state update, dB, stores; folds the bin into the running arg-max key (ordered dB bits << 32 | bin:
// larger dB wins, a larger bin wins ties, state.rs:321-325).  Branch-free.  `cand`: the bin is a peak candidate.
template <int kMode>
__device__ __forceinline__ void fused_bin(const FusedConsts& c, float p, float& st, unsigned bin, float aw, float* ow, float* orw, bool cand,
                                          unsigned long long& best) {
  if (kMode == kPowerOnly) {  // the smoothing kernel (spectrum.cu) takes it from here
    *ow = p;
    return;
  }
  float v;
  if (kMode == OMB_AVG_EXPONENTIAL) {  // spectrum/processor.rs:366-377
    st = st <= 0.0f ? p : __fadd_rn(__fmul_rn(st, c.alpha), __fmul_rn(p, c.one_minus_alpha));
    st = st < c.state_floor ? 0.0f : st;
    v = st;
  } else if (kMode == OMB_AVG_PEAK_HOLD) {  // :380-388
    st = fmaxf(__fmul_rn(st, c.decay), p);
    st = st < c.state_floor ? 0.0f : st;
    v = st;
  } else {
    v = p < c.state_floor ? 0.0f : p;
  }
  // :392-401 — values below the state floor are 0 here: lg2(0) = -inf and both maxima return the floor; NaN powers take
  // the dB path as in the reference (f32::max ignores NaN, so does fmaxf)
  const float db = __fmul_rn(fast_ln(v), kLnToDb);
  const float raw = fmaxf(db, c.floor_db);
  const float weighted = fmaxf(__fadd_rn(db, aw), c.floor_db);
  *ow = weighted;
  *orw = raw;
  const float pv = c.peak_raw ? raw : weighted;
  const unsigned u = __float_as_uint(pv);
  const unsigned ob = u ^ ((unsigned)((int)u >> 31) | 0x80000000u);  // monotone float -> uint map
  const bool ok = cand & ((u & 0x7fffffffu) < 0x7f800000u);          // candidate bin with a finite value
  const unsigned long long key = ok ? (((unsigned long long)ob << 32) | bin) : 0ull;
  best = key > best ? key : best;
}

template <int kMode, bool kPlanar = false>
__global__ void __launch_bounds__(kThreads, 1) k_spectrum_fused_16k(SpecFusedArgs fa) {
  OMB_DYN_SMEM(unsigned char, smem_raw);
  SmemF& sm = *reinterpret_cast<SmemF*>(smem_raw);
  float* ring = reinterpret_cast<float*>(smem_raw + sizeof(SmemF));
  const SpectrumPowerArgs& a = fa.a;
  const int tid = threadIdx.x, t = tid & (kT - 1), lane_id = tid & 31;
  const int g = __shfl_sync(0xffffffffu, tid >> 8, 0);
  const int hop = (int)a.hop, L = (int)fa.ring_len;
  for (int i = tid; i < 4 * kT; i += kThreads) {
    const int row = (1 << (i >> 8)) - 1;
    sm.tw1[i] = __ldg(&fa.tw1[row * kT + (i & (kT - 1))]);
  }
  for (int i = tid; i < 15 * 16; i += kThreads) sm.tw2[i] = __ldg(&fa.tw2[i]);
  if (kPlanar) {  // window in the same two planes as the ring: plane g holds (h[4m + 2g], h[4m + 2g + 1])
    float2* wp = reinterpret_cast<float2*>(sm.win);
    for (int i = tid; i < kN / 4; i += kThreads) {
      wp[i] = make_float2(__ldg(&a.win[4 * i]), __ldg(&a.win[4 * i + 1]));
      wp[kN / 4 + i] = make_float2(__ldg(&a.win[4 * i + 2]), __ldg(&a.win[4 * i + 3]));
    }
  } else {
    for (int i = tid; i < kN; i += kThreads) sm.win[i] = __ldg(&a.win[i]);
  }
  for (int i = tid; i <= kM; i += kThreads) sm.adb_lo[i] = __ldg(&fa.a_db[i]);
  float aw_hi[kSlots][2];  // A-weights of this thread's bins 8192 - aa and 4096 + aa
#pragma unroll
  for (int i = 0; i < kSlots; ++i) {
    aw_hi[i][0] = __ldg(&fa.a_db[2 * kM - (tid + kThreads * i)]);
    aw_hi[i][1] = __ldg(&fa.a_db[kM + tid + kThreads * i]);
  }
  const float aw_mid = __ldg(&fa.a_db[kM + kM / 2]);  // bin 6144 (thread 0's extra pair)
  Addr ad;
  ad.pA = t + (t >> 4);
  ad.pB = 273 * (t >> 4) + (t & 15);
  ad.pC = 273 * (t & 15) + 17 * (t >> 4);
  const float2* tw1t = sm.tw1 + t;
  const float2* tw2o = sm.tw2 + (t & 15);
  float cw, sw;  // W_16384^tid = cw - j sw
  sincospif((float)tid / (float)(kN / 2), &sw, &cw);
  float2 wsl[kSlots];  // W_16384^aa for this thread's four slots (aa = tid + 512 i)
#pragma unroll
  for (int i = 0; i < kSlots; ++i) wsl[i] = make_float2(cw * f16::kCos32[i] - sw * f16::kSin32[i], -(sw * f16::kCos32[i] + cw * f16::kSin32[i]));
  const FusedConsts fc{fa.alpha, __fsub_rn(1.0f, fa.alpha), fa.decay, fa.state_floor, fa.floor_db, fa.peak_raw != 0};
  // shared-memory positions of this thread's pair inputs: posC(tid + 512 i) = pa + 2 i; posC(4096 - tid - 512 i) = pb - 2 i
  // (tid = 0: 4096 - 512 i wraps to 0 for i = 0 and sits at 16 - 2 i otherwise)
  const int pa = posC(tid);
  const int pb0 = posC((kM - tid) & (kM - 1));
  const int pb = tid == 0 ? 16 : pb0;
  // peak candidates among this thread's bins, bit 4 i + b (b: aa, 8192 - aa, 4096 - aa, 4096 + aa); bits 16, 17: bins 2048, 6144
  unsigned cand = 0;
#pragma unroll
  for (int i = 0; i < kSlots; ++i) {
    const int aa = tid + kThreads * i;
    const int bins4[4] = {aa, 2 * kM - aa, kM - aa, kM + aa};
#pragma unroll
    for (int b = 0; b < 4; ++b) cand |= (bins4[b] >= fa.peak_lo && bins4[b] <= fa.peak_hi) ? (1u << (4 * i + b)) : 0u;
  }
  cand |= (kM / 2 >= fa.peak_lo && kM / 2 <= fa.peak_hi) ? (1u << 16) : 0u;
  cand |= (kM + kM / 2 >= fa.peak_lo && kM + kM / 2 <= fa.peak_hi) ? (1u << 17) : 0u;
  if (!fa.peak_bin) cand = 0;
  // |X|^2 = |2X|^2 / 4: the split below works on doubled spectra (exact: powers of two)
  const float norm_ac = 0.25f * fa.norm_ac, norm_dc = 0.25f * fa.norm_dc;
  __syncthreads();

  // (the smoothing modes see compile-time segs = 1, h_begin = 0: their code is the whole-lane loop of round 1)
  const uint64_t n_items = kMode == kPowerOnly ? (uint64_t)a.n_lanes * fa.segs : (uint64_t)a.n_lanes;
  for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
    const uint32_t lane = kMode == kPowerOnly ? (uint32_t)(item / fa.segs) : (uint32_t)item;
    const uint64_t h_begin = kMode == kPowerOnly ? (item % fa.segs) * fa.seg_len : 0;
    const uint64_t h_end = kMode == kPowerOnly ? (h_begin + fa.seg_len < a.hops ? h_begin + fa.seg_len : a.hops) : a.hops;
    if (kMode == kPowerOnly && h_begin >= h_end) continue;
    const float* x = a.lanes + (uint64_t)lane * a.lane_stride + h_begin * (uint64_t)hop;
    const float* means = fa.means + (uint64_t)lane * (kMode == kPowerOnly ? fa.means_stride : a.hops);
    float st[kSlots][4];
    float st_mid[2] = {0.0f, 0.0f};  // bins 2048 and 6144 (aa = 2048), thread 0 only
#pragma unroll
    for (int i = 0; i < kSlots; ++i)
#pragma unroll
      for (int b = 0; b < 4; ++b) st[i][b] = 0.0f;
    int r0 = 0;
    float2* const plane_e = reinterpret_cast<float2*>(ring);  // planar ring: L / 4 float2 per plane
    float2* const plane_o = plane_e + L / 4;
    if (kPlanar) f_ring_fetch_planar(plane_e, plane_o, L, 0, x, kN);
    else f_ring_fetch(ring, L, 0, x, kN);
    f_async_commit_wait(false);
    float mean_next = __ldg(&means[h_begin]);
    for (uint64_t h = h_begin; h < h_end; ++h) {
      f_async_commit_wait(true);
      __syncthreads();  // ring holds frame h; everybody is done with frame h - 1 (ring, W, wkey[(h - 1) & 1] complete)
      if (h + 1 < h_end) {
        int p0 = r0 + kN;
        p0 -= (p0 >= L) ? L : 0;
        if (kPlanar) f_ring_fetch_planar(plane_e, plane_o, L, p0, x + (h - h_begin) * (uint64_t)hop + kN, hop);
        else f_ring_fetch(ring, L, p0, x + (h - h_begin) * (uint64_t)hop + kN, hop);
      }
      f_async_commit_wait(false);
      if (fa.peak_bin && h > h_begin && tid == 0) {  // finish the previous hop's arg-max
        unsigned long long best = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) best = sm.wkey[(h - 1) & 1][w] > best ? sm.wkey[(h - 1) & 1][w] : best;
        fa.peak_bin[(uint64_t)lane * a.hops + h - 1] = best ? (int32_t)(best & 0xffffffffu) : -1;
      }
      const float mean = mean_next;
      if (h + 1 < h_end) mean_next = __ldg(&means[h + 1]);
      const float* lo = ring + r0;
      const float* hi = lo - L;
      const int split = L - r0;  // multiple of 4 (hop % 4 == 0): a float2 at an even offset never straddles the wrap
      // group g transforms z[2m + g], z[n] = (r[2n], r[2n+1]); thread t owns m = t + 256 j -> samples 4m + 2g, +1
      float2 v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float2 xv, wv;
        if (kPlanar) {  // plane g, quad (r0 / 4 + m) mod (L / 4), m = t + 256 j: unit stride across the warp
          const int m = t + kT * j;
          int q = (r0 >> 2) + m;
          q -= (q >= (L >> 2)) ? (L >> 2) : 0;
          xv = (g ? plane_o : plane_e)[q];
          wv = reinterpret_cast<const float2*>(sm.win)[g * (kN / 4) + m];
        } else {
          const int i0 = 4 * (t + kT * j) + 2 * g;
          xv = *reinterpret_cast<const float2*>((i0 < split ? lo : hi) + i0);
          wv = *reinterpret_cast<const float2*>(sm.win + i0);
        }
        v[j] = make_float2((xv.x - mean) * wv.x, (xv.y - mean) * wv.y);
      }
      fft_forward<f16::kAll, 0>(v, sm.W[g], tw1t, tw2o, ad, g);
      float2* wc = sm.W[g] + ad.pC;
#pragma unroll
      for (int q = 0; q < 16; ++q) wc[q] = v[q];
      __syncthreads();
      // epilogue: combine + real split + |X|^2 norm + smoothing + dB, four bins per (thread, slot)
      float* owp = fa.out_weighted + ((uint64_t)lane * a.hops + h) * (uint64_t)(kN / 2 + 1) + tid;  // bins tid + c
      float* orp = fa.out_raw + ((uint64_t)lane * a.hops + h) * (uint64_t)(kN / 2 + 1) + tid;
      float* owm = owp - 2 * tid;                                                                     // bins c - tid
      float* orm = orp - 2 * tid;
      unsigned long long best = 0;
#pragma unroll
      for (int i = 0; i <= kSlots; ++i) {
        if (i == kSlots && tid != 0) break;
        // i < 4: aa = tid + 512 i, bb = 4096 - aa;  i = 4 (thread 0 only): aa = bb = 2048
        const int ia = (i < kSlots) ? pa + 2 * i : posC(kM / 2);
        const int ib = (i < kSlots) ? (i == 0 ? pb0 : pb - 2 * i) : posC(kM / 2);
        const float2 Ea = sm.W[0][ia], Oa = sm.W[1][ia];
        const float2 Eb = sm.W[0][ib], Ob = sm.W[1][ib];
        const float2 w = (i < kSlots) ? wsl[i < kSlots ? i : 0] : make_float2(f16::kH, -f16::kH);  // W_16384^aa (aa = 2048: W_8)
        const float2 w8 = cmul(w, w);
        const float2 ta = cmul(w8, Oa);
        const float2 tb = cmul(make_float2(-w8.x, w8.y), Ob);
        const float2 Za = f16::cadd2(Ea, ta), Za2 = f16::csub2(Ea, ta);
        const float2 Zb = f16::cadd2(Eb, tb), Zb2 = f16::csub2(Eb, tb);
        float p0, p1, p2, p3;
        {
          const float2 A = make_float2(Za.x + Zb2.x, Za.y - Zb2.y);   // 2 A
          const float2 B = make_float2(Za.y + Zb2.y, Zb2.x - Za.x);   // 2 B = (P - conj Q) / j
          const float2 T = cmul(w, B);
          const float2 X0 = f16::cadd2(A, T), X1 = f16::csub2(A, T);
          const float nrm = (i == 0 && tid == 0) ? norm_dc : norm_ac;
          p0 = (X0.x * X0.x + X0.y * X0.y) * nrm;   // bin aa
          p1 = (X1.x * X1.x + X1.y * X1.y) * nrm;   // bin 8192 - aa
        }
        {
          const float2 A = make_float2(Zb.x + Za2.x, Zb.y - Za2.y);
          const float2 B = make_float2(Zb.y + Za2.y, Za2.x - Zb.x);
          const float2 wk = make_float2(-w.y, -w.x);
          const float2 T = cmul(wk, B);
          const float2 X0 = f16::cadd2(A, T), X1 = f16::csub2(A, T);
          p2 = (X0.x * X0.x + X0.y * X0.y) * norm_ac;   // bin 4096 - aa
          p3 = (X1.x * X1.x + X1.y * X1.y) * norm_ac;   // bin 4096 + aa
        }
        if (i < kSlots) {
          const int c = kThreads * i;
          const unsigned aa = (unsigned)(tid + c);
          fused_bin<kMode>(fc, p0, st[i][0], aa, sm.adb_lo[aa], owp + c, orp + c, (cand >> (4 * i)) & 1u, best);
          fused_bin<kMode>(fc, p1, st[i][1], 2 * kM - aa, aw_hi[i][0], owm + (2 * kM - c), orm + (2 * kM - c), (cand >> (4 * i + 1)) & 1u, best);
          fused_bin<kMode>(fc, p2, st[i][2], kM - aa, sm.adb_lo[kM - aa], owm + (kM - c), orm + (kM - c), (cand >> (4 * i + 2)) & 1u, best);
          // aa = 0: bin 4096 a second time (the same bin from the mirrored pair, as in k_spectrum_power_16k where the
          // later store wins too) — cheaper than a divergent branch for one thread
          fused_bin<kMode>(fc, p3, st[i][3], kM + aa, aw_hi[i][1], owp + (kM + c), orp + (kM + c), (cand >> (4 * i + 3)) & 1u, best);
        } else {  // aa = 2048 (thread 0): the two pairs coincide (bins 2048 and 6144)
          fused_bin<kMode>(fc, p0, st_mid[0], kM / 2, sm.adb_lo[kM / 2], owp + kM / 2, orp + kM / 2, (cand >> 16) & 1u, best);
          fused_bin<kMode>(fc, p1, st_mid[1], kM + kM / 2, aw_mid, owp + (kM + kM / 2), orp + (kM + kM / 2), (cand >> 17) & 1u, best);
        }
      }
      if (fa.peak_bin) {
        const unsigned hi32 = __reduce_max_sync(0xffffffffu, (unsigned)(best >> 32));
        const unsigned lo32 = __reduce_max_sync(0xffffffffu, ((unsigned)(best >> 32) == hi32) ? (unsigned)best : 0u);
        if (lane_id == 0) sm.wkey[h & 1][tid >> 5] = hi32 ? (((unsigned long long)hi32 << 32) | lo32) : 0ull;
      }
      r0 += hop;
      r0 -= (r0 >= L) ? L : 0;
    }
    __syncthreads();
    if (fa.peak_bin && tid == 0) {
      unsigned long long best = 0;
      for (int w = 0; w < kThreads / 32; ++w) best = sm.wkey[(h_end - 1) & 1][w] > best ? sm.wkey[(h_end - 1) & 1][w] : best;
      fa.peak_bin[(uint64_t)lane * a.hops + h_end - 1] = best ? (int32_t)(best & 0xffffffffu) : -1;
    }
    __syncthreads();
  }
}

size_t fused_smem_bytes(uint64_t hop) { return sizeof(SmemF) + (size_t)(kN + hop) * sizeof(float); }

}  // namespace

bool spectrum_fast_supported(const SpectrumConfigN& cfg, const DeviceInfo& dev) {
  if (cfg.fft_size != (uint64_t)kN) return false;
  if (cfg.hop == 0 || (kN % cfg.hop) != 0 || (cfg.hop & 1)) return false;
  return dev.max_smem_optin == 0 || sizeof(SmemS) + 256 <= (size_t)dev.max_smem_optin;
}

int spectrum_fast_prepare(SpectrumPlan& p) {
  std::vector<float2> tab(15 * kT + 15 * 16 + (kM / 2 + 1));
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
  for (int a = 0; a <= kM / 2; ++a) {
    const double ang = -tau * (double)a / (double)kN;
    tab[15 * kT + 15 * 16 + a] = make_float2((float)std::cos(ang), (float)std::sin(ang));
  }
  OMB_TRY(p.d_fast_tables.upload(tab, p.stream));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_spectrum_power_16k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemS)));
  p.fused16k = false;
  if ((p.cfg.hop % 4) == 0 && (p.dev.max_smem_optin == 0 || fused_smem_bytes(p.cfg.hop) <= (size_t)p.dev.max_smem_optin)) {
    const int fs = (int)fused_smem_bytes(p.cfg.hop);
    OMB_CUDA_TRY(cudaFuncSetAttribute(k_spectrum_fused_16k<OMB_AVG_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs));
    OMB_CUDA_TRY(cudaFuncSetAttribute(k_spectrum_fused_16k<OMB_AVG_EXPONENTIAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs));
    OMB_CUDA_TRY(cudaFuncSetAttribute(k_spectrum_fused_16k<OMB_AVG_PEAK_HOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize, fs));
    p.fused16k = true;
  }
  return OMB_OK;
}

// Whole-batch fused path (FFT + smoothing + dB + arg-max in one kernel, one CTA per lane). Zero initial state.
int launch_spectrum_fused(SpectrumPlan& p, const float* d_lanes, uint32_t n_lanes, uint64_t hops, uint64_t lane_stride, float* d_weighted,
                          float* d_raw, int32_t* d_peak_bin, cudaStream_t s) {
  if (!hops || !n_lanes) return OMB_OK;
  const SpectrumConfigN& cfg = p.cfg;
  OMB_TRY(spectrum_fast_block_sums(p, d_lanes, lane_stride, n_lanes, hops, s));
  const uint64_t n_blocks = hops - 1 + (uint64_t)kN / cfg.hop;
  OMB_TRY(p.d_means.reserve((size_t)(hops * n_lanes)));
  OMB_LAUNCH(k_frame_means, dim3((unsigned)((hops * n_lanes + 255) / 256)), dim3(256), 0, s, p.d_bsum.ptr, n_blocks, hops, n_lanes,
             (uint32_t)(kN / cfg.hop), p.d_means.ptr);
  OMB_CHECK_LAUNCH();
  SpecFusedArgs fa{};
  fa.a.lanes = d_lanes;
  fa.a.lane_stride = lane_stride;
  fa.a.n_lanes = n_lanes;
  fa.a.hops = hops;
  fa.a.hop = (uint32_t)cfg.hop;
  fa.a.win = p.d_win.ptr;
  fa.tw1 = p.d_fast_tables.ptr;
  fa.tw2 = fa.tw1 + 15 * kT;
  fa.means = p.d_means.ptr;
  fa.a_db = p.d_adb.ptr;
  fa.out_weighted = d_weighted;
  fa.out_raw = d_raw;
  fa.peak_bin = d_peak_bin;
  fa.peak_raw = p.peak_spec.trace == 1u ? 1 : 0;
  fa.peak_lo = p.peak_lo;
  fa.peak_hi = p.peak_hi;
  fa.mode = (int)cfg.averaging;
  fa.alpha = std::min(std::max(cfg.averaging_param, 0.0f), 0.9999f);
  fa.decay = db_to_power_host(-std::fmax(cfg.averaging_param, 0.0f) * ((float)cfg.hop / cfg.sample_rate));
  fa.state_floor = p.state_floor;
  fa.floor_db = cfg.floor_db;
  fa.norm_ac = p.h_norm.size() > 1 ? p.h_norm[1] : p.h_norm[0];
  fa.norm_dc = p.h_norm[0];
  fa.ring_len = (uint32_t)(kN + cfg.hop);
  fa.segs = 1;
  fa.seg_len = hops;
  fa.means_stride = hops;
  const unsigned grid = (unsigned)std::min<uint64_t>(n_lanes, (uint64_t)std::max(p.dev.sm_count, 1));
  const size_t fs = fused_smem_bytes(cfg.hop);
  // Planar ring (conflict-free frame loads): measured on B200 in round 2, 2.307e7 -> 2.481e7 lane-hops/s (+7.5 %, cfg4, 128 lanes;
  // profiles/r02a_cfg4_planar.json), the spectrum GPU tests pass with it: default.  OMB_SPECTRUM_PLANAR=0 restores the interleaved ring.
  const char* planar_env = getenv("OMB_SPECTRUM_PLANAR");
  const bool planar = !(planar_env && planar_env[0] == '0') && (cfg.hop % 4) == 0;
  if (planar) {
    auto kp = k_spectrum_fused_16k<OMB_AVG_PEAK_HOLD, true>;
    auto ke = k_spectrum_fused_16k<OMB_AVG_EXPONENTIAL, true>;
    auto kn = k_spectrum_fused_16k<OMB_AVG_NONE, true>;
    auto k = fa.mode == OMB_AVG_PEAK_HOLD ? kp : (fa.mode == OMB_AVG_EXPONENTIAL ? ke : kn);
    OMB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs));
    OMB_LAUNCH(k, dim3(grid), dim3(kThreads), fs, s, fa);
  } else if (fa.mode == OMB_AVG_PEAK_HOLD) {
    OMB_LAUNCH(k_spectrum_fused_16k<OMB_AVG_PEAK_HOLD>, dim3(grid), dim3(kThreads), fs, s, fa);
  } else if (fa.mode == OMB_AVG_EXPONENTIAL) {
    OMB_LAUNCH(k_spectrum_fused_16k<OMB_AVG_EXPONENTIAL>, dim3(grid), dim3(kThreads), fs, s, fa);
  } else {
    OMB_LAUNCH(k_spectrum_fused_16k<OMB_AVG_NONE>, dim3(grid), dim3(kThreads), fs, s, fa);
  }
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

// The front half of the fused kernel as the power stage of the two-kernel path (few lanes, streaming): hop-overlapped staging ring,
// one pass over the PCM, (lane, hop segment) work items that fill the GPU whatever the lane count.  `d_lanes` points at hop h0 of the
// batch, `hops` is the chunk length; the frame means of the WHOLE batch were computed once (spectrum_fast_frame_means) and the
// chunk starts at hop `h0` of them.
int launch_spectrum_fused_power(SpectrumPlan& p, const float* d_lanes, uint32_t n_lanes, uint64_t hops, uint64_t lane_stride, float* d_power,
                                uint64_t h0, uint64_t hops_total, cudaStream_t s) {
  if (!hops || !n_lanes) return OMB_OK;
  const SpectrumConfigN& cfg = p.cfg;
  SpecFusedArgs fa{};
  fa.a.lanes = d_lanes;
  fa.a.lane_stride = lane_stride;
  fa.a.n_lanes = n_lanes;
  fa.a.hops = hops;
  fa.a.hop = (uint32_t)cfg.hop;
  fa.a.win = p.d_win.ptr;
  fa.tw1 = p.d_fast_tables.ptr;
  fa.tw2 = fa.tw1 + 15 * kT;
  fa.means = p.d_means.ptr + h0;
  fa.means_stride = hops_total;
  fa.a_db = p.d_adb.ptr;
  fa.out_weighted = d_power;
  fa.out_raw = d_power;
  fa.peak_bin = nullptr;
  fa.mode = kPowerOnly;
  fa.state_floor = p.state_floor;
  fa.floor_db = cfg.floor_db;
  fa.norm_ac = p.h_norm.size() > 1 ? p.h_norm[1] : p.h_norm[0];
  fa.norm_dc = p.h_norm[0];
  fa.ring_len = (uint32_t)(kN + cfg.hop);
  // segments: about two work items per SM, at least 8 hops each (a segment primes its ring with a whole frame)
  const uint64_t sms = (uint64_t)std::max(p.dev.sm_count, 1);
  uint64_t segs = std::max<uint64_t>(1, (2 * sms + n_lanes - 1) / n_lanes);
  uint64_t seg_len = std::max<uint64_t>(std::min<uint64_t>(8, hops), (hops + segs - 1) / segs);
  segs = (hops + seg_len - 1) / seg_len;
  fa.segs = (uint32_t)segs;
  fa.seg_len = seg_len;
  const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)n_lanes * segs, sms);
  const size_t fs = fused_smem_bytes(cfg.hop);
  const bool planar = (cfg.hop % 4) == 0;
  if (planar) {
    auto k = k_spectrum_fused_16k<kPowerOnly, true>;
    OMB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs));
    OMB_LAUNCH(k, dim3(grid), dim3(kThreads), fs, s, fa);
  } else {
    auto k = k_spectrum_fused_16k<kPowerOnly, false>;
    OMB_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fs));
    OMB_LAUNCH(k, dim3(grid), dim3(kThreads), fs, s, fa);
  }
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

// Frame means of a whole batch (f64 block sums -> f32), once; the fused kernels index them by absolute hop.
int spectrum_fast_frame_means(SpectrumPlan& p, const float* d_lanes, uint64_t lane_stride, uint32_t n_lanes, uint64_t hops, cudaStream_t s) {
  OMB_TRY(spectrum_fast_block_sums(p, d_lanes, lane_stride, n_lanes, hops, s));
  const uint64_t n_blocks = hops - 1 + (uint64_t)kN / p.cfg.hop;
  OMB_TRY(p.d_means.reserve((size_t)(hops * n_lanes)));
  OMB_LAUNCH(k_frame_means, dim3((unsigned)((hops * n_lanes + 255) / 256)), dim3(256), 0, s, p.d_bsum.ptr, n_blocks, hops, n_lanes,
             (uint32_t)(kN / p.cfg.hop), p.d_means.ptr);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

int spectrum_fast_block_sums(SpectrumPlan& p, const float* d_lanes, uint64_t lane_stride, uint32_t n_lanes, uint64_t hops, cudaStream_t s) {
  const uint64_t n_blocks = hops - 1 + (uint64_t)kN / p.cfg.hop;
  OMB_TRY(p.d_bsum.reserve((size_t)(n_blocks * n_lanes)));
  const uint64_t warps = n_blocks * n_lanes;
  OMB_LAUNCH(k_block_sums, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, s, d_lanes, lane_stride, n_lanes, n_blocks,
             (uint32_t)p.cfg.hop, p.d_bsum.ptr);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

int launch_spectrum_power_fast(SpectrumPlan& p, SpectrumPowerArgs& a, cudaStream_t s) {
  const uint64_t total = a.hops * a.n_lanes;
  if (!total) return OMB_OK;
  if ((reinterpret_cast<uintptr_t>(a.lanes) & 7u) != 0 || (a.lane_stride & 1))
    return fail(OMB_ERR_INVALID, "specialised spectrum kernel needs 8-byte aligned lanes");
  SpecFastArgs fa{};
  fa.a = a;
  fa.tw1 = p.d_fast_tables.ptr;
  fa.tw2 = fa.tw1 + 15 * kT;
  fa.w16 = fa.tw2 + 15 * 16;
  fa.blocks_per_frame = (uint32_t)(kN / a.hop);
  if (p.ext_bsum_blocks) {  // batch-wide sums computed by spectrum_fast_block_sums(); this launch starts at ext_block_off
    fa.n_blocks = p.ext_bsum_blocks;
    fa.bsum = p.d_bsum.ptr + p.ext_block_off;
  } else {
    OMB_TRY(spectrum_fast_block_sums(p, a.lanes, a.lane_stride, a.n_lanes, a.hops, s));
    fa.n_blocks = a.hops - 1 + fa.blocks_per_frame;
    fa.bsum = p.d_bsum.ptr;
  }
  const unsigned grid = (unsigned)std::min<uint64_t>(total, (uint64_t)std::max(p.dev.sm_count, 1));
  OMB_LAUNCH(k_spectrum_power_16k, dim3(grid), dim3(kThreads), sizeof(SmemS), s, fa);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
