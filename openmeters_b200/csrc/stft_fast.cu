// stft_fast.cu — specialised sm_100a kernel for the reassigned STFT (BASELINE configs[1]: N = 4096,
// hop 1024, Blackman-Harris; any window kind and any hop = N/2^k >= 256 take the same kernel).
//
// One 256-thread CTA walks a run of consecutive frames of one lane.  All data of a frame stays on chip:
//
//   ring   : hop-overlapped staging — the H = 2N samples of a frame live in a shared-memory ring of H + hop
//            floats; only the `hop` new samples per frame are fetched (async copy, overlapped with the
//            previous frame's compute), so every PCM sample crosses L2->SM once per run.
//   W      : one padded complex work buffer (M = N points) in which every transform runs in place as three
//            radix-16 register passes (fft16.cuh).
//
// Per frame (restructured — same mathematics, fewer flops than the literal restatement):
//   F  : Z = FFT_M(x[2n] + j x[2n+1])                      (real 2M-point FFT via one M-point complex FFT)
//   X  : pair step Z[k], Z[M-k] -> Q[k]: real-FFT split, Hilbert mask (DC and negative bins zeroed,
//        Nyquist kept) and the complex-to-real packing of the inverse fused into one pointwise step
//   I  : q = IFFT_M(Q): q[m] = Im a[2m] + j Im a[2m+1] of the analytic signal a
//        (Re a[n] = (H x[n] - X[0] + (-1)^n X[H/2]) / 2 needs no transform)
//   G  : c = centre N samples of a; S, D, T = FFT_M(c*h), FFT_M(c*dh), FFT_M(c*th)
//   R  : per-bin reassignment + order-preserving compaction (ascending bin) -> points, count
// 5 complex M-point FFTs per frame instead of the literal 2 x 2M + 3 x M.
// Packed FP32x2 switches of this translation unit (common.h; measured in profiles/r02b_packed_ab.md): packed complex adds
// only — packed products cost this FMA-pipe-bound kernel 2-13 %.
#ifndef OMB_F32X2_CMUL
#define OMB_F32X2_CMUL 0
#endif
#include "device_math.cuh"
#include "fft16.cuh"
#include "stft.h"

#include <cstdlib>

namespace omb {

namespace {

constexpr int kM = 4096;             // complex transform length (= window N = H/2)
constexpr int kT = kM / 16;          // 256 threads
constexpr int kWarps = kT / 32;
constexpr int kWSize = f16::phys_size(kM);  // float2 elements of the padded work buffer
constexpr int kGroups = 9;           // bins t + 256*j for j = 0..7, plus bin 2048 (thread 0, j = 8)

// cos / sin of 2*pi*j/32, j = 0..15 (W_32^j = c - i s)
__device__ constexpr float kW32c[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128673848f, 0.83146961230254523708f,
                                        0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508978178f, 0.19509032201612826785f,
                                        0.0f, -0.19509032201612826785f, -0.38268343236508978178f, -0.55557023301960222474f,
                                        -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128673848f, -0.98078528040323044913f};
__device__ constexpr float kW32s[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508978178f, 0.55557023301960222474f,
                                        0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128673848f, 0.98078528040323044913f,
                                        1.0f, 0.98078528040323044913f, 0.92387953251128673848f, 0.83146961230254523708f,
                                        0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508978178f, 0.19509032201612826785f};

struct FastArgs {
  StftKernelArgs a;
  const float2* tw1;   // [15][256]  W_4096^{b*q}, q = 1..15
  const float2* tw2;   // [15][16]   W_256^{o*q}
  const float2* twh;   // [256]      W_8192^{t}
  uint32_t frames_per_run;
  uint32_t runs_per_lane;
  uint32_t ring_len;   // H + hop floats
};

struct Smem {
  float2 W[kWSize];
  int warp_cnt[kGroups * kWarps];
  int offs[kGroups * kWarps + 1];
  float x0_xm[2];
  // float ring[H + hop] follows (dynamic)
};

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
  asm volatile("cp.async.wait_group 0;\n" ::);
#endif
}

// Loads `count` floats (multiple of 4, 16-byte aligned both sides) of the lane into ring position `pos`
// (no wrap inside the copied span).
__device__ __forceinline__ void ring_fetch(float* ring, int pos, const float* src, int count) {
  for (int i = threadIdx.x * 4; i < count; i += kT * 4) async_copy16(ring + pos + i, src + i);
}

__device__ __forceinline__ int ring_wrap(int i, int ring_len) { return i >= ring_len ? i - ring_len : i; }

// Per-thread shared-memory bases of the three access patterns (fft16.cuh). With phys(i) = i + i/16 + i/256:
//   A: phys(t + 256 q)              = pA + 273 q,   pA = t + t/16
//   B: phys(blk*256 + o + 16 j)     = pB + 17 j,    pB = 273*blk + o        (blk = t/16, o = t%16)
//   C: phys(q1*256 + q2*16 + j)     = pC + j,       pC = 273*q1 + 17*q2     (q1 = t%16, q2 = t/16)
// so every element address is a per-thread base plus a compile-time constant (immediate offsets in LDS/STS).
struct Addr {
  int pA, pB, pC;
};
__device__ __forceinline__ Addr make_addr(int t) {
  Addr a;
  a.pA = t + (t >> 4);
  a.pB = 273 * (t >> 4) + (t & 15);
  a.pC = 273 * (t & 15) + 17 * (t >> 4);
  return a;
}

// External twiddles of one pass: v[q] *= W^{idx*q} (conjugated for the inverse), q = 1..15.
// kTw = 0: 15 table loads.  kTw = 1: 4 loads (q = 1, 2, 4, 8) + 11 products (each derived value is at most two
// multiplications away from a correctly rounded table entry).
template <bool INV, int kTw>
__device__ __forceinline__ void twiddle15(float2 (&v)[16], const float2* __restrict__ tab, int stride, int idx) {
  if (kTw == 0) {
#pragma unroll
    for (int q = 1; q < 16; ++q) v[q] = f16::mul_tw<INV>(v[q], __ldg(&tab[(q - 1) * stride + idx]));
  } else {
    float2 w[16];
    w[1] = __ldg(&tab[0 * stride + idx]);
    w[2] = __ldg(&tab[1 * stride + idx]);
    w[4] = __ldg(&tab[3 * stride + idx]);
    w[8] = __ldg(&tab[7 * stride + idx]);
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

// DIF forward passes 1..3 over W. Input already in v (access A: element b + 256 j in v[j]).
// On return thread t holds frequencies t + 256*j in v[j].
template <int kTw>
__device__ __forceinline__ void fft_forward(float2 (&v)[16], float2* W, const FastArgs& fa, const Addr& ad) {
  const int t = threadIdx.x;
  // pass 1: butterfly over j (stride 256), twiddle W_M^{t*q}, store in place
  f16::dft16<false>(v);
  twiddle15<false, kTw>(v, fa.tw1, kT, t);
  float2* wa = W + ad.pA;
#pragma unroll
  for (int q = 0; q < 16; ++q) wa[273 * q] = v[q];
  __syncthreads();
  // pass 2: thread (blk = t>>4, o = t&15): elements blk*256 + o + 16 j, twiddle W_256^{o*q}
  float2* wb = W + ad.pB;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = wb[17 * j];
  f16::dft16<false>(v);
  twiddle15<false, kTw>(v, fa.tw2, 16, t & 15);
#pragma unroll
  for (int q = 0; q < 16; ++q) wb[17 * q] = v[q];
  __syncthreads();
  // pass 3 (access C): thread t owns the group whose outputs are frequencies t + 256*q
  const float2* wc = W + ad.pC;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = wc[j];
  f16::dft16<false>(v);
}

// DIT inverse passes 1..3. Input: thread t holds Q[t + 256*j] in v[j]. Output (in W, natural order):
// q[m] at phys(m); thread t wrote m = t + 256*j.
template <int kTw>
__device__ __forceinline__ void fft_inverse_to_smem(float2 (&v)[16], float2* W, const FastArgs& fa, const Addr& ad) {
  const int t = threadIdx.x;
  // pass 1 (access C), no twiddles
  f16::dft16<true>(v);
  float2* wc = W + ad.pC;
#pragma unroll
  for (int q = 0; q < 16; ++q) wc[q] = v[q];
  __syncthreads();
  // pass 2 (access B): thread (q1 = t>>4, m0 = t&15): elements q1*256 + m0 + 16*q2, pre-twiddle conj W_256^{m0*q2}
  float2* wb = W + ad.pB;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = wb[17 * j];
  twiddle15<true, kTw>(v, fa.tw2, 16, t & 15);
  f16::dft16<true>(v);
#pragma unroll
  for (int q = 0; q < 16; ++q) wb[17 * q] = v[q];
  __syncthreads();
  // pass 3 (access A): thread b: elements b + 256*q1, pre-twiddle conj W_M^{b*q1}
  float2* wa = W + ad.pA;
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = wa[273 * j];
  twiddle15<true, kTw>(v, fa.tw1, kT, t);
  f16::dft16<true>(v);
#pragma unroll
  for (int q = 0; q < 16; ++q) wa[273 * q] = v[q];
  __syncthreads();
}

template <int kMinBlocks, int kTw>
__global__ void __launch_bounds__(kT, kMinBlocks) k_reassigned_fast(FastArgs fa) {
  OMB_DYN_SMEM(unsigned char, smem_raw);
  Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
  float* ring = reinterpret_cast<float*>(smem_raw + sizeof(Smem));
  const StftKernelArgs& a = fa.a;
  const int t = threadIdx.x, lane_id = t & 31, warp = t >> 5;
  const int hop = (int)a.hop, H = 2 * kM, ring_len = (int)fa.ring_len;
  const int off = (H - kM) / 2;  // centre offset of the analysis window inside the Hilbert frame
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total_runs = (uint64_t)fa.runs_per_lane * a.n_lanes;
  const float2 wh = __ldg(&fa.twh[t]);  // W_8192^t
  const ReassignConsts rc{a.bin_hz, a.max_hz, a.inv_2pi, a.inv_hop, a.latency_hops};
  const float sign = (t & 1) ? -1.0f : 1.0f;  // (-1)^(off + n), off even, n = t + 256 j
  const Addr ad = make_addr(t);

  for (uint64_t run = blockIdx.x; run < total_runs; run += gridDim.x) {
    const uint64_t lane = run / fa.runs_per_lane;
    const uint64_t r_in_lane = run % fa.runs_per_lane;
    const uint64_t f_begin = a.first_frame + r_in_lane * fa.frames_per_run;
    const uint64_t f_end = (f_begin + fa.frames_per_run < a.frames_per_lane) ? f_begin + fa.frames_per_run : a.frames_per_lane;
    const float* x = a.lanes + lane * a.lane_stride;
    (void)per_lane;
    // prime the ring with the first frame of the run: samples [f_begin*hop, f_begin*hop + H) at ring position
    // (f*hop) mod ring_len, copied in two spans around the wrap point
    {
      const int r0 = (int)((f_begin * (uint64_t)hop) % (uint64_t)ring_len);
      const int first = H < ring_len - r0 ? H : ring_len - r0;
      ring_fetch(ring, r0, x + f_begin * hop, first);
      if (first < H) ring_fetch(ring, 0, x + f_begin * hop + first, H - first);
      async_commit();
    }
    for (uint64_t f = f_begin; f < f_end; ++f) {
      const int r0 = (int)((f * (uint64_t)hop) % (uint64_t)ring_len);
      async_wait_all();
      __syncthreads();  // ring complete for frame f; previous frame's readers of W / offs are done
      if (f + 1 < f_end) {  // prefetch the hop new samples of frame f+1 into the slot frame f-1 vacated
        const int dst = ring_wrap(r0 + H, ring_len);
        ring_fetch(ring, dst, x + f * hop + H, hop);
      }
      async_commit();

      float2 v[16];
      // ---- F: z[n] = x[2n] + j x[2n+1], n = t + 256 j  -> sample 2t + 512 j
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int sj = ring_wrap(r0 + 2 * kT * j, ring_len);  // CTA-uniform: r0, ring_len are multiples of 512
        v[j] = *reinterpret_cast<const float2*>(ring + sj + 2 * t);
      }
      fft_forward<kTw>(v, sm.W, fa, ad);
      // thread t holds Z[t + 256 j]. Publish for the pair step (in place, access C) and X[0], X[H/2].
#pragma unroll
      for (int q = 0; q < 16; ++q) sm.W[ad.pC + q] = v[q];
      if (t == 0) {
        sm.x0_xm[0] = v[0].x + v[0].y;  // X[0]   = Re Z0 + Im Z0
        sm.x0_xm[1] = v[0].x - v[0].y;  // X[H/2] = Re Z0 - Im Z0
      }
      __syncthreads();
      // ---- X: partner frequencies M - k. k = t + 256 j -> M - k = (256 - t) + 256 (15 - j) for t > 0,
      //         and 256 (16 - j) for t = 0 (j = 0 pairs with itself: Q[0] = 0).
      {
        const int pt = (kT - t) & (kT - 1);
        const float2* wp = sm.W + 273 * (pt & 15) + 17 * (pt >> 4);
        float2 zp[16];
        if (t == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) zp[j] = wp[(16 - j) & 15];
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) zp[j] = wp[15 - j];
        }
        __syncthreads();  // everyone has read its partner values; W may be overwritten
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          // E = (Z[k] + conj Z[M-k]) / 2, O = (Z[k] - conj Z[M-k]) / (2j); w = W_H^k = wh * W_32^j
          const float2 z = v[j], c = make_float2(zp[j].x, -zp[j].y);
          const float2 E = make_float2(0.5f * (z.x + c.x), 0.5f * (z.y + c.y));
          const float2 d = make_float2(z.x - c.x, z.y - c.y);
          const float2 O = make_float2(0.5f * d.y, -0.5f * d.x);  // d / (2j)
          const float2 w32 = make_float2(kW32c[j], -kW32s[j]);  // W_32^j, literal constants after unrolling
          const float2 w = cmul(wh, w32);
          const float2 P1 = cmul_conj(E, w);                      // conj(w) * E
          const float2 wO = cmul(w, O);
          const float2 P2 = make_float2(-wO.y, wO.x);             // j * w * O
          v[j] = make_float2(P1.x - P2.x, P1.y - P2.y);           // Q[k]
        }
        if (t == 0) v[0] = make_float2(0.0f, 0.0f);                // Q[0] = 0 (DC removed)
      }
      // ---- I: q = IFFT_M(Q) (unnormalised), natural order in W
      fft_inverse_to_smem<kTw>(v, sm.W, fa, ad);
      // ---- G0: c[n] for n = t + 256 j. Re = (H x[off+n] - X0 + (-1)^n XM) / 2, Im = y[off+n],
      //          y[2m] = Re q[m], y[2m+1] = Im q[m]
      float2 c[16];
      {
        const float half_x0 = 0.5f * sm.x0_xm[0], half_xm = 0.5f * sm.x0_xm[1];
        const float bias = sign * half_xm - half_x0;
        // q[m], m = (off + n)/2 = 1024 + 128 j + u, u = t/2: phys(m) = (1092 + u + u/16) + 136 j + j/2
        const int u = t >> 1;
        const float* Wf = reinterpret_cast<const float*>(sm.W) + 2 * (1092 + u + (u >> 4)) + (t & 1);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int sj = ring_wrap(r0 + off + kT * j, ring_len);  // CTA-uniform
          c[j].x = fmaf((float)kM, ring[sj + t], bias);           // H/2 = M
          c[j].y = Wf[2 * (136 * j + (j >> 1))];
        }
      }
      __syncthreads();  // all q reads done before the first window transform overwrites W
      // ---- G: three windowed transforms; keep bins t + 256 j (j < 8) and bin 2048 (j = 8, thread 0)
      float2 S[kGroups];
      float nd[kGroups];  // Im(D conj S) = D.im S.re - D.re S.im, spectrogram/processor.rs:470
#pragma unroll 1
      for (int wsel = 0; wsel < 3; ++wsel) {
        const float* win = wsel == 1 ? a.dwin : a.win;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = t + kT * j;
          float wv = __ldg(&win[n]);
          if (wsel == 2) wv = ((float)n - (float)(kM - 1) * 0.5f) * wv;  // t*h, spectrogram/processor.rs:601-608
          v[j] = make_float2(c[j].x * wv, c[j].y * wv);
        }
        fft_forward<kTw>(v, sm.W, fa, ad);
        if (wsel == 0) {
#pragma unroll
          for (int j = 0; j < kGroups; ++j) S[j] = v[j];
        } else if (wsel == 1) {
#pragma unroll
          for (int j = 0; j < kGroups; ++j) nd[j] = v[j].y * S[j].x - v[j].x * S[j].y;
        }
        __syncthreads();  // pass-3 reads done before the next transform's pass-1 stores
      }
      // ---- R: per-bin reassignment (v = time-ramp spectrum) + ordered compaction
      omb_spectrogram_point pts[kGroups];
      int rank[kGroups];
      unsigned keep = 0;
#pragma unroll
      for (int j = 0; j < kGroups; ++j) {
        const int bin = t + kT * j;
        bool k = (j < 8 || t == 0);
        if (k) k = reassign_bin_nd(S[j], nd[j], v[j], __ldg(&a.bin_norm[bin < (int)a.bins ? bin : 0]), bin, rc, &pts[j]);
        const unsigned m = __ballot_sync(0xffffffffu, k);
        if (lane_id == 0) sm.warp_cnt[j * kWarps + warp] = __popc(m);
        if (k) keep |= 1u << j;
        rank[j] = __popc(m & ((1u << lane_id) - 1u));  // rank inside the warp for this group
      }
      __syncthreads();
      if (warp == 0) {  // exclusive prefix over the 72 (group, warp) counts, in output order
        int c0 = 0, c1 = 0, c2 = 0;
        const int i0 = lane_id * 3;
        if (i0 + 0 < kGroups * kWarps) c0 = sm.warp_cnt[i0 + 0];
        if (i0 + 1 < kGroups * kWarps) c1 = sm.warp_cnt[i0 + 1];
        if (i0 + 2 < kGroups * kWarps) c2 = sm.warp_cnt[i0 + 2];
        const int tot = c0 + c1 + c2;
        int incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int n = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane_id >= o) incl += n;
        }
        const int excl = incl - tot;
        if (i0 + 0 < kGroups * kWarps) sm.offs[i0 + 0] = excl;
        if (i0 + 1 < kGroups * kWarps) sm.offs[i0 + 1] = excl + c0;
        if (i0 + 2 < kGroups * kWarps) sm.offs[i0 + 2] = excl + c0 + c1;
        if (lane_id == 31) sm.offs[kGroups * kWarps] = incl;
      }
      __syncthreads();
      {
        const uint64_t slot = lane * a.frames_per_lane + f;
        omb_spectrogram_point* out = a.out_points + slot * a.point_stride;
#pragma unroll
        for (int j = 0; j < kGroups; ++j)
          if (keep & (1u << j)) out[sm.offs[j * kWarps + warp] + rank[j]] = pts[j];
        if (t == 0) a.out_counts[slot] = (uint32_t)sm.offs[kGroups * kWarps];
      }
    }
    async_wait_all();
    __syncthreads();  // the ring is reused by the next run
  }
}

}  // namespace

bool stft_fast_supported(const StftConfig& cfg, const DeviceInfo& dev) {
  if (!cfg.reassign || cfg.window != (uint64_t)kM || cfg.zero_pad != 1) return false;
  const uint64_t H = 2 * (uint64_t)kM;
  if (cfg.hop < 512 || cfg.hop > (uint64_t)kM || (H % cfg.hop) != 0 || (cfg.hop % 512) != 0) return false;  // ring spans stay contiguous
  const size_t smem = sizeof(Smem) + (size_t)(H + cfg.hop) * sizeof(float);
  return dev.max_smem_optin == 0 || smem <= (size_t)dev.max_smem_optin;
}

int stft_fast_prepare(StftPlan& plan) {
  // twiddle tables: [15*256] W_4096^{b q} | [15*16] W_256^{o q} | [256] W_8192^{t}
  std::vector<float2> tab(15 * kT + 15 * 16 + kT);
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
  for (int t = 0; t < kT; ++t) {
    const double ang = -tau * (double)t / (2.0 * kM);
    tab[15 * kT + 15 * 16 + t] = make_float2((float)std::cos(ang), (float)std::sin(ang));
  }
  OMB_TRY(plan.d_fast_tables.upload(tab, plan.stream));
  const size_t smem = sizeof(Smem) + (size_t)(2 * kM + plan.cfg.hop) * sizeof(float);
  auto k10 = k_reassigned_fast<1, 0>;
  auto k20 = k_reassigned_fast<2, 0>;
  auto k11 = k_reassigned_fast<1, 1>;
  auto k21 = k_reassigned_fast<2, 1>;
  OMB_CUDA_TRY(cudaFuncSetAttribute(k10, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k20, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k11, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k21, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  return OMB_OK;
}

int launch_stft_fast(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s) {
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  if (per_lane == 0 || a.n_lanes == 0) return OMB_OK;
  if ((reinterpret_cast<uintptr_t>(a.lanes) & 15u) != 0 || (a.lane_stride % 4) != 0 || (a.first_frame * a.hop) % 4 != 0)
    return fail(OMB_ERR_INVALID, "specialised STFT kernel needs 16-byte aligned lanes (pointer and lane_stride % 4 == 0)");
  FastArgs fa{};
  fa.a = a;
  fa.tw1 = plan.d_fast_tables.ptr;
  fa.tw2 = fa.tw1 + 15 * kT;
  fa.twh = fa.tw2 + 15 * 16;
  fa.ring_len = (uint32_t)(2 * kM + a.hop);
  // runs: enough of them to balance ~2 CTAs per SM, long enough to amortise the 2N-sample ring prime
  static const int cta_per_sm = [] { const char* e = getenv("OMB_FAST_MINB"); return e ? std::max(1, atoi(e)) : 2; }();
  const uint64_t ctas = (uint64_t)std::max(plan.dev.sm_count, 1) * cta_per_sm;
  uint64_t run = 64;
  while (run > 8 && ((per_lane + run - 1) / run) * a.n_lanes < ctas * 8) run >>= 1;
  fa.frames_per_run = (uint32_t)std::min<uint64_t>(run, per_lane);
  fa.runs_per_lane = (uint32_t)((per_lane + fa.frames_per_run - 1) / fa.frames_per_run);
  const uint64_t total_runs = (uint64_t)fa.runs_per_lane * a.n_lanes;
  const unsigned grid = (unsigned)std::min<uint64_t>(total_runs, ctas);
  const size_t smem = sizeof(Smem) + (size_t)fa.ring_len * sizeof(float);
  // tuning knob (measurement only): OMB_FAST_MINB=1 trades occupancy (1 CTA/SM, no register cap) for zero spills
  static const int minb = [] { const char* e = getenv("OMB_FAST_MINB"); return e ? atoi(e) : 2; }();
  static const int twmode = [] { const char* e = getenv("OMB_FAST_TW"); return e ? atoi(e) : 0; }();
  auto k10 = k_reassigned_fast<1, 0>;
  auto k20 = k_reassigned_fast<2, 0>;
  auto k11 = k_reassigned_fast<1, 1>;
  auto k21 = k_reassigned_fast<2, 1>;
  if (minb == 1 && twmode == 0) {
    OMB_LAUNCH(k10, dim3(grid), dim3(kT), smem, s, fa);
  } else if (minb == 1) {
    OMB_LAUNCH(k11, dim3(grid), dim3(kT), smem, s, fa);
  } else if (twmode == 0) {
    OMB_LAUNCH(k20, dim3(grid), dim3(kT), smem, s, fa);
  } else {
    OMB_LAUNCH(k21, dim3(grid), dim3(kT), smem, s, fa);
  }
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
