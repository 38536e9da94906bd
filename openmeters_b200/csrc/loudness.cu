// loudness.cu — BS.1770 K-weighting, sliding mean squares and true peak (loudness/processor.rs, dsp.rs:264-371).
//
// Two implementations of the same arithmetic:
//
//  * k_loudness_stream_seq — the streaming LoudnessProcessor::process_block as the reference writes it: one thread per
//    channel walks the block sample by sample (TDF-II in f64, Neumaier compensated window sums with periodic re-base,
//    polyphase true-peak FIR in tap order).  Every step carries three dependent chains and four global-memory loads of
//    ring values, ~0.9 us per frame: 8.5 ms for a 9600-frame block.  Kept as the on-GPU cross-check (OMB_LOUDNESS_STREAM_SEQ=1).
//  * k_loudness_stream — the same arithmetic, operation for operation, with the three chains separated so that each runs at
//    its own dependency depth and the independent work runs in parallel: warp 0 the IIR (lane = channel), 6 warps the true
//    peak (a pure function of 12 / 24 consecutive input samples: parallel over samples, max-reduced), then warp 1 the four
//    window sums of every channel (lane = channel x window, ring loads independent of the chain), then all threads append
//    the block to the ring.  Bit-identical snapshots and state (max is order-free; every sum keeps its order).
//
//  * the batched plan — offline throughput path for long multi-channel streams (BASELINE cfg3).  The IIR
//    is linear, so time is cut into 256-sample chunks: (1) zero-state end state per chunk, (2) a short
//    serial scan propagating true chunk start states with A^256, (3) re-run each chunk from its true
//    start state writing y (f32) and the chunk's sum of squares; window means are then sums of whole
//    chunk sums plus two partial edges (all terms non-negative, f64: no cancellation, unlike a global
//    prefix sum); true peak is a per-sample FIR with an 11/23-sample halo, max-reduced per block.
#include "loudness.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace omb {

double channel_weight_host(uint8_t p) {
  switch (p) {
    case OMB_POS_LOW_FREQUENCY: return 0.0;
    case OMB_POS_REAR_LEFT: case OMB_POS_REAR_RIGHT: case OMB_POS_SIDE_LEFT: case OMB_POS_SIDE_RIGHT: return 1.41;
    default: return 1.0;
  }
}

namespace {

// loudness/processor.rs:153-162 — separate mul/add roundings (Rust never contracts).
__device__ __forceinline__ double kw_step(double x, double* s, const KWeight& kw) {
  const double y = __dadd_rn(__dmul_rn(kw.b[0], x), s[0]);
  s[0] = __dsub_rn(__dadd_rn(__dmul_rn(kw.b[1], x), s[1]), __dmul_rn(kw.a[1], y));
  s[1] = __dsub_rn(__dadd_rn(__dmul_rn(kw.b[2], x), s[2]), __dmul_rn(kw.a[2], y));
  s[2] = __dsub_rn(__dadd_rn(__dmul_rn(kw.b[3], x), s[3]), __dmul_rn(kw.a[3], y));
  s[3] = __dsub_rn(__dmul_rn(kw.b[4], x), __dmul_rn(kw.a[4], y));
  return y;
}

// The same section with fused multiply-adds (9 FP64 instructions instead of 17) for the chunk-parallel batch path:
// its chunk start states come out of a scan and are not bit-identical to the sequential recurrence's anyway, so
// per-operation rounding fidelity buys nothing there (f64 differences ~1e-16 relative, invisible after `y as f32`
// except on rounding boundaries; parity budget 1e-5 on the mean squares).  The streaming kernel keeps kw_step.
__device__ __forceinline__ double kw_step_fma(double x, double* s, const KWeight& kw) {
  const double y = fma(kw.b[0], x, s[0]);
  s[0] = fma(-kw.a[1], y, fma(kw.b[1], x, s[1]));
  s[1] = fma(-kw.a[2], y, fma(kw.b[2], x, s[2]));
  s[2] = fma(-kw.a[3], y, fma(kw.b[3], x, s[3]));
  s[3] = fma(-kw.a[4], y, kw.b[4] * x);
  return y;
}

// dsp.rs:277-285 Kahan-Babuska-Neumaier
__device__ __forceinline__ void neumaier_add(double& sum, double& corr, double v) {
  const double next = __dadd_rn(sum, v);
  corr = __dadd_rn(corr, fabs(sum) >= fabs(v) ? __dadd_rn(__dsub_rn(sum, next), v) : __dadd_rn(__dsub_rn(v, next), sum));
  sum = next;
}

__device__ __forceinline__ float lufs_dev(double ms, float floor_db) {  // loudness/processor.rs:57-66
  if (ms > 0.0) return (float)fmax(fma(log10(ms), 10.0, -0.691), (double)floor_db);
  return floor_db;
}

__device__ __forceinline__ float power_to_db_f(float p, float floor_db) {
  return p > 0.0f ? fmaxf(logf(p) * kLnToDb, floor_db) : floor_db;
}

// One CTA per stream (blockIdx.x; a single stream for omb_loudness, S lock-step streams for omb_loudness_bank).
__global__ void __launch_bounds__(32) k_loudness_stream_seq(LoudStreamArgs a) {
  const uint32_t c = threadIdx.x;
  // re-base the per-stream pointers; everything below is the single-stream code
  a.block += (uint64_t)blockIdx.x * a.block_stride;
  a.state += (uint64_t)blockIdx.x * a.channels;
  a.ring += (uint64_t)blockIdx.x * a.channels * a.ring_len;
  a.out += blockIdx.x;
  if (c < a.channels) {
    LoudChannelState st = a.state[c];
    double* ring = a.ring + (uint64_t)c * a.ring_len;
    const uint32_t dl = a.tp_delay_len;
    for (uint64_t f = 0; f < a.frames; ++f) {
      const float s = a.block[f * a.channels + c];
      if (!st.active) {  // lazy activation, loudness/processor.rs:264-274
        if (__float_as_uint(s) == 0u) {
          st.silent_frames += 1;
          continue;
        }
        st.active = 1;
        st.head = st.silent_frames % a.ring_len;  // WindowedMeans::with_leading_zeros, dsp.rs:359-365
        st.count = st.silent_frames < a.ring_len ? st.silent_frames : a.ring_len;
        for (int w = 0; w < kLoudWindows; ++w) {
          st.refresh[w] = st.silent_frames % a.caps[w];
          st.sums[w][0] = st.sums[w][1] = st.corr[w][0] = st.corr[w][1] = 0.0;
        }
        for (int i = 0; i < 4; ++i) st.filter[i] = 0.0;
        for (int i = 0; i < 48; ++i) st.delay[i] = 0.0f;
        st.write = dl;
        st.peak = 0.0f;
      }
      const float yf = (float)kw_step((double)s, st.filter, a.kw);
      double v = __dmul_rn((double)yf, (double)yf);
      if (!isfinite(v)) v = 0.0;  // dsp.rs:324-333
      // WindowedMeans::push, dsp.rs:334-357
      for (int w = 0; w < kLoudWindows; ++w) {
        const uint64_t cap = a.caps[w];
        const bool has_old = st.count >= cap;
        const double old = has_old ? ring[(st.head + a.ring_len - cap) % a.ring_len] : 0.0;
        neumaier_add(st.sums[w][0], st.corr[w][0], v);
        neumaier_add(st.sums[w][1], st.corr[w][1], v);
        if (has_old) neumaier_add(st.sums[w][0], st.corr[w][0], -old);
        if (++st.refresh[w] == cap) {  // CompensatedPair::refresh
          st.sums[w][0] = st.sums[w][1];
          st.sums[w][1] = 0.0;
          st.corr[w][0] = st.corr[w][1];
          st.corr[w][1] = 0.0;
          st.refresh[w] = 0;
        }
      }
      ring[st.head] = v;
      st.head = (st.head + 1) % a.ring_len;
      st.count = st.count + 1 < a.ring_len ? st.count + 1 : a.ring_len;
      // TruePeakMeter::process, loudness/processor.rs:123-150
      st.peak = fmaxf(st.peak, fabsf(s));
      if (dl) {
        st.write = (st.write == 0 ? dl : st.write) - 1;
        const uint32_t pos = st.write;
        st.delay[pos] = s;
        st.delay[pos + dl] = s;
        if (dl == 12) {
          float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f;
          for (uint32_t i = 0; i < 12; ++i) {
            const float d = st.delay[pos + i];
            o0 = __fadd_rn(o0, __fmul_rn(d, a.fir.fir4[i][0]));
            o1 = __fadd_rn(o1, __fmul_rn(d, a.fir.fir4[i][1]));
            o2 = __fadd_rn(o2, __fmul_rn(d, a.fir.fir4[i][2]));
          }
          st.peak = fmaxf(fmaxf(fmaxf(st.peak, fabsf(o0)), fabsf(o1)), fabsf(o2));
        } else {
          float o = 0.0f;
          for (uint32_t i = 0; i < 24; ++i) o = __fadd_rn(o, __fmul_rn(st.delay[pos + i], a.fir.fir2[i]));
          st.peak = fmaxf(st.peak, fabsf(o));
        }
      }
    }
    if (st.active)
      for (int i = 0; i < 4; ++i)
        if (fabs(st.filter[i]) < 1.0e-30) st.filter[i] = 0.0;  // level.rs:14-18
    a.state[c] = st;
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // snapshot, loudness/processor.rs:287-310
    omb_loudness_snapshot snap;
    const float floor_db = a.floor_db;
    for (int i = 0; i < OMB_MAX_CHANNELS; ++i) {
      snap.rms_fast_db[i] = snap.rms_slow_db[i] = snap.true_peak_db[i] = floor_db;
      snap.positions[i] = a.positions[i];
    }
    double wst = 0.0, wm = 0.0;
    for (uint32_t ch = 0; ch < a.channels; ++ch) {
      LoudChannelState& st = a.state[ch];
      if (!st.active) continue;
      double mean[kLoudWindows];
      for (int w = 0; w < kLoudWindows; ++w) {
        uint64_t n = st.count < a.caps[w] ? st.count : a.caps[w];
        if (n < 1) n = 1;
        mean[w] = (st.sums[w][0] + st.corr[w][0]) / (double)n;
      }
      wst = __dadd_rn(wst, __dmul_rn(mean[0], a.weights[ch]));
      wm = __dadd_rn(wm, __dmul_rn(mean[1], a.weights[ch]));
      snap.rms_fast_db[ch] = power_to_db_f((float)mean[2], floor_db);
      snap.rms_slow_db[ch] = power_to_db_f((float)mean[3], floor_db);
      const float peak = st.peak;
      st.peak = 0.0f;
      snap.true_peak_db[ch] = power_to_db_f(__fmul_rn(peak, peak), floor_db);
    }
    snap.short_term_loudness = lufs_dev(wst, floor_db);
    snap.momentary_loudness = lufs_dev(wm, floor_db);
    snap.channel_count = a.channels;
    *a.out = snap;
  }
}


// ---------------------------------------------------------------------------------- phase-parallel streaming kernel
constexpr int kStreamThreads = 256;   // warp 0: IIR, warp 1: window sums, warps 2-7: true peak

__device__ __forceinline__ void loud_snapshot(const LoudStreamArgs& a) {  // loudness/processor.rs:287-310 (thread 0)
  omb_loudness_snapshot snap;
  const float floor_db = a.floor_db;
  for (int i = 0; i < OMB_MAX_CHANNELS; ++i) {
    snap.rms_fast_db[i] = snap.rms_slow_db[i] = snap.true_peak_db[i] = floor_db;
    snap.positions[i] = a.positions[i];
  }
  double wst = 0.0, wm = 0.0;
  for (uint32_t ch = 0; ch < a.channels; ++ch) {
    LoudChannelState& st = a.state[ch];
    if (!st.active) continue;
    double mean[kLoudWindows];
    for (int w = 0; w < kLoudWindows; ++w) {
      uint64_t n = st.count < a.caps[w] ? st.count : a.caps[w];
      if (n < 1) n = 1;
      mean[w] = (st.sums[w][0] + st.corr[w][0]) / (double)n;
    }
    wst = __dadd_rn(wst, __dmul_rn(mean[0], a.weights[ch]));
    wm = __dadd_rn(wm, __dmul_rn(mean[1], a.weights[ch]));
    snap.rms_fast_db[ch] = power_to_db_f((float)mean[2], floor_db);
    snap.rms_slow_db[ch] = power_to_db_f((float)mean[3], floor_db);
    const float peak = st.peak;
    st.peak = 0.0f;
    snap.true_peak_db[ch] = power_to_db_f(__fmul_rn(peak, peak), floor_db);
  }
  snap.short_term_loudness = lufs_dev(wst, floor_db);
  snap.momentary_loudness = lufs_dev(wm, floor_db);
  snap.channel_count = a.channels;
  *a.out = snap;
}

// Tile pipeline.  A block is cut into tiles of kTile frames; in iteration i
//   warps 2-7  stage tile i + 1 of the input and the ring's old values of tile i in shared memory (coalesced, all loads in
//              flight at once) and then run the true-peak FIR of tile i,
//   warp 0     runs the K-weighting recurrence over tile i   (inputs from shared memory: no load inside the chain),
//   warp 1     runs the four compensated window sums of every channel over tile i - 1,
// one __syncthreads per iteration.  A first version with the three chains merely separated still spent 0.5 ms per 1024-frame
// block: every step of a chain waited for a global load (the ring of one stream is 9 MB: DRAM latency); measured in
// profiles/r02_notes.md.
constexpr int kTile = 128;
struct LoudTileSmem {
  float xs[2][kTile][OMB_MAX_CHANNELS];                        // input tile, [frame][channel] (the block's layout)
  double vs[2][OMB_MAX_CHANNELS][kTile];                       // y^2 of the tile
  double olds[2][OMB_MAX_CHANNELS * kLoudWindows][kTile];      // the ring value each (channel, window) retires at each step
};

__global__ void __launch_bounds__(kStreamThreads) k_loudness_stream(LoudStreamArgs a) {
  __shared__ unsigned long long s_start[OMB_MAX_CHANNELS];   // first frame of the block at which the channel is active
  __shared__ unsigned long long s_head0[OMB_MAX_CHANNELS], s_count0[OMB_MAX_CHANNELS];
  __shared__ unsigned s_peak[OMB_MAX_CHANNELS];               // bits of the block's peak (non-negative floats order as integers)
  __shared__ float s_hist[OMB_MAX_CHANNELS][24];              // [j]: the sample pushed j + 1 steps before the block's first active one
  OMB_DYN_SMEM(unsigned char, smem_raw);
  LoudTileSmem& sm = *reinterpret_cast<LoudTileSmem*>(smem_raw);
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  a.block += (uint64_t)blockIdx.x * a.block_stride;
  a.state += (uint64_t)blockIdx.x * a.channels;
  a.ring += (uint64_t)blockIdx.x * a.channels * a.ring_len;
  a.vnew += (uint64_t)blockIdx.x * a.channels * a.frames;
  a.out += blockIdx.x;
  const uint32_t C = a.channels, dl = a.tp_delay_len;
  const uint64_t frames = a.frames, L = a.ring_len;

  // ---- phase 0: lazy activation (loudness/processor.rs:264-274) and the constants of this block
  if (tid < C) {
    LoudChannelState& st = a.state[tid];
    uint64_t start = 0;
    if (!st.active) {
      uint64_t f = 0;
      while (f < frames && __float_as_uint(a.block[f * C + tid]) == 0u) ++f;
      st.silent_frames += f;
      start = f;
      if (f < frames) {
        st.active = 1;
        st.head = st.silent_frames % L;  // WindowedMeans::with_leading_zeros, dsp.rs:359-365
        st.count = st.silent_frames < L ? st.silent_frames : L;
        for (int w = 0; w < kLoudWindows; ++w) {
          st.refresh[w] = st.silent_frames % a.caps[w];
          st.sums[w][0] = st.sums[w][1] = st.corr[w][0] = st.corr[w][1] = 0.0;
        }
        for (int i = 0; i < 4; ++i) st.filter[i] = 0.0;
        for (int i = 0; i < 48; ++i) st.delay[i] = 0.0f;
        st.write = dl;
        st.peak = 0.0f;
      }
    }
    s_start[tid] = start;
    s_head0[tid] = st.head;
    s_count0[tid] = st.count;
    s_peak[tid] = 0u;
    for (uint32_t j = 0; j < dl; ++j) s_hist[tid][j] = st.delay[st.write + j];  // delay[write + i] = the sample pushed i steps ago
  }
  __syncthreads();

  const uint64_t n_tiles = (frames + kTile - 1) / kTile;
  auto load_x = [&](uint64_t tile, uint32_t first, uint32_t nthr) {  // input tile -> xs[tile & 1]
    const uint64_t f0 = tile * kTile;
    for (uint32_t i = first; i < kTile * C; i += nthr) {
      const uint64_t f = f0 + i / C;
      sm.xs[tile & 1][i / C][i % C] = f < frames ? a.block[f * C + i % C] : 0.0f;
    }
  };
  load_x(0, tid, kStreamThreads);
  // per-lane state of the two sequential roles
  double fz[4] = {0.0, 0.0, 0.0, 0.0};                  // warp 0, lane = channel
  double s0 = 0.0, s1 = 0.0, c0 = 0.0, c1 = 0.0;        // warp 1, lane = (channel, window)
  uint64_t refresh = 0;
  const uint32_t bc = lane / kLoudWindows, bw = lane % kLoudWindows;
  if (warp == 0 && lane < C) {
    const LoudChannelState& st = a.state[lane];
    for (int i = 0; i < 4; ++i) fz[i] = st.filter[i];
  }
  if (warp == 1 && lane < C * kLoudWindows) {
    const LoudChannelState& st = a.state[bc];
    s0 = st.sums[bw][0];
    s1 = st.sums[bw][1];
    c0 = st.corr[bw][0];
    c1 = st.corr[bw][1];
    refresh = st.refresh[bw];
  }
  __syncthreads();

  for (uint64_t it = 0; it <= n_tiles; ++it) {
    if (warp >= 2) {
      const uint32_t lt = tid - 64, nl = kStreamThreads - 64;
      if (it + 1 < n_tiles) load_x(it + 1, lt, nl);
      if (it < n_tiles) {
        // ring values retired during tile `it`: (channel, window) at step kk = k - start reads ring[(head0 + kk + L - cap) % L]
        // while kk < cap (later steps retire this block's own y^2, read by warp 1 from the scratch)
        const uint64_t f0 = it * kTile;
        for (uint32_t i = lt; i < kTile * C * kLoudWindows; i += nl) {
          const uint32_t row = i / kTile, t = i % kTile;
          const uint32_t c = row / kLoudWindows, w = row % kLoudWindows;
          const uint64_t k = f0 + t, start = s_start[c], cap = a.caps[w];
          double v = 0.0;
          if (k < frames && k >= start) {
            const uint64_t kk = k - start;
            if (s_count0[c] + kk >= cap && kk < cap) v = a.ring[(uint64_t)c * L + (s_head0[c] + kk + L - cap) % L];
          }
          sm.olds[it & 1][row][t] = v;
        }
        // ---- TruePeakMeter::process (loudness/processor.rs:123-150) for every (frame, channel) of the tile independently;
        //      each FIR sum in tap order (newest sample first), the block maximum by atomicMax on the bit pattern
        for (uint32_t i = lt; i < kTile * C; i += nl) {
          const uint64_t k = f0 + i / C;
          const uint32_t c = i % C;
          const uint64_t start = s_start[c];
          if (k >= frames || k < start) continue;
          const float s = a.block[k * C + c];
          float m = fabsf(s);
          m = m == m ? m : 0.0f;  // f32::max ignores NaN
          const uint64_t kk = k - start;
          auto tap = [&](uint32_t j) -> float {  // the sample pushed j steps before sample k
            return (uint64_t)j <= kk ? a.block[(k - j) * C + c] : s_hist[c][j - kk - 1];
          };
          if (dl == 12) {
            float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f;
#pragma unroll
            for (uint32_t j = 0; j < 12; ++j) {
              const float d = tap(j);
              o0 = __fadd_rn(o0, __fmul_rn(d, a.fir.fir4[j][0]));
              o1 = __fadd_rn(o1, __fmul_rn(d, a.fir.fir4[j][1]));
              o2 = __fadd_rn(o2, __fmul_rn(d, a.fir.fir4[j][2]));
            }
            m = fmaxf(fmaxf(fmaxf(m, fabsf(o0)), fabsf(o1)), fabsf(o2));
          } else if (dl == 24) {
            float o = 0.0f;
#pragma unroll
            for (uint32_t j = 0; j < 24; ++j) o = __fadd_rn(o, __fmul_rn(tap(j), a.fir.fir2[j]));
            m = fmaxf(m, fabsf(o));
          }
          atomicMax(&s_peak[c], __float_as_uint(m));
        }
      }
    } else if (warp == 0) {
      // ---- K-weighting over tile `it`, sequential per channel (loudness/processor.rs:153-162)
      if (it < n_tiles && lane < C) {
        const uint64_t f0 = it * kTile, start = s_start[lane];
        double* vn = a.vnew + (uint64_t)lane * frames;
#pragma unroll 4
        for (uint32_t t = 0; t < kTile; ++t) {
          const uint64_t k = f0 + t;
          if (k < frames && k >= start) {
            const float yf = (float)kw_step((double)sm.xs[it & 1][t][lane], fz, a.kw);
            double v = __dmul_rn((double)yf, (double)yf);
            if (!isfinite(v)) v = 0.0;  // dsp.rs:324-333
            sm.vs[it & 1][lane][t] = v;
            vn[k - start] = v;
          }
        }
      }
    } else if (it > 0 && lane < C * kLoudWindows) {
      // ---- WindowedMeans::push (dsp.rs:334-357) over tile `it - 1`, lane = (channel, window)
      const uint64_t tile = it - 1, f0 = tile * kTile, start = s_start[bc], cap = a.caps[bw], count0 = s_count0[bc];
      const double* vn = a.vnew + (uint64_t)bc * frames;
      const double* vsr = sm.vs[tile & 1][bc];
      const double* oldr = sm.olds[tile & 1][lane];
#pragma unroll 4
      for (uint32_t t = 0; t < kTile; ++t) {
        const uint64_t k = f0 + t;
        if (k < frames && k >= start) {
          const uint64_t kk = k - start;
          const double v = vsr[t];
          const bool has_old = count0 + kk >= cap;
          neumaier_add(s0, c0, v);
          neumaier_add(s1, c1, v);
          if (has_old) neumaier_add(s0, c0, -(kk >= cap ? vn[kk - cap] : oldr[t]));
          if (++refresh == cap) {  // CompensatedPair::refresh
            s0 = s1;
            s1 = 0.0;
            c0 = c1;
            c1 = 0.0;
            refresh = 0;
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- write-back of the sequential roles' state
  if (warp == 0 && lane < C && s_start[lane] < frames) {
    LoudChannelState& st = a.state[lane];
    for (int i = 0; i < 4; ++i) st.filter[i] = fabs(fz[i]) < 1.0e-30 ? 0.0 : fz[i];  // level.rs:14-18
  }
  if (warp == 1 && lane < C * kLoudWindows && s_start[bc] < frames) {
    LoudChannelState& st = a.state[bc];
    st.sums[bw][0] = s0;
    st.sums[bw][1] = s1;
    st.corr[bw][0] = c0;
    st.corr[bw][1] = c1;
    st.refresh[bw] = refresh;
  }

  // ---- phase C: append the block to the ring (only the last ring_len values of a longer block survive), cursors, delay line
  for (uint32_t c = 0; c < C; ++c) {
    const uint64_t start = s_start[c], n = frames - start;
    const double* vn = a.vnew + (uint64_t)c * frames;
    double* ring = a.ring + (uint64_t)c * L;
    for (uint64_t k = (n > L ? n - L : 0) + tid; k < n; k += kStreamThreads) ring[(s_head0[c] + k) % L] = vn[k];
  }
  if (tid >= 64 && tid < 64 + C && s_start[tid - 64] < frames) {
    const uint32_t c = tid - 64;
    LoudChannelState& st = a.state[c];
    const uint64_t start = s_start[c], n = frames - start;
    st.head = (s_head0[c] + n) % L;
    st.count = s_count0[c] + n < L ? s_count0[c] + n : L;
    st.peak = fmaxf(st.peak, __uint_as_float(s_peak[c]));
    if (dl) {
      // after n decrements-with-wrap the write cursor sits at (write0 - n) mod dl; delay[write + i] (and its copy dl further)
      // holds the sample pushed i steps ago
      const uint32_t w0 = st.write % dl;
      const uint32_t wn = (uint32_t)((w0 + dl - (n % dl)) % dl);
      for (uint32_t i = 0; i < dl; ++i) {
        const float val = (uint64_t)i < n ? a.block[(frames - 1 - i) * C + c] : s_hist[c][i - n];
        const uint32_t p = (wn + i) % dl;
        st.delay[p] = val;
        st.delay[p + dl] = val;
      }
      st.write = wn;
    }
  }
  __syncthreads();
  if (tid == 0) loud_snapshot(a);
}

// ---------------------------------------------------------------------------------- batched
struct LoudBatchArgs {
  const float* in;       // [stream][frame][channel]
  uint64_t stream_stride, frames, n_chunks, block_frames, n_blocks;
  uint32_t n_streams, channels, tp_delay_len;
  KWeight kw;
  double M[16];      // A^kKwChunk
  double Mseg[16];   // A^(kKwChunk*kKwSeg)
  double* seg_state; // [stream][segment][channel][4]
  double* end_state;     // [stream][chunk][channel][4]  zero-state end states
  double* start_state;   // [stream][chunk][channel][4]  true start states
  float* y;              // [stream][frame][channel] (K-weighted samples, the input's layout)
  double* csum;          // [stream][channel][chunk]
  const double* cbase;   // [stream][channel][chunk + 1] exclusive prefix sums of csum (null: sum the chunks directly)
  unsigned* peak;        // [stream][block][channel]  float bits (non-negative)
  uint64_t caps[kLoudWindows];
  double weights[OMB_MAX_CHANNELS];
  uint8_t positions[OMB_MAX_CHANNELS];
  float floor_db;
  omb_loudness_snapshot* out;  // [stream][block]
};

// (1)/(3): one thread per (stream, chunk, channel); channel fastest so a warp reads whole frames.
// The apply pass (kApply) writes y in the input's own interleaved layout [stream][frame][channel]: a warp's store of one
// time step is then whole 32-byte sectors (8 channels x 4 B), exactly like its load — no transposition needed — and the
// snapshot kernel's edge sums read one sector per frame for all channels of a window.
template <bool kApply>
__global__ void __launch_bounds__(128) k_kw_chunks(LoudBatchArgs a) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t total = (uint64_t)a.n_streams * a.n_chunks * a.channels;
  if (idx >= total) return;
  const uint32_t c = (uint32_t)(idx % a.channels);
  const uint64_t chunk = (idx / a.channels) % a.n_chunks;
  const uint64_t stream = idx / ((uint64_t)a.channels * a.n_chunks);
  const float* x = a.in + stream * a.stream_stride;
  const uint64_t t0 = chunk * kKwChunk;
  const uint64_t t1 = t0 + kKwChunk < a.frames ? t0 + kKwChunk : a.frames;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  const uint64_t sidx = ((stream * a.n_chunks + chunk) * a.channels + c) * 4;
  if (!kApply) {
    for (uint64_t t = t0; t < t1; ++t) kw_step((double)__ldg(&x[t * a.channels + c]), s, a.kw);
    a.end_state[sidx + 0] = s[0];
    a.end_state[sidx + 1] = s[1];
    a.end_state[sidx + 2] = s[2];
    a.end_state[sidx + 3] = s[3];
    return;
  }
  s[0] = a.start_state[sidx + 0];
  s[1] = a.start_state[sidx + 1];
  s[2] = a.start_state[sidx + 2];
  s[3] = a.start_state[sidx + 3];
  float* y = a.y + stream * a.frames * a.channels;
  double acc = 0.0;
#pragma unroll 4
  for (uint64_t t = t0; t < t1; ++t) {
    const float yf = (float)kw_step_fma((double)__ldg(&x[t * a.channels + c]), s, a.kw);
    const double v = (double)yf * (double)yf;
    acc += isfinite(v) ? v : 0.0;
    y[t * a.channels + c] = yf;
  }
  a.csum[(stream * a.channels + c) * a.n_chunks + chunk] = acc;
}

// (1'), the shipped zero-state pass: the chunk's end state from a zero start is a linear map of its inputs,
// s_end = sum_k A^(K-1-k) B x[k] — 4 FMAs per sample instead of the 17-operation recurrence.  The response table rides in
// the kernel parameters (constant bank) and the k loop is fully unrolled, so every coefficient is an immediate
// constant-bank operand of its DFMA: no load instruction, no L1 traffic for the table (a first version that read it with
// __ldg was bound by the L1 data pipe at 88 % and slower than the recurrence).  A short last chunk is right-aligned
// (zeros in front), which keeps k uniform across the warp.  Bound: the bandwidth of reading x once.
struct KwRespTable {
  double v[kKwChunk * 4];  // [k][r] = (A^(K-1-k) B)[r]
};

template <int kC>  // channel count at compile time (load offsets become immediates), 0 = any
__global__ void __launch_bounds__(128) k_kw_zero_state(LoudBatchArgs a, const KwRespTable tab) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t C = kC ? (uint32_t)kC : a.channels;
  const uint64_t total = (uint64_t)a.n_streams * a.n_chunks * C;
  if (idx >= total) return;
  const uint32_t c = (uint32_t)(idx % C);
  const uint64_t chunk = (idx / C) % a.n_chunks;
  const uint64_t stream = idx / ((uint64_t)C * a.n_chunks);
  const uint64_t t0 = chunk * kKwChunk;
  const uint64_t t1 = t0 + kKwChunk < a.frames ? t0 + kKwChunk : a.frames;
  const int lead = kKwChunk - (int)(t1 - t0);  // zeros in front of a short last chunk
  // sample k of the right-aligned chunk is stream frame t1 - K + k (a possibly "negative" pointer that is never dereferenced)
  const float* x = a.in + stream * a.stream_stride + c + ((int64_t)t1 - kKwChunk) * (int64_t)C;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  if (lead == 0) {
#pragma unroll
    for (int k = 0; k < kKwChunk; ++k) {
      const double xv = (double)__ldg(&x[(uint32_t)k * C]);
      s0 = fma(tab.v[4 * k + 0], xv, s0);
      s1 = fma(tab.v[4 * k + 1], xv, s1);
      s2 = fma(tab.v[4 * k + 2], xv, s2);
      s3 = fma(tab.v[4 * k + 3], xv, s3);
    }
  } else {
#pragma unroll 8
    for (int k = 0; k < kKwChunk; ++k) {
      const double xv = k >= lead ? (double)__ldg(&x[(uint32_t)k * C]) : 0.0;
      s0 = fma(tab.v[4 * k + 0], xv, s0);
      s1 = fma(tab.v[4 * k + 1], xv, s1);
      s2 = fma(tab.v[4 * k + 2], xv, s2);
      s3 = fma(tab.v[4 * k + 3], xv, s3);
    }
  }
  const uint64_t sidx = ((stream * a.n_chunks + chunk) * C + c) * 4;
  a.end_state[sidx + 0] = s0;
  a.end_state[sidx + 1] = s1;
  a.end_state[sidx + 2] = s2;
  a.end_state[sidx + 3] = s3;
}

// (2): scan over chunks of the affine recurrence start[k+1] = M * start[k] + end0[k], three levels so the
// dependent chain is 64 + n_seg + 64 steps instead of n_chunks (a 30 s stream has 5625 chunks; the first version's
// single serial walk was 3.7 ms of pure load latency):
//   kPhase 0: per (stream, channel, segment of kKwSeg chunks) — segment composite from a zero start  -> seg_state
//   kPhase 1: per (stream, channel) — serial over segments with M^kKwSeg                              -> seg_state (in place: starts)
//   kPhase 2: per (stream, channel, segment) — re-walk the segment from its true start                -> start_state
__device__ __forceinline__ void affine_step(double (&s)[4], const double* __restrict__ M, const double* __restrict__ e) {
  double n[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) n[r] = M[r * 4 + 0] * s[0] + M[r * 4 + 1] * s[1] + M[r * 4 + 2] * s[2] + M[r * 4 + 3] * s[3] + e[r];
#pragma unroll
  for (int r = 0; r < 4; ++r) s[r] = n[r];
}

template <int kPhase>
__global__ void __launch_bounds__(128) k_kw_scan(LoudBatchArgs a) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t n_seg = (a.n_chunks + kKwSeg - 1) / kKwSeg;
  if (kPhase == 1) {
    if (idx >= (uint64_t)a.n_streams * a.channels) return;
    const uint32_t c = (uint32_t)(idx % a.channels);
    const uint64_t stream = idx / a.channels;
    double s[4] = {0.0, 0.0, 0.0, 0.0};
    for (uint64_t g = 0; g < n_seg; ++g) {
      double* p = a.seg_state + ((stream * n_seg + g) * a.channels + c) * 4;
      const double e[4] = {p[0], p[1], p[2], p[3]};
      p[0] = s[0]; p[1] = s[1]; p[2] = s[2]; p[3] = s[3];
      affine_step(s, a.Mseg, e);
    }
    return;
  }
  const uint64_t total = (uint64_t)a.n_streams * n_seg * a.channels;
  if (idx >= total) return;
  const uint32_t c = (uint32_t)(idx % a.channels);
  const uint64_t g = (idx / a.channels) % n_seg;
  const uint64_t stream = idx / ((uint64_t)a.channels * n_seg);
  double* sp = a.seg_state + ((stream * n_seg + g) * a.channels + c) * 4;
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  if (kPhase == 2) { s[0] = sp[0]; s[1] = sp[1]; s[2] = sp[2]; s[3] = sp[3]; }
  const uint64_t k0 = g * kKwSeg, k1 = k0 + kKwSeg < a.n_chunks ? k0 + kKwSeg : a.n_chunks;
  const uint64_t stride = (uint64_t)a.channels * 4;
  const double* __restrict__ e = a.end_state + ((stream * a.n_chunks + k0) * a.channels + c) * 4;
  double* __restrict__ st = a.start_state + ((stream * a.n_chunks + k0) * a.channels + c) * 4;
#pragma unroll 4
  for (uint64_t k = k0; k < k1; ++k, e += stride, st += stride) {
    if (kPhase == 2) { st[0] = s[0]; st[1] = s[1]; st[2] = s[2]; st[3] = s[3]; }
    affine_step(s, a.M, e);
  }
  if (kPhase == 0) { sp[0] = s[0]; sp[1] = s[1]; sp[2] = s[2]; sp[3] = s[3]; }
}

// true peak: a CTA takes 256 frames of one (stream, block): the frames plus a 23-frame halo are loaded coalesced and
// stored channel-major in shared memory, then every thread runs the polyphase FIR of its frame for each channel from
// conflict-free rows (taps in the reference's order, separate mul/add roundings: bit-identical to the reference loop).
constexpr int kTpTile = 256, kTpHalo = 23, kTpRow = kTpTile + kTpHalo + 1;
__global__ void __launch_bounds__(kTpTile) k_true_peak(LoudBatchArgs a, TruePeakFir fir) {
  __shared__ float tile[OMB_MAX_CHANNELS][kTpRow];
  const uint64_t sb = blockIdx.x;  // stream * n_blocks + block
  const uint64_t stream = sb / a.n_blocks, blk = sb % a.n_blocks;
  const float* x = a.in + stream * a.stream_stride;
  const uint64_t f0 = blk * a.block_frames;
  const uint64_t f1 = f0 + a.block_frames < a.frames ? f0 + a.block_frames : a.frames;
  const int lane = threadIdx.x & 31;
  const int C = (int)a.channels;
  for (uint64_t base = f0 + (uint64_t)blockIdx.y * kTpTile; base < f1; base += (uint64_t)gridDim.y * kTpTile) {
    // tile frame index u <-> stream frame base - kTpHalo + u; frames before the stream start are zeros
    const int n_el = (kTpTile + kTpHalo) * C;
    for (int e = threadIdx.x; e < n_el; e += kTpTile) {
      const int u = e / C, c = e - u * C;
      const int64_t fr = (int64_t)base - kTpHalo + u;
      tile[c][u] = (fr >= 0 && (uint64_t)fr < a.frames) ? __ldg(&x[(uint64_t)fr * C + c]) : 0.0f;
    }
    __syncthreads();
    const uint64_t t = base + threadIdx.x;
    const int u = kTpHalo + threadIdx.x;
    for (int c = 0; c < C; ++c) {
      float pk = 0.0f;
      if (t < f1) {
        const float* row = &tile[c][u];
        pk = fabsf(row[0]);
        if (a.tp_delay_len == 12) {
          float o0 = 0.0f, o1 = 0.0f, o2 = 0.0f;
#pragma unroll
          for (int i = 0; i < 12; ++i) {  // delay[pos+i] == x[t-i]
            const float d = row[-i];
            o0 = __fadd_rn(o0, __fmul_rn(d, fir.fir4[i][0]));
            o1 = __fadd_rn(o1, __fmul_rn(d, fir.fir4[i][1]));
            o2 = __fadd_rn(o2, __fmul_rn(d, fir.fir4[i][2]));
          }
          pk = fmaxf(fmaxf(fmaxf(pk, fabsf(o0)), fabsf(o1)), fabsf(o2));
        } else if (a.tp_delay_len == 24) {
          float o = 0.0f;
#pragma unroll
          for (int i = 0; i < 24; ++i) o = __fadd_rn(o, __fmul_rn(row[-i], fir.fir2[i]));
          pk = fmaxf(pk, fabsf(o));
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) pk = fmaxf(pk, __shfl_xor_sync(0xffffffffu, pk, o));
      // NaN input: fmaxf drops NaN exactly like Rust's f32::max in TruePeakMeter::process.
      if (lane == 0 && pk > 0.0f) atomicMax(&a.peak[sb * C + c], __float_as_uint(pk));
    }
    __syncthreads();
  }
}

// 4x true peak, register-blocked (sample rates below 96 kHz — the common case): the tile is staged like above, then
// every thread owns kTpRun consecutive frames of ONE channel, keeps their 11-sample history in registers and runs the
// three 12-tap phases with scalar multiplies and adds in the reference's order (separate roundings, bit-identical sums).
// (Packed FP32x2 does not help: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 — one rounding instead of two, even
// with -fmad=false — and the bit-exact alternative, packed multiplies of two adjacent frames (FMUL2) with scalar adds,
// was measured: 20 % fewer instructions, the same 0.52 ms, because FMUL2 holds the FMA pipe for two cycles and the kernel
// is bound by that pipe (67 % of its peak), not by issue slots; profiles/r01c_notes.md.)
// A warp covers 32 runs of one channel: the smem row is padded 9-for-8 so the stride-8 reads are conflict-free, and
// the max is reduced once per kTpRun frames instead of once per frame.  ~60 instructions per sample-channel instead
// of ~190 (profiles/r01b_notes.md).
constexpr int kTpRun = 8, kTp4Halo = 11;
constexpr int kTp4Row = (kTpTile + kTp4Halo) + (kTpTile + kTp4Halo) / 8 + 2;

template <int kC>  // channel count known at compile time (index arithmetic of the staging loop), 0 = any
__global__ void __launch_bounds__(kTpTile) k_true_peak4(LoudBatchArgs a, TruePeakFir fir) {
  __shared__ float tile[OMB_MAX_CHANNELS][kTp4Row];
  const uint64_t sb = blockIdx.x;  // stream * n_blocks + block
  const uint64_t stream = sb / a.n_blocks, blk = sb % a.n_blocks;
  const float* x = a.in + stream * a.stream_stride;
  const uint64_t f0 = blk * a.block_frames;
  const uint64_t f1 = f0 + a.block_frames < a.frames ? f0 + a.block_frames : a.frames;
  const int C = kC ? kC : (int)a.channels;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;  // warp = channel (C <= 8 warps take part in the FIR)
  const bool vec4 = ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);  // 16-byte aligned stream: frames are float4 pairs
  float best = 0.0f;
  for (uint64_t base = f0 + (uint64_t)blockIdx.y * kTpTile; base < f1; base += (uint64_t)gridDim.y * kTpTile) {
    // tile frame index u <-> stream frame base - kTp4Halo + u; frames before the stream start are zeros
    if (kC == 8 && vec4) {
      // 8 interleaved channels = two 16-byte quads per frame: one bounds check and one load per quad
      for (int e = threadIdx.x; e < (kTpTile + kTp4Halo) * 2; e += kTpTile) {
        const int u = e >> 1, c = (e & 1) * 4;
        const int64_t fr = (int64_t)base - kTp4Halo + u;
        float4 q = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (fr >= 0 && (uint64_t)fr < a.frames) q = __ldg(reinterpret_cast<const float4*>(x + (uint64_t)fr * 8 + c));
        const int col = u + (u >> 3);
        tile[c + 0][col] = q.x;
        tile[c + 1][col] = q.y;
        tile[c + 2][col] = q.z;
        tile[c + 3][col] = q.w;
      }
    } else {
      const int n_el = (kTpTile + kTp4Halo) * C;
      for (int e = threadIdx.x; e < n_el; e += kTpTile) {
        const int u = e / C, c = e - u * C;
        const int64_t fr = (int64_t)base - kTp4Halo + u;
        tile[c][u + (u >> 3)] = (fr >= 0 && (uint64_t)fr < a.frames) ? __ldg(&x[(uint64_t)fr * C + c]) : 0.0f;
      }
    }
    __syncthreads();
    if (warp < C) {
      // this thread's frames: base + 8 lane + r, r < 8; window w[k] = tile frame 8 lane + k, k < 19 (w[11 + r] is frame r)
      const float* row = tile[warp];
      const float* lrow = row + 9 * lane;  // this thread's window starts at tile frame 8 lane -> column 9 lane
      const int nvalid = (int)(f1 - base < (uint64_t)kTpTile ? f1 - base : (uint64_t)kTpTile) - kTpRun * lane;  // frames of this run inside the block
      float w[kTpRun + kTp4Halo];
#pragma unroll
      for (int k = 0; k < kTpRun + kTp4Halo; ++k) w[k] = lrow[k + (k >> 3)];  // tile column of window index k: 9 lane + k + (k >> 3)
#pragma unroll
      for (int r = 0; r < kTpRun; ++r) {
        // first tap: 0.0 + p == p (up to the sign of zero, which |.| discards), so the sums start at the product
        float o0 = __fmul_rn(w[kTp4Halo + r], fir.fir4[0][0]), o1 = __fmul_rn(w[kTp4Halo + r], fir.fir4[0][1]),
              o2 = __fmul_rn(w[kTp4Halo + r], fir.fir4[0][2]);
#pragma unroll
        for (int i = 1; i < 12; ++i) {  // delay[pos + i] == x[t - i]
          const float d = w[kTp4Halo + r - i];
          o0 = __fadd_rn(o0, __fmul_rn(d, fir.fir4[i][0]));
          o1 = __fadd_rn(o1, __fmul_rn(d, fir.fir4[i][1]));
          o2 = __fadd_rn(o2, __fmul_rn(d, fir.fir4[i][2]));
        }
        // NaN input: fmaxf drops NaN exactly like Rust's f32::max in TruePeakMeter::process.
        const float pk = fmaxf(fmaxf(fmaxf(fabsf(w[kTp4Halo + r]), fabsf(o0)), fabsf(o1)), fabsf(o2));
        if (r < nvalid) best = fmaxf(best, pk);
      }
    }
    __syncthreads();
  }
  if (warp < C) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0 && best > 0.0f) atomicMax(&a.peak[sb * C + warp], __float_as_uint(best));
  }
}

// Exclusive prefix sums of the chunk sums, one CTA per (stream, channel) row: cbase[row][k] = sum of csum[row][0..k).
// Every thread owns a contiguous slice (local sum -> block scan of the 256 slice sums -> prefix inside the slice).
// f64: a 0.3 s window out of an hour of audio loses < 1e-12 relative to the subtraction.
__global__ void __launch_bounds__(256) k_csum_prefix(LoudBatchArgs a, double* cbase) {
  __shared__ double part[256];
  const uint64_t row = blockIdx.x;
  const int t = threadIdx.x;
  const double* cs = a.csum + row * a.n_chunks;
  double* out = cbase + row * (a.n_chunks + 1);
  const uint64_t per = (a.n_chunks + 255) / 256;
  const uint64_t k0 = (uint64_t)t * per, k1 = k0 + per < a.n_chunks ? k0 + per : a.n_chunks;
  double local = 0.0;
  for (uint64_t k = k0; k < k1; ++k) local += cs[k];
  part[t] = local;
  __syncthreads();
  if (t < 32) {  // exclusive scan of 256 partial sums by one warp, 8 per lane
    double v[8], run = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = run;
      run += part[t * 8 + i];
    }
    double incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double n = __shfl_up_sync(0xffffffffu, incl, o);
      if (t >= o) incl += n;
    }
    const double base = incl - run;
#pragma unroll
    for (int i = 0; i < 8; ++i) part[t * 8 + i] = base + v[i];
  }
  __syncthreads();
  double acc = part[t];
  for (uint64_t k = k0; k < k1; ++k) {
    out[k] = acc;
    acc += cs[k];
  }
  if (t == 255) out[a.n_chunks] = part[255] + local;  // exclusive prefix of the last slice + the slice itself = total
}

// Sum of y^2 over frames [t0, t1) of one (stream, channel): whole chunk sums + two partial edges.
__device__ double window_sum(const LoudBatchArgs& a, uint64_t stream, uint32_t c, uint64_t t0, uint64_t t1) {
  if (t1 <= t0) return 0.0;
  const float* y = a.y + stream * a.frames * a.channels + c;  // interleaved [frame][channel]
  const uint64_t C = a.channels;
  const double* cs = a.csum + (stream * a.channels + c) * a.n_chunks;
  const uint64_t k0 = (t0 + kKwChunk - 1) / kKwChunk;  // first chunk fully inside
  const uint64_t k1 = t1 / kKwChunk;                   // one past the last chunk fully inside
  double acc = 0.0;
  if (k0 >= k1) {  // no whole chunk inside
    for (uint64_t t = t0; t < t1; ++t) {
      const double v = (double)y[t * C] * (double)y[t * C];
      acc += isfinite(v) ? v : 0.0;
    }
    return acc;
  }
  for (uint64_t t = t0; t < k0 * kKwChunk; ++t) {
    const double v = (double)y[t * C] * (double)y[t * C];
    acc += isfinite(v) ? v : 0.0;
  }
  if (a.cbase) {
    const double* cb = a.cbase + (stream * a.channels + c) * (a.n_chunks + 1);
    acc += cb[k1] - cb[k0];
  } else {
    for (uint64_t k = k0; k < k1; ++k) acc += cs[k];
  }
  for (uint64_t t = k1 * kKwChunk; t < t1; ++t) {
    const double v = (double)y[t * C] * (double)y[t * C];
    acc += isfinite(v) ? v : 0.0;
  }
  return acc;
}

// One warp per snapshot: lane = channel*4 + window computes one sliding mean; lane 0 assembles.
__global__ void __launch_bounds__(128) k_loud_snapshots(LoudBatchArgs a) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const uint64_t total = (uint64_t)a.n_streams * a.n_blocks;
  const bool live = warp < total;
  const uint64_t stream = live ? warp / a.n_blocks : 0, blk = live ? warp % a.n_blocks : 0;
  const uint64_t t_end = (blk + 1) * a.block_frames < a.frames ? (blk + 1) * a.block_frames : a.frames;
  const uint32_t c = (uint32_t)(lane >> 2);
  const int w = lane & 3;
  double mean = 0.0;
  if (live && c < a.channels) {
    // divisor: min(count, cap) with count = frames seen since the stream start (leading silence counts,
    // dsp.rs:359-370), at least 1
    uint64_t n = t_end < a.caps[w] ? t_end : a.caps[w];
    if (n < 1) n = 1;
    const uint64_t t0 = t_end > a.caps[w] ? t_end - a.caps[w] : 0;
    mean = window_sum(a, stream, c, t0, t_end) / (double)n;
  }
  // gather the 4 means of every channel to lane 0
  double m[OMB_MAX_CHANNELS][kLoudWindows];
#pragma unroll
  for (int ch = 0; ch < OMB_MAX_CHANNELS; ++ch)
#pragma unroll
    for (int ww = 0; ww < kLoudWindows; ++ww) m[ch][ww] = __shfl_sync(0xffffffffu, mean, ch * 4 + ww);
  if (live && lane == 0) {
    omb_loudness_snapshot snap;
    const float floor_db = a.floor_db;
    double wst = 0.0, wm = 0.0;
    for (int ch = 0; ch < OMB_MAX_CHANNELS; ++ch) {
      snap.rms_fast_db[ch] = snap.rms_slow_db[ch] = snap.true_peak_db[ch] = floor_db;
      snap.positions[ch] = a.positions[ch];
      if (ch < (int)a.channels) {
        wst += m[ch][0] * a.weights[ch];
        wm += m[ch][1] * a.weights[ch];
        snap.rms_fast_db[ch] = power_to_db_f((float)m[ch][2], floor_db);
        snap.rms_slow_db[ch] = power_to_db_f((float)m[ch][3], floor_db);
        const float peak = __uint_as_float(a.peak[warp * a.channels + ch]);
        snap.true_peak_db[ch] = power_to_db_f(__fmul_rn(peak, peak), floor_db);
      }
    }
    snap.short_term_loudness = lufs_dev(wst, floor_db);
    snap.momentary_loudness = lufs_dev(wm, floor_db);
    snap.channel_count = a.channels;
    a.out[warp] = snap;
  }
}

void mat4_mul(const long double* A, const long double* B, long double* C) {
  long double t[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      long double acc = 0;
      for (int k = 0; k < 4; ++k) acc += A[r * 4 + k] * B[k * 4 + c];
      t[r * 4 + c] = acc;
    }
  for (int i = 0; i < 16; ++i) C[i] = t[i];
}

}  // namespace

int launch_loudness_stream(const LoudStreamArgs& a, cudaStream_t s, uint32_t n_streams) {
  if (!n_streams) return OMB_OK;
  const char* e = getenv("OMB_LOUDNESS_STREAM_SEQ");  // read per launch: the cross-check tests flip it inside one process
  const bool seq = e && e[0] == '1';
  if (seq) {
    OMB_LAUNCH(k_loudness_stream_seq, dim3(n_streams), dim3(32), 0, s, a);
  } else {
    static const bool attr = [] {
      return cudaFuncSetAttribute(k_loudness_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LoudTileSmem)) == cudaSuccess;
    }();
    if (!attr) return fail(OMB_ERR_CUDA, "cudaFuncSetAttribute(k_loudness_stream) failed");
    OMB_LAUNCH(k_loudness_stream, dim3(n_streams), dim3(kStreamThreads), sizeof(LoudTileSmem), s, a);
  }
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

LoudnessPlan::~LoudnessPlan() {
  if (stream) cudaStreamDestroy(stream);
  if (side) cudaStreamDestroy(side);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
}

int LoudnessPlan::init(const omb_loudness_config& c, uint32_t ch, const uint8_t* pos) {
  cfg = c;
  OMB_TRY(current_device(&dev));
  sample_rate = sanitize_sample_rate(c.sample_rate);
  channels = std::min<uint32_t>(std::max<uint32_t>(ch, 1), OMB_MAX_CHANNELS);
  if (pos) std::memcpy(positions, pos, OMB_MAX_CHANNELS);
  else fallback_positions_host(channels, positions);
  k_weighting_host((double)sample_rate, kw.b, kw.a);
  true_peak_fir4_host(fir.fir4);
  true_peak_fir2_host(fir.fir2);
  static const float kWindows[kLoudWindows] = {3.0f, 0.4f, 0.3f, 1.0f};  // loudness/processor.rs:13
  for (int w = 0; w < kLoudWindows; ++w) caps[w] = std::max<uint64_t>(loudness_window_length(sample_rate, kWindows[w]), 1);
  tp_delay_len = (double)sample_rate < 96000.0 ? 12 : ((double)sample_rate < 192000.0 ? 24 : 0);
  // zero-input state transition of the TDF-II section, raised to the chunk length (kKwChunk = 2^8)
  long double A[16] = {-(long double)kw.a[1], 1, 0, 0, -(long double)kw.a[2], 0, 1, 0, -(long double)kw.a[3], 0, 0, 1, -(long double)kw.a[4], 0, 0, 0};
  static_assert(kKwChunk == 256, "chunk matrix uses 8 squarings");
  for (int i = 0; i < 8; ++i) mat4_mul(A, A, A);
  for (int i = 0; i < 16; ++i) chunk_matrix[i] = (double)A[i];
  {  // response of the chunk end state to each input sample: v_{K-1} = B, v_{k-1} = A v_k  (state' = A state + B x)
    const long double b0 = kw.b[0];
    const long double A1[16] = {-(long double)kw.a[1], 1, 0, 0, -(long double)kw.a[2], 0, 1, 0, -(long double)kw.a[3], 0, 0, 1, -(long double)kw.a[4], 0, 0, 0};
    long double v[4] = {(long double)kw.b[1] - (long double)kw.a[1] * b0, (long double)kw.b[2] - (long double)kw.a[2] * b0,
                        (long double)kw.b[3] - (long double)kw.a[3] * b0, (long double)kw.b[4] - (long double)kw.a[4] * b0};
    h_resp.assign((size_t)kKwChunk * 4, 0.0);
    for (int k = kKwChunk - 1; k >= 0; --k) {
      for (int r = 0; r < 4; ++r) h_resp[(size_t)k * 4 + r] = (double)v[r];
      long double n[4];
      for (int r = 0; r < 4; ++r) n[r] = A1[r * 4 + 0] * v[0] + A1[r * 4 + 1] * v[1] + A1[r * 4 + 2] * v[2] + A1[r * 4 + 3] * v[3];
      for (int r = 0; r < 4; ++r) v[r] = n[r];
    }
  }
  static_assert(kKwSeg == 64, "segment matrix uses 6 more squarings");
  for (int i = 0; i < 6; ++i) mat4_mul(A, A, A);
  for (int i = 0; i < 16; ++i) seg_matrix[i] = (double)A[i];
  OMB_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  return OMB_OK;
}

int LoudnessPlan::execute_device(const float* d_interleaved, uint32_t n_streams, uint64_t frames, uint64_t stream_stride,
                                 uint64_t block_frames, omb_loudness_snapshot* d_out_snap, cudaStream_t s) {
  if (!n_streams || !frames) return OMB_OK;
  if (!d_interleaved || !d_out_snap || block_frames == 0) return fail(OMB_ERR_INVALID, "invalid argument");
  OMB_CUDA_TRY(cudaSetDevice(dev.device));
  LoudBatchArgs a{};
  a.in = d_interleaved;
  a.stream_stride = stream_stride;
  a.frames = frames;
  a.n_chunks = (frames + kKwChunk - 1) / kKwChunk;
  a.block_frames = block_frames;
  a.n_blocks = (frames + block_frames - 1) / block_frames;
  a.n_streams = n_streams;
  a.channels = channels;
  a.tp_delay_len = tp_delay_len;
  a.kw = kw;
  std::memcpy(a.M, chunk_matrix, sizeof a.M);
  std::memcpy(a.Mseg, seg_matrix, sizeof a.Mseg);
  const uint64_t n_seg = (a.n_chunks + kKwSeg - 1) / kKwSeg;
  OMB_TRY(d_seg.reserve((size_t)((uint64_t)n_streams * n_seg * channels * 4)));
  a.seg_state = d_seg.ptr;
  const uint64_t n_state = (uint64_t)n_streams * a.n_chunks * channels * 4;
  OMB_TRY(d_end.reserve((size_t)n_state));
  OMB_TRY(d_start.reserve((size_t)n_state));
  OMB_TRY(d_y.reserve((size_t)((uint64_t)n_streams * channels * frames)));
  OMB_TRY(d_csum.reserve((size_t)((uint64_t)n_streams * channels * a.n_chunks)));
  OMB_TRY(d_peak.reserve((size_t)((uint64_t)n_streams * a.n_blocks * channels)));
  OMB_TRY(d_cbase.reserve((size_t)((uint64_t)n_streams * channels * (a.n_chunks + 1))));
  a.end_state = d_end.ptr;
  a.start_state = d_start.ptr;
  a.y = d_y.ptr;
  a.csum = d_csum.ptr;
  a.peak = d_peak.ptr;
  for (int w = 0; w < kLoudWindows; ++w) a.caps[w] = caps[w];
  for (uint32_t c = 0; c < OMB_MAX_CHANNELS; ++c) {
    a.weights[c] = channel_weight_host(positions[c]);
    a.positions[c] = positions[c];
  }
  a.floor_db = cfg.floor_db;
  a.out = d_out_snap;
  OMB_CUDA_TRY(cudaMemsetAsync(d_peak.ptr, 0, sizeof(unsigned) * n_streams * a.n_blocks * channels, s));

  // True peak depends on the PCM (and the zeroed peak array) only: on a side stream, next to the K-weighting chain, joined before the
  // snapshots (OMB_LOUDNESS_OVERLAP=0: one stream, round 1's order)
  const char* ov_env = getenv("OMB_LOUDNESS_OVERLAP");
  const bool overlap = !(ov_env && ov_env[0] == '0');
  cudaStream_t ts = s;
  if (overlap) {
    if (!side) {
      OMB_CUDA_TRY(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
      OMB_CUDA_TRY(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      OMB_CUDA_TRY(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    OMB_CUDA_TRY(cudaEventRecord(ev_fork, s));
    OMB_CUDA_TRY(cudaStreamWaitEvent(side, ev_fork, 0));
    ts = side;
  }
  const unsigned gx = (unsigned)std::min<uint64_t>((block_frames + kTpTile - 1) / kTpTile, 64);
  if (tp_delay_len == 12 && !getenv("OMB_NO_TRUE_PEAK4")) {
    // enough (stream, block) pairs to fill the GPU: one CTA walks all tiles of its block (set-up and the final reduction
    // are paid once per block instead of once per 256 frames)
    const unsigned gy = (uint64_t)n_streams * a.n_blocks >= (uint64_t)std::max(dev.sm_count, 1) * 16 ? 1u : gx;
    const dim3 grid((unsigned)((uint64_t)n_streams * a.n_blocks), gy);
    if (channels == 8) {
      OMB_LAUNCH(k_true_peak4<8>, grid, dim3(kTpTile), 0, ts, a, fir);
    } else if (channels == 2) {
      OMB_LAUNCH(k_true_peak4<2>, grid, dim3(kTpTile), 0, ts, a, fir);
    } else {
      OMB_LAUNCH(k_true_peak4<0>, grid, dim3(kTpTile), 0, ts, a, fir);
    }
  } else {
    OMB_LAUNCH(k_true_peak, dim3((unsigned)((uint64_t)n_streams * a.n_blocks), gx), dim3(kTpTile), 0, ts, a, fir);
  }
  OMB_CHECK_LAUNCH();
  if (overlap) OMB_CUDA_TRY(cudaEventRecord(ev_join, side));
  const uint64_t n_items = (uint64_t)n_streams * a.n_chunks * channels;
  const unsigned g1 = (unsigned)((n_items + 127) / 128);
  auto k_apply = k_kw_chunks<true>;
  if (getenv("OMB_KW_ZERO_RECURRENCE")) {  // the recurrence form of the zero-state pass (A/B measurements)
    auto k_zero = k_kw_chunks<false>;
    OMB_LAUNCH(k_zero, dim3(g1), dim3(128), 0, s, a);
  } else {
    KwRespTable tab;
    std::memcpy(tab.v, h_resp.data(), sizeof tab.v);
    if (channels == 8) {
      OMB_LAUNCH(k_kw_zero_state<8>, dim3(g1), dim3(128), 0, s, a, tab);
    } else if (channels == 2) {
      OMB_LAUNCH(k_kw_zero_state<2>, dim3(g1), dim3(128), 0, s, a, tab);
    } else {
      OMB_LAUNCH(k_kw_zero_state<0>, dim3(g1), dim3(128), 0, s, a, tab);
    }
  }
  OMB_CHECK_LAUNCH();
  {
    auto k_s0 = k_kw_scan<0>;
    auto k_s1 = k_kw_scan<1>;
    auto k_s2 = k_kw_scan<2>;
    const uint64_t n_segments = (a.n_chunks + kKwSeg - 1) / kKwSeg;
    const unsigned gs = (unsigned)(((uint64_t)n_streams * n_segments * channels + 127) / 128);
    OMB_LAUNCH(k_s0, dim3(gs), dim3(128), 0, s, a);
    OMB_CHECK_LAUNCH();
    OMB_LAUNCH(k_s1, dim3((unsigned)(((uint64_t)n_streams * channels + 127) / 128)), dim3(128), 0, s, a);
    OMB_CHECK_LAUNCH();
    OMB_LAUNCH(k_s2, dim3(gs), dim3(128), 0, s, a);
    OMB_CHECK_LAUNCH();
  }
  OMB_LAUNCH(k_apply, dim3(g1), dim3(128), 0, s, a);
  OMB_CHECK_LAUNCH();
  OMB_LAUNCH(k_csum_prefix, dim3((unsigned)((uint64_t)n_streams * channels)), dim3(256), 0, s, a, d_cbase.ptr);
  OMB_CHECK_LAUNCH();
  a.cbase = d_cbase.ptr;
  if (overlap) OMB_CUDA_TRY(cudaStreamWaitEvent(s, ev_join, 0));
  const uint64_t n_snap = (uint64_t)n_streams * a.n_blocks;
  OMB_LAUNCH(k_loud_snapshots, dim3((unsigned)((n_snap * 32 + 127) / 128)), dim3(128), 0, s, a);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

int LoudnessPlan::execute_host(const float* h_interleaved, uint32_t n_streams, uint64_t frames, uint64_t stream_stride,
                               uint64_t block_frames, omb_loudness_snapshot* h_out) {
  if (!n_streams || !frames) return OMB_OK;
  if (!h_interleaved || !h_out || block_frames == 0) return fail(OMB_ERR_INVALID, "invalid argument");
  OMB_CUDA_TRY(cudaSetDevice(dev.device));
  const uint64_t per = frames * channels;
  OMB_TRY(d_in.reserve((size_t)(per * n_streams)));
  for (uint32_t i = 0; i < n_streams; ++i)
    OMB_CUDA_TRY(cudaMemcpyAsync(d_in.ptr + i * per, h_interleaved + i * stream_stride, sizeof(float) * per, cudaMemcpyHostToDevice, stream));
  const uint64_t n_blocks = (frames + block_frames - 1) / block_frames;
  OMB_TRY(d_out.reserve((size_t)(n_blocks * n_streams)));
  OMB_TRY(execute_device(d_in.ptr, n_streams, frames, per, block_frames, d_out.ptr, stream));
  OMB_CUDA_TRY(cudaMemcpyAsync(h_out, d_out.ptr, sizeof(omb_loudness_snapshot) * n_blocks * n_streams, cudaMemcpyDeviceToHost, stream));
  OMB_CUDA_TRY(cudaStreamSynchronize(stream));
  return OMB_OK;
}

}  // namespace omb
