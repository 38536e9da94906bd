// spectrum.cu — spectrum-analyzer kernels and plan (spectrum/processor.rs:179-253,332-425; state.rs:321-325).
//
//   k_spectrum_power_generic : per (lane, hop) CTA — DC-remove + window -> FFT -> |X|^2 * norm      (rows a4, a12)
//   k_spectrum_smooth        : per (lane, bin) thread, sequential over hops — None / Exponential /
//                              PeakHold smoothing, state floor, raw + A-weighted dB, fused arg-max      (rows a13, f3)
//
// The smoothing recurrence is sequential in time per bin, parallel over (lane, bin): coalesced over
// bins, HBM-bound (4 B in, 8 B out per bin-hop).  Hops are processed in chunks sized so the power
// scratch written by the first kernel is still L2-resident when the second one reads it.
#include "device_math.cuh"
#include "fft_stockham.cuh"
#include "spectrum.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace omb {

namespace {

constexpr int kThreads = 256;

__device__ void block_fft_radix2_sp(float2* data, int n, int logn, const float2* __restrict__ tw) {
  const int tid = threadIdx.x, nt = blockDim.x;
  if (n <= 1) return;
  for (int i = tid; i < n; i += nt) {
    const int j = (int)(__brev((unsigned)i) >> (32 - logn));
    if (i < j) {
      const float2 a = data[i], b = data[j];
      data[i] = b;
      data[j] = a;
    }
  }
  __syncthreads();
  for (int s = 1; s <= logn; ++s) {
    const int half = 1 << (s - 1), step = n >> s;
    for (int i = tid; i < (n >> 1); i += nt) {
      const int k = i & (half - 1);
      const int base = ((i >> (s - 1)) << s) + k;
      const float2 w = __ldg(&tw[k * step]);
      const float2 a = data[base], b = data[base + half];
      const float2 t = cmul(b, w);
      data[base] = cadd(a, t);
      data[base + half] = csub(a, t);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) k_spectrum_power_generic(SpectrumPowerArgs a) {
  __shared__ float red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  float2* work = a.scratch + (uint64_t)blockIdx.x * a.scratch_stride;
  const uint64_t total = a.hops * a.n_lanes;
  const int N = (int)a.fft_size;
  for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const uint64_t lane = item / a.hops, h = item % a.hops;
    const float* x = a.lanes + lane * a.lane_stride + h * a.hop;
    float part = 0.0f;
    for (int i = tid; i < N; i += nt) part += __ldg(&x[i]);
    const float mean = block_sum(part, red) / (float)N;
    for (int i = tid; i < N; i += nt) work[i] = make_float2((__ldg(&x[i]) - mean) * __ldg(&a.win[i]), 0.0f);
    __syncthreads();
    block_fft_radix2_sp(work, N, (int)a.log2_fft, a.tw);
    float* out = a.power + item * a.bins;
    for (int k = tid; k < (int)a.bins; k += nt) {
      const float2 z = work[k];
      out[k] = (z.x * z.x + z.y * z.y) * __ldg(&a.bin_norm[k]);
    }
    __syncthreads();
  }
}

// Shared-memory variant for N <= 16384: the real N-point transform as one complex N/2-point Stockham FFT of
// (r[2n], r[2n+1]) + the split X[k] = E[k] + W_N^k O[k] (same scheme as k_classic_smem).
__global__ void __launch_bounds__(1024) k_spectrum_power_smem(SpectrumPowerArgs a) {
  OMB_DYN_SMEM(float2, smem);
  __shared__ float red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = (int)a.fft_size, M = N >> 1, logM = (int)a.log2_fft - 1;
  float2* A = smem;
  float2* B = smem + M;
  const uint64_t total = a.hops * a.n_lanes;
  for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const uint64_t lane = item / a.hops, h = item % a.hops;
    const float* x = a.lanes + lane * a.lane_stride + h * a.hop;
    float part = 0.0f;
    for (int i = tid; i < N; i += nt) part += __ldg(&x[i]);
    const float mean = block_sum(part, red) / (float)N;
    for (int n = tid; n < M; n += nt) {
      const float r0 = (__ldg(&x[2 * n]) - mean) * __ldg(&a.win[2 * n]);
      const float r1 = (__ldg(&x[2 * n + 1]) - mean) * __ldg(&a.win[2 * n + 1]);
      A[n] = make_float2(r0, r1);
    }
    __syncthreads();
    const float2* Z = stockham_fft(A, B, M, logM, a.tw, 2);
    float* out = a.power + item * a.bins;
    for (int k = tid; k <= M; k += nt) {
      const float2 zk = Z[k & (M - 1)];
      const float2 zm = Z[(M - k) & (M - 1)];
      const float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
      const float2 O = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
      const float2 w = k < M ? __ldg(&a.tw[k]) : make_float2(-1.0f, 0.0f);
      const float2 X = cadd(E, cmul(w, O));
      out[k] = (X.x * X.x + X.y * X.y) * __ldg(&a.bin_norm[k]);
    }
    __syncthreads();
  }
}

// Monotone map float -> uint so that unsigned compare == float compare (finite values).
__device__ __forceinline__ unsigned ordered_bits(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct SmoothLayout {
  uint64_t out_hops_total, out_hop0;  // where this chunk's hops land in the caller's [lane][hop][bin] arrays
};

__global__ void __launch_bounds__(kThreads) k_spectrum_smooth(SpectrumSmoothArgs a, SmoothLayout lay) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = blockIdx.y;
  const bool live = k < (int)a.bins;
  const int lane_id = threadIdx.x & 31;
  float st = 0.0f;
  if (live && a.state && a.mode != OMB_AVG_NONE) st = a.state[(uint64_t)lane * a.bins + k];
  const float aw = live ? __ldg(&a.a_db[k]) : 0.0f;
  const float one_minus_alpha = __fsub_rn(1.0f, a.alpha);
  const bool interior = live && k >= a.peak_lo && k <= a.peak_hi;  // peak_bin candidates (state.rs:321-324)
  constexpr int kAhead = 8;
  for (uint64_t h0 = 0; h0 < a.hops; h0 += kAhead) {
    // the recurrence is sequential over hops, the loads are not: fetch kAhead hops of power before using them
    float pw[kAhead];
#pragma unroll
    for (int i = 0; i < kAhead; ++i)
      pw[i] = (live && h0 + i < a.hops) ? __ldg(&a.power[((uint64_t)lane * a.hops + h0 + i) * a.bins + k]) : 0.0f;
#pragma unroll
    for (int i = 0; i < kAhead; ++i) {
      const uint64_t h = h0 + i;
      if (h >= a.hops) break;
      const bool emit = a.write_all || h + 1 == a.hops;  // dB values are only needed where something is stored
      float raw = a.floor_db, weighted = a.floor_db;
      if (live) {
        const float p = pw[i];
        float v = p;
        if (a.mode == OMB_AVG_EXPONENTIAL) {  // spectrum/processor.rs:366-377 (re-seed when the average hit zero)
          st = st <= 0.0f ? p : __fadd_rn(__fmul_rn(st, a.alpha), __fmul_rn(p, one_minus_alpha));
          if (st < a.state_floor) st = 0.0f;
          v = st;
        } else if (a.mode == OMB_AVG_PEAK_HOLD) {  // :380-388
          st = fmaxf(__fmul_rn(st, a.decay), p);
          if (st < a.state_floor) st = 0.0f;
          v = st;
        }
        if (emit) {
          if (!(v < a.state_floor)) {  // :392-401
            const float db = __fmul_rn(logf(v), kLnToDb);
            raw = fmaxf(db, a.floor_db);
            weighted = fmaxf(__fadd_rn(db, aw), a.floor_db);
          }
          const uint64_t o = a.write_all ? (((uint64_t)lane * lay.out_hops_total + lay.out_hop0 + h) * a.bins + k)
                                         : ((uint64_t)lane * a.bins + k);
          a.out_weighted[o] = weighted;
          a.out_raw[o] = raw;
        }
      }
      if (a.peak_keys && emit) {  // spectrum/state.rs:321-325: bins 1..len-2 inside [min_f, max_f], finite, last maximum wins
        // two warp-wide integer max reductions (REDUX): the largest ordered dB value, then the largest bin holding it
        const float pv = a.peak_raw ? raw : weighted;
        const unsigned ob = (interior && isfinite(pv)) ? ordered_bits(pv) : 0u;
        const unsigned best = __reduce_max_sync(0xffffffffu, ob);
        const unsigned bin = __reduce_max_sync(0xffffffffu, (ob == best) ? (unsigned)k : 0u);
        if (lane_id == 0 && best)
          atomicMax(&a.peak_keys[(uint64_t)lane * lay.out_hops_total + lay.out_hop0 + h], ((unsigned long long)best << 32) | bin);
      }
    }
  }
  if (live && a.state && a.mode != OMB_AVG_NONE) a.state[(uint64_t)lane * a.bins + k] = st;
}

// interpolated_peak (spectrum/state.rs:327-356), one thread per row; every operation is a separately rounded f32
// operation in the reference's order (no contraction), so identical dB inputs give identical bits.
__global__ void k_peak_interpolate(const float* db, const int32_t* peak_bin, uint64_t rows, uint32_t bins, float bin_hz, float* out_freq,
                                   float* out_level) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float nan = __int_as_float(0x7fc00000);
  float freq = nan, level = nan;
  const int32_t bin = peak_bin[i];
  // :328-337 — bin == 0 or no right neighbour, unusable bin spacing, non-finite centre => None
  if (bin >= 1 && (uint32_t)bin + 1u < bins && isfinite(bin_hz) && bin_hz > 0.0f) {
    const float* row = db + i * (uint64_t)bins;
    const float center = row[bin];
    const float center_freq = __fmul_rn((float)bin, bin_hz);  // frequency_bins[bin] (spectrum/processor.rs:141-146)
    if (isfinite(center) && isfinite(center_freq)) {
      const float left = row[bin - 1], right = row[bin + 1];
      float offset = 0.0f;
      if (isfinite(left) && isfinite(right)) {  // :340-350
        const float denom = __fadd_rn(__fsub_rn(left, __fmul_rn(2.0f, center)), right);
        if (denom < -1e-6f) offset = fminf(fmaxf(__fdiv_rn(__fmul_rn(0.5f, __fsub_rn(left, right)), denom), -0.5f), 0.5f);
      }
      level = offset == 0.0f ? center  // :351-355
                             : fmaxf(__fsub_rn(center, __fmul_rn(__fmul_rn(0.25f, __fsub_rn(left, right)), offset)), center);
      freq = fmaxf(__fadd_rn(center_freq, __fmul_rn(offset, bin_hz)), 0.0f);
    }
  }
  out_freq[i] = freq;
  out_level[i] = level;
}

__global__ void k_peak_keys_to_bins(const unsigned long long* keys, uint64_t n, int32_t* out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long key = keys[i];
  out[i] = key ? (int32_t)(key & 0xffffffffu) : -1;
}

}  // namespace

SpectrumConfigN SpectrumConfigN::from_c(const omb_spectrum_config& c) {
  SpectrumConfigN o;
  o.sample_rate = sanitize_sample_rate(c.sample_rate);
  o.window_kind = c.window <= OMB_WINDOW_BLACKMAN_HARRIS ? c.window : (uint32_t)OMB_WINDOW_RECTANGULAR;
  o.fft_size = std::max<uint64_t>(c.fft_size, 1);
  o.hop = c.hop_size == 0 ? std::max<uint64_t>(o.fft_size / 16, 1) : c.hop_size;
  o.averaging = c.averaging <= OMB_AVG_PEAK_HOLD ? c.averaging : (uint32_t)OMB_AVG_NONE;
  o.averaging_param = c.averaging_param;
  o.source = c.source <= OMB_CHANNEL_NONE ? c.source : (uint32_t)OMB_CHANNEL_NONE;
  o.secondary = c.secondary_source <= OMB_CHANNEL_NONE ? c.secondary_source : (uint32_t)OMB_CHANNEL_NONE;
  o.floor_db = sanitize_negative_db(c.floor_db, -100.0f);
  return o;
}

void SpectrumConfigN::to_c(omb_spectrum_config* out) const {
  std::memset(out, 0, sizeof *out);
  out->sample_rate = sample_rate;
  out->window = window_kind;
  out->fft_size = fft_size;
  out->hop_size = hop;
  out->averaging = averaging;
  out->averaging_param = averaging_param;
  out->source = source;
  out->secondary_source = secondary;
  out->floor_db = floor_db;
}

SpectrumPlan::~SpectrumPlan() {
  if (stream) cudaStreamDestroy(stream);
}

int SpectrumPlan::init(const omb_spectrum_config& c) {
  cfg = SpectrumConfigN::from_c(c);
  OMB_TRY(current_device(&dev));
  const uint64_t N = cfg.fft_size;
  if (!is_pow2(N) || N > (1ull << 24))
    return fail(OMB_ERR_UNSUPPORTED, "spectrum fft_size %llu: only power-of-two lengths have kernels (no CPU fallback)", (unsigned long long)N);
  h_win = make_window((int)cfg.window_kind, (size_t)N);
  h_norm = make_bin_norm(h_win.data(), (size_t)N, (size_t)N);
  const size_t bins = (size_t)cfg.bins();
  const float bin_hz = cfg.sample_rate / (float)N;  // spectrum/processor.rs:138-146
  h_freq.resize(bins);
  h_adb.resize(bins);
  for (size_t b = 0; b < bins; ++b) {
    h_freq[b] = (float)b * bin_hz;
    h_adb[b] = a_weight_host(h_freq[b]);
  }
  state_floor = smoothing_state_floor_host(h_adb, cfg.floor_db);
  OMB_TRY(set_peak_spec(peak_spec));
  OMB_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  OMB_TRY(d_win.upload(h_win, stream));
  OMB_TRY(d_norm.upload(h_norm, stream));
  OMB_TRY(d_adb.upload(h_adb, stream));
  OMB_TRY(d_tw.upload(make_twiddles((size_t)N, (size_t)std::max<uint64_t>(N / 2, 1)), stream));
  fast16k = false;
  if (spectrum_fast_supported(cfg, dev) && !getenv("OMB_NO_FAST_SPECTRUM")) {
    OMB_TRY(spectrum_fast_prepare(*this));
    fast16k = true;
  }
  OMB_CUDA_TRY(cudaStreamSynchronize(stream));
  return OMB_OK;
}

int SpectrumPlan::power_device(const float* d_lanes, uint32_t n_lanes, uint64_t hops, uint64_t lane_stride, float* d_power_out,
                               cudaStream_t s) {
  const uint64_t total = hops * n_lanes;
  if (!total) return OMB_OK;
  OMB_CUDA_TRY(cudaSetDevice(dev.device));
  SpectrumPowerArgs a{};
  a.lanes = d_lanes;
  a.lane_stride = lane_stride;
  a.n_lanes = n_lanes;
  a.hops = hops;
  a.fft_size = (uint32_t)cfg.fft_size;
  a.hop = (uint32_t)cfg.hop;
  a.bins = (uint32_t)cfg.bins();
  a.log2_fft = (uint32_t)ilog2(cfg.fft_size);
  a.win = d_win.ptr;
  a.bin_norm = d_norm.ptr;
  a.tw = d_tw.ptr;
  a.power = d_power_out;
  if (fast16k && (reinterpret_cast<uintptr_t>(d_lanes) & 7u) == 0 && (lane_stride & 1) == 0) return launch_spectrum_power_fast(*this, a, s);
  const uint64_t N = cfg.fft_size;
  const size_t smem = (size_t)N * sizeof(float2);  // two buffers of N/2 complex
  if (N >= 16 && N <= 16384 && !getenv("OMB_NO_SMEM_KERNEL") && (dev.max_smem_optin == 0 || smem + 1024 <= (size_t)dev.max_smem_optin)) {
    OMB_CUDA_TRY(cudaFuncSetAttribute(k_spectrum_power_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)200 * 1024 / smem));
    const unsigned grid = (unsigned)std::min<uint64_t>(total, (uint64_t)std::max(dev.sm_count, 1) * per_sm);
    // whole warps only: block_sum / block_rank use full-mask warp collectives (2048 / 6 = 341 threads broke them)
    const uint64_t want = std::max<uint64_t>(256, (2048 / (uint64_t)per_sm + 31) / 32 * 32);
    const unsigned threads = (unsigned)std::min<uint64_t>(std::min<uint64_t>(1024, want), std::max<uint64_t>(32, N / 8));
    OMB_LAUNCH(k_spectrum_power_smem, dim3(grid), dim3(threads), smem, s, a);
    OMB_CHECK_LAUNCH();
    return OMB_OK;
  }
  const unsigned grid = (unsigned)std::min<uint64_t>(total, (uint64_t)std::max(dev.sm_count, 1) * 4);
  a.scratch_stride = cfg.fft_size;
  OMB_TRY(d_scratch.reserve((size_t)(a.scratch_stride * grid)));
  a.scratch = d_scratch.ptr;
  OMB_LAUNCH(k_spectrum_power_generic, dim3(grid), dim3(kThreads), 0, s, a);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

static int smooth_launch(SpectrumPlan& p, const float* d_power_in, uint32_t n_lanes, uint64_t hops, float* d_state, float* d_weighted,
                         float* d_raw, unsigned long long* keys, bool write_all, uint64_t hops_total, uint64_t hop0, cudaStream_t s) {
  if (!hops || !n_lanes) return OMB_OK;
  const SpectrumConfigN& cfg = p.cfg;
  SpectrumSmoothArgs a{};
  a.power = d_power_in;
  a.n_lanes = n_lanes;
  a.hops = hops;
  a.bins = (uint32_t)cfg.bins();
  a.mode = (int)cfg.averaging;
  a.alpha = std::min(std::max(cfg.averaging_param, 0.0f), 0.9999f);
  const float dt = (float)cfg.hop / cfg.sample_rate;  // spectrum/processor.rs:184
  a.decay = db_to_power_host(-std::fmax(cfg.averaging_param, 0.0f) * dt);
  a.state_floor = p.state_floor;
  a.floor_db = cfg.floor_db;
  a.a_db = p.d_adb.ptr;
  a.state = d_state;
  a.out_weighted = d_weighted;
  a.out_raw = d_raw;
  a.write_all = write_all ? 1 : 0;
  a.peak_keys = keys;  // caller-owned [n_lanes * hops_total], zeroed
  a.peak_raw = p.peak_spec.trace == 1u ? 1 : 0;
  a.peak_lo = p.peak_lo;
  a.peak_hi = p.peak_hi;
  if (keys && !write_all) return fail(OMB_ERR_INVALID, "peak bins are only produced together with per-hop outputs");
  SmoothLayout lay{hops_total, hop0};
  const dim3 grid((unsigned)((a.bins + kThreads - 1) / kThreads), n_lanes);
  OMB_LAUNCH(k_spectrum_smooth, grid, dim3(kThreads), 0, s, a, lay);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

static int keys_begin(SpectrumPlan& p, uint64_t n, cudaStream_t s) {
  OMB_TRY(p.d_keys.reserve((size_t)n));
  OMB_CUDA_TRY(cudaMemsetAsync(p.d_keys.ptr, 0, sizeof(unsigned long long) * n, s));
  return OMB_OK;
}
static int keys_finish(SpectrumPlan& p, uint64_t n, int32_t* d_peak_bin, cudaStream_t s) {
  OMB_LAUNCH(k_peak_keys_to_bins, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, p.d_keys.ptr, n, d_peak_bin);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

int SpectrumPlan::smooth_device(const float* d_power_in, uint32_t n_lanes, uint64_t hops, float* d_state_io, float* d_weighted,
                                float* d_raw, int32_t* d_peak_bin, bool write_all, cudaStream_t s) {
  OMB_CUDA_TRY(cudaSetDevice(dev.device));  // plans are bound to the device they were created on
  const uint64_t n = hops * n_lanes;
  if (d_peak_bin) OMB_TRY(keys_begin(*this, n, s));
  OMB_TRY(smooth_launch(*this, d_power_in, n_lanes, hops, d_state_io, d_weighted, d_raw, d_peak_bin ? d_keys.ptr : nullptr, write_all, hops, 0, s));
  if (d_peak_bin && n) OMB_TRY(keys_finish(*this, n, d_peak_bin, s));
  return OMB_OK;
}

int SpectrumPlan::execute_device(const float* d_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride,
                                 float* d_weighted, float* d_raw, int32_t* d_peak_bin, cudaStream_t s) {
  const uint64_t hops = cfg.hops_for(samples_per_lane);
  if (!hops || !n_lanes) return OMB_OK;
  if (!d_lanes || !d_weighted || !d_raw) return fail(OMB_ERR_INVALID, "null argument");
  if (n_lanes > 65535u) return fail(OMB_ERR_UNSUPPORTED, "at most 65535 lanes per spectrum call (got %u): split the batch", n_lanes);
  OMB_CUDA_TRY(cudaSetDevice(dev.device));  // plans are bound to the device they were created on
  const uint64_t bins = cfg.bins();
  // Enough lanes to fill the GPU with one CTA per lane: the fused kernel (power spectrum never leaves the SM).
  // OMB_SPECTRUM_FUSED=0/1 pins the choice (measurement / cross-checks).
  {
    const char* e = getenv("OMB_SPECTRUM_FUSED");
    const int pin = e ? atoi(e) : -1;
    const bool aligned = (reinterpret_cast<uintptr_t>(d_lanes) & 15u) == 0 && (lane_stride % 4) == 0;
    const bool enough = (uint64_t)n_lanes * 2 >= (uint64_t)std::max(dev.sm_count, 1);
    if (fast16k && fused16k && aligned && pin != 0 && (enough || pin == 1))
      return launch_spectrum_fused(*this, d_lanes, n_lanes, hops, lane_stride, d_weighted, d_raw, d_peak_bin, s);
  }
  // chunk of hops whose power scratch (n_lanes*chunk*bins*4 B) stays comfortably inside the 126 MB L2
  const uint64_t budget_floats = (64ull << 20) / 4;
  uint64_t chunk = std::max<uint64_t>(1, budget_floats / std::max<uint64_t>(1, (uint64_t)n_lanes * bins));
  chunk = std::min(chunk, hops);
  // Power stage: with at least a few hops per lane, the front half of the fused kernel (staging ring, one pass over the PCM, hop
  // segments as work items; OMB_SPECTRUM_RING_POWER=0 keeps the frame-per-CTA kernel, which re-reads every sample N / hop times)
  const char* rp_env = getenv("OMB_SPECTRUM_RING_POWER");
  const bool ring_power_env = !(rp_env && rp_env[0] == '0');
  const bool aligned16 = (reinterpret_cast<uintptr_t>(d_lanes) & 15u) == 0 && (lane_stride % 4) == 0 && (cfg.hop % 4) == 0;
  const bool ring_power = fast16k && fused16k && aligned16 && ring_power_env && hops >= 4;
  // (Running the power stage of chunk i + 1 on a side stream under the smoothing stage of chunk i was measured: nothing — the
  // power kernel's CTAs take every SM's registers, the two stages cannot co-reside; profiles/r02_notes.md.)
  OMB_TRY(d_power.reserve((size_t)((uint64_t)n_lanes * chunk * bins)));
  float* state = nullptr;
  if (cfg.averaging != OMB_AVG_NONE) {
    OMB_TRY(d_state.reserve((size_t)((uint64_t)n_lanes * bins)));
    OMB_CUDA_TRY(cudaMemsetAsync(d_state.ptr, 0, sizeof(float) * n_lanes * bins, s));
    state = d_state.ptr;
  }
  if (d_peak_bin) OMB_TRY(keys_begin(*this, hops * n_lanes, s));
  if (ring_power) {
    OMB_TRY(spectrum_fast_frame_means(*this, d_lanes, lane_stride, n_lanes, hops, s));
  } else if (fast16k) {  // hop-block sums for DC removal: once for the whole batch, chunks index into them
    OMB_TRY(spectrum_fast_block_sums(*this, d_lanes, lane_stride, n_lanes, hops, s));
    ext_bsum_blocks = hops - 1 + cfg.fft_size / cfg.hop;
  }
  int rc = OMB_OK;
  for (uint64_t h0 = 0; h0 < hops && rc >= 0; h0 += chunk) {
    const uint64_t n = std::min(chunk, hops - h0);
    ext_block_off = h0;
    if (ring_power) rc = launch_spectrum_fused_power(*this, d_lanes + h0 * cfg.hop, n_lanes, n, lane_stride, d_power.ptr, h0, hops, s);
    else rc = power_device(d_lanes + h0 * cfg.hop, n_lanes, n, lane_stride, d_power.ptr, s);
    if (rc >= 0) rc = smooth_launch(*this, d_power.ptr, n_lanes, n, state, d_weighted, d_raw, d_peak_bin ? d_keys.ptr : nullptr, true, hops, h0, s);
  }
  ext_bsum_blocks = 0;
  ext_block_off = 0;
  OMB_TRY(rc);
  if (d_peak_bin) OMB_TRY(keys_finish(*this, hops * n_lanes, d_peak_bin, s));
  return OMB_OK;
}

// Resolves the peak spec to the inclusive bin range the kernels filter on, with the reference's f32 comparisons:
// i in 1..bins-1 (exclusive) with (min_f..=max_f).contains(&frequency_bins[i]) (state.rs:106-107,321-324).
int SpectrumPlan::set_peak_spec(const omb_spectrum_peak_spec& spec) {
  if (spec.trace > 1u) return fail(OMB_ERR_INVALID, "peak spec: trace must be 0 (A-weighted) or 1 (raw)");
  peak_spec = spec;
  peak_lo = 1;
  peak_hi = 0;
  const size_t bins = h_freq.size();
  if (bins < 3) return OMB_OK;
  const float min_f = spec.min_hz;
  const float max_f = spec.max_hz > 0.0f ? spec.max_hz : std::fmax(h_freq[bins - 1], min_f * 1.02f);
  // frequency_bins is non-decreasing, so the members of the range are contiguous
  int lo = -1, hi = -1;
  for (size_t i = 1; i + 1 < bins; ++i)
    if (h_freq[i] >= min_f && h_freq[i] <= max_f) {
      if (lo < 0) lo = (int)i;
      hi = (int)i;
    }
  if (lo >= 0) { peak_lo = lo; peak_hi = hi; }
  return OMB_OK;
}

int SpectrumPlan::interpolate_peaks_device(const float* d_db, const int32_t* d_peak_bin, uint64_t rows, float* d_freq, float* d_level,
                                           cudaStream_t s) {
  if (!rows) return OMB_OK;
  if (!d_db || !d_peak_bin || !d_freq || !d_level) return fail(OMB_ERR_INVALID, "null argument");
  OMB_CUDA_TRY(cudaSetDevice(dev.device));
  const float bin_hz = h_freq.size() > 1 ? h_freq[1] - h_freq[0] : 0.0f;  // state.rs:330
  OMB_LAUNCH(k_peak_interpolate, dim3((unsigned)((rows + 255) / 256)), dim3(256), 0, s, d_db, d_peak_bin, rows, (uint32_t)cfg.bins(), bin_hz,
             d_freq, d_level);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

int SpectrumPlan::execute_host(const float* h_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride,
                               float* h_weighted, float* h_raw, int32_t* h_peak_bin, float* h_peak_freq, float* h_peak_level) {
  const uint64_t hops = cfg.hops_for(samples_per_lane);
  if (!hops || !n_lanes) return OMB_OK;
  if (!h_lanes || !h_weighted || !h_raw) return fail(OMB_ERR_INVALID, "null argument");
  OMB_CUDA_TRY(cudaSetDevice(dev.device));
  OMB_TRY(d_in.reserve((size_t)(samples_per_lane * n_lanes)));
  for (uint32_t l = 0; l < n_lanes; ++l)
    OMB_CUDA_TRY(cudaMemcpyAsync(d_in.ptr + l * samples_per_lane, h_lanes + l * lane_stride, sizeof(float) * samples_per_lane,
                                 cudaMemcpyHostToDevice, stream));
  const uint64_t n = hops * n_lanes * cfg.bins();
  OMB_TRY(d_w.reserve((size_t)n));
  OMB_TRY(d_r.reserve((size_t)n));
  const bool want_interp = h_peak_freq || h_peak_level;
  const bool want_bins = h_peak_bin || want_interp;
  if (want_bins) OMB_TRY(d_peak.reserve((size_t)(hops * n_lanes)));
  OMB_TRY(execute_device(d_in.ptr, n_lanes, samples_per_lane, samples_per_lane, d_w.ptr, d_r.ptr, want_bins ? d_peak.ptr : nullptr, stream));
  if (want_interp) {
    OMB_TRY(d_peak_freq.reserve((size_t)(hops * n_lanes)));
    OMB_TRY(d_peak_level.reserve((size_t)(hops * n_lanes)));
    OMB_TRY(interpolate_peaks_device(peak_spec.trace == 1u ? d_r.ptr : d_w.ptr, d_peak.ptr, hops * n_lanes, d_peak_freq.ptr, d_peak_level.ptr,
                                     stream));
    if (h_peak_freq)
      OMB_CUDA_TRY(cudaMemcpyAsync(h_peak_freq, d_peak_freq.ptr, sizeof(float) * hops * n_lanes, cudaMemcpyDeviceToHost, stream));
    if (h_peak_level)
      OMB_CUDA_TRY(cudaMemcpyAsync(h_peak_level, d_peak_level.ptr, sizeof(float) * hops * n_lanes, cudaMemcpyDeviceToHost, stream));
  }
  OMB_CUDA_TRY(cudaMemcpyAsync(h_weighted, d_w.ptr, sizeof(float) * n, cudaMemcpyDeviceToHost, stream));
  OMB_CUDA_TRY(cudaMemcpyAsync(h_raw, d_r.ptr, sizeof(float) * n, cudaMemcpyDeviceToHost, stream));
  if (h_peak_bin) OMB_CUDA_TRY(cudaMemcpyAsync(h_peak_bin, d_peak.ptr, sizeof(int32_t) * hops * n_lanes, cudaMemcpyDeviceToHost, stream));
  OMB_CUDA_TRY(cudaStreamSynchronize(stream));
  return OMB_OK;
}

}  // namespace omb
