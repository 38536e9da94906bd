// stft_generic.cu — any-power-of-two STFT kernels (classic + reassigned).
//
// Correctness-first path used for sizes that have no specialised sm_100a kernel
// (and as an on-GPU cross-check of the specialised ones).  One CTA owns one frame
// at a time and works in a private global-memory scratch area (L1/L2 resident),
// so there is no shared-memory size limit: N up to 16384 with zero-padding x32
// still runs.  Algorithm rows: SURVEY.md §8 a4, a7-a10 / §9.
#include "device_math.cuh"
#include "stft.h"

namespace omb {

namespace {

constexpr int kGenericThreads = 256;

// In-place radix-2 DIT FFT of `n` = 2^logn complex points living in global scratch.
// tw[k] = W_n^k (k < n/2); inverse uses conj(tw). Unnormalised. Block-cooperative.
__device__ void block_fft_radix2(float2* data, int n, int logn, const float2* __restrict__ tw, bool inverse) {
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
    const int half = 1 << (s - 1);
    const int step = n >> s;
    for (int i = tid; i < (n >> 1); i += nt) {
      const int k = i & (half - 1);
      const int base = ((i >> (s - 1)) << s) + k;
      float2 w = __ldg(&tw[k * step]);
      if (inverse) w.y = -w.y;
      const float2 a = data[base], b = data[base + half];
      const float2 t = cmul(b, w);
      data[base] = cadd(a, t);
      data[base + half] = csub(a, t);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kGenericThreads) k_classic_generic(StftKernelArgs a) {
  __shared__ float red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  float2* work = a.scratch + (uint64_t)blockIdx.x * a.scratch_stride;
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total = per_lane * a.n_lanes;
  const int N = (int)a.window, F = (int)a.fft_len;
  for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const uint64_t lane = item / per_lane, frame = a.first_frame + item % per_lane;
    const float* x = a.lanes + lane * a.lane_stride + frame * a.hop;
    // a4: DC removal (mean) + window, zero-pad to F
    float part = 0.0f;
    for (int i = tid; i < N; i += nt) part += __ldg(&x[i]);
    const float mean = block_sum(part, red) / (float)N;
    for (int i = tid; i < F; i += nt) {
      const float v = i < N ? (__ldg(&x[i]) - mean) * __ldg(&a.win[i]) : 0.0f;
      work[i] = make_float2(v, 0.0f);
    }
    __syncthreads();
    block_fft_radix2(work, F, (int)a.log2_fft, a.tw_fft, false);
    // a7: power * norm -> dB -> u16
    uint16_t* out = a.out_classic + (lane * a.frames_per_lane + frame) * a.bins;
    for (int k = tid; k < (int)a.bins; k += nt) {
      const float2 z = work[k];
      const float p = (z.x * z.x + z.y * z.y) * __ldg(&a.bin_norm[k]);
      out[k] = pack_classic_db_dev(power_to_db_dev(p, kDbFloor));
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kGenericThreads) k_reassigned_generic(StftKernelArgs a) {
  __shared__ int cnt[33];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = (int)a.window, F = (int)a.fft_len, H = (int)a.hilbert_len;
  float2* A = a.scratch + (uint64_t)blockIdx.x * a.scratch_stride;  // H
  float2* S = A + H;                                                   // F
  float2* D = S + F;                                                   // F
  float2* T = D + F;                                                   // F
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total = per_lane * a.n_lanes;
  const int off = (H - N) / 2;
  ReassignConsts rc{a.bin_hz, a.max_hz, a.inv_2pi, a.inv_hop, a.latency_hops};
  for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const uint64_t lane = item / per_lane, frame = a.first_frame + item % per_lane;
    const float* x = a.lanes + lane * a.lane_stride + frame * a.hop;
    // a8: analytic signal via FFT_H -> zero DC and negative frequencies (Nyquist kept, no doubling) -> IFFT_H
    for (int i = tid; i < H; i += nt) A[i] = make_float2(__ldg(&x[i]), 0.0f);
    __syncthreads();
    block_fft_radix2(A, H, (int)a.log2_hilbert, a.tw_hil, false);
    for (int i = tid; i < H; i += nt)
      if (i == 0 || i > H / 2) A[i] = make_float2(0.0f, 0.0f);
    __syncthreads();
    block_fft_radix2(A, H, (int)a.log2_hilbert, a.tw_hil, true);
    // a9: three windowed copies of the centre N samples, zero-padded to F
    for (int i = tid; i < F; i += nt) {
      float2 s = make_float2(0.f, 0.f), d = s, t = s;
      if (i < N) {
        const float2 c = A[off + i];
        s = cscale(c, __ldg(&a.win[i]));
        d = cscale(c, __ldg(&a.dwin[i]));
        t = cscale(c, __ldg(&a.twin[i]));
      }
      S[i] = s;
      D[i] = d;
      T[i] = t;
    }
    __syncthreads();
    block_fft_radix2(S, F, (int)a.log2_fft, a.tw_fft, false);
    block_fft_radix2(D, F, (int)a.log2_fft, a.tw_fft, false);
    block_fft_radix2(T, F, (int)a.log2_fft, a.tw_fft, false);
    // a10: per-bin reassignment + order-preserving compaction (ascending bin)
    const uint64_t slot = lane * a.frames_per_lane + frame;
    omb_spectrogram_point* out = a.out_points + slot * a.point_stride;
    int base = 0;
    for (int k0 = 0; k0 < (int)a.bins; k0 += nt) {
      const int k = k0 + tid;
      omb_spectrogram_point p;
      bool keep = false;
      if (k < (int)a.bins) keep = reassign_bin(S[k], D[k], T[k], __ldg(&a.bin_norm[k]), k, rc, &p);
      int tot;
      const int rank = block_rank(keep, cnt, &tot);
      if (keep) out[base + rank] = p;
      base += tot;
    }
    if (tid == 0) a.out_counts[slot] = (uint32_t)base;
    __syncthreads();
  }
}

}  // namespace

int launch_stft_generic(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s, DeviceBuffer<float2>& scratch) {
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total = per_lane * a.n_lanes;
  if (total == 0) return OMB_OK;
  const uint64_t max_blocks = (uint64_t)std::max(plan.dev.sm_count, 1) * 4;
  const unsigned grid = (unsigned)std::min<uint64_t>(total, max_blocks);
  a.scratch_stride = plan.cfg.reassign ? (uint64_t)a.hilbert_len + 3ull * a.fft_len : (uint64_t)a.fft_len;
  OMB_TRY(scratch.reserve((size_t)(a.scratch_stride * grid)));
  a.scratch = scratch.ptr;
  if (plan.cfg.reassign) {
    OMB_LAUNCH(k_reassigned_generic, dim3(grid), dim3(kGenericThreads), 0, s, a);
  } else {
    OMB_LAUNCH(k_classic_generic, dim3(grid), dim3(kGenericThreads), 0, s, a);
  }
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
