// stft_plan.cu — batched STFT plan: table set-up, kernel selection, device/host execution.
#include "stft.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace omb {

StftConfig StftConfig::from_c(const omb_spectrogram_config& c) {
  StftConfig o;  // spectrogram/processor.rs:71-82 normalize()
  o.sample_rate = sanitize_sample_rate(c.sample_rate);
  o.window_kind = c.window <= OMB_WINDOW_BLACKMAN_HARRIS ? c.window : (uint32_t)OMB_WINDOW_RECTANGULAR;
  o.window = c.fft_size == 0 ? 2048 : c.fft_size;
  o.hop = c.hop_size == 0 ? std::max<uint64_t>(std::min<uint64_t>(64, o.window), 1) : c.hop_size;
  o.history_length = c.history_length;
  o.zero_pad = std::max<uint64_t>(c.zero_padding_factor, 1);
  o.reassign = c.use_reassignment != 0;
  return o;
}

void StftConfig::to_c(omb_spectrogram_config* out) const {
  std::memset(out, 0, sizeof *out);
  out->sample_rate = sample_rate;
  out->window = window_kind;
  out->fft_size = window;
  out->hop_size = hop;
  out->history_length = history_length;
  out->zero_padding_factor = zero_pad;
  out->use_reassignment = reassign ? 1 : 0;
}

StftPlan::~StftPlan() {
  if (stream) cudaStreamDestroy(stream);
  for (auto& ps : pipe)
    if (ps) cudaStreamDestroy(ps);
}

int StftPlan::init(const omb_spectrogram_config& c, int choice) {
  cfg = StftConfig::from_c(c);
  kernel_choice = choice;
  OMB_TRY(current_device(&dev));
  const uint64_t N = cfg.window, F = cfg.fft_len(), H = cfg.hilbert_len();
  if (!is_pow2(N) || !is_pow2(F))
    return fail(OMB_ERR_UNSUPPORTED,
                "fft_size %llu x zero_padding %llu: only power-of-two transform lengths have kernels (no CPU fallback)",
                (unsigned long long)N, (unsigned long long)cfg.zero_pad);
  if (F > (1ull << 24) || H > (1ull << 24)) return fail(OMB_ERR_UNSUPPORTED, "transform length %llu too large", (unsigned long long)F);

  h_win = make_window((int)cfg.window_kind, (size_t)N);
  h_norm = make_bin_norm(h_win.data(), (size_t)N, (size_t)F);
  power_scale = 1.0f;
  if (cfg.reassign) {
    const float inv_h = 1.0f / (float)H;  // spectrogram/processor.rs:263-266: IFFT left unnormalised
    for (auto& v : h_norm) v *= inv_h * inv_h;
    h_dwin = make_derivative_window(h_win.data(), (size_t)N);
    h_twin = make_time_weighted_window(h_win.data(), (size_t)N);
    power_scale = make_power_scale(h_win.data(), (size_t)N, (size_t)F);
  }
  OMB_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  OMB_TRY(d_win.upload(h_win, stream));
  OMB_TRY(d_norm.upload(h_norm, stream));
  if (cfg.reassign) {
    OMB_TRY(d_dwin.upload(h_dwin, stream));
    OMB_TRY(d_twin.upload(h_twin, stream));
    OMB_TRY(d_tw_hil.upload(make_twiddles((size_t)H, (size_t)std::max<uint64_t>(H / 2, 1)), stream));
  }
  OMB_TRY(d_tw_fft.upload(make_twiddles((size_t)F, (size_t)std::max<uint64_t>(F / 2, 1)), stream));

  fast = false;
  fast_kind = 0;
  // OMB_FAST_KERNEL=1|2|3 pins the specialised kernel generation (measurement / cross-checks); default: newest that fits
  const char* pin = getenv("OMB_FAST_KERNEL");
  const int want = pin ? atoi(pin) : 3;
  const bool gen1 = stft_fast_supported(cfg, dev), gen2 = want >= 2 && stft_fast2_supported(cfg, dev);
  // Generation 3 (stft_r64.cu) has no staging ring: any hop that is a multiple of 4.  Generation 2 with its contiguous work
  // assignment is 3-4 % faster at every hop its ring handles (multiples of 512 up to 2048, multiples of 4 below 512;
  // profiles/r02_notes.md), so unpinned generation 3 serves the hops generation 2 cannot.
  const bool gen3 = want >= 3 && stft_r64_supported(cfg, dev) && ((pin != nullptr && atoi(pin) >= 3) || !gen2);
  if (choice != OMB_KERNEL_GENERIC && gen3) {  // N = 4096, any hop % 4 == 0: two-pass radix-64 teams (stft_r64.cu)
    OMB_TRY(stft_r64_prepare(*this));
    fast = true;
    fast_kind = 7;
  } else if (choice != OMB_KERNEL_GENERIC && (gen1 || gen2)) {
    OMB_TRY(stft_fast_prepare(*this));  // uploads the twiddle tables both generations use
    fast = true;
    fast_kind = 1;
    if (gen2) {  // also the only specialised kernel for hops below 512
      OMB_TRY(stft_fast2_prepare(*this));
      fast_kind = 2;
    }
  } else if (choice != OMB_KERNEL_GENERIC && stft_r64x_supported(cfg, dev) && !getenv("OMB_NO_R64X")) {
    OMB_TRY(stft_r64x_prepare(*this));
    fast = true;
    fast_kind = 8;
  } else if (choice != OMB_KERNEL_GENERIC && stft_fast8k_supported(cfg, dev) && !getenv("OMB_NO_FAST8K")) {
    OMB_TRY(stft_fast8k_prepare(*this));
    fast = true;
    fast_kind = 4;
  } else if (choice != OMB_KERNEL_GENERIC && stft_fast2k_supported(cfg, dev) && !getenv("OMB_NO_FAST2K")) {
    OMB_TRY(stft_fast2k_prepare(*this));
    fast = true;
    fast_kind = 5;
  } else if (choice != OMB_KERNEL_GENERIC && stft_fast1k_supported(cfg, dev) && !getenv("OMB_NO_FAST1K")) {
    OMB_TRY(stft_fast1k_prepare(*this));
    fast = true;
    fast_kind = 6;
  } else if (choice != OMB_KERNEL_GENERIC && stft_classic_fast_supported(cfg, dev)) {
    OMB_TRY(stft_classic_fast_prepare(*this));
    fast = true;
    fast_kind = 3;
  } else if (choice == OMB_KERNEL_FAST) {
    return fail(OMB_ERR_UNSUPPORTED, "no specialised kernel for window %llu hop %llu zp %llu reassign %d",
                (unsigned long long)N, (unsigned long long)cfg.hop, (unsigned long long)cfg.zero_pad, (int)cfg.reassign);
  }
  // The shared-memory tier is prepared whenever it applies: it is the kernel of choice when no specialised kernel
  // exists, and the fallback of the specialised kernels when a caller's lanes are not 16-byte aligned (e.g. the
  // streaming FIFO after an odd-sized drain).
  if (choice != OMB_KERNEL_GENERIC && stft_smem_supported(cfg, dev) && !getenv("OMB_NO_SMEM_KERNEL")) {
    OMB_TRY(stft_smem_prepare(*this));
    smem_kernel = true;
  }
  OMB_CUDA_TRY(cudaStreamSynchronize(stream));
  return OMB_OK;
}

int StftPlan::execute_device(const float* d_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride,
                             omb_spectrogram_point* out_points, uint64_t point_stride, uint32_t* out_counts,
                             uint16_t* out_classic, cudaStream_t s, uint64_t first_frame) {
  const uint64_t frames = cfg.frames_for(samples_per_lane);
  if (frames == 0 || n_lanes == 0 || first_frame >= frames) return OMB_OK;
  if (!d_lanes) return fail(OMB_ERR_INVALID, "null lanes pointer");
  OMB_CUDA_TRY(cudaSetDevice(dev.device));  // plans are bound to the device they were created on
  if (cfg.reassign) {
    if (!out_points || !out_counts) return fail(OMB_ERR_INVALID, "reassigned plan needs out_points and out_counts");
    if (point_stride < cfg.bins()) return fail(OMB_ERR_INVALID, "point_stride %llu < bins %llu", (unsigned long long)point_stride, (unsigned long long)cfg.bins());
  } else if (!out_classic) {
    return fail(OMB_ERR_INVALID, "classic plan needs out_classic");
  }
  StftKernelArgs a{};
  a.lanes = d_lanes;
  a.lane_stride = lane_stride;
  a.n_lanes = n_lanes;
  a.frames_per_lane = frames;
  a.first_frame = first_frame;
  a.window = (uint32_t)cfg.window;
  a.fft_len = (uint32_t)cfg.fft_len();
  a.hilbert_len = (uint32_t)cfg.hilbert_len();
  a.hop = (uint32_t)cfg.hop;
  a.bins = (uint32_t)cfg.bins();
  a.log2_fft = (uint32_t)ilog2(a.fft_len);
  a.log2_hilbert = (uint32_t)ilog2(a.hilbert_len);
  a.win = d_win.ptr;
  a.dwin = d_dwin.ptr;
  a.twin = d_twin.ptr;
  a.bin_norm = d_norm.ptr;
  a.tw_fft = d_tw_fft.ptr;
  a.tw_hil = d_tw_hil.ptr;
  // spectrogram/processor.rs:446-450 — all f32
  const float sr = cfg.sample_rate;
  a.bin_hz = sr / (float)a.fft_len;
  a.max_hz = sr * 0.5f;
  a.inv_2pi = sr / 6.28318530717958647692f;
  a.inv_hop = 1.0f / (float)cfg.hop;
  a.latency_hops = (float)((a.hilbert_len - a.window) / 2) * a.inv_hop;
  a.out_points = out_points;
  a.point_stride = point_stride;
  a.out_counts = out_counts;
  a.out_classic = out_classic;
  const bool aligned16 = (reinterpret_cast<uintptr_t>(d_lanes) & 15u) == 0 && (lane_stride % 4) == 0;
  const bool aligned8 = (reinterpret_cast<uintptr_t>(d_lanes) & 7u) == 0 && (lane_stride % 2) == 0;
  if (fast_kind == 3 && aligned8) return launch_stft_classic_fast(*this, a, s);
  if (fast_kind == 7 && aligned16) return launch_stft_r64(*this, a, s);
  if (fast_kind == 8 && aligned16) return launch_stft_r64x(*this, a, s);
  if (fast_kind == 4 && aligned16) return launch_stft_fast8k(*this, a, s);
  if (fast_kind == 5 && aligned16) return launch_stft_fast2k(*this, a, s);
  if (fast_kind == 6 && aligned16) return launch_stft_fast1k(*this, a, s);
  if (fast && fast_kind <= 2 && aligned16) return fast_kind == 2 ? launch_stft_fast2(*this, a, s) : launch_stft_fast(*this, a, s);
  if (fast && kernel_choice == OMB_KERNEL_FAST)
    return fail(OMB_ERR_INVALID, "OMB_KERNEL_FAST was forced but the lanes are not aligned (reassigned: 16 bytes and lane_stride % 4 == 0; classic: 8 bytes and lane_stride % 2 == 0)");
  if (smem_kernel) return launch_stft_smem(*this, a, s, d_scratch);
  return launch_stft_generic(*this, a, s, d_scratch);
}

int StftPlan::execute_host(const float* h_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride,
                           omb_spectrogram_point* h_points, uint64_t point_stride, uint32_t* h_counts,
                           uint16_t* h_classic) {
  const uint64_t frames = cfg.frames_for(samples_per_lane);
  if (frames == 0 || n_lanes == 0) return OMB_OK;
  if (!h_lanes) return fail(OMB_ERR_INVALID, "null lanes pointer");
  // Device lanes are laid out with a stride rounded up to 4 samples so that every lane is 16-byte aligned for the
  // specialised kernels; one H2D copy per lane when that differs from the caller's stride.
  const uint64_t ds = (samples_per_lane + 3) & ~(uint64_t)3;
  OMB_CUDA_TRY(cudaSetDevice(dev.device));
  OMB_TRY(d_in.reserve((size_t)(ds * n_lanes)));
  auto upload = [&](uint32_t l0, uint32_t nl, cudaStream_t st) -> int {
    if (lane_stride == ds) {
      OMB_CUDA_TRY(cudaMemcpyAsync(d_in.ptr + (uint64_t)l0 * ds, h_lanes + (uint64_t)l0 * lane_stride,
                                   sizeof(float) * (ds * (nl - 1) + samples_per_lane), cudaMemcpyHostToDevice, st));
    } else {
      for (uint32_t l = l0; l < l0 + nl; ++l)
        OMB_CUDA_TRY(cudaMemcpyAsync(d_in.ptr + (uint64_t)l * ds, h_lanes + (uint64_t)l * lane_stride, sizeof(float) * samples_per_lane,
                                     cudaMemcpyHostToDevice, st));
    }
    return OMB_OK;
  };
  const bool pipelined = cfg.reassign && fast_kind > 0 && n_lanes >= 4;
  if (!pipelined) OMB_TRY(upload(0, n_lanes, stream));
  const uint64_t slots = frames * n_lanes;
  if (cfg.reassign) {
    if (!h_points || !h_counts) return fail(OMB_ERR_INVALID, "reassigned plan needs out_points and out_counts");
    OMB_TRY(d_points.reserve((size_t)(slots * point_stride)));
    OMB_TRY(d_counts.reserve((size_t)slots));
    if (pipelined) {
      // Specialised kernels need no scratch, so lane chunks can be pipelined over three streams: the H2D of chunk
      // i+1, the kernel of chunk i and the D2H of chunk i-1 overlap (PCIe is full duplex; D2H of 12-byte points
      // dominates: ~24.6 KB per frame).
      OMB_CUDA_TRY(cudaStreamSynchronize(stream));
      for (auto& ps : pipe)
        if (!ps) OMB_CUDA_TRY(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
      const uint32_t chunk = std::max<uint32_t>(1, (n_lanes + 11) / 12);
      int k = 0;
      for (uint32_t l0 = 0; l0 < n_lanes; l0 += chunk, ++k) {
        const uint32_t nl = std::min(chunk, n_lanes - l0);
        cudaStream_t ps = pipe[k % 3];
        OMB_TRY(upload(l0, nl, ps));
        const uint64_t s0 = (uint64_t)l0 * frames;
        OMB_TRY(execute_device(d_in.ptr + (uint64_t)l0 * ds, nl, samples_per_lane, ds,
                               d_points.ptr + s0 * point_stride, point_stride, d_counts.ptr + s0, nullptr, ps));
        OMB_CUDA_TRY(cudaMemcpyAsync(h_counts + s0, d_counts.ptr + s0, sizeof(uint32_t) * frames * nl, cudaMemcpyDeviceToHost, ps));
        OMB_CUDA_TRY(cudaMemcpyAsync(h_points + s0 * point_stride, d_points.ptr + s0 * point_stride,
                                     sizeof(omb_spectrogram_point) * frames * nl * point_stride, cudaMemcpyDeviceToHost, ps));
      }
      for (auto& ps : pipe) OMB_CUDA_TRY(cudaStreamSynchronize(ps));
      return OMB_OK;
    }
    OMB_TRY(execute_device(d_in.ptr, n_lanes, samples_per_lane, ds, d_points.ptr, point_stride, d_counts.ptr,
                           nullptr, stream));
    OMB_CUDA_TRY(cudaMemcpyAsync(h_counts, d_counts.ptr, sizeof(uint32_t) * slots, cudaMemcpyDeviceToHost, stream));
    OMB_CUDA_TRY(cudaMemcpyAsync(h_points, d_points.ptr, sizeof(omb_spectrogram_point) * slots * point_stride,
                                 cudaMemcpyDeviceToHost, stream));
  } else {
    if (!h_classic) return fail(OMB_ERR_INVALID, "classic plan needs out_classic");
    OMB_TRY(d_classic.reserve((size_t)(slots * cfg.bins())));
    OMB_TRY(execute_device(d_in.ptr, n_lanes, samples_per_lane, ds, nullptr, 0, nullptr, d_classic.ptr, stream));
    OMB_CUDA_TRY(cudaMemcpyAsync(h_classic, d_classic.ptr, sizeof(uint16_t) * slots * cfg.bins(), cudaMemcpyDeviceToHost, stream));
  }
  OMB_CUDA_TRY(cudaStreamSynchronize(stream));
  return OMB_OK;
}

}  // namespace omb
