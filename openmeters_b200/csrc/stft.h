// stft.h — batched STFT plan (classic + reassigned), shared by the generic and the specialised kernels.
#pragma once
#include "common.h"
#include "tables.h"

namespace omb {

struct StftConfig {  // normalised copy of omb_spectrogram_config (spectrogram/processor.rs:71-82)
  float sample_rate = kDefaultSampleRate;
  uint32_t window_kind = OMB_WINDOW_HANN;
  uint64_t window = 2048;  // "fft_size" of the reference config = window length N
  uint64_t hop = 64;
  uint64_t history_length = 0;
  uint64_t zero_pad = 1;
  bool reassign = true;
  static StftConfig from_c(const omb_spectrogram_config& c);
  void to_c(omb_spectrogram_config* out) const;
  uint64_t fft_len() const { return window * zero_pad; }                 // F
  uint64_t hilbert_len() const { return hilbert_len_for((size_t)window); }  // H
  uint64_t read_len() const { return reassign ? hilbert_len() : window; }
  uint64_t bins() const { return fft_len() / 2 + 1; }
  uint64_t frames_for(uint64_t samples) const { return samples >= read_len() ? (samples - read_len()) / hop + 1 : 0; }
};

// Everything the kernels read. All pointers are device pointers.
struct StftKernelArgs {
  const float* lanes;
  uint64_t lane_stride;
  uint32_t n_lanes;
  uint64_t frames_per_lane;
  uint64_t first_frame;  // frames [first_frame, frames_per_lane) of every lane are computed
  uint32_t window;       // N
  uint32_t fft_len;      // F = N * zp
  uint32_t hilbert_len;  // H
  uint32_t hop;
  uint32_t bins;
  uint32_t log2_fft, log2_hilbert;
  const float* win;      // h[N]
  const float* dwin;     // dh[N]
  const float* twin;     // t*h[N]
  const float* bin_norm; // [bins] (already / H^2 for reassigned)
  const float2* tw_fft;  // W_F^k, k < F/2
  const float2* tw_hil;  // W_H^k, k < H/2
  float bin_hz, max_hz, inv_2pi, inv_hop, latency_hops;
  // outputs
  omb_spectrogram_point* out_points;
  uint64_t point_stride;
  uint32_t* out_counts;
  uint16_t* out_classic;
  // generic-kernel scratch: per-block complex work area
  float2* scratch;
  uint64_t scratch_stride;  // float2 elements per block
};

struct StftPlan {
  StftConfig cfg;
  DeviceInfo dev;
  int kernel_choice = OMB_KERNEL_AUTO;
  bool fast = false;
  int fast_kind = 0;  // 0 generic, 1 = stft_fast.cu, 2 = stft_fast2.cu, 3 = stft_classic_fast.cu, 4 = stft_fast8k.cu, 5 = stft_fast2k.cu, 6 = stft_fast1k.cu, 7 = stft_r64.cu, 8 = stft_r64x.cu
  bool smem_kernel = false;  // stft_smem.cu (used when no specialised kernel applies and OMB_KERNEL_GENERIC was not forced)
  float power_scale = 1.0f;
  std::vector<float> h_win, h_dwin, h_twin, h_norm;
  DeviceBuffer<float> d_win, d_dwin, d_twin, d_norm;
  DeviceBuffer<float2> d_tw_fft, d_tw_hil, d_scratch;
  DeviceBuffer<float2> d_fast_tables;  // specialised-kernel twiddle tables
  DeviceBuffer<float2> d_r64_tables;   // stft_r64.cu: W_4096^{t q} rows
  DeviceBuffer<float> d_r64_scratch;   // stft_r64.cu: global register park (only with OMB_R64_PARK=global)
  bool r64 = false;                    // stft_r64.cu prepared (N = 4096 reassigned, any hop % 4 == 0)
  // host-path staging
  DeviceBuffer<float> d_in;
  DeviceBuffer<omb_spectrogram_point> d_points;
  DeviceBuffer<uint32_t> d_counts;
  DeviceBuffer<uint16_t> d_classic;
  DeviceBuffer<float> d_img_accum, d_img_db;  // render_host: accumulation / dB images of the lane chunks in flight
  cudaStream_t stream = nullptr;  // owned, for the host path
  cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};  // host path: H2D / kernel / D2H overlap across lane chunks

  ~StftPlan();
  int init(const omb_spectrogram_config& c, int kernel_choice);
  // Computes frames [first_frame, frames_per_lane) of each lane. Outputs are indexed by absolute frame.
  int execute_device(const float* d_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride,
                     omb_spectrogram_point* d_points, uint64_t point_stride, uint32_t* d_counts, uint16_t* d_classic,
                     cudaStream_t s, uint64_t first_frame = 0);
  int execute_host(const float* h_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride,
                   omb_spectrogram_point* h_points, uint64_t point_stride, uint32_t* h_counts, uint16_t* h_classic);
  // stft_render.cu: STFT -> splat accumulate -> resolve on the device; only the dB images (and optionally counts) come back
  int render_host(const float* h_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride,
                  const omb_splat_params& view, float* h_db, uint32_t* h_counts);
};

// kernel launchers (stft_generic.cu / stft_fast.cu)
int launch_stft_generic(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s, DeviceBuffer<float2>& scratch);
bool stft_fast_supported(const StftConfig& cfg, const DeviceInfo& dev);
int stft_fast_prepare(StftPlan& plan);
int launch_stft_fast(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s);
bool stft_fast2_supported(const StftConfig& cfg, const DeviceInfo& dev);
int stft_fast2_prepare(StftPlan& plan);
int launch_stft_fast2(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s);
// stft_r64.cu: reassigned N = 4096, two-pass radix-64 transforms, one 64-thread team per frame (third generation)
bool stft_r64_supported(const StftConfig& cfg, const DeviceInfo& dev);
int stft_r64_prepare(StftPlan& plan);
int launch_stft_r64(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s);
// stft_r64x.cu: reassigned N = 16384 on chip (one 256-thread CTA per frame, 64 x 64 x 4 transforms, TMEM park)
bool stft_r64x_supported(const StftConfig& cfg, const DeviceInfo& dev);
int stft_r64x_prepare(StftPlan& plan);
int launch_stft_r64x(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s);
// stft_fast2k.cu: reassigned N = 2048 (two interleaved frames per 4096-point transform)
bool stft_fast2k_supported(const StftConfig& cfg, const DeviceInfo& dev);
int stft_fast2k_prepare(StftPlan& plan);
int launch_stft_fast2k(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s);
// stft_fast1k.cu: reassigned N = 1024 (four interleaved frames per 4096-point transform)
bool stft_fast1k_supported(const StftConfig& cfg, const DeviceInfo& dev);
int stft_fast1k_prepare(StftPlan& plan);
int launch_stft_fast1k(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s);
// stft_fast8k.cu: reassigned N = 8192 (two 4096-point sub-transforms per CTA)
bool stft_fast8k_supported(const StftConfig& cfg, const DeviceInfo& dev);
int stft_fast8k_prepare(StftPlan& plan);
int launch_stft_fast8k(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s);
// stft_classic_fast.cu: warp-per-frame classic column, N = 1024
bool stft_classic_fast_supported(const StftConfig& cfg, const DeviceInfo& dev);
int stft_classic_fast_prepare(StftPlan& plan);
int launch_stft_classic_fast(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s);
// stft_smem.cu: shared-memory Stockham kernels for any power-of-two size that fits on chip
bool stft_smem_supported(const StftConfig& cfg, const DeviceInfo& dev);
int stft_smem_prepare(StftPlan& plan);
int launch_stft_smem(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s, DeviceBuffer<float2>& scratch);

}  // namespace omb
