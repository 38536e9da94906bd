// spectrum.h — batched spectrum-analyzer plan.
#pragma once
#include "common.h"
#include "tables.h"

namespace omb {

struct SpectrumConfigN {  // normalised omb_spectrum_config (spectrum/processor.rs:53-62)
  float sample_rate = kDefaultSampleRate;
  uint32_t window_kind = OMB_WINDOW_HANN;
  uint64_t fft_size = 16384;
  uint64_t hop = 1024;
  uint32_t averaging = OMB_AVG_NONE;
  float averaging_param = 0.0f;
  uint32_t source = OMB_CHANNEL_MID, secondary = OMB_CHANNEL_NONE;
  float floor_db = -100.0f;
  static SpectrumConfigN from_c(const omb_spectrum_config& c);
  void to_c(omb_spectrum_config* out) const;
  uint64_t bins() const { return fft_size / 2 + 1; }
  uint64_t hops_for(uint64_t samples) const { return samples >= fft_size ? (samples - fft_size) / hop + 1 : 0; }
};

struct SpectrumPowerArgs {
  const float* lanes;
  uint64_t lane_stride;
  uint32_t n_lanes;
  uint64_t hops;       // hops per lane in this launch
  uint32_t fft_size, hop, bins, log2_fft;
  const float* win;
  const float* bin_norm;
  const float2* tw;    // W_N^k, k < N/2
  float* power;        // [(lane*hops + h)*bins + k]
  float2* scratch;
  uint64_t scratch_stride;
};

struct SpectrumSmoothArgs {
  const float* power;  // [(lane*hops + h)*bins + k]
  uint32_t n_lanes;
  uint64_t hops;
  uint32_t bins;
  int mode;
  float alpha;         // exponential: clamp(factor, 0, 0.9999)
  float decay;         // peak hold: db_to_power(-max(rate,0) * hop/sr)
  float state_floor, floor_db;
  const float* a_db;   // [bins]
  float* state;        // [lane*bins] smoothed power, in/out (may be null for mode None)
  float* out_weighted; // write_all: [(lane*hops+h)*bins+k], else [lane*bins+k] (last hop only)
  float* out_raw;
  unsigned long long* peak_keys;  // [(lane*hops+h)] packed (ordered dB of the peak trace, bin) or null
  int write_all;
  int peak_raw;        // peak spec: 1 = raw trace, 0 = weighted trace
  int peak_lo, peak_hi;  // peak spec: candidate bins peak_lo..peak_hi (inclusive; empty if lo > hi)
};

struct SpectrumPlan {
  SpectrumConfigN cfg;
  DeviceInfo dev;
  std::vector<float> h_win, h_norm, h_freq, h_adb;
  float state_floor = 0.0f;
  DeviceBuffer<float> d_win, d_norm, d_adb;
  DeviceBuffer<float2> d_tw, d_scratch;
  DeviceBuffer<float> d_power, d_state;
  DeviceBuffer<unsigned long long> d_keys;
  bool fast16k = false;                 // spectrum_fast.cu applies (N = 16384)
  DeviceBuffer<float2> d_fast_tables;
  DeviceBuffer<double> d_bsum;
  bool fused16k = false;                // whole-batch fused kernel available (spectrum_fast.cu, k_spectrum_fused_16k)
  DeviceBuffer<float> d_means;
  uint64_t ext_bsum_blocks = 0;         // > 0: d_bsum already holds [lane][ext_bsum_blocks] sums for the whole batch
  uint64_t ext_block_off = 0;           //      and this launch starts at that block
  // host-path staging
  DeviceBuffer<float> d_in, d_w, d_r;
  DeviceBuffer<int32_t> d_peak;
  DeviceBuffer<float> d_peak_freq, d_peak_level;
  cudaStream_t stream = nullptr;
  // row f3: peak label spec (spectrum/state.rs:106-107,134-136,321-325) resolved to a bin range on the host
  omb_spectrum_peak_spec peak_spec{0u, 20.0f, 0.0f};
  int peak_lo = 1, peak_hi = 0;
  int set_peak_spec(const omb_spectrum_peak_spec& spec);
  int interpolate_peaks_device(const float* d_db, const int32_t* d_peak_bin, uint64_t rows, float* d_freq, float* d_level, cudaStream_t s);

  ~SpectrumPlan();
  int init(const omb_spectrum_config& c);
  // Power spectra of `hops` consecutive hops of each lane, starting at each lane's sample 0.
  int power_device(const float* d_lanes, uint32_t n_lanes, uint64_t hops, uint64_t lane_stride, float* d_power_out, cudaStream_t s);
  // Smoothing + dB. state: [n_lanes*bins] (in/out) or null => zero-initialised internal state.
  int smooth_device(const float* d_power_in, uint32_t n_lanes, uint64_t hops, float* d_state, float* d_weighted, float* d_raw,
                    int32_t* d_peak_bin, bool write_all, cudaStream_t s);
  int execute_device(const float* d_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride, float* d_weighted,
                     float* d_raw, int32_t* d_peak_bin, cudaStream_t s);
  int execute_host(const float* h_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride, float* h_weighted,
                   float* h_raw, int32_t* h_peak_bin, float* h_peak_freq = nullptr, float* h_peak_level = nullptr);
};

// spectrum_fast.cu
bool spectrum_fast_supported(const SpectrumConfigN& cfg, const DeviceInfo& dev);
int spectrum_fast_prepare(SpectrumPlan& p);
int launch_spectrum_power_fast(SpectrumPlan& p, SpectrumPowerArgs& a, cudaStream_t s);
int launch_spectrum_fused(SpectrumPlan& p, const float* d_lanes, uint32_t n_lanes, uint64_t hops, uint64_t lane_stride, float* d_weighted,
                          float* d_raw, int32_t* d_peak_bin, cudaStream_t s);
int launch_spectrum_fused_power(SpectrumPlan& p, const float* d_lanes, uint32_t n_lanes, uint64_t hops, uint64_t lane_stride, float* d_power,
                                uint64_t h0, uint64_t hops_total, cudaStream_t s);
int spectrum_fast_frame_means(SpectrumPlan& p, const float* d_lanes, uint64_t lane_stride, uint32_t n_lanes, uint64_t hops, cudaStream_t s);
int spectrum_fast_block_sums(SpectrumPlan& p, const float* d_lanes, uint64_t lane_stride, uint32_t n_lanes, uint64_t hops, cudaStream_t s);

}  // namespace omb
