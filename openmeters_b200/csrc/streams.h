// streams.h — the three streaming processors behind omb_spectrogram / omb_spectrum / omb_loudness.
// Host logic mirrors the reference's process_block state machines; all arithmetic is in kernels.
#pragma once
#include <memory>

#include "downmix.h"
#include "loudness.h"
#include "spectrum.h"
#include "stft.h"

namespace omb {

struct SpectrogramStream {
  StftConfig config;
  bool prepared = false;
  bool reset = true;
  uint64_t pending_skip = 0;
  DeviceLane pending;
  std::unique_ptr<StftPlan> plan;
  omb_spectrogram_config plan_cfg{};
  DeviceInfo dev;
  cudaStream_t stream = nullptr;
  DeviceBuffer<float> d_block;
  DeviceBuffer<omb_spectrogram_point> d_points;
  DeviceBuffer<uint32_t> d_counts;
  DeviceBuffer<uint16_t> d_classic;
  PinnedBuffer<omb_spectrogram_point> h_points;
  PinnedBuffer<uint32_t> h_counts;
  std::vector<uint32_t> offsets;
  std::vector<omb_spectrogram_point> points;
  std::vector<uint16_t> classic;

  explicit SpectrogramStream(const omb_spectrogram_config& c);
  ~SpectrogramStream();
  int ensure_stream();
  int sync_plan();
  int rebuild_fft();
  int prepare();
  void reset_audio();
  void advance_audio(uint64_t count);
  int update_config(const omb_spectrogram_config& c);
  int push_audio(const float* samples, size_t n_samples, uint32_t channels, const uint8_t* positions);
  int process_block(const float* samples, size_t n_samples, uint32_t channels, float sample_rate, const uint8_t* positions,
                    omb_spectrogram_update* out);
};

struct SpectrumStream {
  SpectrumConfigN config;
  bool prepared = false;
  uint64_t pending_skip = 0;
  DeviceLane pcm[2];
  std::unique_ptr<SpectrumPlan> plan;
  omb_spectrum_config plan_cfg{};
  DeviceInfo dev;
  cudaStream_t stream = nullptr;
  DeviceBuffer<float> d_block, d_power, d_state[2], d_out[2][2];
  std::vector<float> h_traces[2][2];
  std::vector<float> h_freq;

  explicit SpectrumStream(const omb_spectrum_config& c);
  ~SpectrumStream();
  int ensure_stream();
  int sync_plan();
  void active_traces(bool a[2]) const;
  int reset_level_buffers();
  int reset_buffers();
  int rebuild_fft();
  int prepare();
  int reset_audio();
  int update_config(const omb_spectrum_config& c);
  int process_block(const float* samples, size_t n_samples, uint32_t channels, float sample_rate, const uint8_t* positions,
                    omb_spectrum_snapshot* out);
  void fill(omb_spectrum_snapshot* out);
};

struct LoudnessStream {
  omb_loudness_config cfg;
  float sample_rate;          // sanitised rate the state was built for
  uint32_t channels = 0;      // 0 until the first block
  KWeight kw;
  TruePeakFir fir;
  uint64_t caps[kLoudWindows];
  uint64_t ring_len = 1;
  uint32_t tp_delay_len = 0;
  LoudnessStreamCore core;
  DeviceInfo dev;
  cudaStream_t stream = nullptr;

  uint32_t n_streams = 1;     // > 1: a bank of lock-step streams sharing config, channel layout and rate (omb_loudness_bank)

  explicit LoudnessStream(const omb_loudness_config& c);
  ~LoudnessStream();
  int ensure_stream();
  void configure_rate(float sr);
  int ensure_state(uint32_t requested_channels, float sr);
  int reset_audio();
  int process_block(const float* samples, size_t n_samples, uint32_t channels, float sample_rate, const uint8_t* positions,
                    omb_loudness_snapshot* out);
  // n_streams x process_block in one launch: stream s reads samples + s * stream_stride, out[s] receives its snapshot
  int process_bank(const float* samples, uint64_t stream_stride, size_t n_samples, uint32_t channels, float sample_rate,
                   const uint8_t* positions, omb_loudness_snapshot* out);
};

}  // namespace omb
