// stream_spectrum.cu — streaming SpectrumProcessor (spectrum/processor.rs:88-323) over the spectrum plan.
// Host logic only: source projection bookkeeping, hop/skip accounting, reset rules; arithmetic is in
// downmix.cu / spectrum.cu kernels. Smoothing state stays resident on the device between calls.
#include "streams.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace omb {

SpectrumStream::SpectrumStream(const omb_spectrum_config& c) { config = SpectrumConfigN::from_c(c); }

SpectrumStream::~SpectrumStream() {
  if (stream) cudaStreamDestroy(stream);
}

int SpectrumStream::ensure_stream() {
  if (!stream) {
    OMB_TRY(current_device(&dev));
    OMB_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  }
  return OMB_OK;
}

int SpectrumStream::sync_plan() {
  omb_spectrum_config c;
  config.to_c(&c);
  if (plan && std::memcmp(&c, &plan_cfg, sizeof c) == 0) return OMB_OK;
  plan.reset(new SpectrumPlan());
  const int rc = plan->init(c);
  if (rc < 0) {
    plan.reset();
    return rc;
  }
  plan_cfg = c;
  return OMB_OK;
}

void SpectrumStream::active_traces(bool a[2]) const {  // processor.rs:174-177
  a[0] = config.source != OMB_CHANNEL_NONE;
  a[1] = config.secondary != OMB_CHANNEL_NONE && config.secondary != config.source;
}

int SpectrumStream::reset_level_buffers() {  // processor.rs:152-168
  OMB_TRY(ensure_stream());
  const size_t bins = (size_t)config.bins();
  for (auto& t : h_traces)
    for (auto& v : t) v.assign(bins, config.floor_db);
  bool act[2];
  active_traces(act);
  for (int t = 0; t < 2; ++t) {
    if (act[t] && config.averaging != OMB_AVG_NONE) {
      OMB_TRY(d_state[t].reserve(bins));
      OMB_CUDA_TRY(cudaMemsetAsync(d_state[t].ptr, 0, bins * sizeof(float), stream));
    }
  }
  return OMB_OK;
}

int SpectrumStream::reset_buffers() {  // processor.rs:138-150
  OMB_TRY(ensure_stream());
  OMB_TRY(sync_plan());
  h_freq = plan->h_freq;
  OMB_TRY(reset_level_buffers());
  pcm[0].clear();
  pcm[1].clear();
  pending_skip = 0;
  return OMB_OK;
}

int SpectrumStream::rebuild_fft() {  // processor.rs:126-136
  OMB_TRY(ensure_stream());
  OMB_TRY(sync_plan());
  prepared = true;
  return reset_buffers();
}

int SpectrumStream::prepare() { return prepared ? OMB_OK : rebuild_fft(); }  // processor.rs:120-124

int SpectrumStream::reset_audio() {  // processor.rs:112-118
  if (prepared) OMB_TRY(reset_level_buffers());
  pcm[0].clear();
  pcm[1].clear();
  pending_skip = 0;
  return OMB_OK;
}

int SpectrumStream::update_config(const omb_spectrum_config& c) {  // processor.rs:300-322
  const SpectrumConfigN old = config;
  const SpectrumConfigN next = SpectrumConfigN::from_c(c);
  // sizes without a kernel are refused and the previous, working configuration is kept (see stream_spectrogram.cu)
  if (prepared && (!is_pow2(next.fft_size) || next.fft_size > (1ull << 24)))
    return fail(OMB_ERR_UNSUPPORTED, "update_config: spectrum fft_size %llu has no kernel (power-of-two lengths only, no CPU fallback); previous config kept",
                (unsigned long long)next.fft_size);
  config = next;
  if (!prepared) return OMB_OK;
  const bool mode_changed = old.averaging != config.averaging;
  if (old.fft_size != config.fft_size || old.window_kind != config.window_kind) return rebuild_fft();
  if (old.sample_rate != config.sample_rate || old.hop != config.hop || old.source != config.source ||
      old.secondary != config.secondary)
    return reset_buffers();
  if (mode_changed || std::fabs(old.floor_db - config.floor_db) > std::numeric_limits<float>::epsilon()) {
    OMB_TRY(sync_plan());
    return reset_level_buffers();
  }
  return OMB_OK;
}

void SpectrumStream::fill(omb_spectrum_snapshot* out) {
  out->bins = (uint32_t)h_freq.size();
  out->frequency_bins = h_freq.data();
  for (int t = 0; t < 2; ++t)
    for (int w = 0; w < 2; ++w) out->traces[t][w] = h_traces[t][w].data();
}

int SpectrumStream::process_block(const float* samples, size_t n_samples, uint32_t channels, float sample_rate,
                                  const uint8_t* positions, omb_spectrum_snapshot* out) {
  channels = std::min<uint32_t>(std::max<uint32_t>(channels, 1), OMB_MAX_CHANNELS);
  if (n_samples < channels) return OMB_NO_DATA;
  if (!samples || !out) return fail(OMB_ERR_INVALID, "null argument");
  const float sr = sanitize_sample_rate(sample_rate);
  if (sr != config.sample_rate) {  // processor.rs:258-263
    config.sample_rate = sr;
    if (prepared) OMB_TRY(reset_buffers());
  }
  OMB_TRY(prepare());
  OMB_TRY(sync_plan());
  bool act[2];
  active_traces(act);

  // push_sources, processor.rs:271-298
  const size_t frames = n_samples / channels;
  const size_t skip = (size_t)std::min<uint64_t>(pending_skip, frames);
  pending_skip -= skip;
  if (skip < frames && (act[0] || act[1])) {
    const size_t fresh = frames - skip;
    OMB_TRY(d_block.upload(samples, frames * channels, stream));
    for (int t = 0; t < 2; ++t)
      if (act[t]) OMB_TRY(pcm[t].make_room(fresh, stream));
    const StereoMatrix m = make_stereo_matrix(channels, positions);
    OMB_TRY(launch_downmix(d_block.ptr, skip, fresh, channels, m, (int)config.source, act[0] ? pcm[0].tail() : nullptr,
                           (int)config.secondary, act[1] ? pcm[1].tail() : nullptr, dev.sm_count, stream));
    for (int t = 0; t < 2; ++t)
      if (act[t]) pcm[t].commit(fresh);
  }

  // process_ready_windows, processor.rs:179-213
  if (!act[0] && !act[1]) return OMB_NO_DATA;
  const uint64_t N = config.fft_size, hop = config.hop, bins = config.bins();
  uint64_t n = ~0ull;
  for (int t = 0; t < 2; ++t)
    if (act[t]) n = std::min<uint64_t>(n, pcm[t].len >= N ? (pcm[t].len - N) / hop + 1 : 0);
  if (n == 0) return OMB_NO_DATA;
  OMB_TRY(d_power.reserve((size_t)(n * bins)));
  for (int t = 0; t < 2; ++t) {
    if (!act[t]) continue;
    OMB_TRY(plan->power_device(pcm[t].data(), 1, n, pcm[t].len, d_power.ptr, stream));
    for (int w = 0; w < 2; ++w) OMB_TRY(d_out[t][w].reserve((size_t)bins));
    float* state = config.averaging != OMB_AVG_NONE ? d_state[t].ptr : nullptr;
    OMB_TRY(plan->smooth_device(d_power.ptr, 1, n, state, d_out[t][0].ptr, d_out[t][1].ptr, nullptr, false, stream));
    for (int w = 0; w < 2; ++w) {
      h_traces[t][w].resize((size_t)bins);
      OMB_CUDA_TRY(cudaMemcpyAsync(h_traces[t][w].data(), d_out[t][w].ptr, bins * sizeof(float), cudaMemcpyDeviceToHost, stream));
    }
  }
  OMB_CUDA_TRY(cudaStreamSynchronize(stream));
  uint64_t drained = n * hop;
  for (int t = 0; t < 2; ++t)
    if (act[t]) {
      const uint64_t count = std::min<uint64_t>(n * hop, pcm[t].len);
      pcm[t].drain((size_t)count);
      drained = std::min(drained, count);
    }
  pending_skip += n * hop - drained;
  fill(out);
  return OMB_OK;
}

}  // namespace omb
