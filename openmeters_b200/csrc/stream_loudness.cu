// stream_loudness.cu — streaming LoudnessProcessor (loudness/processor.rs:218-311).
// Per-channel filter / window / true-peak state lives on the device between calls.
#include "streams.h"

#include <algorithm>

namespace omb {

LoudnessStream::LoudnessStream(const omb_loudness_config& c) : cfg(c), sample_rate(c.sample_rate) {
  configure_rate(sanitize_sample_rate(c.sample_rate));  // processor.rs:225-232: weighting from the sanitised rate,
  sample_rate = c.sample_rate;                           // config kept as given (a NaN rate re-derives on first block)
  true_peak_fir4_host(fir.fir4);
  true_peak_fir2_host(fir.fir2);
}

LoudnessStream::~LoudnessStream() {
  if (stream) cudaStreamDestroy(stream);
}

int LoudnessStream::ensure_stream() {
  if (!stream) {
    OMB_TRY(current_device(&dev));
    OMB_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  }
  return OMB_OK;
}

void LoudnessStream::configure_rate(float sr) {
  sample_rate = sr;
  k_weighting_host((double)sr, kw.b, kw.a);
}

int LoudnessStream::reset_audio() {  // processor.rs:234-236
  if (!channels) return OMB_OK;
  OMB_TRY(ensure_stream());
  OMB_CUDA_TRY(cudaMemsetAsync(core.d_state.ptr, 0, sizeof(LoudChannelState) * channels * n_streams, stream));
  OMB_CUDA_TRY(cudaMemsetAsync(core.d_ring.ptr, 0, sizeof(double) * ring_len * channels * n_streams, stream));
  return OMB_OK;
}

int LoudnessStream::ensure_state(uint32_t requested, float sr_in) {  // processor.rs:238-251
  const uint32_t ch = std::min<uint32_t>(std::max<uint32_t>(requested, 1), OMB_MAX_CHANNELS);
  const float sr = sanitize_sample_rate(sr_in);
  const bool rate_changed = !(sample_rate == sr);
  if (rate_changed) configure_rate(sr);
  static const float kWindows[kLoudWindows] = {3.0f, 0.4f, 0.3f, 1.0f};
  uint64_t longest = 1;
  for (int w = 0; w < kLoudWindows; ++w) {
    caps[w] = std::max<uint64_t>(loudness_window_length(sample_rate, kWindows[w]), 1);
    longest = std::max(longest, caps[w]);
  }
  tp_delay_len = (double)sample_rate < 96000.0 ? 12 : ((double)sample_rate < 192000.0 ? 24 : 0);
  if (rate_changed || channels != ch || ring_len != longest) {
    channels = ch;
    ring_len = longest;
    OMB_TRY(core.d_state.reserve((size_t)channels * n_streams));
    OMB_TRY(core.d_ring.reserve((size_t)(ring_len * channels) * n_streams));
    OMB_TRY(reset_audio());
  }
  return OMB_OK;
}

int LoudnessStream::process_block(const float* samples, size_t n_samples, uint32_t ch_in, float sr_in, const uint8_t* positions,
                                  omb_loudness_snapshot* out) {
  if (n_streams != 1) return fail(OMB_ERR_INVALID, "process_block on a bank: use process_bank");
  return process_bank(samples, 0, n_samples, ch_in, sr_in, positions, out);
}

int LoudnessStream::process_bank(const float* samples, uint64_t stream_stride, size_t n_samples, uint32_t ch_in, float sr_in,
                                 const uint8_t* positions, omb_loudness_snapshot* out) {
  const uint32_t ch = std::min<uint32_t>(std::max<uint32_t>(ch_in, 1), OMB_MAX_CHANNELS);
  if (n_samples < ch) return OMB_NO_DATA;
  if (!samples || !out) return fail(OMB_ERR_INVALID, "null argument");
  OMB_TRY(ensure_stream());
  OMB_TRY(ensure_state(ch, sr_in));
  const uint64_t frames = n_samples / ch;
  const size_t per = (size_t)(frames * ch);
  OMB_TRY(core.d_block.reserve(per * n_streams));
  if (n_streams == 1) {
    OMB_CUDA_TRY(cudaMemcpyAsync(core.d_block.ptr, samples, per * sizeof(float), cudaMemcpyHostToDevice, stream));
  } else {  // one strided copy for all streams
    OMB_CUDA_TRY(cudaMemcpy2DAsync(core.d_block.ptr, per * sizeof(float), samples, (size_t)stream_stride * sizeof(float), per * sizeof(float),
                                   n_streams, cudaMemcpyHostToDevice, stream));
  }
  OMB_TRY(core.d_snap.reserve(n_streams));
  LoudStreamArgs a{};
  a.block = core.d_block.ptr;
  a.block_stride = per;
  a.frames = frames;
  a.channels = channels;
  a.state = core.d_state.ptr;
  a.ring = core.d_ring.ptr;
  a.ring_len = ring_len;
  for (int w = 0; w < kLoudWindows; ++w) a.caps[w] = caps[w];
  a.tp_delay_len = tp_delay_len;
  a.kw = kw;
  a.fir = fir;
  a.floor_db = cfg.floor_db;
  uint8_t fb[OMB_MAX_CHANNELS];
  if (!positions) {
    fallback_positions_host(ch, fb);
    positions = fb;
  }
  for (int i = 0; i < OMB_MAX_CHANNELS; ++i) {
    a.positions[i] = positions[i];
    a.weights[i] = channel_weight_host(positions[i]);
  }
  a.out = core.d_snap.ptr;
  OMB_TRY(core.d_vnew.reserve(per * n_streams));
  a.vnew = core.d_vnew.ptr;
  OMB_TRY(launch_loudness_stream(a, stream, n_streams));
  OMB_CUDA_TRY(cudaMemcpyAsync(out, core.d_snap.ptr, sizeof(omb_loudness_snapshot) * n_streams, cudaMemcpyDeviceToHost, stream));
  OMB_CUDA_TRY(cudaStreamSynchronize(stream));
  return OMB_OK;
}

}  // namespace omb
