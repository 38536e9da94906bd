// stream_spectrogram.cu — streaming SpectrogramProcessor over the batched STFT plan.
//
// Host logic only: carry-over of `read_len - hop` samples, hop > window
// skip accounting, history retention, reset flag — exactly spectrogram/processor.rs:187-543.  The
// pending audio lives in a device-resident FIFO; every ready column of a call is computed by ONE
// launch of the batched kernel (the reference loops column by column on the CPU).
#include "streams.h"

#include <algorithm>

namespace omb {

SpectrogramStream::SpectrogramStream(const omb_spectrogram_config& c) { config = StftConfig::from_c(c); }

SpectrogramStream::~SpectrogramStream() {
  if (stream) cudaStreamDestroy(stream);
}

int SpectrogramStream::ensure_stream() {
  if (!stream) {
    OMB_TRY(current_device(&dev));
    OMB_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  }
  return OMB_OK;
}

// The plan bakes every config field the kernels read; rebuilt whenever one of them changes.
int SpectrogramStream::sync_plan() {
  omb_spectrogram_config c;
  config.to_c(&c);
  if (plan && std::memcmp(&c, &plan_cfg, sizeof c) == 0) return OMB_OK;
  plan.reset(new StftPlan());
  int rc = plan->init(c, OMB_KERNEL_AUTO);
  if (rc < 0) {
    plan.reset();
    return rc;
  }
  plan_cfg = c;
  return OMB_OK;
}

void SpectrogramStream::reset_audio() {  // processor.rs:212-217
  pending.clear();
  pending_skip = 0;
  reset = true;
}

int SpectrogramStream::rebuild_fft() {  // processor.rs:229-279 (buffer semantics; tables live in the plan)
  OMB_TRY(ensure_stream());
  OMB_TRY(sync_plan());
  prepared = true;
  const uint64_t active_len = config.reassign ? config.hilbert_len() : config.fft_len();
  const uint64_t buffered = active_len * 2;
  if (pending.len > buffered) pending.drain(pending.len - (size_t)buffered);
  OMB_TRY(pending.realign(stream));  // keeps the specialised kernels' 16-byte alignment across odd-sized drains
  pending_skip = 0;
  return OMB_OK;
}

int SpectrogramStream::prepare() {  // processor.rs:219-223
  if (!prepared) return rebuild_fft();
  return OMB_OK;
}

void SpectrogramStream::advance_audio(uint64_t count) {  // processor.rs:406-410
  const uint64_t missing = count > pending.len ? count - pending.len : 0;
  pending.drain((size_t)std::min<uint64_t>(count, pending.len));
  pending_skip += missing;
}

int SpectrogramStream::update_config(const omb_spectrogram_config& c) {  // processor.rs:518-543
  const StftConfig next = StftConfig::from_c(c);
  const StftConfig prev = config;
  // Deliberate difference from the reference (rustfft is mixed-radix, any size is legal there): transform lengths without a
  // kernel are refused HERE and the handle keeps its previous, working configuration — instead of accepting the config and
  // failing every later process_block.
  if (prepared && (!is_pow2(next.window) || !is_pow2(next.fft_len()) || next.fft_len() > (1ull << 24) || next.hilbert_len() > (1ull << 24)))
    return fail(OMB_ERR_UNSUPPORTED,
                "update_config: fft_size %llu x zero_padding %llu has no kernel (power-of-two lengths only, no CPU fallback); previous config kept",
                (unsigned long long)next.window, (unsigned long long)next.zero_pad);
  config = next;
  const bool rate_changed = prev.sample_rate != next.sample_rate;
  const bool rebuild = prev.window != next.window || prev.zero_pad != next.zero_pad || prev.window_kind != next.window_kind ||
                       prev.reassign != next.reassign || rate_changed;
  int rc = OMB_OK;
  if (rebuild && prepared) {
    rc = rebuild_fft();
    if (rate_changed) pending.clear();
  }
  const bool hop_changed = prev.hop != next.hop;
  if (hop_changed) pending_skip = 0;
  reset = reset || rebuild || hop_changed;
  return rc;
}

int SpectrogramStream::push_audio(const float* samples, size_t n_samples, uint32_t channels, const uint8_t* positions) {
  // processor.rs:412-437
  const size_t frames = n_samples / channels;
  const size_t skip = (size_t)std::min<uint64_t>(pending_skip, frames);
  pending_skip -= skip;
  if (skip == frames) return OMB_OK;
  const size_t fresh = frames - skip;
  OMB_TRY(pending.make_room(fresh, stream));
  if (channels == 1) {  // mono blocks bypass the fold-down (processor.rs:420-428)
    OMB_CUDA_TRY(cudaMemcpyAsync(pending.tail(), samples + skip, fresh * sizeof(float), cudaMemcpyHostToDevice, stream));
  } else {
    OMB_TRY(d_block.upload(samples, frames * channels, stream));
    const StereoMatrix m = make_stereo_matrix(channels, positions);
    OMB_TRY(launch_downmix(d_block.ptr, skip, fresh, channels, m, OMB_CHANNEL_MID, pending.tail(), OMB_CHANNEL_NONE, nullptr,
                           dev.sm_count, stream));
  }
  pending.commit(fresh);
  return OMB_OK;
}

int SpectrogramStream::process_block(const float* samples, size_t n_samples, uint32_t channels, float sample_rate,
                                     const uint8_t* positions, omb_spectrogram_update* out) {
  // processor.rs:490-516
  channels = std::min<uint32_t>(std::max<uint32_t>(channels, 1), OMB_MAX_CHANNELS);
  if (n_samples < channels) return OMB_NO_DATA;  // AudioBlock::is_empty (dsp.rs:259-261)
  if (!samples || !out) return fail(OMB_ERR_INVALID, "null argument");
  const float sr = sanitize_sample_rate(sample_rate);
  if (config.sample_rate != sr) {
    config.sample_rate = sr;
    OMB_TRY(rebuild_fft());
    pending.clear();
    reset = true;
  }
  OMB_TRY(prepare());
  OMB_TRY(sync_plan());
  OMB_TRY(push_audio(samples, n_samples, channels, positions));

  // process_ready_windows, processor.rs:281-388
  const uint64_t hop = config.hop, read_len = config.read_len(), bins = config.bins();
  const uint64_t ready = pending.len >= read_len ? (pending.len - read_len) / hop + 1 : 0;
  const uint64_t retained = history_columns(config.reassign, (uint32_t)bins, (size_t)config.history_length);
  const uint64_t skip_cols = ready > retained ? ready - retained : 0;
  advance_audio(skip_cols * hop);
  const uint64_t n = ready - skip_cols;
  if (n == 0) return OMB_NO_DATA;

  offsets.assign(1, 0);
  points.clear();
  classic.clear();
  if (config.reassign) {
    OMB_TRY(d_points.reserve((size_t)(n * bins)));
    OMB_TRY(d_counts.reserve((size_t)n));
    OMB_TRY(plan->execute_device(pending.data(), 1, pending.len, (pending.len + 3) & ~(size_t)3, d_points.ptr, bins, d_counts.ptr, nullptr, stream));
    OMB_TRY(h_counts.reserve((size_t)n));
    OMB_TRY(h_points.reserve((size_t)(n * bins)));
    OMB_CUDA_TRY(cudaMemcpyAsync(h_counts.ptr, d_counts.ptr, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    OMB_CUDA_TRY(cudaMemcpyAsync(h_points.ptr, d_points.ptr, n * bins * sizeof(omb_spectrogram_point), cudaMemcpyDeviceToHost, stream));
    OMB_CUDA_TRY(cudaStreamSynchronize(stream));
    for (uint64_t c = 0; c < n; ++c) {
      const omb_spectrogram_point* src = h_points.ptr + c * bins;
      points.insert(points.end(), src, src + h_counts.ptr[c]);
      offsets.push_back((uint32_t)points.size());
    }
  } else {
    OMB_TRY(d_classic.reserve((size_t)(n * bins)));
    OMB_TRY(plan->execute_device(pending.data(), 1, pending.len, (pending.len + 3) & ~(size_t)3, nullptr, 0, nullptr, d_classic.ptr, stream));
    classic.resize((size_t)(n * bins));
    OMB_CUDA_TRY(cudaMemcpyAsync(classic.data(), d_classic.ptr, n * bins * sizeof(uint16_t), cudaMemcpyDeviceToHost, stream));
    OMB_CUDA_TRY(cudaStreamSynchronize(stream));
    for (uint64_t c = 1; c <= n; ++c) offsets.push_back((uint32_t)(c * bins));
  }
  advance_audio(n * hop);

  out->fft_size = config.fft_len();
  out->hop_size = config.hop;
  out->history_length = config.history_length;
  out->sample_rate = config.sample_rate;
  out->reassigned_power_scale = plan->power_scale;
  out->reset = reset ? 1 : 0;
  reset = false;
  out->kind = config.reassign ? OMB_COLUMN_REASSIGNED : OMB_COLUMN_CLASSIC;
  out->n_columns = (uint32_t)n;
  out->bins = (uint32_t)bins;
  out->column_offsets = offsets.data();
  out->points = points.data();
  out->classic_db = classic.data();
  return OMB_OK;
}

}  // namespace omb
