// stream_bank.cu — the device-side multi-stream ring of SURVEY.md §8 row f1: S spectrogram streams with one config,
// advanced in lock-step.  Same state machine as SpectrogramStream (spectrogram/processor.rs:281-437,490-516: carry-over
// of read_len - hop samples, hop > window skip accounting, history retention, reset flag), but the pending audio of all
// streams lives in ONE device buffer [stream][capacity] and every push costs one H2D copy, one batched fold-down launch
// and one launch of the batched STFT kernel for all streams — instead of S x (copy + 2 launches + sync) — which is what
// makes the batched kernels a drop-in for live ring-buffer input at thousands of streams.
#include <algorithm>
#include <memory>
#include <new>

#include "downmix.h"
#include "stft.h"

namespace omb {

namespace {

// dsp.rs:223-257 for every stream of the bank: grid.y = stream.
__global__ void __launch_bounds__(256) k_downmix_bank(const float* __restrict__ in, uint64_t in_stride, uint64_t first_frame, uint64_t frames,
                                                      int channels, StereoMatrix m, float* __restrict__ out, uint64_t out_stride) {
  const float* src = in + (uint64_t)blockIdx.y * in_stride;
  float* dst = out + (uint64_t)blockIdx.y * out_stride;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < frames; f += stride) {
    const float* fr = src + (first_frame + f) * (uint64_t)channels;
    float l = 0.0f, r = 0.0f;
    for (int c = 0; c < channels; ++c) {
      const float s = __ldg(&fr[c]);
      l = __fadd_rn(l, __fmul_rn(s, m.w[c][0]));
      r = __fadd_rn(r, __fmul_rn(s, m.w[c][1]));
    }
    dst[f] = __fmul_rn(__fadd_rn(l, r), 0.5f);  // Channel::Mid (util/audio/channel.rs:12-21)
  }
}

}  // namespace

struct SpectrogramBank {
  StftConfig config;
  uint32_t S = 0;
  bool reset = true;
  uint64_t pending_skip = 0;
  // ring: [S][cap] floats, the pending samples of stream s are ring[cur] + s * cap + begin .. + len
  DeviceBuffer<float> ring[2];
  int cur = 0;
  size_t cap = 0, begin = 0, len = 0;
  std::unique_ptr<StftPlan> plan;
  omb_spectrogram_config plan_cfg{};
  DeviceInfo dev;
  cudaStream_t stream = nullptr;
  DeviceBuffer<float> d_block;
  DeviceBuffer<omb_spectrogram_point> d_points;
  DeviceBuffer<uint32_t> d_counts;
  DeviceBuffer<uint16_t> d_classic;
  PinnedBuffer<float> h_block;
  PinnedBuffer<omb_spectrogram_point> h_points;
  PinnedBuffer<uint32_t> h_counts;
  PinnedBuffer<uint16_t> h_classic;

  ~SpectrogramBank() {
    if (stream) cudaStreamDestroy(stream);
  }

  int ensure() {
    if (!stream) {
      OMB_TRY(current_device(&dev));
      OMB_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    }
    omb_spectrogram_config c;
    config.to_c(&c);
    if (plan && std::memcmp(&c, &plan_cfg, sizeof c) == 0) return OMB_OK;
    plan.reset(new StftPlan());
    const int rc = plan->init(c, OMB_KERNEL_AUTO);
    if (rc < 0) {
      plan.reset();
      return rc;
    }
    plan_cfg = c;
    return OMB_OK;
  }

  const float* data() const { return ring[cur].ptr + begin; }

  // Moves the pending samples of every stream to offset 0 of the other buffer (grown if needed).
  int relocate(size_t need_cap) {
    DeviceBuffer<float>& other = ring[1 - cur];
    size_t new_cap = std::max(cap, (size_t)4096);
    while (new_cap < need_cap) new_cap *= 2;
    new_cap = (new_cap + 3) & ~(size_t)3;
    if (other.cap < new_cap * S) OMB_TRY(other.reserve(new_cap * S));
    if (len)
      OMB_CUDA_TRY(cudaMemcpy2DAsync(other.ptr, new_cap * sizeof(float), data(), cap * sizeof(float), len * sizeof(float), S,
                                     cudaMemcpyDeviceToDevice, stream));
    cur = 1 - cur;
    cap = new_cap;
    begin = 0;
    // keep both buffers at the same pitch so the next relocation can ping-pong without reallocating
    return OMB_OK;
  }

  int make_room(size_t extra) {
    if (cap && begin + len + extra <= cap && ring[cur].cap >= cap * S) return OMB_OK;
    return relocate(len + extra);
  }

  void drain(size_t n) {
    n = std::min(n, len);
    begin += n;
    len -= n;
    if (len == 0) begin = 0;
  }

  void advance_audio(uint64_t count) {  // processor.rs:406-410
    const uint64_t missing = count > len ? count - len : 0;
    drain((size_t)std::min<uint64_t>(count, len));
    pending_skip += missing;
  }

  void reset_audio() {  // processor.rs:212-217
    begin = len = 0;
    pending_skip = 0;
    reset = true;
  }

  int push(const float* samples, uint64_t stream_stride, size_t frames, uint32_t channels, float sample_rate, const uint8_t* positions,
           omb_spectrogram_bank_update* out) {
    channels = std::min<uint32_t>(std::max<uint32_t>(channels, 1), OMB_MAX_CHANNELS);
    if (frames == 0) return OMB_NO_DATA;
    if (!samples || !out) return fail(OMB_ERR_INVALID, "null argument");
    if (stream_stride < (uint64_t)frames * channels) return fail(OMB_ERR_INVALID, "stream_stride smaller than one block");
    const float sr = sanitize_sample_rate(sample_rate);
    if (config.sample_rate != sr) {  // processor.rs:494-499
      config.sample_rate = sr;
      begin = len = 0;
      pending_skip = 0;
      reset = true;
    }
    OMB_TRY(ensure());

    // push_audio (processor.rs:412-437) for all streams
    const size_t skip = (size_t)std::min<uint64_t>(pending_skip, frames);
    pending_skip -= skip;
    const size_t fresh = frames - skip;
    if (fresh) {
      OMB_TRY(make_room(fresh));
      float* tail = ring[cur].ptr + begin + len;
      const size_t block = frames * channels;
      if (channels == 1) {  // mono blocks bypass the fold-down
        OMB_CUDA_TRY(cudaMemcpy2DAsync(tail, cap * sizeof(float), samples + skip, stream_stride * sizeof(float), fresh * sizeof(float), S,
                                       cudaMemcpyHostToDevice, stream));
      } else {
        OMB_TRY(d_block.reserve(block * S));
        OMB_CUDA_TRY(cudaMemcpy2DAsync(d_block.ptr, block * sizeof(float), samples, stream_stride * sizeof(float), block * sizeof(float), S,
                                       cudaMemcpyHostToDevice, stream));
        const StereoMatrix m = make_stereo_matrix(channels, positions);
        const unsigned gx = (unsigned)std::min<uint64_t>((fresh + 255) / 256, 64);
        OMB_LAUNCH(k_downmix_bank, dim3(gx, S), dim3(256), 0, stream, d_block.ptr, (uint64_t)block, (uint64_t)skip, (uint64_t)fresh, (int)channels,
                   m, tail, (uint64_t)cap);
        OMB_CHECK_LAUNCH();
      }
      len += fresh;
    }

    // process_ready_windows (processor.rs:281-388), the same for every stream
    const uint64_t hop = config.hop, read_len = config.read_len(), bins = config.bins();
    const uint64_t ready = len >= read_len ? (len - read_len) / hop + 1 : 0;
    const uint64_t retained = history_columns(config.reassign, (uint32_t)bins, (size_t)config.history_length);
    const uint64_t skip_cols = ready > retained ? ready - retained : 0;
    advance_audio(skip_cols * hop);
    const uint64_t n = ready - skip_cols;
    if (n == 0) return OMB_NO_DATA;
    if (begin % 4 != 0) OMB_TRY(relocate(len));  // the specialised kernels want 16-byte aligned lanes
    const uint64_t slots = n * S;
    if (config.reassign) {
      OMB_TRY(d_points.reserve((size_t)(slots * bins)));
      OMB_TRY(d_counts.reserve((size_t)slots));
      OMB_TRY(h_points.reserve((size_t)(slots * bins)));
      OMB_TRY(h_counts.reserve((size_t)slots));
      OMB_TRY(plan->execute_device(data(), S, len, cap, d_points.ptr, bins, d_counts.ptr, nullptr, stream));
      OMB_CUDA_TRY(cudaMemcpyAsync(h_counts.ptr, d_counts.ptr, slots * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
      OMB_CUDA_TRY(cudaMemcpyAsync(h_points.ptr, d_points.ptr, slots * bins * sizeof(omb_spectrogram_point), cudaMemcpyDeviceToHost, stream));
    } else {
      OMB_TRY(d_classic.reserve((size_t)(slots * bins)));
      OMB_TRY(h_classic.reserve((size_t)(slots * bins)));
      OMB_TRY(plan->execute_device(data(), S, len, cap, nullptr, 0, nullptr, d_classic.ptr, stream));
      OMB_CUDA_TRY(cudaMemcpyAsync(h_classic.ptr, d_classic.ptr, slots * bins * sizeof(uint16_t), cudaMemcpyDeviceToHost, stream));
    }
    OMB_CUDA_TRY(cudaStreamSynchronize(stream));
    advance_audio(n * hop);

    out->fft_size = config.fft_len();
    out->hop_size = config.hop;
    out->history_length = config.history_length;
    out->sample_rate = config.sample_rate;
    out->reassigned_power_scale = plan->power_scale;
    out->reset = reset ? 1 : 0;
    reset = false;
    out->kind = config.reassign ? OMB_COLUMN_REASSIGNED : OMB_COLUMN_CLASSIC;
    out->n_streams = S;
    out->n_columns = (uint32_t)n;
    out->bins = (uint32_t)bins;
    out->_pad = 0;
    out->counts = config.reassign ? h_counts.ptr : nullptr;
    out->points = config.reassign ? h_points.ptr : nullptr;
    out->classic_db = config.reassign ? nullptr : h_classic.ptr;
    return OMB_OK;
  }
};

}  // namespace omb

using namespace omb;

struct omb_spectrogram_bank {
  SpectrogramBank b;
};

extern "C" {

int omb_spectrogram_bank_create(const omb_spectrogram_config* cfg, uint32_t n_streams, omb_spectrogram_bank** out) {
  if (!cfg || !out || n_streams == 0) return fail(OMB_ERR_INVALID, "invalid argument");
  // the batched fold-down / smoothing launches address the stream (x trace) through gridDim.y
  if (n_streams > 32767u) return fail(OMB_ERR_UNSUPPORTED, "spectrogram bank: at most 32767 streams per bank (got %u)", n_streams);
  try {
    auto* h = new omb_spectrogram_bank();
    h->b.config = StftConfig::from_c(*cfg);
    h->b.S = n_streams;
    *out = h;
    return OMB_OK;
  } catch (const std::bad_alloc&) {
    return fail(OMB_ERR_NOMEM, "host allocation failed");
  }
}
void omb_spectrogram_bank_destroy(omb_spectrogram_bank* b) { delete b; }
int omb_spectrogram_bank_reset_audio(omb_spectrogram_bank* b) {
  if (!b) return fail(OMB_ERR_INVALID, "null argument");
  b->b.reset_audio();
  return OMB_OK;
}
int omb_spectrogram_bank_push(omb_spectrogram_bank* b, const float* samples, uint64_t stream_stride, size_t frames, uint32_t channels,
                              float sample_rate, const uint8_t positions[OMB_MAX_CHANNELS], omb_spectrogram_bank_update* out) {
  if (!b) return fail(OMB_ERR_INVALID, "null argument");
  try {
    return b->b.push(samples, stream_stride, frames, channels, sample_rate, positions, out);
  } catch (const std::bad_alloc&) {
    return fail(OMB_ERR_NOMEM, "host allocation failed");
  } catch (...) {
    return fail(OMB_ERR_INVALID, "unexpected C++ exception");
  }
}
size_t omb_spectrogram_bank_pending(const omb_spectrogram_bank* b) { return b ? b->b.len : 0; }

}  // extern "C"
