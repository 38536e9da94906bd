// spectrum_bank.cu — the device-side multi-stream ring of SURVEY.md §8 row f1 for the spectrum analyzer: S
// SpectrumProcessors (spectrum/processor.rs:88-323) with one config, advanced in lock-step.  Same state machine as
// SpectrumStream (stream_spectrum.cu: source projections, hop / skip accounting, reset rules, smoothing state resident
// on the device), but the pending audio of all streams' traces lives in ONE device buffer [stream * traces][capacity] and a
// push costs one strided H2D copy, one batched fold-down launch and one power + one smoothing launch for all streams.
#include <algorithm>
#include <memory>
#include <new>

#include "downmix.h"
#include "spectrum.h"

namespace omb {

namespace {

__device__ __forceinline__ float bank_project(int channel, float l, float r) {  // util/audio/channel.rs:12-21
  switch (channel) {
    case OMB_CHANNEL_LEFT: return l;
    case OMB_CHANNEL_RIGHT: return r;
    case OMB_CHANNEL_MID: return __fmul_rn(__fadd_rn(l, r), 0.5f);
    case OMB_CHANNEL_SIDE: return __fmul_rn(__fsub_rn(l, r), 0.5f);
    default: return 0.0f;
  }
}

// dsp.rs:223-257 + Channel::project for every stream of the bank: grid.y = stream; up to two projected lanes per stream,
// lane (stream * n_traces + i) of the ring.
__global__ void __launch_bounds__(256) k_downmix_spectrum_bank(const float* __restrict__ in, uint64_t in_stride, uint64_t first_frame,
                                                               uint64_t frames, int channels, StereoMatrix m, int proj0, int proj1, int n_traces,
                                                               float* __restrict__ out, uint64_t out_stride) {
  const float* src = in + (uint64_t)blockIdx.y * in_stride;
  float* dst0 = out + (uint64_t)blockIdx.y * n_traces * out_stride;
  float* dst1 = dst0 + out_stride;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < frames; f += stride) {
    const float* fr = src + (first_frame + f) * (uint64_t)channels;
    float l = 0.0f, r = 0.0f;
    for (int c = 0; c < channels; ++c) {
      const float s = __ldg(&fr[c]);
      l = __fadd_rn(l, __fmul_rn(s, m.w[c][0]));
      r = __fadd_rn(r, __fmul_rn(s, m.w[c][1]));
    }
    dst0[f] = bank_project(proj0, l, r);
    if (n_traces > 1) dst1[f] = bank_project(proj1, l, r);
  }
}

}  // namespace

struct SpectrumBank {
  SpectrumConfigN config;
  uint32_t S = 0;
  bool prepared = false;
  uint64_t pending_skip = 0;
  // ring: [S * T][cap] floats; the pending samples of lane l are ring[cur] + l * cap + begin .. + len
  DeviceBuffer<float> ring[2];
  int cur = 0;
  size_t cap = 0, begin = 0, len = 0;
  uint32_t ring_lanes = 0;
  std::unique_ptr<SpectrumPlan> plan;
  omb_spectrum_config plan_cfg{};
  DeviceInfo dev;
  cudaStream_t stream = nullptr;
  DeviceBuffer<float> d_block, d_power, d_state, d_w, d_r;
  PinnedBuffer<float> h_w, h_r;
  std::vector<float> h_freq;

  ~SpectrumBank() {
    if (stream) cudaStreamDestroy(stream);
  }

  // processor.rs:174-177 — trace 0: the primary source, trace 1: a distinct secondary source
  void active(bool a[2]) const {
    a[0] = config.source != OMB_CHANNEL_NONE;
    a[1] = config.secondary != OMB_CHANNEL_NONE && config.secondary != config.source;
  }
  uint32_t traces() const {
    bool a[2];
    active(a);
    return (a[0] ? 1u : 0u) + (a[1] ? 1u : 0u);
  }

  int ensure() {
    if (!stream) {
      OMB_TRY(current_device(&dev));
      OMB_CUDA_TRY(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    }
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
    h_freq = plan->h_freq;
    return OMB_OK;
  }

  const float* data() const { return ring[cur].ptr + begin; }

  int relocate(size_t need_cap, uint32_t lanes) {
    DeviceBuffer<float>& other = ring[1 - cur];
    size_t new_cap = std::max(cap, (size_t)4096);
    while (new_cap < need_cap) new_cap *= 2;
    new_cap = (new_cap + 3) & ~(size_t)3;
    if (other.cap < new_cap * lanes) OMB_TRY(other.reserve(new_cap * lanes));
    if (len)
      OMB_CUDA_TRY(cudaMemcpy2DAsync(other.ptr, new_cap * sizeof(float), data(), cap * sizeof(float), len * sizeof(float), lanes,
                                     cudaMemcpyDeviceToDevice, stream));
    cur = 1 - cur;
    cap = new_cap;
    begin = 0;
    return OMB_OK;
  }

  int make_room(size_t extra, uint32_t lanes) {
    if (cap && begin + len + extra <= cap && ring[cur].cap >= cap * lanes) return OMB_OK;
    return relocate(len + extra, lanes);
  }

  void drain(size_t n) {
    n = std::min(n, len);
    begin += n;
    len -= n;
    if (len == 0) begin = 0;
  }

  int reset_levels() {  // processor.rs:152-168: traces back to the floor, smoothing state to zero
    const uint32_t lanes = S * traces();
    const size_t bins = (size_t)config.bins();
    if (lanes && config.averaging != OMB_AVG_NONE) {
      OMB_TRY(d_state.reserve(bins * lanes));
      OMB_CUDA_TRY(cudaMemsetAsync(d_state.ptr, 0, bins * lanes * sizeof(float), stream));
    }
    return OMB_OK;
  }

  int reset_buffers() {  // processor.rs:138-150
    OMB_TRY(ensure());
    OMB_TRY(reset_levels());
    begin = len = 0;
    pending_skip = 0;
    return OMB_OK;
  }

  int reset_audio() {  // processor.rs:112-118
    if (prepared) {
      OMB_TRY(ensure());
      OMB_TRY(reset_levels());
    }
    begin = len = 0;
    pending_skip = 0;
    return OMB_OK;
  }

  int push(const float* samples, uint64_t stream_stride, size_t frames, uint32_t channels, float sample_rate, const uint8_t* positions,
           omb_spectrum_bank_snapshot* out) {
    channels = std::min<uint32_t>(std::max<uint32_t>(channels, 1), OMB_MAX_CHANNELS);
    if (frames == 0) return OMB_NO_DATA;
    if (!samples || !out) return fail(OMB_ERR_INVALID, "null argument");
    if (stream_stride < (uint64_t)frames * channels) return fail(OMB_ERR_INVALID, "stream_stride smaller than one block");
    const float sr = sanitize_sample_rate(sample_rate);
    if (sr != config.sample_rate) {  // processor.rs:258-263
      config.sample_rate = sr;
      if (prepared) OMB_TRY(reset_buffers());
    }
    if (!prepared) {  // prepare -> rebuild_fft, processor.rs:120-136
      OMB_TRY(reset_buffers());
      prepared = true;
    }
    OMB_TRY(ensure());
    bool act[2];
    active(act);
    const uint32_t T = traces();
    if (T == 0) return OMB_NO_DATA;
    const uint32_t lanes = S * T;
    if (ring_lanes != lanes) {  // first use (the source selection is fixed for the life of a bank)
      begin = len = 0;
      cap = 0;
      ring_lanes = lanes;
    }

    // push_sources, processor.rs:271-298
    const size_t skip = (size_t)std::min<uint64_t>(pending_skip, frames);
    pending_skip -= skip;
    const size_t fresh = frames - skip;
    if (fresh) {
      OMB_TRY(make_room(fresh, lanes));
      float* tail = ring[cur].ptr + begin + len;
      const size_t block = frames * channels;
      OMB_TRY(d_block.reserve(block * S));
      OMB_CUDA_TRY(cudaMemcpy2DAsync(d_block.ptr, block * sizeof(float), samples, stream_stride * sizeof(float), block * sizeof(float), S,
                                     cudaMemcpyHostToDevice, stream));
      const StereoMatrix m = make_stereo_matrix(channels, positions);
      const int p0 = act[0] ? (int)config.source : (int)config.secondary;
      const int p1 = (int)config.secondary;
      const unsigned gx = (unsigned)std::min<uint64_t>((fresh + 255) / 256, 64);
      OMB_LAUNCH(k_downmix_spectrum_bank, dim3(gx, S), dim3(256), 0, stream, d_block.ptr, (uint64_t)block, (uint64_t)skip, (uint64_t)fresh,
                 (int)channels, m, p0, p1, (int)T, tail, (uint64_t)cap);
      OMB_CHECK_LAUNCH();
      len += fresh;
    }

    // process_ready_windows, processor.rs:179-213 — the same hop count for every stream
    const uint64_t N = config.fft_size, hop = config.hop, bins = config.bins();
    const uint64_t n = len >= N ? (len - N) / hop + 1 : 0;
    if (n == 0) return OMB_NO_DATA;
    if (begin % 4 != 0) OMB_TRY(relocate(len, lanes));  // the specialised kernel wants aligned lanes
    OMB_TRY(d_power.reserve((size_t)(n * bins * lanes)));
    OMB_TRY(d_w.reserve((size_t)(bins * lanes)));
    OMB_TRY(d_r.reserve((size_t)(bins * lanes)));
    OMB_TRY(h_w.reserve((size_t)(bins * lanes)));
    OMB_TRY(h_r.reserve((size_t)(bins * lanes)));
    OMB_TRY(plan->power_device(data(), lanes, n, cap, d_power.ptr, stream));
    float* state = config.averaging != OMB_AVG_NONE ? d_state.ptr : nullptr;
    OMB_TRY(plan->smooth_device(d_power.ptr, lanes, n, state, d_w.ptr, d_r.ptr, nullptr, false, stream));
    OMB_CUDA_TRY(cudaMemcpyAsync(h_w.ptr, d_w.ptr, bins * lanes * sizeof(float), cudaMemcpyDeviceToHost, stream));
    OMB_CUDA_TRY(cudaMemcpyAsync(h_r.ptr, d_r.ptr, bins * lanes * sizeof(float), cudaMemcpyDeviceToHost, stream));
    OMB_CUDA_TRY(cudaStreamSynchronize(stream));
    const uint64_t count = std::min<uint64_t>(n * hop, len);
    drain((size_t)count);
    pending_skip += n * hop - count;

    out->bins = (uint32_t)bins;
    out->n_streams = S;
    out->n_traces = T;
    out->trace_index[0] = act[0] ? 0 : 1;  // which reference trace (0 primary, 1 secondary) each bank trace is
    out->trace_index[1] = 1;
    out->frequency_bins = h_freq.data();
    out->weighted = h_w.ptr;
    out->raw = h_r.ptr;
    return OMB_OK;
  }
};

}  // namespace omb

using namespace omb;

struct omb_spectrum_bank {
  SpectrumBank b;
};

extern "C" {

int omb_spectrum_bank_create(const omb_spectrum_config* cfg, uint32_t n_streams, omb_spectrum_bank** out) {
  if (!cfg || !out || n_streams == 0) return fail(OMB_ERR_INVALID, "invalid argument");
  // the batched fold-down / smoothing launches address the stream (x trace) through gridDim.y
  if (n_streams > 32767u) return fail(OMB_ERR_UNSUPPORTED, "spectrum bank: at most 32767 streams per bank (got %u)", n_streams);
  try {
    auto* h = new omb_spectrum_bank();
    h->b.config = SpectrumConfigN::from_c(*cfg);
    h->b.S = n_streams;
    *out = h;
    return OMB_OK;
  } catch (const std::bad_alloc&) {
    return fail(OMB_ERR_NOMEM, "host allocation failed");
  }
}
void omb_spectrum_bank_destroy(omb_spectrum_bank* b) { delete b; }
int omb_spectrum_bank_reset_audio(omb_spectrum_bank* b) {
  if (!b) return fail(OMB_ERR_INVALID, "null argument");
  try {
    return b->b.reset_audio();
  } catch (...) {
    return fail(OMB_ERR_INVALID, "unexpected C++ exception");
  }
}
int omb_spectrum_bank_push(omb_spectrum_bank* b, const float* samples, uint64_t stream_stride, size_t frames, uint32_t channels,
                           float sample_rate, const uint8_t positions[OMB_MAX_CHANNELS], omb_spectrum_bank_snapshot* out) {
  if (!b) return fail(OMB_ERR_INVALID, "null argument");
  try {
    return b->b.push(samples, stream_stride, frames, channels, sample_rate, positions, out);
  } catch (const std::bad_alloc&) {
    return fail(OMB_ERR_NOMEM, "host allocation failed");
  } catch (...) {
    return fail(OMB_ERR_INVALID, "unexpected C++ exception");
  }
}
size_t omb_spectrum_bank_pending(const omb_spectrum_bank* b) { return b ? b->b.len : 0; }

}  // extern "C"
