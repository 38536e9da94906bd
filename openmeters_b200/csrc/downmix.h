// downmix.h — AudioBlock fold-down (dsp.rs:135-257).
#pragma once
#include "common.h"
#include "tables.h"

namespace omb {

struct StereoMatrix { float w[OMB_MAX_CHANNELS][2]; };

// positions == nullptr => ChannelPosition::fallback(channels) (dsp.rs:36-47), as AudioBlock::new does.
StereoMatrix make_stereo_matrix(uint32_t channels, const uint8_t* positions);

// Folds frames [first_frame, first_frame+frames) of a device-resident interleaved block to up to two
// projected mono lanes (either output may be null).
int launch_downmix(const float* d_interleaved, uint64_t first_frame, uint64_t frames, uint32_t channels,
                   const StereoMatrix& m, int proj_a, float* d_out_a, int proj_b, float* d_out_b, int sm_count,
                   cudaStream_t s);

// Device-resident FIFO of mono samples (the reference's VecDeque<f32>): append at the tail, drain
// at the head, contiguous view for the batched kernels. Ping-pong storage, never overlapping copies.
struct DeviceLane {
  DeviceBuffer<float> buf[2];
  int cur = 0;
  size_t begin = 0, len = 0;
  const float* data() const { return buf[cur].ptr + begin; }
  float* tail() { return buf[cur].ptr + begin + len; }
  int make_room(size_t extra, cudaStream_t s) {
    if (begin + len + extra <= buf[cur].cap) return OMB_OK;
    DeviceBuffer<float>& other = buf[1 - cur];
    const size_t need = len + extra;
    if (other.cap < need) OMB_TRY(other.reserve(std::max<size_t>(need * 2, 8192)));
    if (len) OMB_CUDA_TRY(cudaMemcpyAsync(other.ptr, data(), len * sizeof(float), cudaMemcpyDeviceToDevice, s));
    cur = 1 - cur;
    begin = 0;
    return OMB_OK;
  }
  // Moves the pending samples to offset 0 of the other buffer so that data() is 16-byte aligned again.
  int realign(cudaStream_t s) {
    if (begin % 4 == 0) return OMB_OK;
    DeviceBuffer<float>& other = buf[1 - cur];
    if (other.cap < len) OMB_TRY(other.reserve(std::max<size_t>(len * 2, 8192)));
    if (len) OMB_CUDA_TRY(cudaMemcpyAsync(other.ptr, data(), len * sizeof(float), cudaMemcpyDeviceToDevice, s));
    cur = 1 - cur;
    begin = 0;
    return OMB_OK;
  }
  void commit(size_t n) { len += n; }
  void drain(size_t n) {
    n = std::min(n, len);
    begin += n;
    len -= n;
    if (len == 0) begin = 0;
  }
  void clear() { begin = len = 0; }
};

}  // namespace omb
