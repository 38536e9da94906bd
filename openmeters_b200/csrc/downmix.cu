// downmix.cu — AudioBlock fold-down + projection.
//
// dsp.rs:223-257: every interleaved frame is folded to stereo in channel order starting
// from 0.0 (`left + sample * weight`, no FMA — the reference asserts the fold bits,
// dsp.rs:591-624), then projected (util/audio/channel.rs:12-21).  HBM-bound: one
// coalesced read of the frame, one or two coalesced lane writes.
#include "downmix.h"

namespace omb {

namespace {

__device__ __forceinline__ float project_dev(int channel, float l, float r) {
  switch (channel) {
    case OMB_CHANNEL_LEFT: return l;
    case OMB_CHANNEL_RIGHT: return r;
    case OMB_CHANNEL_MID: return __fmul_rn(__fadd_rn(l, r), 0.5f);
    case OMB_CHANNEL_SIDE: return __fmul_rn(__fsub_rn(l, r), 0.5f);
    default: return 0.0f;
  }
}

__global__ void __launch_bounds__(256) k_downmix(const float* __restrict__ in, uint64_t first_frame, uint64_t frames,
                                                 int channels, StereoMatrix m, int proj_a, int proj_b,
                                                 float* __restrict__ out_a, float* __restrict__ out_b) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; f < frames; f += stride) {
    const float* fr = in + (first_frame + f) * (uint64_t)channels;
    float l = 0.0f, r = 0.0f;
    for (int c = 0; c < channels; ++c) {
      const float s = __ldg(&fr[c]);
      l = __fadd_rn(l, __fmul_rn(s, m.w[c][0]));
      r = __fadd_rn(r, __fmul_rn(s, m.w[c][1]));
    }
    if (out_a) out_a[f] = project_dev(proj_a, l, r);
    if (out_b) out_b[f] = project_dev(proj_b, l, r);
  }
}

}  // namespace

StereoMatrix make_stereo_matrix(uint32_t channels, const uint8_t* positions) {
  StereoMatrix m;
  uint8_t fb[OMB_MAX_CHANNELS];
  if (!positions) {
    fallback_positions_host(channels, fb);
    positions = fb;
  }
  stereo_matrix_host(channels, positions, m.w);
  return m;
}

int launch_downmix(const float* d_interleaved, uint64_t first_frame, uint64_t frames, uint32_t channels,
                   const StereoMatrix& m, int proj_a, float* d_out_a, int proj_b, float* d_out_b, int sm_count,
                   cudaStream_t s) {
  if (frames == 0) return OMB_OK;
  const unsigned grid = (unsigned)std::min<uint64_t>((frames + 255) / 256, (uint64_t)std::max(sm_count, 1) * 8);
  OMB_LAUNCH(k_downmix, dim3(grid), dim3(256), 0, s, d_interleaved, first_frame, frames, (int)channels, m, proj_a, proj_b,
             d_out_a, d_out_b);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
