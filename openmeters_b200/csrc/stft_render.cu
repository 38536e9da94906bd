// stft_render.cu — the reassigned STFT chained into the view's splat accumulation and resolve on the device
// (SURVEY.md §8 rows a8-a10 -> f2), so that only an IMAGE crosses PCIe instead of 24.6 KB of points per column.
//
// The reference does exactly this hand-over per frame tick: SpectrogramUpdate.new_columns are written into the point ring
// (spectrogram/render.rs:557-598), every point is drawn as a scale_factor-sized quad into the accumulation texture
// (render/shaders/spectrogram.wgsl:126-147,215-225) and fs_resolve turns it into dB (wgsl:227-237).  Here each lane is one
// "view" whose ring holds all of the lane's columns (ring_capacity = frames per lane, newest column = last frame): lane
// chunks are pipelined over three streams — H2D of the PCM, STFT kernel, k_splat_accumulate, k_splat_resolve, D2H of the
// dB image (and, if asked for, of the per-column point counts).  The points never leave HBM.
#include <algorithm>

#include "stft.h"

namespace omb {

int StftPlan::render_host(const float* h_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride,
                          const omb_splat_params& view_in, float* h_db, uint32_t* h_counts) {
  if (!cfg.reassign) return fail(OMB_ERR_INVALID, "render needs a reassigned plan (classic columns are drawn without splats)");
  const uint64_t frames = cfg.frames_for(samples_per_lane);
  if (frames == 0 || n_lanes == 0) return OMB_OK;
  if (!h_lanes || !h_db) return fail(OMB_ERR_INVALID, "null argument");
  if (frames > 0xffffffffull) return fail(OMB_ERR_UNSUPPORTED, "too many columns per lane");
  omb_splat_params view = view_in;
  view.ring_capacity = (uint32_t)frames;
  view.col_count = (uint32_t)frames;
  view.newest_col = (uint32_t)frames - 1;
  view.reassigned_power_scale = power_scale;
  uint32_t w = 0, h = 0;
  omb_splat_image_size(&view, &w, &h);
  const uint64_t pixels = (uint64_t)w * h;
  const uint64_t stride = cfg.bins();
  const uint64_t ds = (samples_per_lane + 3) & ~(uint64_t)3;
  OMB_CUDA_TRY(cudaSetDevice(dev.device));
  // lanes per chunk: pipeline depth 3; the call is bound by the H2D of the PCM, so what the chunking controls is the fill / drain
  // of the pipeline (first upload + last kernels and download are exposed): many small chunks (measured: 12 chunks 6.0 ms,
  // H2D alone 4.9 ms for 64 lanes x 2^20 samples).  The splat launch addresses (ring, slot) through gridDim.y (<= 65535).
  uint32_t chunk = std::max<uint32_t>(1, (n_lanes + 31) / 32);
  chunk = (uint32_t)std::min<uint64_t>(chunk, std::max<uint64_t>(1, 65535 / frames));
  if (frames > 65535) return fail(OMB_ERR_UNSUPPORTED, "more than 65535 columns per lane in one render call");
  OMB_TRY(d_in.reserve((size_t)(ds * n_lanes)));
  // points / counts / images of the three chunks in flight
  OMB_TRY(d_points.reserve((size_t)(3ull * chunk * frames * stride)));
  OMB_TRY(d_counts.reserve((size_t)(3ull * chunk * frames)));
  OMB_TRY(d_img_accum.reserve((size_t)(3ull * chunk * pixels)));
  OMB_TRY(d_img_db.reserve((size_t)(3ull * chunk * pixels)));
  OMB_CUDA_TRY(cudaStreamSynchronize(stream));
  for (auto& ps : pipe)
    if (!ps) OMB_CUDA_TRY(cudaStreamCreateWithFlags(&ps, cudaStreamNonBlocking));
  int k = 0;
  for (uint32_t l0 = 0; l0 < n_lanes; l0 += chunk, ++k) {
    const uint32_t nl = std::min(chunk, n_lanes - l0);
    const int slot = k % 3;
    cudaStream_t ps = pipe[slot];
    if (lane_stride == ds) {
      OMB_CUDA_TRY(cudaMemcpyAsync(d_in.ptr + (uint64_t)l0 * ds, h_lanes + (uint64_t)l0 * lane_stride,
                                   sizeof(float) * (ds * (nl - 1) + samples_per_lane), cudaMemcpyHostToDevice, ps));
    } else {
      for (uint32_t l = l0; l < l0 + nl; ++l)
        OMB_CUDA_TRY(cudaMemcpyAsync(d_in.ptr + (uint64_t)l * ds, h_lanes + (uint64_t)l * lane_stride, sizeof(float) * samples_per_lane,
                                     cudaMemcpyHostToDevice, ps));
    }
    omb_spectrogram_point* pts = d_points.ptr + (uint64_t)slot * chunk * frames * stride;
    uint32_t* cnt = d_counts.ptr + (uint64_t)slot * chunk * frames;
    float* acc = d_img_accum.ptr + (uint64_t)slot * chunk * pixels;
    float* db = d_img_db.ptr + (uint64_t)slot * chunk * pixels;
    OMB_TRY(execute_device(d_in.ptr + (uint64_t)l0 * ds, nl, samples_per_lane, ds, pts, stride, cnt, nullptr, ps));
    OMB_TRY(omb_splat_accumulate_device(pts, stride, cnt, nl, &view, acc, ps));
    OMB_TRY(omb_splat_resolve_device(acc, nl, &view, db, ps));
    OMB_CUDA_TRY(cudaMemcpyAsync(h_db + (uint64_t)l0 * pixels, db, sizeof(float) * nl * pixels, cudaMemcpyDeviceToHost, ps));
    if (h_counts)
      OMB_CUDA_TRY(cudaMemcpyAsync(h_counts + (uint64_t)l0 * frames, cnt, sizeof(uint32_t) * nl * frames, cudaMemcpyDeviceToHost, ps));
  }
  for (auto& ps : pipe) OMB_CUDA_TRY(cudaStreamSynchronize(ps));
  return OMB_OK;
}

}  // namespace omb
