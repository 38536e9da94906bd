// splat.cu — row f2 of SURVEY.md §8: the step right after the reassigned STFT.  The reference draws every
// SpectrogramPoint as a scale_factor-sized quad, additively blended into a power accumulation texture, then resolves
// that texture to dB (spectrogram/render.rs:104-165, render/shaders/spectrogram.wgsl:126-147,215-237).  Here the
// accumulation is a scatter-add: one thread per point slot entry, f32 atomics into the image (RED.ADD.F32 in L2),
// the resolve an elementwise pass.  Images are in accumulation space (rotation, palette and the final clip
// transform are the GUI's).
#include <algorithm>
#include <cmath>

#include "common.h"
#include "tables.h"

namespace omb {

namespace {

constexpr float kLogKneeHz = 20.0f;                 // wgsl:8
constexpr float kLog10E = 0.4342944819f;            // wgsl:1
constexpr float kLnToDbWgsl = 4.342944819f;         // wgsl:2
constexpr float kDbToLog2 = 0.3321928095f;          // wgsl:3
constexpr float kAnalysisPowerEps = 1.0023052e-14f; // wgsl:20
constexpr float kDbAnalysisFloor = -140.0f;         // wgsl:16

struct SplatArgs {
  const omb_spectrogram_point* rings;
  const uint32_t* counts;
  uint64_t point_stride;
  uint32_t n_rings, slots, hl, newest;
  uint32_t freq_scale;
  float axis_lo, axis_inv;     // Uniforms.freq_axis (render.rs:212-225)
  float uv_y0, inv_uv_range;
  float ext_w, ext_h, sf, tilt_db;
  uint32_t width, height;
  float* accum;
};

__host__ __device__ inline float freq_scaled(uint32_t scale, float hz) {  // frequency.rs:25-31 / wgsl:63-72
  switch (scale) {
    case OMB_FREQ_LOG: return asinhf(hz / kLogKneeHz);
    case OMB_FREQ_ERB: return 21.4f * logf(1.0f + hz / 228.8f) * kLog10E;
    default: return hz;
  }
}

// One thread per (ring, slot, point index): grid.y = ring * slots + slot.
__global__ void __launch_bounds__(256) k_splat_accumulate(SplatArgs a) {
  const uint32_t rs = blockIdx.y;
  const uint32_t ring = rs / a.slots, slot = rs % a.slots;
  const uint32_t cnt = a.counts[(uint64_t)ring * a.hl + slot];
  const uint32_t count = cnt < (uint32_t)a.point_stride ? cnt : (uint32_t)a.point_stride;
  const uint32_t age = (a.newest + a.hl - slot) % a.hl;
  float* img = a.accum + (uint64_t)ring * a.width * a.height;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
    const omb_spectrogram_point pt = a.rings[((uint64_t)ring * a.hl + slot) * a.point_stride + i];
    float power = pt.power;
    const float zoomed = (__fmul_rn(freq_scaled(a.freq_scale, pt.freq_hz) - a.axis_lo, a.axis_inv) - a.uv_y0) * a.inv_uv_range;
    if (!(power > 0.0f) || zoomed < -0.01f || zoomed > 1.01f) continue;  // wgsl:136-138
    if (a.tilt_db != 0.0f) {                                             // wgsl:217-223
      if (!(power > kAnalysisPowerEps)) continue;
      if (pt.freq_hz > 0.0f) power *= exp2f(a.tilt_db * log2f(pt.freq_hz / 1000.0f) * kDbToLog2);
    }
    // wgsl:139-145: quad centre in accumulation pixels; the quad covers pixel centres in [c - sf/2, c + sf/2)
    const float px = a.ext_w - ((float)age - pt.time_offset) * a.sf;
    const float py = (1.0f - zoomed) * a.ext_h;
    const float h = 0.5f * a.sf;
    int x0 = (int)ceilf(px - h - 0.5f), x1 = (int)ceilf(px + h - 0.5f);
    int y0 = (int)ceilf(py - h - 0.5f), y1 = (int)ceilf(py + h - 0.5f);
    x0 = x0 < 0 ? 0 : x0;
    y0 = y0 < 0 ? 0 : y0;
    x1 = x1 > (int)a.width ? (int)a.width : x1;
    y1 = y1 > (int)a.height ? (int)a.height : y1;
    for (int y = y0; y < y1; ++y)
      for (int x = x0; x < x1; ++x) atomicAdd(&img[(uint64_t)y * a.width + x], power);
  }
}

__global__ void __launch_bounds__(256) k_splat_resolve(const float* accum, uint64_t n, float power_scale, float* db) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const float p = accum[i] * power_scale;  // wgsl:229-232 without the f16 low-power channel
    db[i] = p > 0.0f ? fmaxf(logf(fmaxf(p, 1e-20f)) * kLnToDbWgsl, kDbAnalysisFloor) : -INFINITY;
  }
}

int fill_args(const omb_splat_params& p, SplatArgs* a) {
  if (p.ring_capacity == 0) return fail(OMB_ERR_INVALID, "ring_capacity must be > 0");
  if (p.freq_scale > OMB_FREQ_ERB) return fail(OMB_ERR_INVALID, "unknown frequency scale %u", p.freq_scale);
  uint32_t w = 0, h = 0;
  omb_splat_image_size(&p, &w, &h);
  a->hl = p.ring_capacity;
  a->slots = std::min(p.col_count, p.ring_capacity);  // render.rs:113 visible_slots
  a->newest = p.newest_col % p.ring_capacity;
  a->freq_scale = p.freq_scale;
  const float lo = freq_scaled(p.freq_scale, p.freq_min), hi = freq_scaled(p.freq_scale, p.freq_max);
  a->axis_lo = lo;
  a->axis_inv = 1.0f / std::max(hi - lo, 1e-12f);
  a->uv_y0 = p.uv_y_range[0];
  a->inv_uv_range = 1.0f / std::max(p.uv_y_range[1] - p.uv_y_range[0], 1e-12f);
  a->sf = std::max(p.scale_factor, 1.0f);
  a->ext_w = p.ext_w;
  a->ext_h = p.ext_h;
  a->tilt_db = p.tilt_db;
  a->width = w;
  a->height = h;
  return OMB_OK;
}

}  // namespace

}  // namespace omb

using namespace omb;

extern "C" {

void omb_splat_image_size(const omb_splat_params* p, uint32_t* width, uint32_t* height) {
  // spectrogram.wgsl:146 `ceil(max(ext, 1))` == render.rs:522 for scale-multiplied bounds
  const float w = p ? std::ceil(std::max(p->ext_w, 1.0f)) : 1.0f, h = p ? std::ceil(std::max(p->ext_h, 1.0f)) : 1.0f;
  if (width) *width = (uint32_t)std::min(w, 65536.0f);
  if (height) *height = (uint32_t)std::min(h, 65536.0f);
}

int omb_splat_accumulate_device(const omb_spectrogram_point* d_rings, uint64_t point_stride, const uint32_t* d_slot_counts,
                                uint32_t n_rings, const omb_splat_params* p, float* d_accum, void* cuda_stream) {
  if (!p || !d_accum || (n_rings && (!d_rings || !d_slot_counts))) return fail(OMB_ERR_INVALID, "null argument");
  DeviceInfo dev;
  OMB_TRY(current_device(&dev));
  SplatArgs a{};
  OMB_TRY(fill_args(*p, &a));
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  const uint64_t pixels = (uint64_t)a.width * a.height * n_rings;
  OMB_CUDA_TRY(cudaMemsetAsync(d_accum, 0, sizeof(float) * pixels, s));  // LoadOp::Clear (render.rs:124)
  if (!n_rings || !a.slots || !point_stride) return OMB_OK;
  if ((uint64_t)n_rings * a.slots > 65535) return fail(OMB_ERR_UNSUPPORTED, "n_rings * slots = %llu exceeds one launch",
                                                       (unsigned long long)((uint64_t)n_rings * a.slots));
  a.rings = d_rings;
  a.counts = d_slot_counts;
  a.point_stride = point_stride;
  a.n_rings = n_rings;
  a.accum = d_accum;
  const unsigned bx = (unsigned)std::min<uint64_t>((point_stride + 255) / 256, 64);
  OMB_LAUNCH(k_splat_accumulate, dim3(bx, n_rings * a.slots), dim3(256), 0, s, a);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

int omb_splat_resolve_device(const float* d_accum, uint32_t n_rings, const omb_splat_params* p, float* d_db, void* cuda_stream) {
  if (!p || !d_accum || !d_db) return fail(OMB_ERR_INVALID, "null argument");
  DeviceInfo dev;
  OMB_TRY(current_device(&dev));
  uint32_t w = 0, h = 0;
  omb_splat_image_size(p, &w, &h);
  const uint64_t n = (uint64_t)w * h * n_rings;
  if (!n) return OMB_OK;
  const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)std::max(dev.sm_count, 1) * 16);
  OMB_LAUNCH(k_splat_resolve, dim3(grid), dim3(256), 0, static_cast<cudaStream_t>(cuda_stream), d_accum, n, p->reassigned_power_scale, d_db);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

int omb_splat_render_host(const omb_spectrogram_point* h_rings, uint64_t point_stride, const uint32_t* h_slot_counts,
                          uint32_t n_rings, const omb_splat_params* p, float* h_accum, float* h_db) {
  if (!p || !h_db || (n_rings && (!h_rings || !h_slot_counts))) return fail(OMB_ERR_INVALID, "null argument");
  DeviceInfo dev;
  OMB_TRY(current_device(&dev));
  uint32_t w = 0, h = 0;
  omb_splat_image_size(p, &w, &h);
  const uint64_t pixels = (uint64_t)w * h * n_rings;
  const uint64_t pts = (uint64_t)n_rings * p->ring_capacity * point_stride;
  DeviceBuffer<omb_spectrogram_point> d_pts;
  DeviceBuffer<uint32_t> d_cnt;
  DeviceBuffer<float> d_acc, d_db;
  OMB_TRY(d_pts.reserve((size_t)pts));
  OMB_TRY(d_cnt.reserve((size_t)n_rings * p->ring_capacity));
  OMB_TRY(d_acc.reserve((size_t)pixels));
  OMB_TRY(d_db.reserve((size_t)pixels));
  cudaStream_t s = nullptr;
  OMB_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  int rc = OMB_OK;
  auto run = [&]() -> int {
    if (pts) OMB_CUDA_TRY(cudaMemcpyAsync(d_pts.ptr, h_rings, sizeof(omb_spectrogram_point) * pts, cudaMemcpyHostToDevice, s));
    if (n_rings) OMB_CUDA_TRY(cudaMemcpyAsync(d_cnt.ptr, h_slot_counts, sizeof(uint32_t) * n_rings * p->ring_capacity, cudaMemcpyHostToDevice, s));
    OMB_TRY(omb_splat_accumulate_device(d_pts.ptr, point_stride, d_cnt.ptr, n_rings, p, d_acc.ptr, s));
    OMB_TRY(omb_splat_resolve_device(d_acc.ptr, n_rings, p, d_db.ptr, s));
    if (h_accum) OMB_CUDA_TRY(cudaMemcpyAsync(h_accum, d_acc.ptr, sizeof(float) * pixels, cudaMemcpyDeviceToHost, s));
    OMB_CUDA_TRY(cudaMemcpyAsync(h_db, d_db.ptr, sizeof(float) * pixels, cudaMemcpyDeviceToHost, s));
    OMB_CUDA_TRY(cudaStreamSynchronize(s));
    return OMB_OK;
  };
  rc = run();
  cudaStreamDestroy(s);
  return rc;
}

}  // extern "C"
