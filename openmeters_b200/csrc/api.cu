// api.cu — the extern "C" surface declared in include/omb200.h.  No exceptions cross this boundary.
#include <new>

#include "streams.h"

using namespace omb;

struct omb_spectrogram { SpectrogramStream s; explicit omb_spectrogram(const omb_spectrogram_config& c) : s(c) {} };
struct omb_spectrum { SpectrumStream s; explicit omb_spectrum(const omb_spectrum_config& c) : s(c) {} };
struct omb_loudness { LoudnessStream s; explicit omb_loudness(const omb_loudness_config& c) : s(c) {} };
struct omb_stft_plan { StftPlan p; };
struct omb_spectrum_plan { SpectrumPlan p; };
struct omb_loudness_plan { LoudnessPlan p; };

#define OMB_GUARD_BEGIN try {
#define OMB_GUARD_END                                                   \
  }                                                                     \
  catch (const std::bad_alloc&) { return fail(OMB_ERR_NOMEM, "host allocation failed"); } \
  catch (...) { return fail(OMB_ERR_INVALID, "unexpected C++ exception"); }

extern "C" {

const char* omb_last_error(void) { return last_error_ref().c_str(); }
const char* omb_version(void) {
#ifdef OMB_EMU
  return "omb200 0.1.0 EMULATED (tests only)";
#else
  return "omb200 0.1.0 sm_100a";
#endif
}
int omb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
int omb_set_device(int device) {
  OMB_CUDA_TRY(cudaSetDevice(device));
  return OMB_OK;
}
uint64_t omb_kernel_launch_count(void) { return launch_count().load(); }
int omb_probe_fp32_tflops(double* out_tflops) {
  OMB_GUARD_BEGIN
  return probe_fp32_tflops(out_tflops);
  OMB_GUARD_END
}

// ---- multi-GPU ingest over peer memory (one process per GPU): CUDA IPC export / import of a device buffer
int omb_peer_alloc(size_t bytes, void** d_ptr, uint8_t handle[OMB_PEER_HANDLE_BYTES]) {
  OMB_GUARD_BEGIN
  if (!d_ptr || !handle || !bytes) return fail(OMB_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == OMB_PEER_HANDLE_BYTES, "CUDA IPC handle size");
  DeviceInfo dev;
  OMB_TRY(current_device(&dev));
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    cudaGetLastError();
    return fail(OMB_ERR_NOMEM, "cudaMalloc of %zu bytes failed", bytes);
  }
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    cudaGetLastError();
    return fail(OMB_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  std::memcpy(handle, &h, sizeof(h));
  *d_ptr = p;
  return OMB_OK;
  OMB_GUARD_END
}
int omb_peer_open(const uint8_t handle[OMB_PEER_HANDLE_BYTES], void** d_ptr) {
  OMB_GUARD_BEGIN
  if (!d_ptr || !handle) return fail(OMB_ERR_INVALID, "null argument");
  DeviceInfo dev;
  OMB_TRY(current_device(&dev));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(OMB_ERR_CUDA, "cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  }
  *d_ptr = p;
  return OMB_OK;
  OMB_GUARD_END
}
int omb_peer_close(void* d_ptr) {
  if (!d_ptr) return OMB_OK;
  OMB_CUDA_TRY(cudaIpcCloseMemHandle(d_ptr));
  return OMB_OK;
}
int omb_peer_free(void* d_ptr) {
  if (!d_ptr) return OMB_OK;
  OMB_CUDA_TRY(cudaFree(d_ptr));
  return OMB_OK;
}
int omb_copy_async(void* dst, const void* src, size_t bytes, void* cuda_stream) {
  if (!bytes) return OMB_OK;
  if (!dst || !src) return fail(OMB_ERR_INVALID, "null argument");
  OMB_CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(cuda_stream)));
  return OMB_OK;
}

// ---- plan set-up pieces
int omb_window_coefficients(int kind, size_t len, float* out) {
  OMB_GUARD_BEGIN
  if (!out && len) return fail(OMB_ERR_INVALID, "null out");
  const auto w = make_window(kind, len);
  std::copy(w.begin(), w.end(), out);
  return OMB_OK;
  OMB_GUARD_END
}
int omb_fft_bin_normalization(const float* window, size_t wlen, size_t fft_size, float* out) {
  OMB_GUARD_BEGIN
  if (!out) return fail(OMB_ERR_INVALID, "null out");
  const auto n = make_bin_norm(window, wlen, fft_size);
  std::copy(n.begin(), n.end(), out);
  return OMB_OK;
  OMB_GUARD_END
}
int omb_reassignment_windows(const float* window, size_t len, float* derivative, float* time_weighted) {
  OMB_GUARD_BEGIN
  if (!window || !derivative || !time_weighted) return fail(OMB_ERR_INVALID, "null argument");
  const auto d = make_derivative_window(window, len);
  const auto t = make_time_weighted_window(window, len);
  std::copy(d.begin(), d.end(), derivative);
  std::copy(t.begin(), t.end(), time_weighted);
  return OMB_OK;
  OMB_GUARD_END
}
float omb_reassigned_power_scale(const float* window, size_t len, size_t fft_size) { return make_power_scale(window, len, fft_size); }
uint16_t omb_pack_classic_db(float db) { return pack_classic_db_host(db); }
float omb_a_weight(float f) { return a_weight_host(f); }
int omb_k_weighting_coefficients(double fs, double* b, double* a) {
  if (!b || !a) return fail(OMB_ERR_INVALID, "null argument");
  k_weighting_host(fs, b, a);
  return OMB_OK;
}
int omb_true_peak_fir(int factor, float* out) {
  if (!out) return fail(OMB_ERR_INVALID, "null out");
  if (factor == 4) { true_peak_fir4_host(reinterpret_cast<float(*)[3]>(out)); return OMB_OK; }
  if (factor == 2) { true_peak_fir2_host(out); return OMB_OK; }
  return fail(OMB_ERR_INVALID, "factor must be 2 or 4");
}
void omb_fallback_positions(uint32_t channels, uint8_t positions[OMB_MAX_CHANNELS]) { fallback_positions_host(channels, positions); }
void omb_stereo_matrix(uint32_t channels, const uint8_t positions[OMB_MAX_CHANNELS], float out[OMB_MAX_CHANNELS][2]) {
  stereo_matrix_host(channels, positions, out);
}

int omb_downmix_project(const float* interleaved, size_t frames, uint32_t channels, const uint8_t positions[OMB_MAX_CHANNELS],
                        int channel, float* out_lane) {
  OMB_GUARD_BEGIN
  if (frames == 0) return OMB_OK;
  if (!interleaved || !out_lane) return fail(OMB_ERR_INVALID, "null argument");
  channels = std::min<uint32_t>(std::max<uint32_t>(channels, 1), OMB_MAX_CHANNELS);
  DeviceInfo dev;
  OMB_TRY(current_device(&dev));
  DeviceBuffer<float> d_in, d_out;
  OMB_TRY(d_in.upload(interleaved, frames * channels, nullptr));
  OMB_TRY(d_out.reserve(frames));
  const StereoMatrix m = make_stereo_matrix(channels, positions);
  OMB_TRY(launch_downmix(d_in.ptr, 0, frames, channels, m, channel, d_out.ptr, OMB_CHANNEL_NONE, nullptr, dev.sm_count, nullptr));
  OMB_CUDA_TRY(cudaMemcpy(out_lane, d_out.ptr, frames * sizeof(float), cudaMemcpyDeviceToHost));
  return OMB_OK;
  OMB_GUARD_END
}

// ---- spectrogram
void omb_spectrogram_default_config(omb_spectrogram_config* out) { if (out) StftConfig().to_c(out); }
int omb_spectrogram_create(const omb_spectrogram_config* cfg, omb_spectrogram** out) {
  OMB_GUARD_BEGIN
  if (!cfg || !out) return fail(OMB_ERR_INVALID, "null argument");
  *out = new omb_spectrogram(*cfg);
  return OMB_OK;
  OMB_GUARD_END
}
void omb_spectrogram_destroy(omb_spectrogram* h) { delete h; }
int omb_spectrogram_get_config(const omb_spectrogram* h, omb_spectrogram_config* out) {
  if (!h || !out) return fail(OMB_ERR_INVALID, "null argument");
  h->s.config.to_c(out);
  return OMB_OK;
}
int omb_spectrogram_update_config(omb_spectrogram* h, const omb_spectrogram_config* cfg) {
  OMB_GUARD_BEGIN
  if (!h || !cfg) return fail(OMB_ERR_INVALID, "null argument");
  return h->s.update_config(*cfg);
  OMB_GUARD_END
}
int omb_spectrogram_prepare(omb_spectrogram* h) {
  OMB_GUARD_BEGIN
  if (!h) return fail(OMB_ERR_INVALID, "null handle");
  return h->s.prepare();
  OMB_GUARD_END
}
int omb_spectrogram_reset_audio(omb_spectrogram* h) {
  if (!h) return fail(OMB_ERR_INVALID, "null handle");
  h->s.reset_audio();
  return OMB_OK;
}
int omb_spectrogram_process_block(omb_spectrogram* h, const float* samples, size_t n_samples, uint32_t channels, float sample_rate,
                                  const uint8_t positions[OMB_MAX_CHANNELS], omb_spectrogram_update* out) {
  OMB_GUARD_BEGIN
  if (!h) return fail(OMB_ERR_INVALID, "null handle");
  return h->s.process_block(samples, n_samples, channels, sample_rate, positions, out);
  OMB_GUARD_END
}

// ---- spectrum
void omb_spectrum_default_config(omb_spectrum_config* out) { if (out) SpectrumConfigN().to_c(out); }
int omb_spectrum_create(const omb_spectrum_config* cfg, omb_spectrum** out) {
  OMB_GUARD_BEGIN
  if (!cfg || !out) return fail(OMB_ERR_INVALID, "null argument");
  *out = new omb_spectrum(*cfg);
  return OMB_OK;
  OMB_GUARD_END
}
void omb_spectrum_destroy(omb_spectrum* h) { delete h; }
int omb_spectrum_get_config(const omb_spectrum* h, omb_spectrum_config* out) {
  if (!h || !out) return fail(OMB_ERR_INVALID, "null argument");
  h->s.config.to_c(out);
  return OMB_OK;
}
int omb_spectrum_update_config(omb_spectrum* h, const omb_spectrum_config* cfg) {
  OMB_GUARD_BEGIN
  if (!h || !cfg) return fail(OMB_ERR_INVALID, "null argument");
  return h->s.update_config(*cfg);
  OMB_GUARD_END
}
int omb_spectrum_prepare(omb_spectrum* h) {
  OMB_GUARD_BEGIN
  if (!h) return fail(OMB_ERR_INVALID, "null handle");
  return h->s.prepare();
  OMB_GUARD_END
}
int omb_spectrum_reset_audio(omb_spectrum* h) {
  OMB_GUARD_BEGIN
  if (!h) return fail(OMB_ERR_INVALID, "null handle");
  return h->s.reset_audio();
  OMB_GUARD_END
}
int omb_spectrum_process_block(omb_spectrum* h, const float* samples, size_t n_samples, uint32_t channels, float sample_rate,
                               const uint8_t positions[OMB_MAX_CHANNELS], omb_spectrum_snapshot* out) {
  OMB_GUARD_BEGIN
  if (!h) return fail(OMB_ERR_INVALID, "null handle");
  return h->s.process_block(samples, n_samples, channels, sample_rate, positions, out);
  OMB_GUARD_END
}

// ---- loudness
void omb_loudness_default_config(omb_loudness_config* out) {
  if (out) { out->sample_rate = kDefaultSampleRate; out->floor_db = -99.9f; }
}
int omb_loudness_create(const omb_loudness_config* cfg, omb_loudness** out) {
  OMB_GUARD_BEGIN
  if (!cfg || !out) return fail(OMB_ERR_INVALID, "null argument");
  *out = new omb_loudness(*cfg);
  return OMB_OK;
  OMB_GUARD_END
}
void omb_loudness_destroy(omb_loudness* h) { delete h; }
int omb_loudness_get_config(const omb_loudness* h, omb_loudness_config* out) {
  if (!h || !out) return fail(OMB_ERR_INVALID, "null argument");
  out->sample_rate = h->s.sample_rate;
  out->floor_db = h->s.cfg.floor_db;
  return OMB_OK;
}
int omb_loudness_reset_audio(omb_loudness* h) {
  OMB_GUARD_BEGIN
  if (!h) return fail(OMB_ERR_INVALID, "null handle");
  return h->s.reset_audio();
  OMB_GUARD_END
}
int omb_loudness_process_block(omb_loudness* h, const float* samples, size_t n_samples, uint32_t channels, float sample_rate,
                               const uint8_t positions[OMB_MAX_CHANNELS], omb_loudness_snapshot* out) {
  OMB_GUARD_BEGIN
  if (!h) return fail(OMB_ERR_INVALID, "null handle");
  return h->s.process_block(samples, n_samples, channels, sample_rate, positions, out);
  OMB_GUARD_END
}

// ---- loudness bank (row f1, device side)
struct omb_loudness_bank { LoudnessStream s; explicit omb_loudness_bank(const omb_loudness_config& c, uint32_t n) : s(c) { s.n_streams = n; } };
int omb_loudness_bank_create(const omb_loudness_config* cfg, uint32_t n_streams, omb_loudness_bank** out) {
  OMB_GUARD_BEGIN
  if (!cfg || !out) return fail(OMB_ERR_INVALID, "null argument");
  if (n_streams == 0 || n_streams > (1u << 20)) return fail(OMB_ERR_INVALID, "loudness bank: n_streams must be in 1..2^20");
  *out = new omb_loudness_bank(*cfg, n_streams);
  return OMB_OK;
  OMB_GUARD_END
}
void omb_loudness_bank_destroy(omb_loudness_bank* b) { delete b; }
int omb_loudness_bank_reset_audio(omb_loudness_bank* b) {
  OMB_GUARD_BEGIN
  if (!b) return fail(OMB_ERR_INVALID, "null handle");
  return b->s.reset_audio();
  OMB_GUARD_END
}
int omb_loudness_bank_push(omb_loudness_bank* b, const float* samples, uint64_t stream_stride, size_t n_samples, uint32_t channels,
                           float sample_rate, const uint8_t positions[OMB_MAX_CHANNELS], omb_loudness_snapshot* out_snapshots) {
  OMB_GUARD_BEGIN
  if (!b) return fail(OMB_ERR_INVALID, "null handle");
  if (b->s.n_streams > 1 && stream_stride < n_samples) return fail(OMB_ERR_INVALID, "loudness bank: stream_stride < n_samples");
  return b->s.process_bank(samples, stream_stride, n_samples, channels, sample_rate, positions, out_snapshots);
  OMB_GUARD_END
}

// ---- batched STFT
uint64_t omb_stft_frames_per_lane(const omb_spectrogram_config* cfg, uint64_t samples) {
  if (!cfg) return 0;
  return StftConfig::from_c(*cfg).frames_for(samples);
}
int omb_stft_plan_create(const omb_spectrogram_config* cfg, int kernel_choice, omb_stft_plan** out) {
  OMB_GUARD_BEGIN
  if (!cfg || !out) return fail(OMB_ERR_INVALID, "null argument");
  omb_stft_plan* p = new omb_stft_plan();
  const int rc = p->p.init(*cfg, kernel_choice);
  if (rc < 0) { delete p; return rc; }
  *out = p;
  return OMB_OK;
  OMB_GUARD_END
}
void omb_stft_plan_destroy(omb_stft_plan* p) { delete p; }
uint32_t omb_stft_plan_bins(const omb_stft_plan* p) { return p ? (uint32_t)p->p.cfg.bins() : 0; }
int omb_stft_plan_is_fast(const omb_stft_plan* p) { return p ? p->p.fast_kind : 0; }
float omb_stft_plan_power_scale(const omb_stft_plan* p) { return p ? p->p.power_scale : 0.0f; }
int omb_stft_execute_device(omb_stft_plan* p, const float* d_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride,
                            omb_spectrogram_point* d_out_points, uint64_t point_stride, uint32_t* d_out_counts,
                            uint16_t* d_out_classic, void* cuda_stream) {
  OMB_GUARD_BEGIN
  if (!p) return fail(OMB_ERR_INVALID, "null plan");
  return p->p.execute_device(d_lanes, n_lanes, samples_per_lane, lane_stride, d_out_points, point_stride, d_out_counts, d_out_classic,
                             (cudaStream_t)cuda_stream);
  OMB_GUARD_END
}
int omb_stft_execute_host(omb_stft_plan* p, const float* h_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride,
                          omb_spectrogram_point* h_out_points, uint64_t point_stride, uint32_t* h_out_counts, uint16_t* h_out_classic) {
  OMB_GUARD_BEGIN
  if (!p) return fail(OMB_ERR_INVALID, "null plan");
  return p->p.execute_host(h_lanes, n_lanes, samples_per_lane, lane_stride, h_out_points, point_stride, h_out_counts, h_out_classic);
  OMB_GUARD_END
}

int omb_stft_render_host(omb_stft_plan* p, const float* h_lanes, uint32_t n_lanes, uint64_t samples_per_lane, uint64_t lane_stride,
                         const omb_splat_params* view, float* h_db, uint32_t* h_out_counts) {
  OMB_GUARD_BEGIN
  if (!p || !view) return fail(OMB_ERR_INVALID, "null argument");
  return p->p.render_host(h_lanes, n_lanes, samples_per_lane, lane_stride, *view, h_db, h_out_counts);
  OMB_GUARD_END
}

// ---- batched spectrum
uint64_t omb_spectrum_hops_per_lane(const omb_spectrum_config* cfg, uint64_t samples) {
  if (!cfg) return 0;
  return SpectrumConfigN::from_c(*cfg).hops_for(samples);
}
int omb_spectrum_plan_create(const omb_spectrum_config* cfg, omb_spectrum_plan** out) {
  OMB_GUARD_BEGIN
  if (!cfg || !out) return fail(OMB_ERR_INVALID, "null argument");
  omb_spectrum_plan* p = new omb_spectrum_plan();
  const int rc = p->p.init(*cfg);
  if (rc < 0) { delete p; return rc; }
  *out = p;
  return OMB_OK;
  OMB_GUARD_END
}
void omb_spectrum_plan_destroy(omb_spectrum_plan* p) { delete p; }
int omb_spectrum_execute_device(omb_spectrum_plan* p, const float* d_lanes, uint32_t n_lanes, uint64_t samples_per_lane,
                                uint64_t lane_stride, float* d_out_weighted, float* d_out_raw, int32_t* d_out_peak_bin,
                                void* cuda_stream) {
  OMB_GUARD_BEGIN
  if (!p) return fail(OMB_ERR_INVALID, "null plan");
  return p->p.execute_device(d_lanes, n_lanes, samples_per_lane, lane_stride, d_out_weighted, d_out_raw, d_out_peak_bin,
                             (cudaStream_t)cuda_stream);
  OMB_GUARD_END
}
int omb_spectrum_execute_host(omb_spectrum_plan* p, const float* h_lanes, uint32_t n_lanes, uint64_t samples_per_lane,
                              uint64_t lane_stride, float* h_out_weighted, float* h_out_raw, int32_t* h_out_peak_bin) {
  OMB_GUARD_BEGIN
  if (!p) return fail(OMB_ERR_INVALID, "null plan");
  return p->p.execute_host(h_lanes, n_lanes, samples_per_lane, lane_stride, h_out_weighted, h_out_raw, h_out_peak_bin);
  OMB_GUARD_END
}

void omb_spectrum_default_peak_spec(omb_spectrum_peak_spec* out) {
  if (!out) return;
  out->trace = 0u;      // SpectrumWeightingMode::AWeighted (visuals.rs:98)
  out->min_hz = 20.0f;  // MIN_FREQUENCY (spectrum/state.rs:21)
  out->max_hz = 0.0f;   // frequency_bins[last].max(min * 1.02) (state.rs:107)
}
int omb_spectrum_plan_set_peak_spec(omb_spectrum_plan* p, const omb_spectrum_peak_spec* spec) {
  OMB_GUARD_BEGIN
  if (!p || !spec) return fail(OMB_ERR_INVALID, "null argument");
  return p->p.set_peak_spec(*spec);
  OMB_GUARD_END
}
int omb_spectrum_plan_get_peak_spec(const omb_spectrum_plan* p, omb_spectrum_peak_spec* out) {
  if (!p || !out) return fail(OMB_ERR_INVALID, "null argument");
  *out = p->p.peak_spec;
  return OMB_OK;
}
int omb_spectrum_interpolate_peaks_device(omb_spectrum_plan* p, const float* d_db, const int32_t* d_peak_bin, uint64_t rows,
                                          float* d_out_freq_hz, float* d_out_level_db, void* cuda_stream) {
  OMB_GUARD_BEGIN
  if (!p) return fail(OMB_ERR_INVALID, "null plan");
  return p->p.interpolate_peaks_device(d_db, d_peak_bin, rows, d_out_freq_hz, d_out_level_db, (cudaStream_t)cuda_stream);
  OMB_GUARD_END
}
int omb_spectrum_execute_host_peaks(omb_spectrum_plan* p, const float* h_lanes, uint32_t n_lanes, uint64_t samples_per_lane,
                                    uint64_t lane_stride, float* h_out_weighted, float* h_out_raw, int32_t* h_out_peak_bin,
                                    float* h_out_peak_freq_hz, float* h_out_peak_level_db) {
  OMB_GUARD_BEGIN
  if (!p) return fail(OMB_ERR_INVALID, "null plan");
  return p->p.execute_host(h_lanes, n_lanes, samples_per_lane, lane_stride, h_out_weighted, h_out_raw, h_out_peak_bin, h_out_peak_freq_hz,
                           h_out_peak_level_db);
  OMB_GUARD_END
}

// ---- batched loudness
int omb_loudness_plan_create(const omb_loudness_config* cfg, uint32_t channels, const uint8_t positions[OMB_MAX_CHANNELS],
                             omb_loudness_plan** out) {
  OMB_GUARD_BEGIN
  if (!cfg || !out) return fail(OMB_ERR_INVALID, "null argument");
  omb_loudness_plan* p = new omb_loudness_plan();
  const int rc = p->p.init(*cfg, channels, positions);
  if (rc < 0) { delete p; return rc; }
  *out = p;
  return OMB_OK;
  OMB_GUARD_END
}
void omb_loudness_plan_destroy(omb_loudness_plan* p) { delete p; }
int omb_loudness_execute_device(omb_loudness_plan* p, const float* d_interleaved, uint32_t n_streams, uint64_t frames,
                                uint64_t stream_stride, uint64_t block_frames, omb_loudness_snapshot* d_out, void* cuda_stream) {
  OMB_GUARD_BEGIN
  if (!p) return fail(OMB_ERR_INVALID, "null plan");
  return p->p.execute_device(d_interleaved, n_streams, frames, stream_stride, block_frames, d_out, (cudaStream_t)cuda_stream);
  OMB_GUARD_END
}
int omb_loudness_execute_host(omb_loudness_plan* p, const float* h_interleaved, uint32_t n_streams, uint64_t frames,
                              uint64_t stream_stride, uint64_t block_frames, omb_loudness_snapshot* h_out) {
  OMB_GUARD_BEGIN
  if (!p) return fail(OMB_ERR_INVALID, "null plan");
  return p->p.execute_host(h_interleaved, n_streams, frames, stream_stride, block_frames, h_out);
  OMB_GUARD_END
}

}  // extern "C"
