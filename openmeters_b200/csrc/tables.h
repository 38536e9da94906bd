// tables.h — host-side plan set-up (SURVEY.md §8 rows a2, a3, a5, a14, a15, a18 and the
// dsp.rs channel-layout helpers).  One-off work per (config) — runs on the host, results are
// uploaded once and stay resident in HBM/L2 for the kernels.
#pragma once
#include <complex>
#include <vector>

#include "common.h"

namespace omb {

constexpr float kDefaultSampleRate = 48000.0f;
constexpr float kMaxSampleRate = 768000.0f;
constexpr float kDbFloor = -140.0f;            // util/audio/level.rs:4
constexpr float kLnToDb = 4.3429448f;          // util/audio/level.rs:5
constexpr float kAnalysisFloorPower = 1e-14f;  // spectrogram/processor.rs:69
constexpr float kClassicDbLo = -144.0f;        // spectrogram/processor.rs:66
constexpr float kClassicDbRange = 156.0f;      // spectrogram/processor.rs:66-68

float sanitize_sample_rate(float sr);                 // util/audio/rate.rs:6-13
float sanitize_negative_db(float db, float dflt);     // util/audio/level.rs:20-26
float db_to_power_host(float db);                     // util/audio/level.rs:36-39

std::vector<float> make_window(int kind, size_t len);                                         // window.rs:20-43
std::vector<float> make_bin_norm(const float* window, size_t wlen, size_t fft_size);         // window.rs:90-109
std::vector<float> make_derivative_window(const float* window, size_t n);                    // spectrogram/processor.rs:569-599
std::vector<float> make_time_weighted_window(const float* window, size_t n);                 // spectrogram/processor.rs:601-608
float make_power_scale(const float* window, size_t n, size_t fft_size);                      // spectrogram/processor.rs:111-117
uint16_t pack_classic_db_host(float db);                                                     // spectrogram/processor.rs:103-108
float a_weight_host(float freq_hz);                                                          // spectrum/processor.rs:410-425
float smoothing_state_floor_host(const std::vector<float>& weighting_db, float floor_db);    // spectrum/processor.rs:332-336
void k_weighting_host(double fs, double b[5], double a[5]);                                  // loudness/processor.rs:22-55
void true_peak_fir4_host(float out[12][3]);                                                  // loudness/processor.rs:79-97
void true_peak_fir2_host(float out[24]);
size_t loudness_window_length(float sample_rate, float secs);                                // loudness/processor.rs:68-71

void fallback_positions_host(size_t channels, uint8_t pos[OMB_MAX_CHANNELS]);                // dsp.rs:36-47
void stereo_matrix_host(size_t channels, const uint8_t* pos, float m[OMB_MAX_CHANNELS][2]);  // dsp.rs:117-176

// W_n^k = exp(-2*pi*i*k/n) for k < count, evaluated in f64 and rounded once.
std::vector<float2> make_twiddles(size_t n, size_t count);

// spectrogram/processor.rs:144-158
size_t history_columns(bool reassigned, uint32_t points, size_t requested);
static inline size_t hilbert_len_for(size_t window) { size_t h = (size_t)next_pow2(window * 2); return h < 2 ? 2 : h; }

}  // namespace omb
