// tables.cu — host-side plan set-up. See tables.h for the reference rows each function replaces.
#include "tables.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace omb {

namespace {
constexpr double kPi = 3.14159265358979323846;
constexpr float kTauF = 6.28318530717958647692f;

// Set-up-only complex FFT (f32 like the reference's planner output; n is a power of two,
// otherwise an O(n^2) DFT with f64 accumulation). Not used on the data path.
void setup_fft(std::vector<std::complex<float>>& x, bool inverse) {
  const size_t n = x.size();
  if (n <= 1) return;
  const double sign = inverse ? 2.0 : -2.0;
  if (!is_pow2(n)) {
    std::vector<std::complex<float>> out(n);
    for (size_t k = 0; k < n; ++k) {
      std::complex<double> acc(0, 0);
      for (size_t j = 0; j < n; ++j) {
        const double a = sign * kPi * (double)((k * j) % n) / (double)n;
        acc += std::complex<double>(x[j]) * std::complex<double>(std::cos(a), std::sin(a));
      }
      out[k] = std::complex<float>((float)acc.real(), (float)acc.imag());
    }
    x.swap(out);
    return;
  }
  for (size_t i = 1, j = 0; i < n; ++i) {  // bit reversal
    size_t bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) std::swap(x[i], x[j]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const size_t half = len / 2;
    std::vector<std::complex<float>> w(half);
    for (size_t k = 0; k < half; ++k) {
      const double a = sign * kPi * (double)k / (double)len;
      w[k] = std::complex<float>((float)std::cos(a), (float)std::sin(a));
    }
    for (size_t base = 0; base < n; base += len)
      for (size_t k = 0; k < half; ++k) {
        const std::complex<float> t = x[base + k + half] * w[k];
        const std::complex<float> u = x[base + k];
        x[base + k] = u + t;
        x[base + k + half] = u - t;
      }
  }
}
}  // namespace

float sanitize_sample_rate(float sr) {
  const float v = (std::isfinite(sr) && sr > 0.0f) ? sr : kDefaultSampleRate;
  return std::min(std::max(v, 1.0f), kMaxSampleRate);
}

float sanitize_negative_db(float db, float dflt) { return (std::isfinite(db) && db < 0.0f) ? db : dflt; }

float db_to_power_host(float db) {
  const float k = 0.1f * 3.32192809488736234787f;
  return std::exp2(db * k);
}

std::vector<float> make_window(int kind, size_t len) {
  std::vector<float> w(len, 1.0f);
  if (len <= 1) return w;
  float c[4] = {0, 0, 0, 0};
  int terms = 0;
  switch (kind) {
    case OMB_WINDOW_HANN: c[0] = 0.5f; c[1] = -0.5f; terms = 2; break;
    case OMB_WINDOW_HAMMING: c[0] = 25.0f / 46.0f; c[1] = -21.0f / 46.0f; terms = 2; break;
    case OMB_WINDOW_BLACKMAN: c[0] = 0.42f; c[1] = -0.5f; c[2] = 0.08f; terms = 3; break;
    case OMB_WINDOW_BLACKMAN_HARRIS: c[0] = 0.35875f; c[1] = -0.48829f; c[2] = 0.14128f; c[3] = -0.01168f; terms = 4; break;
    default: return w;  // rectangular
  }
  const float step = kTauF / (float)len;
  for (size_t n = 0; n < len; ++n) {
    const float phi = (float)n * step;
    float acc = 0.0f;
    for (int k = 0; k < terms; ++k) {
      const float term = c[k] * std::cos(phi * (float)k);  // kept as separate mul + add (no contraction: -fmad=false for host code)
      acc = acc + term;
    }
    w[n] = acc;
  }
  return w;
}

std::vector<float> make_bin_norm(const float* window, size_t wlen, size_t fft_size) {
  const size_t bins = fft_size / 2 + 1;
  float sum = 0.0f;
  for (size_t i = 0; i < wlen; ++i) sum += window[i];
  float inv = 0.0f;
  if (std::fabs(sum) > std::numeric_limits<float>::epsilon()) inv = 1.0f / sum;
  else if (fft_size > 0) inv = 1.0f / (float)fft_size;
  const float dc = inv * inv;
  std::vector<float> norm(bins, 4.0f * dc);
  norm[0] = dc;
  if (fft_size % 2 == 0 && bins > 1) norm[bins - 1] = dc;
  return norm;
}

std::vector<float> make_derivative_window(const float* window, size_t n) {
  std::vector<float> out(n, 0.0f);
  if (n <= 1) return out;
  std::vector<std::complex<float>> buf(n);
  for (size_t i = 0; i < n; ++i) buf[i] = std::complex<float>(window[i], 0.0f);
  setup_fft(buf, false);
  const float scale = kTauF / (float)n;
  const size_t half = n / 2;
  buf[0] = 0;
  if (n % 2 == 0) buf[half] = 0;
  for (size_t k = 1; k < n; ++k) {
    const float omega = scale * ((float)k - (k > half ? (float)n : 0.0f));
    buf[k] = std::complex<float>(-omega * buf[k].imag(), omega * buf[k].real());
  }
  setup_fft(buf, true);
  const float inv_n = 1.0f / (float)n;
  for (size_t i = 0; i < n; ++i) out[i] = buf[i].real() * inv_n;
  return out;
}

std::vector<float> make_time_weighted_window(const float* window, size_t n) {
  std::vector<float> out(n);
  const float center = (float)(n ? n - 1 : 0) * 0.5f;
  for (size_t i = 0; i < n; ++i) out[i] = ((float)i - center) * window[i];
  return out;
}

float make_power_scale(const float* window, size_t n, size_t fft_size) {
  double s = 0.0, q = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double x = window[i];
    s += x;
    q += x * x;
  }
  return (float)(s * s / ((double)fft_size * q));
}

uint16_t pack_classic_db_host(float db) {
  const float scale = 65535.0f / kClassicDbRange;
  const float v = std::round((db - kClassicDbLo) * scale);
  return (uint16_t)std::min(std::max(v, 0.0f), 65535.0f);
}

float a_weight_host(float freq_hz) {
  if (freq_hz <= 0.0f) return -std::numeric_limits<float>::infinity();
  const double c1 = 20.598997 * 20.598997, c2 = 107.65265 * 107.65265;
  const double c3 = 737.86223 * 737.86223, c4 = 12194.217 * 12194.217;
  const double f2 = (double)freq_hz * (double)freq_hz;
  const double ra = (c4 * f2 * f2) / ((f2 + c1) * std::sqrt((f2 + c2) * (f2 + c3)) * (f2 + c4));
  return (float)(20.0 * std::log10(ra) + 2.0);
}

float smoothing_state_floor_host(const std::vector<float>& weighting_db, float floor_db) {
  float headroom = 0.0f;
  for (float w : weighting_db) headroom = std::fmax(headroom, w);
  return std::fmax(db_to_power_host(floor_db - headroom), std::numeric_limits<float>::min());
}

void k_weighting_host(double fs, double b[5], double a[5]) {
  // high-shelf stage
  const double f_shelf = 1681.974450955533, gain_db = 3.999843853973347, q_shelf = 0.7071752369554196;
  double k = std::tan(kPi * f_shelf / fs);
  const double vh = std::pow(10.0, gain_db / 20.0);
  const double vb = std::pow(vh, 0.4996667741545416);
  double a0 = 1.0 + k / q_shelf + k * k;
  const double sb[3] = {(vh + vb * k / q_shelf + k * k) / a0, 2.0 * (k * k - vh) / a0, (vh - vb * k / q_shelf + k * k) / a0};
  const double sa[3] = {1.0, 2.0 * (k * k - 1.0) / a0, (1.0 - k / q_shelf + k * k) / a0};
  // high-pass stage
  const double f_hp = 38.13547087602444, q_hp = 0.5003270373238773;
  k = std::tan(kPi * f_hp / fs);
  a0 = 1.0 + k / q_hp + k * k;
  const double hb[3] = {1.0, -2.0, 1.0};
  const double ha[3] = {1.0, 2.0 * (k * k - 1.0) / a0, (1.0 - k / q_hp + k * k) / a0};
  // polynomial product, written out in the reference's association order (processor.rs:45-53)
  auto poly = [](const double p[3], const double r[3], double o[5]) {
    o[0] = p[0] * r[0];
    o[1] = p[0] * r[1] + p[1] * r[0];
    o[2] = p[0] * r[2] + p[1] * r[1] + p[2] * r[0];
    o[3] = p[1] * r[2] + p[2] * r[1];
    o[4] = p[2] * r[2];
  };
  poly(sb, hb, b);
  poly(sa, ha, a);
}

static float true_peak_tap(int j, int factor) {
  const double taps = 48.0;
  const double x = ((double)j - taps * 0.5) * kPi / (double)factor;
  const double hann = 0.5 * (1.0 - std::cos(2.0 * kPi * (double)j / taps));
  return (float)(hann * std::sin(x) / x);
}
void true_peak_fir4_host(float out[12][3]) {
  for (int tap = 0; tap < 12; ++tap)
    for (int phase = 0; phase < 3; ++phase) out[tap][phase] = true_peak_tap(tap * 4 + phase + 1, 4);
}
void true_peak_fir2_host(float out[24]) {
  for (int tap = 0; tap < 24; ++tap) out[tap] = true_peak_tap(tap * 2 + 1, 2);
}

size_t loudness_window_length(float sample_rate, float secs) {
  const float len = sample_rate * secs;
  return len < 1.0f ? 1 : (size_t)len;
}

void fallback_positions_host(size_t channels, uint8_t pos[OMB_MAX_CHANNELS]) {
  channels = std::min<size_t>(channels, OMB_MAX_CHANNELS);
  for (size_t i = 0; i < OMB_MAX_CHANNELS; ++i) pos[i] = i < channels ? (uint8_t)i : (uint8_t)OMB_POS_UNKNOWN;
  switch (channels) {
    case 1: pos[0] = OMB_POS_MONO; break;
    case 4: pos[2] = OMB_POS_REAR_LEFT; pos[3] = OMB_POS_REAR_RIGHT; break;
    case 5: pos[3] = OMB_POS_REAR_LEFT; pos[4] = OMB_POS_REAR_RIGHT; break;
    default: break;
  }
}

void stereo_matrix_host(size_t channels, const uint8_t* pos, float m[OMB_MAX_CHANNELS][2]) {
  channels = std::min<size_t>(std::max<size_t>(channels, 1), OMB_MAX_CHANNELS);
  const float g = 0.70710678118654752440f;
  for (size_t i = 0; i < OMB_MAX_CHANNELS; ++i) m[i][0] = m[i][1] = 0.0f;
  bool has[2] = {false, false};
  for (size_t i = 0; i < channels; ++i) {
    float l = 0.0f, r = 0.0f;
    switch (pos[i]) {
      case OMB_POS_FRONT_LEFT: l = 1.0f; break;
      case OMB_POS_FRONT_RIGHT: r = 1.0f; break;
      case OMB_POS_FRONT_CENTER: l = r = g; break;
      case OMB_POS_REAR_LEFT: case OMB_POS_SIDE_LEFT: l = g; break;
      case OMB_POS_REAR_RIGHT: case OMB_POS_SIDE_RIGHT: r = g; break;
      case OMB_POS_MONO: l = r = 1.0f; break;
      default: break;
    }
    m[i][0] = l;
    m[i][1] = r;
    has[0] |= l != 0.0f;
    has[1] |= r != 0.0f;
  }
  if (has[0] && has[1]) return;
  if (!has[0] && !has[1]) {
    // dsp.rs:117-133 stereo_indices: explicit FL/FR, then Mono, then first free channels
    auto find = [&](uint8_t p) { for (size_t i = 0; i < channels; ++i) if (pos[i] == p) return (int)i; return -1; };
    const int explicit_right = find(OMB_POS_FRONT_RIGHT);
    int left = find(OMB_POS_FRONT_LEFT);
    if (left < 0) left = find(OMB_POS_MONO);
    if (left < 0) for (size_t i = 0; i < channels; ++i) if ((int)i != explicit_right) { left = (int)i; break; }
    if (left < 0) left = 0;
    int right = (explicit_right >= 0 && explicit_right != left) ? explicit_right : -1;
    if (right < 0) for (size_t i = 0; i < channels; ++i) if ((int)i != left) { right = (int)i; break; }
    if (right < 0) right = left;
    m[left][0] = 1.0f;
    m[right][1] = 1.0f;
    return;
  }
  const int from = has[0] ? 0 : 1, to = 1 - from;
  for (size_t i = 0; i < OMB_MAX_CHANNELS; ++i) m[i][to] = m[i][from];
}

std::vector<float2> make_twiddles(size_t n, size_t count) {
  std::vector<float2> t(count);
  for (size_t k = 0; k < count; ++k) {
    const double a = -2.0 * kPi * (double)k / (double)n;
    t[k] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  return t;
}

size_t history_columns(bool reassigned, uint32_t points, size_t requested) {
  const size_t req = std::min<size_t>(std::max<size_t>(requested, 1), 8192);
  const uint64_t stride = reassigned ? (uint64_t)points * 12u : ((uint64_t)points + 1) / 2 * 4;
  const size_t budget = (size_t)(128u * 1024u * 1024u) * (reassigned ? 2 : 1) / (size_t)std::max<uint64_t>(stride, 1);
  return std::min(req, budget);
}

}  // namespace omb
