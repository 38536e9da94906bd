// omb_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A plain C++17 restatement of the reference's DSP hot path, used only as the
// checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// `--impl reference` leg.  Nothing under openmeters_b200/ may link, load or
// call it.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference/).  Arithmetic placement (f32 vs f64), summation order and
// thresholds follow the reference exactly; compile with -ffp-contract=off so
// no FMA contraction changes rounding (Rust never contracts).
//
// Parity pin: the FFT arithmetic of the reference lives in the un-vendored
// crates rustfft 6.4.1 / realfft 3.5.0 (Cargo.lock:2344-2350,2448-2459), so
// FFT *bits* are unpinned; the oracle uses its own f32 radix-2 FFT and is
// pinned semantically against every known-answer test the reference holds for
// this path (tests/test_oracle_kat.py restates them one by one).
//
// Exports use the `ombo_` prefix and the struct types of include/omb200.h so
// one ctypes harness can drive oracle and product alike.

#include "../include/omb200.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <deque>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <thread>
#include <vector>

namespace ombo {

using cf = std::complex<float>;
static constexpr float TAU_F = 6.28318530717958647692f;
static constexpr double PI_D = 3.14159265358979323846;

// ---------------------------------------------------------------------------
// util/audio/level.rs:4-39
// ---------------------------------------------------------------------------
static constexpr float DB_FLOOR = -140.0f;
static constexpr float LN_TO_DB = 4.3429448f;

static inline float power_to_db(float power, float floor) {  // level.rs:28-34
  return power > 0.0f ? std::max(std::log(power) * LN_TO_DB, floor) : floor;
}
static inline float db_to_power(float db) {  // level.rs:36-39
  const float DB_TO_LOG2 = 0.1f * 3.32192809488736234787f;
  return std::exp2(db * DB_TO_LOG2);
}
static inline float sanitize_negative_db(float db, float dflt) {  // level.rs:20-26
  return (std::isfinite(db) && db < 0.0f) ? db : dflt;
}
static inline void flush_denormal_f64(double& v) {  // level.rs:14-18
  if (std::fabs(v) < 1.0e-30) v = 0.0;
}

// util/audio/rate.rs:6-13
static constexpr float DEFAULT_SAMPLE_RATE = 48000.0f;
static constexpr float MAX_SAMPLE_RATE = 768000.0f;
static inline float sanitize_sample_rate(float sr) {
  float v = (std::isfinite(sr) && sr > 0.0f) ? sr : DEFAULT_SAMPLE_RATE;
  return std::min(std::max(v, 1.0f), MAX_SAMPLE_RATE);
}

// ---------------------------------------------------------------------------
// util/audio/window.rs:20-43 — periodic cosine-sum windows, all f32
// ---------------------------------------------------------------------------
static std::vector<float> window_coefficients(int kind, size_t len) {
  if (len <= 1) return std::vector<float>(len, 1.0f);
  static const float hann[] = {0.5f, -0.5f};
  static const float hamming[] = {25.0f / 46.0f, -21.0f / 46.0f};
  static const float blackman[] = {0.42f, -0.5f, 0.08f};
  static const float bh[] = {0.35875f, -0.48829f, 0.14128f, -0.01168f};
  const float* c = nullptr;
  int nc = 0;
  switch (kind) {
    case OMB_WINDOW_HANN: c = hann; nc = 2; break;
    case OMB_WINDOW_HAMMING: c = hamming; nc = 2; break;
    case OMB_WINDOW_BLACKMAN: c = blackman; nc = 3; break;
    case OMB_WINDOW_BLACKMAN_HARRIS: c = bh; nc = 4; break;
    default: return std::vector<float>(len, 1.0f);  // Rectangular
  }
  const float step = TAU_F / (float)len;
  std::vector<float> w(len);
  for (size_t n = 0; n < len; ++n) {
    const float phi = (float)n * step;
    float sum = 0.0f;
    for (int k = 0; k < nc; ++k) sum = sum + c[k] * std::cos(phi * (float)k);
    w[n] = sum;
  }
  return w;
}

// window.rs:90-109
static std::vector<float> compute_fft_bin_normalization(const float* window, size_t wlen, size_t fft_size) {
  const size_t bins = fft_size / 2 + 1;
  float window_sum = 0.0f;
  for (size_t i = 0; i < wlen; ++i) window_sum += window[i];
  float inv_sum;
  if (std::fabs(window_sum) > std::numeric_limits<float>::epsilon()) inv_sum = 1.0f / window_sum;
  else if (fft_size > 0) inv_sum = 1.0f / (float)fft_size;
  else inv_sum = 0.0f;
  const float dc = inv_sum * inv_sum;
  const float ac = 4.0f * dc;
  std::vector<float> norms(bins, ac);
  norms[0] = dc;
  if (fft_size % 2 == 0 && bins > 1) norms[bins - 1] = dc;
  return norms;
}

// window.rs:66-88 — mean is a sequential f32 sum over the frame.
static void copy_dc_removed_windowed(float* dst, const float* src, const float* window, size_t len) {
  if (len == 0) return;
  float sum = 0.0f;
  for (size_t i = 0; i < len; ++i) sum += src[i];
  const float mean = sum / (float)len;
  for (size_t i = 0; i < len; ++i) dst[i] = (src[i] - mean) * window[i];
}

// ---------------------------------------------------------------------------
// FFT — stands in for rustfft/realfft (third-party, absent).  Unnormalised,
// forward kernel e^{-j2πkn/n}.  f32 butterflies, twiddles rounded from f64.
// ---------------------------------------------------------------------------
struct FftPlan {
  size_t n = 0;
  bool pow2 = false;
  std::vector<uint32_t> rev;
  std::vector<cf> tw;  // W_n^k, k < n/2 (pow2) or k < n (generic)
};

static const FftPlan& fft_plan(size_t n) {
  static std::mutex mu;
  static std::map<size_t, std::unique_ptr<FftPlan>> cache;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(n);
  if (it != cache.end()) return *it->second;
  auto p = std::make_unique<FftPlan>();
  p->n = n;
  p->pow2 = n >= 1 && (n & (n - 1)) == 0;
  if (p->pow2) {
    int lg = 0;
    while ((size_t(1) << lg) < n) ++lg;
    p->rev.resize(n);
    for (size_t i = 0; i < n; ++i) {
      uint32_t r = 0;
      for (int b = 0; b < lg; ++b) if (i & (size_t(1) << b)) r |= 1u << (lg - 1 - b);
      p->rev[i] = r;
    }
    p->tw.resize(n / 2 + 1);
    for (size_t k = 0; k < n / 2 + 1; ++k) {
      const double a = -2.0 * PI_D * (double)k / (double)n;
      p->tw[k] = cf((float)std::cos(a), (float)std::sin(a));
    }
  } else {
    p->tw.resize(n);
    for (size_t k = 0; k < n; ++k) {
      const double a = -2.0 * PI_D * (double)k / (double)n;
      p->tw[k] = cf((float)std::cos(a), (float)std::sin(a));
    }
  }
  auto& ref = *p;
  cache.emplace(n, std::move(p));
  return ref;
}

static void fft_inplace(cf* x, size_t n, bool inverse, const FftPlan* plan = nullptr) {
  if (n <= 1) return;
  const FftPlan& p = plan ? *plan : fft_plan(n);  // hot loops pass a cached plan (the lookup takes a lock)
  if (!p.pow2) {  // tiny non power-of-two sizes only: direct DFT, f64 accumulation
    std::vector<cf> out(n);
    for (size_t k = 0; k < n; ++k) {
      double re = 0, im = 0;
      for (size_t j = 0; j < n; ++j) {
        cf w = p.tw[(k * j) % n];
        if (inverse) w = std::conj(w);
        re += (double)x[j].real() * w.real() - (double)x[j].imag() * w.imag();
        im += (double)x[j].real() * w.imag() + (double)x[j].imag() * w.real();
      }
      out[k] = cf((float)re, (float)im);
    }
    std::copy(out.begin(), out.end(), x);
    return;
  }
  for (size_t i = 0; i < n; ++i) {
    const size_t r = p.rev[i];
    if (i < r) std::swap(x[i], x[r]);
  }
  for (size_t len = 2; len <= n; len <<= 1) {
    const size_t half = len >> 1, step = n / len;
    for (size_t base = 0; base < n; base += len) {
      for (size_t k = 0; k < half; ++k) {
        cf w = p.tw[k * step];
        if (inverse) w = std::conj(w);
        const cf a = x[base + k], b = x[base + k + half];
        const cf t(b.real() * w.real() - b.imag() * w.imag(), b.real() * w.imag() + b.imag() * w.real());
        x[base + k] = a + t;
        x[base + k + half] = a - t;
      }
    }
  }
}

// realfft::RealToComplex stand-in: n real -> n/2+1 complex.
static void real_fft(const float* in, size_t n, cf* out, std::vector<cf>& scratch, const FftPlan* plan = nullptr) {
  scratch.resize(n);
  for (size_t i = 0; i < n; ++i) scratch[i] = cf(in[i], 0.0f);
  fft_inplace(scratch.data(), n, false, plan);
  for (size_t k = 0; k < n / 2 + 1; ++k) out[k] = scratch[k];
}

// ---------------------------------------------------------------------------
// dsp.rs:8-262 — channel positions, stereo fold-down, AudioBlock
// ---------------------------------------------------------------------------
static void fallback_positions(size_t channels, uint8_t pos[OMB_MAX_CHANNELS]) {  // dsp.rs:36-47
  channels = std::min<size_t>(channels, OMB_MAX_CHANNELS);
  for (int i = 0; i < OMB_MAX_CHANNELS; ++i) pos[i] = OMB_POS_UNKNOWN;
  for (size_t i = 0; i < channels; ++i) pos[i] = (uint8_t)i;  // SURROUND order
  if (channels == 1) pos[0] = OMB_POS_MONO;
  else if (channels == 4) { pos[2] = OMB_POS_REAR_LEFT; pos[3] = OMB_POS_REAR_RIGHT; }
  else if (channels == 5) { pos[3] = OMB_POS_REAR_LEFT; pos[4] = OMB_POS_REAR_RIGHT; }
}

static void stereo_indices(size_t channels, const uint8_t* pos, size_t out[2]) {  // dsp.rs:117-133
  auto find = [&](uint8_t p) -> std::optional<size_t> {
    for (size_t i = 0; i < channels; ++i) if (pos[i] == p) return i;
    return std::nullopt;
  };
  const auto explicit_right = find(OMB_POS_FRONT_RIGHT);
  std::optional<size_t> left = find(OMB_POS_FRONT_LEFT);
  if (!left) left = find(OMB_POS_MONO);
  if (!left) for (size_t i = 0; i < channels; ++i) if (!(explicit_right && *explicit_right == i)) { left = i; break; }
  const size_t l = left.value_or(0);
  std::optional<size_t> right;
  if (explicit_right && *explicit_right != l) right = explicit_right;
  if (!right) for (size_t i = 0; i < channels; ++i) if (i != l) { right = i; break; }
  out[0] = l;
  out[1] = right.value_or(l);
}

static void stereo_matrix(size_t channels, const uint8_t* pos, float m[OMB_MAX_CHANNELS][2]) {  // dsp.rs:135-176
  channels = std::min<size_t>(std::max<size_t>(channels, 1), OMB_MAX_CHANNELS);
  const float s = 0.70710678118654752440f;  // FRAC_1_SQRT_2
  for (int i = 0; i < OMB_MAX_CHANNELS; ++i) m[i][0] = m[i][1] = 0.0f;
  for (size_t i = 0; i < channels; ++i) {
    switch (pos[i]) {
      case OMB_POS_FRONT_LEFT: m[i][0] = 1.0f; break;
      case OMB_POS_FRONT_RIGHT: m[i][1] = 1.0f; break;
      case OMB_POS_FRONT_CENTER: m[i][0] = m[i][1] = s; break;
      case OMB_POS_REAR_LEFT: case OMB_POS_SIDE_LEFT: m[i][0] = s; break;
      case OMB_POS_REAR_RIGHT: case OMB_POS_SIDE_RIGHT: m[i][1] = s; break;
      case OMB_POS_MONO: m[i][0] = m[i][1] = 1.0f; break;
      default: break;  // LFE, Aux, Unknown
    }
  }
  auto populated = [&](int side) {
    for (size_t i = 0; i < channels; ++i) if (m[i][side] != 0.0f) return true;
    return false;
  };
  const bool pl = populated(0), pr = populated(1);
  if (!pl && !pr) {
    size_t idx[2];
    stereo_indices(channels, pos, idx);
    m[idx[0]][0] = 1.0f;
    m[idx[1]][1] = 1.0f;
  } else if (!pl && pr) {
    for (int i = 0; i < OMB_MAX_CHANNELS; ++i) m[i][0] = m[i][1];
  } else if (pl && !pr) {
    for (int i = 0; i < OMB_MAX_CHANNELS; ++i) m[i][1] = m[i][0];
  }
}

static inline uint32_t f32_bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

struct AudioBlock {  // dsp.rs:108-262
  const float* samples = nullptr;
  size_t len = 0;
  size_t channels = 1;
  float sample_rate = DEFAULT_SAMPLE_RATE;
  uint8_t positions[OMB_MAX_CHANNELS];
  float stereo[OMB_MAX_CHANNELS][2];
  size_t stereo_channels = 1;

  static AudioBlock with_positions(const float* s, size_t len, size_t channels, float sr, const uint8_t* pos) {
    AudioBlock b;
    b.samples = s;
    b.len = len;
    b.channels = std::min<size_t>(std::max<size_t>(channels, 1), OMB_MAX_CHANNELS);
    uint8_t fb[OMB_MAX_CHANNELS];
    if (!pos) { fallback_positions(b.channels, fb); pos = fb; }
    std::memcpy(b.positions, pos, OMB_MAX_CHANNELS);
    // dsp.rs:197-204: trailing channels whose samples are all zero *bits* are trimmed.
    size_t sc = std::min<size_t>(b.channels, 2);
    const size_t hi = std::min(b.channels, len);
    for (size_t ch = hi; ch-- > 2;) {
      bool any = false;
      for (size_t i = ch; i < len; i += b.channels) if (f32_bits(s[i]) != 0) { any = true; break; }
      if (any) { sc = ch + 1; break; }
    }
    b.stereo_channels = sc;
    b.sample_rate = sanitize_sample_rate(sr);
    stereo_matrix(b.channels, b.positions, b.stereo);
    return b;
  }
  size_t frame_count() const { return len / std::max<size_t>(channels, 1); }
  bool is_empty() const { return len < std::max<size_t>(channels, 1); }
  // dsp.rs:223-249: fold from 0.0 in channel order.
  void stereo_frame(size_t f, float out[2]) const {
    const float* fr = samples + f * channels;
    float l = 0.0f, r = 0.0f;
    for (size_t c = 0; c < stereo_channels; ++c) {
      l = l + fr[c] * stereo[c][0];
      r = r + fr[c] * stereo[c][1];
    }
    out[0] = l;
    out[1] = r;
  }
};

static inline float project(int channel, const float st[2]) {  // channel.rs:12-21
  switch (channel) {
    case OMB_CHANNEL_LEFT: return st[0];
    case OMB_CHANNEL_RIGHT: return st[1];
    case OMB_CHANNEL_MID: return (st[0] + st[1]) * 0.5f;
    case OMB_CHANNEL_SIDE: return (st[0] - st[1]) * 0.5f;
    default: return 0.0f;
  }
}

// ---------------------------------------------------------------------------
// spectrogram/processor.rs
// ---------------------------------------------------------------------------
static constexpr size_t DEFAULT_SPECTROGRAM_FFT_SIZE = 2048;
static constexpr size_t DEFAULT_SPECTROGRAM_HOP_SIZE = 64;
static constexpr size_t MAX_SPECTROGRAM_HISTORY_COLUMNS = 8192;
static constexpr size_t SPECTROGRAM_HISTORY_BYTE_BUDGET = 128u * 1024u * 1024u;
static constexpr float CLASSIC_DB_STORE_LO = -144.0f;
static constexpr float CLASSIC_DB_STORE_HI = 12.0f;
static constexpr float CLASSIC_DB_STORE_RANGE = CLASSIC_DB_STORE_HI - CLASSIC_DB_STORE_LO;
static constexpr float ANALYSIS_FLOOR_POWER = 1e-14f;

static inline uint16_t pack_classic_db(float db) {  // processor.rs:103-108
  const float SCALE = 65535.0f / CLASSIC_DB_STORE_RANGE;
  float v = std::round((db - CLASSIC_DB_STORE_LO) * SCALE);
  v = std::min(std::max(v, 0.0f), 65535.0f);
  return (uint16_t)v;
}

static float reassigned_power_scale(const float* w, size_t n, size_t fft_size) {  // processor.rs:111-117
  double sum = 0, sq = 0;
  for (size_t i = 0; i < n; ++i) { const double x = w[i]; sum += x; sq += x * x; }
  return (float)(sum * sum / ((double)fft_size * sq));
}

static size_t next_pow2(size_t v) { size_t p = 1; while (p < v) p <<= 1; return p; }
static size_t hilbert_len_for(size_t window) { return std::max<size_t>(next_pow2(window * 2), 2); }  // :225-227

static uint64_t col_byte_stride(bool reassigned, uint32_t points) {  // :144-151
  return reassigned ? (uint64_t)points * 12u : ((uint64_t)points + 1) / 2 * 4;
}
static size_t history_columns(bool reassigned, uint32_t points, size_t requested) {  // :153-158
  const size_t req = std::min(std::max<size_t>(requested, 1), MAX_SPECTROGRAM_HISTORY_COLUMNS);
  const size_t budget = SPECTROGRAM_HISTORY_BYTE_BUDGET * (1 + (reassigned ? 1 : 0)) /
                        (size_t)std::max<uint64_t>(col_byte_stride(reassigned, points), 1);
  return std::min(req, budget);
}

static std::vector<float> compute_derivative_spectral(const float* window, size_t n) {  // :569-599
  if (n <= 1) return std::vector<float>(n, 0.0f);
  std::vector<cf> buf(n);
  for (size_t i = 0; i < n; ++i) buf[i] = cf(window[i], 0.0f);
  fft_inplace(buf.data(), n, false);
  const float scale = TAU_F / (float)n;
  const size_t half = n / 2;
  buf[0] = cf(0, 0);
  if (n % 2 == 0) buf[half] = cf(0, 0);
  for (size_t k = 1; k < n; ++k) {
    const float omega = scale * ((float)k - (k > half ? (float)n : 0.0f));
    buf[k] = cf(-omega * buf[k].imag(), omega * buf[k].real());
  }
  fft_inplace(buf.data(), n, true);
  const float inv_n = 1.0f / (float)n;
  std::vector<float> out(n);
  for (size_t i = 0; i < n; ++i) out[i] = buf[i].real() * inv_n;
  return out;
}

static std::vector<float> compute_time_weighted(const float* window, size_t n) {  // :601-608
  const float center = (float)(n > 0 ? n - 1 : 0) * 0.5f;
  std::vector<float> out(n);
  for (size_t i = 0; i < n; ++i) out[i] = ((float)i - center) * window[i];
  return out;
}

struct SpectrogramConfig {
  float sample_rate = DEFAULT_SAMPLE_RATE;
  size_t fft_size = DEFAULT_SPECTROGRAM_FFT_SIZE;
  size_t hop_size = DEFAULT_SPECTROGRAM_HOP_SIZE;
  int window = OMB_WINDOW_HANN;
  size_t history_length = 0;
  bool use_reassignment = true;
  size_t zero_padding_factor = 1;
  void normalize() {  // :71-82
    sample_rate = sanitize_sample_rate(sample_rate);
    if (fft_size == 0) fft_size = DEFAULT_SPECTROGRAM_FFT_SIZE;
    if (hop_size == 0) hop_size = std::max<size_t>(std::min(DEFAULT_SPECTROGRAM_HOP_SIZE, fft_size), 1);
    zero_padding_factor = std::max<size_t>(zero_padding_factor, 1);
  }
};

// The per-column math shared by the streaming processor and the batch entry.
struct ColumnEngine {
  SpectrogramConfig cfg;
  size_t fft_size = 0;  // window * zp
  size_t hilbert_len = 0;
  size_t bins = 0;
  std::vector<float> window, dwin, twin, bin_norm;
  float power_scale = 1.0f;
  // scratch
  std::vector<cf> analytic, spectra, scratch, spec;
  std::vector<float> real;
  const FftPlan* plan_f = nullptr;  // cached transform plans (rustfft's Arc<dyn Fft>)
  const FftPlan* plan_h = nullptr;

  void rebuild(const SpectrogramConfig& c) {  // rebuild_fft :229-279 (buffers part)
    cfg = c;
    const size_t ws = cfg.fft_size;
    fft_size = ws * cfg.zero_padding_factor;
    hilbert_len = hilbert_len_for(ws);
    bins = fft_size / 2 + 1;
    plan_f = fft_size > 1 ? &fft_plan(fft_size) : nullptr;
    plan_h = hilbert_len > 1 ? &fft_plan(hilbert_len) : nullptr;
    spec.assign(bins, cf(0, 0));
    window = window_coefficients(cfg.window, ws);
    bin_norm = compute_fft_bin_normalization(window.data(), ws, fft_size);
    if (cfg.use_reassignment) {
      const float inv = 1.0f / (float)hilbert_len;
      for (auto& n : bin_norm) n *= inv * inv;
      dwin = compute_derivative_spectral(window.data(), ws);
      twin = compute_time_weighted(window.data(), ws);
      power_scale = reassigned_power_scale(window.data(), ws, fft_size);
      analytic.assign(hilbert_len, cf(0, 0));
      spectra.assign(fft_size * 3, cf(0, 0));
    } else {
      dwin.clear(); twin.clear();
      power_scale = 1.0f;
      real.assign(fft_size, 0.0f);
    }
  }
  size_t read_len() const { return cfg.use_reassignment ? hilbert_len : cfg.fft_size; }

  // :349-380 — frame points at `window` samples; out has `bins` codes.
  void classic_column(const float* frame, uint16_t* out) {
    const size_t ws = cfg.fft_size;
    copy_dc_removed_windowed(real.data(), frame, window.data(), ws);
    std::fill(real.begin() + ws, real.end(), 0.0f);
    real_fft(real.data(), fft_size, spec.data(), scratch, plan_f);
    for (size_t k = 0; k < bins; ++k) {
      const float p = (spec[k].real() * spec[k].real() + spec[k].imag() * spec[k].imag()) * bin_norm[k];
      out[k] = pack_classic_db(power_to_db(p, DB_FLOOR));
    }
  }

  // :318-348,439-488,546-567 — frame points at hilbert_len samples. Returns point count.
  size_t reassigned_column(const float* frame, omb_spectrogram_point* out) {
    const size_t ws = cfg.fft_size, H = hilbert_len, F = fft_size;
    const size_t center_offset = (H - ws) / 2;
    for (size_t i = 0; i < H; ++i) analytic[i] = cf(frame[i], 0.0f);
    // hilbert_transform :546-557
    fft_inplace(analytic.data(), H, false, plan_h);
    analytic[0] = cf(0, 0);
    for (size_t i = H / 2 + 1; i < H; ++i) analytic[i] = cf(0, 0);
    fft_inplace(analytic.data(), H, true, plan_h);
    const cf* a = analytic.data() + center_offset;
    cf* S = spectra.data();
    cf* D = S + F;
    cf* T = D + F;
    for (size_t i = 0; i < ws; ++i) {  // apply_complex_window :559-567
      S[i] = a[i] * window[i];
      D[i] = a[i] * dwin[i];
      T[i] = a[i] * twin[i];
    }
    for (size_t i = ws; i < F; ++i) S[i] = D[i] = T[i] = cf(0, 0);
    fft_inplace(S, F, false, plan_f);
    fft_inplace(D, F, false, plan_f);
    fft_inplace(T, F, false, plan_f);
    // reassigned_points :439-488
    const float sr = cfg.sample_rate;
    const float bin_hz = sr / (float)F;
    const float max_hz = sr * 0.5f;
    const float inv_2pi = sr / TAU_F;
    const float inv_hop = 1.0f / (float)cfg.hop_size;
    const float latency_hops = (float)center_offset * inv_hop;
    size_t n = 0;
    for (size_t i = 0; i < bins; ++i) {
      const cf b = S[i];
      const float pow = b.real() * b.real() + b.imag() * b.imag();
      const float scaled = pow * bin_norm[i];
      if (scaled < ANALYSIS_FLOOR_POWER) continue;
      const cf d = D[i], t = T[i];
      const float inv_pow = 1.0f / pow;
      const float d_omega = -(d.imag() * b.real() - d.real() * b.imag()) * inv_pow;
      const float freq = (float)i * bin_hz + d_omega * inv_2pi;
      if (!(freq > 0.0f && max_hz - freq > 0.0f)) continue;
      out[n].time_offset = (t.real() * b.real() + t.imag() * b.imag()) * inv_pow * inv_hop - latency_hops;
      out[n].freq_hz = freq;
      out[n].power = scaled;
      ++n;
    }
    return n;
  }
};

struct SpectrogramProcessor {
  SpectrogramConfig config;
  bool prepared = false;
  ColumnEngine eng;
  std::deque<float> audio;
  size_t pending_skip = 0;
  std::optional<size_t> last_nonzero;
  bool reset = true;
  // flattened output of the last process_block
  std::vector<uint32_t> offsets;
  std::vector<omb_spectrogram_point> points;
  std::vector<uint16_t> classic;

  explicit SpectrogramProcessor(SpectrogramConfig c) { c.normalize(); config = c; }

  void reset_audio() {  // :212-217
    audio.clear(); pending_skip = 0; last_nonzero.reset(); reset = true;
  }
  void prepare() { if (!prepared) rebuild_fft(); }  // :219-223
  void drain_audio(size_t count) {  // :397-404
    count = std::min(count, audio.size());
    if (count == 0) return;
    audio.erase(audio.begin(), audio.begin() + (ptrdiff_t)count);
    if (last_nonzero) { if (*last_nonzero >= count) last_nonzero = *last_nonzero - count; else last_nonzero.reset(); }
  }
  void advance_audio(size_t count) {  // :406-410
    const size_t missing = count > audio.size() ? count - audio.size() : 0;
    drain_audio(count);
    pending_skip += missing;
  }
  void rebuild_fft() {  // :229-279
    eng.rebuild(config);
    prepared = true;
    const size_t active_len = config.use_reassignment ? eng.hilbert_len : eng.fft_size;
    const size_t buffered = active_len * 2;
    drain_audio(audio.size() > buffered ? audio.size() - buffered : 0);
    pending_skip = 0;
  }
  void push_audio(const AudioBlock& b) {  // :412-437
    const size_t frames = b.frame_count();
    const size_t skip = std::min(pending_skip, frames);
    pending_skip -= skip;
    if (skip == frames) return;
    if (b.channels == 1) {
      const size_t base = audio.size();
      for (size_t i = frames; i-- > skip;) if (b.samples[i] != 0.0f) { last_nonzero = base + (i - skip); break; }
      audio.insert(audio.end(), b.samples + skip, b.samples + frames);
      return;
    }
    for (size_t f = skip; f < frames; ++f) {
      float st[2];
      b.stereo_frame(f, st);
      const float s = project(OMB_CHANNEL_MID, st);
      if (s != 0.0f) last_nonzero = audio.size();
      audio.push_back(s);
    }
  }
  size_t process_ready_windows() {  // :281-388, returns column count
    const size_t hop = config.hop_size;
    const bool re = config.use_reassignment;
    const size_t bins = eng.bins;
    const size_t read_len = eng.read_len();
    const size_t pending = audio.size();
    const size_t ready = pending >= read_len ? (pending - read_len) / hop + 1 : 0;
    const size_t retained = history_columns(re, (uint32_t)bins, config.history_length);
    const size_t skip = ready > retained ? ready - retained : 0;
    advance_audio(skip * hop);
    offsets.assign(1, 0);
    points.clear();
    classic.clear();
    std::vector<float> frame(read_len);
    std::vector<omb_spectrogram_point> col(bins);
    std::vector<uint16_t> ccol(bins);
    for (size_t it = skip; it < ready; ++it) {
      if (!last_nonzero) {  // :307-316
        if (re) offsets.push_back((uint32_t)points.size());
        else {
          classic.insert(classic.end(), bins, pack_classic_db(DB_FLOOR));
          offsets.push_back((uint32_t)classic.size());
        }
        advance_audio(hop);
        continue;
      }
      std::copy(audio.begin(), audio.begin() + (ptrdiff_t)read_len, frame.begin());
      if (re) {
        const size_t n = eng.reassigned_column(frame.data(), col.data());
        points.insert(points.end(), col.begin(), col.begin() + (ptrdiff_t)n);
        offsets.push_back((uint32_t)points.size());
      } else {
        eng.classic_column(frame.data(), ccol.data());
        classic.insert(classic.end(), ccol.begin(), ccol.end());
        offsets.push_back((uint32_t)classic.size());
      }
      advance_audio(hop);
    }
    return offsets.size() - 1;
  }
  // :490-516
  int process_block(const AudioBlock& b, omb_spectrogram_update* out) {
    if (b.is_empty()) return OMB_NO_DATA;
    if (config.sample_rate != b.sample_rate) {
      config.sample_rate = b.sample_rate;
      rebuild_fft();
      audio.clear();
      last_nonzero.reset();
      reset = true;
    }
    prepare();
    push_audio(b);
    const size_t cols = process_ready_windows();
    if (cols == 0) return OMB_NO_DATA;
    out->fft_size = eng.fft_size;
    out->hop_size = config.hop_size;
    out->history_length = config.history_length;
    out->sample_rate = config.sample_rate;
    out->reassigned_power_scale = eng.power_scale;
    out->reset = reset ? 1 : 0;
    reset = false;
    out->kind = config.use_reassignment ? OMB_COLUMN_REASSIGNED : OMB_COLUMN_CLASSIC;
    out->n_columns = (uint32_t)cols;
    out->bins = (uint32_t)eng.bins;
    out->column_offsets = offsets.data();
    out->points = points.data();
    out->classic_db = classic.data();
    return OMB_OK;
  }
  void update_config(SpectrogramConfig c) {  // :518-543
    c.normalize();
    const SpectrogramConfig prev = config;
    config = c;
    const bool rate_changed = prev.sample_rate != c.sample_rate;
    const bool rebuild = prev.fft_size != c.fft_size || prev.zero_padding_factor != c.zero_padding_factor ||
                         prev.window != c.window || prev.use_reassignment != c.use_reassignment || rate_changed;
    if (rebuild && prepared) {
      rebuild_fft();
      if (rate_changed) { audio.clear(); last_nonzero.reset(); }
    }
    const bool hop_changed = prev.hop_size != c.hop_size;
    if (hop_changed) pending_skip = 0;
    reset = reset || rebuild || hop_changed;
    // The column engine caches the config at rebuild() time; the reference reads self.config.hop_size LIVE in
    // process_ready_windows / reassigned_points (:283, :446-450), so a hop-only change (no rebuild) must reach it too.
    eng.cfg.hop_size = c.hop_size;
    eng.cfg.history_length = c.history_length;
  }
};

// ---------------------------------------------------------------------------
// spectrum/processor.rs
// ---------------------------------------------------------------------------
static float a_weight(float freq_hz) {  // :410-425
  const double C1 = 20.598997 * 20.598997, C2 = 107.65265 * 107.65265;
  const double C3 = 737.86223 * 737.86223, C4 = 12194.217 * 12194.217;
  if (freq_hz <= 0.0f) return -std::numeric_limits<float>::infinity();
  const double f = freq_hz, f2 = f * f;
  const double num = C4 * f2 * f2;
  const double den = (f2 + C1) * std::sqrt((f2 + C2) * (f2 + C3)) * (f2 + C4);
  return (float)(20.0 * std::log10(num / den) + 2.0);
}

static float smoothing_state_floor(const std::vector<float>& weighting, float floor) {  // :332-336
  float head = 0.0f;
  for (float w : weighting) head = std::max(head, w);  // f32::max ignores NaN like fmax; -inf never wins
  return std::max(db_to_power(floor - head), std::numeric_limits<float>::min());
}

struct SpectrumConfig {
  float sample_rate = DEFAULT_SAMPLE_RATE;
  size_t fft_size = 16384;
  size_t hop_size = 16384 / 16;
  int window = OMB_WINDOW_HANN;
  int averaging = OMB_AVG_NONE;
  float averaging_param = 0.0f;
  int source = OMB_CHANNEL_MID;
  int secondary_source = OMB_CHANNEL_NONE;
  float floor_db = -100.0f;
  void normalize() {  // :53-62
    sample_rate = sanitize_sample_rate(sample_rate);
    fft_size = std::max<size_t>(fft_size, 1);
    if (hop_size == 0) hop_size = std::max<size_t>(fft_size / 16, 1);
    floor_db = sanitize_negative_db(floor_db, -100.0f);
  }
};

struct SpectrumLevelBuffers {  // :325-403
  std::vector<float> smoothed, scratch_power;
  float state_floor = 0.0f;
  void reset(size_t bins, float sf, bool smoothing) {
    state_floor = sf;
    if (smoothing) smoothed.assign(bins, 0.0f); else smoothed.clear();
    scratch_power.assign(bins, 0.0f);
  }
  void clear() { smoothed.clear(); scratch_power.clear(); state_floor = 0.0f; }
  void update_outputs(int mode, float param, std::vector<float> out[2], const std::vector<float>& weighting,
                      float dt_seconds, float floor) {
    const size_t bins = scratch_power.size();
    for (int o = 0; o < 2; ++o) if (out[o].size() != bins) out[o].resize(bins, floor);
    const std::vector<float>* powers = &scratch_power;
    if (mode == OMB_AVG_EXPONENTIAL) {
      const float alpha = std::min(std::max(param, 0.0f), 0.9999f);
      for (size_t i = 0; i < bins; ++i) {
        float& avg = smoothed[i];
        const float p = scratch_power[i];
        avg = avg <= 0.0f ? p : avg * alpha + p * (1.0f - alpha);
        if (avg < state_floor) avg = 0.0f;
      }
      powers = &smoothed;
    } else if (mode == OMB_AVG_PEAK_HOLD) {
      const float decay = db_to_power(-std::max(param, 0.0f) * dt_seconds);
      for (size_t i = 0; i < bins; ++i) {
        float& hold = smoothed[i];
        hold = std::max(hold * decay, scratch_power[i]);
        if (hold < state_floor) hold = 0.0f;
      }
      powers = &smoothed;
    }
    std::vector<float>& weighted_out = out[0];
    std::vector<float>& raw_out = out[1];
    for (size_t i = 0; i < bins; ++i) {
      const float p = (*powers)[i];
      if (p < state_floor) { raw_out[i] = floor; weighted_out[i] = floor; continue; }
      const float db = std::log(p) * LN_TO_DB;
      raw_out[i] = std::max(db, floor);
      weighted_out[i] = std::max(db + weighting[i], floor);
    }
  }
};

struct SpectrumProcessor {
  SpectrumConfig config;
  bool prepared = false;
  std::vector<float> window, real, bin_norm, freq_bins, a_db;
  std::vector<cf> spec, scratch;
  const FftPlan* plan = nullptr;
  std::deque<float> pcm[2];
  size_t pending_skip = 0;
  SpectrumLevelBuffers levels[2];
  std::vector<float> traces[2][2];

  explicit SpectrumProcessor(SpectrumConfig c) { c.normalize(); config = c; }
  void active_traces(bool a[2]) const {  // :174-177
    a[0] = config.source != OMB_CHANNEL_NONE;
    a[1] = config.secondary_source != OMB_CHANNEL_NONE && config.secondary_source != config.source;
  }
  void reset_level_buffers() {  // :152-168
    const size_t bins = config.fft_size / 2 + 1;
    const float floor = config.floor_db;
    for (auto& t : traces) for (auto& v : t) v.assign(bins, floor);
    const float sf = smoothing_state_floor(a_db, floor);
    bool act[2];
    active_traces(act);
    const bool smoothing = config.averaging != OMB_AVG_NONE;
    for (int i = 0; i < 2; ++i) { if (act[i]) levels[i].reset(bins, sf, smoothing); else levels[i].clear(); }
  }
  void reset_buffers() {  // :138-150
    const size_t bins = config.fft_size / 2 + 1;
    const float bin_hz = config.sample_rate / (float)config.fft_size;
    freq_bins.resize(bins);
    a_db.resize(bins);
    for (size_t b = 0; b < bins; ++b) { const float f = (float)b * bin_hz; freq_bins[b] = f; a_db[b] = a_weight(f); }
    reset_level_buffers();
    pcm[0].clear(); pcm[1].clear();
    pending_skip = 0;
  }
  void rebuild_fft() {  // :126-136
    const size_t n = config.fft_size;
    window = window_coefficients(config.window, n);
    real.assign(n, 0.0f);
    spec.assign(n / 2 + 1, cf(0, 0));
    plan = n > 1 ? &fft_plan(n) : nullptr;
    prepared = true;
    bin_norm = compute_fft_bin_normalization(window.data(), n, n);
    reset_buffers();
  }
  void prepare() { if (!prepared) rebuild_fft(); }
  void reset_audio() {  // :112-118
    if (prepared) reset_level_buffers();
    pcm[0].clear(); pcm[1].clear();
    pending_skip = 0;
  }
  void process_trace_window(int trace, float dt, float floor) {  // :215-253
    const size_t n = config.fft_size;
    std::vector<float> frame(pcm[trace].begin(), pcm[trace].begin() + (ptrdiff_t)n);
    copy_dc_removed_windowed(real.data(), frame.data(), window.data(), n);
    real_fft(real.data(), n, spec.data(), scratch, plan);
    auto& lvl = levels[trace];
    for (size_t k = 0; k < spec.size(); ++k)
      lvl.scratch_power[k] = (spec[k].real() * spec[k].real() + spec[k].imag() * spec[k].imag()) * bin_norm[k];
    lvl.update_outputs(config.averaging, config.averaging_param, traces[trace], a_db, dt, floor);
  }
  bool process_ready_windows() {  // :179-213
    const size_t n = config.fft_size, hop = config.hop_size;
    const float floor = config.floor_db;
    const float dt = (float)hop / config.sample_rate;
    bool act[2];
    active_traces(act);
    bool produced = false;
    if (!act[0] && !act[1]) return false;
    for (;;) {
      bool ok = true;
      for (int t = 0; t < 2; ++t) if (act[t] && pcm[t].size() < n) ok = false;
      if (!ok) break;
      for (int t = 0; t < 2; ++t) if (act[t]) process_trace_window(t, dt, floor);
      size_t drained = hop;
      for (int t = 0; t < 2; ++t) if (act[t]) {
        const size_t count = std::min(hop, pcm[t].size());
        pcm[t].erase(pcm[t].begin(), pcm[t].begin() + (ptrdiff_t)count);
        drained = std::min(drained, count);
      }
      pending_skip += hop - drained;
      produced = true;
    }
    return produced;
  }
  void push_sources(const AudioBlock& b) {  // :271-298
    const size_t frames = b.frame_count();
    const size_t skip = std::min(pending_skip, frames);
    pending_skip -= skip;
    if (skip == frames) return;
    bool act[2];
    active_traces(act);
    const int src[2] = {config.source, config.secondary_source};
    for (size_t f = skip; f < frames; ++f) {
      float st[2];
      b.stereo_frame(f, st);
      for (int t = 0; t < 2; ++t) if (act[t]) pcm[t].push_back(project(src[t], st));
    }
  }
  int process_block(const AudioBlock& b, omb_spectrum_snapshot* out) {  // :255-269
    if (b.is_empty()) return OMB_NO_DATA;
    if (b.sample_rate != config.sample_rate) {
      config.sample_rate = b.sample_rate;
      if (prepared) reset_buffers();
    }
    prepare();
    push_sources(b);
    if (!process_ready_windows()) return OMB_NO_DATA;
    fill(out);
    return OMB_OK;
  }
  void fill(omb_spectrum_snapshot* out) {
    out->bins = (uint32_t)freq_bins.size();
    out->frequency_bins = freq_bins.data();
    for (int t = 0; t < 2; ++t) for (int w = 0; w < 2; ++w) out->traces[t][w] = traces[t][w].data();
  }
  void update_config(SpectrumConfig c) {  // :300-322
    const SpectrumConfig old = config;
    c.normalize();
    config = c;
    if (!prepared) return;
    const bool mode_changed = old.averaging != c.averaging;
    if (old.fft_size != c.fft_size || old.window != c.window) rebuild_fft();
    else if (old.sample_rate != c.sample_rate || old.hop_size != c.hop_size || old.source != c.source ||
             old.secondary_source != c.secondary_source) reset_buffers();
    else if (mode_changed || std::fabs(old.floor_db - c.floor_db) > std::numeric_limits<float>::epsilon())
      reset_level_buffers();
  }
};

// ---------------------------------------------------------------------------
// dsp.rs:264-371 — CompensatedPair / WindowedMeans<1,W>
// ---------------------------------------------------------------------------
struct CompensatedPair {
  double sums[2] = {0, 0}, corr[2] = {0, 0};
  void add(int i, double v) {  // Kahan-Babuska-Neumaier :277-285
    const double next = sums[i] + v;
    corr[i] += (std::fabs(sums[i]) >= std::fabs(v)) ? (sums[i] - next) + v : (v - next) + sums[i];
    sums[i] = next;
  }
  void refresh() { sums[0] = sums[1]; sums[1] = 0; corr[0] = corr[1]; corr[1] = 0; }
  double value() const { return sums[0] + corr[0]; }
};

struct WindowedMeans {  // VALUES = 1
  std::vector<double> buffer;
  std::vector<size_t> capacities, refresh_counts;
  std::vector<CompensatedPair> sums;
  size_t head = 0, count = 0;
  explicit WindowedMeans(const std::vector<size_t>& caps) {  // :311-322
    capacities = caps;
    size_t len = 1;
    for (auto& c : capacities) { c = std::max<size_t>(c, 1); len = std::max(len, c); }
    buffer.assign(len, 0.0);
    sums.assign(caps.size(), CompensatedPair());
    refresh_counts.assign(caps.size(), 0);
  }
  static WindowedMeans with_leading_zeros(const std::vector<size_t>& caps, size_t cnt) {  // :359-365
    WindowedMeans m(caps);
    m.head = cnt % m.buffer.size();
    m.count = std::min(cnt, m.buffer.size());
    for (size_t w = 0; w < caps.size(); ++w) m.refresh_counts[w] = cnt % m.capacities[w];
    return m;
  }
  void push(double value) {  // :324-357
    double mapped = value;
    if (!std::isfinite(value)) { value = 0.0; mapped = 0.0; }
    const size_t len = buffer.size();
    for (size_t w = 0; w < capacities.size(); ++w) {
      const size_t cap = capacities[w];
      const bool has_old = count >= cap;
      const double old = has_old ? buffer[(head + len - cap) % len] : 0.0;
      sums[w].add(0, mapped);
      sums[w].add(1, mapped);
      if (has_old) sums[w].add(0, -old);
      if (++refresh_counts[w] == cap) { sums[w].refresh(); refresh_counts[w] = 0; }
    }
    buffer[head] = value;
    head = (head + 1) % len;
    count = std::min(count + 1, len);
  }
  double mean(size_t w) const {  // :367-370
    const size_t c = std::max<size_t>(std::min(count, capacities[w]), 1);
    return sums[w].value() / (double)c;
  }
};

// ---------------------------------------------------------------------------
// loudness/processor.rs
// ---------------------------------------------------------------------------
struct KWeighting { double b[5], a[5]; };

static KWeighting k_weighting_coefficients(double fs) {  // :22-55
  KWeighting kw;
  double f0 = 1681.974450955533, g = 3.999843853973347, q = 0.7071752369554196;
  double k = std::tan(PI_D * f0 / fs);
  const double vh = std::pow(10.0, g / 20.0);
  const double vb = std::pow(vh, 0.4996667741545416);
  double a0 = 1.0 + k / q + k * k;
  const double pb[3] = {(vh + vb * k / q + k * k) / a0, 2.0 * (k * k - vh) / a0, (vh - vb * k / q + k * k) / a0};
  const double pa[3] = {1.0, 2.0 * (k * k - 1.0) / a0, (1.0 - k / q + k * k) / a0};
  f0 = 38.13547087602444; q = 0.5003270373238773;
  k = std::tan(PI_D * f0 / fs);
  a0 = 1.0 + k / q + k * k;
  const double rb[3] = {1.0, -2.0, 1.0};
  const double ra[3] = {1.0, 2.0 * (k * k - 1.0) / a0, (1.0 - k / q + k * k) / a0};
  auto conv = [](const double p[3], const double r[3], double o[5]) {
    o[0] = p[0] * r[0];
    o[1] = p[0] * r[1] + p[1] * r[0];
    o[2] = p[0] * r[2] + p[1] * r[1] + p[2] * r[0];
    o[3] = p[1] * r[2] + p[2] * r[1];
    o[4] = p[2] * r[2];
  };
  conv(pb, rb, kw.b);
  conv(pa, ra, kw.a);
  return kw;
}

static float mean_square_to_lufs(double ms, float floor) {  // :57-66
  if (ms > 0.0) return (float)std::max(std::fma(std::log10(ms), 10.0, -0.691), (double)floor);
  return floor;
}
static size_t window_length(float sr, float secs) {  // :68-71
  const float len = sr * secs;
  return len < 1.0f ? 1 : (size_t)len;
}

static constexpr int TRUE_PEAK_TAPS = 48;
static float true_peak_coefficient(int j, int factor) {  // :79-84
  const double offset = (double)j - TRUE_PEAK_TAPS * 0.5;
  const double window = 0.5 * (1.0 - std::cos(2.0 * PI_D * (double)j / (double)TRUE_PEAK_TAPS));
  const double x = offset * PI_D / (double)factor;
  return (float)(window * std::sin(x) / x);
}
struct TruePeakFirs {
  float fir4[12][3];
  float fir2[24];
  TruePeakFirs() {  // :90-97
    for (int tap = 0; tap < 12; ++tap) for (int ph = 0; ph < 3; ++ph) fir4[tap][ph] = true_peak_coefficient(tap * 4 + ph + 1, 4);
    for (int tap = 0; tap < 24; ++tap) fir2[tap] = true_peak_coefficient(tap * 2 + 1, 2);
  }
};
static const TruePeakFirs& firs() { static TruePeakFirs f; return f; }

struct TruePeakMeter {  // :99-151
  float delay[48] = {0};
  size_t write = 0, delay_len = 0;
  float peak = 0.0f;
  explicit TruePeakMeter(double sr) {
    delay_len = sr < 96000.0 ? 12 : (sr < 192000.0 ? 24 : 0);
    write = delay_len;
  }
  void process(float s) {
    peak = std::max(peak, std::fabs(s));
    if (delay_len == 0) return;
    write = (write == 0 ? delay_len : write) - 1;
    const size_t pos = write;
    delay[pos] = s;
    delay[pos + delay_len] = s;
    if (delay_len == 12) {
      float o[3] = {0, 0, 0};
      for (size_t i = 0; i < 12; ++i) {
        const float v = delay[pos + i];
        for (int ph = 0; ph < 3; ++ph) o[ph] += v * firs().fir4[i][ph];
      }
      for (int ph = 0; ph < 3; ++ph) peak = std::max(peak, std::fabs(o[ph]));
    } else {
      float o = 0;
      for (size_t i = 0; i < 24; ++i) o += delay[pos + i] * firs().fir2[i];
      peak = std::max(peak, std::fabs(o));
    }
  }
};

static inline float k_weighted(float sample, double st[4], const KWeighting& kw) {  // :153-162
  const double x = sample;
  const double y = kw.b[0] * x + st[0];
  st[0] = kw.b[1] * x + st[1] - kw.a[1] * y;
  st[1] = kw.b[2] * x + st[2] - kw.a[2] * y;
  st[2] = kw.b[3] * x + st[3] - kw.a[3] * y;
  st[3] = kw.b[4] * x - kw.a[4] * y;
  return (float)y;
}

static double channel_weight(uint8_t pos) {  // :174-183
  switch (pos) {
    case OMB_POS_LOW_FREQUENCY: return 0.0;
    case OMB_POS_REAR_LEFT: case OMB_POS_REAR_RIGHT: case OMB_POS_SIDE_LEFT: case OMB_POS_SIDE_RIGHT: return 1.41;
    default: return 1.0;
  }
}

struct LoudnessChannel {
  std::unique_ptr<WindowedMeans> windows;
  double filter[4] = {0, 0, 0, 0};
  std::unique_ptr<TruePeakMeter> tp;
  size_t silent_frames = 0;
  bool active() const { return (bool)windows; }
};

struct LoudnessProcessor {
  float cfg_sample_rate, floor_db;
  std::vector<LoudnessChannel> channels;
  KWeighting weighting;
  LoudnessProcessor(float sr, float floor) : cfg_sample_rate(sr), floor_db(floor) {  // :225-232
    weighting = k_weighting_coefficients((double)sanitize_sample_rate(sr));
  }
  void reset_audio() { for (auto& c : channels) c = LoudnessChannel(); }  // :234-236
  void ensure_state(size_t requested, float sr) {  // :238-251
    const size_t ch = std::min<size_t>(std::max<size_t>(requested, 1), OMB_MAX_CHANNELS);
    sr = sanitize_sample_rate(sr);
    const bool rate_changed = cfg_sample_rate != sr;
    if (rate_changed) { cfg_sample_rate = sr; weighting = k_weighting_coefficients((double)sr); }
    if (rate_changed || channels.size() != ch) { channels.clear(); channels.resize(ch); }
  }
  int process_block(const AudioBlock& b, omb_loudness_snapshot* out) {  // :253-311
    if (b.is_empty()) return OMB_NO_DATA;
    ensure_state(b.channels, b.sample_rate);
    static const float DEFAULT_WINDOWS[4] = {3.0f, 0.4f, 0.3f, 1.0f};
    std::vector<size_t> caps(4);
    for (int i = 0; i < 4; ++i) caps[i] = window_length(cfg_sample_rate, DEFAULT_WINDOWS[i]);
    const double sr = cfg_sample_rate;
    const size_t frames = b.frame_count();
    for (size_t f = 0; f < frames; ++f) {
      const float* fr = b.samples + f * b.channels;
      for (size_t c = 0; c < channels.size() && c < b.channels; ++c) {
        LoudnessChannel& ch = channels[c];
        const float s = fr[c];
        if (!ch.active()) {
          if (f32_bits(s) == 0) { ch.silent_frames += 1; continue; }
          ch.windows = std::make_unique<WindowedMeans>(WindowedMeans::with_leading_zeros(caps, ch.silent_frames));
          ch.tp = std::make_unique<TruePeakMeter>(sr);
          for (double& v : ch.filter) v = 0.0;
        }
        const double filtered = (double)k_weighted(s, ch.filter, weighting);
        ch.windows->push(filtered * filtered);
        ch.tp->process(s);
      }
    }
    for (auto& ch : channels) if (ch.active()) for (double& v : ch.filter) flush_denormal_f64(v);
    const float floor = floor_db;
    out->short_term_loudness = floor;
    out->momentary_loudness = floor;
    for (int i = 0; i < OMB_MAX_CHANNELS; ++i) {
      out->rms_fast_db[i] = out->rms_slow_db[i] = out->true_peak_db[i] = floor;
      out->positions[i] = OMB_POS_UNKNOWN;
    }
    double wst = 0.0, wm = 0.0;
    for (size_t c = 0; c < channels.size(); ++c) {
      LoudnessChannel& ch = channels[c];
      if (!ch.active()) continue;
      const double weight = channel_weight(b.positions[c]);
      wst += ch.windows->mean(0) * weight;
      wm += ch.windows->mean(1) * weight;
      out->rms_fast_db[c] = power_to_db((float)ch.windows->mean(2), floor);
      out->rms_slow_db[c] = power_to_db((float)ch.windows->mean(3), floor);
      const float peak = ch.tp->peak;
      ch.tp->peak = 0.0f;
      out->true_peak_db[c] = power_to_db(peak * peak, floor);
    }
    out->short_term_loudness = mean_square_to_lufs(wst, floor);
    out->momentary_loudness = mean_square_to_lufs(wm, floor);
    out->channel_count = (uint32_t)channels.size();
    std::memcpy(out->positions, b.positions, OMB_MAX_CHANNELS);
    return OMB_OK;
  }
};

// ---------------------------------------------------------------------------
// helpers for the C layer
// ---------------------------------------------------------------------------
static SpectrogramConfig from_c(const omb_spectrogram_config& c) {
  SpectrogramConfig o;
  o.sample_rate = c.sample_rate; o.fft_size = (size_t)c.fft_size; o.hop_size = (size_t)c.hop_size;
  o.window = (int)c.window; o.history_length = (size_t)c.history_length;
  o.use_reassignment = c.use_reassignment != 0; o.zero_padding_factor = (size_t)c.zero_padding_factor;
  return o;
}
static void to_c(const SpectrogramConfig& c, omb_spectrogram_config* o) {
  std::memset(o, 0, sizeof *o);
  o->sample_rate = c.sample_rate; o->fft_size = c.fft_size; o->hop_size = c.hop_size; o->window = (uint32_t)c.window;
  o->history_length = c.history_length; o->use_reassignment = c.use_reassignment ? 1 : 0;
  o->zero_padding_factor = c.zero_padding_factor;
}
static SpectrumConfig from_c(const omb_spectrum_config& c) {
  SpectrumConfig o;
  o.sample_rate = c.sample_rate; o.fft_size = (size_t)c.fft_size; o.hop_size = (size_t)c.hop_size;
  o.window = (int)c.window; o.averaging = (int)c.averaging; o.averaging_param = c.averaging_param;
  o.source = (int)c.source; o.secondary_source = (int)c.secondary_source; o.floor_db = c.floor_db;
  return o;
}
static void to_c(const SpectrumConfig& c, omb_spectrum_config* o) {
  std::memset(o, 0, sizeof *o);
  o->sample_rate = c.sample_rate; o->fft_size = c.fft_size; o->hop_size = c.hop_size; o->window = (uint32_t)c.window;
  o->averaging = (uint32_t)c.averaging; o->averaging_param = c.averaging_param; o->source = (uint32_t)c.source;
  o->secondary_source = (uint32_t)c.secondary_source; o->floor_db = c.floor_db;
}

static int hw_threads() {
  const unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

template <class F>
static void parallel_for(size_t n, int threads, F&& fn) {
  if (threads <= 0) threads = hw_threads();
  threads = (int)std::min<size_t>((size_t)threads, std::max<size_t>(n, 1));
  if (threads <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t)
    pool.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); });
  for (auto& th : pool) th.join();
}

}  // namespace ombo

// ===========================================================================
// C exports (ombo_ prefix)
// ===========================================================================
using namespace ombo;

struct ombo_spectrogram { SpectrogramProcessor p; explicit ombo_spectrogram(SpectrogramConfig c) : p(c) {} };
struct ombo_spectrum { SpectrumProcessor p; explicit ombo_spectrum(SpectrumConfig c) : p(c) {} };
struct ombo_loudness { LoudnessProcessor p; omb_loudness_config cfg; ombo_loudness(float sr, float fl) : p(sr, fl) {} };

extern "C" {

const char* ombo_version(void) { return "omb200-oracle 0.1 (cpu restatement)"; }
int ombo_hw_threads(void) { return hw_threads(); }

int ombo_window_coefficients(int kind, size_t len, float* out) {
  auto w = window_coefficients(kind, len);
  std::copy(w.begin(), w.end(), out);
  return OMB_OK;
}
int ombo_fft_bin_normalization(const float* window, size_t wlen, size_t fft_size, float* out) {
  auto n = compute_fft_bin_normalization(window, wlen, fft_size);
  std::copy(n.begin(), n.end(), out);
  return OMB_OK;
}
int ombo_reassignment_windows(const float* window, size_t len, float* derivative, float* time_weighted) {
  auto d = compute_derivative_spectral(window, len);
  auto t = compute_time_weighted(window, len);
  std::copy(d.begin(), d.end(), derivative);
  std::copy(t.begin(), t.end(), time_weighted);
  return OMB_OK;
}
float ombo_reassigned_power_scale(const float* w, size_t len, size_t fft_size) { return reassigned_power_scale(w, len, fft_size); }
uint16_t ombo_pack_classic_db(float db) { return pack_classic_db(db); }
float ombo_power_to_db(float p, float floor) { return power_to_db(p, floor); }
float ombo_db_to_power(float db) { return db_to_power(db); }
float ombo_a_weight(float f) { return a_weight(f); }
float ombo_sanitize_sample_rate(float sr) { return sanitize_sample_rate(sr); }
int ombo_k_weighting_coefficients(double fs, double* b, double* a) {
  const KWeighting kw = k_weighting_coefficients(fs);
  std::copy(kw.b, kw.b + 5, b);
  std::copy(kw.a, kw.a + 5, a);
  return OMB_OK;
}
int ombo_true_peak_fir(int factor, float* out) {
  if (factor == 4) { std::memcpy(out, firs().fir4, sizeof(float) * 36); return OMB_OK; }
  if (factor == 2) { std::memcpy(out, firs().fir2, sizeof(float) * 24); return OMB_OK; }
  return OMB_ERR_INVALID;
}
void ombo_fallback_positions(uint32_t channels, uint8_t positions[OMB_MAX_CHANNELS]) { fallback_positions(channels, positions); }
void ombo_stereo_matrix(uint32_t channels, const uint8_t positions[OMB_MAX_CHANNELS], float out[OMB_MAX_CHANNELS][2]) {
  stereo_matrix(channels, positions, out);
}
size_t ombo_history_columns(int reassigned, uint32_t points, size_t requested) { return history_columns(reassigned != 0, points, requested); }

// complex FFT exposed for tests (interleaved re,im)
int ombo_fft(float* interleaved, size_t n, int inverse) {
  fft_inplace(reinterpret_cast<cf*>(interleaved), n, inverse != 0);
  return OMB_OK;
}

// WindowedMeans test helper (dsp.rs:626-656, loudness/processor.rs:323-336): push `n`
// values through a WindowedMeans with `nw` windows and report the means.
int ombo_windowed_means(const size_t* capacities, size_t nw, size_t leading_zeros, const double* values, size_t n, double* means_out) {
  std::vector<size_t> caps(capacities, capacities + nw);
  WindowedMeans m = leading_zeros ? WindowedMeans::with_leading_zeros(caps, leading_zeros) : WindowedMeans(caps);
  for (size_t i = 0; i < n; ++i) m.push(values[i]);
  for (size_t w = 0; w < nw; ++w) means_out[w] = m.mean(w);
  return OMB_OK;
}

int ombo_downmix_project(const float* interleaved, size_t frames, uint32_t channels, const uint8_t positions[OMB_MAX_CHANNELS],
                         int channel, float* out_lane) {
  AudioBlock b = AudioBlock::with_positions(interleaved, frames * std::max<uint32_t>(channels, 1), channels, 48000.0f, positions);
  for (size_t f = 0; f < frames; ++f) { float st[2]; b.stereo_frame(f, st); out_lane[f] = project(channel, st); }
  return OMB_OK;
}
// both stereo lanes (dsp.rs:223-249), out[frames][2]
int ombo_stereo_frames(const float* interleaved, size_t frames, uint32_t channels, const uint8_t positions[OMB_MAX_CHANNELS], float* out) {
  AudioBlock b = AudioBlock::with_positions(interleaved, frames * std::max<uint32_t>(channels, 1), channels, 48000.0f, positions);
  for (size_t f = 0; f < frames; ++f) b.stereo_frame(f, out + 2 * f);
  return OMB_OK;
}

// ---- spectrogram
void ombo_spectrogram_default_config(omb_spectrogram_config* out) { to_c(SpectrogramConfig(), out); }
int ombo_spectrogram_create(const omb_spectrogram_config* cfg, ombo_spectrogram** out) {
  if (!cfg || !out) return OMB_ERR_INVALID;
  *out = new ombo_spectrogram(from_c(*cfg));
  return OMB_OK;
}
void ombo_spectrogram_destroy(ombo_spectrogram* h) { delete h; }
int ombo_spectrogram_get_config(const ombo_spectrogram* h, omb_spectrogram_config* out) { to_c(h->p.config, out); return OMB_OK; }
int ombo_spectrogram_update_config(ombo_spectrogram* h, const omb_spectrogram_config* cfg) { h->p.update_config(from_c(*cfg)); return OMB_OK; }
int ombo_spectrogram_prepare(ombo_spectrogram* h) { h->p.prepare(); return OMB_OK; }
int ombo_spectrogram_reset_audio(ombo_spectrogram* h) { h->p.reset_audio(); return OMB_OK; }
int ombo_spectrogram_process_block(ombo_spectrogram* h, const float* samples, size_t n_samples, uint32_t channels, float sample_rate,
                                   const uint8_t positions[OMB_MAX_CHANNELS], omb_spectrogram_update* out) {
  AudioBlock b = AudioBlock::with_positions(samples, n_samples, channels, sample_rate, positions);
  return h->p.process_block(b, out);
}
// test hook: pending audio length (fft_rebuild_keeps_newest_pending_audio, processor.rs:794-805)
size_t ombo_spectrogram_pending(const ombo_spectrogram* h, float* out, size_t cap) {
  size_t n = h->p.audio.size();
  for (size_t i = 0; i < n && i < cap; ++i) out[i] = h->p.audio[i];
  return n;
}

// ---- spectrum
void ombo_spectrum_default_config(omb_spectrum_config* out) { to_c(SpectrumConfig(), out); }
int ombo_spectrum_create(const omb_spectrum_config* cfg, ombo_spectrum** out) {
  if (!cfg || !out) return OMB_ERR_INVALID;
  *out = new ombo_spectrum(from_c(*cfg));
  return OMB_OK;
}
void ombo_spectrum_destroy(ombo_spectrum* h) { delete h; }
int ombo_spectrum_get_config(const ombo_spectrum* h, omb_spectrum_config* out) { to_c(h->p.config, out); return OMB_OK; }
int ombo_spectrum_update_config(ombo_spectrum* h, const omb_spectrum_config* cfg) { h->p.update_config(from_c(*cfg)); return OMB_OK; }
int ombo_spectrum_prepare(ombo_spectrum* h) { h->p.prepare(); return OMB_OK; }
int ombo_spectrum_reset_audio(ombo_spectrum* h) { h->p.reset_audio(); return OMB_OK; }
int ombo_spectrum_process_block(ombo_spectrum* h, const float* samples, size_t n_samples, uint32_t channels, float sample_rate,
                                const uint8_t positions[OMB_MAX_CHANNELS], omb_spectrum_snapshot* out) {
  AudioBlock b = AudioBlock::with_positions(samples, n_samples, channels, sample_rate, positions);
  return h->p.process_block(b, out);
}
// test hook: snapshot as currently held (floor_change_reseeds..., processor.rs:459-478)
int ombo_spectrum_peek(ombo_spectrum* h, omb_spectrum_snapshot* out) { h->p.fill(out); return OMB_OK; }
size_t ombo_spectrum_pending(const ombo_spectrum* h, int trace, float* out, size_t cap) {
  const auto& d = h->p.pcm[trace];
  for (size_t i = 0; i < d.size() && i < cap; ++i) out[i] = d[i];
  return d.size();
}
// update_outputs exposed stand-alone (processor.rs:613-651 tests poke SpectrumLevelBuffers directly)
int ombo_spectrum_update_outputs(int mode, float param, float state_floor_weighting_max_unused, const float* weighting, size_t bins,
                                 float floor, float dt, float* smoothed_inout, const float* scratch_power, float* weighted_out,
                                 float* raw_out) {
  (void)state_floor_weighting_max_unused;
  SpectrumLevelBuffers l;
  std::vector<float> w(weighting, weighting + bins);
  l.reset(bins, smoothing_state_floor(w, floor), true);
  std::copy(smoothed_inout, smoothed_inout + bins, l.smoothed.begin());
  std::copy(scratch_power, scratch_power + bins, l.scratch_power.begin());
  std::vector<float> out[2];
  l.update_outputs(mode, param, out, w, dt, floor);
  std::copy(l.smoothed.begin(), l.smoothed.end(), smoothed_inout);
  std::copy(out[0].begin(), out[0].end(), weighted_out);
  std::copy(out[1].begin(), out[1].end(), raw_out);
  return OMB_OK;
}

// ---- loudness
void ombo_loudness_default_config(omb_loudness_config* out) { out->sample_rate = DEFAULT_SAMPLE_RATE; out->floor_db = -99.9f; }
int ombo_loudness_create(const omb_loudness_config* cfg, ombo_loudness** out) {
  if (!cfg || !out) return OMB_ERR_INVALID;
  *out = new ombo_loudness(cfg->sample_rate, cfg->floor_db);
  (*out)->cfg = *cfg;
  return OMB_OK;
}
void ombo_loudness_destroy(ombo_loudness* h) { delete h; }
int ombo_loudness_get_config(const ombo_loudness* h, omb_loudness_config* out) {
  out->sample_rate = h->p.cfg_sample_rate; out->floor_db = h->p.floor_db; return OMB_OK;
}
int ombo_loudness_reset_audio(ombo_loudness* h) { h->p.reset_audio(); return OMB_OK; }
int ombo_loudness_process_block(ombo_loudness* h, const float* samples, size_t n_samples, uint32_t channels, float sample_rate,
                                const uint8_t positions[OMB_MAX_CHANNELS], omb_loudness_snapshot* out) {
  AudioBlock b = AudioBlock::with_positions(samples, n_samples, channels, sample_rate, positions);
  return h->p.process_block(b, out);
}
// test hook for leading_silence_matches_eager_channel_state (processor.rs:400-417)
int ombo_loudness_force_eager(ombo_loudness* h, uint32_t channels, float sample_rate) {
  h->p.ensure_state(channels, sample_rate);
  static const float W[4] = {3.0f, 0.4f, 0.3f, 1.0f};
  std::vector<size_t> caps(4);
  for (int i = 0; i < 4; ++i) caps[i] = window_length(h->p.cfg_sample_rate, W[i]);
  for (auto& ch : h->p.channels) {
    ch.windows = std::make_unique<WindowedMeans>(caps);
    ch.tp = std::make_unique<TruePeakMeter>((double)h->p.cfg_sample_rate);
    for (double& v : ch.filter) v = 0.0;
  }
  return OMB_OK;
}

// ---- batched (same layouts as omb_*_execute_host). `threads` <= 0: all hardware threads.
uint64_t ombo_stft_frames_per_lane(const omb_spectrogram_config* cfg, uint64_t samples) {
  SpectrogramConfig c = from_c(*cfg);
  c.normalize();
  const size_t read_len = c.use_reassignment ? hilbert_len_for(c.fft_size) : c.fft_size;
  return samples >= read_len ? (samples - read_len) / c.hop_size + 1 : 0;
}

int ombo_stft_batch(const omb_spectrogram_config* cfg, const float* lanes, uint32_t n_lanes, uint64_t samples_per_lane,
                    uint64_t lane_stride, omb_spectrogram_point* out_points, uint64_t point_stride, uint32_t* out_counts,
                    uint16_t* out_classic, int threads, uint64_t frame_begin, uint64_t frame_end) {
  SpectrogramConfig c = from_c(*cfg);
  c.normalize();
  const uint64_t frames = ombo_stft_frames_per_lane(cfg, samples_per_lane);
  if (frame_end > frames) frame_end = frames;
  if (frame_begin > frame_end) frame_begin = frame_end;
  const uint64_t span = frame_end - frame_begin;
  // work items: (lane, chunk of frames) so few long lanes still use all threads
  const uint64_t chunk = 64;
  const uint64_t chunks_per_lane = (span + chunk - 1) / chunk;
  const size_t items = (size_t)(chunks_per_lane * n_lanes);
  if (threads <= 0) threads = hw_threads();
  std::vector<std::unique_ptr<ColumnEngine>> engines((size_t)threads);
  std::atomic<size_t> next{0};
  auto worker = [&](int tid) {
    auto& eng = engines[(size_t)tid];
    eng = std::make_unique<ColumnEngine>();
    eng->rebuild(c);
    const size_t bins = eng->bins;
    std::vector<omb_spectrogram_point> col(bins);
    for (size_t it; (it = next.fetch_add(1)) < items;) {
      const uint64_t lane = it / chunks_per_lane, ck = it % chunks_per_lane;
      const uint64_t f0 = frame_begin + ck * chunk, f1 = std::min(frame_end, f0 + chunk);
      const float* x = lanes + lane * lane_stride;
      for (uint64_t f = f0; f < f1; ++f) {
        const float* frame = x + f * c.hop_size;
        const uint64_t slot = lane * frames + f;
        if (c.use_reassignment) {
          const size_t n = eng->reassigned_column(frame, col.data());
          std::copy(col.begin(), col.begin() + (ptrdiff_t)n, out_points + slot * point_stride);
          out_counts[slot] = (uint32_t)n;
        } else {
          eng->classic_column(frame, out_classic + slot * bins);
        }
      }
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; ++t) pool.emplace_back(worker, t);
  worker(0);
  for (auto& th : pool) th.join();
  return OMB_OK;
}

uint64_t ombo_spectrum_hops_per_lane(const omb_spectrum_config* cfg, uint64_t samples) {
  SpectrumConfig c = from_c(*cfg);
  c.normalize();
  return samples >= c.fft_size ? (samples - c.fft_size) / c.hop_size + 1 : 0;
}

// spectrum/state.rs:321-325 — (1..len-1).filter(in [min_f, max_f] && finite).max_by(total_cmp): the last maximum wins.
static int32_t peak_bin(const float* bins_hz, const float* db, size_t n, float min_f, float max_f) {
  int32_t best = -1;
  for (size_t i = 1; i + 1 < n; ++i) {
    if (!(bins_hz[i] >= min_f && bins_hz[i] <= max_f) || !std::isfinite(db[i])) continue;
    if (best < 0) { best = (int32_t)i; continue; }
    // total_cmp on finite values: numeric order, and -0.0 < +0.0
    const float a = db[i], b = db[(size_t)best];
    const bool less = a < b || (a == b && std::signbit(a) && !std::signbit(b));
    if (!less) best = (int32_t)i;
  }
  return best;
}

// spectrum/state.rs:327-356 — returns false for None.
static bool interpolated_peak(const float* bins_hz, const float* db, size_t n, int64_t bin, float* out_freq, float* out_level) {
  const float kEps = 1e-6f;  // state.rs:20
  if (bin <= 0 || (size_t)bin + 1 >= n) return false;
  const float bin_hz = bins_hz[1] - bins_hz[0];
  const float center_freq = bins_hz[bin], center = db[bin];
  if (!(std::isfinite(bin_hz) && bin_hz > 0.0f) || !std::isfinite(center_freq) || !std::isfinite(center)) return false;
  const float left = db[bin - 1], right = db[bin + 1];
  float offset = 0.0f;
  if (std::isfinite(left) && std::isfinite(right)) {
    const float denom = left - 2.0f * center + right;
    if (denom < -kEps) offset = std::fmin(std::fmax(0.5f * (left - right) / denom, -0.5f), 0.5f);
  }
  const float level = offset == 0.0f ? center : std::fmax(center - 0.25f * (left - right) * offset, center);
  *out_freq = std::fmax(center_freq + offset * bin_hz, 0.0f);
  *out_level = level;
  return true;
}

// peak_bin + interpolated_peak over `rows` rows of one dB trace [rows][n_bins]; NaN outputs where the reference has None.
int ombo_spectrum_peaks(const float* bins_hz, const float* db, uint32_t n_bins, uint64_t rows, float min_f, float max_f, int32_t* out_bin,
                        float* out_freq, float* out_level) {
  for (uint64_t r = 0; r < rows; ++r) {
    const float* row = db + r * n_bins;
    const int32_t b = peak_bin(bins_hz, row, n_bins, min_f, max_f);
    float f = std::nanf(""), m = std::nanf("");
    if (b >= 0) interpolated_peak(bins_hz, row, n_bins, b, &f, &m);
    if (out_bin) out_bin[r] = b;
    if (out_freq) out_freq[r] = f;
    if (out_level) out_level[r] = m;
  }
  return OMB_OK;
}
// interpolated_peak alone, for given bins (to check a device result on the device's own dB values bit for bit).
int ombo_spectrum_interpolate_peaks(const float* bins_hz, const float* db, uint32_t n_bins, uint64_t rows, const int32_t* bin,
                                    float* out_freq, float* out_level) {
  for (uint64_t r = 0; r < rows; ++r) {
    float f = std::nanf(""), m = std::nanf("");
    interpolated_peak(bins_hz, db + r * n_bins, n_bins, bin[r], &f, &m);
    out_freq[r] = f;
    out_level[r] = m;
  }
  return OMB_OK;
}

int ombo_spectrum_batch(const omb_spectrum_config* cfg, const float* lanes, uint32_t n_lanes, uint64_t samples_per_lane,
                        uint64_t lane_stride, float* out_weighted, float* out_raw, int32_t* out_peak_bin, int threads) {
  SpectrumConfig c = from_c(*cfg);
  c.normalize();
  const uint64_t hops = ombo_spectrum_hops_per_lane(cfg, samples_per_lane);
  const size_t n = c.fft_size, bins = n / 2 + 1;
  parallel_for(n_lanes, threads, [&](size_t lane) {
    // One lane == one mono SpectrumProcessor trace (source Left on a mono block is the identity).
    SpectrumConfig lc = c;
    lc.source = OMB_CHANNEL_LEFT;
    lc.secondary_source = OMB_CHANNEL_NONE;
    SpectrumProcessor p(lc);
    p.prepare();
    const float* x = lanes + lane * lane_stride;
    const float dt = (float)c.hop_size / c.sample_rate;
    for (uint64_t h = 0; h < hops; ++h) {
      p.pcm[0].assign(x + h * c.hop_size, x + h * c.hop_size + n);
      p.process_trace_window(0, dt, c.floor_db);
      const uint64_t slot = lane * hops + h;
      std::copy(p.traces[0][0].begin(), p.traces[0][0].end(), out_weighted + slot * bins);
      std::copy(p.traces[0][1].begin(), p.traces[0][1].end(), out_raw + slot * bins);
      if (out_peak_bin) {  // spectrum/state.rs:106-107,134-136: the default peak label (A-weighted trace, 20 Hz..)
        const float min_f = 20.0f, max_f = std::fmax(p.freq_bins[bins - 1], min_f * 1.02f);
        out_peak_bin[slot] = peak_bin(p.freq_bins.data(), p.traces[0][0].data(), bins, min_f, max_f);
      }
    }
  });
  return OMB_OK;
}

int ombo_loudness_batch(const omb_loudness_config* cfg, uint32_t channels, const uint8_t positions[OMB_MAX_CHANNELS],
                        const float* interleaved, uint32_t n_streams, uint64_t frames, uint64_t stream_stride,
                        uint64_t block_frames, omb_loudness_snapshot* out, int threads) {
  const uint64_t n_blocks = (frames + block_frames - 1) / block_frames;
  parallel_for(n_streams, threads, [&](size_t s) {
    LoudnessProcessor p(cfg->sample_rate, cfg->floor_db);
    const float* x = interleaved + s * stream_stride;
    for (uint64_t b = 0; b < n_blocks; ++b) {
      const uint64_t f0 = b * block_frames, f1 = std::min(frames, f0 + block_frames);
      AudioBlock blk = AudioBlock::with_positions(x + f0 * channels, (size_t)((f1 - f0) * channels), channels, cfg->sample_rate, positions);
      omb_loudness_snapshot snap;
      std::memset(&snap, 0, sizeof snap);
      p.process_block(blk, &snap);
      out[s * n_blocks + b] = snap;
    }
  });
  return OMB_OK;
}

}  // extern "C"
