// omb200.hpp — C++17 host-side mirror of the reference's processor surface over the C ABI (omb200.h).
//
// Same names and meaning as src/visuals/{spectrogram,spectrum,loudness}/processor.rs and dsp.rs:108-262:
// Processor::new(config) -> constructor; config(); update_config(); prepare(); process_block(block) ->
// std::optional<snapshot> (nullopt = the reference's None); reset_audio().  A failing CUDA call throws
// omb::Error (the reference would panic=abort).  Header-only; link with -lomb200.
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "omb200.h"

namespace omb {

struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& what) : std::runtime_error(what + ": " + omb_last_error()), status(s) {}
};
inline int check(int rc, const char* what) {
  if (rc < 0) throw Error(rc, what);
  return rc;
}

// dsp.rs:108-115 — a borrowed interleaved block. positions empty => ChannelPosition::fallback(channels).
struct AudioBlock {
  const float* samples = nullptr;
  size_t len = 0;  // interleaved sample count
  uint32_t channels = 1;
  float sample_rate = 48000.0f;
  std::optional<std::array<uint8_t, OMB_MAX_CHANNELS>> positions;
  size_t frame_count() const { return len / (channels ? channels : 1); }
  bool is_empty() const { return len < (channels ? channels : 1); }
  const uint8_t* pos_ptr() const { return positions ? positions->data() : nullptr; }
};

using SpectrogramConfig = omb_spectrogram_config;
using SpectrogramPoint = omb_spectrogram_point;

struct SpectrogramColumn {  // spectrogram/processor.rs:124-127
  std::vector<SpectrogramPoint> reassigned;
  std::vector<uint16_t> classic;
};
struct SpectrogramUpdate {  // spectrogram/processor.rs:160-168
  size_t fft_size, hop_size;
  float sample_rate;
  size_t history_length;
  bool reset;
  float reassigned_power_scale;
  bool is_reassigned;
  std::vector<SpectrogramColumn> new_columns;
};

class SpectrogramProcessor {
 public:
  static SpectrogramConfig default_config() { SpectrogramConfig c; omb_spectrogram_default_config(&c); return c; }
  explicit SpectrogramProcessor(const SpectrogramConfig& cfg = default_config()) { check(omb_spectrogram_create(&cfg, &h_), "omb_spectrogram_create"); }
  ~SpectrogramProcessor() { omb_spectrogram_destroy(h_); }
  SpectrogramProcessor(const SpectrogramProcessor&) = delete;
  SpectrogramProcessor& operator=(const SpectrogramProcessor&) = delete;
  SpectrogramConfig config() const { SpectrogramConfig c; check(omb_spectrogram_get_config(h_, &c), "get_config"); return c; }
  void update_config(const SpectrogramConfig& c) { check(omb_spectrogram_update_config(h_, &c), "update_config"); }
  void prepare() { check(omb_spectrogram_prepare(h_), "prepare"); }
  void reset_audio() { check(omb_spectrogram_reset_audio(h_), "reset_audio"); }
  std::optional<SpectrogramUpdate> process_block(const AudioBlock& b) {
    omb_spectrogram_update up{};
    if (check(omb_spectrogram_process_block(h_, b.samples, b.len, b.channels, b.sample_rate, b.pos_ptr(), &up), "process_block") == OMB_NO_DATA)
      return std::nullopt;
    SpectrogramUpdate u{(size_t)up.fft_size, (size_t)up.hop_size, up.sample_rate, (size_t)up.history_length, up.reset != 0,
                        up.reassigned_power_scale, up.kind == OMB_COLUMN_REASSIGNED, {}};
    u.new_columns.resize(up.n_columns);
    for (uint32_t c = 0; c < up.n_columns; ++c) {
      if (u.is_reassigned) u.new_columns[c].reassigned.assign(up.points + up.column_offsets[c], up.points + up.column_offsets[c + 1]);
      else u.new_columns[c].classic.assign(up.classic_db + (size_t)c * up.bins, up.classic_db + (size_t)(c + 1) * up.bins);
    }
    return u;
  }

  omb_spectrogram* handle() const { return h_; }

 private:
  omb_spectrogram* h_ = nullptr;
};

using SpectrumConfig = omb_spectrum_config;
struct SpectrumSnapshot {  // spectrum/processor.rs:33-37; traces[trace][0] weighted, [trace][1] raw
  std::vector<float> frequency_bins;
  std::array<std::array<std::vector<float>, 2>, 2> traces;
};

class SpectrumProcessor {
 public:
  static SpectrumConfig default_config() { SpectrumConfig c; omb_spectrum_default_config(&c); return c; }
  explicit SpectrumProcessor(const SpectrumConfig& cfg = default_config()) { check(omb_spectrum_create(&cfg, &h_), "omb_spectrum_create"); }
  ~SpectrumProcessor() { omb_spectrum_destroy(h_); }
  SpectrumProcessor(const SpectrumProcessor&) = delete;
  SpectrumProcessor& operator=(const SpectrumProcessor&) = delete;
  SpectrumConfig config() const { SpectrumConfig c; check(omb_spectrum_get_config(h_, &c), "get_config"); return c; }
  void update_config(const SpectrumConfig& c) { check(omb_spectrum_update_config(h_, &c), "update_config"); }
  void prepare() { check(omb_spectrum_prepare(h_), "prepare"); }
  void reset_audio() { check(omb_spectrum_reset_audio(h_), "reset_audio"); }
  // The reference returns Option<&SpectrumSnapshot> borrowed from the processor; so does this.
  const SpectrumSnapshot* process_block(const AudioBlock& b) {
    omb_spectrum_snapshot s{};
    if (check(omb_spectrum_process_block(h_, b.samples, b.len, b.channels, b.sample_rate, b.pos_ptr(), &s), "process_block") == OMB_NO_DATA)
      return nullptr;
    snap_.frequency_bins.assign(s.frequency_bins, s.frequency_bins + s.bins);
    for (int t = 0; t < 2; ++t)
      for (int w = 0; w < 2; ++w) snap_.traces[t][w].assign(s.traces[t][w], s.traces[t][w] + s.bins);
    return &snap_;
  }

  omb_spectrum* handle() const { return h_; }

 private:
  omb_spectrum* h_ = nullptr;
  SpectrumSnapshot snap_;
};

using LoudnessConfig = omb_loudness_config;
using LoudnessSnapshot = omb_loudness_snapshot;  // loudness/processor.rs:185-194 (Copy)

class LoudnessProcessor {
 public:
  static LoudnessConfig default_config() { LoudnessConfig c; omb_loudness_default_config(&c); return c; }
  explicit LoudnessProcessor(const LoudnessConfig& cfg = default_config()) { check(omb_loudness_create(&cfg, &h_), "omb_loudness_create"); }
  ~LoudnessProcessor() { omb_loudness_destroy(h_); }
  LoudnessProcessor(const LoudnessProcessor&) = delete;
  LoudnessProcessor& operator=(const LoudnessProcessor&) = delete;
  LoudnessConfig config() const { LoudnessConfig c; check(omb_loudness_get_config(h_, &c), "get_config"); return c; }
  void reset_audio() { check(omb_loudness_reset_audio(h_), "reset_audio"); }
  std::optional<LoudnessSnapshot> process_block(const AudioBlock& b) {
    LoudnessSnapshot s{};
    if (check(omb_loudness_process_block(h_, b.samples, b.len, b.channels, b.sample_rate, b.pos_ptr(), &s), "process_block") == OMB_NO_DATA)
      return std::nullopt;
    return s;
  }

  omb_loudness* handle() const { return h_; }

 private:
  omb_loudness* h_ = nullptr;
};

// ---------------------------------------------------------------------------------------------------------------
// Rows f1 / f4: the ordered audio timeline (meter.rs, infra/pipewire/transport.rs)
using AudioFormat = omb_audio_format;  // dsp.rs:79-85

// transport.rs:39-54 CapturedSpan
struct CapturedSpan {
  int kind = OMB_SPAN_RESET;          // OMB_SPAN_*
  std::vector<float> samples;         // Pcm
  uint64_t frames = 0;                // Silence
  AudioFormat format{};
};

// AudioReader's packet timeline (transport.rs:573-657): accept / flush / reset_timeline.
class PacketTimeline {
 public:
  explicit PacketTimeline(const AudioFormat& initial) { check(omb_timeline_create(&initial, &h_), "omb_timeline_create"); }
  ~PacketTimeline() { omb_timeline_destroy(h_); }
  PacketTimeline(const PacketTimeline&) = delete;
  PacketTimeline& operator=(const PacketTimeline&) = delete;
  // samples == nullptr: a silence packet. Spans are appended to `out` in emission order.
  void accept(const float* samples, uint64_t frames, const AudioFormat& f, uint64_t start_ns, uint64_t end_ns, std::vector<CapturedSpan>& out) {
    check(omb_timeline_accept(h_, samples, frames, &f, start_ns, end_ns, &PacketTimeline::collect, &out), "omb_timeline_accept");
  }
  void flush(std::vector<CapturedSpan>& out) { check(omb_timeline_flush(h_, &PacketTimeline::collect, &out), "omb_timeline_flush"); }
  void reset_timeline(uint64_t cursor_ns) { check(omb_timeline_reset(h_, cursor_ns), "omb_timeline_reset"); }
  uint64_t cursor() const { return omb_timeline_cursor(h_); }

 private:
  static void collect(void* user, int kind, const float* samples, size_t n, uint64_t frames, const omb_audio_format* f) {
    CapturedSpan s;
    s.kind = kind;
    if (samples) s.samples.assign(samples, samples + n);
    s.frames = frames;
    if (f) s.format = *f;
    static_cast<std::vector<CapturedSpan>*>(user)->push_back(std::move(s));
  }
  omb_timeline* h_ = nullptr;
};

// DspBatcher (meter.rs:27-84) + ingest_silence (:143-165) + VisualManager::ingest_samples (registry.rs:396-418) over
// borrowed processors.  `on_ingest` (optional) sees every ingested chunk with the processors' outputs.
class DspBatcher {
 public:
  DspBatcher(SpectrogramProcessor* sg, SpectrumProcessor* sp, LoudnessProcessor* ld, omb_ingest_fn on_ingest = nullptr, void* user = nullptr) {
    check(omb_meter_create(&h_), "omb_meter_create");
    check(omb_meter_attach(h_, sg ? sg->handle() : nullptr, sp ? sp->handle() : nullptr, ld ? ld->handle() : nullptr), "omb_meter_attach");
    if (on_ingest) check(omb_meter_set_callback(h_, on_ingest, user), "omb_meter_set_callback");
  }
  ~DspBatcher() { omb_meter_destroy(h_); }
  DspBatcher(const DspBatcher&) = delete;
  DspBatcher& operator=(const DspBatcher&) = delete;
  uint32_t push(const float* samples, size_t n, const AudioFormat& f) { uint32_t c = 0; check(omb_meter_push(h_, samples, n, &f, &c), "omb_meter_push"); return c; }
  uint32_t ingest_silence(uint64_t frames, const AudioFormat& f) { uint32_t c = 0; check(omb_meter_push_silence(h_, frames, &f, &c), "omb_meter_push_silence"); return c; }
  // MeterEngine::advance's dispatch (meter.rs:115-124)
  uint32_t consume(const CapturedSpan& s) {
    uint32_t c = 0;
    check(omb_meter_consume_span(h_, s.kind, s.samples.data(), s.samples.size(), s.frames, &s.format, &c), "omb_meter_consume_span");
    return c;
  }
  void reset() { check(omb_meter_reset(h_), "omb_meter_reset"); }
  void clear() { check(omb_meter_clear(h_), "omb_meter_clear"); }
  size_t pending_samples() const { return omb_meter_pending_samples(h_); }

 private:
  omb_meter* h_ = nullptr;
};

// Row f1, device side: S lock-step processors of one kind behind one handle (one launch chain per push for all streams).
// Outputs are the library's dense views (valid until the next push); see omb200.h for the layouts.
class SpectrogramBank {
 public:
  SpectrogramBank(const SpectrogramConfig& cfg, uint32_t n_streams) { check(omb_spectrogram_bank_create(&cfg, n_streams, &h_), "omb_spectrogram_bank_create"); }
  ~SpectrogramBank() { omb_spectrogram_bank_destroy(h_); }
  SpectrogramBank(const SpectrogramBank&) = delete;
  SpectrogramBank& operator=(const SpectrogramBank&) = delete;
  void reset_audio() { check(omb_spectrogram_bank_reset_audio(h_), "reset_audio"); }
  // false: no stream has a new column (the reference's None)
  bool push(const float* samples, uint64_t stream_stride, size_t frames, uint32_t channels, float sample_rate, const uint8_t* positions,
            omb_spectrogram_bank_update& out) {
    return check(omb_spectrogram_bank_push(h_, samples, stream_stride, frames, channels, sample_rate, positions, &out), "push") != OMB_NO_DATA;
  }
  size_t pending() const { return omb_spectrogram_bank_pending(h_); }

 private:
  omb_spectrogram_bank* h_ = nullptr;
};

class SpectrumBank {
 public:
  SpectrumBank(const SpectrumConfig& cfg, uint32_t n_streams) { check(omb_spectrum_bank_create(&cfg, n_streams, &h_), "omb_spectrum_bank_create"); }
  ~SpectrumBank() { omb_spectrum_bank_destroy(h_); }
  SpectrumBank(const SpectrumBank&) = delete;
  SpectrumBank& operator=(const SpectrumBank&) = delete;
  void reset_audio() { check(omb_spectrum_bank_reset_audio(h_), "reset_audio"); }
  bool push(const float* samples, uint64_t stream_stride, size_t frames, uint32_t channels, float sample_rate, const uint8_t* positions,
            omb_spectrum_bank_snapshot& out) {
    return check(omb_spectrum_bank_push(h_, samples, stream_stride, frames, channels, sample_rate, positions, &out), "push") != OMB_NO_DATA;
  }
  size_t pending() const { return omb_spectrum_bank_pending(h_); }

 private:
  omb_spectrum_bank* h_ = nullptr;
};

class LoudnessBank {
 public:
  LoudnessBank(const LoudnessConfig& cfg, uint32_t n_streams) : n_(n_streams) { check(omb_loudness_bank_create(&cfg, n_streams, &h_), "omb_loudness_bank_create"); }
  ~LoudnessBank() { omb_loudness_bank_destroy(h_); }
  LoudnessBank(const LoudnessBank&) = delete;
  LoudnessBank& operator=(const LoudnessBank&) = delete;
  void reset_audio() { check(omb_loudness_bank_reset_audio(h_), "reset_audio"); }
  // one LoudnessSnapshot per stream; empty when the block holds less than one frame
  std::vector<LoudnessSnapshot> push(const float* samples, uint64_t stream_stride, size_t n_samples, uint32_t channels, float sample_rate,
                                     const uint8_t* positions) {
    std::vector<LoudnessSnapshot> out(n_);
    if (check(omb_loudness_bank_push(h_, samples, stream_stride, n_samples, channels, sample_rate, positions, out.data()), "push") == OMB_NO_DATA)
      out.clear();
    return out;
  }

 private:
  omb_loudness_bank* h_ = nullptr;
  uint32_t n_ = 0;
};

// Row f2: accumulation + resolve passes of the spectrogram view (render.rs:104-165, spectrogram.wgsl) on host buffers.
using SplatParams = omb_splat_params;
struct SplatImages {
  uint32_t width = 0, height = 0;
  std::vector<float> accum, db;  // [ring][height][width]
};
inline SplatImages splat_render(const SpectrogramPoint* rings, uint64_t point_stride, const uint32_t* slot_counts, uint32_t n_rings,
                                const SplatParams& p) {
  SplatImages im;
  omb_splat_image_size(&p, &im.width, &im.height);
  im.accum.resize((size_t)im.width * im.height * n_rings);
  im.db.resize(im.accum.size());
  check(omb_splat_render_host(rings, point_stride, slot_counts, n_rings, &p, im.accum.data(), im.db.data()), "omb_splat_render_host");
  return im;
}

// registry.rs:247-256 VisualModule — the trait the reference drives its visuals through.
struct VisualModule {
  virtual ~VisualModule() = default;
  virtual void ingest(const AudioBlock& block) = 0;
  virtual void reset_audio() = 0;
  virtual void prepare() {}
};

// registry.rs:106-118: `if let Some(snap) = processor.process_block(block) { state.apply_snapshot(snap) }`
template <class Processor, class State>
struct Visual final : VisualModule {
  Processor processor;
  State state;
  void ingest(const AudioBlock& block) override {
    if (auto snap = processor.process_block(block)) state.apply_snapshot(*snap);
  }
  void reset_audio() override { processor.reset_audio(); state.reset_audio(); }
  void prepare() override { if constexpr (requires(Processor& p) { p.prepare(); }) processor.prepare(); }
};

}  // namespace omb
