// meter.cu — the ordered audio timeline either side of the processors (rows f1 and f4 of SURVEY.md §8).
//
//   omb_timeline : AudioReader's packet timeline — gap -> Silence span, overlap -> skipped frames, PCM coalesced
//                  in a scratch vector (infra/pipewire/transport.rs:573-657).
//   omb_meter    : DspBatcher (meter.rs:27-84: carry-over of partial batches, sample-rate-scaled batch size,
//                  never mixing format generations), ingest_silence (meter.rs:143-165) and
//                  VisualManager::ingest_samples (visuals/registry.rs:396-418: reset_audio when the format
//                  generation changes, then one process_block per enabled module).
//
// Host-side state machines: they hold a few KB of carry-over at most and hand whole batches to the streaming
// processors, whose FIFOs and kernels live on the device.  Nothing here touches CUDA directly.
#include <algorithm>
#include <cmath>
#include <new>
#include <vector>

#include "common.h"

using namespace omb;

namespace {

constexpr size_t kSilenceChunkFrames = 4096;      // meter.rs:15
constexpr size_t kDspBatchFramesAt48k = 256;      // meter.rs:16
constexpr size_t kMaxDspIngestFramesAt48k = 1024; // meter.rs:17
constexpr uint64_t kMaxSilenceSeconds = 2;        // meter.rs:18
constexpr double kDefaultRate = 48000.0;          // util/audio DEFAULT_SAMPLE_RATE

bool same_format(const omb_audio_format& a, const omb_audio_format& b) {
  // #[derive(PartialEq)] over every field (dsp.rs:79-85); f32 compare as Rust does (NaN != NaN)
  return a.channels == b.channels && a.sample_rate == b.sample_rate && a.generation == b.generation &&
         std::equal(a.positions, a.positions + OMB_MAX_CHANNELS, b.positions);
}

// meter.rs:20-25
size_t scaled_samples(size_t frames_at_48k, const omb_audio_format& f) {
  const double frames = std::max(std::round((double)frames_at_48k * (double)f.sample_rate / kDefaultRate), 1.0);
  return (size_t)frames * std::max<size_t>(f.channels, 1);
}

// dsp.rs:103-105 AudioFormat::rate
uint64_t format_rate(const omb_audio_format& f) { return (uint64_t)std::max(std::round(f.sample_rate), 1.0f); }

// transport.rs:105-126
uint64_t scale_u64(uint64_t value, uint64_t num, uint64_t den) {
  const unsigned __int128 r = (unsigned __int128)value * num / std::max<uint64_t>(den, 1);
  return r > (unsigned __int128)UINT64_MAX ? UINT64_MAX : (uint64_t)r;
}
uint64_t ns_frames(uint64_t ns, uint64_t rate) { return scale_u64(ns, rate, 1000000000ull); }
uint64_t ns_frames_ceil(uint64_t ns, uint64_t rate) {
  const unsigned __int128 p = (unsigned __int128)ns * rate;
  const unsigned __int128 r = (p + 999999999ull) / 1000000000ull;
  return r > (unsigned __int128)UINT64_MAX ? UINT64_MAX : (uint64_t)r;
}

}  // namespace

struct omb_timeline {
  std::vector<float> scratch;
  omb_audio_format format{};
  uint64_t cursor = 0;
  bool align_next_packet = true;

  void flush(omb_span_fn consume, void* user) {  // transport.rs:634-643
    if (scratch.empty()) return;
    if (consume) consume(user, OMB_SPAN_PCM, scratch.data(), scratch.size(), 0, &format);
    scratch.clear();
  }
};

struct omb_meter {
  std::vector<float> samples;   // DspBatcher.samples
  bool has_format = false;      // DspBatcher.format
  omb_audio_format format{};
  bool has_generation = false;  // VisualManager.format_generation
  uint64_t generation = 0;
  omb_spectrogram* sg = nullptr;
  omb_spectrum* sp = nullptr;
  omb_loudness* ld = nullptr;
  omb_ingest_fn cb = nullptr;
  void* user = nullptr;
  std::vector<float> silence;   // MeterEngine.silence

  int reset_audio() {  // VisualManager::reset_audio -> every module
    if (sg) OMB_TRY(omb_spectrogram_reset_audio(sg));
    if (sp) OMB_TRY(omb_spectrum_reset_audio(sp));
    if (ld) OMB_TRY(omb_loudness_reset_audio(ld));
    return OMB_OK;
  }

  // visuals/registry.rs:396-418
  int ingest(const float* x, size_t n, const omb_audio_format& f) {
    if (n == 0) return OMB_OK;
    if (has_generation && generation != f.generation) OMB_TRY(reset_audio());
    has_generation = true;
    generation = f.generation;
    omb_spectrogram_update up{};
    omb_spectrum_snapshot sn{};
    omb_loudness_snapshot ls{};
    const omb_spectrogram_update* pup = nullptr;
    const omb_spectrum_snapshot* psn = nullptr;
    const omb_loudness_snapshot* pls = nullptr;
    if (sg) {
      const int rc = omb_spectrogram_process_block(sg, x, n, f.channels, f.sample_rate, f.positions, &up);
      if (rc < 0) return rc;
      if (rc == OMB_OK) pup = &up;
    }
    if (sp) {
      const int rc = omb_spectrum_process_block(sp, x, n, f.channels, f.sample_rate, f.positions, &sn);
      if (rc < 0) return rc;
      if (rc == OMB_OK) psn = &sn;
    }
    if (ld) {
      const int rc = omb_loudness_process_block(ld, x, n, f.channels, f.sample_rate, f.positions, &ls);
      if (rc < 0) return rc;
      if (rc == OMB_OK) pls = &ls;
    }
    if (cb) cb(user, x, n, &f, pup, psn, pls);
    return OMB_OK;
  }

  // meter.rs:40-73
  int push(const float* x, size_t n, const omb_audio_format& f, uint32_t* count_out) {
    if (has_format && !same_format(format, f)) samples.clear();
    has_format = true;
    format = f;
    const size_t batch = scaled_samples(kDspBatchFramesAt48k, f);
    uint32_t count = 0;
    if (!samples.empty()) {
      const size_t take = std::min(batch - samples.size(), n);
      samples.insert(samples.end(), x, x + take);
      x += take;
      n -= take;
      if (samples.size() == batch) {
        OMB_TRY(ingest(samples.data(), samples.size(), f));
        samples.clear();
        ++count;
      }
    }
    const size_t ready = n / batch * batch;
    const size_t chunk = scaled_samples(kMaxDspIngestFramesAt48k, f);
    for (size_t o = 0; o < ready; o += chunk) {
      OMB_TRY(ingest(x + o, std::min(chunk, ready - o), f));
      ++count;
    }
    samples.insert(samples.end(), x + ready, x + n);
    if (count_out) *count_out += count;
    return OMB_OK;
  }

  void clear() {  // meter.rs:80-83
    samples.clear();
    has_format = false;
  }
  int reset() {   // meter.rs:75-78
    clear();
    return reset_audio();
  }

  // meter.rs:143-165
  int push_silence(uint64_t frames, const omb_audio_format& f, uint32_t* count_out) {
    const uint64_t limit = (uint64_t)std::max(std::round((double)kMaxSilenceSeconds * (double)f.sample_rate), 1.0);
    if (frames > limit) return reset();
    const size_t ch = std::max<size_t>(f.channels, 1);
    if (silence.empty()) silence.assign(kSilenceChunkFrames * OMB_MAX_CHANNELS, 0.0f);
    const uint64_t capacity = silence.size() / ch;
    uint64_t remaining = frames;
    while (remaining > 0) {
      const size_t chunk = (size_t)std::min<uint64_t>(remaining, capacity);
      OMB_TRY(push(silence.data(), chunk * f.channels, f, count_out));
      remaining -= chunk;
    }
    return OMB_OK;
  }
};

#define OMB_GUARD_BEGIN try {
#define OMB_GUARD_END                                                                      \
  }                                                                                        \
  catch (const std::bad_alloc&) { return fail(OMB_ERR_NOMEM, "host allocation failed"); } \
  catch (...) { return fail(OMB_ERR_INVALID, "unexpected C++ exception"); }

extern "C" {

int omb_timeline_create(const omb_audio_format* initial_format, omb_timeline** out) {
  if (!out) return fail(OMB_ERR_INVALID, "null argument");
  OMB_GUARD_BEGIN
  auto* t = new omb_timeline();
  if (initial_format) t->format = *initial_format;
  *out = t;
  return OMB_OK;
  OMB_GUARD_END
}
void omb_timeline_destroy(omb_timeline* t) { delete t; }

int omb_timeline_accept(omb_timeline* t, const float* samples, uint64_t frames, const omb_audio_format* format,
                        uint64_t start_ns, uint64_t end_ns, omb_span_fn consume, void* user) {
  if (!t || !format) return fail(OMB_ERR_INVALID, "null argument");
  OMB_GUARD_BEGIN
  // ::switch (transport.rs:627-632)
  if (!same_format(t->format, *format)) {
    t->flush(consume, user);
    t->format = *format;
  }
  if (t->align_next_packet) {
    t->align_next_packet = false;
    t->cursor = start_ns;
  }
  const uint64_t rate = format_rate(*format);
  const bool has_gap = start_ns > t->cursor;
  const uint64_t gap = has_gap ? ns_frames(start_ns - t->cursor, rate) : 0;
  uint64_t skip = 0;
  if (t->cursor > start_ns) skip = std::min(ns_frames_ceil(std::min(t->cursor, end_ns) - start_ns, rate), frames);
  t->cursor = std::max(t->cursor, end_ns);
  if (has_gap && gap > 0) {
    t->flush(consume, user);
    if (consume) consume(user, OMB_SPAN_SILENCE, nullptr, 0, gap, format);
  }
  if (samples) {
    if (skip < frames) t->scratch.insert(t->scratch.end(), samples + skip * format->channels, samples + frames * format->channels);
  } else if (skip < frames) {
    t->flush(consume, user);
    if (consume) consume(user, OMB_SPAN_SILENCE, nullptr, 0, frames - skip, format);
  }
  return OMB_OK;
  OMB_GUARD_END
}

int omb_timeline_flush(omb_timeline* t, omb_span_fn consume, void* user) {
  if (!t) return fail(OMB_ERR_INVALID, "null argument");
  OMB_GUARD_BEGIN
  t->flush(consume, user);
  return OMB_OK;
  OMB_GUARD_END
}

int omb_timeline_reset(omb_timeline* t, uint64_t cursor_ns) {
  if (!t) return fail(OMB_ERR_INVALID, "null argument");
  t->scratch.clear();
  t->cursor = cursor_ns;
  t->align_next_packet = true;
  return OMB_OK;
}
uint64_t omb_timeline_cursor(const omb_timeline* t) { return t ? t->cursor : 0; }
size_t omb_timeline_pending_samples(const omb_timeline* t) { return t ? t->scratch.size() : 0; }

int omb_meter_create(omb_meter** out) {
  if (!out) return fail(OMB_ERR_INVALID, "null argument");
  OMB_GUARD_BEGIN
  auto* m = new omb_meter();
  m->samples.reserve(kDspBatchFramesAt48k * OMB_MAX_CHANNELS);  // meter.rs:35
  *out = m;
  return OMB_OK;
  OMB_GUARD_END
}
void omb_meter_destroy(omb_meter* m) { delete m; }

int omb_meter_attach(omb_meter* m, omb_spectrogram* spectrogram, omb_spectrum* spectrum, omb_loudness* loudness) {
  if (!m) return fail(OMB_ERR_INVALID, "null argument");
  m->sg = spectrogram;
  m->sp = spectrum;
  m->ld = loudness;
  return OMB_OK;
}
int omb_meter_set_callback(omb_meter* m, omb_ingest_fn fn, void* user) {
  if (!m) return fail(OMB_ERR_INVALID, "null argument");
  m->cb = fn;
  m->user = user;
  return OMB_OK;
}
int omb_meter_push(omb_meter* m, const float* samples, size_t n_samples, const omb_audio_format* format, uint32_t* n_ingests) {
  if (!m || !format || (!samples && n_samples)) return fail(OMB_ERR_INVALID, "null argument");
  if (n_ingests) *n_ingests = 0;
  OMB_GUARD_BEGIN
  return m->push(samples, n_samples, *format, n_ingests);
  OMB_GUARD_END
}
int omb_meter_push_silence(omb_meter* m, uint64_t frames, const omb_audio_format* format, uint32_t* n_ingests) {
  if (!m || !format) return fail(OMB_ERR_INVALID, "null argument");
  if (n_ingests) *n_ingests = 0;
  OMB_GUARD_BEGIN
  return m->push_silence(frames, *format, n_ingests);
  OMB_GUARD_END
}
int omb_meter_reset(omb_meter* m) {
  if (!m) return fail(OMB_ERR_INVALID, "null argument");
  return m->reset();
}
int omb_meter_clear(omb_meter* m) {
  if (!m) return fail(OMB_ERR_INVALID, "null argument");
  m->clear();
  return OMB_OK;
}
int omb_meter_consume_span(omb_meter* m, int kind, const float* samples, size_t n_samples, uint64_t frames,
                           const omb_audio_format* format, uint32_t* n_ingests) {
  if (!m) return fail(OMB_ERR_INVALID, "null argument");
  if (n_ingests) *n_ingests = 0;
  OMB_GUARD_BEGIN
  switch (kind) {  // meter.rs:115-124
    case OMB_SPAN_PCM:
      if (!format || (!samples && n_samples)) return fail(OMB_ERR_INVALID, "null argument");
      return m->push(samples, n_samples, *format, n_ingests);
    case OMB_SPAN_SILENCE:
      if (!format) return fail(OMB_ERR_INVALID, "null argument");
      return m->push_silence(frames, *format, n_ingests);
    case OMB_SPAN_RESET:
      return m->reset();
    default:
      return fail(OMB_ERR_INVALID, "unknown span kind %d", kind);
  }
  OMB_GUARD_END
}
size_t omb_meter_pending_samples(const omb_meter* m) { return m ? m->samples.size() : 0; }
int omb_meter_has_format(const omb_meter* m) { return m && m->has_format ? 1 : 0; }

}  // extern "C"
