/*
 * omb200.h — C ABI of the B200-native OpenMeters DSP hot path.
 *
 * This is the drop-in boundary: every entry point replaces one method of the
 * reference's three processors (all citations are into /root/reference/):
 *
 *   SpectrogramProcessor  src/visuals/spectrogram/processor.rs:187-223,490-543
 *   SpectrumProcessor     src/visuals/spectrum/processor.rs:88-124,255-322
 *   LoudnessProcessor     src/visuals/loudness/processor.rs:224-311
 *
 * which the reference drives through `VisualModule::ingest(&AudioBlock)`
 * (src/visuals/registry.rs:106-115,247-256).  The reference crate is
 * `#![forbid(unsafe_code)]` (src/main.rs:4), so the Rust binding lives in a
 * separate `-sys` crate; see INTEGRATION.md for the stub.
 *
 * Conventions
 *   - plain C types only; no exceptions cross the boundary.
 *   - return value: OMB_OK (0), OMB_NO_DATA (1, the reference's `None`), or a
 *     negative omb_status; omb_last_error() gives a thread-local message.
 *   - handles are not thread-safe (the reference's processors are `!Send`,
 *     registry.rs:23); distinct handles are independent.
 *   - output structs point into library-owned host memory that stays valid
 *     until the next call on the same handle (the reference moves an owned
 *     SpectrogramUpdate / borrows &SpectrumSnapshot; the caller copies).
 *   - there is NO CPU fallback: every compute entry point fails with
 *     OMB_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef OMB200_H
#define OMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OMB_MAX_CHANNELS 8 /* dsp.rs:6 MAX_AUDIO_CHANNELS */

typedef enum omb_status {
  OMB_OK = 0,
  OMB_NO_DATA = 1,          /* process_block returned None */
  OMB_ERR_INVALID = -1,     /* null pointer / malformed argument */
  OMB_ERR_UNSUPPORTED = -2, /* e.g. non power-of-two fft_size: no kernel, and no CPU fallback */
  OMB_ERR_CUDA = -3,        /* CUDA runtime failure or no device */
  OMB_ERR_NOMEM = -4
} omb_status;

/* util/audio/window.rs:9-18 WindowKind */
typedef enum omb_window_kind {
  OMB_WINDOW_RECTANGULAR = 0,
  OMB_WINDOW_HANN = 1,
  OMB_WINDOW_HAMMING = 2,
  OMB_WINDOW_BLACKMAN = 3,
  OMB_WINDOW_BLACKMAN_HARRIS = 4
} omb_window_kind;

/* util/audio/channel.rs:4-10 Channel */
typedef enum omb_channel {
  OMB_CHANNEL_LEFT = 0,
  OMB_CHANNEL_RIGHT = 1,
  OMB_CHANNEL_MID = 2,
  OMB_CHANNEL_SIDE = 3,
  OMB_CHANNEL_NONE = 4
} omb_channel;

/* dsp.rs:8-22 ChannelPosition, one byte per channel. Aux(n) = OMB_POS_AUX0 + n. */
enum {
  OMB_POS_FRONT_LEFT = 0,
  OMB_POS_FRONT_RIGHT = 1,
  OMB_POS_FRONT_CENTER = 2,
  OMB_POS_LOW_FREQUENCY = 3,
  OMB_POS_REAR_LEFT = 4,
  OMB_POS_REAR_RIGHT = 5,
  OMB_POS_SIDE_LEFT = 6,
  OMB_POS_SIDE_RIGHT = 7,
  OMB_POS_MONO = 8,
  OMB_POS_UNKNOWN = 9,
  OMB_POS_AUX0 = 16
};

/* spectrum/processor.rs:64-70 AveragingMode */
typedef enum omb_averaging_mode {
  OMB_AVG_NONE = 0,
  OMB_AVG_EXPONENTIAL = 1, /* param = factor */
  OMB_AVG_PEAK_HOLD = 2    /* param = decay_per_second */
} omb_averaging_mode;

/* ------------------------------------------------------------------------ */
/* Library / device                                                          */
/* ------------------------------------------------------------------------ */

/* Thread-local description of the last failure on this thread ("" if none). */
const char* omb_last_error(void);
/* "omb200 <version> sm_100a" */
const char* omb_version(void);
/* Number of usable CUDA devices (0 if none; never fails). */
int omb_device_count(void);
/* Select the device used by handles/plans created afterwards on this thread. */
int omb_set_device(int device);
/* Total kernel launches issued by this library in this process (bench.py's gpu_launches). */
uint64_t omb_kernel_launch_count(void);
/* Measured FP32 peak of the current device in TFLOP/s: independent FFMA chains on every SM, timed with CUDA events
 * (the second roofline of SURVEY.md 8(d): the reassigned path is FP32-issue bound, not HBM bound). */
int omb_probe_fp32_tflops(double* out_tflops);

/* ------------------------------------------------------------------------ */
/* Multi-GPU ingest over peer memory (no reference equivalent: */
/* the reference is one process on one machine).  One process per GPU: the    */
/* ingest rank allocates the job's PCM with omb_peer_alloc and hands the       */
/* 64-byte handle to the other ranks (any transport: torch.distributed, a      */
/* pipe); they map it with omb_peer_open and either pull their lanes with      */
/* their own copy engines (omb_copy_async, no SMs involved, overlaps kernels)  */
/* or pass the mapped pointer straight to omb_*_execute_device — the kernels'  */
/* staging copies (bulk / 16-byte async) then read the ingest rank's HBM over  */
/* NVLink while they compute: no separate scatter step exists.                 */
/* ------------------------------------------------------------------------ */
#define OMB_PEER_HANDLE_BYTES 64
/* cudaMalloc + cudaIpcGetMemHandle on the current device. */
int omb_peer_alloc(size_t bytes, void** d_ptr, uint8_t handle[OMB_PEER_HANDLE_BYTES]);
/* cudaIpcOpenMemHandle in ANOTHER process (peer access is enabled lazily); the pointer is valid on the current device. */
int omb_peer_open(const uint8_t handle[OMB_PEER_HANDLE_BYTES], void** d_ptr);
int omb_peer_close(void* d_ptr);   /* importer side */
int omb_peer_free(void* d_ptr);    /* exporter side, after every importer has closed */
/* cudaMemcpyAsync(cudaMemcpyDefault): device, peer-mapped or pinned host pointers; stream may be NULL. */
int omb_copy_async(void* dst, const void* src, size_t bytes, void* cuda_stream);

/* ------------------------------------------------------------------------ */
/* Plan set-up pieces (rows a2,a3,a5,a14,a15 of SURVEY.md §8) — host-side,    */
/* exported so the parity tests can pin them against the oracle.              */
/* ------------------------------------------------------------------------ */

/* window.rs:20-43 — periodic cosine-sum window, f32. */
int omb_window_coefficients(int kind, size_t len, float* out);
/* window.rs:90-109 — out has fft_size/2+1 entries. */
int omb_fft_bin_normalization(const float* window, size_t window_len, size_t fft_size, float* out);
/* spectrogram/processor.rs:569-608 — derivative (dh) and time-ramp (t*h) windows. */
int omb_reassignment_windows(const float* window, size_t len, float* derivative, float* time_weighted);
/* spectrogram/processor.rs:111-117 */
float omb_reassigned_power_scale(const float* window, size_t len, size_t fft_size);
/* spectrogram/processor.rs:103-108 */
uint16_t omb_pack_classic_db(float db);
/* spectrum/processor.rs:410-425 */
float omb_a_weight(float freq_hz);
/* loudness/processor.rs:22-55 — b[5], a[5] of the 4th-order K-weighting section. */
int omb_k_weighting_coefficients(double sample_rate, double* b, double* a);
/* loudness/processor.rs:79-97 — factor 4: out[12*3] (tap-major), factor 2: out[24]. */
int omb_true_peak_fir(int factor, float* out);
/* dsp.rs:36-47 ChannelPosition::fallback(channels) -> positions[8]. */
void omb_fallback_positions(uint32_t channels, uint8_t positions[OMB_MAX_CHANNELS]);
/* dsp.rs:135-176 stereo fold-down matrix, out[8][2]. */
void omb_stereo_matrix(uint32_t channels, const uint8_t positions[OMB_MAX_CHANNELS], float out[OMB_MAX_CHANNELS][2]);

/* ------------------------------------------------------------------------ */
/* AudioBlock down-mix (row a1): dsp.rs:190-257, channel.rs:12-21             */
/* ------------------------------------------------------------------------ */

/* Folds `frames` interleaved frames of `channels` channels to one projected
 * mono lane on the GPU (host buffers in, host buffer out). Bit-exact with the
 * reference fold order. */
int omb_downmix_project(const float* interleaved, size_t frames, uint32_t channels,
                        const uint8_t positions[OMB_MAX_CHANNELS], int channel /*omb_channel*/,
                        float* out_lane);

/* ------------------------------------------------------------------------ */
/* SpectrogramProcessor (rows a4-a11)                                        */
/* ------------------------------------------------------------------------ */

/* spectrogram/processor.rs:45-56 SpectrogramConfig (defaults: 48000, 2048, 64, Hann, 0, true, 1) */
typedef struct omb_spectrogram_config {
  float sample_rate;
  uint32_t window; /* omb_window_kind */
  uint64_t fft_size;
  uint64_t hop_size;
  uint64_t history_length;
  uint64_t zero_padding_factor;
  int32_t use_reassignment;
  int32_t _pad;
} omb_spectrogram_config;

/* spectrogram/processor.rs:37-43 SpectrogramPoint, #[repr(C)] 12 bytes */
typedef struct omb_spectrogram_point {
  float time_offset;
  float freq_hz;
  float power;
} omb_spectrogram_point;

enum { OMB_COLUMN_REASSIGNED = 0, OMB_COLUMN_CLASSIC = 1 };

/* spectrogram/processor.rs:160-168 SpectrogramUpdate, flattened.
 * Column c of a reassigned update is points[column_offsets[c] .. column_offsets[c+1])
 * (ascending bin order); of a classic update, classic_db[c*bins .. (c+1)*bins). */
typedef struct omb_spectrogram_update {
  uint64_t fft_size; /* window * zero_padding_factor */
  uint64_t hop_size;
  uint64_t history_length;
  float sample_rate;
  float reassigned_power_scale;
  int32_t reset;
  int32_t kind; /* OMB_COLUMN_* */
  uint32_t n_columns;
  uint32_t bins;
  const uint32_t* column_offsets;        /* n_columns + 1 */
  const omb_spectrogram_point* points;   /* reassigned only */
  const uint16_t* classic_db;            /* classic only */
} omb_spectrogram_update;

typedef struct omb_spectrogram omb_spectrogram;

void omb_spectrogram_default_config(omb_spectrogram_config* out);
/* ::new (processor.rs:188) — normalises the config, never rejects it (processor.rs:71-82). */
int omb_spectrogram_create(const omb_spectrogram_config* cfg, omb_spectrogram** out);
void omb_spectrogram_destroy(omb_spectrogram* h);
/* ::config (processor.rs:208) */
int omb_spectrogram_get_config(const omb_spectrogram* h, omb_spectrogram_config* out);
/* ::update_config (processor.rs:518-543) */
int omb_spectrogram_update_config(omb_spectrogram* h, const omb_spectrogram_config* cfg);
/* ::prepare (processor.rs:219-223) */
int omb_spectrogram_prepare(omb_spectrogram* h);
/* ::reset_audio (processor.rs:212-217) */
int omb_spectrogram_reset_audio(omb_spectrogram* h);
/* ::process_block (processor.rs:490-516). `n_samples` = interleaved sample count
 * (AudioBlock.samples.len()); positions may be NULL => ChannelPosition::fallback. */
int omb_spectrogram_process_block(omb_spectrogram* h, const float* samples, size_t n_samples,
                                  uint32_t channels, float sample_rate,
                                  const uint8_t positions[OMB_MAX_CHANNELS],
                                  omb_spectrogram_update* out);

/* ------------------------------------------------------------------------ */
/* SpectrumProcessor (rows a12-a14) + peak_bin (row f3)                      */
/* ------------------------------------------------------------------------ */

/* spectrum/processor.rs:39-51 SpectrumConfig (defaults: 48000, 16384, 1024, Hann, None, Mid, None, -100) */
typedef struct omb_spectrum_config {
  float sample_rate;
  uint32_t window;          /* omb_window_kind */
  uint64_t fft_size;
  uint64_t hop_size;
  uint32_t averaging;       /* omb_averaging_mode */
  float averaging_param;    /* factor / decay_per_second */
  uint32_t source;          /* omb_channel */
  uint32_t secondary_source;/* omb_channel */
  float floor_db;
  int32_t _pad;
} omb_spectrum_config;

/* spectrum/processor.rs:33-37 SpectrumSnapshot; traces[t][0]=weighted, [t][1]=raw. */
typedef struct omb_spectrum_snapshot {
  uint32_t bins;
  int32_t _pad;
  const float* frequency_bins;
  const float* traces[2][2];
} omb_spectrum_snapshot;

typedef struct omb_spectrum omb_spectrum;

void omb_spectrum_default_config(omb_spectrum_config* out);
int omb_spectrum_create(const omb_spectrum_config* cfg, omb_spectrum** out);       /* processor.rs:89 */
void omb_spectrum_destroy(omb_spectrum* h);
int omb_spectrum_get_config(const omb_spectrum* h, omb_spectrum_config* out);      /* :108 */
int omb_spectrum_update_config(omb_spectrum* h, const omb_spectrum_config* cfg);   /* :300-322 */
int omb_spectrum_prepare(omb_spectrum* h);                                          /* :120-124 */
int omb_spectrum_reset_audio(omb_spectrum* h);                                      /* :112-118 */
int omb_spectrum_process_block(omb_spectrum* h, const float* samples, size_t n_samples,
                               uint32_t channels, float sample_rate,
                               const uint8_t positions[OMB_MAX_CHANNELS],
                               omb_spectrum_snapshot* out);                         /* :255-269 */

/* ------------------------------------------------------------------------ */
/* LoudnessProcessor (rows a15-a19)                                          */
/* ------------------------------------------------------------------------ */

typedef struct omb_loudness_config {
  float sample_rate; /* default 48000 */
  float floor_db;    /* default -99.9 */
} omb_loudness_config;

/* loudness/processor.rs:185-194 LoudnessSnapshot */
typedef struct omb_loudness_snapshot {
  float short_term_loudness;
  float momentary_loudness;
  float rms_fast_db[OMB_MAX_CHANNELS];
  float rms_slow_db[OMB_MAX_CHANNELS];
  float true_peak_db[OMB_MAX_CHANNELS];
  uint32_t channel_count;
  uint8_t positions[OMB_MAX_CHANNELS];
} omb_loudness_snapshot;

typedef struct omb_loudness omb_loudness;

void omb_loudness_default_config(omb_loudness_config* out);
int omb_loudness_create(const omb_loudness_config* cfg, omb_loudness** out);       /* processor.rs:225 */
void omb_loudness_destroy(omb_loudness* h);
int omb_loudness_get_config(const omb_loudness* h, omb_loudness_config* out);
int omb_loudness_reset_audio(omb_loudness* h);                                      /* :234 */
int omb_loudness_process_block(omb_loudness* h, const float* samples, size_t n_samples,
                               uint32_t channels, float sample_rate,
                               const uint8_t positions[OMB_MAX_CHANNELS],
                               omb_loudness_snapshot* out);                         /* :253-311 */

/* ------------------------------------------------------------------------ */
/* Batched offline entry points — what the streaming calls are wrappers over. */
/* Unit of work = (lane, frame).  `_device` variants take device pointers and */
/* a cudaStream_t (as void*); `_host` variants take host pointers and include */
/* the H2D / D2H copies.                                                      */
/* ------------------------------------------------------------------------ */

/* Plans own their scratch and host-path staging buffers: ONE execute call in flight per plan at a time (calls on different
 * plans, or on different streams with different plans, are independent).  A plan is bound to the device that was current when
 * it was created; every execute entry point selects that device itself. */
typedef struct omb_stft_plan omb_stft_plan;

/* Columns produced for a lane of `samples` samples: processor.rs:294-299. */
uint64_t omb_stft_frames_per_lane(const omb_spectrogram_config* cfg, uint64_t samples);

/* Which kernel family a plan may use. AUTO picks the specialised sm_100a
 * kernel when one exists for the size, else the generic one. */
enum { OMB_KERNEL_AUTO = 0, OMB_KERNEL_GENERIC = 1, OMB_KERNEL_FAST = 2 };

int omb_stft_plan_create(const omb_spectrogram_config* cfg, int kernel_choice, omb_stft_plan** out);
void omb_stft_plan_destroy(omb_stft_plan* p);
/* bins = fft_size*zp/2+1 */
uint32_t omb_stft_plan_bins(const omb_stft_plan* p);
/* 0: generic kernel; > 0: generation of the specialised sm_100a kernel the plan resolved to
 * (1 = stft_fast.cu, 2 = stft_fast2.cu, 7 = stft_r64.cu: reassigned N = 4096; 3 = stft_classic_fast.cu: classic N = 1024;
 * 4 = stft_fast8k.cu: N = 8192; 5 = stft_fast2k.cu: N = 2048; 6 = stft_fast1k.cu: N = 1024; 8 = stft_r64x.cu: N = 16384 / 8192). */
int omb_stft_plan_is_fast(const omb_stft_plan* p);
float omb_stft_plan_power_scale(const omb_stft_plan* p);

/* lanes: n_lanes planar mono lanes, lane l at lanes + l*lane_stride (floats).
 * Reassigned: out_points[(l*frames + f)*point_stride + i], i < out_counts[l*frames+f],
 *             ascending bin; point_stride >= bins (in points).
 * Classic:    out_classic[(l*frames + f)*bins + k].
 * Exactly one of out_points / out_classic is used, by cfg.use_reassignment. */
int omb_stft_execute_device(omb_stft_plan* p, const float* d_lanes, uint32_t n_lanes,
                            uint64_t samples_per_lane, uint64_t lane_stride,
                            omb_spectrogram_point* d_out_points, uint64_t point_stride,
                            uint32_t* d_out_counts, uint16_t* d_out_classic, void* cuda_stream);
int omb_stft_execute_host(omb_stft_plan* p, const float* h_lanes, uint32_t n_lanes,
                          uint64_t samples_per_lane, uint64_t lane_stride,
                          omb_spectrogram_point* h_out_points, uint64_t point_stride,
                          uint32_t* h_out_counts, uint16_t* h_out_classic);

typedef struct omb_spectrum_plan omb_spectrum_plan;

/* Hops produced for a lane: hops = samples >= N ? (samples-N)/hop+1 : 0. */
uint64_t omb_spectrum_hops_per_lane(const omb_spectrum_config* cfg, uint64_t samples);
int omb_spectrum_plan_create(const omb_spectrum_config* cfg, omb_spectrum_plan** out);
void omb_spectrum_plan_destroy(omb_spectrum_plan* p);
/* Every lane is one trace with its own smoothing state (zero-initialised).
 * out_weighted/out_raw: [(l*hops + h)*bins + k] f32; out_peak_bin (may be NULL):
 * [(l*hops+h)] peak_bin of the plan's peak spec (below; default: A-weighted trace, 20 Hz..Nyquist), last max wins
 * (state.rs:321-325), -1 if none. */
int omb_spectrum_execute_device(omb_spectrum_plan* p, const float* d_lanes, uint32_t n_lanes,
                                uint64_t samples_per_lane, uint64_t lane_stride,
                                float* d_out_weighted, float* d_out_raw, int32_t* d_out_peak_bin,
                                void* cuda_stream);
int omb_spectrum_execute_host(omb_spectrum_plan* p, const float* h_lanes, uint32_t n_lanes,
                              uint64_t samples_per_lane, uint64_t lane_stride,
                              float* h_out_weighted, float* h_out_raw, int32_t* h_out_peak_bin);

/* Row f3 — the peak label of the spectrum view (spectrum/state.rs:98-140, 180-205, 321-356).
 *
 * peak_bin(bins, db, min_f, max_f) (state.rs:321-325): arg-max of the selected trace over bins 1..bins-2 whose
 * frequency lies in [min_hz, max_hz] and whose dB value is finite; the LAST maximum wins (Iterator::max_by with
 * total_cmp). apply_snapshot (state.rs:106-107,134-136) calls it with min_f = MIN_FREQUENCY = 20 Hz,
 * max_f = frequency_bins[last].max(min_f * 1.02) on trace_db(traces[primary], weighting_mode), whose default is
 * A-weighted (visuals.rs:98). That is the default spec of every plan; the arg-max itself is fused into the
 * smoothing epilogue (out_peak_bin of omb_spectrum_execute_*). */
typedef struct omb_spectrum_peak_spec {
  uint32_t trace;  /* 0 = A-weighted trace (default), 1 = raw trace: SpectrumWeightingMode, state.rs:426-431 */
  float min_hz;    /* default 20 (state.rs:21) */
  float max_hz;    /* <= 0 (default): frequency_bins[last].max(min_hz * 1.02) (state.rs:107) */
} omb_spectrum_peak_spec;
void omb_spectrum_default_peak_spec(omb_spectrum_peak_spec* out);
int omb_spectrum_plan_set_peak_spec(omb_spectrum_plan* p, const omb_spectrum_peak_spec* spec);
int omb_spectrum_plan_get_peak_spec(const omb_spectrum_plan* p, omb_spectrum_peak_spec* out);

/* interpolated_peak(bins, db, bin) (state.rs:327-356): parabolic refinement of a peak bin from its two
 * neighbours, in the reference's f32 operation order. d_db: one dB trace [rows][bins] (the one the peak bins were
 * taken from); d_peak_bin: [rows]. Outputs [rows]: frequency in Hz (>= 0) and level in dB; both NaN where the
 * reference returns None (bin < 1, bin + 1 >= bins, non-finite centre value). */
int omb_spectrum_interpolate_peaks_device(omb_spectrum_plan* p, const float* d_db, const int32_t* d_peak_bin,
                                          uint64_t rows, float* d_out_freq_hz, float* d_out_level_db,
                                          void* cuda_stream);
/* omb_spectrum_execute_host + the interpolated peaks of the plan's peak spec, taken on the device before the
 * traces are copied back. Any of the three peak outputs may be NULL. */
int omb_spectrum_execute_host_peaks(omb_spectrum_plan* p, const float* h_lanes, uint32_t n_lanes,
                                    uint64_t samples_per_lane, uint64_t lane_stride,
                                    float* h_out_weighted, float* h_out_raw, int32_t* h_out_peak_bin,
                                    float* h_out_peak_freq_hz, float* h_out_peak_level_db);

typedef struct omb_loudness_plan omb_loudness_plan;

int omb_loudness_plan_create(const omb_loudness_config* cfg, uint32_t channels,
                             const uint8_t positions[OMB_MAX_CHANNELS], omb_loudness_plan** out);
void omb_loudness_plan_destroy(omb_loudness_plan* p);
/* n_streams interleaved streams of `frames` frames x `channels`; stream s at
 * base + s*stream_stride (floats). One snapshot per `block_frames` frames
 * (the process_block cadence, meter.rs:15-18): out[s*n_blocks + b],
 * n_blocks = ceil(frames / block_frames). */
int omb_loudness_execute_device(omb_loudness_plan* p, const float* d_interleaved, uint32_t n_streams,
                                uint64_t frames, uint64_t stream_stride, uint64_t block_frames,
                                omb_loudness_snapshot* d_out, void* cuda_stream);
int omb_loudness_execute_host(omb_loudness_plan* p, const float* h_interleaved, uint32_t n_streams,
                              uint64_t frames, uint64_t stream_stride, uint64_t block_frames,
                              omb_loudness_snapshot* h_out);


/* ------------------------------------------------------------------------ */
/* Rows f1 / f4 of SURVEY.md §8: the ordered audio timeline either side of   */
/* the processors.  Host-side state machines (no device work of their own);  */
/* the processors they feed keep their device-resident FIFOs.                */
/* ------------------------------------------------------------------------ */

/* dsp.rs:79-85 AudioFormat. Two formats are equal when every field is (PartialEq derive). */
typedef struct omb_audio_format {
  uint32_t channels;
  float sample_rate;
  uint64_t generation;
  uint8_t positions[OMB_MAX_CHANNELS];
} omb_audio_format;

/* infra/pipewire/transport.rs:39-54 CapturedSpan */
enum { OMB_SPAN_PCM = 0, OMB_SPAN_SILENCE = 1, OMB_SPAN_RESET = 2 };
/* kind PCM: samples/n_samples valid until the callback returns; SILENCE: frames; RESET: nothing else. */
typedef void (*omb_span_fn)(void* user, int kind, const float* samples, size_t n_samples, uint64_t frames,
                            const omb_audio_format* format);

/* Row f4 — AudioReader's packet timeline (transport.rs:573-657): packets carry [start_ns, end_ns) on the capture
 * clock; a gap before a packet becomes a Silence span, an overlap with what was already delivered is skipped,
 * PCM is coalesced in a scratch vector until flushed.  The lock-free queue, epochs and fault watchdog around it
 * are PipeWire plumbing and stay out of scope. */
typedef struct omb_timeline omb_timeline;
int omb_timeline_create(const omb_audio_format* initial_format, omb_timeline** out);
void omb_timeline_destroy(omb_timeline* t);
/* ::accept (transport.rs:573-625). samples == NULL is a silence packet of `frames` frames. */
int omb_timeline_accept(omb_timeline* t, const float* samples, uint64_t frames, const omb_audio_format* format,
                        uint64_t start_ns, uint64_t end_ns, omb_span_fn consume, void* user);
/* ::flush (transport.rs:634-643) */
int omb_timeline_flush(omb_timeline* t, omb_span_fn consume, void* user);
/* ::reset_timeline (transport.rs:645-656): drops the scratch, moves the cursor, aligns to the next packet. */
int omb_timeline_reset(omb_timeline* t, uint64_t cursor_ns);
uint64_t omb_timeline_cursor(const omb_timeline* t);
size_t omb_timeline_pending_samples(const omb_timeline* t);

/* Row f1 — DspBatcher (meter.rs:27-84) + ingest_silence (meter.rs:143-165) + VisualManager::ingest_samples
 * (visuals/registry.rs:396-418) for the three hot-path processors.  After every ingest the callback receives the
 * chunk that was ingested and the three processors' outputs (NULL where a processor is not attached or returned
 * None); the pointers are library-owned and valid until the callback returns. */
typedef struct omb_meter omb_meter;
typedef void (*omb_ingest_fn)(void* user, const float* samples, size_t n_samples, const omb_audio_format* format,
                              const omb_spectrogram_update* spectrogram, const omb_spectrum_snapshot* spectrum,
                              const omb_loudness_snapshot* loudness);
int omb_meter_create(omb_meter** out);
void omb_meter_destroy(omb_meter* m);
/* Borrowed handles (any may be NULL = module disabled); they must outlive the meter. */
int omb_meter_attach(omb_meter* m, omb_spectrogram* spectrogram, omb_spectrum* spectrum, omb_loudness* loudness);
int omb_meter_set_callback(omb_meter* m, omb_ingest_fn fn, void* user);
/* DspBatcher::push (meter.rs:40-73); *n_ingests (may be NULL) = its return value. */
int omb_meter_push(omb_meter* m, const float* samples, size_t n_samples, const omb_audio_format* format, uint32_t* n_ingests);
/* ingest_silence (meter.rs:143-165): more than 2 s of silence resets instead of replaying zeros. */
int omb_meter_push_silence(omb_meter* m, uint64_t frames, const omb_audio_format* format, uint32_t* n_ingests);
/* DspBatcher::reset (meter.rs:75-78): clear + reset_audio on every attached processor. */
int omb_meter_reset(omb_meter* m);
/* DspBatcher::clear (meter.rs:80-83) */
int omb_meter_clear(omb_meter* m);
/* MeterEngine::advance's span dispatch (meter.rs:115-124): feed one CapturedSpan. */
int omb_meter_consume_span(omb_meter* m, int kind, const float* samples, size_t n_samples, uint64_t frames,
                           const omb_audio_format* format, uint32_t* n_ingests);
size_t omb_meter_pending_samples(const omb_meter* m);
/* 1 when the batcher currently holds a format (DspBatcher.format.is_some()). */
int omb_meter_has_format(const omb_meter* m);


/* ------------------------------------------------------------------------ */
/* Row f2 of SURVEY.md §8: what consumes the reassigned points — the splat   */
/* accumulation of the spectrogram view, as a CUDA scatter-add.              */
/* ------------------------------------------------------------------------ */

enum { OMB_FREQ_LINEAR = 0, OMB_FREQ_LOG = 1, OMB_FREQ_ERB = 2 }; /* util/audio/frequency.rs:25-31 */

/* The subset of spectrogram/render.rs:187-252 `Uniforms` the accumulation and resolve passes read. The palette /
 * rotation / clip transform of the final composite stay with the GUI: images here are in accumulation space. */
typedef struct omb_splat_params {
  uint32_t freq_scale;          /* OMB_FREQ_* */
  float freq_min, freq_max;     /* display axis in Hz (spectrogram/state.rs:49-52) */
  float uv_y_range[2];          /* zoom / pan window into the [0,1] frequency axis */
  float ext_w, ext_h;           /* spectrogram.wgsl extents(): widget size in physical pixels (swapped when rotated) */
  float scale_factor;           /* >= 1; one column = scale_factor pixels, splats are scale_factor^2 */
  float tilt_db;                /* dB / octave re 1 kHz; 0 = off */
  uint32_t ring_capacity;       /* history_length: slots in the point ring */
  uint32_t newest_col;          /* (write_slot + ring_capacity - 1) % ring_capacity */
  uint32_t col_count;           /* columns received so far; min(col_count, ring_capacity) slots are drawn */
  float reassigned_power_scale; /* omb_spectrogram_update.reassigned_power_scale */
} omb_splat_params;

/* Accumulation-texture size (render.rs:511-531 resize_accum with rotation folded into ext): ceil(max(ext, 1)). */
void omb_splat_image_size(const omb_splat_params* p, uint32_t* width, uint32_t* height);

/* vs_accum_splat + fs_accum with additive blending (spectrogram.wgsl:126-147,215-225): every point of every drawn
 * slot adds its (tilted) power to the pixels its scale_factor-sized quad covers.  d_rings: n_rings point rings laid
 * out [ring][slot][point_stride]; d_slot_counts[ring][slot] = points in the slot (clamped to point_stride);
 * d_accum: n_rings images [height][width] f32, cleared by this call (LoadOp::Clear).  The Rg16Float dual-scale
 * trick (wgsl:4-7) exists only because of the f16 attachment and is not reproduced: accumulation is f32. */
int omb_splat_accumulate_device(const omb_spectrogram_point* d_rings, uint64_t point_stride, const uint32_t* d_slot_counts,
                                uint32_t n_rings, const omb_splat_params* p, float* d_accum, void* cuda_stream);
/* fs_resolve (wgsl:227-237) up to the palette: dB = max(ln(max(power*scale, 1e-20)) * LN_TO_DB, -140), and -inf
 * where nothing was accumulated (the transparent pixels). */
int omb_splat_resolve_device(const float* d_accum, uint32_t n_rings, const omb_splat_params* p, float* d_db, void* cuda_stream);
/* Host-pointer convenience (H2D of the ring, both passes, D2H of the images). h_accum may be NULL. */
int omb_splat_render_host(const omb_spectrogram_point* h_rings, uint64_t point_stride, const uint32_t* h_slot_counts,
                          uint32_t n_rings, const omb_splat_params* p, float* h_accum, float* h_db);

/* The reassigned STFT chained into the splat passes on the device: what the reference does per frame tick when
 * SpectrogramUpdate.new_columns go into the point ring and are drawn (spectrogram/render.rs:557-598 -> wgsl:126-147,
 * 215-237).  Every lane is one view whose ring holds all of the lane's columns: `view`'s ring_capacity / col_count /
 * newest_col / reassigned_power_scale are SET by the call (frames per lane, last frame newest, the plan's power scale);
 * its axis, extent, scale_factor and tilt are the caller's.  h_db: n_lanes images [height][width] f32 dB (-inf where nothing
 * was drawn), omb_splat_image_size(view); h_out_counts (may be NULL): [lane][frame] points per column.  Only PCM goes
 * up and only images (+ counts) come back: the 24.6 KB of points per column never cross PCIe.  Lane chunks are pipelined
 * (H2D / kernels / D2H overlap).  Reassigned plans only; at most 65535 columns per lane per call. */
int omb_stft_render_host(omb_stft_plan* p, const float* h_lanes, uint32_t n_lanes, uint64_t samples_per_lane,
                         uint64_t lane_stride, const omb_splat_params* view, float* h_db, uint32_t* h_out_counts);


/* ------------------------------------------------------------------------ */
/* Row f1, device side: the multi-stream ring.  S independent streams that    */
/* share one spectrogram config and advance in lock-step (one capture clock,  */
/* e.g. the DspBatcher cadence): every push appends one block to EVERY stream */
/* in a device-resident ring [stream][pending] and all newly ready columns of */
/* all streams come out of ONE launch of the batched kernel.  Column for      */
/* column identical to S separate omb_spectrogram handles fed the same blocks.*/
/* ------------------------------------------------------------------------ */
typedef struct omb_spectrogram_bank omb_spectrogram_bank;

/* Dense form of S SpectrogramUpdates (processor.rs:160-168): every stream has n_columns new columns. */
typedef struct omb_spectrogram_bank_update {
  uint64_t fft_size, hop_size, history_length;
  float sample_rate, reassigned_power_scale;
  int32_t reset;
  int32_t kind;                         /* OMB_COLUMN_* */
  uint32_t n_streams, n_columns, bins, _pad;
  const uint32_t* counts;               /* reassigned: [stream][column] points in the column */
  const omb_spectrogram_point* points;  /* reassigned: [stream][column][bins] slots, the first counts[..] are valid (ascending bin) */
  const uint16_t* classic_db;           /* classic:    [stream][column][bins] */
} omb_spectrogram_bank_update;

int omb_spectrogram_bank_create(const omb_spectrogram_config* cfg, uint32_t n_streams, omb_spectrogram_bank** out);
void omb_spectrogram_bank_destroy(omb_spectrogram_bank* b);
/* ::reset_audio of every stream (processor.rs:212-217) */
int omb_spectrogram_bank_reset_audio(omb_spectrogram_bank* b);
/* ::process_block of every stream (processor.rs:490-516) with one block each: stream s reads `frames` interleaved frames
 * of `channels` channels at samples + s * stream_stride (floats).  OMB_NO_DATA when no stream has a new column.
 * Output pointers are library-owned pinned host memory, valid until the next call. */
int omb_spectrogram_bank_push(omb_spectrogram_bank* b, const float* samples, uint64_t stream_stride, size_t frames,
                              uint32_t channels, float sample_rate, const uint8_t positions[OMB_MAX_CHANNELS],
                              omb_spectrogram_bank_update* out);
/* Samples pending per stream (the VecDeque length of processor.rs:412-437). */
size_t omb_spectrogram_bank_pending(const omb_spectrogram_bank* b);

/* ------------------------------------------------------------------------ */
/* Row f1, device side, loudness: S lock-step LoudnessProcessors               */
/* (loudness/processor.rs:218-311) sharing config, channel layout and rate.   */
/* Filter / window / true-peak state of all streams lives on the device; one   */
/* push = one strided H2D copy + ONE kernel launch (a CTA per stream) + one    */
/* D2H copy of S snapshots, instead of S x (copy + launch + sync).             */
/* Snapshot for snapshot identical to S separate omb_loudness handles.         */
/* ------------------------------------------------------------------------ */
typedef struct omb_loudness_bank omb_loudness_bank;
int omb_loudness_bank_create(const omb_loudness_config* cfg, uint32_t n_streams, omb_loudness_bank** out);
void omb_loudness_bank_destroy(omb_loudness_bank* b);
/* ::reset_audio of every stream (processor.rs:234-236) */
int omb_loudness_bank_reset_audio(omb_loudness_bank* b);
/* ::process_block of every stream (processor.rs:253-311): stream s reads n_samples interleaved samples at
 * samples + s * stream_stride (floats); out_snapshots[s] (caller-owned, n_streams entries) receives its snapshot.
 * OMB_NO_DATA when the block holds less than one frame. */
int omb_loudness_bank_push(omb_loudness_bank* b, const float* samples, uint64_t stream_stride, size_t n_samples,
                           uint32_t channels, float sample_rate, const uint8_t positions[OMB_MAX_CHANNELS],
                           omb_loudness_snapshot* out_snapshots);

/* ------------------------------------------------------------------------ */
/* Row f1, device side, spectrum analyzer: S lock-step SpectrumProcessors      */
/* (spectrum/processor.rs:88-323) with one config.  Pending audio of every     */
/* stream's traces in one device ring [stream * traces][pending], smoothing    */
/* state resident on the device; one push = one strided H2D copy, one batched  */
/* fold-down, one power and one smoothing launch for all streams.              */
/* Trace for trace identical to S separate omb_spectrum handles.               */
/* ------------------------------------------------------------------------ */
typedef struct omb_spectrum_bank omb_spectrum_bank;

/* Dense form of S SpectrumSnapshots (processor.rs:31-37), active traces only (processor.rs:174-177). */
typedef struct omb_spectrum_bank_snapshot {
  uint32_t bins, n_streams, n_traces;   /* n_traces = 1 or 2 */
  uint32_t trace_index[2];              /* reference trace (0 = source, 1 = secondary_source) of bank trace i */
  const float* frequency_bins;          /* [bins] */
  const float* weighted;                /* [stream][trace][bins] A-weighted dB */
  const float* raw;                     /* [stream][trace][bins] dB */
} omb_spectrum_bank_snapshot;

int omb_spectrum_bank_create(const omb_spectrum_config* cfg, uint32_t n_streams, omb_spectrum_bank** out);
void omb_spectrum_bank_destroy(omb_spectrum_bank* b);
/* ::reset_audio of every stream (processor.rs:112-118) */
int omb_spectrum_bank_reset_audio(omb_spectrum_bank* b);
/* ::process_block of every stream (processor.rs:255-269) with one block each: stream s reads `frames` interleaved frames of
 * `channels` channels at samples + s * stream_stride (floats).  OMB_NO_DATA when no hop completed.  Output pointers are
 * library-owned pinned host memory, valid until the next call. */
int omb_spectrum_bank_push(omb_spectrum_bank* b, const float* samples, uint64_t stream_stride, size_t frames,
                           uint32_t channels, float sample_rate, const uint8_t positions[OMB_MAX_CHANNELS],
                           omb_spectrum_bank_snapshot* out);
/* Samples pending per trace (the VecDeque length of processor.rs:271-298). */
size_t omb_spectrum_bank_pending(const omb_spectrum_bank* b);

#ifdef __cplusplus
}
#endif
#endif /* OMB200_H */
