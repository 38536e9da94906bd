// loudness.h — BS.1770 K-weighted loudness / true-peak.
#pragma once
#include "common.h"
#include "tables.h"

namespace omb {

constexpr int kLoudWindows = 4;        // short-term 3 s, momentary 0.4 s, rms fast 0.3 s, rms slow 1 s
constexpr int kKwChunk = 256;          // samples per chunk of the chunk-parallel IIR
constexpr int kKwSeg = 64;             // chunks per segment of the three-level state scan

struct KWeight { double b[5], a[5]; };
struct TruePeakFir { float fir4[12][3]; float fir2[24]; };

// ---- streaming (exact, sequential per channel) state, one per channel, device resident
struct LoudChannelState {
  double filter[4];
  double sums[kLoudWindows][2];
  double corr[kLoudWindows][2];
  unsigned long long refresh[kLoudWindows];
  unsigned long long head, count;        // WindowedMeans ring cursor / fill (dsp.rs:298-305)
  unsigned long long silent_frames;
  float delay[48];
  unsigned int write;
  float peak;
  int active;
  int _pad;
};

struct LoudStreamArgs {
  const float* block;                    // interleaved [frames][channels]; stream s at block + s * block_stride
  uint64_t block_stride;                 // floats between streams (0 for a single stream)
  uint64_t frames;
  uint32_t channels;
  LoudChannelState* state;               // [stream][channels]
  double* ring;                          // [stream][channels][ring_len] squared K-weighted samples
  uint64_t ring_len;                     // longest window
  uint64_t caps[kLoudWindows];
  uint32_t tp_delay_len;                 // 12 (4x), 24 (2x) or 0
  KWeight kw;
  TruePeakFir fir;
  float floor_db;
  double weights[OMB_MAX_CHANNELS];      // channel_weight(position)
  uint8_t positions[OMB_MAX_CHANNELS];
  omb_loudness_snapshot* out;            // [stream]
  double* vnew;                          // [stream][channels][frames] scratch: this block's squared K-weighted samples
};

struct LoudnessStreamCore {              // device objects behind omb_loudness (stream_loudness.cu)
  DeviceBuffer<LoudChannelState> d_state;
  DeviceBuffer<double> d_ring;
  DeviceBuffer<float> d_block;
  DeviceBuffer<omb_loudness_snapshot> d_snap;
  DeviceBuffer<double> d_vnew;           // scratch of the phase-parallel streaming kernel
};
int launch_loudness_stream(const LoudStreamArgs& a, cudaStream_t s, uint32_t n_streams = 1);

// ---- batched plan
struct LoudnessPlan {
  omb_loudness_config cfg;
  float sample_rate = kDefaultSampleRate;
  uint32_t channels = 0;
  uint8_t positions[OMB_MAX_CHANNELS];
  DeviceInfo dev;
  KWeight kw;
  TruePeakFir fir;
  double chunk_matrix[16];               // A^kKwChunk, row-major (state transition over one chunk)
  double seg_matrix[16];                 // A^(kKwChunk*kKwSeg)
  std::vector<double> h_resp;            // [kKwChunk][4]: A^(kKwChunk-1-k) B, the end state's response to input sample k
  uint64_t caps[kLoudWindows];
  uint32_t tp_delay_len = 0;
  DeviceBuffer<double> d_end, d_start, d_csum, d_cbase, d_seg;
  DeviceBuffer<float> d_y;
  DeviceBuffer<unsigned> d_peak;
  DeviceBuffer<float> d_in;
  DeviceBuffer<omb_loudness_snapshot> d_out;
  cudaStream_t stream = nullptr;
  // batch path: the true-peak kernel (FP32-bound, needs only the PCM) runs on `side` next to the K-weighting chain (FP64 / DRAM-bound)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  ~LoudnessPlan();
  int init(const omb_loudness_config& c, uint32_t channels, const uint8_t* positions);
  int execute_device(const float* d_interleaved, uint32_t n_streams, uint64_t frames, uint64_t stream_stride, uint64_t block_frames,
                     omb_loudness_snapshot* d_out, cudaStream_t s);
  int execute_host(const float* h_interleaved, uint32_t n_streams, uint64_t frames, uint64_t stream_stride, uint64_t block_frames,
                   omb_loudness_snapshot* h_out);
};

double channel_weight_host(uint8_t position);  // loudness/processor.rs:174-183

}  // namespace omb
