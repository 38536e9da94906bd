// Compile-and-run check of include/omb200.hpp (the C++ host mirror) against libomb200.so: host-only logic, no GPU.
// Restates meter.rs:221-232 (dsp_batches_coalesce_large_capture_backlogs) and transport.rs:729-770 through the mirror.
#include <cstdio>
#include <vector>

#include "omb200.hpp"

#define EXPECT(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main() {
  omb::AudioFormat f{};
  f.channels = 2;
  f.sample_rate = 48000.0f;
  f.generation = 1;
  omb_fallback_positions(2, f.positions);
  omb::DspBatcher b(nullptr, nullptr, nullptr);
  std::vector<float> x((256 * 6 + 17) * 2, 0.25f);
  EXPECT(b.push(x.data(), x.size(), f) == 2);
  EXPECT(b.pending_samples() == 17 * 2);
  EXPECT(b.push(x.data(), 239 * 2, f) == 1);
  EXPECT(b.pending_samples() == 0);

  omb::AudioFormat m = f;
  m.channels = 1;
  m.sample_rate = 1000.0f;
  omb::PacketTimeline t(m);
  std::vector<omb::CapturedSpan> out;
  const float p[4] = {1, 1, 1, 1};
  const uint64_t ms = 1000000ull;
  t.accept(p, 4, m, 0 * ms, 4 * ms, out);
  t.accept(p, 4, m, 6 * ms, 10 * ms, out);
  t.accept(p, 4, m, 8 * ms, 12 * ms, out);
  t.flush(out);
  EXPECT(out.size() == 3);
  EXPECT(out[0].kind == OMB_SPAN_PCM && out[0].samples.size() == 4);
  EXPECT(out[1].kind == OMB_SPAN_SILENCE && out[1].frames == 2);
  EXPECT(out[2].kind == OMB_SPAN_PCM && out[2].samples.size() == 6);
  std::printf("ok\n");
  return 0;
}
