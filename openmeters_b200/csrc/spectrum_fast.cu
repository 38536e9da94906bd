// spectrum_fast.cu — specialised power-spectrum kernel for the spectrum analyzer's default size N = 16384 (BASELINE cfg4).
//
// The real 16384-point transform is one complex 8192-point FFT of (r[2n], r[2n+1]); that FFT is split radix-2
// (decimation in time) into TWO independent 4096-point FFTs which the two 256-thread groups of a CTA run side by
// side with the shared radix-16 engine (fft4096.cuh).  The radix-2 combine, the real-FFT split and |X|^2 * norm are
// fused into one epilogue that turns each quadruple (E[a], O[a], E[4096-a], O[4096-a]) into the four bins
// a, 4096-a, 4096+a, 8192-a:
//     Z[a] = E + w8 O, Z[a+4096] = E - w8 O                       (w8 = W_8192^a)
//     X[k] = A + W_16384^k B,  X[8192-k] = conj(A - W_16384^k B),  A = (P + conj Q)/2, B = (P - conj Q)/(2j)
// DC removal uses per-hop f64 block sums from a small pre-kernel (mean = sum of N/hop block sums / N), so the PCM is
// read once by the FFT kernel.  Rows a4, a12 of SURVEY.md §8; spectrum/processor.rs:215-244.
#include "fft4096.cuh"
#include "spectrum.h"

namespace omb {

namespace {

using namespace f4k;

constexpr int kN = 16384;
constexpr int kM = 4096;                  // sub-transform length
constexpr int kThreads = 2 * kT;
constexpr int kWSize = f16::phys_size(kM);

struct SpecFastArgs {
  SpectrumPowerArgs a;
  const float2* tw1;     // [15][256] W_4096^{b q} (rows 0,1,3,7 used)
  const float2* tw2;     // [15][16]  W_256^{o q}
  const float2* w16;     // [2049]    W_16384^a
  const double* bsum;    // [lane][n_blocks] f64 sums of hop-sized blocks
  uint64_t n_blocks;
  uint32_t blocks_per_frame;  // N / hop
};

struct SmemS {
  float2 tw1[4 * kT];
  float2 tw2[15 * 16];
  float2 W[2][kWSize];
  float mean;
};

// one warp per (lane, block): f64 sum of `hop` samples
__global__ void __launch_bounds__(256) k_block_sums(const float* lanes, uint64_t lane_stride, uint32_t n_lanes, uint64_t n_blocks,
                                                    uint32_t hop, double* out) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane_id = threadIdx.x & 31;
  const uint64_t total = n_blocks * n_lanes;
  double acc = 0.0;
  if (warp < total) {
    const uint64_t l = warp / n_blocks, b = warp % n_blocks;
    const float* x = lanes + l * lane_stride + b * hop;
    for (uint32_t i = lane_id; i < hop; i += 32) acc += (double)__ldg(&x[i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (warp < total && lane_id == 0) out[warp] = acc;
}

__device__ __forceinline__ int posC(int k) { return 273 * (k & 15) + 17 * ((k >> 4) & 15) + (k >> 8); }

__global__ void __launch_bounds__(kThreads, 1) k_spectrum_power_16k(SpecFastArgs fa) {
  OMB_DYN_SMEM(unsigned char, smem_raw);
  SmemS& sm = *reinterpret_cast<SmemS*>(smem_raw);
  const SpectrumPowerArgs& a = fa.a;
  const int tid = threadIdx.x, t = tid & (kT - 1);
  const int g = __shfl_sync(0xffffffffu, tid >> 8, 0);
  for (int i = tid; i < 4 * kT; i += kThreads) {
    const int row = (1 << (i >> 8)) - 1;
    sm.tw1[i] = __ldg(&fa.tw1[row * kT + (i & (kT - 1))]);
  }
  for (int i = tid; i < 15 * 16; i += kThreads) sm.tw2[i] = __ldg(&fa.tw2[i]);
  Addr ad;
  ad.pA = t + (t >> 4);
  ad.pB = 273 * (t >> 4) + (t & 15);
  ad.pC = 273 * (t & 15) + 17 * (t >> 4);
  const float2* tw1t = sm.tw1 + t;
  const float2* tw2o = sm.tw2 + (t & 15);
  __syncthreads();

  const uint64_t total = a.hops * a.n_lanes;
  for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const uint64_t lane = item / a.hops, h = item % a.hops;
    const float* x = a.lanes + lane * a.lane_stride + h * a.hop;
    if (tid == 0) {
      const double* bs = fa.bsum + lane * fa.n_blocks + h;
      double s = 0.0;
      for (uint32_t i = 0; i < fa.blocks_per_frame; ++i) s += bs[i];
      sm.mean = (float)(s / (double)kN);
    }
    __syncthreads();
    const float mean = sm.mean;
    // group g transforms z[2m + g], z[n] = (r[2n], r[2n+1]); thread t owns m = t + 256 j -> samples 4m + 2g, +1
    float2 v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int i0 = 4 * (t + kT * j) + 2 * g;
      const float2 xv = __ldg(reinterpret_cast<const float2*>(x + i0));
      const float2 wv = __ldg(reinterpret_cast<const float2*>(a.win + i0));
      v[j] = make_float2((xv.x - mean) * wv.x, (xv.y - mean) * wv.y);
    }
    fft_forward<f16::kAll, 0>(v, sm.W[g], tw1t, tw2o, ad, g);
    float2* wc = sm.W[g] + ad.pC;
#pragma unroll
    for (int q = 0; q < 16; ++q) wc[q] = v[q];
    __syncthreads();
    // epilogue over a = 0..2048
    float* out = a.power + item * a.bins;
    for (int aa = tid; aa <= kM / 2; aa += kThreads) {
      const int bb = kM - aa;
      const float2 Ea = sm.W[0][posC(aa)], Oa = sm.W[1][posC(aa)];
      const float2 Eb = sm.W[0][posC(bb & (kM - 1))], Ob = sm.W[1][posC(bb & (kM - 1))];
      const float2 w = __ldg(&fa.w16[aa]);        // W_16384^a
      const float2 w8 = cmul(w, w);               // W_8192^a
      const float2 ta = cmul(w8, Oa);
      const float2 tb = cmul(make_float2(-w8.x, w8.y), Ob);  // W_8192^{4096-a} = -conj(w8)
      const float2 Za = cadd(Ea, ta), Za2 = csub(Ea, ta);
      const float2 Zb = cadd(Eb, tb), Zb2 = csub(Eb, tb);
      // pair 1: P = Z[a], Q = Z[8192-a] = Zb2, twiddle W_16384^a -> bins a and 8192-a
      {
        const float2 A = make_float2(0.5f * (Za.x + Zb2.x), 0.5f * (Za.y - Zb2.y));
        const float2 d = make_float2(Za.x - Zb2.x, Za.y + Zb2.y);          // P - conj Q
        const float2 B = make_float2(0.5f * d.y, -0.5f * d.x);             // d / (2j)
        const float2 T = cmul(w, B);
        const float2 X0 = cadd(A, T), X1 = csub(A, T);
        out[aa] = (X0.x * X0.x + X0.y * X0.y) * __ldg(&a.bin_norm[aa]);
        out[2 * kM - aa] = (X1.x * X1.x + X1.y * X1.y) * __ldg(&a.bin_norm[2 * kM - aa]);
      }
      // pair 2: P = Z[b], Q = Z[4096+a] = Za2, twiddle W_16384^{4096-a} = -j conj(w) -> bins 4096-a and 4096+a
      {
        const float2 A = make_float2(0.5f * (Zb.x + Za2.x), 0.5f * (Zb.y - Za2.y));
        const float2 d = make_float2(Zb.x - Za2.x, Zb.y + Za2.y);
        const float2 B = make_float2(0.5f * d.y, -0.5f * d.x);
        const float2 wk = make_float2(-w.y, -w.x);                          // -j * conj(w), w = (c, -s) -> (s... ) see below
        const float2 T = cmul(wk, B);
        const float2 X0 = cadd(A, T), X1 = csub(A, T);
        out[bb] = (X0.x * X0.x + X0.y * X0.y) * __ldg(&a.bin_norm[bb]);
        out[kM + aa] = (X1.x * X1.x + X1.y * X1.y) * __ldg(&a.bin_norm[kM + aa]);
      }
    }
    __syncthreads();
  }
}

}  // namespace

bool spectrum_fast_supported(const SpectrumConfigN& cfg, const DeviceInfo& dev) {
  if (cfg.fft_size != (uint64_t)kN) return false;
  if (cfg.hop == 0 || (kN % cfg.hop) != 0 || (cfg.hop & 1)) return false;
  return dev.max_smem_optin == 0 || sizeof(SmemS) + 256 <= (size_t)dev.max_smem_optin;
}

int spectrum_fast_prepare(SpectrumPlan& p) {
  std::vector<float2> tab(15 * kT + 15 * 16 + (kM / 2 + 1));
  const double tau = 6.28318530717958647692;
  for (int q = 1; q < 16; ++q)
    for (int b = 0; b < kT; ++b) {
      const double ang = -tau * (double)((b * q) % kM) / (double)kM;
      tab[(q - 1) * kT + b] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  for (int q = 1; q < 16; ++q)
    for (int o = 0; o < 16; ++o) {
      const double ang = -tau * (double)((o * q) % 256) / 256.0;
      tab[15 * kT + (q - 1) * 16 + o] = make_float2((float)std::cos(ang), (float)std::sin(ang));
    }
  for (int a = 0; a <= kM / 2; ++a) {
    const double ang = -tau * (double)a / (double)kN;
    tab[15 * kT + 15 * 16 + a] = make_float2((float)std::cos(ang), (float)std::sin(ang));
  }
  OMB_TRY(p.d_fast_tables.upload(tab, p.stream));
  OMB_CUDA_TRY(cudaFuncSetAttribute(k_spectrum_power_16k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemS)));
  return OMB_OK;
}

int spectrum_fast_block_sums(SpectrumPlan& p, const float* d_lanes, uint64_t lane_stride, uint32_t n_lanes, uint64_t hops, cudaStream_t s) {
  const uint64_t n_blocks = hops - 1 + (uint64_t)kN / p.cfg.hop;
  OMB_TRY(p.d_bsum.reserve((size_t)(n_blocks * n_lanes)));
  const uint64_t warps = n_blocks * n_lanes;
  OMB_LAUNCH(k_block_sums, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, s, d_lanes, lane_stride, n_lanes, n_blocks,
             (uint32_t)p.cfg.hop, p.d_bsum.ptr);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

int launch_spectrum_power_fast(SpectrumPlan& p, SpectrumPowerArgs& a, cudaStream_t s) {
  const uint64_t total = a.hops * a.n_lanes;
  if (!total) return OMB_OK;
  if ((reinterpret_cast<uintptr_t>(a.lanes) & 7u) != 0 || (a.lane_stride & 1))
    return fail(OMB_ERR_INVALID, "specialised spectrum kernel needs 8-byte aligned lanes");
  SpecFastArgs fa{};
  fa.a = a;
  fa.tw1 = p.d_fast_tables.ptr;
  fa.tw2 = fa.tw1 + 15 * kT;
  fa.w16 = fa.tw2 + 15 * 16;
  fa.blocks_per_frame = (uint32_t)(kN / a.hop);
  if (p.ext_bsum_blocks) {  // batch-wide sums computed by spectrum_fast_block_sums(); this launch starts at ext_block_off
    fa.n_blocks = p.ext_bsum_blocks;
    fa.bsum = p.d_bsum.ptr + p.ext_block_off;
  } else {
    OMB_TRY(spectrum_fast_block_sums(p, a.lanes, a.lane_stride, a.n_lanes, a.hops, s));
    fa.n_blocks = a.hops - 1 + fa.blocks_per_frame;
    fa.bsum = p.d_bsum.ptr;
  }
  const unsigned grid = (unsigned)std::min<uint64_t>(total, (uint64_t)std::max(p.dev.sm_count, 1));
  OMB_LAUNCH(k_spectrum_power_16k, dim3(grid), dim3(kThreads), sizeof(SmemS), s, fa);
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
