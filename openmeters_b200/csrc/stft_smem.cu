// stft_smem.cu — shared-memory STFT kernels for any power-of-two size that fits on chip.
//
// Middle tier between the specialised N = 4096 kernels (stft_fast*.cu) and the any-size global-scratch kernels
// (stft_generic.cu): one CTA per frame, Stockham radix-4 FFTs ping-ponging between two shared buffers
// (fft_stockham.cuh), natural-order spectra.
//   classic    : the real F-point transform runs as ONE complex F/2-point FFT of (r[2n], r[2n+1]) plus the
//                split X[k] = E[k] + W_F^k O[k]                                       (rows a4, a7)
//   reassigned : same restructuring as the specialised kernel — packed real forward FFT (N points), fused
//                Hilbert pair step, one inverse, three windowed F-point FFTs — instead of the literal
//                2 x 2N + 3 x F                                                        (rows a8-a10)
// Limits: on-chip complex lengths up to 8192 (classic F <= 16384; reassigned N*zp <= 8192) run as one transform.
// Longer ZERO-PADDED analysis lengths F = N*zp (N <= 8192) are split into R = F / 8192 residue classes,
//     X[R m + r] = sum_{n<N} (a[n] W_F^{n r}) W_8192^{n m} = FFT_8192(a[n] W_F^{n r})[m],
// i.e. R on-chip transforms of the modulated, zero-padded frame (same flops as one F-point transform, no global-memory
// passes); only N = 16384 with F > 16384 (classic) / any reassigned N = 16384 is left to the generic kernels.
#include "fft_stockham.cuh"
#include "stft.h"

namespace omb {

namespace {

constexpr int kMaxComplex = 8192;

__global__ void __launch_bounds__(1024) k_classic_smem(StftKernelArgs a) {
  OMB_DYN_SMEM(float2, smem);
  __shared__ float red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = (int)a.window, F = (int)a.fft_len, M = F >> 1, logM = (int)a.log2_fft - 1;
  float2* A = smem;
  float2* B = smem + M;
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total = per_lane * a.n_lanes;
  for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const uint64_t lane = item / per_lane, frame = a.first_frame + item % per_lane;
    const float* x = a.lanes + lane * a.lane_stride + frame * a.hop;
    float part = 0.0f;
    for (int i = tid; i < N; i += nt) part += __ldg(&x[i]);
    const float mean = block_sum(part, red) / (float)N;
    for (int n = tid; n < M; n += nt) {
      const int i0 = 2 * n, i1 = 2 * n + 1;
      const float r0 = i0 < N ? (__ldg(&x[i0]) - mean) * __ldg(&a.win[i0]) : 0.0f;
      const float r1 = i1 < N ? (__ldg(&x[i1]) - mean) * __ldg(&a.win[i1]) : 0.0f;
      A[n] = make_float2(r0, r1);
    }
    __syncthreads();
    const float2* Z = stockham_fft(A, B, M, logM, a.tw_fft, 2);  // W_M^i = W_F^{2i}
    uint16_t* out = a.out_classic + (lane * a.frames_per_lane + frame) * a.bins;
    for (int k = tid; k <= M; k += nt) {
      const float2 zk = Z[k & (M - 1)];
      const float2 zm = Z[(M - k) & (M - 1)];
      const float2 E = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
      const float2 O = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));  // (zk - conj zm) / (2j)
      const float2 w = k < M ? __ldg(&a.tw_fft[k]) : make_float2(-1.0f, 0.0f);
      const float2 X = cadd(E, cmul(w, O));
      const float p = (X.x * X.x + X.y * X.y) * __ldg(&a.bin_norm[k]);
      out[k] = classic_code_dev(p);
    }
    __syncthreads();
  }
}

// Classic column for F = N * zp > 16384 with N <= 8192: R = F / 8192 residue transforms of the real windowed frame
// (complex after the modulation).  Bins k = R m + r <= F/2 come from m <= (F/2 - r) / R of residue r.
__global__ void __launch_bounds__(1024) k_classic_residue(StftKernelArgs a, int R) {
  OMB_DYN_SMEM(float2, smem);
  __shared__ float red[32];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = (int)a.window, F = (int)a.fft_len, Fs = F / R, logFs = (int)a.log2_fft - (31 - __clz(R));
  float2* A = smem;
  float2* B = smem + Fs;
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total = per_lane * a.n_lanes;
  for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const uint64_t lane = item / per_lane, frame = a.first_frame + item % per_lane;
    const float* x = a.lanes + lane * a.lane_stride + frame * a.hop;
    float part = 0.0f;
    for (int i = tid; i < N; i += nt) part += __ldg(&x[i]);
    const float mean = block_sum(part, red) / (float)N;
    uint16_t* out = a.out_classic + (lane * a.frames_per_lane + frame) * a.bins;
    for (int r = 0; r < R; ++r) {
      for (int n = tid; n < Fs; n += nt) {
        float2 v = make_float2(0.0f, 0.0f);
        if (n < N) {
          const float rv = (__ldg(&x[n]) - mean) * __ldg(&a.win[n]);
          const int idx = (n * r) & (F - 1);  // W_F^{n r}; the table holds the upper half circle
          float2 w = __ldg(&a.tw_fft[idx & (F / 2 - 1)]);
          if (idx >= F / 2) w = make_float2(-w.x, -w.y);
          v = make_float2(rv * w.x, rv * w.y);
        }
        A[n] = v;
      }
      __syncthreads();
      const float2* Z = stockham_fft(A, B, Fs, logFs, a.tw_fft, R);  // W_Fs^i = W_F^{R i}
      for (int m = tid; m < Fs; m += nt) {
        const int k = R * m + r;
        if (k >= (int)a.bins) break;
        const float2 X = Z[m];
        out[k] = classic_code_dev((X.x * X.x + X.y * X.y) * __ldg(&a.bin_norm[k]));
      }
      __syncthreads();
    }
  }
}

struct SmemReassignScratch {
  float2* S;   // [bins]
  float* nd;   // [bins]
};

// nres = 1: the analysis transforms (F points) run on chip.  nres = R > 1 (F = N * zp > 8192): R residue transforms of
// Fs = F / R points per window, X[R m + r] = FFT_Fs(c[n] w[n] W_F^{n r})[m]; the third spectrum then also goes through the scratch.
template <bool kResidue>  // false: nres_arg == 1 and the residue machinery compiles away (the on-chip path as it was)
__global__ void __launch_bounds__(1024) k_reassigned_smem(StftKernelArgs a, float* gscratch, uint64_t gscratch_stride, int nres_arg) {
  const int nres = kResidue ? nres_arg : 1;
  OMB_DYN_SMEM(float2, smem);
  __shared__ int cnt[33];
  __shared__ float x0_xm[2];
  const int tid = threadIdx.x, nt = blockDim.x;
  const int N = (int)a.window, F = (int)a.fft_len, H = 2 * N, Fs = F / nres;
  const int logN = (int)a.log2_hilbert - 1, logF = (int)a.log2_fft - (31 - __clz(nres));
  const int off = (H - N) / 2;
  float2* A = smem;
  float2* B = smem + Fs;
  float* Y = reinterpret_cast<float*>(smem + 2 * Fs);  // N floats
  float2* Sg = reinterpret_cast<float2*>(gscratch + (uint64_t)blockIdx.x * gscratch_stride);
  float* ndg = reinterpret_cast<float*>(Sg + a.bins);
  float2* Tg = reinterpret_cast<float2*>(ndg + a.bins + (a.bins & 1));  // 8-byte aligned; used when nres > 1
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total = per_lane * a.n_lanes;
  const ReassignConsts rc{a.bin_hz, a.max_hz, a.inv_2pi, a.inv_hop, a.latency_hops};
  for (uint64_t item = blockIdx.x; item < total; item += gridDim.x) {
    const uint64_t lane = item / per_lane, frame = a.first_frame + item % per_lane;
    const float* x = a.lanes + lane * a.lane_stride + frame * a.hop;
    // F: packed real transform of the 2N-sample frame
    for (int n = tid; n < N; n += nt) A[n] = make_float2(__ldg(&x[2 * n]), __ldg(&x[2 * n + 1]));
    __syncthreads();
    float2* Z = stockham_fft(A, B, N, logN, a.tw_hil, 2);  // W_N^i = W_H^{2i}
    float2* other = (Z == A) ? B : A;
    if (tid == 0) {
      x0_xm[0] = Z[0].x + Z[0].y;
      x0_xm[1] = Z[0].x - Z[0].y;
    }
    // X: Q[k] = cos(th) conj(Z[N-k]) + j sin(th) Z[k], th = 2 pi k / H; written as conj(Q) for the inverse
    for (int k = tid; k <= N / 2; k += nt) {
      if (k == 0) {
        Z[0] = make_float2(0.0f, 0.0f);
        continue;
      }
      const float2 w = __ldg(&a.tw_hil[k]);  // (cos, -sin)
      const float c = w.x, s = -w.y;
      const float2 zk = Z[k], zm = Z[N - k];
      const float2 qk = make_float2(c * zm.x - s * zk.y, s * zk.x - c * zm.y);
      const float2 qm = make_float2(-c * zk.x - s * zm.y, s * zm.x + c * zk.y);
      Z[k] = make_float2(qk.x, -qk.y);
      if (k != N - k) Z[N - k] = make_float2(qm.x, -qm.y);
    }
    __syncthreads();
    // I: q = conj(FFT(conj Q)); keep y[off + n]: Y[2m'] = Re q[m], Y[2m'+1] = Im q[m], m = off/2 + m'
    const float2* R = stockham_fft(Z, other, N, logN, a.tw_hil, 2);
    for (int n = tid; n < N; n += nt) {
      const float2 q = R[(off + n) >> 1];
      Y[n] = (n & 1) ? -q.y : q.x;
    }
    __syncthreads();
    const float half_x0 = 0.5f * x0_xm[0], half_xm = 0.5f * x0_xm[1];
    // G: three windowed F-point transforms (R residue transforms of Fs points each when F does not fit on chip)
    const float2* T = nullptr;
    for (int r = 0; r < nres; ++r) {
      for (int wsel = 0; wsel < 3; ++wsel) {
        const float* win = wsel == 1 ? a.dwin : a.win;
        for (int n = tid; n < Fs; n += nt) {
          float2 v = make_float2(0.0f, 0.0f);
          if (n < N) {
            float wv = __ldg(&win[n]);
            if (wsel == 2) wv *= (float)n - (float)(N - 1) * 0.5f;
            const float bias = ((n & 1) ? -half_xm : half_xm) - half_x0;  // off is even for every N >= 4
            const float cx = fmaf((float)N, __ldg(&x[off + n]), bias);
            v = make_float2(cx * wv, Y[n] * wv);
            if (r > 0) {  // modulation W_F^{n r}; the table holds the upper half circle
              const int idx = (n * r) & (F - 1);
              float2 w = __ldg(&a.tw_fft[idx & (F / 2 - 1)]);
              if (idx >= F / 2) w = make_float2(-w.x, -w.y);
              v = cmul(v, w);
            }
          }
          A[n] = v;
        }
        __syncthreads();
        const float2* Rw = stockham_fft(A, B, Fs, logF, a.tw_fft, nres);  // W_Fs^i = W_F^{nres i}
        if (nres == 1) {
          if (wsel == 0) {
            for (int k = tid; k < (int)a.bins; k += nt) Sg[k] = Rw[k];
          } else if (wsel == 1) {
            for (int k = tid; k < (int)a.bins; k += nt) {
              const float2 s = Sg[k], d = Rw[k];
              ndg[k] = d.y * s.x - d.x * s.y;
            }
          } else {
            T = Rw;
          }
        } else {
          for (int m = tid; m < Fs; m += nt) {
            const int k = nres * m + r;
            if (k >= (int)a.bins) break;
            if (wsel == 0) {
              Sg[k] = Rw[m];
            } else if (wsel == 1) {
              const float2 s = Sg[k], d = Rw[m];
              ndg[k] = d.y * s.x - d.x * s.y;
            } else {
              Tg[k] = Rw[m];
            }
          }
          T = Tg;
        }
        __syncthreads();
      }
    }
    // R
    const uint64_t slot = lane * a.frames_per_lane + frame;
    omb_spectrogram_point* out = a.out_points + slot * a.point_stride;
    int base = 0;
    for (int k0 = 0; k0 < (int)a.bins; k0 += nt) {
      const int k = k0 + tid;
      omb_spectrogram_point p;
      bool keep = false;
      if (k < (int)a.bins) keep = reassign_bin_nd(Sg[k], ndg[k], T[k], __ldg(&a.bin_norm[k]), k, rc, &p);
      int tot;
      const int rank = block_rank(keep, cnt, &tot);
      if (keep) out[base + rank] = p;
      base += tot;
    }
    if (tid == 0) a.out_counts[slot] = (uint32_t)base;
    __syncthreads();
  }
}

}  // namespace

// Residue classes of the analysis transform: 1 when it fits on chip, else F / 8192 (needs the frame itself on chip: N <= 8192).
static int residues_for(const StftConfig& cfg) {
  const uint64_t F = cfg.fft_len();
  const uint64_t on_chip = cfg.reassign ? (uint64_t)kMaxComplex : 2ull * kMaxComplex;  // classic packs the real frame into F/2 points
  return F <= on_chip ? 1 : (int)(F / (uint64_t)kMaxComplex);
}

static size_t smem_for(const StftConfig& cfg) {
  const uint64_t N = cfg.window, F = cfg.fft_len();
  const uint64_t Fs = F / (uint64_t)residues_for(cfg);
  if (cfg.reassign) return (size_t)(2 * Fs * sizeof(float2) + N * sizeof(float));
  return residues_for(cfg) == 1 ? (size_t)(F * sizeof(float2)) /* two buffers of F/2 */ : (size_t)(2 * Fs * sizeof(float2));
}

bool stft_smem_supported(const StftConfig& cfg, const DeviceInfo& dev) {
  const uint64_t N = cfg.window, F = cfg.fft_len();
  if (!is_pow2(N) || !is_pow2(F)) return false;
  if (cfg.reassign ? N < 8 : F < 16) return false;
  if (residues_for(cfg) > 1 && (N > (uint64_t)kMaxComplex || F > (1ull << 24))) return false;  // the frame must fit one residue transform
  if (cfg.reassign && N > (uint64_t)kMaxComplex) return false;                                   // so must the Hilbert stage
  return dev.max_smem_optin == 0 || smem_for(cfg) + 1024 <= (size_t)dev.max_smem_optin;
}

int stft_smem_prepare(StftPlan& plan) {
  const int smem = (int)smem_for(plan.cfg);
  if (plan.cfg.reassign) {
    if (residues_for(plan.cfg) == 1) OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_smem<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    else OMB_CUDA_TRY(cudaFuncSetAttribute(k_reassigned_smem<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  } else if (residues_for(plan.cfg) == 1) {
    OMB_CUDA_TRY(cudaFuncSetAttribute(k_classic_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  } else {
    OMB_CUDA_TRY(cudaFuncSetAttribute(k_classic_residue, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  return OMB_OK;
}

int launch_stft_smem(const StftPlan& plan, StftKernelArgs& a, cudaStream_t s, DeviceBuffer<float2>& scratch) {
  const uint64_t per_lane = a.frames_per_lane - a.first_frame;
  const uint64_t total = per_lane * a.n_lanes;
  if (total == 0) return OMB_OK;
  const size_t smem = smem_for(plan.cfg);
  const int R = residues_for(plan.cfg);
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (size_t)200 * 1024 / std::max<size_t>(smem, 1)));
  const unsigned grid = (unsigned)std::min<uint64_t>(total, (uint64_t)std::max(plan.dev.sm_count, 1) * per_sm);
  // one radix-4 butterfly per thread per stage when possible; when shared memory limits the SM to one or two
  // CTAs, make them wide (up to 1024 threads) so the SM still has 16-32 warps to hide latency
  const uint64_t work = (plan.cfg.reassign || R > 1) ? a.fft_len / (uint64_t)R : a.fft_len / 2;
  // whole warps only: block_sum / block_rank use full-mask warp collectives (2048 / 6 = 341 threads broke them)
  const uint64_t want = std::max<uint64_t>(256, (2048 / (uint64_t)per_sm + 31) / 32 * 32);
  const unsigned threads = (unsigned)std::min<uint64_t>(std::min<uint64_t>(1024, want), std::max<uint64_t>(32, work / 4));
  if (plan.cfg.reassign) {
    // floats per CTA: S (2 per bin) + nd (1 per bin, padded to an even count) + T (2 per bin, residue mode only)
    const uint64_t stride = R == 1 ? 3ull * a.bins + 1 : 5ull * a.bins + (a.bins & 1);
    OMB_TRY(scratch.reserve((size_t)((stride * grid + 1) / 2)));
    if (R == 1) {
      OMB_LAUNCH(k_reassigned_smem<false>, dim3(grid), dim3(threads), smem, s, a, reinterpret_cast<float*>(scratch.ptr), stride, 1);
    } else {
      OMB_LAUNCH(k_reassigned_smem<true>, dim3(grid), dim3(threads), smem, s, a, reinterpret_cast<float*>(scratch.ptr), stride, R);
    }
  } else if (R == 1) {
    OMB_LAUNCH(k_classic_smem, dim3(grid), dim3(threads), smem, s, a);
  } else {
    OMB_LAUNCH(k_classic_residue, dim3(grid), dim3(threads), smem, s, a, R);
  }
  OMB_CHECK_LAUNCH();
  return OMB_OK;
}

}  // namespace omb
