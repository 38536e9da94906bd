// fft16.cuh — in-register radix-16 butterfly and the shared-memory layout helpers of the
// specialised STFT kernels.
//
// A length-M complex FFT (M = 256*R3, here R3 = 16 -> M = 4096) is three in-place radix-16 passes over a
// padded shared-memory buffer, T = M/16 threads, 16 elements per thread per pass:
//
//   access A (stride T)  : thread b owns logical indices b + T*j            (DIF pass 1 / DIT pass 3)
//   access B (stride 16) : thread u = (blk, o) owns blk*T + o + 16*j        (pass 2 in both directions)
//   access C (contiguous): thread t owns base(t)*16 + j, base(t) = digit swap so that after a DIF
//                          transform thread t holds frequencies t + 256*j   (DIF pass 3 / DIT pass 1)
//
// phys(i) = i + i/16 + i/256 (in float2 units) makes all three patterns bank-conflict free for 64-bit
// accesses (each half-warp touches 16 distinct 8-byte bank pairs).
#pragma once
#include "common.h"

namespace omb {
namespace f16 {

constexpr float kC1 = 0.92387953251128673848f;  // cos(pi/8)
constexpr float kS1 = 0.38268343236508978178f;  // sin(pi/8)
constexpr float kH = 0.70710678118654752440f;   // sqrt(1/2)

// cos / sin of 2*pi*j/32, j = 0..15
__device__ constexpr float kCos32[16] = {1.0f, 0.98078528040323044913f, 0.92387953251128673848f, 0.83146961230254523708f,
                                         0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508978178f, 0.19509032201612826785f,
                                         0.0f, -0.19509032201612826785f, -0.38268343236508978178f, -0.55557023301960222474f,
                                         -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128673848f, -0.98078528040323044913f};
__device__ constexpr float kSin32[16] = {0.0f, 0.19509032201612826785f, 0.38268343236508978178f, 0.55557023301960222474f,
                                         0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128673848f, 0.98078528040323044913f,
                                         1.0f, 0.98078528040323044913f, 0.92387953251128673848f, 0.83146961230254523708f,
                                         0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508978178f, 0.19509032201612826785f};

// Packed FP32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2: both lanes of an aligned register pair in ONE issue slot).
// The SASS operand modifiers make every complex primitive of an FFT a one- or two-instruction sequence with no
// register shuffling: `.F32x2.LO_HI` swaps the lanes of a source pair, `.NP` negates one lane, `.F32` broadcasts a
// scalar register (or an immediate) to both lanes.  So
//     a + b, a - b                       1 FADD2
//     a -+ j b   (radix-4 rotation)      1 FADD2  (b.LO_HI.NP)              — two scalar FADDs before
//     a * w      (complex)               1 FMUL2 (a, w.x bcast) + 1 FFMA2 (a.LO_HI, w.y bcast with .NP, acc)  — four before
// A packed instruction holds the FMA pipe for two cycles, so the pipe time is unchanged; what halves is the ISSUE cost.
// ptxas folds the make_float2(...) swizzles below into those modifiers (checked with cuobjdump: no MOV / PRMT).
// MEASURED (B200, profiles/r02b_packed_ab.md): packing the additions pays (round 1), packing the PRODUCTS does not — the
// radix-16 kernels lose 7-13 % (static instructions 3440 -> 2944 for the cfg2 kernel, yet slower): they are bound by the
// FMA pipe inside the butterfly phases, not by issue slots, and the two-instruction packed product is a longer dependent
// chain than the four scalar ones it replaces.  So the switches of common.h default to packed adds only; the packed
// product forms stay here behind OMB_F32X2_MUL / OMB_F32X2_ROT for A/B builds.
// (switches: common.h.)
__device__ __forceinline__ float2 cadd2(float2 a, float2 b) {
#if OMB_F32X2_ADD
  return __fadd2_rn(a, b);
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
__device__ __forceinline__ float2 csub2(float2 a, float2 b) {
#if OMB_F32X2_ADD
  return __fadd2_rn(a, make_float2(-b.x, -b.y));
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
// a + (-j) b = (a.x + b.y, a.y - b.x)   and   a + (+j) b = (a.x - b.y, a.y + b.x)
__device__ __forceinline__ float2 cadd_mj(float2 a, float2 b) {
#if OMB_F32X2_ROT
  return __fadd2_rn(a, make_float2(b.y, -b.x));
#else
  return make_float2(a.x + b.y, a.y - b.x);
#endif
}
__device__ __forceinline__ float2 cadd_pj(float2 a, float2 b) {
#if OMB_F32X2_ROT
  return __fadd2_rn(a, make_float2(-b.y, b.x));
#else
  return make_float2(a.x - b.y, a.y + b.x);
#endif
}
// a * (wx + j wy)
__device__ __forceinline__ float2 cmul2(float2 a, float wx, float wy) {
#if OMB_F32X2_MUL
  return __ffma2_rn(make_float2(a.y, a.x), make_float2(-wy, wy), __fmul2_rn(a, make_float2(wx, wx)));
#else
  return make_float2(a.x * wx - a.y * wy, a.y * wx + a.x * wy);
#endif
}
// (ax, ay) * s  (both lanes by one scalar)
__device__ __forceinline__ float2 cscale2(float2 a, float s) {
#if OMB_F32X2_MUL
  return __fmul2_rn(a, make_float2(s, s));
#else
  return make_float2(a.x * s, a.y * s);
#endif
}

// c conj(p) + s (j z) = (c p.x - s z.y, s z.x - c p.y): the Hilbert pair step of the reassigned kernels
__device__ __forceinline__ float2 pair_q(float2 p, float2 z, float c, float s) {
#if OMB_F32X2_MUL
  return __ffma2_rn(make_float2(-z.y, z.x), make_float2(s, s), __fmul2_rn(make_float2(p.x, -p.y), make_float2(c, c)));
#else
  return make_float2(c * p.x - s * z.y, s * z.x - c * p.y);
#endif
}

template <bool INV>
__device__ __forceinline__ float2 rot_mj(float2 a) {  // a * (-j) forward, a * (+j) inverse
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}
// a * (c - j s) forward, a * (c + j s) inverse
template <bool INV>
__device__ __forceinline__ float2 mul_cs(float2 a, float c, float s) {
  return INV ? cmul2(a, c, s) : cmul2(a, c, -s);
}
// a * w forward, a * conj(w) inverse, w = (cos, -sin) table entry
template <bool INV>
__device__ __forceinline__ float2 mul_tw(float2 a, float2 w) {
  return INV ? cmul2(a, w.x, -w.y) : cmul2(a, w.x, w.y);
}

template <bool INV>
__device__ __forceinline__ void radix4(float2& a0, float2& a1, float2& a2, float2& a3) {
  const float2 s0 = cadd2(a0, a2);
  const float2 s1 = csub2(a0, a2);
  const float2 s2 = cadd2(a1, a3);
  const float2 d = csub2(a1, a3);
  a0 = cadd2(s0, s2);
  a2 = csub2(s0, s2);
  // y1 = s1 + rot(d), y3 = s1 - rot(d), rot = multiply by -j (forward) / +j (inverse): lane swap + one negation,
  // both operand modifiers of the packed add
  if (INV) {
    a1 = cadd_pj(s1, d);
    a3 = cadd_mj(s1, d);
  } else {
    a1 = cadd_mj(s1, d);
    a3 = cadd_pj(s1, d);
  }
}

// v[q] <- sum_j v[j] * W16^{+-jq}; natural order in, natural order out.
template <bool INV>
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
  // step 1: radix-4 over j1 for each j0 (elements j0, j0+4, j0+8, j0+12) -> t[j0][q0] left in v[j0 + 4*q0]
#pragma unroll
  for (int j0 = 0; j0 < 4; ++j0) radix4<INV>(v[j0], v[j0 + 4], v[j0 + 8], v[j0 + 12]);
  // step 2: t[j0][q0] *= W16^{j0*q0}
  v[1 + 4 * 1] = mul_cs<INV>(v[1 + 4 * 1], kC1, kS1);                 // W^1
  v[1 + 4 * 2] = mul_cs<INV>(v[1 + 4 * 2], kH, kH);                   // W^2
  v[1 + 4 * 3] = mul_cs<INV>(v[1 + 4 * 3], kS1, kC1);                 // W^3
  v[2 + 4 * 1] = mul_cs<INV>(v[2 + 4 * 1], kH, kH);                   // W^2
  v[2 + 4 * 2] = rot_mj<INV>(v[2 + 4 * 2]);                           // W^4 = -j
  v[2 + 4 * 3] = mul_cs<INV>(v[2 + 4 * 3], -kH, kH);                  // W^6
  v[3 + 4 * 1] = mul_cs<INV>(v[3 + 4 * 1], kS1, kC1);                 // W^3
  v[3 + 4 * 2] = mul_cs<INV>(v[3 + 4 * 2], -kH, kH);                  // W^6
  v[3 + 4 * 3] = mul_cs<INV>(v[3 + 4 * 3], -kC1, -kS1);               // W^9
  // step 3: radix-4 over j0 for each q0: inputs v[j0 + 4*q0], outputs X[q0 + 4*q1] land in v[q1 + 4*q0]
#pragma unroll
  for (int q0 = 0; q0 < 4; ++q0) radix4<INV>(v[4 * q0], v[4 * q0 + 1], v[4 * q0 + 2], v[4 * q0 + 3]);
  // transpose the 4x4 index (q1 + 4*q0 -> q0 + 4*q1) so the result is in natural order
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) {
      const float2 t = v[a + 4 * b];
      v[a + 4 * b] = v[b + 4 * a];
      v[b + 4 * a] = t;
    }
}

// Pruned variants: same transform, but only a subset of the 16 outputs is produced (the others are left
// undefined).  kFirst9: outputs 0..8 (bins <= Nyquist of the analysis transforms);  kMid8: outputs 4..11
// (the centre half of the inverse transform).  Only step 3 differs: X[q0 + 4 q1] needs q1 in {0,1} (+ q1 = 2 for
// q0 = 0) resp. q1 in {1,2}.
enum Prune { kAll = 0, kFirst9 = 1, kMid8 = 2 };

template <bool INV, int kPrune>
__device__ __forceinline__ void radix4_part(float2& a0, float2& a1, float2& a2, float2& a3, bool want2) {
  const float2 s0 = cadd2(a0, a2);
  const float2 s1 = csub2(a0, a2);
  const float2 s2 = cadd2(a1, a3);
  const float2 d = csub2(a1, a3);
  if (kPrune == kFirst9) {
    a0 = cadd2(s0, s2);
    if (want2) a2 = csub2(s0, s2);
  } else {  // kMid8
    a2 = csub2(s0, s2);
  }
  a1 = INV ? cadd_pj(s1, d) : cadd_mj(s1, d);
}

template <bool INV, int kPrune>
__device__ __forceinline__ void dft16p(float2 (&v)[16]) {
  if (kPrune == kAll) {
    dft16<INV>(v);
    return;
  }
#pragma unroll
  for (int j0 = 0; j0 < 4; ++j0) radix4<INV>(v[j0], v[j0 + 4], v[j0 + 8], v[j0 + 12]);
  v[1 + 4 * 1] = mul_cs<INV>(v[1 + 4 * 1], kC1, kS1);
  v[1 + 4 * 2] = mul_cs<INV>(v[1 + 4 * 2], kH, kH);
  v[1 + 4 * 3] = mul_cs<INV>(v[1 + 4 * 3], kS1, kC1);
  v[2 + 4 * 1] = mul_cs<INV>(v[2 + 4 * 1], kH, kH);
  v[2 + 4 * 2] = rot_mj<INV>(v[2 + 4 * 2]);
  v[2 + 4 * 3] = mul_cs<INV>(v[2 + 4 * 3], -kH, kH);
  v[3 + 4 * 1] = mul_cs<INV>(v[3 + 4 * 1], kS1, kC1);
  v[3 + 4 * 2] = mul_cs<INV>(v[3 + 4 * 2], -kH, kH);
  v[3 + 4 * 3] = mul_cs<INV>(v[3 + 4 * 3], -kC1, -kS1);
#pragma unroll
  for (int q0 = 0; q0 < 4; ++q0) radix4_part<INV, kPrune>(v[4 * q0], v[4 * q0 + 1], v[4 * q0 + 2], v[4 * q0 + 3], q0 == 0);
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = a + 1; b < 4; ++b) {
      const float2 t = v[a + 4 * b];
      v[a + 4 * b] = v[b + 4 * a];
      v[b + 4 * a] = t;
    }
}

__device__ __forceinline__ int phys(int i) { return i + (i >> 4) + (i >> 8); }
constexpr int phys_size(int m) { return m + (m >> 4) + (m >> 8); }

}  // namespace f16
}  // namespace omb
