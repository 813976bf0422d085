// acro_views.cuh - how kernels see weights, references and SoA trajectories.
#pragma once
#include <cstdint>

#include "acro_device.cuh"

namespace acro {

// Kernel-side copy of AcroWeights.  For shared weights the values sit in the kernel
// parameter (constant) bank, so they cost no registers: DFMA takes them as c[][] operands.
struct KWeights {
  double Q[16], R[4], QT[16];
  double Q2[16], R2[4], QT2[16];  // 2Q, 2R, 2Q_T: the Hessian blocks of tg:100-101,112
  const double *Qb, *Rb, *QTb;    // per-problem overrides [e][B] or nullptr
};

template <bool WPB>
struct WV;

template <>
struct WV<false> {
  const KWeights* k;
  __device__ __forceinline__ WV(const KWeights& kw, int64_t, int64_t) : k(&kw) {}
  __device__ __forceinline__ double Q(int i, int j) const { return k->Q[i * 4 + j]; }
  __device__ __forceinline__ double R(int i, int j) const { return k->R[i * 2 + j]; }
  __device__ __forceinline__ double QT(int i, int j) const { return k->QT[i * 4 + j]; }
  __device__ __forceinline__ double Q2(int i, int j) const { return k->Q2[i * 4 + j]; }
  __device__ __forceinline__ double R2(int i, int j) const { return k->R2[i * 2 + j]; }
  __device__ __forceinline__ double QT2(int i, int j) const { return k->QT2[i * 4 + j]; }
};

template <>
struct WV<true> {
  double q[10], r[3], qt[10];
  __device__ __forceinline__ WV(const KWeights& kw, int64_t B, int64_t b) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = i; j < 4; ++j) {
        q[sym(i, j)] = kw.Qb ? kw.Qb[(i * 4 + j) * B + b] : kw.Q[i * 4 + j];
        qt[sym(i, j)] = kw.QTb ? kw.QTb[(i * 4 + j) * B + b] : kw.QT[i * 4 + j];
      }
    r[0] = kw.Rb ? kw.Rb[0 * B + b] : kw.R[0];
    r[1] = kw.Rb ? kw.Rb[1 * B + b] : kw.R[1];
    r[2] = kw.Rb ? kw.Rb[3 * B + b] : kw.R[3];
  }
  __device__ __forceinline__ double Q(int i, int j) const { return q[sym(i, j)]; }
  __device__ __forceinline__ double R(int i, int j) const { return r[i + j]; }
  __device__ __forceinline__ double QT(int i, int j) const { return qt[sym(i, j)]; }
  __device__ __forceinline__ double Q2(int i, int j) const { return 2.0 * q[sym(i, j)]; }
  __device__ __forceinline__ double R2(int i, int j) const { return 2.0 * r[i + j]; }
  __device__ __forceinline__ double QT2(int i, int j) const { return 2.0 * qt[sym(i, j)]; }
};

// functors handed to riccati_step
template <class W>
struct QhQ2 {
  const W& w;
  __device__ __forceinline__ double operator()(int i, int j) const { return w.Q2(i, j); }
};
template <class W>
struct QhQ {
  const W& w;
  __device__ __forceinline__ double operator()(int i, int j) const { return w.Q(i, j); }
};

// Time-indexed batch arrays are tiled structure-of-arrays, tile-major:
//     A[tile][t][c][lane],   tile = b / 32, lane = b % 32,   T time steps, C components
//     element (t, c, b)  ->  A[((tile*T + t)*C + c)*32 + lane]
// The whole trajectory of a warp's 32 problems is one contiguous block: the C rows of a time step are C*256
// consecutive bytes (component c at the compile-time offset c*256), consecutive time steps follow each other
// (compile-time stride C*256), so a pass streams through memory with one pointer per array and any number of
// consecutive steps can be moved by ONE bulk copy.  `T` is the number of time steps the array holds.
__device__ __forceinline__ int64_t padded(int64_t B) { return (B + 31) & ~int64_t(31); }
__device__ __forceinline__ int64_t soa(int t, int C, int c, int64_t T, int64_t b) {
  return (((b >> 5) * T + t) * C + c) * 32 + (b & 31);
}

// Reference trajectory: shared (N,4)/(N-1,2) row-major, or per problem in the tiled layout (x: T = N, u: T = N-1).
template <bool RPB>
struct RefV {
  const double* x;
  const double* u;
  int64_t N, b;
  __device__ __forceinline__ double X(int t, int c) const { return RPB ? x[soa(t, 4, c, N, b)] : __ldg(x + t * 4 + c); }
  __device__ __forceinline__ double U(int t, int c) const { return RPB ? u[soa(t, 2, c, N - 1, b)] : __ldg(u + t * 2 + c); }
};

// (v' W v) for a symmetric 4x4 / 2x2 given through an accessor
template <class F>
__device__ __forceinline__ double quad4(const double v[4], F W) {
  double acc = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    double s = v[0] * W(0, j);
#pragma unroll
    for (int i = 1; i < 4; ++i) s = fma(v[i], W(i, j), s);
    acc = fma(s, v[j], acc);
  }
  return acc;
}
template <class F>
__device__ __forceinline__ double quad2(const double v[2], F W) {
  const double s0 = fma(v[1], W(1, 0), v[0] * W(0, 0));
  const double s1 = fma(v[1], W(1, 1), v[0] * W(0, 1));
  return fma(s1, v[1], s0 * v[0]);
}

}  // namespace acro
