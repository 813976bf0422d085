// acro_kernels.cu - sm_100a kernels and the C ABI of libacro_b200.so (see include/acro_abi.h).
//
// Thread mapping of the kernels in this file: one thread = one independent problem (lane-per-problem), so every
// global access of a warp is 32 consecutive doubles of one SoA row (256 B, coalesced) and no thread ever waits on
// another.  The time recurrences (RK4 rollout, Riccati sweep) are sequential per problem; the next step's operands
// are prefetched into registers while the current step computes.
// The Newton / Armijo loop has two more implementations, chosen by batch size in acro_newton_solve:
//   acro_newton_ring.cuh  one warp per tile of 32 problems, warp-synchronous, operands through a TMA-fed ring
//   acro_newton_duo.cuh   two warps per tile (recurrence warp + trailer warp) for batches that leave SMs idle
// and acro_mpc_box.cuh holds the box-constrained MPC tracker.  There is no collective anywhere on this path.
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/acro_abi.h"
#include "acro_device.cuh"
#include "acro_views.cuh"

#ifndef ACRO_ROLLOUT_MINB
#define ACRO_ROLLOUT_MINB 3  // resident blocks of 128 threads per SM the closed-loop rollout kernels are compiled for
#endif

namespace acro {

// ---------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
static int cuda_fail(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return ACRO_E_CUDA;
}
#define ACRO_REQUIRE(cond, msg) \
  if (!(cond)) return fail(ACRO_E_INVALID, msg)
#define ACRO_LAUNCH_CHECK(name)                             \
  do {                                                      \
    g_launches.fetch_add(1, std::memory_order_relaxed);     \
    cudaError_t e__ = cudaPeekAtLastError();                \
    if (e__ != cudaSuccess) return cuda_fail(e__, name);    \
  } while (0)

static Model make_model(const AcroParams& p) {
  Model m;
  m.a1 = p.I1 + p.I2 + p.lc1 * p.lc1 * p.m1 + p.m2 * (p.l1 * p.l1 + p.lc2 * p.lc2);
  m.h = p.m2 * p.l1 * p.lc2;
  m.a3 = p.I2 + p.lc2 * p.lc2 * p.m2;
  m.g1 = p.g * (p.lc1 * p.m1 + p.m2 * p.l1);
  m.g2 = p.g * p.m2 * p.lc2;
  m.f1 = p.f1;
  m.f2 = p.f2;
  m.dt = p.dt;
  m.tau1 = p.actuated_tau1 ? 1.0 : 0.0;
  m.h2 = 2.0 * m.h;
  m.det0 = m.a1 * m.a3 - m.a3 * m.a3;
  m.hsq = m.h * m.h;
  const double tc[16] = ACRO_TRIG_CONSTANTS;
  for (int i = 0; i < 16; ++i) m.tc[i] = tc[i];
  return m;
}

static KWeights make_weights(const AcroWeights& w) {
  KWeights k;
  for (int i = 0; i < 16; ++i) {
    k.Q[i] = w.Q[i];
    k.QT[i] = w.QT[i];
    k.Q2[i] = 2.0 * w.Q[i];
    k.QT2[i] = 2.0 * w.QT[i];
  }
  for (int i = 0; i < 4; ++i) {
    k.R[i] = w.R[i];
    k.R2[i] = 2.0 * w.R[i];
  }
  k.Qb = w.Q_b;
  k.Rb = w.R_b;
  k.QTb = w.QT_b;
  return k;
}
static bool per_problem_weights(const AcroWeights& w) { return w.Q_b || w.R_b || w.QT_b; }

struct Cfg {
  int block;
  unsigned grid;
};
// Small batches are latency bound: one warp per block spreads the warps over all SMs.
// Large batches are FP64-throughput bound: wider blocks, several resident per SM.
static Cfg cfg_for(int64_t n) {
  Cfg c;
  c.block = (n >= 148LL * 64 * 8) ? 128 : (n >= 148LL * 32 * 4 ? 64 : 32);
  c.grid = (unsigned)((n + c.block - 1) / c.block);
  return c;
}

#define DISPATCH2(b0, b1, EXPR)                   \
  do {                                            \
    if (b0) {                                     \
      if (b1) { EXPR(true, true); } else { EXPR(true, false); }   \
    } else {                                      \
      if (b1) { EXPR(false, true); } else { EXPR(false, false); } \
    }                                             \
  } while (0)

// ---------------------------------------------------------------------------------------
// D1-D3 point-wise kernels
// ---------------------------------------------------------------------------------------
enum { PT_F = 0, PT_RK4 = 1 };

// PPB: physical parameters per problem (`pb`, see model_per_problem); otherwise the shared model in the constant bank
#define ACRO_MODEL(PPB, m0, pb, B, b) \
  Model m_loc_;                        \
  if (PPB) m_loc_ = model_per_problem(m0, pb, B, b); \
  const Model& m = PPB ? m_loc_ : m0

template <int MODE, bool PPB>
__global__ void k_point(const __grid_constant__ Model m0, const double* __restrict__ pb, int64_t B,
                        const double* __restrict__ x, const double* __restrict__ u, double* __restrict__ out) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  ACRO_MODEL(PPB, m0, pb, B, b);
  double xs[4], o[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) xs[c] = x[c * B + b];
  const double u0 = u[b], u1 = u[B + b];
  if (MODE == PT_F)
    f_eval(m, xs, u0, u1, o);
  else
    rk4_step(m, xs, u0, u1, o);
#pragma unroll
  for (int c = 0; c < 4; ++c) out[c * B + b] = o[c];
}

template <bool PPB>
__global__ void k_linearize(const __grid_constant__ Model m0, const double* __restrict__ pb, int64_t B,
                            const double* __restrict__ x, const double* __restrict__ u, double* __restrict__ A,
                            double* __restrict__ Bm, int discrete) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  ACRO_MODEL(PPB, m0, pb, B, b);
  double xs[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) xs[c] = x[c * B + b];
  const LinC c = linearize_c(m, xs, u[b], u[B + b]);
  double a[2][4], b0[2], b1[2];
  if (discrete) {
    const LinD d = discretize(c, m.dt);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) a[r][j] = d.a[r][j];
      b0[r] = d.b0[r];
      b1[r] = d.b[r];
    }
  } else {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
#pragma unroll
      for (int j = 0; j < 4; ++j) a[r][j] = c.ac[r][j];
      b0[r] = c.bc0[r];
      b1[r] = c.bc1[r];
    }
  }
  const double one = discrete ? 1.0 : 0.0, off = discrete ? m.dt : 1.0;
  // rows 0-1: continuous [0 0 1 0; 0 0 0 1], discrete [1 0 dt 0; 0 1 0 dt]
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    A[(0 * 4 + j) * B + b] = (j == 0) ? one : (j == 2 ? off : 0.0);
    A[(1 * 4 + j) * B + b] = (j == 1) ? one : (j == 3 ? off : 0.0);
    A[(2 * 4 + j) * B + b] = a[0][j];
    A[(3 * 4 + j) * B + b] = a[1][j];
  }
  Bm[0 * B + b] = 0.0;
  Bm[1 * B + b] = 0.0;
  Bm[2 * B + b] = 0.0;
  Bm[3 * B + b] = 0.0;
  Bm[4 * B + b] = b0[0];
  Bm[5 * B + b] = b1[0];
  Bm[6 * B + b] = b0[1];
  Bm[7 * B + b] = b1[1];
}

// ---------------------------------------------------------------------------------------
// compute_equilibrium (tg:22-39): G(theta1, theta2) = u_target, one Newton iteration per thread.
//   g1 sin(th1) + g2 sin(th1 + th2) = u[0],   g2 sin(th1 + th2) = u[1]
// (the reference hands the same two equations to MINPACK's hybr; the root is the same to rounding)
// ---------------------------------------------------------------------------------------
template <bool PPB>
__global__ void k_equilibrium(const __grid_constant__ Model m0, const double* __restrict__ pb, int64_t B,
                              const double* __restrict__ ut, const double* __restrict__ th0, double tol, int max_iter,
                              double* __restrict__ theta, int32_t* __restrict__ n_iter) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  ACRO_MODEL(PPB, m0, pb, B, b);
  double t1 = th0[b], t2 = th0[B + b];
  const double u0 = ut[b], u1 = ut[B + b];
  int n = -max_iter;
  for (int i = 0; i <= max_iter; ++i) {
    double s1, c1, s12, c12;
    sincos(t1, &s1, &c1);
    sincos(t1 + t2, &s12, &c12);
    const double r0 = fma(m.g1, s1, m.g2 * s12) - u0, r1 = m.g2 * s12 - u1;
    if (fmax(fabs(r0), fabs(r1)) < tol) {
      n = i;
      break;
    }
    if (i == max_iter) break;
    const double j11 = m.g2 * c12, j00 = fma(m.g1, c1, j11);
    const Lu2 lu = lu2(j00, j11, j11, j11);
    double d0, d1;
    lu2_solve(lu, r0, r1, d0, d1);
    t1 -= d0;
    t2 -= d1;
  }
  theta[b] = t1;
  theta[B + b] = t2;
  n_iter[b] = n;
}

// ---------------------------------------------------------------------------------------
// G1 open-loop rollout, G8 cost, G3 costate
// ---------------------------------------------------------------------------------------
template <bool PPB>
__global__ void k_rollout_open(const __grid_constant__ Model m0, const double* __restrict__ pb, int64_t B, int N,
                               const double* __restrict__ x0, const double* __restrict__ U, double* __restrict__ X) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  ACRO_MODEL(PPB, m0, pb, B, b);
  double x[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    x[c] = x0[c * B + b];
    X[soa(0, 4, c, N, b)] = x[c];
  }
  double u0 = U ? U[soa(0, 2, 0, N - 1, b)] : 0.0, u1 = U ? U[soa(0, 2, 1, N - 1, b)] : 0.0;
  TrigCarry tc = trig_carry_at(m, x[0], x[1]);
  for (int t = 0; t < N - 1; ++t) {
    double n0 = 0.0, n1 = 0.0;
    if (U && t + 1 < N - 1) {  // prefetch the next input while this step computes
      n0 = U[soa(t + 1, 2, 0, N - 1, b)];
      n1 = U[soa(t + 1, 2, 1, N - 1, b)];
    }
    double xn[4];
    rk4_step_inc(m, x, u0, u1, xn, tc);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      x[c] = xn[c];
      X[soa(t + 1, 4, c, N, b)] = x[c];
    }
    u0 = n0;
    u1 = n1;
  }
}

// total_cost (tg:231-252): stage costs with Q, R accumulated in time order, terminal with Q_T.
template <bool WPB, bool RPB>
__device__ __forceinline__ double total_cost_dev(const WV<WPB>& w, const RefV<RPB>& ref, int N, const double* X,
                                                 const double* U, int64_t ld, int64_t b) {
  double cost = 0.0;
  for (int t = 0; t < N - 1; ++t) {
    double dx[4], du[2];
#pragma unroll
    for (int c = 0; c < 4; ++c) dx[c] = X[soa(t, 4, c, N, b)] - ref.X(t, c);
#pragma unroll
    for (int c = 0; c < 2; ++c) du[c] = U[soa(t, 2, c, N - 1, b)] - ref.U(t, c);
    cost += quad4(dx, [&](int i, int j) { return w.Q(i, j); });
    cost += quad2(du, [&](int i, int j) { return w.R(i, j); });
  }
  double dx[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) dx[c] = X[soa(N - 1, 4, c, N, b)] - ref.X(N - 1, c);
  cost += quad4(dx, [&](int i, int j) { return w.QT(i, j); });
  return cost;
}

template <bool WPB, bool RPB>
__global__ void k_total_cost(const __grid_constant__ KWeights kw, int64_t B, int N, const double* __restrict__ X,
                             const double* __restrict__ U, const double* rx, const double* ru,
                             double* __restrict__ cost) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  const WV<WPB> w(kw, B, b);
  const RefV<RPB> ref{rx, ru, N, b};
  cost[b] = total_cost_dev(w, ref, N, X, U, B, b);
}

// compute_costate_trajectory (tg:138-159)
template <bool WPB, bool RPB>
__global__ void k_costate(const __grid_constant__ Model m, const __grid_constant__ KWeights kw, int64_t B, int N,
                          const double* __restrict__ X, const double* __restrict__ U, const double* rx,
                          const double* ru, double* __restrict__ lam) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  const WV<WPB> w(kw, B, b);
  const RefV<RPB> ref{rx, ru, N, b};
  double l[4], dx[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) dx[c] = X[soa(N - 1, 4, c, N, b)] - ref.X(N - 1, c);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double s = w.QT2(i, 0) * dx[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) s = fma(w.QT2(i, j), dx[j], s);
    l[i] = s;
    lam[soa(N - 1, 4, i, N, b)] = s;
  }
  for (int t = N - 2; t >= 0; --t) {
    double x[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = X[soa(t, 4, c, N, b)];
    const LinD L = linearize_d(m, x, U[soa(t, 2, 0, N - 1, b)], U[soa(t, 2, 1, N - 1, b)]);
#pragma unroll
    for (int c = 0; c < 4; ++c) dx[c] = x[c] - ref.X(t, c);
    double g[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = w.Q2(i, 0) * dx[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) s = fma(w.Q2(i, j), dx[j], s);
      g[i] = s;
    }
    const double l0 = l[0], l1 = l[1], l2 = l[2], l3 = l[3];
    l[0] = g[0] + fma(L.a[1][0], l3, fma(L.a[0][0], l2, l0));
    l[1] = g[1] + fma(L.a[1][1], l3, fma(L.a[0][1], l2, l1));
    l[2] = g[2] + fma(L.a[1][2], l3, fma(L.a[0][2], l2, m.dt * l0));
    l[3] = g[3] + fma(L.a[1][3], l3, fma(L.a[0][3], l2, m.dt * l1));
#pragma unroll
    for (int i = 0; i < 4; ++i) lam[soa(t, 4, i, N, b)] = l[i];
  }
}

// Pull the rows of step t of an SoA array towards L2 ahead of the register prefetch: every lane names its
// own element, so one instruction covers the 256 B row of the warp; no register, no scoreboard.
#define ACRO_PF_DIST 4
template <int C>
__device__ __forceinline__ void l2_prefetch_rows(const double* __restrict__ A, int t, int64_t T, int64_t b) {
  // the C rows of the warp's step are one block of 2*C lines of 128 B: lane l names line l
  const int lane = int(b & 31);
  if (lane < 2 * C) asm volatile("prefetch.global.L2 [%0];" ::"l"(A + soa(t, C, 0, T, b & ~int64_t(31)) + lane * 16));
}

// compact discrete linearisation lin[t][10][ld]: a[0][0..3], a[1][0..3], b[0], b[1]
// T = number of time steps of the tiled array; T == 0 means a single shared linearisation stored plainly as lin[t][10]
__device__ __forceinline__ LinD load_lin(const double* __restrict__ lin, int t, int64_t T, int64_t b) {
  LinD L;
  if (T == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      L.a[0][j] = __ldg(lin + t * 10 + j);
      L.a[1][j] = __ldg(lin + t * 10 + 4 + j);
    }
    L.b[0] = __ldg(lin + t * 10 + 8);
    L.b[1] = __ldg(lin + t * 10 + 9);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      L.a[0][j] = lin[soa(t, 10, j, T, b)];
      L.a[1][j] = lin[soa(t, 10, 4 + j, T, b)];
    }
    L.b[0] = lin[soa(t, 10, 8, T, b)];
    L.b[1] = lin[soa(t, 10, 9, T, b)];
  }
  L.b0[0] = L.b0[1] = 0.0;
  return L;
}
__device__ __forceinline__ void store_lin(double* __restrict__ lin, int t, int64_t T, int64_t b, const LinD& L) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    lin[soa(t, 10, j, T, b)] = L.a[0][j];
    lin[soa(t, 10, 4 + j, T, b)] = L.a[1][j];
  }
  lin[soa(t, 10, 8, T, b)] = L.b[0];
  lin[soa(t, 10, 9, T, b)] = L.b[1];
}

// ---------------------------------------------------------------------------------------
// G2+G4+G5+G6 fused backward pass: linearise, discretise, cost blocks, affine Riccati.
// Reads X, U of one problem (column b of leading dimension ld), writes K, S.
// ---------------------------------------------------------------------------------------
// HAVE_LIN: the discrete linearisation about (X, U) was already written by the forward pass that produced
// the iterate (rk4_step_lin), so this pass is pure Riccati algebra on loaded operands.
// ACT0: the fully-actuated plant (tau_1 = u[0] acts on the first joint, fully_actuated_ref_gen.py:20-73): B_d has two
// columns, G = R + B'PB is a full 2x2; the compact linearisation streamed between passes holds one column only, so
// this variant always linearises itself (HAVE_LIN = false).
template <bool WPB, bool RPB, bool HAVE_LIN, bool ACT0 = false>
__device__ __forceinline__ void backward_pass(const Model& m, const WV<WPB>& w, const RefV<RPB>& ref, int N,
                                              const double* __restrict__ X, const double* __restrict__ U,
                                              const double* __restrict__ lin, double* __restrict__ K,
                                              double* __restrict__ S, int64_t ld, int64_t b, double& dJ_out,
                                              double& sn_out) {
  double P[10], p[4];
  {
    double dx[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) dx[c] = X[soa(N - 1, 4, c, N, b)] - ref.X(N - 1, c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = w.QT2(i, 0) * dx[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) s = fma(w.QT2(i, j), dx[j], s);
      p[i] = s;
#pragma unroll
      for (int j = i; j < 4; ++j) P[sym(i, j)] = w.QT2(i, j);
    }
  }
  double dJ = 0.0, sn = 0.0;
  // operands of step t (state, input, reference) are loaded one iteration ahead, so no load result is
  // consumed in the iteration that issued it (all loads of an iteration share scoreboard slots)
  double x[4], u[2], xr[4], ur[2];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    x[c] = X[soa(N - 2, 4, c, N, b)];
    xr[c] = ref.X(N - 2, c);
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    u[c] = U[soa(N - 2, 2, c, N - 1, b)];
    ur[c] = ref.U(N - 2, c);
  }
  const QhQ2<WV<WPB>> Qh{w};
  const Lu2Col col = lu2_col(w.R2(0, 0), w.R2(0, 1));
  LinD Lc;
  if (HAVE_LIN) Lc = load_lin(lin, N - 2, N - 1, b);
  for (int t = N - 2; t >= 0; --t) {
    if (t > ACRO_PF_DIST) {
      l2_prefetch_rows<4>(X, t - 1 - ACRO_PF_DIST, N, b);
      l2_prefetch_rows<2>(U, t - 1 - ACRO_PF_DIST, N - 1, b);
      if (HAVE_LIN) l2_prefetch_rows<10>(lin, t - 1 - ACRO_PF_DIST, N - 1, b);
    }
    const LinD L = HAVE_LIN ? Lc : linearize_d(m, x, u[0], u[1]);
    double dx[4], du[2], q[4], r[2];
#pragma unroll
    for (int c = 0; c < 4; ++c) dx[c] = x[c] - xr[c];
#pragma unroll
    for (int c = 0; c < 2; ++c) du[c] = u[c] - ur[c];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = w.Q2(i, 0) * dx[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) s = fma(w.Q2(i, j), dx[j], s);
      q[i] = s;
    }
    r[0] = fma(w.R2(0, 1), du[1], w.R2(0, 0) * du[0]);
    r[1] = fma(w.R2(1, 1), du[1], w.R2(1, 0) * du[0]);
    // operands of step t-1 into the registers step t has finished with (x, u, reference); the Riccati algebra
    // below covers the load latency
    if (t > 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        x[c] = X[soa(t - 1, 4, c, N, b)];
        xr[c] = ref.X(t - 1, c);
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        u[c] = U[soa(t - 1, 2, c, N - 1, b)];
        ur[c] = ref.U(t - 1, c);
      }
    }
    double Kt[8], st[2];
    riccati_step<true, ACT0>(P, p, L, m.dt, Qh, col, w.R2(0, 0), w.R2(0, 1), w.R2(1, 1), q, r, Kt, st, dJ);
    if (HAVE_LIN && t > 0) Lc = load_lin(lin, t - 1, N - 1, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) K[soa(t, 8, e, N - 1, b)] = Kt[e];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      S[soa(t, 2, e, N - 1, b)] = st[e];
      const double a = fabs(st[e]);
      sn = (a > sn || a != a) ? a : sn;  // NaN is sticky, like np.max(np.abs(sigma))
    }
  }
  dJ_out = dJ;
  sn_out = sn;
}

template <bool WPB, bool RPB, bool ACT0 = false>
__global__ void k_riccati_affine(const __grid_constant__ Model m, const __grid_constant__ KWeights kw, int64_t B,
                                 int N, const double* __restrict__ X, const double* __restrict__ U,
                                 const double* rx, const double* ru, double* __restrict__ K,
                                 double* __restrict__ S, double* __restrict__ dJ, double* __restrict__ sn) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  const WV<WPB> w(kw, B, b);
  const RefV<RPB> ref{rx, ru, N, b};
  double d, s;
  backward_pass<WPB, RPB, false, ACT0>(m, w, ref, N, X, U, nullptr, K, S, B, b, d, s);
  dJ[b] = d;
  sn[b] = s;
}

// ---------------------------------------------------------------------------------------
// G7+G8 fused forward pass: closed-loop rollout with step size gamma and its total cost.
// Base iterate (X,U,K,S) is column b of leading dimension ld; the candidate trajectory goes
// to column bo of leading dimension ldo of (Xn,Un) when STORE.
// ---------------------------------------------------------------------------------------
struct StepIn {
  double x[4], u[2], k[8], s[2], xr[4], ur[2];
};
template <class REF>
__device__ __forceinline__ StepIn load_step(const double* __restrict__ X, const double* __restrict__ U,
                                            const double* __restrict__ K, const double* __restrict__ S,
                                            const REF& ref, int t, int64_t N, int64_t b) {
  StepIn in;
#pragma unroll
  for (int c = 0; c < 4; ++c) in.xr[c] = ref.X(t, c);
#pragma unroll
  for (int c = 0; c < 2; ++c) in.ur[c] = ref.U(t, c);
#pragma unroll
  for (int c = 0; c < 4; ++c) in.x[c] = X[soa(t, 4, c, N, b)];
#pragma unroll
  for (int c = 0; c < 2; ++c) in.u[c] = U[soa(t, 2, c, N - 1, b)];
#pragma unroll
  for (int c = 0; c < 8; ++c) in.k[c] = K[soa(t, 8, c, N - 1, b)];
#pragma unroll
  for (int c = 0; c < 2; ++c) in.s[c] = S[soa(t, 2, c, N - 1, b)];
  return in;
}

template <bool WPB, bool RPB, bool STORE, bool LIN = false>
__device__ __forceinline__ double forward_pass(const Model& m, const WV<WPB>& w, const RefV<RPB>& ref, int N,
                                               const double* __restrict__ X, const double* __restrict__ U,
                                               const double* __restrict__ K, const double* __restrict__ S,
                                               int64_t ld, int64_t b, double gamma, double* __restrict__ Xn,
                                               double* __restrict__ Un, int64_t ldo, int64_t bo,
                                               double* __restrict__ lin_out = nullptr) {
  double xp[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) xp[c] = X[soa(0, 4, c, N, b)];
  double cost = 0.0;
  StepIn in = load_step(X, U, K, S, ref, 0, N, b);
  TrigCarry tc;
  if (!LIN) tc = trig_carry_at(m, xp[0], xp[1]);
  double xrT[4];  // terminal reference, loaded early
#pragma unroll
  for (int c = 0; c < 4; ++c) xrT[c] = ref.X(N - 1, c);
  for (int t = 0; t < N - 1; ++t) {
    if (t + 1 + ACRO_PF_DIST < N - 1) {
      l2_prefetch_rows<4>(X, t + 1 + ACRO_PF_DIST, N, b);
      l2_prefetch_rows<2>(U, t + 1 + ACRO_PF_DIST, N - 1, b);
      l2_prefetch_rows<8>(K, t + 1 + ACRO_PF_DIST, N - 1, b);
      l2_prefetch_rows<2>(S, t + 1 + ACRO_PF_DIST, N - 1, b);
    }
    double dx[4], up[2];
#pragma unroll
    for (int c = 0; c < 4; ++c) dx[c] = xp[c] - in.x[c];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double kd = in.k[i * 4] * dx[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) kd = fma(in.k[i * 4 + j], dx[j], kd);
      up[i] = (in.u[i] + kd) + gamma * in.s[i];
    }
    if (STORE) {
#pragma unroll
      for (int c = 0; c < 4; ++c) Xn[soa(t, 4, c, N, bo)] = xp[c];
#pragma unroll
      for (int c = 0; c < 2; ++c) Un[soa(t, 2, c, N - 1, bo)] = up[c];
    }
    double ex[4], eu[2];
#pragma unroll
    for (int c = 0; c < 4; ++c) ex[c] = xp[c] - in.xr[c];
#pragma unroll
    for (int c = 0; c < 2; ++c) eu[c] = up[c] - in.ur[c];
    cost += quad4(ex, [&](int i, int j) { return w.Q(i, j); });
    cost += quad2(eu, [&](int i, int j) { return w.R(i, j); });
    // operands of step t+1 straight into the registers step t no longer needs; the four RK4 stages below
    // (about a thousand cycles) cover the load latency
    if (t + 1 < N - 1) in = load_step(X, U, K, S, ref, t + 1, N, b);
    double xn[4];
    if (LIN) {
      LinD L;
      rk4_step_lin(m, xp, up[0], up[1], xn, L);
      store_lin(lin_out, t, N - 1, bo, L);
    } else {
      rk4_step_inc(m, xp, up[0], up[1], xn, tc);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) xp[c] = xn[c];
  }
  double ex[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (STORE) Xn[soa(N - 1, 4, c, N, bo)] = xp[c];
    ex[c] = xp[c] - xrT[c];
  }
  cost += quad4(ex, [&](int i, int j) { return w.QT(i, j); });
  return cost;
}

// one thread per (problem b, candidate g); lanes run over b
template <bool WPB, bool RPB>
__global__ void __launch_bounds__(128, ACRO_ROLLOUT_MINB) k_closed_loop(const __grid_constant__ Model m, const __grid_constant__ KWeights kw, int64_t B,
                              int N, const double* __restrict__ X, const double* __restrict__ U,
                              const double* __restrict__ K, const double* __restrict__ S, const double* rx,
                              const double* ru, int G, const double* __restrict__ gammas, int gpp,
                              double* __restrict__ Xn, double* __restrict__ Un, double* __restrict__ cost) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  const int g = blockIdx.y;
  if (b >= B) return;
  const WV<WPB> w(kw, B, b);
  const RefV<RPB> ref{rx, ru, N, b};
  const double gamma = gpp ? gammas[int64_t(g) * B + b] : gammas[g];
  double c;
  if (Xn)
    c = forward_pass<WPB, RPB, true>(m, w, ref, N, X, U, K, S, B, b, gamma, Xn + int64_t(g) * N * 4 * padded(B),
                                     Un + int64_t(g) * (N - 1) * 2 * padded(B), B, b);
  else
    c = forward_pass<WPB, RPB, false>(m, w, ref, N, X, U, K, S, B, b, gamma, nullptr, nullptr, B, b);
  cost[int64_t(g) * B + b] = c;
}

// G11: P base iterates x S_n step sizes.  A warp = 32 consecutive step sizes of ONE base
// iterate, so the 16 operand loads per step are warp-wide broadcasts; the 4 warps of a
// block take 4 neighbouring iterates, i.e. whole 32-byte sectors of every SoA row.
template <bool WPB, bool RPB>
__global__ void __launch_bounds__(128, ACRO_ROLLOUT_MINB) k_sweep(const __grid_constant__ Model m, const __grid_constant__ KWeights kw, int64_t P, int N,
                        const double* __restrict__ X, const double* __restrict__ U, const double* __restrict__ K,
                        const double* __restrict__ S, const double* rx, const double* ru, int S_n,
                        const double* __restrict__ steps, double* __restrict__ cost) {
  const int64_t p = blockIdx.x * 4LL + (threadIdx.x >> 5);
  const int s = blockIdx.y * 32 + (threadIdx.x & 31);
  if (p >= P || s >= S_n) return;
  const WV<WPB> w(kw, P, p);
  const RefV<RPB> ref{rx, ru, N, p};
  cost[int64_t(s) * P + p] =
      forward_pass<WPB, RPB, false>(m, w, ref, N, X, U, K, S, P, p, steps[s], nullptr, nullptr, P, p);
}

__global__ void k_armijo_select(int64_t B, int G, const double* __restrict__ cost_k, const double* __restrict__ dJ,
                                const double* __restrict__ gammas, int gpp, const double* __restrict__ cc, double c,
                                int32_t* __restrict__ acc) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  int a = -1;
  const double ck = cost_k[b], d = dJ[b];
  for (int g = 0; g < G && a < 0; ++g) {
    const double gam = gpp ? gammas[int64_t(g) * B + b] : gammas[g];
    // cost_k + c*gamma*dJ evaluated as numpy does: ((c*gamma)*dJ) then +, no FMA contraction
    const double thr = __dadd_rn(ck, __dmul_rn(__dmul_rn(c, gam), d));
    if (cc[int64_t(g) * B + b] < thr) a = g;
  }
  acc[b] = a;
}

// ---------------------------------------------------------------------------------------
// G10: the whole Newton / Armijo loop of one problem in one thread (tg:298-398).
// ---------------------------------------------------------------------------------------
struct NewtonArgs {
  Model m;
  KWeights kw;
  AcroNewtonOpts o;
  int64_t B;
  int N;
  const double* x0;
  const double* pb;  // per-problem physical parameters [11][B] or nullptr (k_newton<.., PPB = true> only)
  const double *rx, *ru;
  double *X, *U, *Xw, *Uw, *lin, *K, *S;
  double* spec_ws;  // candidate trajectories of k_newton_spec (AcroNewtonOpts.spec_ws) or nullptr
  double *cost, *dJ, *sn, *gacc;
  int32_t *iters, *status;
  double *h_cost, *h_sn, *h_gamma;
  int32_t* h_ntry;
};

template <bool WPB, bool RPB, bool PPB = false, bool ACT0 = false>
__global__ void k_newton(const __grid_constant__ NewtonArgs a) {
  const int64_t B = a.B, b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  ACRO_MODEL(PPB, a.m, a.pb, B, b);
  const int N = a.N;
  const WV<WPB> w(a.kw, B, b);
  const RefV<RPB> ref{a.rx, a.ru, N, b};
  int it, st;
  double cost_k;
  if (a.o.init) {
    // u = 0 (init = 1; init = 2 keeps the caller's U as a warm start), x = simulate_open_loop(x0, u),
    // cost_k = total_cost(...)   (tg:311-319); the rollout also leaves the linearisation for iteration 0
    double x[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      x[c] = a.x0[c * B + b];
      a.X[soa(0, 4, c, N, b)] = x[c];
    }
    for (int t = 0; t < N - 1; ++t) {
      double u0 = 0.0, u1 = 0.0;
      if (a.o.init == 2) {
        u0 = a.U[soa(t, 2, 0, N - 1, b)];
        u1 = a.U[soa(t, 2, 1, N - 1, b)];
      } else {
        a.U[soa(t, 2, 0, N - 1, b)] = 0.0;
        a.U[soa(t, 2, 1, N - 1, b)] = 0.0;
      }
      double xn[4];
      if (ACT0) {
        rk4_step(m, x, u0, u1, xn);
      } else {
        LinD L;
        rk4_step_lin(m, x, u0, u1, xn, L);
        store_lin(a.lin, t, N - 1, b, L);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        x[c] = xn[c];
        a.X[soa(t + 1, 4, c, N, b)] = x[c];
      }
    }
    cost_k = total_cost_dev(w, ref, N, a.X, a.U, B, b);
    it = 0;
    st = ACRO_RUNNING;
    if (a.h_cost) a.h_cost[b] = cost_k;
  } else {
    it = a.iters[b];
    st = a.status[b];
    cost_k = a.cost[b];
  }
  int cur = 0, done = 0;
  double dJ = a.o.init ? 0.0 : a.dJ[b], sn = a.o.init ? 0.0 : a.sn[b], gacc = a.o.init ? 0.0 : a.gacc[b];
  while (st == ACRO_RUNNING && it < a.o.max_iters && (a.o.chunk_iters <= 0 || done < a.o.chunk_iters)) {
    const double* Xc = cur ? a.Xw : a.X;
    const double* Uc = cur ? a.Uw : a.U;
    double* Xo = cur ? a.X : a.Xw;
    double* Uo = cur ? a.U : a.Uw;
    backward_pass<WPB, RPB, !ACT0, ACT0>(m, w, ref, N, Xc, Uc, a.lin, a.K, a.S, B, b, dJ, sn);
    if (a.h_sn) a.h_sn[int64_t(it) * B + b] = sn;
    double gamma = a.o.gamma_0, cn = 0.0;
    int tries = 0;
    bool ok = false;
    for (int i = 0; i < a.o.max_line_search; ++i) {
      cn = forward_pass<WPB, RPB, true, !ACT0>(m, w, ref, N, Xc, Uc, a.K, a.S, B, b, gamma, Xo, Uo, B, b, a.lin);
      ++tries;
      // accept iff cost_new < cost_k + c*gamma*delta_J  (strict, NaN rejects)   tg:361
      const double thr = __dadd_rn(cost_k, __dmul_rn(__dmul_rn(a.o.c, gamma), dJ));
      if (cn < thr) {
        ok = true;
        break;
      }
      gamma = __dmul_rn(gamma, a.o.beta);  // tg:365
    }
    if (a.h_ntry) a.h_ntry[int64_t(it) * B + b] = tries;
    ++it;
    ++done;
    if (!ok) {  // tg:367-369: keep the current iterate, stop
      st = ACRO_LINE_SEARCH_FAILED;
      if (a.h_gamma) a.h_gamma[int64_t(it - 1) * B + b] = nan("");
      break;
    }
    cur ^= 1;
    cost_k = cn;
    gacc = gamma;
    if (a.h_gamma) a.h_gamma[int64_t(it - 1) * B + b] = gamma;
    if (a.h_cost) a.h_cost[int64_t(it) * B + b] = cost_k;
    if (sn < a.o.tol) st = ACRO_CONVERGED;  // tg:394-396
  }
  if (st == ACRO_RUNNING && it >= a.o.max_iters) st = ACRO_MAX_ITERS;
  if (cur) {  // current iterate sits in the workspace: move it home
    for (int t = 0; t < N; ++t) {
#pragma unroll
      for (int c = 0; c < 4; ++c) a.X[soa(t, 4, c, N, b)] = a.Xw[soa(t, 4, c, N, b)];
      if (t < N - 1) {
#pragma unroll
        for (int c = 0; c < 2; ++c) a.U[soa(t, 2, c, N - 1, b)] = a.Uw[soa(t, 2, c, N - 1, b)];
      }
    }
  }
  a.cost[b] = cost_k;
  a.dJ[b] = dJ;
  a.sn[b] = sn;
  a.gacc[b] = gacc;
  a.iters[b] = it;
  a.status[b] = st;
}

}  // namespace acro
#include "acro_newton_ring.cuh"
#include "acro_newton_duo.cuh"
#include "acro_newton_spec.cuh"
namespace acro {

// ---------------------------------------------------------------------------------------
// Stand-alone pieces of the Newton iteration for the drop-in functions that expose them:
// derivatives_Cost (tg:89-114), discretize_linearization (tg:161-164), build_stage_lists
// (tg:166-181), calculate_K_and_sigma on caller-supplied lists (tg:183-216).
// ---------------------------------------------------------------------------------------
template <bool WPB>
__global__ void k_cost_derivatives(const __grid_constant__ KWeights kw, int64_t B, const double* __restrict__ x,
                                   const double* __restrict__ xr, const double* __restrict__ u,
                                   const double* __restrict__ ur, int terminal, double* __restrict__ l,
                                   double* __restrict__ gx, double* __restrict__ gu) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  const WV<WPB> w(kw, B, b);
  double dx[4], du[2] = {0, 0};
#pragma unroll
  for (int c = 0; c < 4; ++c) dx[c] = x[c * B + b] - xr[c * B + b];
  if (terminal) {
    l[b] = quad4(dx, [&](int i, int j) { return w.QT(i, j); });
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = w.QT2(i, 0) * dx[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) s = fma(w.QT2(i, j), dx[j], s);
      gx[i * B + b] = s;
    }
    return;
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) du[c] = u[c * B + b] - ur[c * B + b];
  l[b] = quad4(dx, [&](int i, int j) { return w.Q(i, j); }) + quad2(du, [&](int i, int j) { return w.R(i, j); });
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double s = w.Q2(i, 0) * dx[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) s = fma(w.Q2(i, j), dx[j], s);
    gx[i * B + b] = s;
  }
  gu[0 * B + b] = fma(w.R2(0, 1), du[1], w.R2(0, 0) * du[0]);
  gu[1 * B + b] = fma(w.R2(1, 1), du[1], w.R2(1, 0) * du[0]);
}

__global__ void k_discretize(int64_t B, const double* __restrict__ Ac, const double* __restrict__ Bc, double dt,
                             double* __restrict__ Ad, double* __restrict__ Bd) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Ad[(i * 4 + j) * B + b] = (i == j ? 1.0 : 0.0) + dt * Ac[(i * 4 + j) * B + b];
#pragma unroll
  for (int e = 0; e < 8; ++e) Bd[e * B + b] = dt * Bc[e * B + b];
}

// one thread per (t, b): dense A_d, B_d, q_t, r_t; threads with t == N-1 write q_T
template <bool WPB, bool RPB>
__global__ void k_stage_lists(const __grid_constant__ Model m, const __grid_constant__ KWeights kw, int64_t B, int N,
                              const double* __restrict__ X, const double* __restrict__ U, const double* rx,
                              const double* ru, double* __restrict__ A, double* __restrict__ Bm,
                              double* __restrict__ q, double* __restrict__ r, double* __restrict__ qT) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (idx >= int64_t(N) * B) return;
  const int t = int(idx / B);
  const int64_t b = idx % B;
  const WV<WPB> w(kw, B, b);
  const RefV<RPB> ref{rx, ru, N, b};
  double x[4], dx[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    x[c] = X[soa(t, 4, c, N, b)];
    dx[c] = x[c] - ref.X(t, c);
  }
  if (t == N - 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = w.QT2(i, 0) * dx[0];
#pragma unroll
      for (int j = 1; j < 4; ++j) s = fma(w.QT2(i, j), dx[j], s);
      qT[i * B + b] = s;
    }
    return;
  }
  const double u0 = U[soa(t, 2, 0, N - 1, b)], u1 = U[soa(t, 2, 1, N - 1, b)];
  const LinD L = linearize_d(m, x, u0, u1);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    A[soa(t, 16, 0 * 4 + j, N - 1, b)] = (j == 0) ? 1.0 : (j == 2 ? m.dt : 0.0);
    A[soa(t, 16, 1 * 4 + j, N - 1, b)] = (j == 1) ? 1.0 : (j == 3 ? m.dt : 0.0);
    A[soa(t, 16, 2 * 4 + j, N - 1, b)] = L.a[0][j];
    A[soa(t, 16, 3 * 4 + j, N - 1, b)] = L.a[1][j];
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) Bm[soa(t, 8, e, N - 1, b)] = 0.0;
  Bm[soa(t, 8, 4, N - 1, b)] = L.b0[0];
  Bm[soa(t, 8, 5, N - 1, b)] = L.b[0];
  Bm[soa(t, 8, 6, N - 1, b)] = L.b0[1];
  Bm[soa(t, 8, 7, N - 1, b)] = L.b[1];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    double s = w.Q2(i, 0) * dx[0];
#pragma unroll
    for (int j = 1; j < 4; ++j) s = fma(w.Q2(i, j), dx[j], s);
    q[soa(t, 4, i, N - 1, b)] = s;
  }
  const double du0 = u0 - ref.U(t, 0), du1 = u1 - ref.U(t, 1);
  r[soa(t, 2, 0, N - 1, b)] = fma(w.R2(0, 1), du1, w.R2(0, 0) * du0);
  r[soa(t, 2, 1, N - 1, b)] = fma(w.R2(1, 1), du1, w.R2(1, 0) * du0);
}

__global__ void k_riccati_lists(int64_t B, int T, const double* __restrict__ A, const double* __restrict__ Bm,
                                const double* __restrict__ Q, const double* __restrict__ R,
                                const double* __restrict__ Sx, const double* __restrict__ q,
                                const double* __restrict__ r, const double* __restrict__ QT,
                                const double* __restrict__ qT, double* __restrict__ K, double* __restrict__ S,
                                double* __restrict__ dJ) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  double P[10], p[4], d = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    p[i] = qT[i * B + b];
#pragma unroll
    for (int j = i; j < 4; ++j) P[sym(i, j)] = QT[(i * 4 + j) * B + b];
  }
  for (int t = T - 1; t >= 0; --t) {
    double Am[16], Bv[8], Qm[16], Rm[4], Sm[8], qv[4], rv[2], Kt[8], st[2];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      Am[e] = A[soa(t, 16, e, T, b)];
      Qm[e] = Q[soa(t, 16, e, T, b)];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      Bv[e] = Bm[soa(t, 8, e, T, b)];
      Sm[e] = Sx ? Sx[soa(t, 8, e, T, b)] : 0.0;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      Rm[e] = R[soa(t, 4, e, T, b)];
      qv[e] = q[soa(t, 4, e, T, b)];
    }
    rv[0] = r[soa(t, 2, 0, T, b)];
    rv[1] = r[soa(t, 2, 1, T, b)];
    riccati_step_lists(P, p, Am, Bv, Qm, Rm, Sm, qv, rv, Kt, st, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) K[soa(t, 8, e, T, b)] = Kt[e];
    S[soa(t, 2, 0, T, b)] = st[0];
    S[soa(t, 2, 1, T, b)] = st[1];
  }
  dJ[b] = d;
}

// ---------------------------------------------------------------------------------------
// T1 LQR gains, T2 LQR tracking
// ---------------------------------------------------------------------------------------
template <bool WPB, bool RPB, bool ACT0 = false>
__global__ void k_lqr_gains(const __grid_constant__ Model m, const __grid_constant__ KWeights kw, int64_t B, int N,
                            const double* rx, const double* ru, double* __restrict__ K) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  const WV<WPB> w(kw, B, b);
  const RefV<RPB> ref{rx, ru, N, b};
  double P[10], p[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) P[sym(i, j)] = w.Q2(i, j);  // Q_T_reg = 2 Q_reg  (tt:175)
  const QhQ<WV<WPB>> Qh{w};
  const Lu2Col col = lu2_col(w.R(0, 0), w.R(0, 1));
  double dummy = 0.0;
  for (int t = N - 2; t >= 0; --t) {
    double x[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = ref.X(t, c);
    const LinD L = linearize_d(m, x, ref.U(t, 0), ref.U(t, 1));
    double Kt[8], st[2];
    riccati_step<false, ACT0>(P, p, L, m.dt, Qh, col, w.R(0, 0), w.R(0, 1), w.R(1, 1), p, p, Kt, st, dummy);
    if (RPB) {
#pragma unroll
      for (int e = 0; e < 8; ++e) K[soa(t, 8, e, N - 1, b)] = Kt[e];
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) K[t * 8 + e] = Kt[e];
    }
  }
}

template <bool RPB, bool PPB>
__global__ void k_lqr_track(const __grid_constant__ Model m0, const double* __restrict__ pb, int64_t B, int N,
                            const double* rx, const double* ru, const double* __restrict__ K,
                            const double* __restrict__ x0, double* __restrict__ Xt, double* __restrict__ Ut) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  ACRO_MODEL(PPB, m0, pb, B, b);
  const RefV<RPB> ref{rx, ru, N, b};
  double x[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    x[c] = x0[c * B + b];
    Xt[soa(0, 4, c, N, b)] = x[c];
  }
  TrigCarry tc = trig_carry_at(m, x[0], x[1]);
  for (int t = 0; t < N - 1; ++t) {
    double u[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double kd = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double k = RPB ? K[soa(t, 8, i * 4 + j, N - 1, b)] : __ldg(K + t * 8 + i * 4 + j);
        kd = (j == 0) ? k * (x[0] - ref.X(t, 0)) : fma(k, x[j] - ref.X(t, j), kd);
      }
      u[i] = ref.U(t, i) + kd;
      Ut[soa(t, 2, i, N - 1, b)] = u[i];
    }
    double xn[4];
    rk4_step_inc(m, x, u[0], u[1], xn, tc);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      x[c] = xn[c];
      Xt[soa(t + 1, 4, c, N, b)] = x[c];
    }
  }
}

// ---------------------------------------------------------------------------------------
// T3 P_inf, T4 MPC solve (dense caller-supplied matrices)
// ---------------------------------------------------------------------------------------
template <bool WPB>
__global__ void k_p_inf(const __grid_constant__ KWeights kw, int64_t B, const double* __restrict__ A,
                        const double* __restrict__ Bm, int max_iter, double tol, double* __restrict__ Pout,
                        int32_t* __restrict__ n_iter) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  const WV<WPB> w(kw, B, b);
  double Am[16], Bv[8], P[10];
#pragma unroll
  for (int e = 0; e < 16; ++e) Am[e] = A[e * B + b];
#pragma unroll
  for (int e = 0; e < 8; ++e) Bv[e] = Bm[e * B + b];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) P[sym(i, j)] = w.Q(i, j);
  const QhQ<WV<WPB>> Qh{w};
  int n = -max_iter;
  for (int i = 0; i < max_iter; ++i) {
    double Pp[10], K[8];
#pragma unroll
    for (int e = 0; e < 10; ++e) Pp[e] = P[e];
    riccati_step_dense(P, Am, Bv, Qh, w.R(0, 0), w.R(0, 1), w.R(1, 1), K);
    double d = 0.0;
#pragma unroll
    for (int e = 0; e < 10; ++e) {
      const double v = fabs(P[e] - Pp[e]);
      d = (v > d || v != v) ? v : d;
    }
    if (d < tol) {  // tt:161-162
      n = i + 1;
      break;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Pout[(i * 4 + j) * B + b] = P[sym(i, j)];
  n_iter[b] = n;
}

template <bool WPB>
__global__ void k_mpc_solve(const __grid_constant__ KWeights kw, int64_t B, int H, const double* __restrict__ x0,
                            const double* __restrict__ Aw, const double* __restrict__ Bw,
                            const double* __restrict__ QT, double* __restrict__ U0, double* __restrict__ Xo,
                            double* __restrict__ Uo, double* __restrict__ Kws) {
  const int64_t b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  const WV<WPB> w(kw, B, b);
  double P[10];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) P[sym(i, j)] = QT[(i * 4 + j) * B + b];
  const QhQ<WV<WPB>> Qh{w};
  double K[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = H - 2; j >= 0; --j) {
    double Am[16], Bv[8];
#pragma unroll
    for (int e = 0; e < 16; ++e) Am[e] = Aw[soa(j, 16, e, H - 1, b)];
#pragma unroll
    for (int e = 0; e < 8; ++e) Bv[e] = Bw[soa(j, 8, e, H - 1, b)];
    riccati_step_dense(P, Am, Bv, Qh, w.R(0, 0), w.R(0, 1), w.R(1, 1), K);
    if (Kws) {
#pragma unroll
      for (int e = 0; e < 8; ++e) Kws[soa(j, 8, e, H - 1, b)] = K[e];
    }
  }
  double x[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) x[c] = x0[c * B + b];
  // U0 = K_0 x0 (tt:135); K holds K_0 after the sweep (H >= 2)
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    double s = K[i * 4] * x[0];
#pragma unroll
    for (int jj = 1; jj < 4; ++jj) s = fma(K[i * 4 + jj], x[jj], s);
    U0[i * B + b] = (H >= 2) ? s : 0.0;
  }
  if (!Xo || !Uo || !Kws) return;
  // forward pass of the predicted trajectory (X_opt, U_opt of tt:136-137)
  for (int j = 0; j < H; ++j) {
#pragma unroll
    for (int c = 0; c < 4; ++c) Xo[soa(j, 4, c, H, b)] = x[c];
    if (j == H - 1) {
      Uo[soa(j, 2, 0, H, b)] = 0.0;  // free, unpenalised variable stays at its initial guess
      Uo[soa(j, 2, 1, H, b)] = 0.0;
      break;
    }
    double u[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double s = Kws[soa(j, 8, i * 4, H - 1, b)] * x[0];
#pragma unroll
      for (int jj = 1; jj < 4; ++jj) s = fma(Kws[soa(j, 8, i * 4 + jj, H - 1, b)], x[jj], s);
      u[i] = s;
      Uo[soa(j, 2, i, H, b)] = s;
    }
    double xn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double s = Aw[soa(j, 16, i * 4, H - 1, b)] * x[0];
#pragma unroll
      for (int jj = 1; jj < 4; ++jj) s = fma(Aw[soa(j, 16, i * 4 + jj, H - 1, b)], x[jj], s);
      s = fma(Bw[soa(j, 8, i * 2, H - 1, b)], u[0], s);
      s = fma(Bw[soa(j, 8, i * 2 + 1, H - 1, b)], u[1], s);
      xn[i] = s;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) x[c] = xn[c];
  }
}

// ---------------------------------------------------------------------------------------
// T5 MPC tracking
// ---------------------------------------------------------------------------------------
// compact discrete linearisation about a trajectory: lin[t][10][ld] = a[0][0..3], a[1][0..3], b[0], b[1]
template <bool RPB>
__global__ void k_lin_compact(const __grid_constant__ Model m, int64_t B, int N, const double* rx, const double* ru,
                              double* __restrict__ lin, const double* __restrict__ pb = nullptr) {
  const int64_t idx = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  // RPB: one thread per (t, b) with b fastest; shared: one thread per t
  const int64_t nb = RPB ? B : 1;
  if (idx >= int64_t(N - 1) * nb) return;
  const int t = int(idx / nb);
  const int64_t b = idx % nb;
  const RefV<RPB> ref{rx, ru, N, b};
  double x[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) x[c] = ref.X(t, c);
  const LinD L = (RPB && pb) ? linearize_d(model_per_problem(m, pb, B, b), x, ref.U(t, 0), ref.U(t, 1))
                             : linearize_d(m, x, ref.U(t, 0), ref.U(t, 1));
  if (RPB) {
    store_lin(lin, t, N - 1, b, L);
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      lin[t * 10 + j] = L.a[0][j];
      lin[t * 10 + 4 + j] = L.a[1][j];
    }
    lin[t * 10 + 8] = L.b[0];
    lin[t * 10 + 9] = L.b[1];
  }
}

struct MpcArgs {
  Model m;
  KWeights kw;
  int64_t B;
  int N, T, H;
  const double *rx, *ru;
  double xf[4], uf[2];
  const double* QT;  // [16] shared or [16][B]
  int qt_per_problem;
  const double* x0;
  const double* lin;  // compact linearisation: shared [N-1][10] or [N-1][10][B]
  double* K0;         // shared mode: [(T-1)][8]
  double *Xr, *Ur;
  const double* pb;   // physical parameters per problem [11][B] (k_mpc_track_pp<.., true>) or null
};

// One receding-horizon solve: (H-1)-step Riccati sweep over the window starting at time t
// (tt:50, 80-117 with the window/padding of tt:64-67) -> first-move gain K_0 (2x4).
template <bool WPB, int SW>
__device__ __forceinline__ void mpc_sweep_sw(const WV<WPB>& w, double dt, const double* __restrict__ lin, int64_t ld,
                                             int64_t b, int n_lin, const LinD& Lf, const double QT[10], int t, int H,
                                             double K[8]) {
  double P[10], inv_u11, qsel;
#pragma unroll
  for (int e = 0; e < 10; ++e) P[e] = QT[e];
  const QhQ<WV<WPB>> Qh{w};
  const Lu2Col col = lu2_col(w.R(0, 0), w.R(0, 1));
  int j = H - 2;
  // padded tail of the window: linearisation about (x_f, u_f).  Only the last step of a sweep (j = 0) needs the
  // complete gain; the steps before it advance P.
  for (; j >= 1 && t + j >= n_lin; --j)
    riccati_chain_step<SW, false>(P, Lf, dt, Qh, col, w.R(0, 1), w.R(1, 1), K, inv_u11, qsel);
  if (j == 0 && t >= n_lin) {
    riccati_chain_step<SW, true>(P, Lf, dt, Qh, col, w.R(0, 1), w.R(1, 1), K, inv_u11, qsel);
    return;
  }
  if (j < 0) return;
  // Window steps t+j, j downwards: the loads run three steps ahead of the arithmetic (a step is ~250 cycles of one
  // warp, a load from L2 under load takes longer), in a rotating set of three register buffers.
  LinD L0 = load_lin(lin, t + j, ld, b);
  LinD L1 = load_lin(lin, t + max(j - 1, 0), ld, b);
  LinD L2 = load_lin(lin, t + max(j - 2, 0), ld, b);
  // (each buffer is refilled right AFTER the step that consumed it: a load into a temporary issued before the step
  // and copied afterwards makes the copy wait for the load - ncu: 30 % of all stall samples on that one MOV)
  for (; j >= 3; j -= 3) {
    riccati_chain_step<SW, false>(P, L0, dt, Qh, col, w.R(0, 1), w.R(1, 1), K, inv_u11, qsel);
    L0 = load_lin(lin, t + j - 3, ld, b);
    riccati_chain_step<SW, false>(P, L1, dt, Qh, col, w.R(0, 1), w.R(1, 1), K, inv_u11, qsel);
    L1 = load_lin(lin, t + max(j - 4, 0), ld, b);
    riccati_chain_step<SW, false>(P, L2, dt, Qh, col, w.R(0, 1), w.R(1, 1), K, inv_u11, qsel);
    L2 = load_lin(lin, t + max(j - 5, 0), ld, b);
  }
  // j in {0, 1, 2} steps left: L0 = step t+j, L1 = t+j-1, L2 = t+j-2
  if (j == 2) {
    riccati_chain_step<SW, false>(P, L0, dt, Qh, col, w.R(0, 1), w.R(1, 1), K, inv_u11, qsel);
    riccati_chain_step<SW, false>(P, L1, dt, Qh, col, w.R(0, 1), w.R(1, 1), K, inv_u11, qsel);
    riccati_chain_step<SW, true>(P, L2, dt, Qh, col, w.R(0, 1), w.R(1, 1), K, inv_u11, qsel);
  } else if (j == 1) {
    riccati_chain_step<SW, false>(P, L0, dt, Qh, col, w.R(0, 1), w.R(1, 1), K, inv_u11, qsel);
    riccati_chain_step<SW, true>(P, L1, dt, Qh, col, w.R(0, 1), w.R(1, 1), K, inv_u11, qsel);
  } else {
    riccati_chain_step<SW, true>(P, L0, dt, Qh, col, w.R(0, 1), w.R(1, 1), K, inv_u11, qsel);
  }
}
// The pivot row of the 2x2 factorisation is the same for every step of every sweep: decide it once (per thread with
// per-problem weights, for the whole batch otherwise).
template <bool WPB>
__device__ __forceinline__ void mpc_sweep(const WV<WPB>& w, double dt, const double* __restrict__ lin, int64_t ld,
                                          int64_t b, int n_lin, const LinD& Lf, const double QT[10], int t, int H,
                                          double K[8]) {
  if (WPB)
    mpc_sweep_sw<WPB, 2>(w, dt, lin, ld, b, n_lin, Lf, QT, t, H, K);
  else if (fabs(w.R(0, 1)) > fabs(w.R(0, 0)))
    mpc_sweep_sw<WPB, 1>(w, dt, lin, ld, b, n_lin, Lf, QT, t, H, K);
  else
    mpc_sweep_sw<WPB, 0>(w, dt, lin, ld, b, n_lin, Lf, QT, t, H, K);
}

// Several consecutive receding-horizon solves in one pass: the windows of steps t, t+1, .. overlap in all but a few
// rows, the first-move gains do not depend on the state, and a single sweep is a dependent chain that leaves the FP64
// pipe idle two thirds of the time (ncu: 36 % of the issue slots wait for loads, 29 % for the previous DFMA).  Row k
// of the window of step t (absolute time t + k) is step k - c of the sweep that starts at t + c: each row is loaded
// once and advances ACRO_MPC_NS independent chains.  Same expressions per chain as mpc_sweep_sw.  H >= NS + 1.
//
// The rows come through a ring of ACRO_MPC_RING rows per warp in shared memory, filled with cp.async (every thread
// copies the ten values of its own problem, so no synchronisation between lanes) and awaited with
// cp.async.wait_group, which counts GROUPS: the wait for row k leaves the copies of the rows after it in flight.  A
// register prefetch cannot do that here: ptxas put the loads of all three rotating register buffers on one scoreboard
// and the first use after the loop's back edge waited for every load in flight, the youngest included (ncu: 23 % of
// all stall samples on that one DFMA; profiles/r2_mpc_pp_ncu_summary.txt).
#ifndef ACRO_MPC_RING
#define ACRO_MPC_RING 6
#endif
#ifndef ACRO_MPC_NS
#define ACRO_MPC_NS 3  // solves per pass (2: 7.56 ms, 3: 7.14 ms, 4: 6.98 ms with spills, at B = 16 384, H = 75)
#endif
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// ring: this thread's column of its warp's ring, element (slot, c) at ring[(slot * 10 + c) * 32].
// NS solves per pass (steps t .. t+NS-1; K[c] = first-move gain of step t + c): row k is step k - c of sweep c.
template <bool WPB, int SW, int NS>
__device__ __forceinline__ void mpc_sweepn_sw(const WV<WPB>& w, double dt, const double* __restrict__ lin, int64_t ld,
                                              int64_t b, int n_lin, const LinD& Lf, const double QT[10], int t, int H,
                                              double* ring, double (&K)[NS][8]) {
  double P[NS][10], iu, qs;
#pragma unroll
  for (int c = 0; c < NS; ++c)
#pragma unroll
    for (int e = 0; e < 10; ++e) P[c][e] = QT[e];
  const QhQ<WV<WPB>> Qh{w};
  const Lu2Col col = lu2_col(w.R(0, 0), w.R(0, 1));
  const double R01 = w.R(0, 1), R11 = w.R(1, 1);
  auto step = [&](const LinD& L, int k) {
    if (k >= NS && k <= H - 2) {  // an inner step of every sweep: one basic block, the chains interleave
#pragma unroll
      for (int c = 0; c < NS; ++c) riccati_chain_step<SW, false>(P[c], L, dt, Qh, col, R01, R11, K[c], iu, qs);
    } else {  // the NS-1 rows at either end of the pass: some sweeps only; the last step of a sweep gives its complete gain
#pragma unroll
      for (int c = 0; c < NS; ++c) {
        if (k == c)
          riccati_chain_step<SW, true>(P[c], L, dt, Qh, col, R01, R11, K[c], iu, qs);
        else if (k > c && k <= H - 2 + c)
          riccati_chain_step<SW, false>(P[c], L, dt, Qh, col, R01, R11, K[c], iu, qs);
      }
    }
  };
  auto fetch = [&](int k, int slot) {  // row k -> ring slot (an empty group below row 0 keeps the group count uniform)
    if (k >= 0) {
#pragma unroll
      for (int c = 0; c < 10; ++c) cp_async8(ring + (slot * 10 + c) * 32, lin + soa(t + k, 10, c, ld, b));
    }
    cp_async_commit();
  };
  const int ktop = H - 2 + NS - 1;               // top row of the last sweep
  const int kl = min(ktop, n_lin - 1 - t);       // highest row with a stored linearisation (>= NS - 1)
#pragma unroll
  for (int i = 0; i < ACRO_MPC_RING; ++i) fetch(kl - i, i);
  // padded tail of the window: linearisation about (x_f, u_f)
  for (int k = ktop; k > kl; --k) step(Lf, k);
  int slot = 0;
  for (int k = kl; k >= 0; --k) {
    cp_async_wait<ACRO_MPC_RING - 1>();
    LinD L;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      L.a[0][j] = ring[(slot * 10 + j) * 32];
      L.a[1][j] = ring[(slot * 10 + 4 + j) * 32];
    }
    L.b[0] = ring[(slot * 10 + 8) * 32];
    L.b[1] = ring[(slot * 10 + 9) * 32];
    L.b0[0] = L.b0[1] = 0.0;
    step(L, k);
    fetch(k - ACRO_MPC_RING, slot);  // after the step: the row has been consumed
    slot = (slot + 1 == ACRO_MPC_RING) ? 0 : slot + 1;
  }
  cp_async_wait<0>();
}
template <bool WPB, int NS>
__device__ __forceinline__ void mpc_sweepn(const WV<WPB>& w, double dt, const double* __restrict__ lin, int64_t ld,
                                           int64_t b, int n_lin, const LinD& Lf, const double QT[10], int t, int H,
                                           double* ring, double (&K)[NS][8]) {
  if (WPB)
    mpc_sweepn_sw<WPB, 2, NS>(w, dt, lin, ld, b, n_lin, Lf, QT, t, H, ring, K);
  else if (fabs(w.R(0, 1)) > fabs(w.R(0, 0)))
    mpc_sweepn_sw<WPB, 1, NS>(w, dt, lin, ld, b, n_lin, Lf, QT, t, H, ring, K);
  else
    mpc_sweepn_sw<WPB, 0, NS>(w, dt, lin, ld, b, n_lin, Lf, QT, t, H, ring, K);
}

// shared reference: one thread per time step computes K0[t]
template <bool WPB>
__global__ void k_mpc_gains_shared(const __grid_constant__ MpcArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.T - 1) return;
  const WV<WPB> w(a.kw, 1, 0);
  const LinD Lf = linearize_d(a.m, a.xf, a.uf[0], a.uf[1]);
  double QT[10];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) QT[sym(i, j)] = a.QT[i * 4 + j];
  double K[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  mpc_sweep(w, a.m.dt, a.lin, 0, 0, a.N - 1, Lf, QT, t, a.H, K);
#pragma unroll
  for (int e = 0; e < 8; ++e) a.K0[t * 8 + e] = K[e];
}

// closed-loop plant simulation with the first-move gains (shared reference):
// u = u_ref_win[0] + K0_t (x - x_ref_win[0]) (tt:46-56); window index 0 is time t (x_f beyond N-1)
__global__ void k_mpc_track_shared(const __grid_constant__ MpcArgs a) {
  const int64_t B = a.B, b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  double x[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    x[c] = a.x0[c * B + b];
    a.Xr[soa(0, 4, c, a.T, b)] = x[c];
  }
  TrigCarry tc = trig_carry_at(a.m, x[0], x[1]);
  for (int t = 0; t < a.T - 1; ++t) {
    double u[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double kd = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double xr = (t < a.N) ? __ldg(a.rx + t * 4 + j) : a.xf[j];
        const double k = __ldg(a.K0 + t * 8 + i * 4 + j);
        kd = (j == 0) ? k * (x[0] - xr) : fma(k, x[j] - xr, kd);
      }
      const double ur = (t < a.N - 1) ? __ldg(a.ru + t * 2 + i) : a.uf[i];
      u[i] = ur + kd;
      a.Ur[soa(t, 2, i, a.T - 1, b)] = u[i];
    }
    double xn[4];
    rk4_step_inc(a.m, x, u[0], u[1], xn, tc);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      x[c] = xn[c];
      a.Xr[soa(t + 1, 4, c, a.T, b)] = x[c];
    }
  }
}

// per-problem reference / weights: every problem runs its own T-1 sweeps.  PPB: every problem its own physical
// parameters (model_per_problem): the plant step and the padding linearisation about (x_f, u_f) use them, as did
// k_lin_compact for the window rows.
template <bool WPB, bool PPB = false>
__global__ void k_mpc_track_pp(const __grid_constant__ MpcArgs a) {
  const int64_t B = a.B, b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  const WV<WPB> w(a.kw, B, b);
  const RefV<true> ref{a.rx, a.ru, a.N, b};
  Model m_loc_;
  if (PPB) m_loc_ = model_per_problem(a.m, a.pb, B, b);
  const Model& mm = PPB ? m_loc_ : a.m;
  const LinD Lf = linearize_d(mm, a.xf, a.uf[0], a.uf[1]);
  double QT[10];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) QT[sym(i, j)] = a.qt_per_problem ? a.QT[(i * 4 + j) * B + b] : a.QT[i * 4 + j];
  double x[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    x[c] = a.x0[c * B + b];
    a.Xr[soa(0, 4, c, a.T, b)] = x[c];
  }
  // closed-loop step of the plant with the first-move gain of the solve at step t (tt:46-56)
  auto plant = [&](int t, const double K[8]) {
    double u[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double kd = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const double xr = (t < a.N) ? ref.X(t, j) : a.xf[j];
        kd = (j == 0) ? K[i * 4] * (x[0] - xr) : fma(K[i * 4 + j], x[j] - xr, kd);
      }
      const double ur = (t < a.N - 1) ? ref.U(t, i) : a.uf[i];
      u[i] = ur + kd;
      a.Ur[soa(t, 2, i, a.T - 1, b)] = u[i];
    }
    double xn[4];
    rk4_step(mm, x, u[0], u[1], xn);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      x[c] = xn[c];
      a.Xr[soa(t + 1, 4, c, a.T, b)] = x[c];
    }
  };
  extern __shared__ double mpc_ring[];  // ACRO_MPC_RING rows of 10 x 32 doubles per warp
  double* const ring = mpc_ring + (threadIdx.x >> 5) * (ACRO_MPC_RING * 320) + (threadIdx.x & 31);
  int t = 0;
  if (a.H >= ACRO_MPC_NS + 1) {  // ACRO_MPC_NS solves per pass over their common window rows
    for (; t + ACRO_MPC_NS - 1 < a.T - 1; t += ACRO_MPC_NS) {
      double Kn[ACRO_MPC_NS][8];
#pragma unroll
      for (int c = 0; c < ACRO_MPC_NS; ++c)
#pragma unroll
        for (int e = 0; e < 8; ++e) Kn[c][e] = 0.0;
      mpc_sweepn<WPB, ACRO_MPC_NS>(w, a.m.dt, a.lin, a.N - 1, b, a.N - 1, Lf, QT, t, a.H, ring, Kn);
#pragma unroll
      for (int c = 0; c < ACRO_MPC_NS; ++c) plant(t + c, Kn[c]);
    }
  }
  for (; t < a.T - 1; ++t) {
    double K[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    mpc_sweep(w, a.m.dt, a.lin, a.N - 1, b, a.N - 1, Lf, QT, t, a.H, K);
    plant(t, K);
  }
}

}  // namespace acro
#include "acro_mpc_box.cuh"
namespace acro {

// ---------------------------------------------------------------------------------------
// layout helpers, tiled through shared memory
// ---------------------------------------------------------------------------------------
__global__ void k_transpose(int64_t rows, int64_t cols, const double* __restrict__ src, double* __restrict__ dst) {
  // src (rows, cols) row-major -> dst (cols, rows) row-major
  __shared__ double tile[32][33];
  const int64_t tiles_c = (cols + 31) / 32;
  const int64_t c0 = (blockIdx.x % tiles_c) * 32LL, r0 = (blockIdx.x / tiles_c) * 32LL;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = src[r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[c * rows + r] = tile[threadIdx.x][i];
  }
}

// batch-major (B, T, C) <-> tiled [tile][t][c][lane].  One block = one tile of 32 problems x 32 consecutive
// flattened (t, c) indices; both sides are read / written 256 bytes at a time.  Padding lanes are zero-filled.
template <bool PACK>
__global__ void k_tiled(int64_t B, int T, int C, const double* __restrict__ src, double* __restrict__ dst) {
  __shared__ double tile[32][33];
  const int64_t TC = int64_t(T) * C, chunks = (TC + 31) / 32;
  const int64_t k0 = (blockIdx.x % chunks) * 32LL, b0 = (blockIdx.x / chunks) * 32LL;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    if (PACK) {  // row i = problem b0+i, column = flattened index k0 + x
      const int64_t b = b0 + i, k = k0 + threadIdx.x;
      tile[i][threadIdx.x] = (b < B && k < TC) ? src[b * TC + k] : 0.0;
    } else {     // row i = flattened index k0+i, column = lane x
      const int64_t k = k0 + i, b = b0 + threadIdx.x;
      if (k < TC) tile[i][threadIdx.x] = src[soa(int(k / C), C, int(k % C), T, b)];
    }
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    if (PACK) {
      const int64_t k = k0 + i, b = b0 + threadIdx.x;
      if (k < TC) dst[soa(int(k / C), C, int(k % C), T, b)] = tile[threadIdx.x][i];
    } else {
      const int64_t b = b0 + i, k = k0 + threadIdx.x;
      if (b < B && k < TC) dst[b * TC + k] = tile[threadIdx.x][i];
    }
  }
}

// ---------------------------------------------------------------------------------------
// FP64 pipe peak: 8 independent DFMA chains per thread (the roofline denominator bench.py
// measures in the same run, since MEASURED_PEAKS.json has no FP64 entry).
// ---------------------------------------------------------------------------------------
// one dependent DFMA chain per thread: with one warp per block this measures the DFMA latency
// `chains` (1..8) independent chains per thread; only lanes < active_lanes of each warp work
template <int CH>
__device__ __forceinline__ double chain_body(int iters, double b, double c) {
  double a[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) a[k] = 1e-3 * (threadIdx.x + k);
  for (int i = 0; i < iters; i += 4) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int k = 0; k < CH; ++k) a[k] = fma(a[k], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < CH; ++k) s += a[k];
  return s;
}
__global__ void k_fp64_chain(double* __restrict__ out, int iters, double b, double c, long long* __restrict__ cycles,
                             int chains, int active_lanes) {
  double a = 0.0;
  const long long t0 = clock64();
  if ((threadIdx.x & 31) < active_lanes) {
    if (chains == 1) a = chain_body<1>(iters, b, c);
    else if (chains == 2) a = chain_body<2>(iters, b, c);
    else if (chains == 4) a = chain_body<4>(iters, b, c);
    else a = chain_body<8>(iters, b, c);
  }
  const long long t1 = clock64();
  out[blockIdx.x * int64_t(blockDim.x) + threadIdx.x] = a;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void k_fp64_peak(double* __restrict__ out, int iters, double b, double c) {
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 1e-3 * (threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = fma(a[k], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += a[k];
  out[blockIdx.x * int64_t(blockDim.x) + threadIdx.x] = s;
}

}  // namespace acro

namespace acro {
// Which Newton kernel a call runs: the one place that decides (acro_newton_solve and acro_newton_describe share it).
struct NewtonPlan {
  int kernel, stage_steps, recompute_lin;
};
#define ACRO_SPEC_AUTO_GAMMA 0.5  // initial step sizes from here on select the speculative kernel automatically
static int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    n = 148;  // B200 (no device: acro_newton_describe on a CPU-only box)
  return n;
}
static int newton_plan(const AcroNewtonOpts& o, int64_t B, bool rpb, bool wpb, bool ppb, bool act, bool tma_ok, NewtonPlan& plan) {
  (void)wpb;
  const int64_t tiles = (B + 31) / 32;
  const int n_sm = sm_count();
  int k = o.kernel, sg = o.stage_steps, rl = o.recompute_lin;
  ACRO_REQUIRE(k >= ACRO_NEWTON_AUTO && k <= ACRO_NEWTON_SPEC, "acro_newton_solve: unknown kernel");
  ACRO_REQUIRE(sg == 0 || sg == 2 || sg == 4 || sg == 8 || sg == 16, "acro_newton_solve: stage_steps must be 0, 2, 4, 8 or 16");
  ACRO_REQUIRE(rl >= 0 && rl <= 2, "acro_newton_solve: recompute_lin must be 0, 1 or 2");
  if (k == ACRO_NEWTON_AUTO) {
    // the fully-actuated plant runs on the one-thread-per-problem kernel; buffers that are not 128-byte aligned cannot
    // be the source of bulk copies: same kernel.
    if (act || !tma_ok) k = ACRO_NEWTON_THREAD;
    // at most two tiles per SM (config 2: 128 tiles): split every tile between two warps on two SM sub-partitions;
    // at most one tile per SM and a large initial step size (the reference's default gamma_0 = 1 back-tracks several times
    // per iteration; its shipped recipes use 0.05 and 0.1, which never do): evaluate the Armijo candidates in parallel.
    // Both kernels take the same decisions; the speculative one is 2.3x faster when the line search back-tracks and 25 %
    // slower when it does not (profiles/r2_spec_probe.txt).
    else if (tiles <= n_sm && !ppb && o.spec_ws != nullptr && o.gamma_0 >= ACRO_SPEC_AUTO_GAMMA) k = ACRO_NEWTON_SPEC;
    else k = (tiles <= 2 * int64_t(n_sm)) ? ACRO_NEWTON_DUO : ACRO_NEWTON_RING;
  }
  if (k != ACRO_NEWTON_THREAD) {
    ACRO_REQUIRE(tma_ok, "acro_newton_solve: the duo / ring kernels need 128-byte aligned X, U, Xw, Uw, lin_ws, K, S and reference buffers");
    ACRO_REQUIRE(!act, "acro_newton_solve: the fully-actuated plant runs on ACRO_NEWTON_THREAD");
  }
  if (k == ACRO_NEWTON_SPEC) {
    // eight warps per tile, one block per SM (acro_newton_spec.cuh); needs the candidate workspace
    ACRO_REQUIRE(tiles <= n_sm, "acro_newton_solve: the speculative kernel runs one tile per SM (B <= 32 x SM count)");
#ifndef ACRO_SPEC_TIMING
    ACRO_REQUIRE(!ppb, "acro_newton_solve: the speculative kernel takes shared physical parameters");
#endif
    ACRO_REQUIRE(o.spec_ws != nullptr, "acro_newton_solve: the speculative kernel needs AcroNewtonOpts.spec_ws (acro_newton_spec_ws_doubles)");
    ACRO_REQUIRE(o.speculate >= 0 && o.speculate <= 8, "acro_newton_solve: speculate must be 0 (adaptive) or 1..8");
    ACRO_REQUIRE(sg == 0 || (sg == 16 && !rpb) || (sg == 8 && rpb), "acro_newton_solve: speculative kernel: stage_steps 16 (8 with per-problem references)");
    sg = rpb ? 8 : 16;
    rl = 0;
  } else if (k == ACRO_NEWTON_DUO) {
    // one block per SM: deep stages (16 time steps per bulk copy, 220 KB of shared memory); two blocks per SM: 4-step
    // stages (70-90 KB).  Per-problem references need 50 % more shared memory per stage: 8 instead of 16.
    if (sg == 0) sg = (tiles > n_sm) ? 4 : (rpb ? 8 : 16);
    ACRO_REQUIRE(sg == 4 || sg == 8 || (sg == 16 && !rpb), "acro_newton_solve: duo kernel: stage_steps 4, 8 (or 16 with a shared reference)");
    rl = 0;
  } else if (k == ACRO_NEWTON_RING) {
    // at most one block per SM: 16 steps per stage; up to four: 4 steps; beyond: 2 steps (25 KB per block, eight blocks
    // = two warps per sub-partition)
    if (sg == 0) sg = (tiles <= n_sm && !rpb) ? 16 : (tiles > 4 * int64_t(n_sm) ? 2 : 4);
    ACRO_REQUIRE(sg == 2 || sg == 4 || (sg == 16 && !rpb), "acro_newton_solve: ring kernel: stage_steps 2, 4 (or 16 with a shared reference)");
    if (ppb && sg == 16) sg = 4;  // (per-problem physical parameters: 4- and 2-step stages are instantiated)
    rl = (rl == 0) ? (sg == 2) : (rl == 1);
    ACRO_REQUIRE(!rl || sg == 2, "acro_newton_solve: recompute_lin needs stage_steps = 2");
    ACRO_REQUIRE(!ppb || rl || sg == 4, "acro_newton_solve: ring kernel with per-problem physical parameters: stage_steps 4, or 2 with recompute_lin");
  } else {
    sg = 0;
    rl = 0;
  }
  plan.kernel = k;
  plan.stage_steps = sg;
  plan.recompute_lin = rl;
  return ACRO_OK;
}
}  // namespace acro

// =========================================================================================
// C ABI
// =========================================================================================
using namespace acro;

extern "C" {

const char* acro_version(void) { return "acro_b200 0.2.0 (sm_100a, abi 2)"; }
const char* acro_last_error_string(void) { return g_err; }
int64_t acro_launch_count(void) { return g_launches.load(); }

// ---- D1-D3, G1, T2 with shared (params_b = NULL) or per-problem physical parameters
int acro_continuous_dynamics_pp(const AcroParams* p, const double* params_b, int64_t B, const double* x, const double* u,
                                double* xdot, void* stream) {
  ACRO_REQUIRE(p && x && u && xdot && B > 0, "acro_continuous_dynamics: bad argument");
  const Cfg c = cfg_for(B);
  if (params_b)
    k_point<PT_F, true><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(make_model(*p), params_b, B, x, u, xdot);
  else
    k_point<PT_F, false><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(make_model(*p), nullptr, B, x, u, xdot);
  ACRO_LAUNCH_CHECK("acro_continuous_dynamics");
  return ACRO_OK;
}
int acro_continuous_dynamics(const AcroParams* p, int64_t B, const double* x, const double* u, double* xdot,
                             void* stream) {
  return acro_continuous_dynamics_pp(p, nullptr, B, x, u, xdot, stream);
}

int acro_rk4_step_pp(const AcroParams* p, const double* params_b, int64_t B, const double* x, const double* u,
                     double* xnext, void* stream) {
  ACRO_REQUIRE(p && x && u && xnext && B > 0, "acro_rk4_step: bad argument");
  const Cfg c = cfg_for(B);
  if (params_b)
    k_point<PT_RK4, true><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(make_model(*p), params_b, B, x, u, xnext);
  else
    k_point<PT_RK4, false><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(make_model(*p), nullptr, B, x, u, xnext);
  ACRO_LAUNCH_CHECK("acro_rk4_step");
  return ACRO_OK;
}
int acro_rk4_step(const AcroParams* p, int64_t B, const double* x, const double* u, double* xnext, void* stream) {
  return acro_rk4_step_pp(p, nullptr, B, x, u, xnext, stream);
}

int acro_linearize_pp(const AcroParams* p, const double* params_b, int64_t B, const double* x, const double* u,
                      double* A, double* Bm, int discrete, void* stream) {
  ACRO_REQUIRE(p && x && u && A && Bm && B > 0, "acro_linearize: bad argument");
  const Cfg c = cfg_for(B);
  if (params_b)
    k_linearize<true><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(make_model(*p), params_b, B, x, u, A, Bm, discrete);
  else
    k_linearize<false><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(make_model(*p), nullptr, B, x, u, A, Bm, discrete);
  ACRO_LAUNCH_CHECK("acro_linearize");
  return ACRO_OK;
}
int acro_linearize(const AcroParams* p, int64_t B, const double* x, const double* u, double* A, double* Bm,
                   int discrete, void* stream) {
  return acro_linearize_pp(p, nullptr, B, x, u, A, Bm, discrete, stream);
}

int acro_equilibrium(const AcroParams* p, const double* params_b, int64_t B, const double* u_target, const double* theta_guess,
                     double tol, int max_iter, double* theta, int32_t* n_iter, void* stream) {
  ACRO_REQUIRE(p && u_target && theta_guess && theta && n_iter && B > 0 && max_iter > 0 && tol > 0.0, "acro_equilibrium: bad argument");
  const Cfg c = cfg_for(B);
  if (params_b)
    k_equilibrium<true><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(make_model(*p), params_b, B, u_target, theta_guess, tol, max_iter, theta, n_iter);
  else
    k_equilibrium<false><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(make_model(*p), nullptr, B, u_target, theta_guess, tol, max_iter, theta, n_iter);
  ACRO_LAUNCH_CHECK("acro_equilibrium");
  return ACRO_OK;
}

int acro_rollout_open_loop_pp(const AcroParams* p, const double* params_b, int64_t B, int N, const double* x0,
                              const double* U, double* X, void* stream) {
  ACRO_REQUIRE(p && x0 && X && B > 0 && N >= 1, "acro_rollout_open_loop: bad argument");
  const Cfg c = cfg_for(B);
  if (params_b)
    k_rollout_open<true><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(make_model(*p), params_b, B, N, x0, U, X);
  else
    k_rollout_open<false><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(make_model(*p), nullptr, B, N, x0, U, X);
  ACRO_LAUNCH_CHECK("acro_rollout_open_loop");
  return ACRO_OK;
}
int acro_rollout_open_loop(const AcroParams* p, int64_t B, int N, const double* x0, const double* U, double* X,
                           void* stream) {
  return acro_rollout_open_loop_pp(p, nullptr, B, N, x0, U, X, stream);
}

int acro_total_cost(const AcroWeights* w, int64_t B, int N, const double* X, const double* U, const AcroRef* ref,
                    double* cost, void* stream) {
  ACRO_REQUIRE(w && X && U && ref && ref->x && ref->u && cost && B > 0 && N >= 2, "acro_total_cost: bad argument");
  const Cfg c = cfg_for(B);
  const KWeights kw = make_weights(*w);
#define EXPR(WPB, RPB) \
  k_total_cost<WPB, RPB><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(kw, B, N, X, U, ref->x, ref->u, cost)
  DISPATCH2(per_problem_weights(*w), ref->per_problem != 0, EXPR);
#undef EXPR
  ACRO_LAUNCH_CHECK("acro_total_cost");
  return ACRO_OK;
}

int acro_costate(const AcroParams* p, const AcroWeights* w, int64_t B, int N, const double* X, const double* U,
                 const AcroRef* ref, double* lam, void* stream) {
  ACRO_REQUIRE(p && w && X && U && ref && ref->x && ref->u && lam && B > 0 && N >= 2, "acro_costate: bad argument");
  const Cfg c = cfg_for(B);
  const KWeights kw = make_weights(*w);
  const Model m = make_model(*p);
#define EXPR(WPB, RPB) \
  k_costate<WPB, RPB><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(m, kw, B, N, X, U, ref->x, ref->u, lam)
  DISPATCH2(per_problem_weights(*w), ref->per_problem != 0, EXPR);
#undef EXPR
  ACRO_LAUNCH_CHECK("acro_costate");
  return ACRO_OK;
}

int acro_riccati_affine(const AcroParams* p, const AcroWeights* w, int64_t B, int N, const double* X,
                        const double* U, const AcroRef* ref, double* K, double* S, double* delta_J,
                        double* sigma_norm, void* stream) {
  ACRO_REQUIRE(p && w && X && U && ref && ref->x && ref->u && K && S && delta_J && sigma_norm && B > 0 && N >= 2,
               "acro_riccati_affine: bad argument");
  const Cfg c = cfg_for(B);
  const KWeights kw = make_weights(*w);
  const Model m = make_model(*p);
  if (p->actuated_tau1) {
#define EXPR(WPB, RPB)                                                                                   \
  k_riccati_affine<WPB, RPB, true><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(m, kw, B, N, X, U, ref->x, ref->u, K, S, \
                                                                                 delta_J, sigma_norm)
    DISPATCH2(per_problem_weights(*w), ref->per_problem != 0, EXPR);
#undef EXPR
  } else {
#define EXPR(WPB, RPB)                                                                                   \
  k_riccati_affine<WPB, RPB><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(m, kw, B, N, X, U, ref->x, ref->u, K, S, \
                                                                           delta_J, sigma_norm)
    DISPATCH2(per_problem_weights(*w), ref->per_problem != 0, EXPR);
#undef EXPR
  }
  ACRO_LAUNCH_CHECK("acro_riccati_affine");
  return ACRO_OK;
}

int acro_closed_loop_rollout_cost(const AcroParams* p, const AcroWeights* w, int64_t B, int N, const double* X,
                                  const double* U, const double* K, const double* S, const AcroRef* ref, int G,
                                  const double* gammas, int gammas_per_problem, double* Xn, double* Un,
                                  double* cost, void* stream) {
  ACRO_REQUIRE(p && w && X && U && K && S && ref && ref->x && ref->u && gammas && cost && B > 0 && N >= 2 && G > 0,
               "acro_closed_loop_rollout_cost: bad argument");
  ACRO_REQUIRE((Xn == nullptr) == (Un == nullptr), "acro_closed_loop_rollout_cost: Xn and Un go together");
  ACRO_REQUIRE(G <= 65535, "acro_closed_loop_rollout_cost: G too large");
  Cfg c = cfg_for(B * G);
  if (c.block > 32 && B < 148LL * 32 * 4) c.block = 32;
  const dim3 grid((unsigned)((B + c.block - 1) / c.block), (unsigned)G);
  const KWeights kw = make_weights(*w);
  const Model m = make_model(*p);
#define EXPR(WPB, RPB)                                                                                      \
  k_closed_loop<WPB, RPB><<<grid, c.block, 0, (cudaStream_t)stream>>>(m, kw, B, N, X, U, K, S, ref->x, ref->u, G, \
                                                                      gammas, gammas_per_problem, Xn, Un, cost)
  DISPATCH2(per_problem_weights(*w), ref->per_problem != 0, EXPR);
#undef EXPR
  ACRO_LAUNCH_CHECK("acro_closed_loop_rollout_cost");
  return ACRO_OK;
}

int acro_armijo_select(int64_t B, int G, const double* cost_k, const double* delta_J, const double* gammas,
                       int gammas_per_problem, const double* cost_cand, double c, int32_t* accepted, void* stream) {
  ACRO_REQUIRE(cost_k && delta_J && gammas && cost_cand && accepted && B > 0 && G > 0, "acro_armijo_select: bad argument");
  const int block = 128;
  k_armijo_select<<<(unsigned)((B + block - 1) / block), block, 0, (cudaStream_t)stream>>>(
      B, G, cost_k, delta_J, gammas, gammas_per_problem, cost_cand, c, accepted);
  ACRO_LAUNCH_CHECK("acro_armijo_select");
  return ACRO_OK;
}

int acro_newton_solve(const AcroParams* p, const AcroWeights* w, const AcroNewtonOpts* opts, int64_t B, int N,
                      const double* x0, const AcroRef* ref, double* X, double* U, double* Xw, double* Uw,
                      double* lin_ws, double* K, double* S, double* cost, double* delta_J, double* sigma_norm, double* gamma_acc,
                      int32_t* iters, int32_t* status, double* hist_cost, double* hist_sigma_norm,
                      double* hist_gamma, int32_t* hist_ntry, void* stream) {
  return acro_newton_solve_pp(p, nullptr, w, opts, B, N, x0, ref, X, U, Xw, Uw, lin_ws, K, S, cost, delta_J, sigma_norm,
                              gamma_acc, iters, status, hist_cost, hist_sigma_norm, hist_gamma, hist_ntry, stream);
}

int acro_newton_solve_pp(const AcroParams* p, const double* params_b, const AcroWeights* w, const AcroNewtonOpts* opts,
                         int64_t B, int N, const double* x0, const AcroRef* ref, double* X, double* U, double* Xw,
                         double* Uw, double* lin_ws, double* K, double* S, double* cost, double* delta_J,
                         double* sigma_norm, double* gamma_acc, int32_t* iters, int32_t* status, double* hist_cost,
                         double* hist_sigma_norm, double* hist_gamma, int32_t* hist_ntry, void* stream) {
  ACRO_REQUIRE(p && w && opts && ref && ref->x && ref->u && X && U && Xw && Uw && lin_ws && K && S && cost && delta_J &&
                   sigma_norm && gamma_acc && iters && status && B > 0 && N >= 2,
               "acro_newton_solve: bad argument");
  ACRO_REQUIRE(!opts->init || x0, "acro_newton_solve: x0 required when init");
  ACRO_REQUIRE(opts->max_iters >= 0 && opts->max_line_search >= 1, "acro_newton_solve: bad options");
  ACRO_REQUIRE(!(p->actuated_tau1 && params_b), "acro_newton_solve: the fully-actuated plant takes shared physical parameters");
  NewtonArgs a;
  a.m = make_model(*p);
  a.kw = make_weights(*w);
  a.o = *opts;
  a.B = B;
  a.N = N;
  a.x0 = x0;
  a.pb = params_b;
  a.rx = ref->x;
  a.ru = ref->u;
  a.X = X;
  a.U = U;
  a.Xw = Xw;
  a.Uw = Uw;
  a.lin = lin_ws;
  a.spec_ws = opts->spec_ws;
  a.K = K;
  a.S = S;
  a.cost = cost;
  a.dJ = delta_J;
  a.sn = sigma_norm;
  a.gacc = gamma_acc;
  a.iters = iters;
  a.status = status;
  a.h_cost = hist_cost;
  a.h_sn = hist_sigma_norm;
  a.h_gamma = hist_gamma;
  a.h_ntry = hist_ntry;
  auto aligned = [](const void* q, uintptr_t al) { return (reinterpret_cast<uintptr_t>(q) % al) == 0; };
  const bool tma_ok = aligned(X, 128) && aligned(U, 128) && aligned(Xw, 128) && aligned(Uw, 128) &&
                      aligned(lin_ws, 128) && aligned(K, 128) && aligned(S, 128) &&
                      aligned(ref->x, ref->per_problem ? 128 : 32) && aligned(ref->u, ref->per_problem ? 128 : 16);
  NewtonPlan plan;
  // the fully-actuated plant runs on the one-thread-per-problem kernel (planned like per-problem physical parameters)
  const int rc = newton_plan(*opts, B, ref->per_problem != 0, per_problem_weights(*w), params_b != nullptr,
                             p->actuated_tau1 != 0, tma_ok, plan);
  if (rc != ACRO_OK) return rc;
  const int64_t tiles = (B + 31) / 32;
  const bool wpb = per_problem_weights(*w), rpb = ref->per_problem != 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (plan.kernel == ACRO_NEWTON_THREAD) {
    const Cfg c = cfg_for(B);
    if (p->actuated_tau1) {
#define EXPR(WPB, RPB) k_newton<WPB, RPB, false, true><<<c.grid, c.block, 0, s>>>(a)
      DISPATCH2(wpb, rpb, EXPR);
#undef EXPR
    } else if (params_b) {
#define EXPR(WPB, RPB) k_newton<WPB, RPB, true><<<c.grid, c.block, 0, s>>>(a)
      DISPATCH2(wpb, rpb, EXPR);
#undef EXPR
    } else {
#define EXPR(WPB, RPB) k_newton<WPB, RPB><<<c.grid, c.block, 0, s>>>(a)
      DISPATCH2(wpb, rpb, EXPR);
#undef EXPR
    }
  } else if (plan.kernel == ACRO_NEWTON_SPEC) {
#define LAUNCH_SPEC(WPB, RPB, SG)                                                                                   \
  do {                                                                                                              \
    constexpr int smem = SpecSmem<RPB, SG>::total;                                                                  \
    cudaError_t e_ = cudaFuncSetAttribute(k_newton_spec<WPB, RPB, SG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          smem);                                                                    \
    if (e_ != cudaSuccess) return cuda_fail(e_, "acro_newton_solve/smem");                                          \
    k_newton_spec<WPB, RPB, SG><<<(unsigned)tiles, ACRO_SPEC_W * 32, smem, s>>>(a);                                 \
  } while (0)
    if (rpb) {
      LAUNCH_SPEC(true, true, 8);
    } else {
#define EXPR(WPB, RPB) LAUNCH_SPEC(WPB, false, 16)
      DISPATCH2(wpb, false, EXPR);
#undef EXPR
    }
#undef LAUNCH_SPEC
  } else if (plan.kernel == ACRO_NEWTON_DUO) {
#define LAUNCH_DUO(WPB, RPB, SG)                                                                                   \
  do {                                                                                                             \
    constexpr int smem = DuoSmem<RPB, SG>::total;                                                                  \
    cudaError_t e_ = cudaFuncSetAttribute(k_newton_duo<WPB, RPB, SG>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          smem);                                                                   \
    if (e_ != cudaSuccess) return cuda_fail(e_, "acro_newton_solve/smem");                                         \
    k_newton_duo<WPB, RPB, SG><<<(unsigned)tiles, 64, smem, s>>>(a);                                               \
  } while (0)
    // Per-problem references always run the variant that keeps the weights in registers (it takes shared weights too:
    // WV<true> falls back to them): ptxas puts YIELDs at the loop heads of k_newton_duo<false, true, *> and of no
    // other variant, which costs 12 % (8.7 against 9.9 M it/s at B = 4096).
    const bool wreg = wpb || rpb;
    if (params_b) {
      // per-problem physical parameters: the variants with the weights in registers (they take shared weights too)
#define LAUNCH_DUO_PPB(RPB, SG)                                                                                          \
  do {                                                                                                                   \
    constexpr int smem = DuoSmem<RPB, SG>::total;                                                                        \
    cudaError_t e_ = cudaFuncSetAttribute(k_newton_duo<true, RPB, SG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          smem);                                                                         \
    if (e_ != cudaSuccess) return cuda_fail(e_, "acro_newton_solve/smem");                                               \
    k_newton_duo<true, RPB, SG, true><<<(unsigned)tiles, 64, smem, s>>>(a);                                              \
  } while (0)
      if (plan.stage_steps == 4) {
        if (rpb) LAUNCH_DUO_PPB(true, 4); else LAUNCH_DUO_PPB(false, 4);
      } else if (rpb) {
        LAUNCH_DUO_PPB(true, 8);
      } else if (plan.stage_steps == 8) {
        LAUNCH_DUO_PPB(false, 8);
      } else {
        LAUNCH_DUO_PPB(false, 16);
      }
#undef LAUNCH_DUO_PPB
    } else if (plan.stage_steps == 4) {
#define EXPR(WPB, RPB) LAUNCH_DUO(WPB, RPB, 4)
      DISPATCH2(wreg, rpb, EXPR);
#undef EXPR
    } else if (plan.stage_steps == 8 && !rpb) {
#define EXPR(WPB, RPB) LAUNCH_DUO(WPB, false, 8)
      DISPATCH2(wpb, false, EXPR);
#undef EXPR
    } else if (plan.stage_steps == 8) {
      LAUNCH_DUO(true, true, 8);
    } else {
#define EXPR(WPB, RPB) LAUNCH_DUO(WPB, false, 16)
      DISPATCH2(wpb, false, EXPR);
#undef EXPR
    }
#undef LAUNCH_DUO
  } else {
#define LAUNCH_RING(WPB, RPB, SG, RL)                                                                               \
  do {                                                                                                              \
    constexpr int smem = ACRO_RING_D * stage_bytes<RPB, SG>() + ACRO_RING_D * 8;                                    \
    cudaError_t e_ = cudaFuncSetAttribute(k_newton_ring<WPB, RPB, SG, RL>,                                          \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                       \
    if (e_ != cudaSuccess) return cuda_fail(e_, "acro_newton_solve/smem");                                          \
    k_newton_ring<WPB, RPB, SG, RL><<<(unsigned)tiles, 32, smem, s>>>(a);                                           \
  } while (0)
    if (params_b) {
#define LAUNCH_RING_PPB(RPB, SG, RL)                                                                                  \
  do {                                                                                                                  \
    constexpr int smem = ACRO_RING_D * stage_bytes<RPB, SG>() + ACRO_RING_D * 8;                                        \
    cudaError_t e_ = cudaFuncSetAttribute(k_newton_ring<true, RPB, SG, RL, true>,                                       \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem);                           \
    if (e_ != cudaSuccess) return cuda_fail(e_, "acro_newton_solve/smem");                                              \
    k_newton_ring<true, RPB, SG, RL, true><<<(unsigned)tiles, 32, smem, s>>>(a);                                        \
  } while (0)
      if (plan.stage_steps == 2) {
        if (rpb) LAUNCH_RING_PPB(true, 2, true); else LAUNCH_RING_PPB(false, 2, true);
      } else {
        if (rpb) LAUNCH_RING_PPB(true, 4, false); else LAUNCH_RING_PPB(false, 4, false);
      }
#undef LAUNCH_RING_PPB
    } else if (plan.stage_steps == 16) {
#define EXPR(WPB, RPB) LAUNCH_RING(WPB, false, 16, false)
      DISPATCH2(wpb, false, EXPR);
#undef EXPR
    } else if (plan.stage_steps == 2 && plan.recompute_lin) {
      // more tiles than SM sub-partitions: this regime is HBM bound, and the backward pass recomputes the linearisation
      // about (x_t, u_t) instead of streaming the 80 B the forward pass would have stored for it (304 instead of 464 B
      // per problem-step-iteration)
#define EXPR(WPB, RPB) LAUNCH_RING(WPB, RPB, 2, true)
      DISPATCH2(wpb, rpb, EXPR);
#undef EXPR
    } else if (plan.stage_steps == 2) {
#define EXPR(WPB, RPB) LAUNCH_RING(WPB, RPB, 2, false)
      DISPATCH2(wpb, rpb, EXPR);
#undef EXPR
    } else {
#define EXPR(WPB, RPB) LAUNCH_RING(WPB, RPB, 4, false)
      DISPATCH2(wpb, rpb, EXPR);
#undef EXPR
    }
#undef LAUNCH_RING
  }
  ACRO_LAUNCH_CHECK("acro_newton_solve");
  return ACRO_OK;
}

int64_t acro_newton_spec_ws_doubles(int64_t B, int N) {
  if (B <= 0 || N < 2) return 0;
  return int64_t(ACRO_SPEC_W) * ((B + 31) / 32) * 32 * (int64_t(N) * 4 + int64_t(N - 1) * 2);
}

int acro_newton_describe(const AcroNewtonOpts* opts, int64_t B, int ref_per_problem, int weights_per_problem,
                         int params_per_problem, char* buf, int buf_len) {
  ACRO_REQUIRE(opts && buf && buf_len > 0 && B > 0, "acro_newton_describe: bad argument");
  NewtonPlan plan;
  AcroNewtonOpts o = *opts;
  if (!o.spec_ws && o.kernel == ACRO_NEWTON_SPEC) o.spec_ws = reinterpret_cast<double*>(uintptr_t(128));  // (planning only)
  const int rc = newton_plan(o, B, ref_per_problem != 0, weights_per_problem != 0, params_per_problem != 0, false, true, plan);
  if (rc != ACRO_OK) return rc;
  const bool wpb = weights_per_problem != 0, rpb = ref_per_problem != 0;
  const char* tf[2] = {"false", "true"};
  if (plan.kernel == ACRO_NEWTON_THREAD)
    snprintf(buf, buf_len, "acro::k_newton<%s,%s,%s>", tf[wpb], tf[rpb], tf[params_per_problem != 0]);
  else if (plan.kernel == ACRO_NEWTON_DUO && params_per_problem)
    snprintf(buf, buf_len, "acro::k_newton_duo<true,%s,%d,true>", tf[rpb], plan.stage_steps);
  else if (plan.kernel == ACRO_NEWTON_RING && params_per_problem)
    snprintf(buf, buf_len, "acro::k_newton_ring<true,%s,%d,%s,true>", tf[rpb], plan.stage_steps, tf[plan.recompute_lin]);
  else if (plan.kernel == ACRO_NEWTON_DUO)
    snprintf(buf, buf_len, "acro::k_newton_duo<%s,%s,%d>", tf[wpb || rpb], tf[rpb], plan.stage_steps);
  else if (plan.kernel == ACRO_NEWTON_SPEC)
    snprintf(buf, buf_len, "acro::k_newton_spec<%s,%s,%d>", tf[wpb || rpb], tf[rpb], plan.stage_steps);
  else
    snprintf(buf, buf_len, "acro::k_newton_ring<%s,%s,%d,%s>", tf[wpb], tf[rpb], plan.stage_steps, tf[plan.recompute_lin]);
  return ACRO_OK;
}

int acro_stepsize_sweep(const AcroParams* p, const AcroWeights* w, int64_t P, int N, const double* X,
                        const double* U, const double* K, const double* S, const AcroRef* ref, int S_n,
                        const double* steps, double* cost, void* stream) {
  ACRO_REQUIRE(p && w && X && U && K && S && ref && ref->x && ref->u && steps && cost && P > 0 && N >= 2 && S_n > 0,
               "acro_stepsize_sweep: bad argument");
  const dim3 grid((unsigned)((P + 3) / 4), (unsigned)((S_n + 31) / 32));
  ACRO_REQUIRE(grid.y <= 65535, "acro_stepsize_sweep: too many step sizes");
  const KWeights kw = make_weights(*w);
  const Model m = make_model(*p);
#define EXPR(WPB, RPB) \
  k_sweep<WPB, RPB><<<grid, 128, 0, (cudaStream_t)stream>>>(m, kw, P, N, X, U, K, S, ref->x, ref->u, S_n, steps, cost)
  DISPATCH2(per_problem_weights(*w), ref->per_problem != 0, EXPR);
#undef EXPR
  ACRO_LAUNCH_CHECK("acro_stepsize_sweep");
  return ACRO_OK;
}

int acro_cost_derivatives(const AcroWeights* w, int64_t B, const double* x, const double* x_ref, const double* u,
                          const double* u_ref, int terminal, double* l, double* grad_x, double* grad_u,
                          void* stream) {
  ACRO_REQUIRE(w && x && x_ref && l && grad_x && B > 0, "acro_cost_derivatives: bad argument");
  ACRO_REQUIRE(terminal || (u && u_ref && grad_u), "acro_cost_derivatives: stage cost needs u, u_ref, grad_u");
  const Cfg c = cfg_for(B);
  const KWeights kw = make_weights(*w);
  if (per_problem_weights(*w))
    k_cost_derivatives<true><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(kw, B, x, x_ref, u, u_ref, terminal, l, grad_x, grad_u);
  else
    k_cost_derivatives<false><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(kw, B, x, x_ref, u, u_ref, terminal, l, grad_x, grad_u);
  ACRO_LAUNCH_CHECK("acro_cost_derivatives");
  return ACRO_OK;
}

int acro_discretize(int64_t B, const double* Ac, const double* Bc, double dt, double* Ad, double* Bd, void* stream) {
  ACRO_REQUIRE(Ac && Bc && Ad && Bd && B > 0, "acro_discretize: bad argument");
  const Cfg c = cfg_for(B);
  k_discretize<<<c.grid, c.block, 0, (cudaStream_t)stream>>>(B, Ac, Bc, dt, Ad, Bd);
  ACRO_LAUNCH_CHECK("acro_discretize");
  return ACRO_OK;
}

int acro_stage_lists(const AcroParams* p, const AcroWeights* w, int64_t B, int N, const double* X, const double* U,
                     const AcroRef* ref, double* A, double* Bm, double* q, double* r, double* q_T, void* stream) {
  ACRO_REQUIRE(p && w && X && U && ref && ref->x && ref->u && A && Bm && q && r && q_T && B > 0 && N >= 2,
               "acro_stage_lists: bad argument");
  const int64_t n = int64_t(N) * B;
  const KWeights kw = make_weights(*w);
  const Model m = make_model(*p);
#define EXPR(WPB, RPB)                                                                                         \
  k_stage_lists<WPB, RPB><<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(m, kw, B, N, X, U, ref->x, \
                                                                                         ref->u, A, Bm, q, r, q_T)
  DISPATCH2(per_problem_weights(*w), ref->per_problem != 0, EXPR);
#undef EXPR
  ACRO_LAUNCH_CHECK("acro_stage_lists");
  return ACRO_OK;
}

int acro_riccati_lists(int64_t B, int T, const double* A, const double* Bm, const double* Q, const double* R,
                       const double* S_cross, const double* q, const double* r, const double* Q_T,
                       const double* q_T, double* K, double* S, double* delta_J, void* stream) {
  ACRO_REQUIRE(A && Bm && Q && R && q && r && Q_T && q_T && K && S && delta_J && B > 0 && T >= 1,
               "acro_riccati_lists: bad argument");
  const Cfg c = cfg_for(B);
  k_riccati_lists<<<c.grid, c.block, 0, (cudaStream_t)stream>>>(B, T, A, Bm, Q, R, S_cross, q, r, Q_T, q_T, K, S, delta_J);
  ACRO_LAUNCH_CHECK("acro_riccati_lists");
  return ACRO_OK;
}

int acro_lqr_gains(const AcroParams* p, const AcroWeights* w, int64_t B, int N, const AcroRef* traj, double* K,
                   void* stream) {
  ACRO_REQUIRE(p && w && traj && traj->x && traj->u && K && B > 0 && N >= 2, "acro_lqr_gains: bad argument");
  ACRO_REQUIRE(traj->per_problem || B == 1, "acro_lqr_gains: a shared trajectory is one problem (B = 1)");
  const Cfg c = cfg_for(B);
  const KWeights kw = make_weights(*w);
  const Model m = make_model(*p);
  if (p->actuated_tau1) {
#define EXPR(WPB, RPB) \
  k_lqr_gains<WPB, RPB, true><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(m, kw, B, N, traj->x, traj->u, K)
    DISPATCH2(per_problem_weights(*w), traj->per_problem != 0, EXPR);
#undef EXPR
  } else {
#define EXPR(WPB, RPB) \
  k_lqr_gains<WPB, RPB><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(m, kw, B, N, traj->x, traj->u, K)
    DISPATCH2(per_problem_weights(*w), traj->per_problem != 0, EXPR);
#undef EXPR
  }
  ACRO_LAUNCH_CHECK("acro_lqr_gains");
  return ACRO_OK;
}

int acro_lqr_track_pp(const AcroParams* p, const double* params_b, int64_t B, int N, const AcroRef* traj,
                      const double* K, const double* x0, double* Xt, double* Ut, void* stream) {
  ACRO_REQUIRE(p && traj && traj->x && traj->u && K && x0 && Xt && Ut && B > 0 && N >= 2, "acro_lqr_track: bad argument");
  const Cfg c = cfg_for(B);
  const Model m = make_model(*p);
#define EXPR(RPB, PPB) \
  k_lqr_track<RPB, PPB><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(m, params_b, B, N, traj->x, traj->u, K, x0, Xt, Ut)
  DISPATCH2(traj->per_problem != 0, params_b != nullptr, EXPR);
#undef EXPR
  ACRO_LAUNCH_CHECK("acro_lqr_track");
  return ACRO_OK;
}
int acro_lqr_track(const AcroParams* p, int64_t B, int N, const AcroRef* traj, const double* K, const double* x0,
                   double* Xt, double* Ut, void* stream) {
  return acro_lqr_track_pp(p, nullptr, B, N, traj, K, x0, Xt, Ut, stream);
}

int acro_p_inf(const AcroWeights* w, int64_t B, const double* A, const double* Bm, int max_iter, double tol,
               double* P, int32_t* n_iter, void* stream) {
  ACRO_REQUIRE(w && A && Bm && P && n_iter && B > 0 && max_iter > 0, "acro_p_inf: bad argument");
  const Cfg c = cfg_for(B);
  const KWeights kw = make_weights(*w);
  if (per_problem_weights(*w))
    k_p_inf<true><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(kw, B, A, Bm, max_iter, tol, P, n_iter);
  else
    k_p_inf<false><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(kw, B, A, Bm, max_iter, tol, P, n_iter);
  ACRO_LAUNCH_CHECK("acro_p_inf");
  return ACRO_OK;
}

int acro_mpc_solve(const AcroWeights* w, int64_t B, int T_pred, const double* x0, const double* A_w,
                   const double* B_w, const double* QT, double* U0, double* X_opt, double* U_opt, double* K_ws,
                   void* stream) {
  ACRO_REQUIRE(w && x0 && A_w && B_w && QT && U0 && B > 0 && T_pred >= 1, "acro_mpc_solve: bad argument");
  ACRO_REQUIRE((X_opt == nullptr) == (U_opt == nullptr), "acro_mpc_solve: X_opt and U_opt go together");
  ACRO_REQUIRE(!X_opt || K_ws, "acro_mpc_solve: K_ws workspace required for X_opt/U_opt");
  const Cfg c = cfg_for(B);
  const KWeights kw = make_weights(*w);
  if (per_problem_weights(*w))
    k_mpc_solve<true><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(kw, B, T_pred, x0, A_w, B_w, QT, U0, X_opt, U_opt, K_ws);
  else
    k_mpc_solve<false><<<c.grid, c.block, 0, (cudaStream_t)stream>>>(kw, B, T_pred, x0, A_w, B_w, QT, U0, X_opt, U_opt, K_ws);
  ACRO_LAUNCH_CHECK("acro_mpc_solve");
  return ACRO_OK;
}

}  // extern "C"
static int mpc_track_impl(const AcroParams* p, const double* params_b, const AcroWeights* w, int64_t B, int N, int T,
                          int T_pred, const AcroRef* ref, const double* x_f, const double* u_f, const double* QT_inf,
                          int qt_per_problem, const double* x0, double* K0, double* lin_ws, double* Xr, double* Ur,
                          int64_t* n_solves, void* stream) {
  ACRO_REQUIRE(p && w && ref && ref->x && ref->u && x_f && u_f && QT_inf && x0 && lin_ws && Xr && Ur && B > 0 &&
                   N >= 2 && T >= 2 && T <= N && T_pred >= 2,
               "acro_mpc_track: bad argument");
  ACRO_REQUIRE(!p->actuated_tau1, "acro_mpc_track: fully-actuated plant not supported here");
  ACRO_REQUIRE(!params_b || ref->per_problem, "acro_mpc_track_pp: per-problem parameters need a per-problem reference layout");
  const bool wpb = per_problem_weights(*w);
  const bool pp = ref->per_problem || wpb || qt_per_problem;
  ACRO_REQUIRE(pp || K0, "acro_mpc_track: K0 workspace required for a shared reference");
  ACRO_REQUIRE(!pp || ref->per_problem, "acro_mpc_track: per-problem weights need a per-problem reference layout");
  MpcArgs a;
  a.m = make_model(*p);
  a.kw = make_weights(*w);
  a.B = B;
  a.N = N;
  a.T = T;
  a.H = T_pred;
  a.rx = ref->x;
  a.ru = ref->u;
  for (int i = 0; i < 4; ++i) a.xf[i] = x_f[i];
  for (int i = 0; i < 2; ++i) a.uf[i] = u_f[i];
  a.QT = QT_inf;
  a.qt_per_problem = qt_per_problem;
  a.x0 = x0;
  a.lin = lin_ws;
  a.K0 = K0;
  a.Xr = Xr;
  a.Ur = Ur;
  a.pb = params_b;
  cudaStream_t s = (cudaStream_t)stream;
  if (!pp) {
    const int nt = N - 1;
    k_lin_compact<false><<<(nt + 63) / 64, 64, 0, s>>>(a.m, B, N, ref->x, ref->u, lin_ws);
    ACRO_LAUNCH_CHECK("acro_mpc_track/linearize");
    k_mpc_gains_shared<false><<<(T - 1 + 31) / 32, 32, 0, s>>>(a);
    ACRO_LAUNCH_CHECK("acro_mpc_track/gains");
    const Cfg c = cfg_for(B);
    k_mpc_track_shared<<<c.grid, c.block, 0, s>>>(a);
    ACRO_LAUNCH_CHECK("acro_mpc_track/track");
    if (n_solves) *n_solves = T - 1;
  } else {
    const int64_t n = int64_t(N - 1) * B;
    k_lin_compact<true><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a.m, B, N, ref->x, ref->u, lin_ws, params_b);
    ACRO_LAUNCH_CHECK("acro_mpc_track/linearize");
    const Cfg c = cfg_for(B);
    const size_t ring_bytes = size_t(c.block / 32) * ACRO_MPC_RING * 320 * sizeof(double);
#define LAUNCH_MPC_PP(WPB_, PPB_)                                                                                    \
  do {                                                                                                               \
    cudaError_t e0 = cudaFuncSetAttribute(k_mpc_track_pp<WPB_, PPB_>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                          (int)ring_bytes);                                                          \
    ACRO_REQUIRE(e0 == cudaSuccess, "acro_mpc_track: cudaFuncSetAttribute failed");                                  \
    k_mpc_track_pp<WPB_, PPB_><<<c.grid, c.block, ring_bytes, s>>>(a);                                               \
  } while (0)
    if (params_b) {
      if (wpb) LAUNCH_MPC_PP(true, true); else LAUNCH_MPC_PP(false, true);
    } else {
      if (wpb) LAUNCH_MPC_PP(true, false); else LAUNCH_MPC_PP(false, false);
    }
#undef LAUNCH_MPC_PP
    ACRO_LAUNCH_CHECK("acro_mpc_track/track");
    if (n_solves) *n_solves = int64_t(T - 1) * B;
  }
  return ACRO_OK;
}

extern "C" {
int acro_mpc_track(const AcroParams* p, const AcroWeights* w, int64_t B, int N, int T, int T_pred,
                   const AcroRef* ref, const double* x_f, const double* u_f, const double* QT_inf,
                   int qt_per_problem, const double* x0, double* K0, double* lin_ws, double* Xr, double* Ur,
                   int64_t* n_solves, void* stream) {
  return mpc_track_impl(p, nullptr, w, B, N, T, T_pred, ref, x_f, u_f, QT_inf, qt_per_problem, x0, K0, lin_ws, Xr, Ur,
                        n_solves, stream);
}
int acro_mpc_track_pp(const AcroParams* p, const double* params_b, const AcroWeights* w, int64_t B, int N, int T,
                      int T_pred, const AcroRef* ref, const double* x_f, const double* u_f, const double* QT_inf,
                      int qt_per_problem, const double* x0, double* lin_ws, double* Xr, double* Ur, int64_t* n_solves,
                      void* stream) {
  return mpc_track_impl(p, params_b, w, B, N, T, T_pred, ref, x_f, u_f, QT_inf, qt_per_problem, x0, nullptr, lin_ws, Xr,
                        Ur, n_solves, stream);
}

int64_t acro_mpc_box_ws_doubles(int64_t B, int T, int T_pred) {
  return mpc_box_ws_per_problem(T_pred) * B + mpc_box_table_doubles(T, T_pred);
}

}  // extern "C"
static int mpc_track_box_impl(const AcroParams* p, const double* params_b, const AcroWeights* w, int64_t B, int N, int T,
                              int T_pred, const AcroRef* ref, const double* x_f, const double* u_f, const double* QT_inf,
                              int qt_per_problem, const double* x0, double tau_max, int max_iter, double* lin_ws,
                              double* ws, double* Xr, double* Ur, int32_t* n_sweeps, int32_t* n_active, int32_t* status,
                              void* stream) {
  ACRO_REQUIRE(p && w && ref && ref->x && ref->u && x_f && u_f && QT_inf && x0 && lin_ws && ws && Xr && Ur && B > 0 &&
                   N >= 2 && T >= 2 && T <= N && T_pred >= 2 && tau_max > 0.0,
               "acro_mpc_track_box: bad argument");
  ACRO_REQUIRE(!params_b || ref->per_problem, "acro_mpc_track_box_pp: per-problem parameters need a per-problem reference layout");
  ACRO_REQUIRE(!p->actuated_tau1, "acro_mpc_track_box: fully-actuated plant not supported here");
  ACRO_REQUIRE(!per_problem_weights(*w), "acro_mpc_track_box: per-problem weights not supported here");
  ACRO_REQUIRE(w->R[1] == 0.0 && w->R[2] == 0.0, "acro_mpc_track_box: R must be diagonal");
  MpcBoxArgs a;
  a.m = make_model(*p);
  a.kw = make_weights(*w);
  a.B = B;
  a.N = N;
  a.T = T;
  a.H = T_pred;
  a.rx = ref->x;
  a.ru = ref->u;
  for (int i = 0; i < 4; ++i) a.xf[i] = x_f[i];
  for (int i = 0; i < 2; ++i) a.uf[i] = u_f[i];
  a.QT = QT_inf;
  a.qt_per_problem = qt_per_problem;
  a.x0 = x0;
  a.lin = lin_ws;
  a.ws = ws;
  a.ktab = nullptr;
  a.wtab = nullptr;
  a.pb = params_b;
  a.tau = tau_max;
  a.max_iter = max_iter > 0 ? max_iter : 6 * (T_pred - 1) + 20;
  a.Xr = Xr;
  a.Ur = Ur;
  a.n_sweeps = n_sweeps;
  a.n_active = n_active;
  a.status = status;
  cudaStream_t s = (cudaStream_t)stream;
  const Cfg c = cfg_for(B);
  if (ref->per_problem) {
    const int64_t n = int64_t(N - 1) * B;
    k_lin_compact<true><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(a.m, B, N, ref->x, ref->u, lin_ws, params_b);
    ACRO_LAUNCH_CHECK("acro_mpc_track_box/linearize");
    const size_t ring_bytes = size_t(c.block / 32) * ACRO_BOX_RING * BoxSlot<true>::N * 32 * sizeof(double);
    if (params_b) {
      ACRO_REQUIRE(cudaFuncSetAttribute(k_mpc_track_box<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)ring_bytes) == cudaSuccess, "acro_mpc_track_box: cudaFuncSetAttribute failed");
      k_mpc_track_box<true, true><<<c.grid, c.block, ring_bytes, s>>>(a);
    } else {
      ACRO_REQUIRE(cudaFuncSetAttribute(k_mpc_track_box<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes) ==
                       cudaSuccess, "acro_mpc_track_box: cudaFuncSetAttribute failed");
      k_mpc_track_box<true><<<c.grid, c.block, ring_bytes, s>>>(a);
    }
  } else {
    k_lin_compact<false><<<(N - 1 + 63) / 64, 64, 0, s>>>(a.m, B, N, ref->x, ref->u, lin_ws);
    ACRO_LAUNCH_CHECK("acro_mpc_track_box/linearize");
    if (!qt_per_problem) {  // gains of the empty working set, shared by the batch: one sweep per MPC step
      a.ktab = ws + mpc_box_ws_per_problem(T_pred) * B;
      a.wtab = a.ktab + mpc_box_ktab_doubles(T, T_pred);
      k_mpc_box_gains<<<(T - 1 + 31) / 32, 32, 0, s>>>(a);
      ACRO_LAUNCH_CHECK("acro_mpc_track_box/gains");
    }
    // per warp: the operand ring and, when the tables are used, their staged copies (gains 2 x 4n, window rows 11 (n+1))
    const int nb = T_pred - 1;
    size_t ring_bytes = size_t(c.block / 32) * (ACRO_BOX_RING * BoxSlot<false>::N * 32 + (a.ktab ? 8 * nb + 11 * (nb + 1) : 0)) *
                        sizeof(double);
    if (ring_bytes > 200 * 1024) {  // a horizon too long to stage: every solve runs its own backward sweep, rows from global memory
      a.ktab = nullptr;
      ring_bytes = size_t(c.block / 32) * ACRO_BOX_RING * BoxSlot<false>::N * 32 * sizeof(double);
    }
    ACRO_REQUIRE(cudaFuncSetAttribute(k_mpc_track_box<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes) ==
                     cudaSuccess, "acro_mpc_track_box: cudaFuncSetAttribute failed");
    k_mpc_track_box<false><<<c.grid, c.block, ring_bytes, s>>>(a);
  }
  ACRO_LAUNCH_CHECK("acro_mpc_track_box");
  return ACRO_OK;
}

extern "C" {
int acro_mpc_track_box(const AcroParams* p, const AcroWeights* w, int64_t B, int N, int T, int T_pred,
                       const AcroRef* ref, const double* x_f, const double* u_f, const double* QT_inf,
                       int qt_per_problem, const double* x0, double tau_max, int max_iter, double* lin_ws, double* ws,
                       double* Xr, double* Ur, int32_t* n_sweeps, int32_t* n_active, int32_t* status, void* stream) {
  return mpc_track_box_impl(p, nullptr, w, B, N, T, T_pred, ref, x_f, u_f, QT_inf, qt_per_problem, x0, tau_max, max_iter,
                            lin_ws, ws, Xr, Ur, n_sweeps, n_active, status, stream);
}
int acro_mpc_track_box_pp(const AcroParams* p, const double* params_b, const AcroWeights* w, int64_t B, int N, int T,
                          int T_pred, const AcroRef* ref, const double* x_f, const double* u_f, const double* QT_inf,
                          int qt_per_problem, const double* x0, double tau_max, int max_iter, double* lin_ws, double* ws,
                          double* Xr, double* Ur, int32_t* n_sweeps, int32_t* n_active, int32_t* status, void* stream) {
  return mpc_track_box_impl(p, params_b, w, B, N, T, T_pred, ref, x_f, u_f, QT_inf, qt_per_problem, x0, tau_max, max_iter,
                            lin_ws, ws, Xr, Ur, n_sweeps, n_active, status, stream);
}

int acro_bench_fp64_peak(int blocks, int threads, int iters, double* out, void* stream) {
  ACRO_REQUIRE(out && blocks > 0 && threads > 0 && threads <= 1024 && iters > 0, "acro_bench_fp64_peak: bad argument");
  k_fp64_peak<<<blocks, threads, 0, (cudaStream_t)stream>>>(out, iters, 0.999999, 1e-7);
  ACRO_LAUNCH_CHECK("acro_bench_fp64_peak");
  return ACRO_OK;
}

int acro_bench_fp64_chain(int blocks, int threads, int iters, int chains, int active_lanes, double* out,
                          long long* cycles, void* stream) {
  ACRO_REQUIRE(out && cycles && blocks > 0 && threads > 0 && threads <= 1024 && iters > 0, "acro_bench_fp64_chain: bad argument");
  k_fp64_chain<<<blocks, threads, 0, (cudaStream_t)stream>>>(out, iters, 0.999999, 1e-7, cycles, chains, active_lanes);
  ACRO_LAUNCH_CHECK("acro_bench_fp64_chain");
  return ACRO_OK;
}

int acro_transpose(int64_t rows, int64_t cols, const double* src, double* dst, void* stream) {
  ACRO_REQUIRE(src && dst && rows > 0 && cols > 0, "acro_transpose: bad argument");
  const int64_t tiles = ((cols + 31) / 32) * ((rows + 31) / 32);
  ACRO_REQUIRE(tiles < (1LL << 31), "acro_transpose: array too large for one call");
  k_transpose<<<(unsigned)tiles, dim3(32, 8), 0, (cudaStream_t)stream>>>(rows, cols, src, dst);
  ACRO_LAUNCH_CHECK("acro_transpose");
  return ACRO_OK;
}

static int tiled(bool pack, int64_t B, int T, int C, const double* src, double* dst, void* stream, const char* name) {
  ACRO_REQUIRE(src && dst && B > 0 && T > 0 && C > 0, "acro_pack_soa/acro_unpack_soa: bad argument");
  const int64_t blocks = ((int64_t(T) * C + 31) / 32) * ((B + 31) / 32);
  ACRO_REQUIRE(blocks < (1LL << 31), "acro_pack_soa/acro_unpack_soa: array too large for one call");
  if (pack)
    k_tiled<true><<<(unsigned)blocks, dim3(32, 8), 0, (cudaStream_t)stream>>>(B, T, C, src, dst);
  else
    k_tiled<false><<<(unsigned)blocks, dim3(32, 8), 0, (cudaStream_t)stream>>>(B, T, C, src, dst);
  ACRO_LAUNCH_CHECK(name);
  return ACRO_OK;
}

int acro_pack_soa(int64_t B, int T, int C, const double* src, double* dst, void* stream) {
  return tiled(true, B, T, C, src, dst, stream, "acro_pack_soa");
}
int acro_unpack_soa(int64_t B, int T, int C, const double* src, double* dst, void* stream) {
  return tiled(false, B, T, C, src, dst, stream, "acro_unpack_soa");
}

}  // extern "C"
