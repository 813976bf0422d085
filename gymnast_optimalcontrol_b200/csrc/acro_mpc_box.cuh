// acro_mpc_box.cuh - receding-horizon MPC with the input box the reference keeps behind `test_constraints`
// (trajectory_tracking.py:87-91, 102-104, 112-114):   -tau_max <= U[:, j] + u_ref[j] <= tau_max,  j < T_pred - 1.
//
// With the box the QP of solver_mpc (tt:80-117) is no longer a plain LQ problem.  The first input does not act on the
// plant (B[:, 0] = 0, dynamics.py:153) and R is diagonal, so it separates: u0_j = the point of its interval closest
// to 0.  What remains is a box-constrained LQ problem in the scalar input v_j = U[1, j], solved EXACTLY by a primal
// active-set method whose linear algebra is Riccati sweeps (one thread per problem, everything in FP64):
//
//   working set W (inputs held at a bound), feasible iterate v, its state trajectory xb
//   repeat
//     backward sweep j = n-1 .. 0 : value function x'Px + 2p'x with the inputs in W fixed and the others free
//                                   (free: K_j = -F/G, k_j = -g/G;  fixed: P <- Q + A'PA, p <- A'(p + P b v_j))
//     forward sweep (closed loop) : the minimiser v* over the free inputs and its states xb; largest step alpha <= 1
//                                   from v towards v* that stays inside the box, and the input that blocks it
//     blocked   -> v += alpha (v* - v), the blocking input joins W
//     unblocked -> v = v*; costate sweep of (xb, v*) and with it the multipliers dJ/dv_j of the inputs in W:
//                  drop the input whose multiplier has the wrong sign (the most violating), or stop
//
// Finite termination, no tolerance on the answer other than rounding; consecutive MPC steps warm-start from the
// shifted solution (typically one or two sweeps per step).  Per-problem scratch lives in a caller-supplied workspace
// ws[e][B] (e fastest across problems: coalesced).  The tests compare it with the same QP condensed and solved with
// dense matrices; the reference itself would hand it to IPOPT (tol 1e-6): parity against IPOPT is unpinned.
//
// How a solve spends its time (profiles/r2_box_probe.txt): one thread per problem, a few warps per SM, so every loop
// over the window that waits for its own loads pays a full memory latency per step (2770 cycles per window step in
// round 1, of which ~400 are arithmetic).  Hence
//   * the general sweeps get their per-problem operands through a cp.async ring in shared memory (below);
//   * the window slides by renaming: v and the working-set flags live in circular slots (window index j of step t is
//     slot (t + j) mod n), the candidate v* is written to a second buffer and a full step swaps the two pointers,
//     the flags are bytes, the number of held inputs is carried along - no copy loops;
//   * with a shared reference and an empty working set the gains of the backward sweep are the same for every
//     problem: k_mpc_box_gains computes them once per MPC step (table [T-1][n][4]) and a solve whose working set
//     is empty runs only the forward sweep (the feasibility check of the unconstrained minimiser) against the table;
//     table and window rows of the step are staged in shared memory per warp, the next step's prefetched.
#pragma once
#include "acro_device.cuh"
#include "acro_views.cuh"

namespace acro {

struct MpcBoxArgs {
  Model m;
  KWeights kw;
  int64_t B;
  int N, T, H;
  const double *rx, *ru;
  double xf[4], uf[2];
  const double* QT;  // [16] shared or [16][B]
  int qt_per_problem;
  const double* x0;
  const double* lin;  // compact linearisation: shared [N-1][10] or per problem {N-1 x 10}
  double* ws;         // [12 (H-1) + 4][B] per-problem scratch, then the gain table
  double* ktab;       // [(T-1)][H-1][4]: gains of the empty working set (shared reference, shared Q_T), else null
  double* wtab;       // [T + H - 3][11]: linearisation and second reference input of every absolute time a window
                      // reaches, the (x_f, u_f) padding of tt:64-67 filled in (shared reference), else null
  double tau;
  int max_iter;       // active-set iterations per solve
  double *Xr, *Ur;
  int32_t* n_sweeps;  // [B] total active-set iterations of the problem (out, may be null)
  int32_t* n_active;  // [T-1][B] inputs at a bound in the solution of step t (out, may be null)
  int32_t* status;    // [B] 0, or 1 if some step ran into max_iter (out, may be null)
  const double* pb;   // physical parameters per problem [11][B] (k_mpc_track_box<true, true>) or null
};

// doubles of workspace per problem: v, v* (2n), gains (5n), states (4n + 4), flags (n int32, n doubles reserved)
__host__ __device__ inline int64_t mpc_box_ws_per_problem(int H) { return 12LL * (H - 1) + 4; }
__host__ __device__ inline int64_t mpc_box_ktab_doubles(int T, int H) { return 4LL * (T - 1) * (H - 1); }
__host__ __device__ inline int64_t mpc_box_table_doubles(int T, int H) {
  return mpc_box_ktab_doubles(T, H) + 11LL * (T + H - 3);
}
// row tj of the window table
__device__ __forceinline__ LinD box_load_row(const double* __restrict__ wtab, int tj) {
  const double* r = wtab + int64_t(tj) * 11;
  LinD L;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    L.a[0][j] = __ldg(r + j);
    L.a[1][j] = __ldg(r + 4 + j);
  }
  L.b[0] = __ldg(r + 8);
  L.b[1] = __ldg(r + 9);
  L.b0[0] = L.b0[1] = 0.0;
  return L;
}

// x+ = A_d x + b v on the structured linearisation
__device__ __forceinline__ void box_plant(const LinD& L, double dt, const double x[4], double v, double xn[4]) {
  xn[0] = fma(dt, x[2], x[0]);
  xn[1] = fma(dt, x[3], x[1]);
  xn[2] = fma(L.b[0], v, fma(L.a[0][3], x[3], fma(L.a[0][2], x[2], fma(L.a[0][1], x[1], L.a[0][0] * x[0]))));
  xn[3] = fma(L.b[1], v, fma(L.a[1][3], x[3], fma(L.a[1][2], x[2], fma(L.a[1][1], x[1], L.a[1][0] * x[0]))));
}
// A_d' y
__device__ __forceinline__ void box_At(const LinD& L, double dt, const double y[4], double o[4]) {
  o[0] = fma(L.a[1][0], y[3], fma(L.a[0][0], y[2], y[0]));
  o[1] = fma(L.a[1][1], y[3], fma(L.a[0][1], y[2], y[1]));
  o[2] = fma(L.a[1][2], y[3], fma(L.a[0][2], y[2], dt * y[0]));
  o[3] = fma(L.a[1][3], y[3], fma(L.a[0][3], y[2], dt * y[1]));
}
// S = A_d' P A_d (upper triangle), F = b' P A_d, Pb = P b
__device__ __forceinline__ void box_products(const double P[10], const LinD& L, double dt, double S[10], double F[4],
                                             double Pb[4]) {
  double M[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double pi0 = P[sym(i, 0)], pi1 = P[sym(i, 1)], pi2 = P[sym(i, 2)], pi3 = P[sym(i, 3)];
    M[i][0] = fma(pi3, L.a[1][0], fma(pi2, L.a[0][0], pi0));
    M[i][1] = fma(pi3, L.a[1][1], fma(pi2, L.a[0][1], pi1));
    M[i][2] = fma(pi3, L.a[1][2], fma(pi2, L.a[0][2], dt * pi0));
    M[i][3] = fma(pi3, L.a[1][3], fma(pi2, L.a[0][3], dt * pi1));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int c = i; c < 4; ++c) {
      const double top = (i < 2) ? M[i][c] : dt * M[i - 2][c];
      S[sym(i, c)] = fma(L.a[1][i], M[3][c], fma(L.a[0][i], M[2][c], top));
    }
    F[i] = fma(L.b[1], M[3][i], L.b[0] * M[2][i]);
    Pb[i] = fma(P[sym(i, 3)], L.b[1], P[sym(i, 2)] * L.b[0]);
  }
}
// backward step over a FREE input: gain row Kr = -F/G, P <- Q + S + Kr F'; returns 1/G
template <class W>
__device__ __forceinline__ double box_free_step(double P[10], const LinD& L, const W& w, double R11, const double S[10],
                                                const double F[4], const double Pb[4], double Kr[4]) {
  const double G = R11 + fma(L.b[1], Pb[3], L.b[0] * Pb[2]);
  const double iG = 1.0 / G;
#pragma unroll
  for (int c = 0; c < 4; ++c) Kr[c] = -F[c] * iG;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = i; c < 4; ++c) P[sym(i, c)] = w.Q(i, c) + S[sym(i, c)] + Kr[i] * F[c];
  return iG;
}

// Gains of the EMPTY working set for every MPC step of a shared reference: thread t sweeps its window once.
__global__ void k_mpc_box_gains(const __grid_constant__ MpcBoxArgs a) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.T - 1) return;
  const WV<false> w(a.kw, 1, 0);
  const double dt = a.m.dt, R11 = w.R(1, 1);
  const int n = a.H - 1, n_lin = a.N - 1;
  const LinD Lf = linearize_d(a.m, a.xf, a.uf[0], a.uf[1]);
  double P[10];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) P[sym(i, j)] = a.QT[i * 4 + j];
  // window table: rows t (every thread its own) and, by the last thread, the padded tail
  for (int tj = t; tj < a.T + n - 2; tj += (t == a.T - 2 ? 1 : a.T)) {
    const LinD L = (tj < n_lin) ? load_lin(a.lin, tj, 0, 0) : Lf;
    double* const r = a.wtab + int64_t(tj) * 11;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      r[j] = L.a[0][j];
      r[4 + j] = L.a[1][j];
    }
    r[8] = L.b[0];
    r[9] = L.b[1];
    r[10] = (tj < a.N - 1) ? __ldg(a.ru + tj * 2 + 1) : a.uf[1];
  }
  double* const row = a.ktab + int64_t(t) * n * 4;
  for (int j = n - 1; j >= 0; --j) {
    const LinD L = (t + j < n_lin) ? load_lin(a.lin, t + j, 0, 0) : Lf;
    double S[10], F[4], Pb[4], Kr[4];
    box_products(P, L, dt, S, F, Pb);
    box_free_step(P, L, w, R11, S, F, Pb, Kr);
#pragma unroll
    for (int c = 0; c < 4; ++c) row[j * 4 + c] = Kr[c];
  }
}

// Per-problem operands of the general sweeps come through a ring of ACRO_BOX_RING window steps per warp in shared memory,
// filled with cp.async (every thread copies the values of its own problem: no synchronisation between lanes) and awaited
// with cp.async.wait_group, which counts groups - the wait for step j leaves the copies of the steps after it in flight.
// (Loads into rotating register records all landed on one scoreboard: every first use waited for the youngest load too,
// 41 % of the stall samples of the first round-2 version, profiles/r2_mpc_box_ncu_summary.txt.)
#define ACRO_BOX_RING 4
template <bool RPB>
struct BoxSlot {
  static constexpr int N = RPB ? 18 : 8;  // doubles per lane and step: v, flag, K[4] | xb[4], u_ref, k (+ lin[10])
};
__device__ __forceinline__ void box_cp8(double* smem_dst, const double* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void box_cp4(double* smem_dst, const int32_t* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void box_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void box_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// PPB (with RPB): every problem its own physical parameters - plant step and padding linearisation with its own model
template <bool RPB, bool PPB = false>
__global__ void k_mpc_track_box(const __grid_constant__ MpcBoxArgs a) {
  const int64_t B = a.B, b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  const WV<false> w(a.kw, B, b);
  const RefV<RPB> ref{a.rx, a.ru, a.N, b};
  Model m_loc_;
  if (PPB) m_loc_ = model_per_problem(a.m, a.pb, B, b);
  const Model& m = PPB ? m_loc_ : a.m;
  const double dt = m.dt, R11 = w.R(1, 1), tau = a.tau;
  const int n = a.H - 1, n_lin = a.N - 1;
  const int64_t ld = RPB ? int64_t(a.N - 1) : 0;
  const bool tab = !RPB && a.ktab != nullptr && !a.qt_per_problem;
  const LinD Lf = linearize_d(m, a.xf, a.uf[0], a.uf[1]);
  double QT[10];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) QT[sym(i, j)] = a.qt_per_problem ? a.QT[(i * 4 + j) * B + b] : a.QT[i * 4 + j];
  // workspace views (element e of this problem at ws[e * B + b]); v / v* / flags in circular slots
  double* const ws = a.ws + b;
  double* vcur = ws;                   // the feasible iterate v
  double* voth = ws + int64_t(n) * B;  // the minimiser v* of the current working set
  int32_t* const wk = reinterpret_cast<int32_t*>(a.ws + (11LL * n + 4) * B) + b;  // +1 / -1: held at the upper / lower bound
  auto Kg = [&](int j, int c) -> double& { return ws[int64_t(2 * n + 4 * j + c) * B]; };
  auto kg = [&](int j) -> double& { return ws[int64_t(6 * n + j) * B]; };
  auto Xb = [&](int j, int c) -> double& { return ws[int64_t(7 * n + 4 * j + c) * B]; };
  // Shared reference: the tables of the current MPC step live in shared memory, one copy per warp - the gains of the
  // empty working set (n x 4, double-buffered: the next step's are fetched while this one is solved) and the window rows
  // (circular, n + 1 rows of 11: the window slides by one row per step).  Filled by the lanes of the warp together with
  // cp.async at the top of every step, where the lanes meet (__syncwarp); read-only in between.  (Reading them from
  // global memory at their use - broadcast loads that hit in L1 - left ~35 % of the stall samples on their first use.)
  extern __shared__ double box_ring[];
  constexpr int SL = BoxSlot<RPB>::N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int na = int(min(int64_t(32), B - (b - lane)));  // lanes of this warp that own a problem (the others have left)
  const unsigned lanes = na == 32 ? 0xffffffffu : ((1u << na) - 1u);
  double* const ksm = box_ring + nwarps * (ACRO_BOX_RING * SL * 32) + warp * (8 * n + 11 * (n + 1));
  double* const wsm = ksm + 8 * n;
  int t_now = 0, w0 = 0;  // the MPC step being solved; slot of its first window row (= t mod (n + 1))
  auto wrow = [&](int tj) -> const double* {
    int i = w0 + (tj - t_now);
    if (i >= n + 1) i -= n + 1;
    return wsm + i * 11;
  };
  auto lin_at = [&](int tj) {
    if (tab) {
      const double* r = wrow(tj);
      LinD L;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        L.a[0][c] = r[c];
        L.a[1][c] = r[4 + c];
      }
      L.b[0] = r[8];
      L.b[1] = r[9];
      L.b0[0] = L.b0[1] = 0.0;
      return L;
    }
    if (!RPB && a.wtab) return box_load_row(a.wtab, tj);
    return tj < n_lin ? load_lin(a.lin, tj, ld, b) : Lf;
  };
  auto uref1 = [&](int tj) {
    if (tab) return wrow(tj)[10];
    if (!RPB && a.wtab) return __ldg(a.wtab + int64_t(tj) * 11 + 10);
    return tj < a.N - 1 ? ref.U(tj, 1) : a.uf[1];
  };
  int s0 = 0;  // slot of window index 0 (= t mod n)
  auto slot = [&](int j) -> int64_t {
    const int s = s0 + j;
    return int64_t(s >= n ? s - n : s) * B;
  };

  double* const ring = box_ring + warp * (ACRO_BOX_RING * SL * 32) + lane;
  auto rs = [&](int s, int e) -> double& { return ring[(s * SL + e) * 32]; };
  // window step j of absolute time tj -> ring slot s: what every general sweep needs (v, flag) and, with per-problem
  // references, the linearisation and the reference input of the step
  // (the rows of a SHARED reference stay broadcast loads at their use: copying them into the ring per lane - 11 more
  // cp.async per lane and step, 32 times redundant - made the kernel 45 % slower)
  constexpr bool rows_in_ring = RPB;
  auto fetch_common = [&](int j, int tj, int s, const double* vbuf) {
    box_cp8(&rs(s, 0), vbuf + slot(j));
    box_cp4(&rs(s, 1), wk + slot(j));
    if (RPB) {
      if (tj < n_lin) {
#pragma unroll
        for (int c = 0; c < 10; ++c) box_cp8(&rs(s, 8 + c), a.lin + soa(tj, 10, c, ld, b));
      }
      if (tj < a.N - 1) box_cp8(&rs(s, 6), a.ru + soa(tj, 2, 1, a.N - 1, b));
    }
  };
  auto ring_lin = [&](int tj, int s) {
    if (!rows_in_ring) return lin_at(tj);
    if (RPB && tj >= n_lin) return Lf;
    LinD L;
    constexpr int o = RPB ? 8 : 0;  // (never read without RPB)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      L.a[0][c] = rs(s, o + c);
      L.a[1][c] = rs(s, o + 4 + c);
    }
    L.b[0] = rs(s, RPB ? 16 : 0);
    L.b[1] = rs(s, RPB ? 17 : 0);
    L.b0[0] = L.b0[1] = 0.0;
    return L;
  };
  auto ring_uref = [&](int tj, int s) {
    if (!rows_in_ring) return uref1(tj);
    if (RPB && tj >= a.N - 1) return a.uf[1];
    return rs(s, 6);
  };
  auto ring_flag = [&](int s) { return *reinterpret_cast<const int32_t*>(&rs(s, 1)); };

  double x[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    x[c] = a.x0[c * B + b];
    a.Xr[soa(0, 4, c, a.T, b)] = x[c];
  }
  int sweeps = 0, stat = 0, held = 0;
  for (int t = 0; t < a.T - 1; ++t) {
    if (tab) {
      auto stage = [&](double* dst, const double* src, int count) {
        for (int i = lane; i < count; i += na) box_cp8(dst + i, src + i);
      };
      if (t == 0) {
        stage(ksm, a.ktab, 4 * n);
        stage(wsm, a.wtab, 11 * n);
        box_commit();
      } else {
        w0 = (w0 + 1 == n + 1) ? 0 : w0 + 1;
      }
      t_now = t;
      box_wait<0>();
      __syncwarp(lanes);  // every lane has finished step t-1 and the tables of step t have landed
      if (t + 1 < a.T - 1) {  // tables of step t+1: its gains into the other buffer, the row that enters its window
        stage(ksm + ((t + 1) & 1) * 4 * n, a.ktab + int64_t(t + 1) * n * 4, 4 * n);
        int i = w0 + n;
        if (i >= n + 1) i -= n + 1;
        stage(wsm + i * 11, a.wtab + int64_t(t + n) * 11, 11);
      }
      box_commit();
    }
    // ---- start: the previous solution shifted by one step (same absolute times, hence still feasible and with the
    // same working set), the new last input at the point of its interval closest to 0
    if (t == 0) {
      for (int j = 0; j < n; ++j) {
        const double ur = uref1(j), lo = -tau - ur, hi = tau - ur;
        const double v = fmin(fmax(0.0, lo), hi);
        const int f = (v >= hi) ? 1 : ((v <= lo) ? -1 : 0);
        vcur[slot(j)] = v;
        wk[slot(j)] = f;
        held += (f != 0);
      }
    } else {
      held -= (wk[slot(0)] != 0);  // the input that leaves the window
      s0 = (s0 + 1 == n) ? 0 : s0 + 1;
      const double ur = uref1(t + n - 1), lo = -tau - ur, hi = tau - ur;
      const double v = fmin(fmax(0.0, lo), hi);
      const int f = (v >= hi) ? 1 : ((v <= lo) ? -1 : 0);
      vcur[slot(n - 1)] = v;
      wk[slot(n - 1)] = f;
      held += (f != 0);
    }
    double x0w[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) x0w[c] = x[c] - ((t < a.N) ? ref.X(t, c) : a.xf[c]);
    int it = 0;
    for (; it < a.max_iter; ++it) {
      const bool from_table = tab && held == 0;  // empty working set, shared reference: the gains are in the table
      if (!from_table) {
        // ---- backward sweep: value function with the held inputs fixed, gains of the free ones
        double P[10], p[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int e = 0; e < 10; ++e) P[e] = QT[e];
        auto bw_step = [&](double rv, int rf, const LinD& Lr, int j) {
          const LinD& L = Lr;
          double S[10], F[4], Pb[4], pn[4];
          box_products(P, L, dt, S, F, Pb);
          if (rf != 0) {  // held at v_j (its gain row is never read)
            double y[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) y[c] = fma(Pb[c], rv, p[c]);
            box_At(L, dt, y, pn);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int c = i; c < 4; ++c) P[sym(i, c)] = w.Q(i, c) + S[sym(i, c)];
          } else {
            const double gg = fma(L.b[1], p[3], L.b[0] * p[2]);
            double Kr[4];
            const double iG = box_free_step(P, L, w, R11, S, F, Pb, Kr);
#pragma unroll
            for (int c = 0; c < 4; ++c) Kg(j, c) = Kr[c];
            kg(j) = -gg * iG;
            box_At(L, dt, p, pn);
#pragma unroll
            for (int c = 0; c < 4; ++c) pn[c] = fma(Kr[c], gg, pn[c]);
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) p[c] = pn[c];
        };
        auto fetch_bw = [&](int j, int s) {
          if (j >= 0) fetch_common(j, t + j, s, vcur);
          box_commit();
        };
#pragma unroll
        for (int i = 0; i < ACRO_BOX_RING; ++i) fetch_bw(n - 1 - i, i);
        int s = 0;
        for (int j = n - 1; j >= 0; --j) {
          box_wait<ACRO_BOX_RING - 1>();
          const double rv = rs(s, 0);
          const int rf = ring_flag(s);
          const LinD Lr = ring_lin(t + j, s);
          bw_step(rv, rf, Lr, j);
          fetch_bw(j - ACRO_BOX_RING, s);  // after the step: the slot has been consumed
          s = (s + 1 == ACRO_BOX_RING) ? 0 : s + 1;
        }
        box_wait<0>();
      }
      // ---- forward sweep (closed loop): minimiser over the free inputs, blocking input
      double alpha = 1.0;
      int jb = -1, sb = 0;
      // the candidate vs of input j: blocking ratio of the step from v towards it
      auto consider = [&](int j, double vj, double vs, double ur) {
        const double lo = -tau - ur, hi = tau - ur, d = vs - vj;
        if (vs > hi && d > 0.0) {
          const double q = (hi - vj) / d;
          if (q < alpha) {
            alpha = q;
            jb = j;
            sb = 1;
          }
        } else if (vs < lo && d < 0.0) {
          const double q = (lo - vj) / d;
          if (q < alpha) {
            alpha = q;
            jb = j;
            sb = -1;
          }
        }
      };
      if (from_table) {
        // nothing is held: gains and window rows from the staged tables, v six steps ahead
        double xs[4] = {x0w[0], x0w[1], x0w[2], x0w[3]};
        const double* const krow = ksm + (t & 1) * 4 * n;
        auto tab_step = [&](double vj, int j) {
          const double* const q = wrow(t + j);
          const double* const K = krow + j * 4;
          const double vs = fma(K[3], xs[3], fma(K[2], xs[2], fma(K[1], xs[1], fma(K[0], xs[0], 0.0))));
          consider(j, vj, vs, q[10]);
          voth[slot(j)] = vs;
          double xn[4];
          xn[0] = fma(dt, xs[2], xs[0]);
          xn[1] = fma(dt, xs[3], xs[1]);
          xn[2] = fma(q[8], vs, fma(q[3], xs[3], fma(q[2], xs[2], fma(q[1], xs[1], q[0] * xs[0]))));
          xn[3] = fma(q[9], vs, fma(q[7], xs[3], fma(q[6], xs[2], fma(q[5], xs[1], q[4] * xs[0]))));
#pragma unroll
          for (int c = 0; c < 4; ++c) xs[c] = xn[c];
        };
        auto v_at = [&](int j) { return vcur[slot(min(j, n - 1))]; };
        double v0 = v_at(0), v1 = v_at(1), v2 = v_at(2), v3 = v_at(3), v4 = v_at(4), v5 = v_at(5);
        for (int j = 0; j < n; j += 2) {
          const double va = v0, vb = v1;
          v0 = v2;
          v1 = v3;
          v2 = v4;
          v3 = v5;
          v4 = v_at(j + 6);
          v5 = v_at(j + 7);
          tab_step(va, j);
          if (j + 1 < n) tab_step(vb, j + 1);
        }
      } else {
        double xs[4] = {x0w[0], x0w[1], x0w[2], x0w[3]};
        const bool keep_states = held > 0;  // the costate sweep below reads them
        auto fetch_fw = [&](int j, int s) {
          if (j < n) {
            fetch_common(j, t + j, s, vcur);
#pragma unroll
            for (int c = 0; c < 4; ++c) box_cp8(&rs(s, 2 + c), &Kg(j, c));
            box_cp8(&rs(s, 7), &kg(j));
          }
          box_commit();
        };
#pragma unroll
        for (int i = 0; i < ACRO_BOX_RING; ++i) fetch_fw(i, i);
        int s = 0;
        for (int j = 0; j < n; ++j) {
          box_wait<ACRO_BOX_RING - 1>();
          const double vj = rs(s, 0);
          const int rf = ring_flag(s);
          double vs = vj;
          if (keep_states) {
#pragma unroll
            for (int c = 0; c < 4; ++c) Xb(j, c) = xs[c];
          }
          if (rf == 0) {
            vs = fma(rs(s, 5), xs[3], fma(rs(s, 4), xs[2], fma(rs(s, 3), xs[1], fma(rs(s, 2), xs[0], rs(s, 7)))));
            consider(j, vj, vs, ring_uref(t + j, s));
          }
          voth[slot(j)] = vs;
          const LinD L = ring_lin(t + j, s);
          double xn[4];
          box_plant(L, dt, xs, vs, xn);
#pragma unroll
          for (int c = 0; c < 4; ++c) xs[c] = xn[c];
          fetch_fw(j + ACRO_BOX_RING, s);  // after the step: the slot has been consumed
          s = (s + 1 == ACRO_BOX_RING) ? 0 : s + 1;
        }
        box_wait<0>();
        if (keep_states) {
#pragma unroll
          for (int c = 0; c < 4; ++c) Xb(n, c) = xs[c];
        }
      }
      if (jb >= 0) {  // blocked: partial step, the blocking input joins the working set
        for (int j0 = 0; j0 < n; j0 += 4) {  // four loads in flight before the first store
          double vo[4], vn[4], ur[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int j = min(j0 + k, n - 1);
            vo[k] = vcur[slot(j)];
            vn[k] = voth[slot(j)];
            ur[k] = uref1(t + j);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k)  // clamped: rounding must not carry an input a unit in the last place outside
            if (j0 + k < n) vcur[slot(j0 + k)] = fmin(fmax(fma(alpha, vn[k] - vo[k], vo[k]), -tau - ur[k]), tau - ur[k]);
        }
        const double ur = uref1(t + jb);
        vcur[slot(jb)] = (sb > 0) ? tau - ur : -tau - ur;
        wk[slot(jb)] = sb;
        ++held;
        continue;
      }
      // not blocked: v* is feasible - it becomes the iterate (v <-> v*)
      {
        double* const tmp = vcur;
        vcur = voth;
        voth = tmp;
      }
      if (held == 0) break;  // nothing is held: the unconstrained minimiser lies inside the box, done
      // ---- v is the minimiser of the working set: multipliers of the held inputs, dJ/dv_j = 2 (R11 v_j + b' lam_{j+1}),
      // from the costate lam_j = Q xb_j + A_d' lam_{j+1} of the trajectory the forward sweep has just stored
      int jw = -1;
      {
        double lam[4], worst = 0.0;
        {
          double xe[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) xe[c] = Xb(n, c);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            double sacc = QT[sym(i, 0)] * xe[0];
#pragma unroll
            for (int c = 1; c < 4; ++c) sacc = fma(QT[sym(i, c)], xe[c], sacc);
            lam[i] = sacc;  // half the gradient of the cost with respect to the state
          }
        }
        auto fetch_co = [&](int j, int s) {
          if (j >= 0) {
            fetch_common(j, t + j, s, vcur);
#pragma unroll
            for (int c = 0; c < 4; ++c) box_cp8(&rs(s, 2 + c), &Xb(j, c));
          }
          box_commit();
        };
#pragma unroll
        for (int i = 0; i < ACRO_BOX_RING; ++i) fetch_co(n - 1 - i, i);
        int s = 0;
        for (int j = n - 1; j >= 0; --j) {
          box_wait<ACRO_BOX_RING - 1>();
          const double rv = rs(s, 0);
          const int rf = ring_flag(s);
          const double xb0 = rs(s, 2), xb1 = rs(s, 3), xb2 = rs(s, 4), xb3 = rs(s, 5);
          const LinD L = ring_lin(t + j, s);
          if (rf != 0) {
            const double t1 = R11 * rv, t2 = L.b[0] * lam[2], t3 = L.b[1] * lam[3];
            const double g = t1 + t2 + t3, scale = fabs(t1) + fabs(t2) + fabs(t3);
            const double viol = double(rf) * g;  // must be <= 0 at the upper bound, >= 0 at the lower one
            if (viol > 1e-10 * scale + 1e-300 && viol / (scale + 1e-300) > worst) {
              worst = viol / (scale + 1e-300);
              jw = j;
            }
          }
          double al[4];
          box_At(L, dt, lam, al);
          const double xbv[4] = {xb0, xb1, xb2, xb3};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            double sacc = w.Q(i, 0) * xbv[0];
#pragma unroll
            for (int c = 1; c < 4; ++c) sacc = fma(w.Q(i, c), xbv[c], sacc);
            lam[i] = sacc + al[i];
          }
          fetch_co(j - ACRO_BOX_RING, s);  // after the step: the slot has been consumed
          s = (s + 1 == ACRO_BOX_RING) ? 0 : s + 1;
        }
        box_wait<0>();
      }
      if (jw < 0) break;  // optimal
      wk[slot(jw)] = 0;   // release the input whose multiplier has the wrong sign (the most violating one)
      --held;
    }
    sweeps += it + 1;
    if (it >= a.max_iter) stat = 1;
    if (a.n_active) a.n_active[int64_t(t) * B + b] = held;
    // ---- apply the first move (tt:53-56)
    const double ur0 = (t < a.N - 1) ? ref.U(t, 0) : a.uf[0], ur1 = uref1(t);
    const double u0 = ur0 + fmin(fmax(0.0, -tau - ur0), tau - ur0), u1 = ur1 + vcur[slot(0)];
    a.Ur[soa(t, 2, 0, a.T - 1, b)] = u0;
    a.Ur[soa(t, 2, 1, a.T - 1, b)] = u1;
    double xn[4];
    rk4_step(m, x, u0, u1, xn);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      x[c] = xn[c];
      a.Xr[soa(t + 1, 4, c, a.T, b)] = x[c];
    }
  }
  if (a.n_sweeps) a.n_sweeps[b] = sweeps;
  if (a.status) a.status[b] = stat;
}

}  // namespace acro
