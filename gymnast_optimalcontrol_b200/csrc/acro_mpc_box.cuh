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
//   * the sweeps read their per-problem operands three window steps ahead into a rotating set of registers;
//   * the window slides by renaming: v and the working-set flags live in circular slots (window index j of step t is
//     slot (t + j) mod n), the candidate v* is written to a second buffer and a full step swaps the two pointers,
//     the flags are bytes, the number of held inputs is carried along - no copy loops;
//   * with a shared reference and an empty working set the gains of the backward sweep are the same for every
//     problem: k_mpc_box_gains computes them once per MPC step (table [T-1][n][4]) and a solve whose working set
//     is empty runs only the forward sweep (the feasibility check of the unconstrained minimiser) against the table.
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
};

// doubles of workspace per problem: v, v* (2n), gains (5n), states (4n + 4), flags (n bytes, n doubles reserved)
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

template <bool RPB>
__global__ void k_mpc_track_box(const __grid_constant__ MpcBoxArgs a) {
  const int64_t B = a.B, b = blockIdx.x * int64_t(blockDim.x) + threadIdx.x;
  if (b >= B) return;
  const WV<false> w(a.kw, B, b);
  const RefV<RPB> ref{a.rx, a.ru, a.N, b};
  const Model& m = a.m;
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
  int8_t* const wk = reinterpret_cast<int8_t*>(a.ws + (11LL * n + 4) * B) + b;  // +1 / -1: held at the upper / lower bound
  auto Kg = [&](int j, int c) -> double& { return ws[int64_t(2 * n + 4 * j + c) * B]; };
  auto kg = [&](int j) -> double& { return ws[int64_t(6 * n + j) * B]; };
  auto Xb = [&](int j, int c) -> double& { return ws[int64_t(7 * n + 4 * j + c) * B]; };
  auto lin_at = [&](int tj) {
    if (!RPB && a.wtab) return box_load_row(a.wtab, tj);
    return tj < n_lin ? load_lin(a.lin, tj, ld, b) : Lf;
  };
  auto uref1 = [&](int tj) {
    if (!RPB && a.wtab) return __ldg(a.wtab + int64_t(tj) * 11 + 10);
    return tj < a.N - 1 ? ref.U(tj, 1) : a.uf[1];
  };
  int s0 = 0;  // slot of window index 0 (= t mod n)
  auto slot = [&](int j) -> int64_t {
    const int s = s0 + j;
    return int64_t(s >= n ? s - n : s) * B;
  };

  double x[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    x[c] = a.x0[c * B + b];
    a.Xr[soa(0, 4, c, a.T, b)] = x[c];
  }
  int sweeps = 0, stat = 0, held = 0;
  for (int t = 0; t < a.T - 1; ++t) {
    // ---- start: the previous solution shifted by one step (same absolute times, hence still feasible and with the
    // same working set), the new last input at the point of its interval closest to 0
    if (t == 0) {
      for (int j = 0; j < n; ++j) {
        const double ur = uref1(j), lo = -tau - ur, hi = tau - ur;
        const double v = fmin(fmax(0.0, lo), hi);
        const int f = (v >= hi) ? 1 : ((v <= lo) ? -1 : 0);
        vcur[slot(j)] = v;
        wk[slot(j)] = int8_t(f);
        held += (f != 0);
      }
    } else {
      held -= (wk[slot(0)] != 0);  // the input that leaves the window
      s0 = (s0 + 1 == n) ? 0 : s0 + 1;
      const double ur = uref1(t + n - 1), lo = -tau - ur, hi = tau - ur;
      const double v = fmin(fmax(0.0, lo), hi);
      const int f = (v >= hi) ? 1 : ((v <= lo) ? -1 : 0);
      vcur[slot(n - 1)] = v;
      wk[slot(n - 1)] = int8_t(f);
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
        struct BwRec {
          LinD L;
          double v;
          int f;
        };
        auto load_bw = [&](int j) {
          BwRec r;
          j = max(j, 0);
          if (RPB) r.L = lin_at(t + j);
          r.v = vcur[slot(j)];
          r.f = wk[slot(j)];
          return r;
        };
        auto bw_step = [&](const BwRec& r, int j) {
          const LinD L = RPB ? r.L : lin_at(t + j);
          double S[10], F[4], Pb[4], pn[4];
          box_products(P, L, dt, S, F, Pb);
          if (r.f != 0) {  // held at v_j (its gain row is never read)
            double y[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) y[c] = fma(Pb[c], r.v, p[c]);
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
        BwRec r0 = load_bw(n - 1), r1 = load_bw(n - 2), r2 = load_bw(n - 3);
        for (int j = n - 1; j >= 0; j -= 3) {
          bw_step(r0, j);  // a record is refilled right after the step that consumed it
          r0 = load_bw(j - 3);
          if (j >= 1) {
            bw_step(r1, j - 1);
            r1 = load_bw(j - 4);
          }
          if (j >= 2) {
            bw_step(r2, j - 2);
            r2 = load_bw(j - 5);
          }
        }
      }
      // ---- forward sweep (closed loop): minimiser over the free inputs, blocking input
      double alpha = 1.0, dmax = 0.0, vmax = 0.0;
      int jb = -1, sb = 0;
      // the candidate vs of input j: distance to v, blocking ratio
      auto consider = [&](int j, double vj, double vs, double ur) {
        const double lo = -tau - ur, hi = tau - ur, d = vs - vj;
        dmax = fmax(dmax, fabs(d));
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
        // nothing is held: gains from the table, table rows one step ahead, v three steps ahead
        double xs[4] = {x0w[0], x0w[1], x0w[2], x0w[3]};
        const double* const krow = a.ktab + int64_t(t) * n * 4;
        const double* const wrow = a.wtab + int64_t(t) * 11;
        struct TabRec {
          double K[4], r[11];
        };
        auto load_tab = [&](int j) {
          TabRec q;
          j = min(j, n - 1);
#pragma unroll
          for (int c = 0; c < 4; ++c) q.K[c] = __ldg(krow + j * 4 + c);
#pragma unroll
          for (int c = 0; c < 11; ++c) q.r[c] = __ldg(wrow + j * 11 + c);
          return q;
        };
        auto tab_step = [&](const TabRec& q, double vj, int j) {
          const double vs = fma(q.K[3], xs[3], fma(q.K[2], xs[2], fma(q.K[1], xs[1], fma(q.K[0], xs[0], 0.0))));
          consider(j, vj, vs, q.r[10]);
          vmax = fmax(vmax, fabs(vj));
          voth[slot(j)] = vs;
          double xn[4];
          xn[0] = fma(dt, xs[2], xs[0]);
          xn[1] = fma(dt, xs[3], xs[1]);
          xn[2] = fma(q.r[8], vs, fma(q.r[3], xs[3], fma(q.r[2], xs[2], fma(q.r[1], xs[1], q.r[0] * xs[0]))));
          xn[3] = fma(q.r[9], vs, fma(q.r[7], xs[3], fma(q.r[6], xs[2], fma(q.r[5], xs[1], q.r[4] * xs[0]))));
#pragma unroll
          for (int c = 0; c < 4; ++c) xs[c] = xn[c];
        };
        auto v_at = [&](int j) { return vcur[slot(min(j, n - 1))]; };
        double v0 = v_at(0), v1 = v_at(1), v2 = v_at(2), v3 = v_at(3), v4 = v_at(4), v5 = v_at(5);
        TabRec qa = load_tab(0), qb = load_tab(1);
        for (int j = 0; j < n; j += 2) {
          const double va = v0, vb = v1;
          v0 = v2;
          v1 = v3;
          v2 = v4;
          v3 = v5;
          v4 = v_at(j + 6);
          v5 = v_at(j + 7);
          tab_step(qa, va, j);
          qa = load_tab(j + 2);
          if (j + 1 < n) {
            tab_step(qb, vb, j + 1);
            qb = load_tab(j + 3);
          }
        }
      } else {
        double xs[4] = {x0w[0], x0w[1], x0w[2], x0w[3]};
        struct FwRec {
          LinD L;
          double v, K[4], k, ur;
          int f;
        };
        auto load_fw = [&](int j) {
          FwRec r;
          j = min(j, n - 1);
          if (RPB) {
            r.L = lin_at(t + j);
            r.ur = uref1(t + j);
          }
          r.v = vcur[slot(j)];
          r.f = wk[slot(j)];
#pragma unroll
          for (int c = 0; c < 4; ++c) r.K[c] = Kg(j, c);
          r.k = kg(j);
          return r;
        };
        const bool keep_states = held > 0;  // the costate sweep below reads them
        auto fw_step = [&](const FwRec& r, int j) {
          const double vj = r.v;
          double vs = vj;
          if (keep_states) {
#pragma unroll
            for (int c = 0; c < 4; ++c) Xb(j, c) = xs[c];
          }
          if (r.f == 0) {
            vs = fma(r.K[3], xs[3], fma(r.K[2], xs[2], fma(r.K[1], xs[1], fma(r.K[0], xs[0], r.k))));
            consider(j, vj, vs, RPB ? r.ur : uref1(t + j));
          }
          vmax = fmax(vmax, fabs(vj));
          voth[slot(j)] = vs;
          const LinD L = RPB ? r.L : lin_at(t + j);
          double xn[4];
          box_plant(L, dt, xs, vs, xn);
#pragma unroll
          for (int c = 0; c < 4; ++c) xs[c] = xn[c];
        };
        FwRec r0 = load_fw(0), r1 = load_fw(1), r2 = load_fw(2);
        for (int j = 0; j < n; j += 3) {
          fw_step(r0, j);  // a record is refilled right after the step that consumed it
          r0 = load_fw(j + 3);
          if (j + 1 < n) {
            fw_step(r1, j + 1);
            r1 = load_fw(j + 4);
          }
          if (j + 2 < n) {
            fw_step(r2, j + 2);
            r2 = load_fw(j + 5);
          }
        }
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
        wk[slot(jb)] = int8_t(sb);
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
        struct CoRec {
          LinD L;
          double v, xb[4];
          int f;
        };
        auto load_co = [&](int j) {
          CoRec r;
          j = max(j, 0);
          if (RPB) r.L = lin_at(t + j);
          r.v = vcur[slot(j)];
          r.f = wk[slot(j)];
#pragma unroll
          for (int c = 0; c < 4; ++c) r.xb[c] = Xb(j, c);
          return r;
        };
        auto co_step = [&](const CoRec& r, int j) {
          const LinD L = RPB ? r.L : lin_at(t + j);
          if (r.f != 0) {
            const double t1 = R11 * r.v, t2 = L.b[0] * lam[2], t3 = L.b[1] * lam[3];
            const double g = t1 + t2 + t3, scale = fabs(t1) + fabs(t2) + fabs(t3);
            const double viol = double(r.f) * g;  // must be <= 0 at the upper bound, >= 0 at the lower one
            if (viol > 1e-10 * scale + 1e-300 && viol / (scale + 1e-300) > worst) {
              worst = viol / (scale + 1e-300);
              jw = j;
            }
          }
          double al[4];
          box_At(L, dt, lam, al);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            double sacc = w.Q(i, 0) * r.xb[0];
#pragma unroll
            for (int c = 1; c < 4; ++c) sacc = fma(w.Q(i, c), r.xb[c], sacc);
            lam[i] = sacc + al[i];
          }
        };
        CoRec r0 = load_co(n - 1), r1 = load_co(n - 2), r2 = load_co(n - 3);
        for (int j = n - 1; j >= 0; j -= 3) {
          co_step(r0, j);  // a record is refilled right after the step that consumed it
          r0 = load_co(j - 3);
          if (j >= 1) {
            co_step(r1, j - 1);
            r1 = load_co(j - 4);
          }
          if (j >= 2) {
            co_step(r2, j - 2);
            r2 = load_co(j - 5);
          }
        }
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
