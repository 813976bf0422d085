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
//                                   (free: K_j = -F/G, k_j = -g/G;  fixed: P <- Q + A'PA, p <- A'(p + P b v_j)),
//                                   costate of (xb, v) and with it the multipliers dJ/dv_j of the inputs in W
//     forward sweep (closed loop) : the minimiser v* over the free inputs; largest step alpha <= 1 from v towards v*
//                                   that stays inside the box, and the input that blocks it
//     blocked          -> v += alpha (v* - v), the blocking input joins W
//     moved, unblocked -> v = v*
//     at the minimiser -> drop the input of W whose multiplier has the wrong sign (the most violating), or stop
//
// Finite termination, no tolerance on the answer other than rounding; consecutive MPC steps warm-start from the
// shifted solution (typically one or two sweeps per step).  Per-problem scratch lives in a caller-supplied workspace
// ws[e][B] (e fastest across problems: coalesced).  The tests compare it with the same QP condensed and solved with
// dense matrices; the reference itself would hand it to IPOPT (tol 1e-6): parity against IPOPT is unpinned.
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
  double* ws;         // [11 (H-1) + 4 + (H-1)][B]
  double tau;
  int max_iter;       // active-set iterations per solve
  double *Xr, *Ur;
  int32_t* n_sweeps;  // [B] total active-set iterations of the problem (out, may be null)
  int32_t* n_active;  // [T-1][B] inputs at a bound in the solution of step t (out, may be null)
  int32_t* status;    // [B] 0, or 1 if some step ran into max_iter (out, may be null)
};

// doubles of workspace per problem
__host__ __device__ inline int64_t mpc_box_ws_per_problem(int H) { return 12LL * (H - 1) + 4; }

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
  const LinD Lf = linearize_d(m, a.xf, a.uf[0], a.uf[1]);
  double QT[10];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = i; j < 4; ++j) QT[sym(i, j)] = a.qt_per_problem ? a.QT[(i * 4 + j) * B + b] : a.QT[i * 4 + j];
  // workspace views (element e of this problem at ws[e * B + b])
  double* const ws = a.ws + b;
  auto V = [&](int j) -> double& { return ws[int64_t(j) * B]; };
  auto Wk = [&](int j) -> double& { return ws[int64_t(n + j) * B]; };
  auto Kg = [&](int j, int c) -> double& { return ws[int64_t(2 * n + 4 * j + c) * B]; };
  auto kg = [&](int j) -> double& { return ws[int64_t(6 * n + j) * B]; };
  auto Xb = [&](int j, int c) -> double& { return ws[int64_t(7 * n + 4 * j + c) * B]; };
  auto Vs = [&](int j) -> double& { return ws[int64_t(11 * n + 4 + j) * B]; };
  auto lin_at = [&](int tj) { return tj < n_lin ? load_lin(a.lin, tj, ld, b) : Lf; };
  auto uref1 = [&](int tj) { return tj < a.N - 1 ? ref.U(tj, 1) : a.uf[1]; };

  double x[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    x[c] = a.x0[c * B + b];
    a.Xr[soa(0, 4, c, a.T, b)] = x[c];
  }
  int sweeps = 0, stat = 0;
  for (int t = 0; t < a.T - 1; ++t) {
    // ---- start: the previous solution shifted by one step (same absolute times, hence still feasible), the new last
    // input at the point of its interval closest to 0; W = the inputs that sit on a bound
    int held = 0;  // inputs in the working set
    for (int j = 0; j < n; ++j) {
      const double ur = uref1(t + j), lo = -tau - ur, hi = tau - ur;
      double v = (t > 0 && j + 1 < n) ? V(j + 1) : 0.0;
      v = fmin(fmax(v, lo), hi);
      V(j) = v;
      const double wj = (v >= hi) ? 1.0 : ((v <= lo) ? -1.0 : 0.0);
      Wk(j) = wj;
      held += (wj != 0.0);
    }
    double x0w[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) x0w[c] = x[c] - ((t < a.N) ? ref.X(t, c) : a.xf[c]);
    auto rollout = [&]() {  // xb <- states of the window under v
      double xs[4] = {x0w[0], x0w[1], x0w[2], x0w[3]};
      for (int j = 0; j < n; ++j) {
#pragma unroll
        for (int c = 0; c < 4; ++c) Xb(j, c) = xs[c];
        const LinD L = lin_at(t + j);
        double xn[4];
        box_plant(L, dt, xs, V(j), xn);
#pragma unroll
        for (int c = 0; c < 4; ++c) xs[c] = xn[c];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) Xb(n, c) = xs[c];
    };
    // the state trajectory of v is only needed for the multipliers of held inputs
    if (held > 0) rollout();
    int it = 0;
    for (; it < a.max_iter; ++it) {
      const bool need_lam = held > 0;
      // ---- backward sweep
      double P[10], p[4] = {0.0, 0.0, 0.0, 0.0}, lam[4];
#pragma unroll
      for (int e = 0; e < 10; ++e) P[e] = QT[e];
      if (need_lam) {
        double xe[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) xe[c] = Xb(n, c);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          double s = P[sym(i, 0)] * xe[0];
#pragma unroll
          for (int c = 1; c < 4; ++c) s = fma(P[sym(i, c)], xe[c], s);
          lam[i] = s;  // half the gradient of the cost with respect to the state
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) lam[i] = 0.0;
      }
      int jw = -1;
      double worst = 0.0;
      for (int j = n - 1; j >= 0; --j) {
        const LinD L = lin_at(t + j);
        const double vj = V(j), wj = Wk(j);
        // multiplier of a held input: dJ/dv_j = 2 (R11 v_j + b' lam_{j+1})
        if (wj != 0.0) {
          const double t1 = R11 * vj, t2 = L.b[0] * lam[2], t3 = L.b[1] * lam[3];
          const double g = t1 + t2 + t3, scale = fabs(t1) + fabs(t2) + fabs(t3);
          const double viol = wj * g;  // must be <= 0 at the upper bound, >= 0 at the lower one
          if (viol > 1e-10 * scale + 1e-300 && viol / (scale + 1e-300) > worst) {
            worst = viol / (scale + 1e-300);
            jw = j;
          }
        }
        // M = P A_d, S = A_d' M (upper triangle), F = b' M, Pb = P b
        double M[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const double pi0 = P[sym(i, 0)], pi1 = P[sym(i, 1)], pi2 = P[sym(i, 2)], pi3 = P[sym(i, 3)];
          M[i][0] = fma(pi3, L.a[1][0], fma(pi2, L.a[0][0], pi0));
          M[i][1] = fma(pi3, L.a[1][1], fma(pi2, L.a[0][1], pi1));
          M[i][2] = fma(pi3, L.a[1][2], fma(pi2, L.a[0][2], dt * pi0));
          M[i][3] = fma(pi3, L.a[1][3], fma(pi2, L.a[0][3], dt * pi1));
        }
        double S[10], F[4], Pb[4];
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
        double pn[4];
        if (wj != 0.0) {  // held at v_j
#pragma unroll
          for (int c = 0; c < 4; ++c) Kg(j, c) = 0.0;
          kg(j) = vj;
          double y[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) y[c] = fma(Pb[c], vj, p[c]);
          box_At(L, dt, y, pn);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = i; c < 4; ++c) P[sym(i, c)] = w.Q(i, c) + S[sym(i, c)];
        } else {
          const double G = R11 + fma(L.b[1], Pb[3], L.b[0] * Pb[2]);
          const double gg = fma(L.b[1], p[3], L.b[0] * p[2]);
          const double iG = 1.0 / G;
          double Kr[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            Kr[c] = -F[c] * iG;
            Kg(j, c) = Kr[c];
          }
          kg(j) = -gg * iG;
          box_At(L, dt, p, pn);
#pragma unroll
          for (int c = 0; c < 4; ++c) pn[c] = fma(Kr[c], gg, pn[c]);
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = i; c < 4; ++c) P[sym(i, c)] = w.Q(i, c) + S[sym(i, c)] + Kr[i] * F[c];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) p[c] = pn[c];
        // costate: lam_j = Q xb_j + A_d' lam_{j+1}
        if (need_lam) {
          double xe[4], al[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) xe[c] = Xb(j, c);
          box_At(L, dt, lam, al);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            double s = w.Q(i, 0) * xe[0];
#pragma unroll
            for (int c = 1; c < 4; ++c) s = fma(w.Q(i, c), xe[c], s);
            lam[i] = s + al[i];
          }
        }
      }
      // ---- forward sweep (closed loop): minimiser over the free inputs, blocking input
      double alpha = 1.0, dmax = 0.0, vmax = 0.0;
      int jb = -1;
      double sb = 0.0;
      {
        double xs[4] = {x0w[0], x0w[1], x0w[2], x0w[3]};
        for (int j = 0; j < n; ++j) {
          const double vj = V(j), wj = Wk(j);
          double vs = vj;
          if (wj == 0.0) {
            vs = fma(Kg(j, 3), xs[3], fma(Kg(j, 2), xs[2], fma(Kg(j, 1), xs[1], fma(Kg(j, 0), xs[0], kg(j)))));
            const double ur = uref1(t + j), lo = -tau - ur, hi = tau - ur, d = vs - vj;
            dmax = fmax(dmax, fabs(d));
            if (vs > hi && d > 0.0) {
              const double r = (hi - vj) / d;
              if (r < alpha) {
                alpha = r;
                jb = j;
                sb = 1.0;
              }
            } else if (vs < lo && d < 0.0) {
              const double r = (lo - vj) / d;
              if (r < alpha) {
                alpha = r;
                jb = j;
                sb = -1.0;
              }
            }
          }
          vmax = fmax(vmax, fabs(vj));
          Vs(j) = vs;
          const LinD L = lin_at(t + j);
          double xn[4];
          box_plant(L, dt, xs, vs, xn);
#pragma unroll
          for (int c = 0; c < 4; ++c) xs[c] = xn[c];
        }
      }
      if (jb >= 0) {  // blocked: partial step, the blocking input joins the working set
        for (int j = 0; j < n; ++j) V(j) = fma(alpha, Vs(j) - V(j), V(j));
        const double ur = uref1(t + jb);
        V(jb) = (sb > 0.0) ? tau - ur : -tau - ur;
        Wk(jb) = sb;
        ++held;
        rollout();
        continue;
      }
      if (dmax > 1e-11 * (1.0 + vmax)) {  // full step to the minimiser of the current working set
        for (int j = 0; j < n; ++j) V(j) = Vs(j);
        if (held == 0) break;  // nothing is held: the unconstrained minimiser lies inside the box, done
        rollout();
        continue;
      }
      if (jw >= 0) {  // at the minimiser: release the input whose multiplier has the wrong sign
        Wk(jw) = 0.0;
        --held;
        continue;
      }
      break;  // optimal
    }
    sweeps += it + 1;
    if (it >= a.max_iter) stat = 1;
    if (a.n_active) {
      int na = 0;
      for (int j = 0; j < n; ++j) na += (Wk(j) != 0.0);
      a.n_active[int64_t(t) * B + b] = na;
    }
    // ---- apply the first move (tt:53-56)
    const double ur0 = (t < a.N - 1) ? ref.U(t, 0) : a.uf[0], ur1 = uref1(t);
    const double u0 = ur0 + fmin(fmax(0.0, -tau - ur0), tau - ur0), u1 = ur1 + V(0);
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
