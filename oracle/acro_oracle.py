"""NumPy CPU oracle for the batched acrobot optimal-control hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(``gymnast_optimalcontrol_b200/``) imports this module; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may use it, and only as the checker or the CPU baseline.

It restates, in closed form, what the reference computes with SymPy-lambdified
closures and per-step NumPy calls.  Every function names the reference lines it
follows (paths relative to the reference root; ``tg`` = trajectory_generation.py,
``tt`` = trajectory_tracking.py).  All functions accept leading batch axes
(``x`` of shape ``(..., 4)``) so the same code checks one problem or a batch.

Pinning (see tests/test_oracle_golden.py and tests/golden/make_golden.py):
the restatement is checked against fixtures produced by importing the
unmodified reference in the build container, and against the two trajectory
files the reference ships.  The MPC QP (tt:73-140) is solved by CasADi/IPOPT in
the reference; CasADi is not installed and not vendored, so ``solver_mpc`` here
is a dense-KKT and a Riccati restatement of that QP checked against each other:
MPC PARITY IS UNPINNED against IPOPT itself.  What the reference ships of its MPC are figures
(figures/mpc/*.png); the restatement reproduces every peak of their error curves for dx = 0.05 (input box), 0.1, 0.15
and 0.2 to the accuracy they can be read at (2-3 %, 0.04 s: tests/mpc_figures.py) - a figure-level pin, not 1e-9.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------
# D4: model constants (dynamics.py:15-61, 173-175)
# --------------------------------------------------------------------------
dt = 2e-2
ns = 4
ni = 2

PARAM_SETS = {
    1: dict(m1=1.0, m2=1.0, l1=1.0, lc1=0.5, l2=1.0, lc2=0.5, I1=0.33, I2=0.33, g=9.81, f1=1.0, f2=1.0),
    2: dict(m1=2.0, m2=2.0, l1=1.5, lc1=0.75, l2=1.5, lc2=0.75, I1=1.5, I2=1.5, g=9.81, f1=1.0, f2=1.0),
    3: dict(m1=1.5, m2=1.5, l1=2.0, lc1=1.0, l2=2.0, lc2=1.0, I1=2.0, I2=2.0, g=9.81, f1=1.0, f2=1.0),
}


class Model:
    """Lumped constants of M(q), C(q,qd), G(q), F (dynamics.py:64-90)."""

    def __init__(self, p=None, step=dt, actuated_tau1=False):
        p = dict(PARAM_SETS[1] if p is None else p)
        self.a1 = p["I1"] + p["I2"] + p["lc1"] ** 2 * p["m1"] + p["m2"] * (p["l1"] ** 2 + p["lc2"] ** 2)
        self.h = p["m2"] * p["l1"] * p["lc2"]
        self.a3 = p["I2"] + p["lc2"] ** 2 * p["m2"]
        self.g1 = p["g"] * (p["lc1"] * p["m1"] + p["m2"] * p["l1"])
        self.g2 = p["g"] * p["m2"] * p["lc2"]
        self.f1 = p["f1"]
        self.f2 = p["f2"]
        self.dt = step
        # dynamics.py:205 forces tau_1 = 0 in the plant; the fully-actuated variant
        # (fully_actuated_ref_gen.py:20-73) lets uu[0] act on the first joint.
        self.tau1_gain = 1.0 if actuated_tau1 else 0.0


DEFAULT = Model()

# weights (tg:16-18, tt:173-175, tt:38-39)
Q_NEWTON = np.diag([130.0, 30.0, 0.0001, 0.0001])
R_NEWTON = np.diag([1e-6, 1.5])
QT_NEWTON = np.diag([130.0, 130.0, 1.0, 1.0])
Q_LQR = np.diag([100.0, 100.0, 10.0, 10.0])
R_LQR = np.diag([1.0, 1.0])
Q_MPC = np.diag([120.0, 100.0, 0.0001, 0.0001])
R_MPC = np.diag([1e-6, 10.0])
X_F = np.array([np.pi, 0.0, 0.0, 0.0])
U_F = np.array([0.0, 0.0])


# --------------------------------------------------------------------------
# D1-D3
# --------------------------------------------------------------------------
def continuous_dynamics(xx, uu, m=DEFAULT):
    """xdot = [qd ; M^-1 (tau - (C+F) qd - G)] with tau = [0, uu[1]]  (dynamics.py:197-213)."""
    xx = np.asarray(xx, dtype=float)
    uu = np.asarray(uu, dtype=float)
    th1, th2, w1, w2 = xx[..., 0], xx[..., 1], xx[..., 2], xx[..., 3]
    s1, s2, c2, s12 = np.sin(th1), np.sin(th2), np.cos(th2), np.sin(th1 + th2)
    M11 = m.a1 + 2.0 * m.h * c2
    M12 = m.a3 + m.h * c2
    M22 = m.a3
    tau1 = m.tau1_gain * uu[..., 0]
    r1 = tau1 + m.h * s2 * w2 * w1 + m.h * s2 * (w1 + w2) * w2 - m.f1 * w1 - (m.g1 * s1 + m.g2 * s12)
    r2 = uu[..., 1] - m.h * s2 * w1 * w1 - m.f2 * w2 - m.g2 * s12
    det = M11 * M22 - M12 * M12
    qdd1 = (M22 * r1 - M12 * r2) / det
    qdd2 = (M11 * r2 - M12 * r1) / det
    return np.stack(np.broadcast_arrays(w1, w2, qdd1, qdd2), axis=-1)


def dynamics(xx, uu, m=DEFAULT):
    """Classic RK4 step with zero-order hold on the input (dynamics.py:177-195)."""
    xx = np.asarray(xx, dtype=float)
    uu = np.asarray(uu, dtype=float)
    h = m.dt
    k1 = continuous_dynamics(xx, uu, m)
    k2 = continuous_dynamics(xx + (h / 2) * k1, uu, m)
    k3 = continuous_dynamics(xx + (h / 2) * k2, uu, m)
    k4 = continuous_dynamics(xx + h * k3, uu, m)
    return xx + h * (k1 + 2 * k2 + 2 * k3 + k4) / 6.0


def Calculate_A_B_matrixes(x_t, u_t, m=DEFAULT):
    """Continuous Jacobians df/dx (4x4), df/du (4x2) (dynamics.py:153-170, 217-226).

    Uses A_c[2:4, j] = M^-1 (d rhs/dx_j - (dM/dx_j) qdd), B_c[2:4, :] = M^-1 d tau/du.
    """
    x_t = np.asarray(x_t, dtype=float)
    u_t = np.asarray(u_t, dtype=float)
    th1, th2, w1, w2 = x_t[..., 0], x_t[..., 1], x_t[..., 2], x_t[..., 3]
    s1, c1 = np.sin(th1), np.cos(th1)
    s2, c2 = np.sin(th2), np.cos(th2)
    s12, c12 = np.sin(th1 + th2), np.cos(th1 + th2)
    M11 = m.a1 + 2.0 * m.h * c2
    M12 = m.a3 + m.h * c2
    M22 = m.a3 + 0.0 * c2
    det = M11 * M22 - M12 * M12
    tau1 = m.tau1_gain * u_t[..., 0]
    r1 = tau1 + m.h * s2 * w2 * w1 + m.h * s2 * (w1 + w2) * w2 - m.f1 * w1 - (m.g1 * s1 + m.g2 * s12)
    r2 = u_t[..., 1] - m.h * s2 * w1 * w1 - m.f2 * w2 - m.g2 * s12
    qdd1 = (M22 * r1 - M12 * r2) / det
    qdd2 = (M11 * r2 - M12 * r1) / det

    def minv(b1, b2):
        return (M22 * b1 - M12 * b2) / det, (M11 * b2 - M12 * b1) / det

    # columns of d rhs / dx, corrected by -(dM/dx_j) qdd for j = theta2
    d1 = (-(m.g1 * c1 + m.g2 * c12), -m.g2 * c12)
    dM11, dM12 = -2.0 * m.h * s2, -m.h * s2
    d2 = (
        m.h * c2 * (w1 * w2 + (w1 + w2) * w2) - m.g2 * c12 - (dM11 * qdd1 + dM12 * qdd2),
        -m.h * c2 * w1 * w1 - m.g2 * c12 - (dM12 * qdd1),
    )
    d3 = (2.0 * m.h * s2 * w2 - m.f1, -2.0 * m.h * s2 * w1)
    d4 = (2.0 * m.h * s2 * (w1 + w2), -m.f2 + 0.0 * s2)
    A = np.zeros(x_t.shape[:-1] + (4, 4))
    A[..., 0, 2] = 1.0
    A[..., 1, 3] = 1.0
    for j, d in enumerate((d1, d2, d3, d4)):
        a, b = minv(d[0], d[1])
        A[..., 2, j] = a
        A[..., 3, j] = b
    B = np.zeros(x_t.shape[:-1] + (4, 2))
    b0 = minv(1.0, 0.0)
    b1 = minv(0.0, 1.0)
    B[..., 2, 0] = m.tau1_gain * b0[0]
    B[..., 3, 0] = m.tau1_gain * b0[1]
    B[..., 2, 1] = b1[0]
    B[..., 3, 1] = b1[1]
    return A, B


def discretize_linearization(Ac, Bc, step=dt):
    """Forward-Euler discretisation A_d = I + dt A_c, B_d = dt B_c (tg:161-164)."""
    return np.eye(4) + step * np.asarray(Ac), step * np.asarray(Bc)


def linearize_discrete(x_t, u_t, m=DEFAULT):
    Ac, Bc = Calculate_A_B_matrixes(x_t, u_t, m)
    return discretize_linearization(Ac, Bc, m.dt)


# --------------------------------------------------------------------------
# G1-G11 (trajectory_generation.py)
# --------------------------------------------------------------------------
def _mv(M, v):
    return np.einsum("...ij,...j->...i", M, v)


def _quad(v, M):
    return np.einsum("...i,...ij,...j->...", v, M, v)


def simulate_open_loop(x0, u_traj, m=DEFAULT):
    """x_{t+1} = dynamics(x_t, u_t)  (tg:74-87).  u_traj: (..., Nc, 2) -> (..., Nc+1, 4)."""
    u_traj = np.asarray(u_traj, dtype=float)
    Nc = u_traj.shape[-2]
    x = np.zeros(u_traj.shape[:-2] + (Nc + 1, 4))
    x[..., 0, :] = x0
    for t in range(Nc):
        x[..., t + 1, :] = dynamics(x[..., t, :], u_traj[..., t, :], m)
    return x


def derivatives_Cost(x, x_ref, u, u_ref, Q, R, Q_T=None, terminal=False):
    """Stage / terminal cost with gradients and Hessians 2Q, 2R, 2Q_T (tg:89-114)."""
    dx = np.asarray(x) - np.asarray(x_ref)
    if terminal:
        return _quad(dx, Q_T), 2.0 * _mv(Q_T, dx), 2.0 * np.asarray(Q_T)
    du = np.asarray(u) - np.asarray(u_ref)
    return _quad(dx, Q) + _quad(du, R), 2.0 * _mv(Q, dx), 2.0 * _mv(R, du), 2.0 * np.asarray(Q), 2.0 * np.asarray(R)


def compute_costate_trajectory(x_traj, u_traj, x_ref, u_ref, Q=Q_NEWTON, R=R_NEWTON, Q_T=QT_NEWTON, m=DEFAULT):
    """lambda_T = 2 Q_T dx_T ; lambda_t = 2 Q dx_t + A_d^T lambda_{t+1}  (tg:138-159)."""
    x_traj = np.asarray(x_traj, dtype=float)
    N = x_traj.shape[-2]
    lam = np.zeros_like(x_traj)
    lam[..., N - 1, :] = 2.0 * _mv(Q_T, x_traj[..., N - 1, :] - x_ref[..., N - 1, :])
    for t in range(N - 2, -1, -1):
        Ad, _ = linearize_discrete(x_traj[..., t, :], u_traj[..., t, :], m)
        gx = 2.0 * _mv(Q, x_traj[..., t, :] - x_ref[..., t, :])
        lam[..., t, :] = gx + np.einsum("...ji,...j->...i", Ad, lam[..., t + 1, :])
    return lam


def build_stage_lists(x_traj, u_traj, x_ref, u_ref, Q=Q_NEWTON, R=R_NEWTON, Q_T=QT_NEWTON, m=DEFAULT):
    """Per-step (A_d, B_d, q_t, r_t) and the terminal (2Q_T, q_T)  (tg:166-181, 116-136).

    The Hessian blocks are the constants 2Q, 2R and S = 0 (tg:122-124); the costate
    handed in by the reference (tg:175) is ignored there, so it is not an input here.
    """
    x_traj = np.asarray(x_traj, dtype=float)
    u_traj = np.asarray(u_traj, dtype=float)
    Ad, Bd = linearize_discrete(x_traj[..., :-1, :], u_traj, m)
    q = 2.0 * _mv(Q, x_traj[..., :-1, :] - x_ref[..., :-1, :])
    r = 2.0 * _mv(R, u_traj - u_ref)
    qT = 2.0 * _mv(Q_T, x_traj[..., -1, :] - x_ref[..., -1, :])
    return Ad, Bd, q, r, 2.0 * np.asarray(Q_T), qT


def calculate_K_and_sigma(Ad, Bd, q, r, Qt2, Rt2, QT2, qT):
    """Backward affine Riccati pass (tg:183-216).

    Ad (..., T, 4, 4), Bd (..., T, 4, 2), q (..., T, 4), r (..., T, 2); Qt2 = 2Q, Rt2 = 2R,
    QT2 = 2Q_T.  Returns K (..., T, 2, 4), sigma (..., T, 2), expected reduction (...,).
    """
    T = Ad.shape[-3]
    lead = Ad.shape[:-3]
    P = np.broadcast_to(QT2, lead + (4, 4)).copy()
    p = np.array(np.broadcast_to(qT, lead + (4,)), dtype=float)
    K = np.zeros(lead + (T, 2, 4))
    sig = np.zeros(lead + (T, 2))
    red = np.zeros(lead)
    for t in range(T - 1, -1, -1):
        A = Ad[..., t, :, :]
        B = Bd[..., t, :, :]
        Bt = np.swapaxes(B, -1, -2)
        At = np.swapaxes(A, -1, -2)
        G = Rt2 + Bt @ P @ B
        F = Bt @ P @ A
        g = r[..., t, :] + _mv(Bt, p)
        Kt = -np.linalg.solve(G, F)
        st = -np.linalg.solve(G, g[..., None])[..., 0]
        red = red + np.einsum("...i,...i->...", g, st)
        Ktt = np.swapaxes(Kt, -1, -2)
        p = q[..., t, :] + _mv(At, p) - _mv(Ktt @ G, st)
        P = Qt2 + At @ P @ A - Ktt @ G @ Kt
        K[..., t, :, :] = Kt
        sig[..., t, :] = st
    return K, sig, red


def forward_closed_loop_update(x_traj, u_traj, K, sigma, gamma=1.0, m=DEFAULT):
    """u+_t = u_t + K_t (x+_t - x_t) + gamma sigma_t ; x+_{t+1} = dynamics(x+_t, u+_t)  (tg:218-229)."""
    x_traj = np.asarray(x_traj, dtype=float)
    u_traj = np.asarray(u_traj, dtype=float)
    gamma = np.asarray(gamma, dtype=float)
    xn = x_traj.copy()
    un = u_traj.copy()
    for t in range(x_traj.shape[-2] - 1):
        dx = xn[..., t, :] - x_traj[..., t, :]
        un[..., t, :] = u_traj[..., t, :] + _mv(K[..., t, :, :], dx) + gamma[..., None] * sigma[..., t, :]
        xn[..., t + 1, :] = dynamics(xn[..., t, :], un[..., t, :], m)
    return xn, un


def total_cost(x_traj, u_traj, x_ref, u_ref, Q=Q_NEWTON, R=R_NEWTON, Q_T=QT_NEWTON):
    """Sum of stage costs with Q, R plus the terminal cost with Q_T, accumulated in time order (tg:231-252)."""
    x_traj = np.asarray(x_traj, dtype=float)
    N = x_traj.shape[-2]
    cost = np.zeros(x_traj.shape[:-2])
    for t in range(N - 1):
        cost = cost + _quad(x_traj[..., t, :] - x_ref[..., t, :], Q)
        cost = cost + _quad(u_traj[..., t, :] - u_ref[..., t, :], R)
    return cost + _quad(x_traj[..., N - 1, :] - x_ref[..., N - 1, :], Q_T)


def armijo_gammas(gamma_0, beta, n=20):
    """The candidate step sizes exactly as the loop produces them: gamma *= beta (tg:344-365)."""
    out = []
    g = gamma_0
    for _ in range(n):
        out.append(g)
        g *= beta
    return out


STATUS_RUNNING, STATUS_CONVERGED, STATUS_MAX_ITERS, STATUS_LINE_SEARCH_FAILED = 0, 1, 2, 3


def newton_iteration(x_traj, u_traj, cost_k, x_ref, u_ref, gamma_0, beta=0.7, c=0.5,
                     Q=Q_NEWTON, R=R_NEWTON, Q_T=QT_NEWTON, m=DEFAULT, max_ls=20):
    """One pass of the loop body of newton_Algorithm for ONE problem (tg:329-384).

    Returns a dict with K, sigma, delta_J, sigma_norm, tried gammas and costs, the
    accepted index (-1 = line search failed) and the new iterate.
    """
    Ad, Bd, q, r, QT2, qT = build_stage_lists(x_traj, u_traj, x_ref, u_ref, Q, R, Q_T, m)
    K, sig, dJ = calculate_K_and_sigma(Ad, Bd, q, r, 2.0 * Q, 2.0 * R, QT2, qT)
    out = dict(K=K, sigma=sig, delta_J=float(dJ), sigma_norm=float(np.max(np.abs(sig))),
               gammas=[], costs=[], accepted=-1, x=x_traj, u=u_traj, cost=cost_k)
    g = gamma_0
    for i in range(max_ls):
        xn, un = forward_closed_loop_update(x_traj, u_traj, K, sig, g, m)
        cn = float(total_cost(xn, un, x_ref, u_ref, Q, R, Q_T))
        out["gammas"].append(g)
        out["costs"].append(cn)
        if cn < cost_k + c * g * dJ:
            out.update(accepted=i, x=xn, u=un, cost=cn)
            break
        g *= beta
    return out


def newton_Algorithm(x0, x_ref, u_ref, max_iters, tol=1e-6, beta=0.7, c=0.5, gamma_0=1,
                     Q=Q_NEWTON, R=R_NEWTON, Q_T=QT_NEWTON, m=DEFAULT, keep_trajs=False):
    """Regularised Newton method with Armijo back-tracking for ONE problem (tg:298-398)."""
    x_ref = np.asarray(x_ref, dtype=float)
    u_ref = np.asarray(u_ref, dtype=float)
    if u_ref.shape[0] == x_ref.shape[0]:
        u_ref = u_ref[:-1]
    if u_ref.shape[0] != x_ref.shape[0] - 1:
        raise ValueError("Incompatible dimensions: x_ref has %d states but u_ref has %d controls"
                         % (x_ref.shape[0], u_ref.shape[0]))
    u = np.zeros_like(u_ref)
    x = simulate_open_loop(x0, u, m)
    cost_k = float(total_cost(x, u, x_ref, u_ref, Q, R, Q_T))
    hist = dict(cost=[cost_k], sigma_norm=[], delta_J=[], gamma=[], n_try=[], x_trajs=[x.copy()] if keep_trajs else [])
    K = sig = None
    status = STATUS_MAX_ITERS
    for k in range(max_iters):
        it = newton_iteration(x, u, cost_k, x_ref, u_ref, gamma_0, beta, c, Q, R, Q_T, m)
        K, sig = it["K"], it["sigma"]
        hist["sigma_norm"].append(it["sigma_norm"])
        hist["delta_J"].append(it["delta_J"])
        if it["accepted"] < 0:
            status = STATUS_LINE_SEARCH_FAILED
            break
        hist["gamma"].append(it["gammas"][it["accepted"]])
        hist["n_try"].append(it["accepted"] + 1)
        x, u, cost_k = it["x"], it["u"], it["cost"]
        hist["cost"].append(cost_k)
        if keep_trajs:
            hist["x_trajs"].append(x.copy())
        if it["sigma_norm"] < tol:
            status = STATUS_CONVERGED
            break
    hist["status"] = status
    hist["iters"] = len(hist["sigma_norm"])
    return x, u, K, sig, hist


def stepsize_sweep(x_traj, u_traj, K, sigma, x_ref, u_ref, steps, Q=Q_NEWTON, R=R_NEWTON, Q_T=QT_NEWTON, m=DEFAULT):
    """Cost along the search direction for each step size (numeric part of tg:257-264)."""
    out = np.zeros(len(steps))
    for i, s in enumerate(steps):
        xn, un = forward_closed_loop_update(x_traj, u_traj, K, sigma, s, m)
        out[i] = total_cost(xn, un, x_ref, u_ref, Q, R, Q_T)
    return out


# --------------------------------------------------------------------------
# T1-T5 (trajectory_tracking.py)
# --------------------------------------------------------------------------
def _inv2(G):
    return np.linalg.inv(G)


def solve_LQR_tracking(x_opt, u_opt, Q_reg=Q_LQR, R_reg=R_LQR, m=DEFAULT):
    """Time-varying LQR gains about (x_opt, u_opt), P_T = 2 Q_reg  (tt:170-203).  -> (..., N-1, 2, 4)"""
    x_opt = np.asarray(x_opt, dtype=float)
    Ad, Bd = linearize_discrete(x_opt[..., :-1, :], np.asarray(u_opt, dtype=float), m)
    T = Ad.shape[-3]
    lead = Ad.shape[:-3]
    P = np.broadcast_to(2.0 * np.asarray(Q_reg), lead + (4, 4)).copy()
    K = np.zeros(lead + (T, 2, 4))
    for t in range(T - 1, -1, -1):
        A = Ad[..., t, :, :]
        B = Bd[..., t, :, :]
        Bt = np.swapaxes(B, -1, -2)
        At = np.swapaxes(A, -1, -2)
        Kt = -_inv2(R_reg + Bt @ P @ B) @ (Bt @ P @ A)
        P = Q_reg + At @ P @ A + (At @ P @ B) @ Kt
        K[..., t, :, :] = Kt
    return K


def simulate_tracking(x_opt, u_opt, K_reg, x0_perturbed, m=DEFAULT):
    """u_t = u_opt_t + K_t (x_t - x_opt_t), x_{t+1} = dynamics(x_t, u_t)  (tt:206-216).

    x_opt/u_opt/K_reg are shared; x0_perturbed may carry leading batch axes.
    """
    x0 = np.asarray(x0_perturbed, dtype=float)
    N = x_opt.shape[0]
    xt = np.zeros(x0.shape[:-1] + (N, 4))
    ut = np.zeros(x0.shape[:-1] + (N - 1, 2))
    xt[..., 0, :] = x0
    with np.errstate(all="ignore"):
        for t in range(N - 1):
            ut[..., t, :] = u_opt[t] + _mv(K_reg[t], xt[..., t, :] - x_opt[t])
            xt[..., t + 1, :] = dynamics(xt[..., t, :], ut[..., t, :], m)
    return xt, ut


def LQR_tracking(x_ref, u_ref, t_ref=None, x0_perturbed=None, m=DEFAULT):
    """tt:219-249."""
    if x0_perturbed is None:
        x0_perturbed = x_ref[0].copy()
    K = solve_LQR_tracking(x_ref, u_ref, m=m)
    return simulate_tracking(x_ref, u_ref, K, x0_perturbed, m)


def compute_P_inf(A, B, Q, R, max_iter=1000, tol=1e-6, return_iters=False):
    """Fixed-point Riccati iteration from P = Q (tt:144-165)."""
    P = np.asarray(Q, dtype=float)
    for i in range(max_iter):
        Pp = P
        K = -_inv2(R + B.T @ P @ B) @ (B.T @ P @ A)
        P = Q + A.T @ P @ A + (A.T @ P @ B) @ K
        if np.abs(P - Pp).max() < tol:
            return (P, i + 1) if return_iters else P
    return (P, max_iter) if return_iters else P


def solver_mpc_kkt(x0, A_list, B_list, Q, R, Q_T, T_pred):
    """The equality-constrained QP of tt:80-117 solved as ONE dense KKT system.

    Variables X (4 x T_pred) and U (2 x T_pred); cost sum_{t<T_pred-1} x'Qx + u'Ru +
    x_{T-1}' Q_T x_{T-1}; constraints X0 = x0, X_{t+1} = A_t X_t + B_t U_t.  U[:, T_pred-1]
    appears in no cost or constraint (tt:83, 95-117): it is left at 0, IPOPT's initial
    guess.  Independent of the Riccati form below; used to cross-check it.
    """
    H = T_pred
    nxv, nuv = 4 * H, 2 * (H - 1)
    nz = nxv + nuv
    Hs = np.zeros((nz, nz))
    for t in range(H - 1):
        Hs[4 * t:4 * t + 4, 4 * t:4 * t + 4] = 2.0 * Q
        Hs[nxv + 2 * t:nxv + 2 * t + 2, nxv + 2 * t:nxv + 2 * t + 2] = 2.0 * R
    Hs[4 * (H - 1):4 * H, 4 * (H - 1):4 * H] = 2.0 * Q_T
    nc = 4 * H
    C = np.zeros((nc, nz))
    d = np.zeros(nc)
    C[0:4, 0:4] = np.eye(4)
    d[0:4] = x0
    for t in range(H - 1):
        rws = slice(4 * (t + 1), 4 * (t + 2))
        C[rws, 4 * (t + 1):4 * (t + 2)] = np.eye(4)
        C[rws, 4 * t:4 * t + 4] = -A_list[t]
        C[rws, nxv + 2 * t:nxv + 2 * t + 2] = -B_list[t]
    KKT = np.block([[Hs, C.T], [C, np.zeros((nc, nc))]])
    rhs = np.concatenate([np.zeros(nz), d])
    sol = np.linalg.solve(KKT, rhs)
    X = sol[:nxv].reshape(H, 4)
    U = np.zeros((H, 2))
    U[:H - 1] = sol[nxv:nz].reshape(H - 1, 2)
    return U[0].copy(), X, U


def mpc_gains(A_list, B_list, Q, R, Q_T, T_pred):
    """Backward Riccati for the QP of tt:80-117: P_{H-1} = Q_T, plain Q and R weights."""
    P = np.asarray(Q_T, dtype=float)
    Ks = [None] * (T_pred - 1)
    for j in range(T_pred - 2, -1, -1):
        A, B = A_list[j], B_list[j]
        K = -_inv2(R + B.T @ P @ B) @ (B.T @ P @ A)
        P = Q + A.T @ P @ A + (A.T @ P @ B) @ K
        Ks[j] = K
    return Ks


def solver_mpc(x0, A_list, B_list, Q, R, Q_T, T_pred, u_ref=None):
    """Riccati restatement of solver_mpc (tt:73-140): returns (U0, X_opt (T_pred,4), U_opt (T_pred,2))."""
    Ks = mpc_gains(A_list, B_list, Q, R, Q_T, T_pred)
    X = np.zeros((T_pred, 4))
    U = np.zeros((T_pred, 2))
    X[0] = x0
    for j in range(T_pred - 1):
        U[j] = Ks[j] @ X[j]
        X[j + 1] = A_list[j] @ X[j] + B_list[j] @ U[j]
    return U[0].copy(), X, U


def solver_mpc_box(x0, A_list, B_list, Q, R, Q_T, T_pred, u_ref, tau_max=18.0, return_info=False):
    """solver_mpc with the input constraints the reference keeps behind `test_constraints` (tt:87-91, 112-114):
    -tau_max <= U[:, t] + u_ref[t] <= tau_max for t < T_pred - 1.

    The reference hands this QP to IPOPT (tol 1e-6).  Here it is solved exactly: states eliminated (condensed QP in
    the inputs), primal active-set method with dense solves.  R must be diagonal; the first input does not act on
    the plant (B[:, 0] = 0, dynamics.py:153), so it is min R00 u0^2 on its interval, i.e. the point closest to 0.
    Returns (U0, X_opt (T_pred,4), U_opt (T_pred,2)); U_opt[T_pred-1] = 0 as in solver_mpc."""
    H = T_pred
    n = H - 1
    Q, R, Q_T = (np.asarray(M, dtype=float) for M in (Q, R, Q_T))
    assert R[0, 1] == 0.0 and R[1, 0] == 0.0, "solver_mpc_box: R must be diagonal"
    u_ref = np.asarray(u_ref, dtype=float)[:n]
    A = [np.asarray(a, dtype=float) for a in A_list[:n]]
    b = [np.asarray(Bm, dtype=float)[:, 1] for Bm in B_list[:n]]
    assert all(np.all(np.asarray(Bm)[:, 0] == 0.0) for Bm in B_list[:n]), "solver_mpc_box: B[:, 0] must vanish"
    x0 = np.asarray(x0, dtype=float)
    Phi = [np.eye(4)]
    for j in range(n):
        Phi.append(A[j] @ Phi[-1])
    G = np.zeros((H, 4, n))
    for j in range(n):
        col = b[j]
        G[j + 1, :, j] = col
        for t in range(j + 2, H):
            col = A[t - 1] @ col
            G[t, :, j] = col
    Hq = np.eye(n) * R[1, 1]
    f = np.zeros(n)
    for t in range(H):
        W = Q_T if t == H - 1 else Q
        Hq += G[t].T @ W @ G[t]
        f += G[t].T @ W @ (Phi[t] @ x0)
    lo, hi = -tau_max - u_ref[:, 1], tau_max - u_ref[:, 1]
    v = np.clip(np.zeros(n), lo, hi)
    act = np.where(v >= hi, 1, np.where(v <= lo, -1, 0))
    iters = 0
    for iters in range(1, 20 * n + 20):
        F = act == 0
        vA = np.where(act == 1, hi, np.where(act == -1, lo, 0.0))
        vn = vA.copy()
        if F.any():
            vn[F] = np.linalg.solve(Hq[np.ix_(F, F)], -(f[F] + Hq[np.ix_(F, ~F)] @ vA[~F]))
        dv = vn - v
        alpha, jb, sg = 1.0, -1, 0
        for j in np.where(F)[0]:
            if vn[j] > hi[j] and dv[j] > 0:
                a_ = (hi[j] - v[j]) / dv[j]
                if a_ < alpha:
                    alpha, jb, sg = a_, j, 1
            elif vn[j] < lo[j] and dv[j] < 0:
                a_ = (lo[j] - v[j]) / dv[j]
                if a_ < alpha:
                    alpha, jb, sg = a_, j, -1
        if jb >= 0:
            v = v + alpha * dv
            act[jb] = sg
            v[jb] = hi[jb] if sg == 1 else lo[jb]
            continue
        v = vn
        grad = 2.0 * (Hq @ v + f)
        scale = 2.0 * (np.abs(Hq) @ np.abs(v) + np.abs(f))
        wrong = ((act == 1) & (grad > 1e-10 * scale)) | ((act == -1) & (grad < -1e-10 * scale))
        if not wrong.any():
            break
        jw = np.where(wrong)[0][np.argmax((np.abs(grad) / scale)[wrong])]
        act[jw] = 0
    X = np.array([Phi[t] @ x0 + G[t] @ v for t in range(H)])
    U = np.zeros((H, 2))
    U[:n, 1] = v
    U[:n, 0] = np.clip(0.0, -tau_max - u_ref[:, 0], tau_max - u_ref[:, 0])
    if return_info:
        return U[0].copy(), X, U, {"active": act, "iterations": iters, "H": Hq, "f": f, "lo": lo, "hi": hi}
    return U[0].copy(), X, U


def solve_mpc_tracking_box(x0, x_ref, u_ref, T, T_pred=75, tau_max=18.0, Q=Q_MPC, R=R_MPC, m=DEFAULT, Q_T=None):
    """solve_mpc_tracking (tt:8-69) with the input box of tt:112-114 switched on, ONE problem (x0 (4,)):
    every step solves solver_mpc_box on the sliding window and applies u_ref[t] + U0 to the plant."""
    x_ref = np.asarray(x_ref, dtype=float)
    u_ref = np.asarray(u_ref, dtype=float)
    N = x_ref.shape[0]
    Ad, Bd = linearize_discrete(x_ref[:-1], u_ref, m)
    A_f, B_f = linearize_discrete(X_F, U_F, m)
    Q_T = compute_P_inf(A_f, B_f, Q, R) if Q_T is None else np.asarray(Q_T, dtype=float)
    xr = np.zeros((T, 4))
    ur = np.zeros((T - 1, 2))
    xr[0] = x0
    n_active = np.zeros(T - 1, dtype=int)
    for t in range(T - 1):
        Aw = [Ad[t + j] if t + j < N - 1 else A_f for j in range(T_pred - 1)]
        Bw = [Bd[t + j] if t + j < N - 1 else B_f for j in range(T_pred - 1)]
        uw = np.array([u_ref[t + j] if t + j < N - 1 else U_F for j in range(T_pred - 1)])
        xw = x_ref[t] if t < N else X_F
        U0, _, _, info = solver_mpc_box(xr[t] - xw, Aw, Bw, Q, R, Q_T, T_pred, uw, tau_max, return_info=True)
        n_active[t] = int((info["active"] != 0).sum())
        ur[t] = uw[0] + U0
        xr[t + 1] = dynamics(xr[t], ur[t], m)
    return xr, ur, n_active


def solve_mpc_tracking(x0, x_ref, u_ref, T, T_pred=75, Q=Q_MPC, R=R_MPC, m=DEFAULT, return_gains=False):
    """Receding-horizon tracking loop (tt:8-69) for ONE reference; x0 may carry batch axes.

    The window is a slice of the linearisation about the reference padded with
    (A_f, B_f) about x_f = [pi,0,0,0] (tt:33-35, 64-67); the first-move gain depends
    only on the window, so it is computed once per time step and applied to every x0.
    """
    x_ref = np.asarray(x_ref, dtype=float)
    u_ref = np.asarray(u_ref, dtype=float)
    N = x_ref.shape[0]
    Ad, Bd = linearize_discrete(x_ref[:-1], u_ref, m)
    A_f, B_f = linearize_discrete(X_F, U_F, m)
    Q_T = compute_P_inf(A_f, B_f, Q, R)
    x0 = np.asarray(x0, dtype=float)
    xr = np.zeros(x0.shape[:-1] + (N, 4))
    ur = np.zeros(x0.shape[:-1] + (N - 1, 2))
    xr[..., 0, :] = x0
    K0s = np.zeros((T - 1, 2, 4))
    for t in range(T - 1):
        Aw = [Ad[t + j] if t + j < N - 1 else A_f for j in range(T_pred - 1)]
        Bw = [Bd[t + j] if t + j < N - 1 else B_f for j in range(T_pred - 1)]
        K0 = mpc_gains(Aw, Bw, Q, R, Q_T, T_pred)[0]
        K0s[t] = K0
        xw = x_ref[t] if t < N else X_F
        uw = u_ref[t] if t < N - 1 else U_F
        ur[..., t, :] = uw + _mv(K0, xr[..., t, :] - xw)
        xr[..., t + 1, :] = dynamics(xr[..., t, :], ur[..., t, :], m)
    if return_gains:
        return xr, ur, K0s, Q_T
    return xr, ur
