"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (``/root/reference`` present):

    python tests/golden/make_golden.py [--only NAME ...]

Every array written here is the output of a reference function (dynamics.py,
trajectory_generation.py, trajectory_tracking.py) imported through
``oracle/ref_import.py`` (matplotlib / casadi stubbed, plot function no-op'ed),
or a bit-for-bit copy of the arrays inside the two trajectory files the reference
ships.  Nothing from the oracle restatement or the CUDA path is used.
The MPC QP needs CasADi/IPOPT and cannot be run: no MPC fixture is made.
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from oracle import ref_import  # noqa: E402

rd, rtg, rtt = ref_import.load()


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, {k: np.asarray(v).shape for k, v in arrs.items()}, flush=True)


class CallLog:
    """Records gamma of every forward_closed_loop_update call and every total_cost value."""

    def __init__(self):
        self.gammas, self.costs, self.u_new = [], [], []
        self._f, self._c = rtg.forward_closed_loop_update, rtg.total_cost

    def __enter__(self):
        def f(x, u, K, s, gamma=1.0):
            self.gammas.append(float(gamma))
            out = self._f(x, u, K, s, gamma=gamma)
            self.u_new.append(out[1].copy())
            return out

        def c(*a, **k):
            v = self._c(*a, **k)
            self.costs.append(float(v))
            return v

        rtg.forward_closed_loop_update, rtg.total_cost = f, c
        return self

    def __exit__(self, *a):
        rtg.forward_closed_loop_update, rtg.total_cost = self._f, self._c


def fully_actuated_ref():
    with ref_import.in_ref_dir():
        return rtg.get_fully_actuated_ref()


def g_shipped():
    with ref_import.in_ref_dir():
        a = np.load("trajectories_npz/fully_actuated_trajectory.npz")
        save("fully_actuated_trajectory", **{k: a[k] for k in a.files})
        b = np.load("trajectories_npz/acrobot_optimal_trajectory.npz")
        save("acrobot_optimal_trajectory", **{k: b[k] for k in b.files})


def g_dyn():
    rng = np.random.default_rng(7)
    n = 256
    x = np.concatenate([rng.uniform(-4, 4, (n, 2)), rng.uniform(-10, 10, (n, 2))], axis=1)
    u = rng.uniform(-25, 25, (n, 2))
    x[0] = [0.3, -0.2, 0.5, -0.7]
    u[0] = [1.0, 2.0]
    x[1] = [np.pi, 0, 0, 0]
    u[1] = 0
    x[2] = 0
    u[2] = 0
    f = np.array([rd.continuous_dynamics(x[i], u[i]) for i in range(n)])
    s = np.array([rd.dynamics(x[i], u[i]) for i in range(n)])
    AB = [rd.Calculate_A_B_matrixes(x[i], u[i]) for i in range(n)]
    save("dyn_kat", x=x, u=u, f=f, step=s, A_c=np.array([a for a, _ in AB]), B_c=np.array([b for _, b in AB]))


def run_newton(x0, x_ref, u_ref, max_iters, tol, gamma_0, keep=4):
    with CallLog() as log:
        x, u, K, sig, h = rtg.newton_Algorithm(x0, x_ref, u_ref, max_iters=max_iters, tol=tol, gamma_0=gamma_0,
                                               plot_armijo_iters=0)
    # reconstruct per-iteration tries: gammas restart at gamma_0 at each iteration
    n_try, gam_acc = [], []
    i = 0
    g = log.gammas
    while i < len(g):
        j = i + 1
        while j < len(g) and g[j] != gamma_0:
            j += 1
        n_try.append(j - i)
        gam_acc.append(g[j - 1])
        i = j
    acc_calls = np.cumsum(n_try) - 1
    u_trajs = [np.zeros_like(u)] + [log.u_new[c] for c in acc_calls[:keep]]
    # the iterate the returned K / sigma were computed on (tg:338 runs before the update at tg:384)
    n_acc = len(h["cost"]) - 1
    x_prev = h["x_trajs"][n_acc - 1]
    u_prev = log.u_new[acc_calls[n_acc - 2]] if n_acc >= 2 else np.zeros_like(u)
    return dict(u_trajs=np.array(u_trajs), x_prev=x_prev, u_prev=u_prev, x=x, u=u, K=np.array(K), sigma=np.array(sig), cost=np.array(h["cost"]),
                sigma_norm=np.array(h["sigma_norm"]), x_trajs=np.array(h["x_trajs"][:keep + 1]),
                sigmas=np.array([np.array(s) for s in h["sigmas"][:keep]]),
                cand_gammas=np.array(log.gammas), cand_costs=np.array(log.costs[1:]),
                n_try=np.array(n_try), gamma_acc=np.array(gam_acc))


def first_iteration_blocks(x_traj, u_traj, x_ref, u_ref):
    lam = rtg.compute_costate_trajectory(x_traj, u_traj, x_ref, u_ref)
    lists = rtg.build_stage_lists(x_traj, u_traj, x_ref, u_ref, lam)
    K, sig, dJ = rtg.calculate_K_and_sigma(*lists)
    return dict(lam=np.array(lam), A_list=np.array(lists[0]), B_list=np.array(lists[1]), q_list=np.array(lists[5]),
                r_list=np.array(lists[6]), Q_T_block=lists[7], q_T=lists[8], K0=np.array(K), sigma0=np.array(sig),
                delta_J0=float(dJ))


def g_task2():
    x_ref, u_ref, t_ref = fully_actuated_ref()
    x0 = np.array([0.0, 0, 0, 0])
    t = time.time()
    out = run_newton(x0, x_ref, u_ref, 5000, 1e-4, 0.1)
    print("task2 solve", time.time() - t, "s, iters", len(out["sigma_norm"]))
    save("newton_task2", x0=x0, x_ref=x_ref, u_ref=u_ref, t_ref=t_ref, **out)


def trim(x_ref, u_ref):
    return u_ref[:-1] if u_ref.shape[0] == x_ref.shape[0] else u_ref  # tg:301-303


def g_blocks():
    """Stage lists, costate and Riccati outputs of the FIRST Newton iteration of task_2."""
    x_ref, u_ref, _ = fully_actuated_ref()
    u_ref = trim(x_ref, u_ref)
    x0 = np.zeros(4)
    u0 = np.zeros_like(u_ref)
    xo = rtg.simulate_open_loop(x0, u0)
    save("newton_task2_blocks", x0=x0, x_open=xo, **first_iteration_blocks(xo, u0, x_ref, u_ref))


def g_gamma1():
    x_ref, u_ref, _ = fully_actuated_ref()
    x0 = np.array([0.0, 0, 0, 0])
    out = run_newton(x0, x_ref, u_ref, 14, 1e-4, 1.0, keep=14)
    save("newton_gamma1", x0=x0, **out)


def g_task1():
    u_t1, u_t2 = np.array([0.0, 0.0]), np.array([0.5, 0.5])
    x_e1, u_e1 = rtg.compute_equilibrium(u_t1, (0.1, -0.1))
    x_e2, u_e2 = rtg.compute_equilibrium(u_t2, (0.35, -0.35))
    t_ref, x_ref, u_ref = rtg.define_reference_piecewise(10.0, x_e1, x_e2, u_e1, u_e2)
    t = time.time()
    out = run_newton(x_e1.copy(), x_ref, u_ref, 5000, 1e-4, 0.05)
    print("task1 solve", time.time() - t, "s, iters", len(out["sigma_norm"]))
    save("newton_task1", x0=x_e1, x_e1=x_e1, x_e2=x_e2, u_e1=u_e1, u_e2=u_e2, x_ref=x_ref, u_ref=u_ref, t_ref=t_ref, **out)


def g_c2():
    """Rows 1-3 of the config-2 batch, 6 iterations each."""
    x_ref, u_ref, _ = fully_actuated_ref()
    x0s = np.random.default_rng(1).uniform(-0.2, 0.2, (4096, 4))
    outs = [run_newton(x0s[i], x_ref, u_ref, 6, 1e-4, 0.1, keep=6) for i in (1, 2, 3)]
    save("newton_c2_rows", rows=np.array([1, 2, 3]), x0=x0s[1:4],
         **{k: np.array([o[k] for o in outs]) for k in ("x", "u", "K", "sigma", "cost", "sigma_norm", "n_try", "gamma_acc")})


def g_lqr():
    with ref_import.in_ref_dir():
        d = np.load("trajectories_npz/acrobot_optimal_trajectory.npz")
    x_opt, u_opt = d["x"], d["u"]
    K = np.array(rtt.solve_LQR_tracking(x_opt, u_opt))
    rng = np.random.default_rng(2)
    x0s = x_opt[0] + rng.uniform(-0.3, 0.3, (24, 4))
    x0s[0] = x_opt[0] + 0.2
    x0s[1] = x_opt[0] + 0.3
    # a few large perturbations: some of these rollouts overflow to non-finite values
    x0s[20:] = x_opt[0] + np.random.default_rng(5).normal(0, 1.5, (4, 4))
    xs, us = [], []
    with np.errstate(all="ignore"):
        for x0 in x0s:
            xt, ut = rtt.simulate_tracking(x_opt, u_opt, list(K), x0)
            xs.append(xt)
            us.append(ut)
    save("lqr_tracking", K_reg=K, x0=x0s, x_track=np.array(xs), u_track=np.array(us))


def g_pinf():
    x_f, u_f = np.array([np.pi, 0, 0, 0]), np.array([0.0, 0.0])
    A_f, B_f = rtg.discretize_linearization(*rd.Calculate_A_B_matrixes(x_f, u_f), 2e-2)
    Q = np.diag([120.0, 100.0, 0.0001, 0.0001])
    R = np.diag([1e-6, 10.0])
    P = rtt.compute_P_inf(A_f, B_f, Q, R)
    # second case: LQR weights about the hanging equilibrium
    A0, B0 = rtg.discretize_linearization(*rd.Calculate_A_B_matrixes(np.zeros(4), np.zeros(2)), 2e-2)
    P0 = rtt.compute_P_inf(A0, B0, np.diag([100.0, 100.0, 10.0, 10.0]), np.eye(2))
    save("p_inf", A_f=A_f, B_f=B_f, Q=Q, R=R, P_inf=P, A0=A0, B0=B0, P0=P0)


def g_sweep():
    x_ref, u_ref, _ = fully_actuated_ref()
    u_ref = trim(x_ref, u_ref)
    x0 = np.zeros(4)
    u = np.zeros_like(u_ref)
    x = rtg.simulate_open_loop(x0, u)
    lam = [None] * len(x)
    lists = rtg.build_stage_lists(x, u, x_ref, u_ref, lam)
    K, sig, dJ = rtg.calculate_K_and_sigma(*lists)
    steps = np.linspace(0, 1.25, 200)
    costs = np.zeros(200)
    for i, s in enumerate(steps):
        xn, un = rtg.forward_closed_loop_update(x, u, K, sig, gamma=s)
        costs[i] = rtg.total_cost(xn, un, x_ref, u_ref, rtg.Q, rtg.R, rtg.Q_T)
    save("sweep_iter0", steps=steps, costs=costs, delta_J=float(dJ))


ALL = dict(shipped=g_shipped, dyn=g_dyn, blocks=g_blocks, pinf=g_pinf, lqr=g_lqr, gamma1=g_gamma1, c2=g_c2, sweep=g_sweep,
           task1=g_task1, task2=g_task2)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    a = ap.parse_args()
    for k in (a.only or list(ALL)):
        t = time.time()
        ALL[k]()
        print("[%s] %.1f s" % (k, time.time() - t), flush=True)
