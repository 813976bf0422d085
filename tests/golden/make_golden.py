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


# ------------------------------------------------------------------------------------------------------------
# Round 2: SURVEY 8(d) sample sizes.  The reference is a per-problem Python loop (3-5 Newton iterations/s/core), so
# these run in a fork()ed pool (the SymPy model is built once in the parent).  --workers sets the pool size.
# ------------------------------------------------------------------------------------------------------------
WORKERS = 6


def _pool_map(fn, items):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(WORKERS) as pool:
        return pool.map(fn, items, chunksize=1)


def c2_x0():
    """The config-2 batch of SURVEY 8(d): seed 1, problem 0 = task_2."""
    x0 = np.random.default_rng(1).uniform(-0.2, 0.2, (4096, 4))
    x0[0] = 0.0
    return x0


# first / last lane of the first / last tile, tile boundaries, and a spread over the batch
C2_CONV_ROWS = [1, 31, 32, 63, 777, 1024, 2047, 2048, 2500, 3000, 3333, 3840, 4000, 4064, 4094, 4095]


def _c2_conv_one(i):
    x_ref, u_ref, _ = fully_actuated_ref()
    t = time.time()
    o = run_newton(c2_x0()[i], x_ref, u_ref, 5000, 1e-4, 0.1, keep=0)
    print("  c2 row %d: %d iterations, %.0f s" % (i, len(o["sigma_norm"]), time.time() - t), flush=True)
    return o


def g_c2conv():
    """16 problems of the config-2 batch to convergence (gamma_0 = 0.1, tol = 1e-4) with the unmodified reference."""
    outs = _pool_map(_c2_conv_one, C2_CONV_ROWS)
    n = max(len(o["cost"]) for o in outs)

    def pad(v, m):
        return np.concatenate([np.asarray(v, dtype=float), np.full(m - len(v), np.nan)])

    save("newton_c2_converged", rows=np.array(C2_CONV_ROWS), x0=c2_x0()[C2_CONV_ROWS],
         iters=np.array([len(o["sigma_norm"]) for o in outs]),
         cost=np.array([pad(o["cost"], n) for o in outs]), sigma_norm=np.array([pad(o["sigma_norm"], n - 1) for o in outs]),
         n_try=np.array([pad(o["n_try"], n - 1) for o in outs]), gamma_acc=np.array([pad(o["gamma_acc"], n - 1) for o in outs]),
         x=np.array([o["x"] for o in outs]), u=np.array([o["u"] for o in outs]), sigma=np.array([o["sigma"] for o in outs]),
         K=np.array([o["K"] for o in outs]), x_prev=np.array([o["x_prev"] for o in outs]),
         u_prev=np.array([o["u_prev"] for o in outs]))


def c2_3it_rows():
    r = np.random.default_rng(11).choice(np.arange(64, 4032), 56, replace=False)
    return np.sort(np.concatenate([[0, 31, 32, 63, 4032, 4063, 4064, 4095], r]))


def _c2_3it_one(args):
    i, gamma_0 = args
    x_ref, u_ref, _ = fully_actuated_ref()
    o = run_newton(c2_x0()[i], x_ref, u_ref, 3, 1e-4, gamma_0, keep=0)
    return o


def g_c2three():
    """64 problems of the config-2 batch x 3 iterations, at gamma_0 = 0.1 and gamma_0 = 1 (back-tracking)."""
    rows = c2_3it_rows()
    arrs = dict(rows=rows, x0=c2_x0()[rows])
    for tag, g0 in (("g01", 0.1), ("g1", 1.0)):
        outs = _pool_map(_c2_3it_one, [(int(i), g0) for i in rows])
        arrs.update({tag + "_x": np.array([o["x"] for o in outs]), tag + "_u": np.array([o["u"] for o in outs]),
                     tag + "_sigma": np.array([o["sigma"] for o in outs]), tag + "_cost": np.array([o["cost"] for o in outs]),
                     tag + "_sigma_norm": np.array([o["sigma_norm"] for o in outs]),
                     tag + "_n_try": np.array([o["n_try"] for o in outs]),
                     tag + "_gamma_acc": np.array([o["gamma_acc"] for o in outs]),
                     # K of 8 of them (the gains do not compress: 32 KB per problem)
                     tag + "_K8": np.array([o["K"] for o in outs[:8]])})
    save("newton_c2_three_iters", **arrs)


def c3_x0():
    """The config-3 batch of SURVEY 8(d): seed 2, problems 0, 1 = the perturbations of main.py:104-110."""
    with ref_import.in_ref_dir():
        d = np.load("trajectories_npz/acrobot_optimal_trajectory.npz")
    x0 = d["x"][0] + np.random.default_rng(2).uniform(-0.3, 0.3, (65536, 4))
    x0[0] = d["x"][0] + 0.2
    x0[1] = d["x"][0] + 0.3
    return x0, d["x"], d["u"]


def c3_rows():
    r = np.random.default_rng(12).choice(np.arange(32, 65504), 248, replace=False)
    return np.sort(np.concatenate([[0, 1, 30, 31, 65504, 65505, 65534, 65535], r]))


_C3 = {}


def _c3_one(i):
    x0, x_opt, u_opt = _C3["x0"], _C3["x_opt"], _C3["u_opt"]
    with np.errstate(all="ignore"):
        xt, ut = rtt.simulate_tracking(x_opt, u_opt, _C3["K"], x0[i])
    return xt[::10], ut[::10], xt.sum(axis=0), ut.sum(axis=0)


def g_lqr256():
    """256 problems of the config-3 batch: tracked rollouts sampled every 10th step + column sums over time."""
    x0, x_opt, u_opt = c3_x0()
    _C3.update(x0=x0, x_opt=x_opt, u_opt=u_opt, K=rtt.solve_LQR_tracking(x_opt, u_opt))
    rows = c3_rows()
    res = _pool_map(_c3_one, [int(i) for i in rows])
    save("lqr_tracking_c3", rows=rows, x0=x0[rows], x_track_10=np.array([r[0] for r in res]),
         u_track_10=np.array([r[1] for r in res]), x_sum=np.array([r[2] for r in res]), u_sum=np.array([r[3] for r in res]))


SWEEP_GAMMA_IDX = np.array([0, 7, 16, 40, 80, 120, 159, 199])  # of linspace(0, 1.25, 200)  (tg:257-258)


def c5_rows():
    return np.sort(np.random.default_rng(13).choice(4096, 256, replace=False))


def _c5_one(args):
    p, k = args
    x_ref, u_ref, _ = fully_actuated_ref()
    u_ref = trim(x_ref, u_ref)
    x0 = c2_x0()[p]
    if k == 0:
        u = np.zeros_like(u_ref)
        x = rtg.simulate_open_loop(x0, u)
    else:
        with CallLog() as log:
            x, u, _, _, h = rtg.newton_Algorithm(x0, x_ref, u_ref, max_iters=k, tol=1e-4, gamma_0=0.1, plot_armijo_iters=0)
        assert len(h["sigma_norm"]) == k
    lists = rtg.build_stage_lists(x, u, x_ref, u_ref, [None] * len(x))
    K, sig, dJ = rtg.calculate_K_and_sigma(*lists)
    steps = np.linspace(0, 1.25, 200)[SWEEP_GAMMA_IDX]
    costs = np.zeros(len(steps))
    for i, s in enumerate(steps):
        xn, un = rtg.forward_closed_loop_update(x, u, K, sig, gamma=s)
        costs[i] = rtg.total_cost(xn, un, x_ref, u_ref, rtg.Q, rtg.R, rtg.Q_T)
    return costs, float(dJ), x[-1].copy(), float(np.max(np.abs(np.array(sig))))


def g_sweep256():
    """Config 5: 256 base iterates = Newton iterate k_p (k_p = p mod 50, gamma_0 = 0.1) of config-2 problem p, each
    with the cost along the search direction at 8 of the 200 step sizes of tg:257-258."""
    rows = c5_rows()
    ks = rows % 50
    res = _pool_map(_c5_one, [(int(p), int(k)) for p, k in zip(rows, ks)])
    save("sweep_c5", rows=rows, k=ks, gamma_idx=SWEEP_GAMMA_IDX, steps=np.linspace(0, 1.25, 200)[SWEEP_GAMMA_IDX],
         costs=np.array([r[0] for r in res]), delta_J=np.array([r[1] for r in res]), x_T=np.array([r[2] for r in res]),
         sigma_norm=np.array([r[3] for r in res]))


ALL_R2 = dict(c2conv=g_c2conv, c2three=g_c2three, lqr256=g_lqr256, sweep256=g_sweep256)


ALL = dict(shipped=g_shipped, dyn=g_dyn, blocks=g_blocks, pinf=g_pinf, lqr=g_lqr, gamma1=g_gamma1, c2=g_c2, sweep=g_sweep,
           task1=g_task1, task2=g_task2)

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--workers", type=int, default=WORKERS)
    a = ap.parse_args()
    WORKERS = a.workers
    ALL.update(ALL_R2)
    for k in (a.only or list(ALL)):
        t = time.time()
        ALL[k]()
        print("[%s] %.1f s" % (k, time.time() - t), flush=True)
