"""Pin the NumPy oracle (oracle/acro_oracle.py) against the reference's own outputs.

The fixtures were produced by tests/golden/make_golden.py from the unmodified reference
(or are the arrays of the trajectory files it ships).  CPU only.
"""
import numpy as np
import pytest

from conftest import golden, rel_err
from oracle import acro_oracle as O

TOL = 1e-9  # contract of BASELINE.json north_star; observed errors are ~1e-13


def test_dynamics_kat():
    g = golden("dyn_kat")
    assert rel_err(O.continuous_dynamics(g["x"], g["u"]), g["f"]) < 1e-12
    assert rel_err(O.dynamics(g["x"], g["u"]), g["step"]) < 1e-12
    A, B = O.Calculate_A_B_matrixes(g["x"], g["u"])
    assert rel_err(A, g["A_c"]) < 1e-12
    assert rel_err(B, g["B_c"]) < 1e-12
    assert np.all(B[:, :, 0] == 0.0) and np.all(B[:, :2, :] == 0.0)


def test_survey_known_answers():
    x = np.array([0.3, -0.2, 0.5, -0.7])
    u = np.array([1.0, 2.0])
    np.testing.assert_allclose(O.continuous_dynamics(x, u), [0.5, -0.7, -8.097062412610704, 18.791857037341483], rtol=1e-13)
    np.testing.assert_allclose(O.dynamics(x, u), [0.30842243714633577, -0.21036198377515106, 0.34430441728616834,
                                                  -0.34211929152559867], rtol=1e-13)


def test_shipped_optimal_trajectory_satisfies_rk4_step():
    """500 known-answer vectors: x[k+1] == dynamics(x[k], u[k]) on the shipped file."""
    d = golden("acrobot_optimal_trajectory")
    nxt = O.dynamics(d["x"][:-1], d["u"])
    assert np.max(np.abs(nxt - d["x"][1:])) < 5e-15


def test_fully_actuated_file_satisfies_actuated_step():
    d = golden("fully_actuated_trajectory")
    m = O.Model(actuated_tau1=True)
    nxt = O.dynamics(d["x"][:-1], d["u"], m)
    assert np.max(np.abs(nxt - d["x"][1:])) < 1e-9


def test_first_iteration_blocks(fa_ref):
    g = golden("newton_task2_blocks")
    x_ref, u_ref, _ = fa_ref
    assert u_ref.shape[0] == x_ref.shape[0] - 1
    u0 = np.zeros_like(u_ref)
    xo = O.simulate_open_loop(g["x0"], u0)
    assert rel_err(xo, g["x_open"]) < 1e-12
    Ad, Bd, q, r, QT2, qT = O.build_stage_lists(xo, u0, x_ref, u_ref)
    assert rel_err(Ad, g["A_list"]) < 1e-12 and rel_err(Bd, g["B_list"]) < 1e-12
    assert rel_err(q, g["q_list"]) < 1e-12 and rel_err(r, g["r_list"]) < 1e-12
    assert rel_err(qT, g["q_T"]) < 1e-12 and rel_err(QT2, g["Q_T_block"]) == 0
    K, s, dJ = O.calculate_K_and_sigma(Ad, Bd, q, r, 2 * O.Q_NEWTON, 2 * O.R_NEWTON, QT2, qT)
    assert rel_err(K, g["K0"]) < 1e-11 and rel_err(s, g["sigma0"]) < 1e-11
    assert abs(dJ - g["delta_J0"]) < 1e-11 * abs(g["delta_J0"])
    lam = O.compute_costate_trajectory(xo, u0, x_ref, u_ref)
    assert rel_err(lam, g["lam"]) < 1e-11
    np.testing.assert_allclose(g["delta_J0"], -169165.70417990963, rtol=1e-12)


def test_newton_task2_first_iterations_and_costs(fa_ref):
    g = golden("newton_task2")
    x_ref, u_ref, _ = fa_ref
    x, u, K, s, h = O.newton_Algorithm(g["x0"], x_ref, u_ref, max_iters=4, tol=1e-4, gamma_0=0.1, keep_trajs=True)
    assert rel_err(h["cost"], g["cost"][:5]) < 1e-12
    assert rel_err(h["sigma_norm"], g["sigma_norm"][:4]) < 1e-12
    assert rel_err(np.array(h["x_trajs"]), g["x_trajs"][:5]) < 1e-11
    assert h["n_try"] == list(g["n_try"][:4])
    np.testing.assert_allclose(g["cost"][:3], [407310.76108107425, 391453.8670466373, 378640.1632285486], rtol=1e-13)
    assert len(g["sigma_norm"]) == 393 and abs(g["cost"][-1] - 28063.21834988143) < 1e-6


def test_newton_task2_golden_matches_shipped_file():
    """The reference re-run here reproduces the file it ships (2.3e-13 / 6.0e-13)."""
    g = golden("newton_task2")
    d = golden("acrobot_optimal_trajectory")
    assert np.max(np.abs(g["x"] - d["x"])) < 1e-11 and np.max(np.abs(g["u"] - d["u"])) < 1e-11


def test_newton_backtracking_gamma1(fa_ref):
    """gamma_0 = 1 exercises the Armijo back-tracking: same tries, same accepted step sizes."""
    g = golden("newton_gamma1")
    x_ref, u_ref, _ = fa_ref
    x, u, K, s, h = O.newton_Algorithm(g["x0"], x_ref, u_ref, max_iters=8, tol=1e-4, gamma_0=1.0, keep_trajs=True)
    assert h["n_try"] == list(g["n_try"][:8])
    assert h["gamma"] == list(g["gamma_acc"][:8])  # bit-identical step sizes (sequential gamma *= beta)
    assert rel_err(h["cost"], g["cost"][:9]) < 1e-10
    assert rel_err(np.array(h["x_trajs"]), g["x_trajs"][:9]) < 1e-9
    assert max(g["n_try"]) > 1


def test_newton_task1_recipe():
    g = golden("newton_task1")
    x, u, K, s, h = O.newton_Algorithm(g["x0"], g["x_ref"], g["u_ref"], max_iters=5, tol=1e-4, gamma_0=0.05, keep_trajs=True)
    assert rel_err(h["cost"], g["cost"][:6]) < 1e-12
    assert rel_err(np.array(h["x_trajs"])[:5], g["x_trajs"][:5]) < 1e-11
    assert len(g["sigma_norm"]) == 173 and abs(g["cost"][-1] - 27.48962661922637) < 1e-9


def test_newton_c2_rows(fa_ref):
    g = golden("newton_c2_rows")
    x_ref, u_ref, _ = fa_ref
    for i in range(2):
        x, u, K, s, h = O.newton_Algorithm(g["x0"][i], x_ref, u_ref, max_iters=6, tol=1e-4, gamma_0=0.1)
        assert rel_err(x, g["x"][i]) < 1e-11 and rel_err(u, g["u"][i]) < 1e-11
        assert rel_err(K, g["K"][i]) < 1e-10 and rel_err(s, g["sigma"][i]) < 1e-10
        assert rel_err(h["cost"], g["cost"][i]) < 1e-12


def test_sweep(fa_ref):
    g = golden("sweep_iter0")
    t2 = golden("newton_task2_blocks")
    x_ref, u_ref, _ = fa_ref
    idx = np.arange(0, 200, 9)
    c = O.stepsize_sweep(t2["x_open"], np.zeros_like(u_ref), t2["K0"], t2["sigma0"], x_ref, u_ref, g["steps"][idx])
    assert rel_err(c, g["costs"][idx]) < 1e-11


def test_lqr_gains_and_tracking():
    g = golden("lqr_tracking")
    d = golden("acrobot_optimal_trajectory")
    K = O.solve_LQR_tracking(d["x"], d["u"])
    assert rel_err(K, g["K_reg"]) < 1e-11
    np.testing.assert_allclose(K[0][1], [1.6085004079175165, -5.314848245482309, -4.725771848391788, -4.1000147935985405], rtol=1e-10)
    xt, ut = O.simulate_tracking(d["x"], d["u"], g["K_reg"], g["x0"])
    ok = np.isfinite(g["x_track"]).all(axis=(1, 2)) & (np.abs(g["x_track"]).max(axis=(1, 2)) < 50)
    assert ok[:20].all()
    assert rel_err(xt[ok], g["x_track"][ok]) < TOL and rel_err(ut[ok], g["u_track"][ok]) < TOL
    np.testing.assert_allclose(xt[0, -1], [3.141507366148879, 3.088315439801646e-05, -2.0527577096561255e-04,
                                           6.888578496876233e-04], atol=1e-10)


def test_p_inf():
    g = golden("p_inf")
    P, it = O.compute_P_inf(g["A_f"], g["B_f"], g["Q"], g["R"], return_iters=True)
    assert it == 434
    assert rel_err(P, g["P_inf"]) < 1e-12
    A_f, B_f = O.linearize_discrete(O.X_F, O.U_F)
    assert rel_err(A_f, g["A_f"]) < 1e-13 and rel_err(B_f, g["B_f"]) < 1e-13
    P0 = O.compute_P_inf(g["A0"], g["B0"], O.Q_LQR, O.R_LQR)
    assert rel_err(P0, g["P0"]) < 1e-9


def test_mpc_riccati_equals_kkt():
    """The QP of solver_mpc (trajectory_tracking.py:80-117): Riccati form == dense KKT solve (unpinned vs IPOPT)."""
    d = golden("acrobot_optimal_trajectory")
    g = golden("p_inf")
    Ad, Bd = O.linearize_discrete(d["x"][:-1], d["u"])
    for H, t0 in ((75, 0), (50, 100), (20, 490)):
        A_f, B_f = g["A_f"], g["B_f"]
        Aw = [Ad[t0 + j] if t0 + j < 500 else A_f for j in range(H - 1)]
        Bw = [Bd[t0 + j] if t0 + j < 500 else B_f for j in range(H - 1)]
        x0 = 0.1 * np.ones(4)
        U0, X, U = O.solver_mpc(x0, Aw, Bw, O.Q_MPC, O.R_MPC, g["P_inf"], H)
        U0k, Xk, Uk = O.solver_mpc_kkt(x0, Aw, Bw, O.Q_MPC, O.R_MPC, g["P_inf"], H)
        assert rel_err(U0, U0k) < 1e-8 and rel_err(X, Xk) < 1e-8 and rel_err(U, Uk) < 1e-8
        if t0 == 0:
            # figures/mpc/tracking_dx_0.1_err.png: initial control error ~0.81 (-0.80644713)
            assert abs(U0[1] + 0.80644713) < 1e-6


def test_gain_conditioning(fa_ref):
    """Evidence for the tolerance used on FREE-RUNNING gains: near convergence the Riccati sweep is ill-conditioned.
    The same oracle code on iterates that differ by 1e-13 gives gains that differ by far more than 1e-9 |K|_inf,
    while sigma stays within 1e-9 (hard parts)."""
    g = golden("newton_task2")
    x_ref, u_ref, _ = fa_ref

    def ks(x, u):
        Ad, Bd, q, r, QT2, qT = O.build_stage_lists(x, u, x_ref, u_ref)
        return O.calculate_K_and_sigma(Ad, Bd, q, r, 2 * O.Q_NEWTON, 2 * O.R_NEWTON, QT2, qT)[:2]

    rng = np.random.default_rng(0)
    K0, S0 = ks(g["x"], g["u"])
    worst = 0.0
    for _ in range(3):
        K1, S1 = ks(g["x"] + 1e-13 * rng.normal(size=g["x"].shape), g["u"] + 1e-13 * rng.normal(size=g["u"].shape))
        worst = max(worst, rel_err(K1, K0))
        assert rel_err(S1, S0) < 1e-9
    assert np.abs(K0).max() > 400
    assert 1e-10 < worst < 1e-6


def test_reference_gains_carry_rounding_noise_near_convergence(fa_ref):
    """The reference's own final-iteration gains differ from an 80-bit evaluation of the same recursion by ~1.6e-8
    |K|_inf (and the float64 oracle by ~6e-8), while at iteration 0 everything agrees to 1e-13: this is what the GPU
    parity test's gain criterion (tests/test_gpu_parity.py::assert_gain_parity) rests on."""
    from extended import riccati_ld
    if np.finfo(np.longdouble).eps > 1e-18:
        pytest.skip("no extended precision long double on this host")
    x_ref, u_ref, _ = fa_ref
    g = golden("newton_task2")
    K, S = riccati_ld(g["x_prev"], g["u_prev"], x_ref, u_ref)
    e = rel_err(g["K"], K.astype(float))
    assert 1e-9 < e < 1e-6
    assert rel_err(g["sigma"], S.astype(float)) < 1e-10
    b = golden("newton_task2_blocks")
    K0, S0 = riccati_ld(b["x_open"], np.zeros((500, 2)), x_ref, u_ref)
    assert rel_err(b["K0"], K0.astype(float)) < 1e-12


def test_box_constrained_mpc_oracle_against_scipy_bvls():
    """solver_mpc_box (dense primal active set) against SciPy's bounded-variable least squares on the same condensed
    QP: min v'Hv + 2f'v = |L'v + L^-1 f|^2, lo <= v <= hi.  Windows of the shipped trajectory where the box binds."""
    from scipy.optimize import lsq_linear
    d = golden("acrobot_optimal_trajectory")
    g = golden("p_inf")
    x_ref, u_ref = d["x"], d["u"]
    N = 501
    Ad, Bd = O.linearize_discrete(x_ref[:-1], u_ref)
    rng = np.random.default_rng(1)
    n_bound = 0
    for t0, H, tau in ((170, 40, 18.0), (227, 50, 12.0), (60, 30, 18.0), (410, 75, 18.0), (300, 20, 8.0)):
        Aw = [Ad[t0 + j] if t0 + j < N - 1 else g["A_f"] for j in range(H - 1)]
        Bw = [Bd[t0 + j] if t0 + j < N - 1 else g["B_f"] for j in range(H - 1)]
        uw = np.array([u_ref[t0 + j] if t0 + j < N - 1 else O.U_F for j in range(H - 1)])
        x0 = rng.uniform(-0.1, 0.1, 4)
        U0, X, U, info = O.solver_mpc_box(x0, Aw, Bw, O.Q_MPC, O.R_MPC, g["P_inf"], H, uw, tau, return_info=True)
        L = np.linalg.cholesky(info["H"])
        r = lsq_linear(L.T, -np.linalg.solve(L, info["f"]), bounds=(info["lo"], info["hi"]), method="bvls", tol=1e-15,
                       max_iter=5000)
        assert np.abs(r.x - U[:H - 1, 1]).max() <= 1e-9 * max(1.0, np.abs(r.x).max())
        assert (np.abs(U[:H - 1] + uw) <= tau * (1 + 1e-12)).all()
        # the states are those of the window dynamics under U
        x = x0.copy()
        for j in range(H - 1):
            assert np.abs(X[j] - x).max() < 1e-9 * max(1.0, np.abs(x).max())
            x = Aw[j] @ x + Bw[j] @ U[j]
        n_bound += int((info["active"] != 0).sum())
        # a box that cannot bind gives solver_mpc
        U0u, Xu, Uu = O.solver_mpc(x0, Aw, Bw, O.Q_MPC, O.R_MPC, g["P_inf"], H)
        U0b, Xb, Ub = O.solver_mpc_box(x0, Aw, Bw, O.Q_MPC, O.R_MPC, g["P_inf"], H, uw, 1e9)
        assert rel_err(Ub, Uu) < 1e-8 and rel_err(Xb, Xu) < 1e-8
    assert n_bound > 20


# ------------------------------------------------------------------------------------- round 2: SURVEY 8(d) sample sizes
def _pool_map(fn, items):
    import multiprocessing as mp
    with mp.get_context("fork").Pool(min(len(items), max(1, (mp.cpu_count() or 2) - 1))) as pool:
        return pool.map(fn, items, chunksize=1)


def _newton(args):
    x0, x_ref, u_ref, kw = args
    x, u, K, s, h = O.newton_Algorithm(x0, x_ref, u_ref, **kw)
    return x, u, np.array(K), np.array(s), h


def test_c2_three_iterations_64_problems(fa_ref):
    """The oracle against the unmodified reference on 64 problems of the config-2 batch, 3 iterations, both step-size
    regimes (fixture: make_golden.py c2three)."""
    g = golden("newton_c2_three_iters")
    x_ref, u_ref, _ = fa_ref
    for tag, g0 in (("g01", 0.1), ("g1", 1.0)):
        res = _pool_map(_newton, [(g["x0"][i], x_ref, u_ref, dict(max_iters=3, tol=1e-4, gamma_0=g0)) for i in range(64)])
        for i, (x, u, K, s, h) in enumerate(res):
            assert h["n_try"] == list(g[tag + "_n_try"][i].astype(int)) and h["gamma"] == list(g[tag + "_gamma_acc"][i])
            assert rel_err(h["cost"], g[tag + "_cost"][i]) < 1e-12
            assert rel_err(x, g[tag + "_x"][i]) < 1e-10 and rel_err(u, g[tag + "_u"][i]) < 1e-10
            assert rel_err(s, g[tag + "_sigma"][i]) < 1e-10
            if i < 8:
                assert rel_err(K, g[tag + "_K8"][i]) < 1e-10


def test_c2_converged_rows(fa_ref):
    """Four of the sixteen config-2 problems the reference solved to convergence (make_golden.py c2conv): same
    iteration counts, cost histories, final trajectories."""
    g = golden("newton_c2_converged")
    x_ref, u_ref, _ = fa_ref
    pick = [0, 5, 10, 15]
    res = _pool_map(_newton, [(g["x0"][i], x_ref, u_ref, dict(max_iters=5000, tol=1e-4, gamma_0=0.1)) for i in pick])
    for i, (x, u, K, s, h) in zip(pick, res):
        n = int(g["iters"][i])
        assert h["iters"] == n and h["status"] == O.STATUS_CONVERGED
        assert rel_err(h["cost"], g["cost"][i, :n + 1]) < 1e-11
        assert rel_err(h["sigma_norm"], g["sigma_norm"][i, :n]) < 1e-9
        assert rel_err(x, g["x"][i]) < 1e-10 and rel_err(u, g["u"][i]) < 1e-10 and rel_err(s, g["sigma"][i]) < 1e-9


def _track(args):
    x_opt, u_opt, K, x0 = args
    with np.errstate(all="ignore"):
        return O.simulate_tracking(x_opt, u_opt, K, x0)


def test_c3_256_rollouts():
    g = golden("lqr_tracking_c3")
    d = golden("acrobot_optimal_trajectory")
    K = O.solve_LQR_tracking(d["x"], d["u"])
    res = _pool_map(_track, [(d["x"], d["u"], K, g["x0"][i]) for i in range(256)])
    X = np.array([r[0] for r in res])
    U = np.array([r[1] for r in res])
    assert rel_err(X[:, ::10], g["x_track_10"]) < 1e-10 and rel_err(U[:, ::10], g["u_track_10"]) < 1e-10
    assert rel_err(X.sum(axis=1), g["x_sum"]) < 1e-10


def _sweep(args):
    x0, k, x_ref, u_ref, steps = args
    if k == 0:
        u = np.zeros_like(u_ref)
        x = O.simulate_open_loop(x0, u)
    else:
        x, u, _, _, h = O.newton_Algorithm(x0, x_ref, u_ref, max_iters=k, tol=1e-4, gamma_0=0.1)
    Ad, Bd, q, r, QT2, qT = O.build_stage_lists(x, u, x_ref, u_ref)
    K, sig, dJ = O.calculate_K_and_sigma(Ad, Bd, q, r, 2.0 * O.Q_NEWTON, 2.0 * O.R_NEWTON, QT2, qT)
    return O.stepsize_sweep(x, u, K, sig, x_ref, u_ref, steps), dJ, x[-1]


def test_c5_base_iterates_from_newton_iterations(fa_ref):
    """Config 5 fixture (make_golden.py sweep256): the base iterates with k_p <= 12 Newton iterations behind them
    (the oracle needs 50 ms per iteration; the GPU test checks all 256)."""
    g = golden("sweep_c5")
    x_ref, u_ref, _ = fa_ref
    x0s = np.random.default_rng(1).uniform(-0.2, 0.2, (4096, 4))
    x0s[0] = 0.0
    pick = [i for i in range(256) if g["k"][i] <= 12][:48]
    res = _pool_map(_sweep, [(x0s[g["rows"][i]], int(g["k"][i]), x_ref, u_ref, g["steps"]) for i in pick])
    for i, (c, dJ, xT) in zip(pick, res):
        assert rel_err(c, g["costs"][i]) < 1e-11 and abs(dJ - g["delta_J"][i]) < 1e-10 * abs(dJ)
        assert rel_err(xT, g["x_T"][i]) < 1e-11


@pytest.mark.parametrize("dx", [0.05, 0.1, 0.15, 0.2])
def test_mpc_oracle_against_the_shipped_figures(dx):
    """The reference's MPC is CasADi + IPOPT (absent here); what it ships are figures.  The Riccati restatement
    (and, for dx = 0.05, the active-set solve of the box-constrained QP) reproduces every peak of their error curves
    (tests/mpc_figures.py: numbers read off figures/mpc/tracking_dx_*_err.png)."""
    import mpc_figures
    d = golden("acrobot_optimal_trajectory")
    tau = mpc_figures.FIGURES[dx][0]
    if tau is None:
        xr, ur = O.solve_mpc_tracking(d["x"][0] + dx, d["x"], d["u"], 501, T_pred=75)
    else:
        xr, ur, _ = O.solve_mpc_tracking_box(d["x"][0] + dx, d["x"], d["u"], 501, T_pred=75, tau_max=tau)
    mpc_figures.check(dx, d["t"], d["x"], d["u"], xr, ur)
