"""The drop-in modules used the way main.py uses the reference (task_1..task_4 recipes), on the GPU."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, golden, rel_err
from oracle import acro_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def mods():
    from gymnast_optimalcontrol_b200 import dynamics as dyn
    from gymnast_optimalcontrol_b200 import trajectory_generation as tg
    from gymnast_optimalcontrol_b200 import trajectory_tracking as tt
    return dyn, tg, tt


def test_dynamics_single_and_batched(mods):
    dyn, tg, tt = mods
    g = golden("dyn_kat")
    x, u = g["x"][0], g["u"][0]
    out = dyn.dynamics(x, u)
    assert isinstance(out, np.ndarray) and out.shape == (4,)
    np.testing.assert_allclose(out, [0.30842243714633577, -0.21036198377515106, 0.34430441728616834, -0.34211929152559867],
                               rtol=1e-12)
    np.testing.assert_allclose(dyn.continuous_dynamics(x, u), [0.5, -0.7, -8.097062412610704, 18.791857037341483], rtol=1e-12)
    A, B = dyn.Calculate_A_B_matrixes(x, u)
    assert A.shape == (4, 4) and B.shape == (4, 2)
    assert rel_err(A, g["A_c"][0]) < TOL and rel_err(B, g["B_c"][0]) < TOL
    assert rel_err(dyn.dynamics(g["x"], g["u"]), g["step"]) < TOL
    Ab, Bb = dyn.Calculate_A_B_matrixes(g["x"], g["u"])
    assert rel_err(Ab, g["A_c"]) < TOL and rel_err(Bb, g["B_c"]) < TOL
    xt = torch.from_numpy(g["x"]).cuda()
    ut = torch.from_numpy(g["u"]).cuda()
    o = dyn.dynamics(xt, ut)
    assert o.is_cuda and rel_err(o.cpu().numpy(), g["step"]) < TOL
    Ad, Bd = tg.discretize_linearization(A, B, dyn.dt)
    assert rel_err(Ad, np.eye(4) + dyn.dt * g["A_c"][0]) < 1e-15 and rel_err(Bd, dyn.dt * g["B_c"][0]) < 1e-15
    assert dyn.set_params(2).m1 == 2.0 and dyn.active_params().m1 == 1.0  # set_params does not rebind (dynamics.py:147)


def test_task2_recipe(mods):
    """main.py:55-71 with a short iteration cap; the pieces are exposed one by one like the reference's functions."""
    dyn, tg, tt = mods
    g = golden("newton_task2")
    b = golden("newton_task2_blocks")
    x_ref, u_ref, t_ref = tg.get_fully_actuated_ref(os.path.join(GOLDEN, "fully_actuated_trajectory.npz"))
    x0 = np.array([0, 0, 0, 0])
    x, u, K, sigma, hist = tg.newton_Algorithm(x0, x_ref, u_ref, max_iters=4, tol=1e-4, gamma_0=0.1, plot_armijo_iters=7,
                                               return_history=True, verbose=False)
    assert x.shape == (501, 4) and u.shape == (500, 2) and len(K) == 500 and K[0].shape == (2, 4) and sigma[0].shape == (2,)
    assert rel_err(hist["cost"], g["cost"][:5]) < TOL and rel_err(hist["sigma_norm"], g["sigma_norm"][:4]) < TOL
    assert rel_err(np.array(hist["x_trajs"]), g["x_trajs"][:5]) < TOL
    assert rel_err(np.array(hist["sigmas"]), g["sigmas"][:4]) < TOL
    # the stand-alone steps
    u0 = np.zeros_like(u_ref)
    xo = tg.simulate_open_loop(x0, u0)
    assert rel_err(xo, b["x_open"]) < TOL
    lam = tg.compute_costate_trajectory(xo, u0, x_ref, u_ref)
    assert len(lam) == 501 and rel_err(np.array(lam), b["lam"]) < TOL
    lists = tg.build_stage_lists(xo, u0, x_ref, u_ref, lam)
    assert rel_err(np.array(lists[0]), b["A_list"]) < TOL and rel_err(np.array(lists[1]), b["B_list"]) < TOL
    assert rel_err(np.array(lists[5]), b["q_list"]) < TOL and rel_err(np.array(lists[6]), b["r_list"]) < TOL
    assert rel_err(lists[7], b["Q_T_block"]) == 0 and rel_err(lists[8], b["q_T"]) < TOL
    K0, s0, dJ = tg.calculate_K_and_sigma(*lists)
    assert rel_err(np.array(K0), b["K0"]) < TOL and rel_err(np.array(s0), b["sigma0"]) < TOL
    assert abs(dJ - b["delta_J0"]) < TOL * abs(b["delta_J0"])
    xn, un = tg.forward_closed_loop_update(xo, u0, K0, s0, gamma=0.1)
    assert rel_err(xn, g["x_trajs"][1]) < TOL
    c = tg.total_cost(xn, un, x_ref, u_ref, tg.Q, tg.R, tg.Q_T)
    assert abs(c - g["cost"][1]) < TOL * g["cost"][1]
    l, gx, gu, hx, hu = tg.derivatives_Cost(xo[3], x_ref[3], u0[3], u_ref[3], tg.Q, tg.R)
    lo, gxo, guo, hxo, huo = O.derivatives_Cost(xo[3], x_ref[3], u0[3], u_ref[3], tg.Q, tg.R)
    assert abs(l - lo) < TOL * max(1, abs(lo)) and rel_err(gx, gxo) < TOL and rel_err(gu, guo) < TOL
    assert np.array_equal(hx, hxo) and np.array_equal(hu, huo)
    lT, gT, hT = tg.derivatives_Cost(xo[-1], x_ref[-1], np.zeros(2), np.zeros(2), Q=None, R=None, Q_T=tg.Q_T, terminal=True)
    lTo, gTo, hTo = O.derivatives_Cost(xo[-1], x_ref[-1], None, None, None, None, Q_T=tg.Q_T, terminal=True)
    assert abs(lT - lTo) < TOL * max(1, abs(lTo)) and rel_err(gT, gTo) < TOL and np.array_equal(hT, hTo)
    s = golden("sweep_iter0")
    steps, costs = tg.armijo_sweep(xo, u0, K0, s0, x_ref, u_ref)
    assert np.array_equal(steps, s["steps"]) and rel_err(costs, s["costs"]) < TOL


def test_task1_recipe_length_mismatch_message(mods, capsys):
    dyn, tg, tt = mods
    g = golden("newton_task1")
    x, u, K, s, h = tg.newton_Algorithm(g["x0"], g["x_ref"], g["u_ref"], max_iters=3, tol=1e-4, gamma_0=0.05)
    assert "u_ref has same length as x_ref (501)" in capsys.readouterr().out  # tg:302
    assert rel_err(h["cost"], g["cost"][:4]) < TOL and u.shape == (500, 2)


def test_batched_newton_through_the_dropin(mods, fa_ref):
    dyn, tg, tt = mods
    x_ref, u_ref, _ = fa_ref
    x0 = np.random.default_rng(1).uniform(-0.2, 0.2, (4096, 4))[:48]
    x, u, K, s, h = tg.newton_Algorithm(x0, x_ref, u_ref, max_iters=6, tol=1e-4, gamma_0=0.1, verbose=False)
    g = golden("newton_c2_rows")
    assert x.shape == (48, 501, 4) and K.shape == (48, 500, 2, 4) and h["cost"].shape == (48, 7)
    assert rel_err(x[1:4], g["x"]) < TOL and rel_err(K[1:4], g["K"]) < TOL and rel_err(h["cost"][1:4], g["cost"]) < TOL
    # per-problem weights through keyword arguments
    Qs = np.repeat(tg.Q[None], 48, 0)
    Qs[5] = np.diag([50.0, 60.0, 0.1, 0.1])
    x2, u2, K2, s2, h2 = tg.newton_Algorithm(x0, x_ref, u_ref, max_iters=3, tol=1e-4, gamma_0=0.1, verbose=False, Q=Qs)
    xo, uo, Ko, so, ho = O.newton_Algorithm(x0[5], x_ref, u_ref, max_iters=3, tol=1e-4, gamma_0=0.1, Q=Qs[5])
    assert rel_err(x2[5], xo) < TOL and rel_err(K2[5], Ko) < TOL
    xo, uo, Ko, so, ho = O.newton_Algorithm(x0[6], x_ref, u_ref, max_iters=3, tol=1e-4, gamma_0=0.1)
    assert rel_err(x2[6], xo) < TOL


def test_task3_recipe(mods):
    """main.py:99-119: LQR tracking with +0.2 / +0.3 on all states."""
    dyn, tg, tt = mods
    g = golden("lqr_tracking")
    d = golden("acrobot_optimal_trajectory")
    x_ref, u_ref, t_ref = d["x"], d["u"], d["t"]
    for i, dist in enumerate((0.2, 0.3)):
        xt, ut = tt.LQR_tracking(x_ref, u_ref, t_ref, x0_perturbed=x_ref[0].copy() + dist)
        assert xt.shape == (501, 4) and ut.shape == (500, 2)
        assert rel_err(xt, g["x_track"][i]) < TOL and rel_err(ut, g["u_track"][i]) < TOL
    K = tt.solve_LQR_tracking(x_ref, u_ref)
    assert len(K) == 500 and rel_err(np.array(K), g["K_reg"]) < TOL
    xt, ut = tt.simulate_tracking(x_ref, u_ref, K, g["x0"][:20])
    assert rel_err(xt, g["x_track"][:20]) < TOL
    xt0, _ = tt.LQR_tracking(x_ref, u_ref, t_ref)
    assert np.max(np.abs(xt0 - x_ref)) < 1e-9


def test_task4_recipe(mods):
    """main.py:122-145: MPC tracking with +0.1; compute_P_inf and solver_mpc stand-alone."""
    dyn, tg, tt = mods
    d = golden("acrobot_optimal_trajectory")
    g = golden("p_inf")
    x_ref, u_ref, t_ref = d["x"], d["u"], d["t"]
    P = tt.compute_P_inf(g["A_f"], g["B_f"], g["Q"], g["R"])
    assert rel_err(P, g["P_inf"]) < TOL
    x0 = x_ref[0].copy() + 0.1
    xr, ur = tt.solve_mpc_tracking(x0, x_ref, u_ref, len(t_ref))
    xo, uo = O.solve_mpc_tracking(x0, x_ref, u_ref, 501)
    assert xr.shape == (501, 4) and ur.shape == (500, 2)
    assert rel_err(xr, xo) < TOL and rel_err(ur, uo) < TOL
    assert abs(abs(ur[0, 1] - u_ref[0, 1]) - 0.80644713) < 1e-6  # figures/mpc/tracking_dx_0.1_err.png (~0.81)
    Ad, Bd = O.linearize_discrete(x_ref[:-1], u_ref)
    U0, X, U = tt.solver_mpc(x0 - x_ref[0], list(Ad[:75]), list(Bd[:75]), tt.Q_mpc, tt.R_mpc, P, 75, u_ref[:75])
    U0o, Xo, Uo = O.solver_mpc_kkt(x0 - x_ref[0], list(Ad[:74]), list(Bd[:74]), O.Q_MPC, O.R_MPC, g["P_inf"], 75)
    assert X.shape == (75, 4) and U.shape == (75, 2)
    assert rel_err(U0, U0o) < 1e-7 and rel_err(X, Xo) < 1e-7 and rel_err(U, Uo) < 1e-7
    assert rel_err(ur[0] - u_ref[0], U0) < TOL


def test_per_problem_parameters_through_the_dropins(mods):
    """Keyword-only extension params_b (B, 11) or a dict of per-problem arrays."""
    dyn, tg, tt = mods
    from gymnast_optimalcontrol_b200.batched import PARAM_SETS, PHYS_FIELDS
    rng = np.random.default_rng(12)
    n = 9
    x, u = rng.uniform(-1, 1, (n, 4)), rng.uniform(-2, 2, (n, 2))
    sets = [dict(PARAM_SETS[1 + b % 3]) for b in range(n)]
    rows = np.array([[s[f] for f in PHYS_FIELDS] for s in sets])
    step = dyn.dynamics(x, u, params_b=rows)
    A, B = dyn.Calculate_A_B_matrixes(x, u, params_b=rows)
    for b in range(n):
        m = O.Model(sets[b])
        assert rel_err(step[b], O.dynamics(x[b], u[b], m)) < TOL
        Ac, Bc = O.Calculate_A_B_matrixes(x[b], u[b], m)
        assert rel_err(A[b], Ac) < TOL and rel_err(B[b], Bc) < TOL
    # dict form: only the masses vary
    m2 = rng.uniform(0.8, 1.2, n)
    step = dyn.dynamics(x, u, params_b={"m2": m2})
    Ut = rng.uniform(-1, 1, (n, 20, 2))
    Xo = tg.simulate_open_loop(x, Ut, params_b={"m2": m2})
    for b in range(n):
        m = O.Model(dict(PARAM_SETS[1], m2=m2[b]))
        assert rel_err(step[b], O.dynamics(x[b], u[b], m)) < TOL
        assert rel_err(Xo[b], O.simulate_open_loop(x[b], Ut[b], m)) < TOL
    with pytest.raises(ValueError):
        dyn.dynamics(x, u, params_b=rows[:3])


# ------------------------------------------------------------------------------------- round 2: ownership, pipelining, shapes
def test_results_of_consecutive_calls_do_not_alias(mods):
    """Pinned-host torch inputs get pinned-host results: every call returns a FRESH buffer (the staging
    buffer used to be cached per shape, so x1 = dynamics(xp, u1); x2 = dynamics(xp, u2) made x1 IS x2)."""
    dyn, tg, tt = mods
    rng = np.random.default_rng(5)
    xp = torch.from_numpy(rng.uniform(-1, 1, (64, 4))).pin_memory()
    u1 = torch.from_numpy(rng.uniform(-2, 2, (64, 2))).pin_memory()
    u2 = torch.from_numpy(rng.uniform(-2, 2, (64, 2))).pin_memory()
    x1 = dyn.dynamics(xp, u1)
    keep = x1.clone()
    x2 = dyn.dynamics(xp, u2)
    assert x1.is_pinned() and x2.is_pinned() and x1.data_ptr() != x2.data_ptr()
    assert torch.equal(x1, keep) and not torch.equal(x1, x2)
    for b in (0, 63):
        assert rel_err(x1[b].numpy(), O.dynamics(xp[b].numpy(), u1[b].numpy())) < TOL
        assert rel_err(x2[b].numpy(), O.dynamics(xp[b].numpy(), u2[b].numpy())) < TOL


def test_pipelined_newton_equals_blocking_newton(mods, fa_ref):
    """newton_Algorithm(block=False): three solves in flight over two solver-state slots; every result equals the
    blocking call bit for bit, results own their buffers, the history of a short solve is not polluted by the longer
    solve that used the slot before it."""
    dyn, tg, tt = mods
    x_ref, u_ref, _ = fa_ref
    rng = np.random.default_rng(9)
    x0s = [torch.from_numpy(rng.uniform(-0.2, 0.2, (96, 4))).pin_memory() for _ in range(3)]
    xr, ur = torch.from_numpy(x_ref).pin_memory(), torch.from_numpy(u_ref).pin_memory()
    kws = [dict(max_iters=5, tol=1e-4, gamma_0=1.0), dict(max_iters=5, tol=1e-4, gamma_0=0.1), dict(max_iters=5, tol=50.0, gamma_0=0.5)]
    blocking = [tg.newton_Algorithm(x0, xr, ur, verbose=False, **kw) for x0, kw in zip(x0s, kws)]
    pend = [tg.newton_Algorithm(x0, xr, ur, verbose=False, block=False, **kw) for x0, kw in zip(x0s, kws)]
    res = [p.result() for p in pend]
    for a, b in zip(blocking, res):
        for i in range(4):
            assert torch.equal(a[i], b[i]) and a[i].data_ptr() != b[i].data_ptr() and b[i].is_pinned()
        for k in ("cost", "sigma_norm", "n_try", "gamma", "iters", "status"):
            assert np.array_equal(a[4][k], b[4][k], equal_nan=True), k
    # the third solve stops early (tol = 50): rows of the history beyond each problem's last iteration are NaN, not
    # left-overs of the first solve that ran in the same slot
    h = res[2][4]
    assert h["iters"].min() < 5
    for b in range(96):
        assert np.isnan(h["cost"][b, h["iters"][b] + 1:]).all()
    # return_gains=False: K and sigma stay on the device
    x, u, K, s, hh = tg.newton_Algorithm(x0s[1], xr, ur, verbose=False, return_gains=False, **kws[1])
    from gymnast_optimalcontrol_b200.batched import Traj
    assert isinstance(K, Traj) and isinstance(s, Traj) and torch.equal(x, blocking[1][0])
    assert torch.equal(K.batch_major().reshape(96, 500, 2, 4).cpu(), blocking[1][2])


def test_history_lists_follow_the_reference(mods, fa_ref):
    """history['x_trajs'] / ['sigmas'] / ['cost'] / ['sigma_norm'] lengths as the reference builds them (tg:322-327,
    341-342, 387-388): a line-search failure appends sigma but no iterate and no cost; progress line every 10th
    iteration (tg:391-392)."""
    dyn, tg, tt = mods
    x_ref, u_ref, _ = fa_ref
    x, u, K, s, h = tg.newton_Algorithm(np.zeros(4), x_ref, u_ref, max_iters=3, tol=1e-4, gamma_0=0.1, return_history=True, verbose=False)
    assert len(h["x_trajs"]) == 4 and len(h["sigmas"]) == 3 and len(h["cost"]) == 4 and len(h["sigma_norm"]) == 3
    g = golden("newton_task2")
    assert rel_err(np.array(h["x_trajs"]), g["x_trajs"][:4]) < TOL and rel_err(np.array(h["sigmas"]), g["sigmas"][:3]) < TOL
    # c = 1e6 can never be satisfied: the first line search fails (tg:367-369)
    x, u, K, s, h = tg.newton_Algorithm(np.zeros(4), x_ref, u_ref, max_iters=3, tol=1e-4, gamma_0=0.1, c=1e6, return_history=True, verbose=False)
    assert h["status"] == 3 and h["iters"] == 1
    assert len(h["x_trajs"]) == 1 and len(h["sigmas"]) == 1 and len(h["cost"]) == 1 and len(h["sigma_norm"]) == 1
    # non-finite initial cost: same lengths (the count comes from iters / status, not from isfinite)
    x, u, K, s, h = tg.newton_Algorithm(np.array([np.nan, 0, 0, 0]), x_ref, u_ref, max_iters=3, tol=1e-4, gamma_0=0.1, verbose=False)
    assert h["status"] == 3 and len(h["cost"]) == 1 and np.isnan(h["cost"][0]) and len(h["sigma_norm"]) == 1


def test_progress_line(mods, fa_ref, capsys):
    dyn, tg, tt = mods
    x_ref, u_ref, _ = fa_ref
    tg.newton_Algorithm(np.zeros(4), x_ref, u_ref, max_iters=12, tol=1e-4, gamma_0=0.1)
    out = capsys.readouterr().out
    g = golden("newton_task2")
    assert "Iter 0: Cost=%.2f, diff_cost=%.2e, " % (g["cost"][1], g["cost"][0] - g["cost"][1]) in out
    assert "Iter 10: Cost=%.2f" % g["cost"][11] in out


def test_mpc_outputs_have_the_shapes_of_the_references(mods):
    """solve_mpc_tracking returns arrays shaped like x_ref / u_ref, zero beyond T (tt:27-28)."""
    dyn, tg, tt = mods
    d = golden("acrobot_optimal_trajectory")
    x0 = d["x"][0] + 0.1
    xr, ur = tt.solve_mpc_tracking(x0, d["x"], d["u"], 101)
    assert xr.shape == (501, 4) and ur.shape == (500, 2)
    assert np.all(xr[101:] == 0.0) and np.all(ur[100:] == 0.0)
    xf, uf = tt.solve_mpc_tracking(x0, d["x"], d["u"], 501)
    assert rel_err(xr[:101], xf[:101]) < TOL and rel_err(ur[:100], uf[:100]) < TOL


def test_batched_reference_builders(mods):
    """task_1's recipe (main.py:33-41) for a batch of target torques: equilibria on the device, piecewise references,
    then the solver on per-problem references."""
    dyn, tg, tt = mods
    g = golden("newton_task1")
    u2 = np.array([[0.5, 0.5], [0.4, 0.4], [0.3, 0.6]])
    x_e1, u_e1 = tg.compute_equilibrium(np.zeros((3, 2)), (0.1, -0.1))
    x_e2, u_e2 = tg.compute_equilibrium(u2, (0.35, -0.35))
    assert x_e1.shape == (3, 4) and np.abs(x_e2[0] - g["x_e2"]).max() < 1e-9 and np.abs(x_e1[0] - g["x_e1"]).max() < 1e-9
    xs, us = tg.compute_equilibrium(u2[0], (0.35, -0.35))  # the single-problem call of the reference
    assert np.abs(xs - g["x_e2"]).max() < 1e-10
    t_ref, x_ref, u_ref = tg.define_reference_piecewise(10.0, x_e1, x_e2, u_e1, u_e2)
    assert x_ref.shape == (3, 501, 4) and u_ref.shape == (3, 501, 2)
    assert rel_err(x_ref[0], g["x_ref"]) < 1e-9 and rel_err(u_ref[0], g["u_ref"]) < 1e-12
    x, u, K, s, h = tg.newton_Algorithm(x_e1, x_ref, u_ref, max_iters=3, tol=1e-4, gamma_0=0.05, verbose=False)
    assert rel_err(h["cost"][0], g["cost"][:4]) < 1e-8
    with pytest.raises(RuntimeError, match="Root finder failed"):
        tg.compute_equilibrium(np.array([[0.0, 50.0]]), (0.1, 0.1))


@pytest.mark.parametrize("tau_max", [None, 18.0])
def test_pipelined_mpc_tracking_equals_blocking(mods, tau_max):
    """solve_mpc_tracking(block=False): uploads on the upload stream, copies back on the copy stream, several calls in
    flight - the same numbers as the blocking calls, for shared and per-problem references, host and device inputs,
    with and without the input box, and with T < N (zero padding)."""
    dyn, tg, tt = mods
    d = golden("acrobot_optimal_trajectory")
    rng = np.random.default_rng(17)
    Bn = 96
    calls = []
    for i in range(4):
        x0 = d["x"][0] + rng.uniform(-0.1, 0.1, (Bn, 4))
        if i % 2:
            xr = np.repeat(d["x"][None], Bn, 0) + rng.uniform(-1e-3, 1e-3, (Bn, 501, 4))
            ur = np.repeat(d["u"][None], Bn, 0)
        else:
            xr, ur = d["x"], d["u"]
        calls.append((x0, xr, ur, 61 if i == 3 else 81))
    want = [tt.solve_mpc_tracking(x0, xr, ur, T, tau_max=tau_max) for x0, xr, ur, T in calls]
    # host (pinned) inputs: everything in flight before the first result is read
    pin = [tuple(torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in c[:3]) + (c[3],) for c in calls]
    pend = [tt.solve_mpc_tracking(x0, xr, ur, T, tau_max=tau_max, block=False) for x0, xr, ur, T in pin]
    for p, (wx, wu) in zip(pend, want):
        gx, gu = p.result()
        assert isinstance(gx, torch.Tensor) and gx.is_pinned() and tuple(gx.shape) == wx.shape
        assert np.array_equal(gx.numpy(), wx) and np.array_equal(gu.numpy(), wu)
        assert p.ready()
    # device inputs stay on the caller's stream; results come back as device tensors
    dev = [tuple(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in c[:3]) + (c[3],) for c in calls]
    pend = [tt.solve_mpc_tracking(x0, xr, ur, T, tau_max=tau_max, block=False) for x0, xr, ur, T in dev]
    for p, (wx, wu) in zip(pend, want):
        gx, gu = p.result()
        assert gx.is_cuda and np.array_equal(gx.cpu().numpy(), wx) and np.array_equal(gu.cpu().numpy(), wu)
    # NumPy in, NumPy out
    gx, gu = tt.solve_mpc_tracking(*calls[1][:3], calls[1][3], tau_max=tau_max, block=False).result()
    assert isinstance(gx, np.ndarray) and np.array_equal(gx, want[1][0]) and np.array_equal(gu, want[1][1])


def test_pipelined_lqr_tracking_equals_blocking(mods):
    dyn, tg, tt = mods
    d = golden("acrobot_optimal_trajectory")
    rng = np.random.default_rng(18)
    x0s = [d["x"][0] + rng.uniform(-0.1, 0.1, (200, 4)) for _ in range(3)]
    want = [tt.LQR_tracking(d["x"], d["u"], d["t"], x0_perturbed=x0) for x0 in x0s]
    pend = [tt.LQR_tracking(d["x"], d["u"], d["t"], x0_perturbed=torch.from_numpy(x0).pin_memory(), block=False) for x0 in x0s]
    for p, (wx, wu) in zip(pend, want):
        gx, gu = p.result()
        assert np.array_equal(gx.numpy(), wx) and np.array_equal(gu.numpy(), wu)


def test_terminal_weight_is_cached_per_weights(mods):
    """P_inf is computed once per (parameters, Q, R) and the cache distinguishes weights."""
    dyn, tg, tt = mods
    d = golden("acrobot_optimal_trajectory")
    x0 = d["x"][0] + 0.05
    a = tt.solve_mpc_tracking(x0, d["x"], d["u"], 41, return_info=True)
    b = tt.solve_mpc_tracking(x0, d["x"], d["u"], 41, Q=np.diag([10.0, 10.0, 1.0, 1.0]), return_info=True)
    c = tt.solve_mpc_tracking(x0, d["x"], d["u"], 41, return_info=True)
    assert np.array_equal(a[2]["P_inf"], c[2]["P_inf"]) and np.array_equal(a[0], c[0])
    assert not np.allclose(a[2]["P_inf"], b[2]["P_inf"])
    assert rel_err(b[2]["P_inf"], tt.compute_P_inf(*[np.asarray(m) for m in _lin_f(dyn)], np.diag([10.0, 10.0, 1.0, 1.0]), tt.R_mpc)) < 1e-9


def _lin_f(dyn):
    A, B = dyn.Calculate_A_B_matrixes(np.array([np.pi, 0, 0, 0]), np.array([0.0, 0.0]))
    return np.eye(4) + 2e-2 * A, 2e-2 * B


def test_constrained_mpc_against_the_shipped_figure(mods):
    """The only artefact the reference ships for its constrained MPC is figures/mpc/tracking_constrained.png
    (`test_constraints = True`, tt:87-91, 112-114; IPOPT is not available here, so this is a FIGURE-LEVEL pin, numbers
    read off the PNG by eye): with |u| <= 18 the elbow torque sits flat on -18 from t ~ 3.45 s to ~ 3.75 s while the
    reference dips to -23.9, peaks at ~ 13.7 at t ~ 2.82 s (reference 12.5), comes back only to ~ -9 at t ~ 3.88 s
    (reference -6.5); theta2 overshoots to ~ 0.58 at t ~ 3.4 s (reference 0.32) and undershoots to ~ -3.27 at ~ 4.22 s
    (reference -3.20); the upright is reached.  The curves start at ~ 0.05 rad: x0 = x_ref[0] + 0.05 (main.py:127)."""
    dyn, tg, tt = mods
    d = golden("acrobot_optimal_trajectory")
    t = d["t"]
    xr, ur = tt.solve_mpc_tracking(d["x"][0] + 0.05, d["x"], d["u"], 501, tau_max=18.0)
    u2 = ur[:, 1]
    assert np.abs(ur).max() <= 18.0 * (1 + 1e-12) and np.abs(ur[:, 0]).max() == 0.0
    sat = np.where(u2 <= -18.0 * (1 - 1e-12))[0]
    assert len(sat) == sat.max() - sat.min() + 1                      # one flat segment
    assert abs(t[sat.min()] - 3.45) <= 0.05 and abs(t[sat.max()] - 3.75) <= 0.05
    assert d["u"][sat, 1].min() < -23.8                                 # where the reference asks for -23.9
    i = int(np.argmax(u2))
    assert abs(t[i] - 2.82) <= 0.04 and abs(u2[i] - 13.7) <= 0.6 and abs(d["u"][:, 1].max() - 12.5) < 0.05
    w = np.where((t[:-1] > 3.7) & (t[:-1] < 4.0))[0]
    j = w[np.argmax(u2[w])]
    assert abs(t[j] - 3.88) <= 0.04 and abs(u2[j] + 9.0) <= 0.5
    w2 = np.where((t > 3.2) & (t < 3.6))[0]
    k = w2[np.argmax(xr[w2, 1])]
    assert abs(t[k] - 3.40) <= 0.04 and abs(xr[k, 1] - 0.58) <= 0.05
    m = int(np.argmin(xr[:, 1]))
    assert abs(t[m] - 4.22) <= 0.06 and abs(xr[m, 1] + 3.27) <= 0.06
    assert np.abs(xr[-1] - np.array([np.pi, 0, 0, 0])).max() < 1e-2


@pytest.mark.parametrize("dx", [0.05, 0.1, 0.15, 0.2])
def test_mpc_dropin_against_the_shipped_figures(mods, dx):
    """main.py's task_4 on the drop-in for the disturbances the reference ships figures for: every peak of the error
    curves of figures/mpc/tracking_dx_*_err.png (tests/mpc_figures.py); dx = 0.05 is the run with the input box."""
    import mpc_figures
    dyn, tg, tt = mods
    d = golden("acrobot_optimal_trajectory")
    xr, ur = tt.solve_mpc_tracking(d["x"][0] + dx, d["x"], d["u"], len(d["t"]), tau_max=mpc_figures.FIGURES[dx][0])
    mpc_figures.check(dx, d["t"], d["x"], d["u"], xr, ur)


def test_mpc_dropin_with_per_problem_parameters(mods):
    """solve_mpc_tracking(params_b=...): shared reference, every problem its own plant; blocking == non-blocking."""
    from gymnast_optimalcontrol_b200.batched import PARAM_SETS, PHYS_FIELDS
    dyn, tg, tt = mods
    d = golden("acrobot_optimal_trajectory")
    n = 20
    rng = np.random.default_rng(5)
    rows = np.array([[PARAM_SETS[1][f] * (1.0 if (f == "g" or b == 0) else rng.uniform(0.98, 1.02)) for f in PHYS_FIELDS]
                     for b in range(n)])
    x0 = d["x"][0] + rng.uniform(-0.05, 0.05, (n, 4))
    xr, ur = tt.solve_mpc_tracking(x0, d["x"], d["u"], 101, params_b=rows)
    assert xr.shape == (n, 501, 4) and ur.shape == (n, 500, 2) and np.all(xr[:, 101:] == 0.0)
    xn, un = tt.solve_mpc_tracking(x0[0], d["x"], d["u"], 101)
    assert rel_err(xr[0, :101], xn[:101]) < TOL and rel_err(ur[0, :100], un[:100]) < TOL
    xo, uo = O.solve_mpc_tracking(x0[3], d["x"], d["u"], 101, m=O.Model(dict(zip(PHYS_FIELDS, rows[3]))))
    assert rel_err(xr[3, :101], xo[:101]) < TOL and rel_err(ur[3, :100], uo[:100]) < TOL
    gx, gu = tt.solve_mpc_tracking(x0, d["x"], d["u"], 101, params_b=rows, block=False).result()
    assert np.array_equal(gx, xr) and np.array_equal(gu, ur)
    # the same with the input box (a stretch where it binds), against the oracle with that problem's model
    t0 = 150
    xs, us = d["x"][t0:t0 + 60], d["u"][t0:t0 + 59]
    x0b = xs[0] + rng.uniform(-0.02, 0.02, (n, 4))
    xb, ub, info = tt.solve_mpc_tracking(x0b, xs, us, 46, T_pred=20, params_b=rows, tau_max=12.0, return_info=True)
    assert info["n_active"].max() > 0 and np.abs(ub).max() <= 12.0 * (1 + 1e-12)
    for b in (0, 7):
        xo, uo, nao = O.solve_mpc_tracking_box(x0b[b], xs, us, 46, T_pred=20, tau_max=12.0, m=O.Model(dict(zip(PHYS_FIELDS, rows[b]))))
        assert rel_err(xb[b, :46], xo) < TOL and rel_err(ub[b, :45], uo) < TOL
