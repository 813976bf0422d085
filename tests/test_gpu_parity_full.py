"""GPU parity at the sample sizes of SURVEY 8(d) and on every kernel the dispatch selects at its own batch size.

* every Newton kernel variant at the batch size where `acro_newton_solve` picks it (B = 4 800 ... 32 768), >= 16
  sampled problems each (first / last lane of first / last tile included) against the oracle, gamma_0 in {0.1, 1};
* config 2 (B = 4096): 16 problems to convergence and 64 problems x 3 iterations against fixtures produced by the
  UNMODIFIED reference (tests/golden/make_golden.py: c2conv, c2three);
* config 3 (B = 65 536): 256 tracked rollouts against the reference (lqr256);
* config 4 (B = 16 384): 16 problems x 500 receding-horizon steps at H = 50, 100, 200, shared and per-problem
  references, against the oracle (Riccati restatement; unpinned against IPOPT);
* config 5: 256 base iterates taken from Newton iterations 0..49 of config-2 problems against the reference (sweep256).

Tolerance 1e-9 relative per array; Armijo selections, iteration counts and status flags identical.
"""
import multiprocessing as mp

import numpy as np
import pytest
import torch

from conftest import golden, rel_err
from oracle import acro_oracle as O
from test_gpu_parity import TOL, aos, assert_gain_parity, dev, kmat, soa

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bt():
    from gymnast_optimalcontrol_b200 import batched
    assert torch.cuda.is_available()
    return batched


# ------------------------------------------------------------------------------------- oracle in a process pool
def _oracle_newton(args):
    x0, x_ref, u_ref, kw = args
    x, u, K, s, h = O.newton_Algorithm(x0, x_ref, u_ref, **kw)
    return x, u, np.array(K), np.array(s), h


def _oracle_mpc(args):
    x0, x_ref, u_ref, H = args
    return O.solve_mpc_tracking(x0, x_ref, u_ref, x_ref.shape[0], T_pred=H)


def pool_map(fn, items):
    """The oracle is a per-problem Python loop: spread the sampled problems over the host cores (fork: the children
    only run NumPy, they never touch CUDA)."""
    n = min(len(items), max(1, (mp.cpu_count() or 2) - 1), 32)
    with mp.get_context("fork").Pool(n) as pool:
        return pool.map(fn, items, chunksize=1)


def sample_rows(Bn, n, seed):
    """first / last lane of the first / last tile + a random spread"""
    edge = [0, 31, Bn - 32, Bn - 1] if Bn % 32 == 0 else [0, 31, 32 * (Bn // 32), Bn - 1]
    rest = np.random.default_rng(seed).choice(np.arange(32, Bn - 32), n - len(edge), replace=False)
    return sorted(set(edge) | set(int(r) for r in rest))


# ------------------------------------------------------------------------------------- dispatch variants
@pytest.mark.parametrize("Bn,variant,expect", [
    (4800, "auto", "acro::k_newton_duo<false,false,4>"),
    (8192, "auto", "acro::k_newton_duo<false,false,4>"),
    (9600, "auto", "acro::k_newton_ring<false,false,4,false>"),
    (16384, "auto", "acro::k_newton_ring<false,false,4,false>"),
    (32768, "auto", "acro::k_newton_ring<false,false,2,true>"),
])
@pytest.mark.parametrize("gamma_0,iters", [(0.1, 4), (1.0, 3)])
def test_dispatch_variants_at_their_batch_sizes(bt, fa_ref, Bn, variant, expect, gamma_0, iters):
    """The kernels the automatic dispatch selects beyond one tile per SM, at batch sizes where it selects them:
    16 sampled problems against the oracle, same tolerances as test_newton_vs_oracle_random_batch, and a second
    launch that must give the same bits."""
    x_ref, u_ref, _ = fa_ref
    name = bt.newton_kernel_name(Bn, kernel=variant)
    print("B = %d -> %s" % (Bn, name))
    assert name == expect
    x0s = np.random.default_rng(100 + Bn % 97).uniform(-0.2, 0.2, (Bn, 4))
    ref = bt.make_ref(x_ref, u_ref)
    kw = dict(max_iters=iters, tol=1e-4, gamma_0=gamma_0)
    st = bt.newton_solve(soa(x0s), ref, kernel=variant, **kw)
    torch.cuda.synchronize()
    rows = sample_rows(Bn, 16, Bn)
    X, U, K, S = (t.batch_major()[rows].cpu().numpy() for t in (st.X, st.U, st.K, st.S))
    hc, hn, hg = (h[:, rows].cpu().numpy() for h in (st.hist_cost, st.hist_ntry, st.hist_gamma))
    its, status = st.iters[rows].cpu().numpy(), st.status[rows].cpu().numpy()
    res = pool_map(_oracle_newton, [(x0s[b], x_ref, u_ref, kw) for b in rows])
    for i, (x, u, Ko, so, h) in enumerate(res):
        assert status[i] == h["status"] and its[i] == h["iters"], (rows[i], status[i], its[i], h["status"], h["iters"])
        n = len(h["n_try"])
        assert list(hn[:n, i]) == h["n_try"], (rows[i], list(hn[:n, i]), h["n_try"])
        assert list(hg[:n, i]) == h["gamma"]
        assert rel_err(hc[:n + 1, i], h["cost"]) < TOL
        assert rel_err(X[i], x) < TOL and rel_err(U[i], u) < TOL, rows[i]
        assert rel_err(S[i], so) < TOL and rel_err(K[i].reshape(-1, 2, 4), Ko) < TOL, rows[i]
    # every problem made progress, the line search did its job
    hcf = st.hist_cost[:iters + 1]
    assert bool(torch.isfinite(hcf).all()) and bool((hcf[1:] < hcf[:-1]).all())
    if gamma_0 == 1.0:
        assert int(st.hist_ntry[:iters].max()) >= 2  # the back-tracking path ran
    st2 = bt.newton_solve(soa(x0s), ref, kernel=variant, **kw)
    torch.cuda.synchronize()
    for a, b in ((st.X, st2.X), (st.U, st2.U), (st.K, st2.K), (st.S, st2.S)):
        assert torch.equal(a.data, b.data)
    assert torch.equal(st.hist_cost, st2.hist_cost) or torch.equal(torch.nan_to_num(st.hist_cost), torch.nan_to_num(st2.hist_cost))


# ------------------------------------------------------------------------------------- config 2
def c2_x0():
    x0 = np.random.default_rng(1).uniform(-0.2, 0.2, (4096, 4))
    x0[0] = 0.0
    return x0


@pytest.mark.parametrize("kernel", ["auto", "spec"])
def test_c2_sixteen_problems_to_convergence_vs_reference(bt, fa_ref, kernel):
    """Config 2 (B = 4096, gamma_0 = 0.1, tol = 1e-4) to convergence: 16 problems spread over the batch against runs
    of the unmodified reference (387-399 iterations each): iteration counts, every cost and max|sigma| of the
    history, Armijo tries and accepted steps, final x, u, sigma at 1e-9; gains by assert_gain_parity."""
    g = golden("newton_c2_converged")
    x_ref, u_ref, _ = fa_ref
    x0s = c2_x0()
    rows = g["rows"]
    assert len(rows) >= 16 and np.array_equal(x0s[rows], g["x0"])
    st = bt.newton_solve(soa(x0s), bt.make_ref(x_ref, u_ref), max_iters=5000, tol=1e-4, gamma_0=0.1, kernel=kernel)
    torch.cuda.synchronize()
    assert bool((st.status == 1).all())
    its = st.iters.cpu().numpy()
    assert np.array_equal(its[rows], g["iters"])
    X, U, K, S = (t.batch_major()[rows].cpu().numpy() for t in (st.X, st.U, st.K, st.S))
    for i, b in enumerate(rows):
        n = int(g["iters"][i])
        assert rel_err(st.hist_cost[:n + 1, b].cpu().numpy(), g["cost"][i, :n + 1]) < TOL
        assert rel_err(st.hist_sigma_norm[:n, b].cpu().numpy(), g["sigma_norm"][i, :n]) < TOL
        assert np.array_equal(st.hist_ntry[:n, b].cpu().numpy(), g["n_try"][i, :n].astype(np.int32))
        assert np.array_equal(st.hist_gamma[:n, b].cpu().numpy(), g["gamma_acc"][i, :n])
        assert rel_err(X[i], g["x"][i]) < TOL and rel_err(U[i], g["u"][i]) < TOL
        assert rel_err(S[i], g["sigma"][i]) < TOL
        assert_gain_parity(K[i].reshape(-1, 2, 4), g["K"][i], g["x_prev"][i], g["u_prev"][i], x_ref, u_ref)


@pytest.mark.parametrize("kernel", ["auto", "spec", "spec8"])
@pytest.mark.parametrize("tag,gamma_0", [("g01", 0.1), ("g1", 1.0)])
def test_c2_sixty_four_problems_three_iterations_vs_reference(bt, fa_ref, tag, gamma_0, kernel):
    """Config 2, 64 problems x 3 iterations (8 of them the corner lanes of the first and last tile) against the
    unmodified reference, at the shipped step size and in the back-tracking regime."""
    g = golden("newton_c2_three_iters")
    x_ref, u_ref, _ = fa_ref
    x0s = c2_x0()
    rows = g["rows"]
    assert len(rows) == 64 and np.array_equal(x0s[rows], g["x0"])
    st = bt.newton_solve(soa(x0s), bt.make_ref(x_ref, u_ref), max_iters=3, tol=1e-4, gamma_0=gamma_0, kernel=kernel)
    torch.cuda.synchronize()
    X, U, K, S = (t.batch_major()[rows].cpu().numpy() for t in (st.X, st.U, st.K, st.S))
    assert np.array_equal(st.hist_ntry[:3, rows].cpu().numpy().T, g[tag + "_n_try"].astype(np.int32))
    assert np.array_equal(st.hist_gamma[:3, rows].cpu().numpy().T, g[tag + "_gamma_acc"])
    assert rel_err(st.hist_cost[:4, rows].cpu().numpy().T, g[tag + "_cost"]) < TOL
    assert rel_err(st.hist_sigma_norm[:3, rows].cpu().numpy().T, g[tag + "_sigma_norm"]) < TOL
    for i in range(64):
        assert rel_err(X[i], g[tag + "_x"][i]) < TOL and rel_err(U[i], g[tag + "_u"][i]) < TOL, rows[i]
        assert rel_err(S[i], g[tag + "_sigma"][i]) < TOL, rows[i]
    assert rel_err(K[:8].reshape(8, -1, 2, 4), g[tag + "_K8"]) < TOL
    if gamma_0 == 1.0:
        assert g[tag + "_n_try"].max() >= 2


# ------------------------------------------------------------------------------------- config 3
def test_c3_256_rollouts_vs_reference(bt):
    """Config 3 (B = 65 536 LQR-tracked rollouts): 256 sampled problems against simulate_tracking of the unmodified
    reference (every 10th step of x and u, and the column sums over all 501 / 500 steps)."""
    g = golden("lqr_tracking_c3")
    d = golden("acrobot_optimal_trajectory")
    Bn = 65536
    x0 = d["x"][0] + np.random.default_rng(2).uniform(-0.3, 0.3, (Bn, 4))
    x0[0] = d["x"][0] + 0.2
    x0[1] = d["x"][0] + 0.3
    rows = g["rows"]
    assert len(rows) == 256 and np.array_equal(x0[rows], g["x0"])
    traj = bt.make_ref(d["x"], d["u"])
    K = bt.lqr_gains(traj)
    Xt, Ut = bt.lqr_track(traj, K, soa(x0))
    X = Xt.batch_major()[rows].cpu().numpy()
    U = Ut.batch_major()[rows].cpu().numpy()
    assert rel_err(X[:, ::10], g["x_track_10"]) < TOL
    assert rel_err(U[:, ::10], g["u_track_10"]) < TOL
    with np.errstate(all="ignore"):
        assert rel_err(X.sum(axis=1), g["x_sum"]) < TOL and rel_err(U.sum(axis=1), g["u_sum"]) < TOL


# ------------------------------------------------------------------------------------- config 4
@pytest.mark.parametrize("H", [50, 100, 200])
def test_c4_sixteen_problems_500_steps(bt, H):
    """Config 4 (B = 16 384 acrobots, 500 receding-horizon steps): 16 sampled problems against the oracle's Riccati
    restatement of solve_mpc_tracking (tt:8-69), with the shared reference (gains computed once per time step) and with
    per-problem references (16 different converged swing-up trajectories spread over the batch, every problem its own
    500 sweeps).  Parity against CasADi/IPOPT itself is unpinned."""
    d = golden("acrobot_optimal_trajectory")
    gp = golden("p_inf")
    Bn = 16384
    rng = np.random.default_rng(3)
    x0 = d["x"][0] + rng.uniform(-0.1, 0.1, (Bn, 4))
    x0[0] = d["x"][0] + 0.1  # main.py:127
    w = bt.mpc_weights()
    QT = dev(gp["P_inf"])
    rows = sample_rows(Bn, 16, 4)
    # shared reference
    ref = bt.make_ref(d["x"], d["u"])
    Xr, Ur, K0, ns = bt.mpc_track(soa(x0), ref, QT, T=501, T_pred=H, w=w)
    assert ns == 500
    xo, uo, K0o, QTo = O.solve_mpc_tracking(x0[rows], d["x"], d["u"], 501, T_pred=H, return_gains=True)
    assert rel_err(gp["P_inf"], QTo) < TOL
    assert rel_err(K0.cpu().numpy().reshape(500, 2, 4), K0o) < TOL
    assert rel_err(Xr.batch_major()[rows].cpu().numpy(), xo) < TOL
    assert rel_err(Ur.batch_major()[rows].cpu().numpy(), uo) < TOL
    # per-problem references: problem b tracks the converged swing-up of config-2 problem rows[b % 16] (the 16 trajectories
    # the unmodified reference solved to convergence, make_golden.py c2conv): dynamically consistent and all different,
    # so every problem linearises about its own points and runs its own 500 sweeps
    gc = golden("newton_c2_converged")
    which = np.arange(Bn) % 16
    which[rows] = np.arange(16)  # the sampled problems cover all 16 references
    xs = torch.from_numpy(gc["x"]).cuda()[torch.from_numpy(which).cuda()]
    us = torch.from_numpy(gc["u"]).cuda()[torch.from_numpy(which).cuda()]
    refp = bt.Ref(bt.Traj.from_batch_major(xs), bt.Traj.from_batch_major(us))
    x0p = gc["x"][which, 0] + (x0 - d["x"][0])
    Xp, Up, _, nsp = bt.mpc_track(soa(x0p), refp, QT, T=501, T_pred=H, w=w)
    assert nsp == 500 * Bn
    res = pool_map(_oracle_mpc, [(x0p[b], gc["x"][which[b]], gc["u"][which[b]], H) for b in rows])
    Xp_s, Up_s = Xp.batch_major()[rows].cpu().numpy(), Up.batch_major()[rows].cpu().numpy()
    for i, (xo, uo) in enumerate(res):
        assert np.isfinite(xo).all()
        assert rel_err(Xp_s[i], xo) < TOL and rel_err(Up_s[i], uo) < TOL, rows[i]
    assert bool(torch.isfinite(Xp.data).all())


# ------------------------------------------------------------------------------------- config 5
def test_c5_256_base_iterates_vs_reference(bt, fa_ref):
    """Config 5: the step-size sweep (tg:257-264) on 256 base iterates, iterate p = Newton iterate k_p (k_p = p mod 50,
    gamma_0 = 0.1) of config-2 problem p.  The iterates are produced here by the Newton kernel (one iteration per
    launch), K and sigma by acro_riccati_affine, the costs by acro_stepsize_sweep; the fixture holds what the unmodified
    reference computes for the same problems: 8 of the 200 step sizes, delta_J, max|sigma|, the final state of the
    base iterate."""
    g = golden("sweep_c5")
    x_ref, u_ref, _ = fa_ref
    rows, ks = g["rows"], g["k"]
    assert len(rows) == 256 and len(set(ks)) >= 40
    x0s = c2_x0()[rows]
    ref = bt.make_ref(x_ref, u_ref)
    w = bt.newton_weights()
    Xb = torch.empty(256, 501, 4, dtype=torch.float64, device="cuda")
    Ub = torch.empty(256, 500, 2, dtype=torch.float64, device="cuda")
    sel = torch.from_numpy(ks == 0).cuda()
    X0 = bt.rollout_open_loop(soa(x0s), None, N=501)
    Xb[sel] = X0.batch_major()[sel]
    Ub[sel] = 0.0
    st = None
    for k in range(1, int(ks.max()) + 1):
        st = bt.newton_solve(soa(x0s), ref, max_iters=50, tol=1e-4, gamma_0=0.1, state=st, chunk_iters=1)
        sel = torch.from_numpy(ks == k).cuda()
        if bool(sel.any()):
            Xb[sel] = st.X.batch_major()[sel]
            Ub[sel] = st.U.batch_major()[sel]
    assert bool((st.iters == int(ks.max())).all())
    X, U = bt.Traj.from_batch_major(Xb), bt.Traj.from_batch_major(Ub)
    K, S, dJ, sn = bt.riccati_affine(X, U, ref, w)
    cost = bt.stepsize_sweep(X, U, K, S, ref, w, dev(g["steps"])).cpu().numpy()  # (8, 256)
    assert rel_err(Xb[:, -1].cpu().numpy(), g["x_T"]) < TOL
    assert rel_err(dJ.cpu().numpy(), g["delta_J"]) < TOL
    assert rel_err(sn.cpu().numpy(), g["sigma_norm"]) < TOL
    for p in range(256):
        assert rel_err(cost[:, p], g["costs"][p]) < TOL, (p, ks[p])
    # the full 200-point curve of every iterate: its value at the sampled step sizes is the one checked above
    full = bt.stepsize_sweep(X, U, K, S, ref, w, dev(np.linspace(0, 1.25, 200))).cpu().numpy()
    assert np.array_equal(full[g["gamma_idx"]], cost)


# ------------------------------------------------------------------------------------- a real 10 000-step horizon
def _oracle_newton_long(args):
    x0, x_ref, u_ref, kw = args
    x, u, K, s, h = O.newton_Algorithm(x0, x_ref, u_ref, **kw)
    return x[::50], u[::50], np.array(s)[::50], h


@pytest.mark.parametrize("Bn,kernel", [(96, "auto"), (96, "ring4"), (96, "ring-rl")])
def test_long_horizon_vs_oracle(bt, fa_ref, Bn, kernel):
    """N = 10 001 (the 10k-step horizon BASELINE.json's target is quoted in; bench.py's `long_horizon` block): the
    swing-up reference followed by the upright equilibrium, 2 Newton iterations, 3 sampled problems against the oracle
    (every 50th step of x, u, sigma; costs; Armijo tries)."""
    x_fa, u_fa, _ = fa_ref
    N = 10001
    x_ref = np.vstack([x_fa, np.repeat(np.array([[np.pi, 0.0, 0.0, 0.0]]), N - x_fa.shape[0], 0)])
    u_ref = np.vstack([u_fa[:500], np.zeros((N - 1 - 500, 2))])
    x0s = np.random.default_rng(77).uniform(-0.2, 0.2, (Bn, 4))
    kw = dict(max_iters=2, tol=1e-4, gamma_0=0.1)
    st = bt.newton_solve(soa(x0s), bt.make_ref(x_ref, u_ref), kernel=kernel, **kw)
    torch.cuda.synchronize()
    rows = [0, 31, Bn - 1]
    X, U, S = (t.batch_major()[rows].cpu().numpy() for t in (st.X, st.U, st.S))
    res = pool_map(_oracle_newton_long, [(x0s[b], x_ref, u_ref, kw) for b in rows])
    for i, (x, u, s, h) in enumerate(res):
        assert int(st.iters[rows[i]]) == h["iters"] and list(st.hist_ntry[:2, rows[i]].cpu().numpy()) == h["n_try"]
        assert rel_err(st.hist_cost[:3, rows[i]].cpu().numpy(), h["cost"]) < TOL
        assert rel_err(X[i][::50], x) < TOL and rel_err(U[i][::50], u) < TOL and rel_err(S[i][::50], s) < TOL
