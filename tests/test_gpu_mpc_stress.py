"""Randomised configurations of the two reworked MPC kernels against the oracle: horizons from 2 to 40, T not a multiple
of the pass width, windows that run past the end of the reference, batch sizes around the warp size, boxes that bind
rarely / often / almost always, shared and per-problem references."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err
from oracle import acro_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bt():
    from gymnast_optimalcontrol_b200 import batched
    return batched


def _soa(bt, a):
    return bt.pack_soa(torch.from_numpy(np.ascontiguousarray(a)).cuda())


def _aos(bt, t):
    return bt.unpack_soa(t).cpu().numpy()


@pytest.mark.parametrize("seed", range(30))
def test_random_mpc_configurations(bt, seed):
    d = golden("acrobot_optimal_trajectory")
    g = golden("p_inf")
    rng = np.random.default_rng(1000 + seed)
    H = int(rng.choice([2, 3, 4, 5, 7, 12, 19, 26, 40]))
    N_ = int(rng.integers(12, 45))
    T = int(rng.integers(max(3, N_ - 6), N_ + 1))
    t0 = int(rng.choice([0, 90, 160, 330, 501 - N_]))
    n = int(rng.choice([1, 5, 31, 33, 64]))
    tau = float(rng.choice([6.0, 12.0, 18.0, 30.0]))
    xs, us = d["x"][t0:t0 + N_], d["u"][t0:t0 + N_ - 1]
    x0 = xs[0] + rng.uniform(-0.02, 0.02, (n, 4))
    QT = torch.from_numpy(g["P_inf"]).cuda()
    w = bt.mpc_weights()
    shared = bt.make_ref(xs, us)
    xsp = np.repeat(xs[None], n, 0) + rng.uniform(-1e-3, 1e-3, (n, N_, 4))
    usp = np.repeat(us[None], n, 0) + rng.uniform(-1e-2, 1e-2, (n, N_ - 1, 2)) * np.array([0.0, 1.0])
    perp = bt.Ref(_soa(bt, xsp), _soa(bt, usp))
    picks = sorted(set([0, n // 2, n - 1]))

    def prefix(xo, T_):  # compare while the closed loop has not run away (short horizons do not stabilise the plant)
        big = np.where(~(np.abs(xo[:T_]).max(axis=1) < 20.0))[0]
        return T_ if len(big) == 0 else int(big[0])

    # ---- unconstrained, per-problem references (several solves per pass + remainders)
    Xp, Up, _, ns = bt.mpc_track(_soa(bt, x0), perp, QT, T=T, T_pred=H, w=w)
    assert ns == (T - 1) * n
    Xp, Up = _aos(bt, Xp), _aos(bt, Up)
    for b in picks:
        xo, uo = O.solve_mpc_tracking(x0[b], xsp[b], usp[b], T, T_pred=H)
        m = prefix(xo, T)
        tol = 1e-9 if m == T else 1e-7
        assert m >= min(T, 4) and rel_err(Xp[b][:m], xo[:m]) < tol and rel_err(Up[b][:m - 1], uo[:m - 1]) < tol, (H, N_, T, t0, n)
    # ---- with the input box, shared and per-problem references
    for ref, xr_b, ur_b in ((shared, lambda b: xs, lambda b: us), (perp, lambda b: xsp[b], lambda b: usp[b])):
        Xb, Ub, info = bt.mpc_track_box(_soa(bt, x0), ref, QT, tau_max=tau, T=T, T_pred=H, w=w)
        assert int(info["status"].max()) == 0
        Xb, Ub = _aos(bt, Xb), _aos(bt, Ub)
        na = info["n_active"].cpu().numpy()
        for b in picks:
            ub = ur_b(b)
            # the box must admit the trivial part of the QP: the first input stays at the point closest to zero
            xo, uo, nao = O.solve_mpc_tracking_box(x0[b], xr_b(b), ub, T, T_pred=H, tau_max=tau, Q_T=g["P_inf"])
            m = prefix(xo, T)
            tol = 1e-9 if m == T else 1e-6
            assert m >= min(T, 4), (H, N_, T, t0, n, tau)
            assert rel_err(Xb[b][:m], xo[:m]) < tol and rel_err(Ub[b][:m - 1], uo[:m - 1]) < tol, (H, N_, T, t0, n, tau)
            assert np.abs(na[:m - 1, b] - nao[:m - 1]).max() <= 1
