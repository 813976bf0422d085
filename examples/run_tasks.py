#!/usr/bin/env python
"""The four task recipes of the reference's main.py (task_1 .. task_4, main.py:28-145) on the GPU drop-ins, followed by
their batched counterparts (configs 2-4 of BASELINE.json).  Plotting / animation are not part of this package.

    python examples/run_tasks.py            (on a B200; about ten seconds)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gymnast_optimalcontrol_b200 import trajectory_generation as tg  # noqa: E402
from gymnast_optimalcontrol_b200 import trajectory_tracking as tt  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def timed(label, f):
    t = time.perf_counter()
    out = f()
    print("  [%s: %.3f s]" % (label, time.perf_counter() - t))
    return out


def task_1():
    """main.py:28-49: step reference between two equilibria, gamma_0 = 0.05."""
    x_e1, u_e1 = tg.compute_equilibrium(np.array([0.0, 0.0]), (0.1, -0.1))
    x_e2, u_e2 = tg.compute_equilibrium(np.array([0.5, 0.5]), (0.35, -0.35))
    t_ref, x_ref, u_ref = tg.define_reference_piecewise(10.0, x_e1, x_e2, u_e1, u_e2)
    x, u, K, s, h = timed("task_1", lambda: tg.newton_Algorithm(x_e1.copy(), x_ref, u_ref, max_iters=5000, tol=1e-4,
                                                                gamma_0=0.05, plot_armijo_iters=5, verbose=False))
    print("task_1: %d iterations, final cost %.9f (reference run: 173, 27.48962661922637)" % (h["iters"], h["cost"][-1]))


def task_2():
    """main.py:52-96: swing-up towards the fully-actuated reference, gamma_0 = 0.1."""
    x_ref, u_ref, t_ref = tg.get_fully_actuated_ref(os.path.join(G, "fully_actuated_trajectory.npz"))
    x, u, K, s, h = timed("task_2", lambda: tg.newton_Algorithm(np.zeros(4), x_ref, u_ref, max_iters=5000, tol=1e-4,
                                                                gamma_0=0.1, plot_armijo_iters=7, verbose=False))
    d = np.load(os.path.join(G, "acrobot_optimal_trajectory.npz"))
    print("task_2: %d iterations, final cost %.8f (reference run: 393, 28063.21834988143); max |x - shipped npz| = %.1e"
          % (h["iters"], h["cost"][-1], np.abs(x - d["x"]).max()))
    return d["x"], d["u"], d["t"]


def task_3(x_opt, u_opt, t_ref):
    """main.py:99-119: LQR tracking from perturbed initial conditions."""
    for dx in (0.2, 0.3):
        xt, ut = tt.LQR_tracking(x_opt, u_opt, t_ref, x0_perturbed=x_opt[0] + dx)
        print("task_3: dx = %.1f, final state error %.2e, first control error %.3f" % (dx, np.abs(xt[-1] - x_opt[-1]).max(),
                                                                                     abs(ut[0, 1] - u_opt[0, 1])))


def task_4(x_opt, u_opt, t_ref):
    """main.py:122-145: receding-horizon MPC tracking (horizon 75), then with the input box of tt:87-91 switched on."""
    x0 = x_opt[0] + 0.1
    xr, ur = timed("task_4", lambda: tt.solve_mpc_tracking(x0, x_opt, u_opt, len(t_ref)))
    print("task_4: final state error %.2e, first control error %.3f (figures/mpc/tracking_dx_0.1_err.png: ~0.81)"
          % (np.abs(xr[-1] - x_opt[-1]).max(), abs(ur[0, 1] - u_opt[0, 1])))
    xb, ub, info = tt.solve_mpc_tracking(x0, x_opt, u_opt, len(t_ref), tau_max=18.0, return_info=True)
    print("task_4 with |u| <= 18: max |u| = %.3f (unconstrained %.3f), final state error %.2e, %d of 500 steps with active bounds"
          % (np.abs(ub).max(), np.abs(ur).max(), np.abs(xb[-1] - x_opt[-1]).max(), int((info["n_active"] > 0).sum())))


def batched(x_opt, u_opt, t_ref):
    rng = np.random.default_rng(1)
    x_ref, u_ref, _ = tg.get_fully_actuated_ref(os.path.join(G, "fully_actuated_trajectory.npz"))
    x0 = rng.uniform(-0.2, 0.2, (4096, 4))
    x, u, K, s, h = timed("4096 swing-up solves to convergence", lambda: tg.newton_Algorithm(
        x0, x_ref, u_ref, max_iters=5000, tol=1e-4, gamma_0=0.1, verbose=False))
    print("config 2: %d problems, iterations %d..%d, all converged: %s" % (len(x0), h["iters"].min(), h["iters"].max(),
                                                                           bool((h["status"] == 1).all())))
    x0 = x_opt[0] + rng.uniform(-0.3, 0.3, (65536, 4))
    xt, ut = timed("65536 LQR-tracked rollouts", lambda: tt.LQR_tracking(x_opt, u_opt, t_ref, x0_perturbed=x0))
    ok = np.isfinite(xt[:, -1]).all(axis=1)
    print("config 3: %.1f %% of the rollouts stay finite, median final error %.2e" % (100 * ok.mean(),
                                                                                    np.median(np.abs(xt[ok, -1] - x_opt[-1]).max(axis=1))))
    x0 = x_opt[0] + rng.uniform(-0.1, 0.1, (16384, 4))
    xr, ur = timed("16384 MPC-tracked acrobots, horizon 75", lambda: tt.solve_mpc_tracking(x0, x_opt, u_opt, len(t_ref)))
    print("config 4: median final error %.2e" % np.median(np.abs(xr[:, -1] - x_opt[-1]).max(axis=1)))


if __name__ == "__main__":
    task_1()
    traj = task_2()
    task_3(*traj)
    task_4(*traj)
    batched(*traj)
