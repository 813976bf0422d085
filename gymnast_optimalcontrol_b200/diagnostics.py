"""What the callers AFTER the hot path consume: trajectory files in the reference's key
layout, and the numbers behind the reference's figures - without matplotlib (plotting itself is out of scope).

    save_optimal_trajectory / load_optimal_trajectory        main.py:74-79, main.py:101,124    keys x, u, t
    save_fully_actuated_reference / load_...                 fully_actuated_ref_gen.py:210-217, tg:513    keys x, u, time, T, N
    report_graph_data(t_ref, x_ref, u_ref, x_opt, u_opt, history)   tg:405-509   the curves of the four report figures
    armijo_plot_data(...)                                    tg:254-296   the curve, the tangent, the Armijo line, the tried points
    tracking_plot_data(...)                                  main.py:147-185   states / inputs / errors of a tracking run
    link_positions, animation_frames                         animation.py:9-16, 18-77   joint and tip positions per frame

Everything is host-side formatting of arrays the kernels produced (NumPy in, NumPy out); batched inputs keep their
leading batch axis.
"""
import numpy as np


def _np(a):
    return a.detach().cpu().numpy() if hasattr(a, "detach") else np.asarray(a)


# ------------------------------------------------------------------------------------------ npz files
def save_optimal_trajectory(path, x, u, t):
    """np.savez(path, x=x_opt, u=u_opt, t=t_ref)  (main.py:74-79).  x (N,4) / u (N-1,2), or batched (B,N,4) / (B,N-1,2)."""
    np.savez(path, x=_np(x), u=_np(u), t=_np(t))


def load_optimal_trajectory(path):
    """-> x_ref, u_ref, t_ref as main.py:101-103, 124-126 read them."""
    d = np.load(path)
    return d["x"], d["u"], d["t"]


def save_fully_actuated_reference(path, x, u, time, T=None, N=None):
    """The file fully_actuated_ref_gen.py:210-217 writes and trajectory_generation.py:513 reads: x, u, time, T, N."""
    x, u, time = _np(x), _np(u), _np(time)
    np.savez(path, x=x, u=u, time=time, T=float(time[-1]) if T is None else T, N=x.shape[-2] if N is None else N)


def load_fully_actuated_reference(path, rescale=True):
    """rescale=True applies trajectory_generation.py:514-518 (tau_1 := 0, torques x 2) like get_fully_actuated_ref."""
    d = np.load(path)
    u = d["u"]
    if rescale:
        u2 = np.zeros(u.shape)
        u2[..., 1] = u[..., 1]
        u = np.multiply(u2, 2)
    return d["x"], u, d["time"]


# ------------------------------------------------------------------------------------------ report figures
def iterations_to_show(num_iters):
    """The iterates the reference draws in its 'intermediate trajectories' figure (tg:438-449)."""
    fixed = [i for i in (0, 1, 5, 10, 100) if 0 <= i < num_iters]
    even = np.linspace(0, num_iters - 1, 5, dtype=int).tolist()
    return sorted(set(fixed + even))


def report_graph_data(t_ref, x_ref, u_ref, x_opt, u_opt, history):
    """The series of generate_report_graphs (tg:405-509) as arrays: one dict per figure.

    history: what newton_Algorithm(..., return_history=True) returns ('cost', 'sigma_norm', 'x_trajs', 'sigmas')."""
    t_ref, x_ref, u_ref, x_opt, u_opt = (_np(a) for a in (t_ref, x_ref, u_ref, x_opt, u_opt))
    u_ref_plot = u_ref[:-1] if u_ref.shape[0] == len(t_ref) else u_ref  # tg:407-410
    out = {"optimal_vs_desired": {
        "t": t_ref, "theta1_opt": x_opt[:, 0], "theta2_opt": x_opt[:, 1], "theta1_des": x_ref[:, 0], "theta2_des": x_ref[:, 1],
        "t_u": t_ref[:-1], "tau1_des": u_ref_plot[:, 0], "tau2_des": u_ref_plot[:, 1], "tau1_opt": u_opt[:, 0], "tau2_opt": u_opt[:, 1]}}
    xs = history.get("x_trajs", [])
    if len(xs):
        show = iterations_to_show(len(xs))
        out["intermediate_trajectories"] = {"t": t_ref, "iterations": show, "theta1": np.array([_np(xs[i])[:, 0] for i in show]),
                                            "theta2": np.array([_np(xs[i])[:, 1] for i in show]),
                                            "theta1_des": x_ref[:, 0], "theta2_des": x_ref[:, 1]}
    sg = history.get("sigmas", [])
    if len(sg):
        it = [i for i in (0, 1, 2, len(sg) - 1) if 0 <= i < len(sg)]  # tg:470-472
        out["descent_direction"] = {"t": t_ref[:-1], "iterations": it, "sigma_tau2": np.array([np.vstack(_np(sg[i]))[:, 1] for i in it])}
    cost = np.asarray(history["cost"], dtype=float)
    out["convergence"] = {"iteration_cost": np.arange(len(cost)), "cost": cost,
                          "iteration_sigma": np.arange(1, len(history["sigma_norm"]) + 1),
                          "sigma_norm": np.asarray(history["sigma_norm"], dtype=float)}
    return out


def armijo_plot_data(steps, costs, cost_k, delta_J, gamma_acc, stepsizes_tested, costs_tested, c=0.5):
    """plot_armijo_line_search (tg:254-296): the cost along the search direction (steps, costs from `armijo_sweep`), the
    tangent J_k + gamma dJ, the Armijo line J_k + c gamma dJ, the candidates the line search tried and the accepted one."""
    steps = _np(steps)
    return {"steps": steps, "cost_curve": _np(costs), "tangent": cost_k + steps * delta_J, "armijo_line": cost_k + c * steps * delta_J,
            "tested_steps": np.asarray(stepsizes_tested, dtype=float), "tested_costs": np.asarray(costs_tested, dtype=float),
            "accepted_step": float(gamma_acc)}


def tracking_plot_data(x_ref, u_ref, x_track, u_track, t_ref):
    """plot_tracking(x_ref, u_ref, x_opt, u_opt, t_ref) (main.py:147-185; same argument order): the curves of the two
    figures - tracked against reference angles, velocities and torques, the state error norm ||x - x_ref||_2 per time
    step and the control error |tau_2 - tau_2_ref| padded with one NaN (main.py:178-181).  The first control error of the
    shipped LQR / MPC figures (2.5 and 0.81) is `control_error[..., 0]`."""
    x_ref, u_ref, x_track, u_track, t_ref = (_np(a) for a in (x_ref, u_ref, x_track, u_track, t_ref))
    nu_ = min(u_ref.shape[-2], u_track.shape[-2])
    ce = np.abs(u_track[..., :nu_, 1] - u_ref[..., :nu_, 1])
    pad = np.full(ce.shape[:-1] + (1,), np.nan)
    return {"t": t_ref, "t_control": t_ref[:-1], "x_track": x_track, "x_ref": x_ref, "u_track": u_track[..., :nu_, :],
            "u_ref": u_ref[..., :nu_, :], "state_error": np.linalg.norm(x_track - x_ref, axis=-1),
            "control_error": np.concatenate([ce, pad], axis=-1)}


# ------------------------------------------------------------------------------------------ animation
def link_positions(theta1, theta2, l1=1.0, l2=1.0):
    """(x, y) of the base, the elbow and the tip (animation.py:9-16); theta arrays broadcast."""
    theta1, theta2 = _np(theta1), _np(theta2)
    x1, y1 = l1 * np.sin(theta1), -l1 * np.cos(theta1)
    x2, y2 = x1 + l2 * np.sin(theta1 + theta2), y1 - l2 * np.cos(theta1 + theta2)
    z = np.zeros_like(x1)
    return np.stack([z, x1, x2], axis=-1), np.stack([z, y1, y2], axis=-1)


def animation_frames(x_opt, x_ref, x_e1=None, x_e2=None, l1=1.0, l2=1.0):
    """The point sets create_and_save_animation (animation.py:18-77) draws: per frame the three joints of the optimal
    acrobot and of the reference 'ghost', the tip trace, and the tip positions of the two equilibria."""
    x_opt, x_ref = _np(x_opt), _np(x_ref)
    ox, oy = link_positions(x_opt[..., 0], x_opt[..., 1], l1, l2)
    rx, ry = link_positions(x_ref[..., 0], x_ref[..., 1], l1, l2)
    out = {"opt_x": ox, "opt_y": oy, "ref_x": rx, "ref_y": ry, "trace_x": ox[..., 2], "trace_y": oy[..., 2],
           "limits": (-(l1 + l2) * 1.1, (l1 + l2) * 1.1)}
    for name, xe in (("e1", x_ref[..., 0, :] if x_e1 is None else _np(x_e1)), ("e2", x_ref[..., -1, :] if x_e2 is None else _np(x_e2))):
        ex, ey = link_positions(xe[..., 0], xe[..., 1], l1, l2)
        out[name + "_tip"] = (ex[..., 2], ey[..., 2])
    return out
