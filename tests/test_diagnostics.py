"""Host-side adapters around the hot path: npz files in the reference's key layout and the
numbers behind the reference's figures.  CPU only.  Where /root/reference is present the adapters are compared with what
the UNMODIFIED reference hands to matplotlib (a recording stub stands in for pyplot)."""
import importlib.util
import os

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden, rel_err
from oracle import ref_import

spec = importlib.util.spec_from_file_location("acro_diagnostics", os.path.join(ROOT, "gymnast_optimalcontrol_b200", "diagnostics.py"))
dg = importlib.util.module_from_spec(spec)
spec.loader.exec_module(dg)  # (loaded by path: importing the package needs the CUDA library)


class Recorder:
    """Stands in for matplotlib.pyplot: remembers every call."""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def f(*a, **k):
            self.calls.append((name, a, k))
            return self
        return f

    def series(self, *names):
        return [c for c in self.calls if c[0] in names]


def _history():
    g = golden("newton_task2")
    return g, {"cost": list(g["cost"]), "sigma_norm": list(g["sigma_norm"]), "x_trajs": list(g["x_trajs"]),
               "sigmas": [list(s) for s in g["sigmas"]]}


def test_npz_round_trip_in_the_reference_key_layout(tmp_path):
    d = golden("acrobot_optimal_trajectory")
    p = str(tmp_path / "opt.npz")
    dg.save_optimal_trajectory(p, d["x"], d["u"], d["t"])
    z = np.load(p)
    assert sorted(z.files) == ["t", "u", "x"]  # main.py:74-79
    x, u, t = dg.load_optimal_trajectory(p)
    assert np.array_equal(x, d["x"]) and np.array_equal(u, d["u"]) and np.array_equal(t, d["t"])
    # the shipped file itself loads through the same reader
    x2, u2, t2 = dg.load_optimal_trajectory(os.path.join(GOLDEN, "acrobot_optimal_trajectory.npz"))
    assert np.array_equal(x2, d["x"])
    f = golden("fully_actuated_trajectory")
    p2 = str(tmp_path / "fa.npz")
    dg.save_fully_actuated_reference(p2, f["x"], f["u"], f["time"])
    z2 = np.load(p2)
    assert sorted(z2.files) == ["N", "T", "time", "u", "x"] and int(z2["N"]) == 501 and float(z2["T"]) == float(f["T"])
    xr, ur, tr = dg.load_fully_actuated_reference(p2)
    assert np.array_equal(ur[:, 0], np.zeros(len(ur))) and np.array_equal(ur[:, 1], 2 * f["u"][:, 1])  # tg:514-518
    # batched trajectories keep their leading axis
    dg.save_optimal_trajectory(p, np.repeat(d["x"][None], 3, 0), np.repeat(d["u"][None], 3, 0), d["t"])
    assert dg.load_optimal_trajectory(p)[0].shape == (3, 501, 4)


def test_report_graph_data_shapes_and_selection():
    g, h = _history()
    out = dg.report_graph_data(g["t_ref"], g["x_ref"], g["u_ref"], g["x"], g["u"], h)
    assert dg.iterations_to_show(394) == [0, 1, 5, 10, 98, 100, 196, 294, 393]
    assert out["intermediate_trajectories"]["iterations"] == [0, 1, 2, 3, 4]
    assert out["descent_direction"]["iterations"] == [0, 1, 2, 3]
    assert out["convergence"]["cost"].shape == (394,) and out["convergence"]["sigma_norm"].shape == (393,)
    assert out["optimal_vs_desired"]["tau2_des"].shape == (500,)  # u_ref has 501 rows in this recipe: trimmed (tg:407-410)


@pytest.mark.skipif(not ref_import.available(), reason="needs the unmodified reference")
def test_report_graph_data_is_what_the_reference_plots():
    rd, rtg, rtt = ref_import.load()
    g, h = _history()
    rec = Recorder()
    old = rtg.plt
    rtg.plt = rec
    try:
        rtg.generate_report_graphs(g["t_ref"], g["x_ref"], g["u_ref"], g["x"], g["u"], h)
    finally:
        rtg.plt = old
    out = dg.report_graph_data(g["t_ref"], g["x_ref"], g["u_ref"], g["x"], g["u"], h)
    plots = rec.series("plot")
    o = out["optimal_vs_desired"]
    for c, key in zip(plots[:4], ("theta1_opt", "theta2_opt", "theta1_des", "theta2_des")):
        assert np.array_equal(c[1][0], o["t"]) and np.array_equal(c[1][1], o[key])
    steps = rec.series("step")
    for c, key in zip(steps, ("tau1_des", "tau2_des", "tau1_opt", "tau2_opt")):
        assert np.array_equal(c[1][0], o["t_u"]) and np.array_equal(c[1][1], o[key])
    it = out["intermediate_trajectories"]
    n = len(it["iterations"])
    th1 = plots[4:4 + n]
    assert [c[2]["label"] for c in th1] == ["Iter %d" % i for i in it["iterations"]]
    for c, y in zip(th1, it["theta1"]):
        assert np.array_equal(c[1][1], y)
    th2 = plots[4 + n + 1:4 + 2 * n + 1]
    for c, y in zip(th2, it["theta2"]):
        assert np.array_equal(c[1][1], y)
    sg = plots[4 + 2 * n + 2:]
    for c, y in zip(sg, out["descent_direction"]["sigma_tau2"]):
        assert np.array_equal(c[1][0], out["descent_direction"]["t"]) and np.array_equal(c[1][1], y)
    semi = rec.series("semilogy")
    cv = out["convergence"]
    assert np.array_equal(np.asarray(list(semi[0][1][0])), cv["iteration_cost"]) and np.array_equal(semi[0][1][1], cv["cost"])
    assert np.array_equal(np.asarray(list(semi[1][1][0])), cv["iteration_sigma"]) and np.array_equal(semi[1][1][1], cv["sigma_norm"])


def test_tracking_plot_data_reproduces_the_shipped_figure_values():
    """figures/LQR/tracking_dx_0.2_err.png: first control error ~2.5 (main.py:178-181 pads the control error with a NaN)."""
    g = golden("lqr_tracking")
    d = golden("acrobot_optimal_trajectory")
    out = dg.tracking_plot_data(d["x"], d["u"], g["x_track"][0], g["u_track"][0], d["t"])
    assert out["state_error"].shape == (501,) and out["control_error"].shape == (501,) and np.isnan(out["control_error"][-1])
    assert abs(out["control_error"][0] - 2.506) < 1e-3
    assert abs(out["state_error"][0] - np.linalg.norm(np.full(4, 0.2))) < 1e-12
    many = dg.tracking_plot_data(d["x"], d["u"], g["x_track"][:5], g["u_track"][:5], d["t"])
    assert many["state_error"].shape == (5, 501) and many["control_error"].shape == (5, 501)


def test_animation_frames():
    d = golden("acrobot_optimal_trajectory")
    f = golden("fully_actuated_trajectory")
    fr = dg.animation_frames(d["x"], f["x"])
    assert fr["opt_x"].shape == (501, 3) and fr["ref_y"].shape == (501, 3)
    # hanging at rest: tip at (0, -2); upright: tip at (0, 2)   (animation.py:9-16)
    assert abs(fr["opt_y"][0, 2] + 2.0) < 1e-9 and abs(fr["opt_y"][-1, 2] - 2.0) < 1e-3
    x, y = dg.link_positions(0.3, -0.2)
    assert abs(x[1] - np.sin(0.3)) < 1e-15 and abs(y[2] - (-np.cos(0.3) - np.cos(0.1))) < 1e-15


def test_armijo_plot_data():
    g = golden("sweep_iter0")
    t2 = golden("newton_task2")
    out = dg.armijo_plot_data(g["steps"], g["costs"], t2["cost"][0], float(g["delta_J"]), 0.1, [0.1], [t2["cost"][1]])
    assert out["cost_curve"].shape == (200,) and out["tangent"][0] == t2["cost"][0]
    # the accepted step satisfies the Armijo condition it draws (tg:361)
    i = np.argmin(np.abs(g["steps"] - 0.1))
    assert t2["cost"][1] < t2["cost"][0] + 0.5 * 0.1 * float(g["delta_J"])
    assert out["armijo_line"][i] > g["costs"][i] - 1.0
