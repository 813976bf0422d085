"""k_mpc_track_box: time against horizon and bound (tau = 1e6: the box never binds) - where does a solve spend its time?"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gymnast_optimalcontrol_b200 import batched as bt  # noqa: E402

opt = np.load(os.path.join(ROOT, "tests", "golden", "acrobot_optimal_trajectory.npz"))
N = 501
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
traj = bt.make_ref(opt["x"], opt["u"])
w = bt.mpc_weights()
xf = bt.upload(np.array([[np.pi], [0], [0], [0]], dtype=np.float64))
A_f, B_f = bt.linearize(xf, bt.upload(np.zeros((2, 1))), discrete=True)
P, n = bt.p_inf(A_f, B_f, w)
QT = P[:, :, 0].contiguous()
x0 = bt.upload(np.ascontiguousarray((opt["x"][0] + np.random.default_rng(3).uniform(-0.1, 0.1, (B, 4))).T))
for tau in (18.0, 1e6):
    for H in (10, 25, 50, 75):
        for T in (N,):
            ts = []
            for rep in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                Xr, Ur, info = bt.mpc_track_box(x0, traj, QT, tau_max=tau, T=T, T_pred=H, w=w)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            sw = info["n_sweeps"].double()
            na = info["n_active"]
            print("B %d tau %g H %d: %.2f ms (%.2f first), %.3g solves/s, sweeps/solve mean %.3f max-problem %.3f, steps with active bounds %.3f, "
                  "max active %d" % (B, tau, H, min(ts), ts[0], B * (T - 1) / min(ts) * 1e3, sw.mean().item() / (T - 1),
                                     sw.max().item() / (T - 1), (na > 0).double().mean().item(), int(na.max().item())), flush=True)
