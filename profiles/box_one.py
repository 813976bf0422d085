"""One launch of k_mpc_track_box (config 4 with the input box: B = 16384, horizon 75, |u| <= 18, shared reference) for ncu."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gymnast_optimalcontrol_b200 import batched as bt  # noqa: E402

opt = np.load(os.path.join(ROOT, "tests", "golden", "acrobot_optimal_trajectory.npz"))
N, B, H = 501, 16384, 75
traj = bt.make_ref(opt["x"], opt["u"])
w = bt.mpc_weights()
xf = bt.upload(np.array([[np.pi], [0], [0], [0]], dtype=np.float64))
A_f, B_f = bt.linearize(xf, bt.upload(np.zeros((2, 1))), discrete=True)
P, n = bt.p_inf(A_f, B_f, w)
x0 = bt.upload(np.ascontiguousarray((opt["x"][0] + np.random.default_rng(3).uniform(-0.1, 0.1, (B, 4))).T))
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    Xr, Ur, info = bt.mpc_track_box(x0, traj, P[:, :, 0].contiguous(), tau_max=18.0, T=N, T_pred=H, w=w)
torch.cuda.synchronize()
print("sweeps per solve", info["n_sweeps"].double().mean().item() / (N - 1))
