"""One launch of k_mpc_track_pp (config 4: B = 16384, horizon 75, per-problem references) for ncu; prints its time."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gymnast_optimalcontrol_b200 import batched as bt  # noqa: E402

opt = np.load(os.path.join(ROOT, "tests", "golden", "acrobot_optimal_trajectory.npz"))
N, B = 501, int(sys.argv[1]) if len(sys.argv) > 1 else 16384
H = int(sys.argv[2]) if len(sys.argv) > 2 else 75
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w = bt.mpc_weights()
xf = bt.upload(np.array([[np.pi], [0], [0], [0]], dtype=np.float64))
A_f, B_f = bt.linearize(xf, bt.upload(np.zeros((2, 1))), discrete=True)
P, n = bt.p_inf(A_f, B_f, w)
QT = P[:, :, 0].contiguous()
x0 = bt.upload(np.ascontiguousarray((opt["x"][0] + np.random.default_rng(3).uniform(-0.1, 0.1, (B, 4))).T))
refp = bt.Ref(bt.Traj.from_batch_major(bt.upload(np.repeat(opt["x"][None], B, 0))),
              bt.Traj.from_batch_major(bt.upload(np.repeat(opt["u"][None], B, 0))))
for rep in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    Xr, Ur, K0, ns = bt.mpc_track(x0, refp, QT, T=N, T_pred=H, w=w)
    e1.record()
    torch.cuda.synchronize()
    print("B %d H %d: %.2f ms, %.3g solves/s" % (B, H, e0.elapsed_time(e1), ns / e0.elapsed_time(e1) * 1e3), flush=True)
