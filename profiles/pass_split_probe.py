"""Time the backward (acro_riccati_affine) and forward (acro_closed_loop_rollout_cost) passes separately at B=4096."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from gymnast_optimalcontrol_b200 import batched as bt
d = np.load('tests/golden/fully_actuated_trajectory.npz')
u_ref = np.zeros(d['u'].shape); u_ref[:, 1] = 2 * d['u'][:, 1]
ref = bt.make_ref(d['x'], u_ref); w = bt.newton_weights()
for B in (4096, 65536):
    x0 = torch.from_numpy(np.random.default_rng(1).uniform(-0.2, 0.2, (4, B))).cuda()
    st = bt.newton_solve(x0, ref, max_iters=3, tol=0.0, gamma_0=0.1, history=False)
    gam = torch.tensor([0.1], dtype=torch.float64, device='cuda')
    def timeit(f, n=5):
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): f()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    tb = timeit(lambda: bt.riccati_affine(st.X, st.U, ref, w))
    tf = timeit(lambda: bt.closed_loop_rollout_cost(st.X, st.U, st.K, st.S, ref, w, gam))
    tfs = timeit(lambda: bt.closed_loop_rollout_cost(st.X, st.U, st.K, st.S, ref, w, gam, store=True))
    tr = timeit(lambda: bt.rollout_open_loop(x0, None, N=501))
    print("B=%d backward %.3f ms, forward(cost only) %.3f ms, forward(store) %.3f ms, open-loop rollout %.3f ms; cycles/step @1.965GHz: bwd %.0f fwd %.0f rollout %.0f"
          % (B, tb, tf, tfs, tr, tb * 1e-3 * 1.965e9 / 500, tfs * 1e-3 * 1.965e9 / 500, tr * 1e-3 * 1.965e9 / 500))
