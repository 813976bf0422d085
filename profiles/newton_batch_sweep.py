"""Newton iterations/s versus batch size for the two kernels (ACRO_NEWTON_KERNEL=ring|ldg)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from gymnast_optimalcontrol_b200 import batched as bt
d = np.load('tests/golden/fully_actuated_trajectory.npz')
u_ref = np.zeros(d['u'].shape); u_ref[:, 1] = 2 * d['u'][:, 1]
ref = bt.make_ref(d['x'], u_ref)
for B in (512, 1024, 2048, 4096, 8192, 16384, 28416, 32768, 65536):
    out = []
    for k in ("ring", "ldg"):
        os.environ["ACRO_NEWTON_KERNEL"] = k
        x0 = torch.from_numpy(np.random.default_rng(1).uniform(-0.2, 0.2, (4, B))).cuda()
        st = bt.newton_alloc(B, 501, 10, history=False)
        def run():
            st.initialised = False
            bt.newton_solve(x0, ref, max_iters=10, tol=0.0, gamma_0=0.1, state=st)
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
        out.append(B * 20 / (e0.elapsed_time(e1) * 1e-3))
    print("B=%6d  ring %.3e it/s   ldg %.3e it/s" % (B, out[0], out[1]), flush=True)
