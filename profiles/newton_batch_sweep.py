"""Newton iterations/s versus batch size for the Newton kernels.
Columns: default dispatch (duo <= 148 tiles, ring with 16/4/2-step stages beyond), and forced variants."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
from gymnast_optimalcontrol_b200 import batched as bt
d = np.load('tests/golden/fully_actuated_trajectory.npz')
u_ref = np.zeros(d['u'].shape); u_ref[:, 1] = 2 * d['u'][:, 1]
ref = bt.make_ref(d['x'], u_ref)
VARIANTS = (("default", "auto"), ("duo", "duo"), ("duo4", "duo4"), ("ring16", "ring16"), ("ring4", "ring4"), ("ring-rl", "ring-rl"),
            ("ldg", "ldg"), ("spec", "spec"))
sizes = [int(v) for v in sys.argv[1:]] or [2048, 4096, 4736, 8192, 9472, 16384, 18944, 32768, 37888, 65536, 131072]
for B in sizes:
    out = []
    for name, k in VARIANTS:
        if name in ("duo", "duo4") and B > 4736 * 4:
            out.append(float('nan')); continue
        if name in ("ring16", "spec") and B > 4736:
            out.append(float('nan')); continue
        x0 = torch.from_numpy(np.random.default_rng(1).uniform(-0.2, 0.2, (4, B))).cuda()
        st = bt.newton_alloc(B, 501, 10, history=False)
        def run():
            st.initialised = False
            bt.newton_solve(x0, ref, max_iters=10, tol=0.0, gamma_0=0.1, state=st, kernel=k)
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
        out.append(B * 20 / (e0.elapsed_time(e1) * 1e-3))
        del st
    print("B=%6d  " % B + "  ".join("%s %.2fM" % (v[0], o / 1e6) for v, o in zip(VARIANTS, out)), flush=True)
