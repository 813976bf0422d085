"""Where a pass of k_newton_spec spends its time (needs a library built with -DACRO_SPEC_TIMING, loaded through
ACRO_B200_LIB): block 0 logs clock64 after barrier A, when the pass function returns and after barrier B, for warp 0
(chain / control) and warp 1 (trailer)."""
import sys
import numpy as np, torch
sys.path.insert(0, '.')
from gymnast_optimalcontrol_b200 import batched as bt
d = np.load('tests/golden/fully_actuated_trajectory.npz')
u_ref = np.zeros(d['u'].shape); u_ref[:, 1] = 2 * d['u'][:, 1]
ref = bt.make_ref(d['x'], u_ref)
B, iters = 4096, 12
g0 = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
spec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
x0 = bt.upload(np.ascontiguousarray(np.random.default_rng(1).uniform(-0.2, 0.2, (B, 4)).T))
tbuf = torch.zeros(2 * 2000 * 4, dtype=torch.float64, device='cuda')
st = bt.newton_alloc(B, 501, iters, history=True)
for rep in range(2):
    st.reset(); tbuf.zero_()
    bt.newton_solve(x0, ref, max_iters=iters, tol=0.0, gamma_0=g0, state=st, kernel="spec", speculate=spec, params_b=tbuf)
    torch.cuda.synchronize()
t = tbuf.view(torch.int64).cpu().numpy().reshape(2, 2000, 4)
names = {1: "backward", 2: "forward(duo)", 3: "forward(spec)", 4: "commit", 0: "exit"}
w0 = t[0][t[0][:, 1] != 0]; w1 = t[1][t[1][:, 1] != 0]
print("warp 0: cmd | work (A -> function returned) | fence + wait at B | logic until next A   [cycles]")
for i in range(min(len(w0), 40)):
    nxt = w0[i + 1][1] - w0[i][3] if i + 1 < len(w0) else 0
    print("  %-14s %9d %9d %9d" % (names.get(int(w0[i][0]) % 16, "?"), w0[i][2] - w0[i][1], w0[i][3] - w0[i][2], nxt))
print("warp 1: cmd | work | fence + wait at B | wait at next A")
for i in range(min(len(w1), 40)):
    nxt = w1[i + 1][1] - w1[i][3] if i + 1 < len(w1) else 0
    print("  %-14s %9d %9d %9d" % (names.get(int(w1[i][0]) % 16, "?"), w1[i][2] - w1[i][1], w1[i][3] - w1[i][2], nxt))
tot = w0[-1][3] - w0[0][1]
print("total cycles %d for %d passes; sum of work %d, B-waits %d" % (tot, len(w0), (w0[:, 2] - w0[:, 1]).sum(), (w0[:, 3] - w0[:, 2]).sum()))
