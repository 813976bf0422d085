"""Speculative Armijo (k_newton_spec) against k_newton_duo on config 2 (B = 4096, N = 501, 50 iterations):
Newton iterations/s at gamma_0 = 0.1 (one try per iteration) and gamma_0 = 1 (back-tracking), for the adaptive
and the fixed numbers of candidates per round; the Armijo decisions of every variant must equal those of k_newton_duo.
Every configuration runs in its own process under a 90 s limit (a deadlocked kernel must not take the others down)."""
import hashlib
import subprocess
import sys

CONFIGS = [(0.1, "duo"), (0.1, "spec1"), (0.1, "spec"), (0.1, "spec4"), (1.0, "duo"), (1.0, "spec1"), (1.0, "spec"),
           (1.0, "spec2"), (1.0, "spec4"), (1.0, "spec6"), (1.0, "spec8")]


def one(g0, name, B, iters):
    import numpy as np, torch
    sys.path.insert(0, '.')
    from gymnast_optimalcontrol_b200 import batched as bt
    d = np.load('tests/golden/fully_actuated_trajectory.npz')
    u_ref = np.zeros(d['u'].shape); u_ref[:, 1] = 2 * d['u'][:, 1]
    ref = bt.make_ref(d['x'], u_ref)
    x0 = bt.upload(np.ascontiguousarray(np.random.default_rng(1).uniform(-0.2, 0.2, (B, 4)).T))
    kw = dict(kernel=name) if not (name.startswith("spec") and name[4:].isdigit()) else dict(kernel="spec", speculate=int(name[4:]))
    st = bt.newton_alloc(B, 501, iters, history=True)

    def run():
        st.reset()
        bt.newton_solve(x0, ref, max_iters=iters, tol=0.0, gamma_0=g0, state=st, **kw)
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    done = int(st.iters.sum())
    nt = st.hist_ntry[:iters]
    h = hashlib.md5()
    for t in (st.hist_ntry, torch.nan_to_num(st.hist_gamma), st.status, st.iters):
        h.update(t.cpu().numpy().tobytes())
    hx = hashlib.md5(st.X.data.cpu().numpy().tobytes()).hexdigest()[:8]
    tm = nt.reshape(iters, -1, 32).max(dim=2).values.double().mean().item() if B % 32 == 0 else float('nan')
    print("gamma_0=%.1f %-6s %8.2f ms  %6.2f M it/s  mean tries %.2f  mean tile-max %.2f  status %s  decisions %s  X %s  cost %.6f"
          % (g0, name, ms, done / ms / 1e3, nt.double().mean().item(), tm, sorted(set(st.status.tolist())), h.hexdigest()[:8], hx,
             float(st.cost.double().mean())), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        one(float(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5]))
    else:
        B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
        iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
        only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
        for g0, name in CONFIGS:
            if only and name not in only:
                continue
            try:
                r = subprocess.run([sys.executable, __file__, "--one", str(g0), name, str(B), str(iters)], capture_output=True, text=True, timeout=90)
                print(r.stdout.strip() or ("gamma_0=%.1f %-6s FAILED: %s" % (g0, name, r.stderr.strip()[-300:])), flush=True)
            except subprocess.TimeoutExpired:
                print("gamma_0=%.1f %-6s TIMEOUT (deadlock?)" % (g0, name), flush=True)
