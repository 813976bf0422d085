"""Where the end-to-end time of a batched solve goes (one GPU): timeline of solve kernel / un-tiling + D2H on the copy
stream for the pipelined drop-in call, host time per call, and the same with return_gains=False.
Also dumps the Armijo tries of config 2 at gamma_0 = 1 (gpurun_out/ntry_gamma1.npz) for the speculation study."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
from gymnast_optimalcontrol_b200 import batched as bt
from gymnast_optimalcontrol_b200 import trajectory_generation as tg

d = np.load('tests/golden/fully_actuated_trajectory.npz')
u_ref = np.zeros(d['u'].shape); u_ref[:, 1] = 2 * d['u'][:, 1]
B = 4096
x0 = np.random.default_rng(1).uniform(-0.2, 0.2, (B, 4))
x0h = torch.from_numpy(x0).pin_memory(); xr = torch.from_numpy(np.ascontiguousarray(d['x'])).pin_memory(); ur = torch.from_numpy(u_ref).pin_memory()
kw = dict(max_iters=50, tol=0.0, gamma_0=0.1, verbose=False)

def timeline(n, **extra):
    evs, pend, host = [], None, []
    torch.cuda.synchronize()
    t00 = time.perf_counter()
    base = torch.cuda.Event(enable_timing=True); base.record()
    for i in range(n):
        a = torch.cuda.Event(enable_timing=True); a.record()
        t0 = time.perf_counter()
        cur = tg.newton_Algorithm(x0h, xr, ur, block=False, **kw, **extra)
        t1 = time.perf_counter()
        b = torch.cuda.Event(enable_timing=True); b.record()
        c = torch.cuda.Event(enable_timing=True); c.record(tg._pipeline.copy_stream)
        if pend is not None:
            pend.result()
        t2 = time.perf_counter()
        pend = cur
        evs.append((a, b, c)); host.append((t0 - t00, t1 - t0, t2 - t1))
    pend.result()
    torch.cuda.synchronize()
    total = time.perf_counter() - t00
    for i, ((a, b, c), h) in enumerate(zip(evs, host)):
        print("  step %d: compute-stream start %.2f end %.2f | copy-stream end %.2f | host: submit at %.2f, submit took %.2f, result(prev) took %.2f ms"
              % (i, base.elapsed_time(a), base.elapsed_time(b), base.elapsed_time(c), 1e3 * h[0], 1e3 * h[1], 1e3 * h[2]))
    print("  total %.2f ms for %d steps = %.2f ms/step" % (1e3 * total, n, 1e3 * total / n))

for name, extra in (("pipelined, full returns", {}), ("pipelined, return_gains=False", dict(return_gains=False))):
    print(name)
    timeline(3, **extra)   # warm-up (pinned allocations)
    timeline(6, **extra)

# blocking call split
for _ in range(2):
    tg.newton_Algorithm(x0h, xr, ur, **kw)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(4):
    tg.newton_Algorithm(x0h, xr, ur, **kw)
torch.cuda.synchronize(); print("blocking: %.2f ms/step" % (1e3 * (time.perf_counter() - t0) / 4))

# raw D2H bandwidth of one 131 MB array into a fresh pinned tensor, alone and while a solve runs
K = torch.empty(4096, 500, 2, 4, dtype=torch.float64, device='cuda')
for conc in (False, True):
    h = torch.empty(K.shape, dtype=K.dtype, pin_memory=True)
    torch.cuda.synchronize()
    if conc:
        st = bt.newton_solve(bt.upload(np.ascontiguousarray(x0.T)), bt.make_ref(d['x'], u_ref), max_iters=50, tol=0.0, gamma_0=0.1)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); h.copy_(K, non_blocking=True); e1.record()
    torch.cuda.synchronize()
    print("D2H 131 MB %s: %.2f ms = %.1f GB/s" % ("during a solve" if conc else "alone", e0.elapsed_time(e1), K.numel() * 8 / e0.elapsed_time(e1) / 1e6))

# Armijo tries at gamma_0 = 1
st = bt.newton_solve(bt.upload(np.ascontiguousarray(x0.T)), bt.make_ref(d['x'], u_ref), max_iters=50, tol=0.0, gamma_0=1.0)
torch.cuda.synchronize()
nt = st.hist_ntry.cpu().numpy()
os.makedirs('gpurun_out', exist_ok=True)
np.savez_compressed('gpurun_out/ntry_gamma1.npz', ntry=nt, status=st.status.cpu().numpy(), iters=st.iters.cpu().numpy(), cost=st.hist_cost.cpu().numpy())
tile_max = nt.reshape(nt.shape[0], -1, 32).max(axis=2)
print("gamma_0=1: mean tries %.2f, mean tile-max %.2f, per-iteration tile-max mean:" % (nt.mean(), tile_max.mean()), np.round(tile_max.mean(axis=1), 1))
