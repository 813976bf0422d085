"""Per-call host timings of the pipelined solve_mpc_tracking loop of bench.py (config 4), to see where a step waits."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gymnast_optimalcontrol_b200 import trajectory_tracking as tt  # noqa: E402
from gymnast_optimalcontrol_b200 import _io as _io0  # noqa: E402

# experiment switches (this probe only): PROBE_PRIO = priority of the side streams, CUDA_DEVICE_MAX_CONNECTIONS from the shell
if "PROBE_PRIO" in os.environ:
    pr = int(os.environ["PROBE_PRIO"])
    _io0._side[torch.cuda.current_device()] = (torch.cuda.Stream(priority=pr), torch.cuda.Stream(priority=pr))
print("config: prio", os.environ.get("PROBE_PRIO", "default(-1)"), "max connections", os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS", "default"))

d = np.load(os.path.join(ROOT, "tests", "golden", "acrobot_optimal_trajectory.npz"))
B, H, N = 16384, 75, 501
x0h = d["x"][0] + np.random.default_rng(3).uniform(-0.1, 0.1, (B, 4))
x0 = torch.from_numpy(x0h).pin_memory()
xr = torch.from_numpy(np.repeat(d["x"][None], B, 0)).pin_memory()
ur = torch.from_numpy(np.repeat(d["u"][None], B, 0)).pin_memory()
for piped in (True, False, True):
    pend, res, rows = None, None, []
    torch.cuda.synchronize()
    t_all = time.perf_counter()
    for i in range(14):
        t0 = time.perf_counter()
        nxt = tt.solve_mpc_tracking(x0, xr, ur, N, T_pred=H, block=not piped)
        t1 = time.perf_counter()
        if piped:
            if pend is not None:
                res = pend.result()
            pend = nxt
        else:
            res = nxt
        t2 = time.perf_counter()
        rows.append("%d: issue %.1f ms, result %.1f ms, reserved %.2f GB" % (i, 1e3 * (t1 - t0), 1e3 * (t2 - t1),
                                                                          torch.cuda.memory_reserved() / 2**30))
    if pend is not None:
        res = pend.result()
    torch.cuda.synchronize()
    print("piped" if piped else "blocking", "total %.1f ms for 14 calls" % (1e3 * (time.perf_counter() - t_all)))
    print("\n".join(rows), flush=True)

# ---- stage timeline of the pipelined loop: events on the upload / main / copy streams ----
from gymnast_optimalcontrol_b200 import _io, batched as bt  # noqa: E402

w = bt.mpc_weights()
P = tt._terminal_weight(tt.active_params(), bt.Weights(tt.Q_mpc, tt.R_mpc))
base = torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
base.record()
marks, pend_q = [], []


def ev(stream):
    e = torch.cuda.Event(enable_timing=True)
    e.record(stream)
    return e


def issue(i):
    main = torch.cuda.current_stream()
    up, cs = _io.side_streams()
    m = {"i": i, "host_issue": time.perf_counter()}
    m["u0"] = ev(up)
    with _io.upload_scope(True, x0, xr, ur) as scope:
        x0d, kind = _io.state_in(x0, 4)
        ref = bt.make_ref(xr, ur)
        scope.keep(x0d, ref)
        m["u1"] = ev(up)
    _io.flush_deferred()
    m["c0"] = ev(main)
    Xr, Ur, K0, ns = bt.mpc_track(x0d, ref, P, T=N, T_pred=H, w=w)
    m["c1"] = ev(main)
    pend = _io.out_async([(Xr, N), (Ur, N - 1)], kind, defer=True)
    marks.append(m)
    return pend


pend = None
t_host0 = time.perf_counter()
for i in range(10):
    nxt = issue(i)
    if pend is not None:
        pend.result()
        marks[i - 1]["host_result"] = time.perf_counter()
    pend = nxt
pend.result()
torch.cuda.synchronize()
print("stage timeline (ms since start; u = upload+pack, c = MPC kernel, d = unpack+copy back)")
for m in marks:
    print("%d: host issue %.1f | u %.1f-%.1f | c %.1f-%.1f | host result %.1f" % (
        m["i"], 1e3 * (m["host_issue"] - t_host0), base.elapsed_time(m["u0"]), base.elapsed_time(m["u1"]),
        base.elapsed_time(m["c0"]), base.elapsed_time(m["c1"]),
        1e3 * (m.get("host_result", 0) - t_host0)))
