#!/usr/bin/env python
"""Secondary benchmark lines: configs 3, 4, 5 of BASELINE.json and config 2 in the throughput regime.

    python bench_configs.py [--quick]

One JSON line per config, device-resident timing with CUDA events (inputs in HBM), FP64 roofline by the flop
convention of SURVEY 8(d).  bench.py (config 2, B = 4096) stays the headline the driver runs.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from gymnast_optimalcontrol_b200 import _abi  # noqa: E402
from gymnast_optimalcontrol_b200 import batched as bt  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def timeit(f, n):
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _abi.launch_count()
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / n, (_abi.launch_count() - l0) // n


def fp64_peak():
    blocks, threads, iters = 148 * 8, 256, 20000
    out = torch.empty(blocks * threads, dtype=torch.float64, device="cuda")
    best = 0.0
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _abi.call("acro_bench_fp64_peak", blocks, threads, iters, C.c_void_p(out.data_ptr()), None)
        e1.record()
        torch.cuda.synchronize()
        if i:
            best = max(best, blocks * threads * iters * 16.0 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def line(name, metric, value, unit, flops_per_unit, t, launches, peak, extra=None):
    ach = value * flops_per_unit / 1e12
    d = {"config": name, "metric": metric, "value": value, "unit": unit, "seconds": t, "gpu_launches": launches,
         "roofline": {"bound": "fp64", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                      "flops_per_unit": flops_per_unit,
                      "convention": "SURVEY 8(d): dense 4x4 algebra without credit for structure, 40 flops per sin or cos; "
                                    "the kernels execute fewer FP64 instructions than that, so frac can exceed 1 (pipe "
                                    "utilisation measured by ncu: DESIGN.md section 4.3)"}}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    reps = 2 if a.quick else 5
    peak = fp64_peak()
    opt = np.load(os.path.join(G, "acrobot_optimal_trajectory.npz"))
    fa = np.load(os.path.join(G, "fully_actuated_trajectory.npz"))
    u_ref = np.zeros(fa["u"].shape)
    u_ref[:, 1] = 2.0 * fa["u"][:, 1]
    N = 501

    # ---- config 3: LQR tracking of the optimal trajectory from 65 536 perturbed initial conditions
    B = 65536
    traj = bt.make_ref(opt["x"], opt["u"])
    K = bt.lqr_gains(traj)
    x0 = bt.upload(np.ascontiguousarray((opt["x"][0] + np.random.default_rng(2).uniform(-0.3, 0.3, (B, 4))).T))
    t, nl = timeit(lambda: bt.lqr_track(traj, K, x0), reps)
    line("C3 LQR tracking, B=65536, N=501", "lqr_rollouts_per_sec", B / t, "tracked problems/s", 878.0 * (N - 1), t, nl, peak,
         {"steps_per_sec": B * (N - 1) / t})
    # the same with every plant its own physical parameters: nominal gains, mismatched plants
    rng = np.random.default_rng(7)
    rows = np.array([[getattr(bt.DEFAULT_PARAMS, f) * (1.0 if f == "g" else 1.0) for f in bt.PHYS_FIELDS]] * B)
    rows[:, :8] *= rng.uniform(0.97, 1.03, (B, 8))
    pb = bt.phys_params(rows)
    t, nl = timeit(lambda: bt.lqr_track(traj, K, x0, params_b=pb), reps)
    line("C3 LQR tracking with per-problem physical parameters (+-3 %), B=65536, N=501", "lqr_rollouts_per_sec", B / t,
         "tracked problems/s", 878.0 * (N - 1), t, nl, peak, {"steps_per_sec": B * (N - 1) / t})
    tg, _ = timeit(lambda: bt.lqr_gains(traj), reps)
    print(json.dumps({"config": "C3 gains (one shared trajectory, 500 sequential Riccati steps, 1 thread)", "seconds": tg}))

    # ---- config 4: receding-horizon MPC for 16 384 acrobots
    B = 16384
    w = bt.mpc_weights()
    xf = bt.upload(np.array([[np.pi], [0], [0], [0]], dtype=np.float64))
    A_f, B_f = bt.linearize(xf, bt.upload(np.zeros((2, 1))), discrete=True)
    P, n = bt.p_inf(A_f, B_f, w)
    QT = P[:, :, 0].contiguous()
    x0 = bt.upload(np.ascontiguousarray((opt["x"][0] + np.random.default_rng(3).uniform(-0.1, 0.1, (B, 4))).T))
    for H in ((75,) if a.quick else (50, 75, 100, 200)):
        res = {}
        t, nl = timeit(lambda: res.__setitem__("s", bt.mpc_track(x0, traj, QT, T=N, T_pred=H, w=w)), reps)
        ns = res["s"][3]
        line("C4 MPC tracking, shared reference, B=16384, H=%d" % H, "mpc_control_updates_per_sec", B * (N - 1) / t,
             "closed-loop MPC steps/s (gains from %d Riccati sweeps shared by the batch)" % ns, 856.0 + 16.0, t, nl, peak,
             {"riccati_sweeps_executed": ns})
    # the same with the input box of tt:87-91, 112-114 (tau_max = 18; the shipped trajectory asks for up to 23.9): every
    # QP solved exactly by an active-set method on Riccati sweeps; "solves" = receding-horizon steps x problems
    for H in ((75,) if a.quick else (50, 75)):
        res = {}
        t, nl = timeit(lambda: res.__setitem__("s", bt.mpc_track_box(x0, traj, QT, tau_max=18.0, T=N, T_pred=H, w=w)), 1)
        info = res["s"][2]
        sweeps = float(info["n_sweeps"].double().sum().item())
        # solves that end with nothing held ran (but for rare exceptions) one forward sweep against the shared gain
        # table (~60 flops per window step); every other active-set iteration is a backward sweep (~300), a forward
        # sweep (~60) and, when it is not blocked, a costate sweep (~90)
        free_solves = float((info["n_active"] == 0).double().sum().item())
        flops = (free_solves * 60.0 + (sweeps - free_solves) * 450.0) * (H - 1) + B * (N - 1) * 856.0
        print(json.dumps({"config": "C4 MPC tracking with the input box |u| <= 18, shared reference, B=16384, H=%d" % H,
                          "metric": "mpc_solves_per_sec", "value": B * (N - 1) / t,
                          "unit": "box-constrained MPC solves/s (exact active-set solve + plant step each)", "seconds": t,
                          "gpu_launches": nl, "active_set_iterations_per_solve": sweeps / (B * (N - 1)),
                          "steps_with_active_bounds": float((info["n_active"] > 0).double().mean().item()),
                          "iteration_limit_hit": int(info["status"].sum().item()),
                          "roofline": {"bound": "fp64", "achieved": flops / t / 1e12, "peak": peak,
                                       "unit": "TFLOP/s", "frac": flops / t / 1e12 / peak,
                                       "flops_per_unit": "per window step: 60 for a solve whose working set stays empty (forward "
                                                         "sweep against the shared gain table), 450 per active-set iteration "
                                                         "otherwise (backward, forward and costate sweeps); + 856 per plant step"}}),
              flush=True)
    refp = bt.Ref(bt.Traj.from_batch_major(bt.upload(np.repeat(opt["x"][None], B, 0))),
                  bt.Traj.from_batch_major(bt.upload(np.repeat(opt["u"][None], B, 0))))
    # the input box with per-problem references: no gain table, every solve runs its own backward sweep
    res = {}
    t, nl = timeit(lambda: res.__setitem__("s", bt.mpc_track_box(x0, refp, QT, tau_max=18.0, T=N, T_pred=75, w=w)), 1)
    info = res["s"][2]
    sweeps = float(info["n_sweeps"].double().sum().item())
    print(json.dumps({"config": "C4 MPC tracking with the input box |u| <= 18, per-problem references, B=16384, H=75",
                      "metric": "mpc_solves_per_sec", "value": B * (N - 1) / t,
                      "unit": "box-constrained MPC solves/s (exact active-set solve + plant step each)", "seconds": t,
                      "gpu_launches": nl, "active_set_iterations_per_solve": sweeps / (B * (N - 1)),
                      "roofline": {"bound": "fp64", "achieved": (sweeps * 74 * 450.0 + B * (N - 1) * 856.0) / t / 1e12, "peak": peak,
                                   "unit": "TFLOP/s", "frac": (sweeps * 74 * 450.0 + B * (N - 1) * 856.0) / t / 1e12 / peak,
                                   "flops_per_unit": "450 per window step and active-set iteration + 856 per plant step"}}),
          flush=True)
    for H in ((75,) if a.quick else (50, 75, 100, 200)):
        res = {}
        t, nl = timeit(lambda: res.__setitem__("s", bt.mpc_track(x0, refp, QT, T=N, T_pred=H, w=w)), max(1, reps // 2))
        ns = res["s"][3]
        line("C4 MPC tracking, per-problem references, B=16384, H=%d" % H, "mpc_solves_per_sec", ns / t,
             "MPC solves/s (one (H-1)-step Riccati sweep + plant step each)", 180.0 * (H - 1) + 16 + 856, t, nl, peak,
             {"riccati_sweeps_executed": ns,
              "survey_convention_frac": ns / t * (550.0 * (H - 1) + 16 + 856) / 1e12 / peak,
              "flops": "executed: a sweep step is 108 FP64 instructions = 180 flops (A_d has two trivial rows, inner steps "
                       "need no gain); SURVEY 8(d)'s dense count is 550 per step (survey_convention_frac)"})

    # per-problem physical parameters (+-3 %) in the MPC tracker: own linearisation, own terminal weight, own plant
    rows = np.array([[getattr(bt.DEFAULT_PARAMS, f) for f in bt.PHYS_FIELDS]] * B)
    rows[:, :8] *= np.random.default_rng(8).uniform(0.97, 1.03, (B, 8))
    pbm = bt.phys_params(rows)
    A_fb, B_fb = bt.linearize(xf.repeat(1, B), bt.upload(np.zeros((2, B))), discrete=True, params_b=pbm)
    Pb, _ = bt.p_inf(A_fb, B_fb, w)
    res = {}
    t, nl = timeit(lambda: res.__setitem__("s", bt.mpc_track(x0, refp, Pb, T=N, T_pred=75, w=w, params_b=pbm)), max(1, reps // 2))
    ns = res["s"][3]
    line("C4 MPC tracking with per-problem physical parameters (+-3 %), per-problem references, B=16384, H=75",
         "mpc_solves_per_sec", ns / t, "MPC solves/s (one (H-1)-step Riccati sweep + plant step each)",
         180.0 * 74 + 16 + 856, t, nl, peak, {"riccati_sweeps_executed": ns})

    # ---- config 5: 5000 base iterates x 200 step sizes = 1M closed-loop rollouts
    Pn = 5000 if not a.quick else 1000
    ref = bt.make_ref(fa["x"], u_ref)
    wn = bt.newton_weights()
    x0 = bt.upload(np.ascontiguousarray(np.random.default_rng(1).uniform(-0.2, 0.2, (Pn, 4)).T))
    st = bt.newton_solve(x0, ref, max_iters=3, tol=0.0, gamma_0=0.1, history=False)
    Kd, Sd, dJ, sn = bt.riccati_affine(st.X, st.U, ref, wn)
    steps = bt.upload(np.linspace(0, 1.25, 200))
    t, nl = timeit(lambda: bt.stepsize_sweep(st.X, st.U, Kd, Sd, ref, wn, steps), reps)
    line("C5 step-size sweep, %d iterates x 200 step sizes" % Pn, "sweep_rollouts_per_sec", Pn * 200 / t, "rollouts/s",
         940.0 * (N - 1), t, nl, peak)

    # ---- config 2 to convergence: B = 4096, tol 1e-4, gamma_0 = 0.1 as task_2 (main.py:55-71), and the
    # back-tracking regime gamma_0 = 1 over a fixed 30 iterations (Armijo tries differ from problem to problem)
    B = 4096
    x0h = np.random.default_rng(1).uniform(-0.2, 0.2, (B, 4))
    x0h[0] = 0.0
    x0 = bt.upload(np.ascontiguousarray(x0h.T))
    state = bt.newton_alloc(B, N, 5000, history=True)

    def run_conv():
        state.initialised = False
        bt.newton_solve(x0, ref, max_iters=5000, tol=1e-4, gamma_0=0.1, state=state)
    t, nl = timeit(run_conv, 1)
    its = state.iters.double()
    done = float(its.sum().item())
    print(json.dumps({"config": "C2 Newton to convergence, B=4096, tol=1e-4, gamma_0=0.1", "metric": "newton_iterations_per_sec",
                      "value": done / t, "unit": "Newton iterations/s (iterations actually executed)", "seconds": t,
                      "gpu_launches": nl, "solves_per_sec": B / t, "iterations_min_mean_max": [float(its.min().item()), done / B, float(its.max().item())],
                      "converged": int((state.status == 1).sum().item()),
                      "roofline": {"bound": "fp64", "achieved": done / t * 2054.0 * (N - 1) / 1e12, "peak": peak, "unit": "TFLOP/s",
                                   "frac": done / t * 2054.0 * (N - 1) / 1e12 / peak, "flops_per_unit": 2054.0 * (N - 1)}}), flush=True)
    state = bt.newton_alloc(B, N, 30, history=True)
    for kern in ("auto", "duo"):
        def run_bt():
            state.reset()
            bt.newton_solve(x0, ref, max_iters=30, tol=0.0, gamma_0=1.0, state=state, kernel=kern)
        t, nl = timeit(run_bt, 2)
        tries = float(state.hist_ntry[:30].double().mean().item())
        done = float(state.iters.double().sum().item())
        print(json.dumps({"config": "C2 Newton with back-tracking, B=4096, gamma_0=1, 30 iterations, kernel=%s" % kern,
                          "kernel": bt.newton_kernel_name(B, gamma_0=1.0, kernel=kern), "metric": "newton_iterations_per_sec",
                          "value": done / t, "unit": "Newton iterations/s (iterations actually executed)", "seconds": t, "gpu_launches": nl,
                          "armijo_tries_mean": tries, "tile_max_tries_mean": float(state.hist_ntry[:30].reshape(30, -1, 32).max(dim=2).values.double().mean().item()),
                          "roofline": {"bound": "fp64", "achieved": done / t * (1114.0 + 940.0 * tries) * (N - 1) / 1e12, "peak": peak,
                                       "unit": "TFLOP/s", "frac": done / t * (1114.0 + 940.0 * tries) * (N - 1) / 1e12 / peak,
                                       "flops_per_unit": (1114.0 + 940.0 * tries) * (N - 1),
                                       "note": "a tile runs a forward pass for every candidate ANY of its 32 problems still needs (tile-max); "
                                               "k_newton_spec evaluates up to 8 of them in parallel"}}), flush=True)

    # ---- SURVEY 8f rank 1 inside the fast kernels: every problem its own physical parameters / its own reference
    rng = np.random.default_rng(7)
    for Bp in ((4096,) if a.quick else (4096, 65536)):
        rows = np.array([[getattr(bt.DEFAULT_PARAMS, f) for f in bt.PHYS_FIELDS]] * Bp)
        rows[:, :8] *= rng.uniform(0.97, 1.03, (Bp, 8))
        pb = bt.phys_params(rows)
        x0p = bt.upload(np.ascontiguousarray(np.random.default_rng(1).uniform(-0.2, 0.2, (Bp, 4)).T))
        stp = bt.newton_alloc(Bp, N, 20, history=False)

        def run_pp():
            stp.reset()
            bt.newton_solve(x0p, ref, max_iters=20, tol=0.0, gamma_0=0.1, state=stp, params_b=pb)
        t, nl = timeit(run_pp, 2)
        line("C2 Newton with per-problem physical parameters (+-3 %%), B=%d, 20 iterations" % Bp, "newton_iterations_per_sec",
             float(stp.iters.double().sum().item()) / t, "Newton iterations/s", 2054.0 * (N - 1), t, nl, peak,
             {"kernel": bt.newton_kernel_name(Bp, params_per_problem=True)})
        del stp
    refp = bt.Ref(bt.Traj.from_batch_major(bt.upload(np.repeat(fa["x"][None], 4096, 0))),
                  bt.Traj.from_batch_major(bt.upload(np.repeat(u_ref[None], 4096, 0))))
    stp = bt.newton_alloc(4096, N, 20, history=False)

    def run_rp():
        stp.reset()
        bt.newton_solve(x0, refp, max_iters=20, tol=0.0, gamma_0=0.1, state=stp)
    t, nl = timeit(run_rp, 2)
    line("C2 Newton with per-problem references, B=4096, 20 iterations", "newton_iterations_per_sec",
         float(stp.iters.double().sum().item()) / t, "Newton iterations/s", 2054.0 * (N - 1), t, nl, peak,
         {"kernel": bt.newton_kernel_name(4096, ref_per_problem=True)})
    del stp, refp
    # the fully-actuated plant (tau_1 live) on the one-thread-per-problem kernel
    pa = bt.make_params(1, actuated_tau1=True)
    refa = bt.make_ref(fa["x"], fa["u"])
    wa = bt.Weights(np.diag([130.0, 30.0, 1e-4, 1e-4]), np.diag([0.5, 1.5]), np.diag([130.0, 130.0, 1.0, 1.0]))
    for Ba in ((4096,) if a.quick else (4096, 65536)):
        x0a = bt.upload(np.ascontiguousarray(np.random.default_rng(1).uniform(-0.2, 0.2, (Ba, 4)).T))
        sta = bt.newton_alloc(Ba, N, 10, history=False)

        def run_act():
            sta.reset()
            bt.newton_solve(x0a, refa, max_iters=10, tol=0.0, gamma_0=0.5, state=sta, params=pa, w=wa)
        t, nl = timeit(run_act, 2)
        line("C2-like Newton on the fully-actuated plant, B=%d, 10 iterations" % Ba, "newton_iterations_per_sec",
             float(sta.iters.double().sum().item()) / t, "Newton iterations/s", 2054.0 * (N - 1), t, nl, peak,
             {"kernel": "acro::k_newton<false,false,false,true>"})
        del sta

    # ---- config 2 in the throughput regime
    hbm = 6551.4
    try:
        hbm = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for B in ((65536,) if a.quick else (8192, 16384, 32768, 65536, 131072)):
        x0 = bt.upload(np.ascontiguousarray(np.random.default_rng(1).uniform(-0.2, 0.2, (B, 4)).T))
        state = bt.newton_alloc(B, N, 10, history=False)

        def run():
            state.initialised = False
            bt.newton_solve(x0, ref, max_iters=10, tol=0.0, gamma_0=0.1, state=state)
        t, nl = timeit(run, max(1, reps // 2))
        rate = B * 10 / t
        line("C2 Newton, B=%d, 10 iterations" % B, "newton_iterations_per_sec", rate, "Newton iterations/s",
             2054.0 * (N - 1), t, nl, peak,
             {"hbm": {"algorithmic_gbs": rate * (N - 1) * 304.0 / 1e9, "implementation_gbs": rate * (N - 1) * 464.0 / 1e9,
                      "peak_gbs": hbm, "frac_algorithmic": rate * (N - 1) * 304.0 / 1e9 / hbm,
                      "frac_implementation": rate * (N - 1) * 464.0 / 1e9 / hbm,
                      "note": "304 B per problem-step-iteration by the SURVEY 8(d) count; the kernels move 464 B (they store and "
                              "re-read the 80 B linearisation instead of recomputing it)"}})


if __name__ == "__main__":
    main()
