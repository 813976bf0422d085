#!/usr/bin/env python
"""bench.py - Newton iterations/s of the batched acrobot hot path (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over one batch: `--iters` regularised-Newton iterations (backward
affine Riccati sweep + Armijo closed-loop rollouts, trajectory_generation.py:329-396) for each of the
`--batch` independent swing-up problems of config 2 (B = 4096 randomised initial states, N = 501 time
steps as shipped, gamma_0 = 0.1, weights of trajectory_generation.py:16-18), starting from the open-loop
rollout (tol = 0: no early exit, every problem runs every iteration).

    value   device-resident throughput: inputs already in HBM, CUDA events around K steps; counted from the
            all-reduced sum of the per-problem iteration counters, not from the nominal B x iters
    e2e     the same through the drop-in trajectory_generation.newton_Algorithm with HOST (pinned) buffers:
            per step H2D of x0 / x_ref / u_ref and D2H of x_traj, u_traj, K, sigma and the history; the calls are
            issued with block=False, so the copies of step i overlap the kernel of step i+1 (double-buffered
            solver state); `e2e.sync` is the same with blocking calls
    roofline / cpu_baseline / clocks: see DESIGN.md section "Measurement"
    backtracking, mpc, strong, long_horizon: further blocks of the same JSON line

Multi-GPU (torchrun): every rank solves its own 4096 problems (weak scaling, no collective on the hot
path); one NCCL gather of the per-problem summary (cost, status, iterations) per step is inside the timed
region.  The `strong` block splits fixed total batches over the ranks instead.
`--impl reference` times the UNMODIFIED reference (oracle/_ref, staged by oracle/build_ref.py) on rank 0's host
cores, one problem per process; if it cannot be imported there, the NumPy port (oracle/acro_oracle.py), and says so.
"""
import argparse
import glob
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_STEPS = 501
FLOPS_FIXED, FLOPS_PER_TRY = 1114.0, 940.0  # per problem-step-iteration, FMA = 2, sin/cos = 40
BYTES_STEP_ITER = 304.0                      # read x,u; write K,sigma; read x,u,K,sigma; write x+,u+
TRAFFIC_STEP_INIT = 176.0                    # design bytes per problem-step of the initial rollout + cost (X, U, lin written, read back once)
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 37.2


def load_reference_trajectory():
    d = np.load(os.path.join(ROOT, "tests", "golden", "fully_actuated_trajectory.npz"))
    u_ref = np.zeros(d["u"].shape)
    u_ref[:, 1] = 2.0 * d["u"][:, 1]  # trajectory_generation.py:511-518
    return np.ascontiguousarray(d["x"]), u_ref


def make_x0(batch, rank):
    x0 = np.random.default_rng(1 + rank).uniform(-0.2, 0.2, (batch, 4))
    if rank == 0:
        x0[0] = 0.0  # problem 0 = task_2 of main.py, whose result the reference ships
    return x0


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def headline_ncu():
    """Newest committed ncu capture of the headline kernel (profiles/r*_newton_headline_ncu.json, written by
    profiles/ncu_to_json.py from an `ncu --set full` report): dram bytes and FP64 pipe % of that launch."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_newton_headline_ncu.json")))
    if not files:
        return None
    try:
        d = json.load(open(files[-1]))
        k = d["kernels"][0]
        k["file"] = os.path.relpath(files[-1], ROOT)
        k["launch"] = d.get("launch", {})
        return k
    except Exception:
        return None


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def wait_first_row(self, timeout=8.0):
        """nvidia-smi needs a second or more to start on an 8-GPU box: block until it streams."""
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        return len(self.rows)

    def summary(self, lo=0, hi=None):
        """Rows lo:hi (the timed region); if the region was too short to be sampled, every row so far (the warm-up
        steps before it run the same kernels on the same inputs)."""
        rows = self.rows[lo:hi]
        window = "timed region"
        if not rows:
            rows, window = list(self.rows), "warm-up + timed region (timed region shorter than the sampling interval)"
        out = self._summary(rows)
        out["window"] = window
        return out

    def _summary(self, rows):
        sm, mx, reasons = [], 0.0, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------- CPU arms
def _quiet_stdout():
    """The reference prints progress lines; stdout of this program carries exactly one JSON line."""
    sys.stdout.flush()
    fd = os.dup(1)
    os.dup2(2, 1)
    return fd


def _cpu_port_worker(args):
    x0, iters = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import acro_oracle as O
    x_ref, u_ref = load_reference_trajectory()
    t = time.perf_counter()
    x, u, K, s, h = O.newton_Algorithm(x0, x_ref, u_ref, max_iters=iters, tol=0.0, gamma_0=0.1)
    return time.perf_counter() - t, h["iters"], float(h["cost"][-1])


_REF = {}


def _cpu_ref_worker(args):
    """newton_Algorithm of the UNMODIFIED reference (trajectory_generation.py:298) on one problem."""
    x0, iters = args
    rtg = _REF["rtg"]
    x_ref, u_ref = load_reference_trajectory()
    import contextlib
    import io
    t = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):  # the reference prints a progress line every 10 iterations (tg:391-392)
        x, u, K, s, h = rtg.newton_Algorithm(x0, x_ref, u_ref, max_iters=iters, tol=0.0, gamma_0=0.1, plot_armijo_iters=0)
    return time.perf_counter() - t, len(h["sigma_norm"]), float(h["cost"][-1])


def reference_available():
    try:
        from oracle import ref_import
        if not ref_import.available():
            return False
        if "rtg" not in _REF:
            _REF["rd"], _REF["rtg"], _REF["rtt"] = ref_import.load()  # SymPy model built once, inherited by fork()
        return True
    except Exception as e:  # noqa: BLE001
        sys.stderr.write("reference not importable (%r): CPU arm falls back to the NumPy port\n" % (e,))
        return False


def cpu_newton_rate(pool, cores, iters_per_problem, seed=0, real=False):
    """`cores` problems of the workload, one per worker, `iters_per_problem` Newton iterations each."""
    x0 = make_x0(4096, 0)[seed * cores:(seed + 1) * cores]
    t = time.perf_counter()
    res = pool.map(_cpu_ref_worker if real else _cpu_port_worker, [(x0[i], iters_per_problem) for i in range(cores)])
    wall = time.perf_counter() - t
    done = sum(r[1] for r in res)
    return done / wall, wall, done


def cpu_baseline_block(cpu_iters, ref_iters):
    """The reference's own CPU path on this box's cores (kind "reference" when oracle/_ref imports), and the NumPy port
    as a second, faster CPU number."""
    cores = os.cpu_count() or 1
    out = None
    real = reference_available()
    with mp.get_context("fork").Pool(cores) as pool:
        if real:
            r, wall, d = cpu_newton_rate(pool, cores, ref_iters, real=True)
            out = {"value": r, "unit": "Newton iterations/s", "cores": cores, "kind": "reference",
                   "sample": "%d problems of the 4096 (one per core) x %d Newton iterations of the unmodified reference "
                             "(oracle/_ref: trajectory_generation.newton_Algorithm), %.1f s of wall time" % (cores, ref_iters, wall)}
        r, wall, d = cpu_newton_rate(pool, cores, cpu_iters)
        port = {"value": r, "unit": "Newton iterations/s", "cores": cores, "kind": "port",
                "sample": "%d problems of the 4096 (one per core) x %d Newton iterations of oracle/acro_oracle.py, %.1f s of wall time"
                          % (cores, cpu_iters, wall)}
    if out is None:
        out = port
        out["note"] = "the reference could not be imported on this box (oracle/_ref or sympy missing): NumPy port"
    else:
        out["port"] = port
    return out


def _cpu_mpc_worker(args):
    """`n_steps` receding-horizon MPC solves (tt:43-67) of one problem with the oracle: (H-1)-step Riccati sweep on the
    sliding window + plant step."""
    x0, n_steps, H = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import acro_oracle as O
    d = np.load(os.path.join(ROOT, "tests", "golden", "acrobot_optimal_trajectory.npz"))
    x_ref, u_ref = d["x"], d["u"]
    N = x_ref.shape[0]
    Ad, Bd = O.linearize_discrete(x_ref[:-1], u_ref)
    A_f, B_f = O.linearize_discrete(O.X_F, O.U_F)
    Q_T = O.compute_P_inf(A_f, B_f, O.Q_MPC, O.R_MPC)
    x = np.array(x0, dtype=float)
    t0 = time.perf_counter()
    for t in range(n_steps):
        Aw = [Ad[t + j] if t + j < N - 1 else A_f for j in range(H - 1)]
        Bw = [Bd[t + j] if t + j < N - 1 else B_f for j in range(H - 1)]
        K0 = O.mpc_gains(Aw, Bw, O.Q_MPC, O.R_MPC, Q_T, H)[0]
        u = u_ref[t] + K0 @ (x - x_ref[t])
        x = O.dynamics(x, u)
    return time.perf_counter() - t0, n_steps


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    json_fd = _quiet_stdout()
    cores = os.cpu_count() or 1
    real = reference_available() and not a.port
    ipp = a.ref_iters if real else a.cpu_iters
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(min(a.warmup, 1)):
            cpu_newton_rate(pool, cores, 1, real=real)
        t0 = time.perf_counter()
        done = 0
        for s in range(a.steps):
            r, wall, d = cpu_newton_rate(pool, cores, ipp, seed=s % 8, real=real)
            done += d
        total = time.perf_counter() - t0
    value = done / total
    what = ("the unmodified reference (oracle/_ref: trajectory_generation.newton_Algorithm, Python + SymPy lambdify + NumPy)"
            if real else "the NumPy port oracle/acro_oracle.py (the reference could not be imported on this box)")
    sample = "%d problems of the 4096 (one per core) x %d Newton iterations per step, %s" % (cores, ipp, what)
    line = {
        "impl": "reference", "metric": "newton_iterations_per_sec", "value": value,
        "unit": "Newton iterations/s (N=501 time steps each)", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, sample_note=sample),
        "cpu_baseline": {"value": value, "unit": "Newton iterations/s", "cores": cores, "kind": "reference" if real else "port",
                         "sample": sample,
                         "note": "one problem per process on every host core, like the reference's per-problem Python loop"},
        "e2e": {"value": value, "unit": "Newton iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    os.write(json_fd, (json.dumps(line) + "\n").encode())


def workload_config(a, sample_note=None):
    c = {"workload": "config 2: batched Newton+Armijo swing-up solves, B=%d problems per GPU from randomised initial states, "
                     "N=501 (T=10 s, dt=0.02 as shipped), %d Newton iterations per step, gamma_0=0.1, tol=0 (no early exit)"
                     % (a.batch, a.iters),
         "batch_per_gpu": a.batch, "horizon_steps": N_STEPS - 1, "newton_iters_per_step": a.iters,
         "l2": "working set 525 MB per GPU (X,U,Xw,Uw,lin,K,S) > 126 MB L2, no flush needed"}
    if sample_note:
        c["cpu_sample"] = sample_note
    return c


# ----------------------------------------------------------------------------------------- GPU arm
def fp64_peak_tflops(bt, torch):
    """DFMA-chain microbenchmark (acro_bench_fp64_peak): achievable FP64 pipe rate of this GPU."""
    import ctypes as C
    blocks, threads, iters = 148 * 8, 256, 20000
    out = torch.empty(blocks * threads, dtype=torch.float64, device="cuda")
    best = 0.0
    for i in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bt.call("acro_bench_fp64_peak", blocks, threads, iters, C.c_void_p(out.data_ptr()),
                C.c_void_p(torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        if i:
            best = max(best, blocks * threads * iters * 16.0 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


class Ctx:
    """What every block needs: torch, the package modules, the process group."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        from gymnast_optimalcontrol_b200 import sharding as _sh
        self.cpus = _sh.bind_to_gpu_numa(self.local) if self.world > 1 else None  # before any pinned allocation
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        from gymnast_optimalcontrol_b200 import _abi
        from gymnast_optimalcontrol_b200 import batched as bt
        from gymnast_optimalcontrol_b200 import sharding
        from gymnast_optimalcontrol_b200 import trajectory_generation as tg
        from gymnast_optimalcontrol_b200 import trajectory_tracking as tt
        self._abi, self.bt, self.sharding, self.tg, self.tt = _abi, bt, sharding, tg, tt

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t]

    def time_device(self, fn, steps, warmup):
        """CUDA events around `steps` calls on the launching stream, barrier + synchronise on both sides, max over ranks."""
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1) * 1e-3)[0]


def newton_block(cx, B, iters, gamma_0, steps, warmup, x0=None, N=N_STEPS, ref_xy=None, gather=True, kernel=None):
    """Device-resident Newton throughput for one batch per rank.  Returns (iterations/s over all ranks, seconds per step,
    iterations done per step over all ranks, mean Armijo tries, state)."""
    bt, torch = cx.bt, cx.torch
    x_ref, u_ref = ref_xy if ref_xy is not None else load_reference_trajectory()
    ref = bt.make_ref(x_ref, u_ref)
    w = bt.newton_weights()
    x0 = make_x0(B, cx.rank) if x0 is None else x0
    x0d = bt.upload(np.ascontiguousarray(x0.T))
    state = bt.newton_alloc(B, N, iters, history=True)

    def step():
        state.initialised = False
        bt.newton_solve(x0d, ref, max_iters=iters, tol=0.0, gamma_0=gamma_0, w=w, state=state, kernel=kernel)
        if gather and cx.world > 1:  # the only collective: the per-problem summary of every shard, over NCCL/NVLink
            summ = cx.sharding.pack_summary(state.cost, state.status, state.iters, state.gamma_acc, state.sigma_norm)
            cx.sharding.gather_summary(summ, B * cx.world)

    t = cx.time_device(step, steps, warmup)
    done = cx.sum_over_ranks(float(state.iters.sum().item()))[0]
    ntry = float(state.hist_ntry[:iters].double().mean().item())
    return done * steps / t, t / steps, done, ntry, state


def headline(cx, a):
    torch, _abi = cx.torch, cx._abi
    with ClockSampler(cx.local) as clk:
        clk.wait_first_row()
        # warm-up outside the marks, then the timed region between two marks of the sampler
        bt = cx.bt
        x_ref, u_ref = load_reference_trajectory()
        ref = bt.make_ref(x_ref, u_ref)
        w = bt.newton_weights()
        x0 = make_x0(a.batch, cx.rank)
        x0d = bt.upload(np.ascontiguousarray(x0.T))
        state = bt.newton_alloc(a.batch, N_STEPS, a.iters, history=True)

        def step():
            state.initialised = False
            bt.newton_solve(x0d, ref, max_iters=a.iters, tol=0.0, gamma_0=0.1, w=w, state=state)
            if cx.world > 1:
                summ = cx.sharding.pack_summary(state.cost, state.status, state.iters, state.gamma_acc, state.sigma_norm)
                cx.sharding.gather_summary(summ, a.batch * cx.world)

        for _ in range(a.warmup):
            step()
        cx.barrier()
        l0 = _abi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = clk.mark()
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        cx.barrier()
        c1 = clk.mark()
    clocks = clk.summary(c0, c1)
    launches = _abi.launch_count() - l0
    t_dev = cx.max_over_ranks(e0.elapsed_time(e1) * 1e-3)[0]
    done_rank = float(state.iters.sum().item())
    done = cx.sum_over_ranks(done_rank)[0]
    return dict(t=t_dev, done_per_step=done, done_rank0=done_rank, launches=launches, clocks=clocks,
                ntry=float(state.hist_ntry[:a.iters].double().mean().item()), cost_mean=float(state.cost.mean().item()),
                x0=x0, x_ref=x_ref, u_ref=u_ref)


def e2e_block(cx, a, h):
    """Through trajectory_generation.newton_Algorithm with pinned host buffers, inputs host->device and results
    device->host inside the timed region of every step.  Three ways of calling it:
      solution  block=False, return_gains=False: x_traj, u_traj and the history come back (what main.py's task_1 / task_2
                consume, main.py:42-48, 65-79); K and sigma of the last iteration stay on the device as handles
      full      block=False: everything the reference's function returns, K and sigma included (twice the bytes)
      sync      blocking calls, everything returned: solve, then copy"""
    torch, tg = cx.torch, cx.tg
    x0_host = torch.from_numpy(h["x0"]).pin_memory()
    xr_host = torch.from_numpy(h["x_ref"]).pin_memory()
    ur_host = torch.from_numpy(h["u_ref"]).pin_memory()
    kw = dict(max_iters=a.iters, tol=0.0, gamma_0=0.1, verbose=False)
    res = {}

    def run(block, gains, steps):
        prev = None
        for _ in range(steps):
            cur = tg.newton_Algorithm(x0_host, xr_host, ur_host, block=block, return_gains=gains, **kw)
            if not block:
                if prev is not None:
                    res["out"] = prev.result()
                prev = cur
            else:
                res["out"] = cur
        if prev is not None:
            res["out"] = prev.result()

    out = {}
    h2d = x0_host.numel() * 8 + xr_host.numel() * 8 + ur_host.numel() * 8
    for mode, block, gains in (("solution", False, False), ("full", False, True), ("sync", True, True)):
        run(block, gains, 3)
        cx.barrier()
        t0 = time.perf_counter()
        run(block, gains, a.steps)
        cx.barrier()
        t = cx.max_over_ranks(time.perf_counter() - t0)[0]
        o = res["out"]
        d2h = sum(v.numel() * v.element_size() for v in o[:4] if isinstance(v, torch.Tensor)) + \
            sum(np.asarray(v).nbytes for k, v in o[4].items() if k in ("cost", "sigma_norm", "iters", "status", "n_try", "gamma"))
        done = cx.sum_over_ranks(float(np.asarray(o[4]["iters"]).sum()))[0]
        out[mode] = dict(t=t, d2h=int(d2h), done_per_step=done, cost_mean=float(o[4]["cost"][:, -1].mean()))
    out["h2d"] = int(h2d)
    return out


def measure_mpc(cx, steps, with_cpu):
    """Second half of the BASELINE.json metric: MPC solves/s on config 4 (B = 16 384 acrobots per GPU, horizon 75,
    500 receding-horizon steps, every problem its own reference copy so that every problem runs its own Riccati
    sweeps: 8.19 M solves per run and GPU)."""
    bt, torch, tt = cx.bt, cx.torch, cx.tt
    d = np.load(os.path.join(ROOT, "tests", "golden", "acrobot_optimal_trajectory.npz"))
    B, H, N = 16384, 75, 501
    w = bt.mpc_weights()
    xf = bt.upload(np.array([[np.pi], [0], [0], [0]], dtype=np.float64))
    A_f, B_f = bt.linearize(xf, bt.upload(np.zeros((2, 1))), discrete=True)
    P, _ = bt.p_inf(A_f, B_f, w)
    QT = P[:, :, 0].contiguous()
    x0h = d["x"][0] + np.random.default_rng(3 + cx.rank).uniform(-0.1, 0.1, (B, 4))
    x0 = bt.upload(np.ascontiguousarray(x0h.T))
    refp = bt.Ref(bt.Traj.from_batch_major(bt.upload(np.repeat(d["x"][None], B, 0))),
                  bt.Traj.from_batch_major(bt.upload(np.repeat(d["u"][None], B, 0))))
    res = {}

    def run():
        res["s"] = bt.mpc_track(x0, refp, QT, T=N, T_pred=H, w=w)

    t = cx.time_device(run, steps, 2)
    ns = res["s"][3]
    value = ns * steps * cx.world / t
    flops = 550.0 * (H - 1) + 16 + 856
    out = {"metric": "mpc_solves_per_sec", "value": value,
           "unit": "MPC solves/s (one (H-1)-step Riccati sweep on the sliding window + plant step each)",
           "config": {"workload": "config 4: receding-horizon MPC tracking of acrobot_optimal_trajectory.npz, B=16384 acrobots per GPU, "
                                  "horizon 75, 500 steps, per-problem references (every solve executed)",
                      "solves_per_run_per_gpu": ns},
           "ms_per_run": 1e3 * t / steps, "flops_per_solve": flops,
           "achieved_tflops_per_gpu": value / cx.world * flops / 1e12}
    # end to end through the drop-in: host x0 (B,4), per-problem references (B,N,4)/(B,N-1,2) in, x_real/u_real out
    x0_host = torch.from_numpy(x0h).pin_memory()
    xr_host = torch.from_numpy(np.repeat(d["x"][None], B, 0)).pin_memory()
    ur_host = torch.from_numpy(np.repeat(d["u"][None], B, 0)).pin_memory()

    def timed_e2e(piped):
        def loop(n):
            inflight = []  # up to three calls in flight: upload of call i+2, kernel of call i+1, copy-back of call i
            for _ in range(n):
                r = tt.solve_mpc_tracking(x0_host, xr_host, ur_host, N, T_pred=H, block=not piped)
                if not piped:
                    res["e"] = r
                    continue
                inflight.append(r)
                if len(inflight) >= 3:
                    res["e"] = inflight.pop(0).result()
            while inflight:
                res["e"] = inflight.pop(0).result()

        loop(8 if piped else 2)  # the first pipelined calls page-lock their result buffers and grow the device pools
        torch.cuda.synchronize()
        cx.barrier()
        t0 = time.perf_counter()
        loop(steps)
        torch.cuda.synchronize()
        cx.barrier()
        return cx.max_over_ranks(time.perf_counter() - t0)[0]

    te = timed_e2e(True)
    ts = timed_e2e(False)
    out["e2e"] = {"value": ns * steps * cx.world / te, "unit": "MPC solves/s", "ms_per_run": 1e3 * te / steps,
                  "h2d_bytes_per_step": int((x0_host.numel() + xr_host.numel() + ur_host.numel()) * 8),
                  "d2h_bytes_per_step": int(sum(t_.numel() * 8 for t_ in res["e"])),
                  "api": "trajectory_tracking.solve_mpc_tracking(x0[B,4], x_ref[B,N,4], u_ref[B,N-1,2] pinned host, T=501, "
                         "block=False), three calls in flight: upload of call i+2, kernel of call i+1, copy-back of call i",
                  "blocking": {"value": ns * steps * cx.world / ts, "ms_per_run": 1e3 * ts / steps,
                               "api": "the same call with block=True (the reference's calling convention)"}}
    if with_cpu:
        cores = os.cpu_count() or 1
        n_steps = 400
        with mp.get_context("fork").Pool(cores) as pool:
            t0 = time.perf_counter()
            r = pool.map(_cpu_mpc_worker, [(x0h[i], n_steps, H) for i in range(cores)])
            wall = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": sum(x[1] for x in r) / wall, "unit": "MPC solves/s", "cores": cores, "kind": "port",
                               "sample": "%d problems (one per core) x %d MPC steps, horizon 75, %.1f s of wall time; the reference "
                                         "solves each QP with CasADi/IPOPT (absent here): the port is the Riccati restatement" % (cores, n_steps, wall)}
    return out


def strong_block(cx, a):
    """Fixed TOTAL work split over the ranks (the batch shards across the GPUs, no hot-path collective): config 2 at its
    own size (4096 problems in total: 16 tiles per GPU at 8 GPUs - a latency-bound 1000-step recurrence per tile does not
    get faster by using fewer tiles per GPU), config 2 at a saturating total batch, and config 5 (5000 base iterates x
    200 step sizes = 10^6 rollouts)."""
    bt, torch, sh = cx.bt, cx.torch, cx.sharding
    out = {"scaling": "strong", "n_gpus": cx.world}
    iters = min(a.iters, 10)
    for total in (4096, 262144):
        lo, hi = sh.shard_bounds(total, cx.world, cx.rank)
        x0 = np.random.default_rng(1).uniform(-0.2, 0.2, (total, 4))[lo:hi]
        rate, t_step, done, ntry, st = newton_block(cx, hi - lo, iters, 0.1, 2, 1, x0=x0, gather=False)
        out["newton_total_%d" % total] = {"value": rate, "unit": "Newton iterations/s", "ms_per_step": 1e3 * t_step,
                                          "problems_per_gpu": hi - lo, "newton_iters_per_step": iters,
                                          "kernel": bt.newton_kernel_name(hi - lo)}
        del st
        torch.cuda.empty_cache()
    # config 5
    P, S = 5000, 200
    lo, hi = sh.shard_bounds(P, cx.world, cx.rank)
    x_ref, u_ref = load_reference_trajectory()
    ref, w = bt.make_ref(x_ref, u_ref), bt.newton_weights()
    x0 = np.random.default_rng(1).uniform(-0.2, 0.2, (P, 4))[lo:hi]
    st = bt.newton_solve(bt.upload(np.ascontiguousarray(x0.T)), ref, max_iters=3, tol=0.0, gamma_0=0.1, history=False)
    K, Sg, dJ, sn = bt.riccati_affine(st.X, st.U, ref, w)
    grid = bt.upload(np.linspace(0.0, 1.25, S))  # tg:257-258
    keep = {}

    def run():
        cost = bt.stepsize_sweep(st.X, st.U, K, Sg, ref, w, grid)
        best = grid[torch.argmin(torch.nan_to_num(cost, nan=float("inf")), dim=0)]
        keep["best"] = sh.gather_summary(best[None], P) if cx.world > 1 else best[None]

    t = cx.time_device(run, 3, 2)
    out["config5_sweep"] = {"value": P * S * 3 / t, "unit": "closed-loop rollouts/s (N=501)", "ms_per_step": 1e3 * t / 3,
                            "rollouts": P * S, "median_minimising_step": float(keep["best"].median().item())}
    return out


def long_horizon_block(cx, a):
    """A REAL 10 000-step horizon (N = 10 001, T = 200 s at dt = 0.02): the shipped swing-up reference followed by holding
    the upright equilibrium.  The target of BASELINE.json is quoted in 10k-step-horizon-equivalent Newton iterations/s."""
    N = 10001
    x_fa, u_fa = load_reference_trajectory()
    x_ref = np.vstack([x_fa, np.repeat(np.array([[np.pi, 0.0, 0.0, 0.0]]), N - x_fa.shape[0], 0)])
    u_ref = np.vstack([u_fa, np.zeros((N - 1 - u_fa.shape[0], 2))])
    out = {"horizon_steps": N - 1, "reference": "fully-actuated swing-up (500 steps) then the upright equilibrium, dt = 0.02",
           "gamma_0": 0.1, "unit": "Newton iterations/s at N = 10 001 (not rescaled)"}
    for B in (4096, 16384, 32768):
        rate, t_step, done, ntry, st = newton_block(cx, B, 10, 0.1, 2, 1, N=N, ref_xy=(x_ref, u_ref), gather=False)
        out["batch_%d" % B] = {"value": rate, "ms_per_step": 1e3 * t_step, "newton_iters_per_step": 10, "armijo_tries_mean": ntry,
                               "kernel": cx.bt.newton_kernel_name(B), "batch_per_gpu": B}
        del st
        cx.torch.cuda.empty_cache()
    return out


def run_native(a):
    json_fd = _quiet_stdout()  # stdout carries the one JSON line and nothing else (NCCL banners etc. go to stderr)
    cx = Ctx()
    torch, bt, _abi = cx.torch, cx.bt, cx._abi
    world, rank = cx.world, cx.rank
    B, iters = a.batch, a.iters

    h = headline(cx, a)
    e = e2e_block(cx, a, h)
    value = h["done_per_step"] * a.steps / h["t"]
    e2e_rate = {m: e[m]["done_per_step"] * a.steps / e[m]["t"] for m in ("solution", "full", "sync")}

    back = None
    if not a.quick:
        rate, t_step, done, ntry, st = newton_block(cx, B, iters, 1.0, max(2, a.steps // 2), 1)
        tile_max = float(st.hist_ntry[:iters].reshape(iters, -1, 32).max(dim=2).values.double().mean().item()) if B % 32 == 0 else None
        back = {"metric": "newton_iterations_per_sec", "value": rate, "ms_per_step": 1e3 * t_step, "gamma_0": 1.0,
                "armijo_tries_mean": ntry, "armijo_tries_tile_max_mean": tile_max, "iterations_done_per_step": done,
                "iterations_nominal_per_step": B * iters * world,
                "kernel": bt.newton_kernel_name(B, gamma_0=1.0),
                "note": "a tile of 32 problems needs a forward pass for every candidate ANY of its problems still has to try (tile-max, "
                        "not the mean); k_newton_spec evaluates up to 8 candidates per round in parallel and takes the decisions of the "
                        "sequential loop; problems whose line search fails (status 3, tg:367-369) stop, so fewer than the nominal "
                        "iterations are done and counted",
                "config": "config 2 with the reference's default gamma_0 = 1 (trajectory_generation.py:298): the back-tracking regime"}
        frac_flops = (FLOPS_FIXED + FLOPS_PER_TRY * ntry) * (N_STEPS - 1) * rate / world / 1e12
        back["achieved_tflops_per_gpu"] = frac_flops
        del st
        torch.cuda.empty_cache()
    mpc = None if a.no_mpc else measure_mpc(cx, max(2, min(a.steps, 5)), with_cpu=(world == 1 and rank == 0 and not a.no_cpu))
    strong = None if a.quick else strong_block(cx, a)
    longh = None if a.quick else long_horizon_block(cx, a)

    if rank == 0:
        peaks = measured_peaks()
        fp64_meas = fp64_peak_tflops(_abi, torch)
        ncu = headline_ncu()
        if mpc:
            ach = mpc.pop("achieved_tflops_per_gpu")
            exe = mpc["value"] / world * (180.0 * 74 + 16 + 856) / 1e12
            mpc["roofline"] = {"bound": "fp64", "achieved": exe, "peak": fp64_meas, "unit": "TFLOP/s", "frac": exe / fp64_meas,
                               "flops_per_solve": 180.0 * 74 + 16 + 856,
                               "survey_convention": {"achieved": ach, "frac": ach / fp64_meas,
                                                     "flops_per_solve": mpc["flops_per_solve"]},
                               "note": "`achieved` counts EXECUTED flops: a sweep step is 108 FP64 instructions = 180 flops "
                                       "(A_d has two trivial rows, the inner steps of a sweep need no gain).  SURVEY 8(d)'s "
                                       "convention (dense 4x4 algebra without credit for structure: 550 (H-1) + 16 + 856 flops "
                                       "per solve) is under `survey_convention`; by it the kernel runs above the DFMA peak, "
                                       "which only says that the dense count is not what has to be computed.  Three "
                                       "consecutive solves share one pass over their common window rows (80 B of compact "
                                       "linearisation per row through a cp.async ring, read once for the three)"}
        if back:
            back["roofline_frac"] = back.pop("achieved_tflops_per_gpu") / fp64_meas
        flops_iter = (FLOPS_FIXED + FLOPS_PER_TRY * h["ntry"]) * (N_STEPS - 1)
        per_gpu_rate = value / world
        ach_tflops = per_gpu_rate * flops_iter / 1e12
        ach_gbs = per_gpu_rate * BYTES_STEP_ITER * (N_STEPS - 1) / 1e9
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic, traffic_src, pipe = None, "no committed ncu capture found under profiles/", None
        if a.traffic_bytes:
            traffic, traffic_src = a.traffic_bytes, "--traffic-bytes"
        elif ncu and ncu.get("launch", {}).get("batch") and ncu["launch"].get("iters"):
            lb, li, ls = ncu["launch"]["batch"], ncu["launch"]["iters"], ncu["launch"].get("n_steps", N_STEPS - 1)
            per_iter = ((ncu["dram_bytes_read"] + ncu["dram_bytes_write"]) / (lb * ls) - TRAFFIC_STEP_INIT) / li
            traffic = B * (N_STEPS - 1) * (per_iter * iters + TRAFFIC_STEP_INIT)
            traffic_src = ("%s (ncu --set full of %s, %d-iteration launch at B = %d: %.0f B per problem-step-iteration + %.0f B per "
                           "problem-step of the initial rollout), scaled to this launch" % (ncu["file"], ncu["kernel"], li, lb, per_iter, TRAFFIC_STEP_INIT))
        if ncu:
            pipe = ncu.get("pipe_fp64_pct")
        roof = {"bound": "fp64", "achieved": ach_tflops, "peak": fp64_meas, "unit": "TFLOP/s", "frac": ach_tflops / fp64_meas,
                "traffic": traffic, "traffic_source": traffic_src,
                "pipe_fp64_pct": pipe, "issue_active_pct": ncu.get("issue_active_pct") if ncu else None,
                "pipe_source": (ncu["file"] + ": sm__pipe_fp64_cycles_active / smsp__issue_active of the captured launch; `frac` counts "
                                "40 flops per sin/cos and dense 4x4 algebra (SURVEY 8d), the kernel spends 7-10 FP64 instructions per sin/cos") if ncu else None,
                "peak_source": "DFMA-chain microbenchmark (acro_bench_fp64_peak) run in this process; nominal %.1f" % FP64_NOMINAL_TFLOPS,
                "frac_of_nominal": ach_tflops / FP64_NOMINAL_TFLOPS,
                "kernel": bt.newton_kernel_name(B), "launch_ms": 1e3 * h["t"] / max(h["launches"], 1),
                "algorithmic_flops_per_launch": flops_iter * B * iters, "algorithmic_bytes_per_launch": BYTES_STEP_ITER * (N_STEPS - 1) * B * iters,
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6650"},
                "note": "B=4096 is 128 tiles of 32 problems on 148 SMs: every tile is a 1000-step dependent FP64 recurrence per iteration, "
                        "split between two warps on two SM sub-partitions; the warp that carries the recurrence issues one FP64 "
                        "instruction every 2 cycles and is the bound (see DESIGN.md section 4). Large batches (bench_configs.py, "
                        "strong.newton_total_262144) are HBM / FP64-issue bound"}
        cpu = None
        if world == 1 and not a.no_cpu:
            cpu = cpu_baseline_block(a.cpu_iters, a.ref_iters)
        target = None
        if longh:
            target = {"target": ">= 1e6 Newton iterations/s, T = 10k-step horizon-equivalent, on one B200 (BASELINE.json)",
                      "n501_reading": value / world,
                      "rescaled_10k_equivalent_at_this_batch": value / world * (N_STEPS - 1) / 1e4,
                      "real_10k_horizon": {k: v["value"] / world for k, v in longh.items() if k.startswith("batch_")},
                      "met_at_batch_4096": bool(longh["batch_4096"]["value"] / world >= 1e6),
                      "met_at_batch_16384": bool(longh["batch_16384"]["value"] / world >= 1e6),
                      "met_at_batch_32768": bool(longh["batch_32768"]["value"] / world >= 1e6)}
        line = {
            "metric": "newton_iterations_per_sec", "value": value, "unit": "Newton iterations/s (N=501 time steps each)",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * h["t"] / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a),
            "step_iters_per_sec": value * (N_STEPS - 1), "newton_iters_per_sec_10k_step_equivalent": value * (N_STEPS - 1) / 1e4,
            "e2e": {"value": e2e_rate["solution"], "unit": "Newton iterations/s", "h2d_bytes_per_step": e["h2d"],
                    "d2h_bytes_per_step": e["solution"]["d2h"], "ms_per_step": 1e3 * e["solution"]["t"] / a.steps,
                    "api": "trajectory_generation.newton_Algorithm(x0[B,4] pinned host, x_ref, u_ref, block=False, return_gains=False): "
                           "every step's inputs go host->device; x_traj, u_traj and the history (cost, max|sigma|, Armijo tries and "
                           "steps, iterations, status) come back to fresh pinned host tensors - what main.py's task_1 / task_2 consume; "
                           "K and sigma of the last iteration stay on the device as handles.  The copies of step i overlap the kernel "
                           "of step i+1 (double-buffered solver state)",
                    "full_returns": {"value": e2e_rate["full"], "ms_per_step": 1e3 * e["full"]["t"] / a.steps,
                                     "d2h_bytes_per_step": e["full"]["d2h"],
                                     "api": "the same call with return_gains=True: K (B,500,2,4) and sigma (B,500,2) come back too, "
                                            "i.e. everything the reference's function returns (twice the bytes; on an 8-GPU box the "
                                            "eight ranks then share the host's device-to-host ceiling, measured 88 GB/s aggregate: "
                                            "DESIGN.md section 5)"},
                    "sync": {"value": e2e_rate["sync"], "ms_per_step": 1e3 * e["sync"]["t"] / a.steps,
                             "d2h_bytes_per_step": e["sync"]["d2h"],
                             "api": "blocking calls (block=True), everything returned: solve, then copy"}},
            "host_binding": {"cpus_of_rank0": len(cx.cpus) if cx.cpus else None,
                             "note": "ranks of a multi-GPU run are pinned to the CPU cores NVML reports next to their GPU before pinned "
                                     "host memory is allocated (sharding.bind_to_gpu_numa)"},
            "gpu_launches": int(h["launches"]), "roofline": roof, "cpu_baseline": cpu, "clocks": h["clocks"],
            "backtracking": back, "mpc": mpc, "strong": strong, "long_horizon": longh, "target": target,
            "check": {"iterations_done_per_step_all_ranks": h["done_per_step"], "iterations_nominal_per_step": B * iters * world,
                      "armijo_tries_mean": h["ntry"], "mean_final_cost": h["cost_mean"], "e2e_mean_final_cost": e["full"]["cost_mean"]},
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        cx.dist.barrier()
        cx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=50, help="Newton iterations per problem per step")
    ap.add_argument("--cpu-iters", type=int, default=40, help="Newton iterations per problem in the CPU sample (NumPy port)")
    ap.add_argument("--ref-iters", type=int, default=3, help="Newton iterations per problem in the CPU sample (unmodified reference)")
    ap.add_argument("--port", action="store_true", help="--impl reference: time the NumPy port even if the reference imports")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-mpc", action="store_true", help="skip the secondary MPC solves/s measurement")
    ap.add_argument("--quick", action="store_true", help="headline + e2e (+ mpc) only: no backtracking / strong / long_horizon blocks")
    ap.add_argument("--traffic-bytes", type=float, default=None, help="dram bytes per launch from an ncu capture (overrides profiles/)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_native(a)


if __name__ == "__main__":
    main()
