#!/usr/bin/env python
"""bench.py - Newton iterations/s of the batched acrobot hot path (BASELINE.json metric) on N B200s.

A "step" is one pass of the hot path over one batch: `--iters` regularised-Newton iterations (backward
affine Riccati sweep + Armijo closed-loop rollouts, trajectory_generation.py:329-396) for each of the
`--batch` independent swing-up problems of config 2 (B = 4096 randomised initial states, N = 501 time
steps as shipped, gamma_0 = 0.1, weights of trajectory_generation.py:16-18), starting from the open-loop
rollout (tol = 0: no early exit, every problem runs every iteration).

    value   device-resident throughput: inputs already in HBM, CUDA events around K steps
    e2e     the same through the drop-in trajectory_generation.newton_Algorithm with HOST (pinned) buffers:
            per step H2D of x0 / x_ref / u_ref and D2H of x_traj, u_traj, K, sigma and the history
    roofline / cpu_baseline / clocks: see DESIGN.md section "Measurement"

Multi-GPU (torchrun): every rank solves its own 4096 problems (weak scaling, no collective on the hot
path); one NCCL gather of the per-problem summary (cost, status, iterations) per step is inside the timed
region.  `--impl reference` times the CPU oracle port (the reference is pure Python / NumPy and cannot
travel to the GPU box) on rank 0's host cores.
"""
import argparse
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
# dram__bytes_read+write of acro::k_newton_duo from the ncu --set full capture in profiles/ (4-iteration launch,
# B = 4096: 3.60 GB): 396 B per problem-step-iteration (the kernel moves 464 B by design, L2 absorbs part of the
# re-reads) + 176 B per problem-step for the initial rollout and cost
TRAFFIC_STEP_ITER, TRAFFIC_STEP_INIT = 396.0, 176.0
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


# ----------------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    x0, iters = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import acro_oracle as O
    x_ref, u_ref = load_reference_trajectory()
    t = time.perf_counter()
    x, u, K, s, h = O.newton_Algorithm(x0, x_ref, u_ref, max_iters=iters, tol=0.0, gamma_0=0.1)
    return time.perf_counter() - t, h["iters"], float(h["cost"][-1])


def cpu_newton_rate(pool, cores, iters_per_problem, seed=0):
    """`cores` problems of the workload, one per worker, `iters_per_problem` Newton iterations each."""
    x0 = make_x0(4096, 0)[seed * cores:(seed + 1) * cores]
    t = time.perf_counter()
    res = pool.map(_cpu_worker, [(x0[i], iters_per_problem) for i in range(cores)])
    wall = time.perf_counter() - t
    done = sum(r[1] for r in res)
    return done / wall, wall, done


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


def measure_mpc(bt, torch, dist, rank, world, steps, with_cpu):
    """Second half of the BASELINE.json metric: MPC solves/s on config 4 (B = 16 384 acrobots per GPU, horizon 75,
    500 receding-horizon steps, every problem its own reference copy so that every problem runs its own Riccati
    sweeps: 8.19 M solves per run and GPU)."""
    d = np.load(os.path.join(ROOT, "tests", "golden", "acrobot_optimal_trajectory.npz"))
    B, H, N = 16384, 75, 501
    w = bt.mpc_weights()
    xf = bt.upload(np.array([[np.pi], [0], [0], [0]], dtype=np.float64))
    A_f, B_f = bt.linearize(xf, bt.upload(np.zeros((2, 1))), discrete=True)
    P, _ = bt.p_inf(A_f, B_f, w)
    QT = P[:, :, 0].contiguous()
    x0h = d["x"][0] + np.random.default_rng(3 + rank).uniform(-0.1, 0.1, (B, 4))
    x0 = bt.upload(np.ascontiguousarray(x0h.T))
    refp = bt.Ref(bt.Traj.from_batch_major(bt.upload(np.repeat(d["x"][None], B, 0))),
                  bt.Traj.from_batch_major(bt.upload(np.repeat(d["u"][None], B, 0))))
    res = {}

    def run():
        res["s"] = bt.mpc_track(x0, refp, QT, T=N, T_pred=H, w=w)
    for _ in range(2):
        run()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ns = res["s"][3]
    value = ns * steps * world / float(t[0])
    flops = 550.0 * (H - 1) + 16 + 856
    out = {"metric": "mpc_solves_per_sec", "value": value,
           "unit": "MPC solves/s (one (H-1)-step Riccati sweep on the sliding window + plant step each)",
           "config": {"workload": "config 4: receding-horizon MPC tracking of acrobot_optimal_trajectory.npz, B=16384 acrobots per GPU, "
                                  "horizon 75, 500 steps, per-problem references (every solve executed)",
                      "solves_per_run_per_gpu": ns},
           "ms_per_run": 1e3 * float(t[0]) / steps, "flops_per_solve": flops,
           "achieved_tflops_per_gpu": value / world * flops / 1e12}
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


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    ipp = a.cpu_iters
    with mp.get_context("fork").Pool(cores) as pool:
        for _ in range(min(a.warmup, 1)):
            cpu_newton_rate(pool, cores, 1)
        t0 = time.perf_counter()
        done = 0
        for s in range(a.steps):
            r, wall, d = cpu_newton_rate(pool, cores, ipp, seed=s % 8)
            done += d
        total = time.perf_counter() - t0
    value = done / total
    sample = "%d problems of the 4096 (one per core) x %d Newton iterations per step" % (cores, ipp)
    print(json.dumps({
        "impl": "reference", "metric": "newton_iterations_per_sec", "value": value,
        "unit": "Newton iterations/s (N=501 time steps each)", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, sample_note=sample),
        "cpu_baseline": {"value": value, "unit": "Newton iterations/s", "cores": cores, "kind": "port", "sample": sample,
                         "note": "NumPy oracle port (oracle/acro_oracle.py), one problem per process like the reference's "
                                 "per-problem Python loop; the reference itself is Python+SymPy and is not on this box"},
        "e2e": {"value": value, "unit": "Newton iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(a, sample_note=None):
    c = {"workload": "config 2: batched Newton+Armijo swing-up solves, B=%d problems per GPU from randomised initial states, "
                     "N=501 (T=10 s, dt=0.02 as shipped), %d Newton iterations per step, gamma_0=0.1, tol=0 (no early exit)"
                     % (a.batch, a.iters),
         "batch_per_gpu": a.batch, "horizon_steps": N_STEPS - 1, "newton_iters_per_step": a.iters,
         "l2": "working set 525 MB per GPU (X,U,Xw,Uw,lin,K,S) > 126 MB L2, no flush needed"}
    if sample_note:
        c["cpu_sample"] = sample_note
    return c


def kernel_name(batch):
    tiles = (batch + 31) // 32
    if tiles <= 296:
        return "acro::k_newton_duo<false,false,%d> (two warps per tile: recurrence warp + trailer warp, TMA-fed ring)" % (16 if tiles <= 148 else 4)
    return "acro::k_newton_ring<false,false,%d> (one warp per tile, TMA-fed shared-memory ring)" % (4 if tiles <= 592 else 2)


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


def run_native(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # stdout carries the one JSON line and nothing else: whatever libraries print to file descriptor 1 (NCCL's
    # version banner when NCCL_DEBUG is set in the environment) goes to stderr; the line is written to the saved fd
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gymnast_optimalcontrol_b200 import _abi
    from gymnast_optimalcontrol_b200 import batched as bt
    from gymnast_optimalcontrol_b200 import sharding
    from gymnast_optimalcontrol_b200 import trajectory_generation as tg

    B, iters = a.batch, a.iters
    x_ref, u_ref = load_reference_trajectory()
    x0 = make_x0(B, rank)
    ref = bt.make_ref(x_ref, u_ref)
    w = bt.newton_weights()
    x0d = bt.upload(np.ascontiguousarray(x0.T))
    state = bt.newton_alloc(B, N_STEPS, iters, history=True)
    gathered = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        state.initialised = False
        bt.newton_solve(x0d, ref, max_iters=iters, tol=0.0, gamma_0=0.1, w=w, state=state)
        if world > 1:  # the only collective: the per-problem summary of every shard, over NCCL/NVLink
            summ = sharding.pack_summary(state.cost, state.status, state.iters, state.gamma_acc, state.sigma_norm)
            gathered["summary"] = sharding.gather_summary(summ, B * world)

    # ---- device-resident throughput (the clock sampler runs from before the warm-up: nvidia-smi takes its time to start)
    with ClockSampler(local) as clk:
        clk.wait_first_row()
        for _ in range(a.warmup):
            step_device()
        barrier()
        l0 = _abi.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0 = clk.mark()
        e0.record()
        for _ in range(a.steps):
            step_device()
        e1.record()
        barrier()
        c1 = clk.mark()
    clocks = clk.summary(c0, c1)
    launches = _abi.launch_count() - l0
    t_dev = e0.elapsed_time(e1) * 1e-3
    done_iters = int(state.iters.sum().item())
    ntry_mean = float(state.hist_ntry[:iters].double().mean().item())
    cost_mean = float(state.cost.mean().item())

    # ---- end to end through the drop-in with host buffers
    x0_host = torch.from_numpy(x0).pin_memory()
    xr_host = torch.from_numpy(x_ref).pin_memory()
    ur_host = torch.from_numpy(u_ref).pin_memory()

    def step_e2e():
        out = tg.newton_Algorithm(x0_host, xr_host, ur_host, max_iters=iters, tol=0.0, gamma_0=0.1, verbose=False)
        return out

    for _ in range(max(1, min(a.warmup, 2))):
        out = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        out = step_e2e()
    barrier()
    t_e2e = time.perf_counter() - t0
    h2d = x0_host.numel() * 8 + xr_host.numel() * 8 + ur_host.numel() * 8
    d2h = sum(o.numel() * 8 for o in out[:4]) + sum(np.asarray(v).nbytes for k, v in out[4].items()
                                                    if k in ("cost", "sigma_norm", "iters", "status", "n_try", "gamma"))
    e2e_cost_mean = float(out[4]["cost"][:, -1].mean())

    times = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_dev, t_e2e = float(times[0]), float(times[1])
    total_iters = B * iters * a.steps * world
    value = total_iters / t_dev
    e2e_value = total_iters / t_e2e

    mpc = None if a.no_mpc else measure_mpc(bt, torch, dist, rank, world, max(2, a.steps), with_cpu=(world == 1 and rank == 0 and not a.no_cpu))
    if rank == 0:
        peaks = measured_peaks()
        fp64_meas = fp64_peak_tflops(_abi, torch)
        if mpc:
            ach = mpc.pop("achieved_tflops_per_gpu")
            mpc["roofline"] = {"bound": "fp64", "achieved": ach, "peak": fp64_meas, "unit": "TFLOP/s", "frac": ach / fp64_meas,
                               "note": "550 (H-1) + 16 + 856 flops per solve (SURVEY 8d); the sweeps re-read 80 B of compact "
                                       "linearisation per window step from L2"}
        flops_iter = (FLOPS_FIXED + FLOPS_PER_TRY * ntry_mean) * (N_STEPS - 1)
        per_gpu_rate = value / world
        ach_tflops = per_gpu_rate * flops_iter / 1e12
        ach_gbs = per_gpu_rate * BYTES_STEP_ITER * (N_STEPS - 1) / 1e9
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        roof = {"bound": "fp64", "achieved": ach_tflops, "peak": fp64_meas, "unit": "TFLOP/s", "frac": ach_tflops / fp64_meas,
                "traffic": a.traffic_bytes if a.traffic_bytes else B * (N_STEPS - 1) * (TRAFFIC_STEP_ITER * iters + TRAFFIC_STEP_INIT),
                "traffic_source": "profiles/r1_newton_duo_ncu_summary.txt (ncu --set full, 4-iteration launch) scaled to this launch",
                "peak_source": "DFMA-chain microbenchmark (acro_bench_fp64_peak) run in this process; nominal %.1f" % FP64_NOMINAL_TFLOPS,
                "frac_of_nominal": ach_tflops / FP64_NOMINAL_TFLOPS,
                "kernel": kernel_name(B), "launch_ms": 1e3 * t_dev / max(launches, 1),
                "algorithmic_flops_per_launch": flops_iter * B * iters, "algorithmic_bytes_per_launch": BYTES_STEP_ITER * (N_STEPS - 1) * B * iters,
                "hbm": {"achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6650"},
                "note": "B=4096 is 128 tiles of 32 problems on 148 SMs: every tile is a 1000-step dependent FP64 recurrence per iteration, "
                        "split between two warps on two SM sub-partitions; the warp that carries the recurrence issues one FP64 "
                        "instruction every 2 cycles and is the bound (see DESIGN.md section 4). Large batches (bench_configs.py) "
                        "are HBM bound at 89% of the measured copy bandwidth"}
        cpu = None
        if world == 1 and not a.no_cpu:
            cores = os.cpu_count() or 1
            with mp.get_context("fork").Pool(cores) as pool:
                r, wall, d = cpu_newton_rate(pool, cores, a.cpu_iters)
            cpu = {"value": r, "unit": "Newton iterations/s", "cores": cores, "kind": "port",
                   "sample": "%d problems of the 4096 (one per core) x %d Newton iterations, %.1f s of wall time" % (cores, a.cpu_iters, wall)}
        line = {
            "metric": "newton_iterations_per_sec", "value": value, "unit": "Newton iterations/s (N=501 time steps each)",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_dev / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(a),
            "step_iters_per_sec": value * (N_STEPS - 1), "newton_iters_per_sec_10k_step_equivalent": value * (N_STEPS - 1) / 1e4,
            "e2e": {"value": e2e_value, "unit": "Newton iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * t_e2e / a.steps, "api": "trajectory_generation.newton_Algorithm(x0[B,4] pinned host, x_ref, u_ref)"},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "mpc": mpc,
            "check": {"iterations_done_last_step_rank0": done_iters, "armijo_tries_mean": ntry_mean, "mean_final_cost": cost_mean,
                      "e2e_mean_final_cost": e2e_cost_mean},
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--iters", type=int, default=50, help="Newton iterations per problem per step")
    ap.add_argument("--cpu-iters", type=int, default=40, help="Newton iterations per problem in the CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-mpc", action="store_true", help="skip the secondary MPC solves/s measurement")
    ap.add_argument("--traffic-bytes", type=float, default=None, help="dram bytes per launch from the ncu capture (profiles/)")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_native(a)


if __name__ == "__main__":
    main()
