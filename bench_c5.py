#!/usr/bin/env python
"""Config 5 of BASELINE.json: the 200-point Armijo step-size sweep (trajectory_generation.py:257-264) for 5000 base
iterates = 1 000 000 closed-loop rollouts, sharded over the GPUs of one box (strong scaling: the 5000 iterates are
split into contiguous blocks, one per rank; no collective on the hot path, one NCCL gather of the per-iterate
minimising step size for the summary).

    python bench_c5.py                                                       (one GPU)
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench_c5.py --gpus 8

One JSON line (rank 0): rollouts/s over all GPUs, device timing (CUDA events, max over ranks)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iterates", type=int, default=5000)
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from gymnast_optimalcontrol_b200 import batched as bt
    from gymnast_optimalcontrol_b200 import sharding

    fa = np.load(os.path.join(ROOT, "tests", "golden", "fully_actuated_trajectory.npz"))
    u_ref = np.zeros(fa["u"].shape)
    u_ref[:, 1] = 2.0 * fa["u"][:, 1]
    ref = bt.make_ref(fa["x"], u_ref)
    w = bt.newton_weights()
    P, S, N = a.iterates, 200, 501
    lo, hi = sharding.shard_bounds(P, world, rank)
    # base iterates: Newton iterate 3 of the config-2 problems lo..hi-1 (x, u, K, sigma stored), as in SURVEY 8(d)
    x0 = np.random.default_rng(1).uniform(-0.2, 0.2, (P, 4))[lo:hi]
    st = bt.newton_solve(bt.upload(np.ascontiguousarray(x0.T)), ref, max_iters=3, tol=0.0, gamma_0=0.1, history=False)
    K, Sg, dJ, sn = bt.riccati_affine(st.X, st.U, ref, w)
    steps = bt.upload(np.linspace(0.0, 1.25, S))  # tg:257-258
    out = {}

    def run():
        cost = bt.stepsize_sweep(st.X, st.U, K, Sg, ref, w, steps)   # (S, p)
        best = steps[torch.argmin(torch.nan_to_num(cost, nan=float("inf")), dim=0)]
        if world > 1:  # summary only: the minimising step size of every base iterate
            out["best"] = sharding.gather_summary(best[None], P)
        else:
            out["best"] = best[None]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(a.warmup):
        run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        run()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        rate = P * S * a.steps / float(t[0])
        line = {"metric": "sweep_rollouts_per_sec", "value": rate, "unit": "closed-loop rollouts/s (N=501)", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * float(t[0]) / a.steps, "higher_is_better": True,
                "scaling": "strong", "dtype": "f64", "data": "synthetic",
                "config": {"workload": "config 5: %d base iterates x %d step sizes = %d closed-loop rollouts + costs, sharded over %d GPU(s)"
                                       % (P, S, P * S, world)},
                "roofline": {"bound": "fp64", "achieved_per_gpu": rate / world * 940.0 * (N - 1) / 1e12, "unit": "TFLOP/s",
                             "flops_per_unit": 940.0 * (N - 1)},
                "check": {"median_minimising_step": float(out["best"].median().item()), "iterates_gathered": int(out["best"].shape[1])}}
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
