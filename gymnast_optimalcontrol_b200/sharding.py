"""Batch sharding over the GPUs of one box (one process per GPU, torch.distributed).

The problems are independent, so the hot path needs no exchange: every rank owns a contiguous block of the
batch and runs the same kernels on it.  The only collective is one gather of the per-problem summary
(final cost, status, iteration count, accepted step, max|sigma|) - 40 bytes per problem - over NCCL/NVLink
(gloo in the CPU tests).  Results for problem i do not depend on which rank owns it; they depend on the world size only
through the Newton kernel variant, which is chosen from the per-rank batch size (the variants agree to 1e-9 relative and
take identical Armijo decisions in the parity tests; `kernel=` pins one for bit-identical results across world sizes).
"""
import torch
import torch.distributed as dist

SUMMARY_FIELDS = ("cost", "status", "iters", "gamma_acc", "sigma_norm")


def shard_bounds(n_problems, world_size, rank):
    """Contiguous block [lo, hi) of rank `rank`: the first (n mod W) ranks get one extra problem."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of size %d" % (rank, world_size))
    base, extra = divmod(n_problems, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_problems, world_size):
    return [shard_bounds(n_problems, world_size, r)[1] - shard_bounds(n_problems, world_size, r)[0] for r in range(world_size)]


def pack_summary(cost, status, iters, gamma_acc, sigma_norm):
    """(5, b) float64: the int32 fields are exactly representable."""
    return torch.stack([cost.double(), status.double(), iters.double(), gamma_acc.double(), sigma_norm.double()])


def unpack_summary(summary):
    return {"cost": summary[0], "status": summary[1].to(torch.int32), "iters": summary[2].to(torch.int32),
            "gamma_acc": summary[3], "sigma_norm": summary[4]}


def gather_summary(local, n_problems, group=None):
    """All-gather the per-problem summaries of every rank's shard into one (F, n_problems) tensor in global
    problem order.  `local` is (F, b_rank) on this rank's device; shards may be ragged (padded for the collective)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_problems, world)
    if local.shape[1] != sizes[dist.get_rank(group)]:
        raise ValueError("this rank holds %d problems, its shard has %d" % (local.shape[1], sizes[dist.get_rank(group)]))
    width = max(sizes)
    F = local.shape[0]
    send = local
    if local.shape[1] != width:
        send = torch.zeros(F, width, dtype=local.dtype, device=local.device)
        send[:, :local.shape[1]] = local
    recv = torch.empty(world, F, width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv.view(world * F, width), send.contiguous(), group=group)
    return torch.cat([recv[r, :, :sizes[r]] for r in range(world)], dim=1)


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPU cores next to GPU `device_index` (NVML's CPU affinity of the device), BEFORE any pinned
    host memory is allocated: page-locked result buffers then live on the GPU's own NUMA node, and the device-to-host
    copies of the ranks of an 8-GPU box do not all cross to one socket.  Returns the CPU list, or None if NVML or the
    affinity call is not available (nothing is changed then)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1 and 64 * i + b < n_cpu]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001  (no NVML, no permission, not Linux: run unbound)
        return None
