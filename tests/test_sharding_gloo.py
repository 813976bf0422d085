"""The N>1 host logic on CPU: world_size-2 (and 3) gloo process groups, ragged shards, global problem order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gymnast_optimalcontrol_b200 import sharding


def test_shard_bounds_cover_the_batch():
    for n, w in ((4096, 8), (4097, 8), (5, 8), (1000000, 8), (7, 2), (1, 1)):
        prev = 0
        for r in range(w):
            lo, hi = sharding.shard_bounds(n, w, r)
            assert lo == prev and hi >= lo
            prev = hi
        assert prev == n
        sizes = sharding.shard_sizes(n, w)
        assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_bounds(n, world, rank)
        idx = torch.arange(lo, hi, dtype=torch.float64)
        # a fake per-problem summary that encodes the global problem index in every field
        local = sharding.pack_summary(cost=1000.0 + idx, status=(idx % 4).to(torch.int32), iters=(idx * 3).to(torch.int32),
                                      gamma_acc=0.1 * idx, sigma_norm=1e-4 * idx)
        full = sharding.gather_summary(local, n)
        out = sharding.unpack_summary(full)
        g = torch.arange(n, dtype=torch.float64)
        ok = (torch.equal(out["cost"], 1000.0 + g) and torch.equal(out["status"], (g % 4).to(torch.int32))
              and torch.equal(out["iters"], (g * 3).to(torch.int32)) and torch.equal(out["gamma_acc"], 0.1 * g)
              and torch.equal(out["sigma_norm"], 1e-4 * g))
        q.put((rank, bool(ok), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 4096), (2, 4097), (3, 10)])
def test_gather_summary_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok and shape == (5, n), (rank, ok, shape)
