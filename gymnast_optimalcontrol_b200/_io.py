"""Host <-> device marshalling for the drop-in modules.

The reference works on NumPy arrays of one problem: state (4,), trajectory (N,4), lists of (2,4) gains.
The drop-ins accept those, the same with a leading batch axis ((B,4), (B,N,4), ...), and torch tensors
(CPU, pinned or not, or CUDA).  Everything is converted to the structure-of-arrays device layout with
the library's own pack/unpack kernels and converted back to the kind of object the caller passed in.
"""
import numpy as np
import torch

from . import batched as bt


class Kind:
    """Remembers what the caller handed in so results go back the same way."""

    def __init__(self, a, batched):
        self.batched = batched
        if isinstance(a, torch.Tensor):
            self.torch = True
            self.cuda = a.is_cuda
            self.pinned = (not a.is_cuda) and a.is_pinned()
        else:
            self.torch = self.cuda = self.pinned = False


class PinnedPool:
    """Pinned host memory for results.  A block is handed out again only when nothing made from it is alive any more
    (tensor views, NumPy arrays: the reference count of its storage is back to that of the pool's own handle), so two
    results never alias - and a loop that drops its previous results never page-locks new memory, which costs
    milliseconds and synchronises the device (torch's own caching host allocator did allocate in the steady state of the
    pipelined solves: profiles/r2_e2e_probe.txt)."""
    LIMIT = 8 << 30  # bytes kept; idle blocks beyond that are released

    def __init__(self):
        self.blocks = []
        probe = torch.empty(8, dtype=torch.uint8)
        self._count = getattr(torch._C, "_storage_Use_Count", None)
        self._base = self._uses(probe) if self._count else 0

    def _uses(self, b):
        return self._count(b.untyped_storage()._cdata)

    def empty(self, shape, dtype=torch.float64):
        shape = tuple(int(v) for v in shape)
        n = int(np.prod(shape, dtype=np.int64)) * torch.empty(0, dtype=dtype).element_size()
        if self._count is None or n == 0:
            return torch.empty(shape, dtype=dtype, pin_memory=True)
        for b in self.blocks:
            if n <= b.numel() <= 2 * n + 4096 and self._uses(b) <= self._base:
                return b[:n].view(dtype).view(shape)
        if sum(b.numel() for b in self.blocks) + n > self.LIMIT:
            self.blocks = [b for b in self.blocks if self._uses(b) > self._base]
        b = torch.empty((n + 4095) // 4096 * 4096, dtype=torch.uint8, pin_memory=True)
        self.blocks.append(b)
        return b[:n].view(dtype).view(shape)


_pool = PinnedPool()


def pinned_empty(shape, dtype=torch.float64):
    """A pinned host tensor for a result that no live result shares memory with (PinnedPool)."""
    return _pool.empty(shape, dtype)


def to_host(t, stream_sync=True):
    """Device tensor -> fresh pinned host tensor (asynchronous copy on the current stream, then one synchronise)."""
    h = pinned_empty(t.shape, t.dtype)
    h.copy_(t, non_blocking=True)
    if stream_sync:
        torch.cuda.current_stream().synchronize()
    return h


def as_device(a):
    return bt.upload(a)


def state_in(a, C):
    """(C,), (C,1), (B,C) -> SoA (C,B) on the device, Kind"""
    if isinstance(a, torch.Tensor):
        nd, shape = a.dim(), tuple(a.shape)
    else:
        a = np.asarray(a, dtype=np.float64)
        nd, shape = a.ndim, a.shape
    if nd == 1 or (nd == 2 and shape == (C, 1)):
        k = Kind(a, False)
        d = as_device(a).reshape(1, C)
    elif nd == 2 and shape[1] == C:
        k = Kind(a, True)
        d = as_device(a)
    else:
        raise ValueError("expected shape (%d,) or (B,%d), got %r" % (C, C, shape))
    return bt.pack_soa(d), k


def traj_in(a, C):
    """(T,C) or (B,T,C) (or a list of T arrays of shape (C,) / (r,c) with r*c = C) -> Traj (tiled device layout), Kind"""
    if isinstance(a, (list, tuple)):
        a = np.asarray([np.asarray(v, dtype=np.float64).reshape(-1) for v in a])
    if isinstance(a, torch.Tensor):
        nd = a.dim()
    else:
        a = np.asarray(a, dtype=np.float64)
        nd = a.ndim
    d = as_device(a)
    if d.shape[-1] != C:  # e.g. gains given as (..., 2, 4)
        d = d.reshape(*d.shape[:-2], C)
        nd -= 1
    if nd == 2:
        return bt.Traj.from_batch_major(d.reshape(1, *d.shape)), Kind(a, False)
    if nd == 3:
        return bt.Traj.from_batch_major(d), Kind(a, True)
    raise ValueError("expected a trajectory of shape (T,%d) or (B,T,%d)" % (C, C))


def out(t_soa, kind, tail=None, key="out"):
    """Device point batch (C,B) or Traj -> caller's kind; tail reshapes the component axis (e.g. (2,4)).
    Host results are always fresh buffers (`key` is kept for call-site readability only)."""
    t = bt.unpack_soa(t_soa)  # (B,C) / (B,T,C)
    if tail is not None:
        t = t.reshape(*t.shape[:-1], *tail)
    if not kind.batched:
        t = t[0]
    if kind.torch and kind.cuda:
        return t
    h = to_host(t)
    return h if kind.torch else h.numpy()


def vec_out(t, kind):
    """per-problem scalars (B,) -> float / array following kind"""
    if not kind.batched:
        return t[0].item()
    if kind.torch and kind.cuda:
        return t
    h = to_host(t)
    return h if kind.torch else h.numpy()
