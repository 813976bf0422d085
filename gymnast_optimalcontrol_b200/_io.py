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


_side = {}


def side_streams():
    """(upload stream, copy stream) of the current device: host->device staging of the next call and device->host
    copies of the previous one run there, next to the kernels on the caller's stream."""
    i = torch.cuda.current_device()
    if i not in _side:
        _side[i] = (torch.cuda.Stream(), torch.cuda.Stream())
    return _side[i]


def _tensors(objs):
    for o in objs:
        if isinstance(o, torch.Tensor):
            yield o
        elif isinstance(o, bt.Traj):
            yield o.data
        elif isinstance(o, bt.Ref):
            yield from _tensors((o.x, o.u))


class upload_scope:
    """`with upload_scope(enabled) as up:` - uploads and layout kernels inside run on the upload stream, so that
    they overlap whatever the caller's stream is still computing; on exit the caller's stream waits for them.
    `up.keep(...)` tells the allocator that the staged tensors are used on the caller's stream."""

    def __init__(self, enabled, *inputs):
        # inputs already on the device were produced on the caller's stream: they stay there
        self.enabled = bool(enabled) and not any(isinstance(a, torch.Tensor) and a.is_cuda for a in inputs)

    def __enter__(self):
        if self.enabled:
            self.main = torch.cuda.current_stream()
            self.up = side_streams()[0]
            self.ctx = torch.cuda.stream(self.up)
            self.ctx.__enter__()
        return self

    def keep(self, *objs):
        if self.enabled:
            for t in _tensors(objs):
                t.record_stream(self.main)

    def __exit__(self, *exc):
        if self.enabled:
            ev = torch.cuda.Event()
            ev.record(self.up)
            self.ctx.__exit__(*exc)
            self.main.wait_event(ev)
        return False


_deferred = []  # copy-back stages recorded but not yet queued (see out_async)


def flush_deferred():
    """Queue the copy-back stages that out_async(..., defer=True) held back."""
    while _deferred:
        _deferred.pop(0)()


class Pending:
    """Handle of a batched call whose results are on their way to the host; ``result()`` waits for the copies and
    returns what the blocking call returns."""

    def __init__(self, done, build):
        self._done, self._build, self._value = done, build, None

    def _event(self):
        if callable(self._done):  # a deferred copy-back: queue it now
            flush_deferred()
            self._done = self._done()
        return self._done

    def ready(self):
        return self._value is not None or self._event().query()

    def result(self):
        if self._value is None:
            self._event().synchronize()
            self._value = self._build()
            self._build = None
        return self._value


def out_async(items, kind, defer=False):
    """items: [(Traj or device tensor, pad_rows or None)] produced on the current stream -> Pending of the list of
    host (or device, for CUDA callers) results.  Un-tiling and the device-to-host copies run on the copy stream.

    defer=True: the copy-back is only RECORDED here and queued by the next `flush_deferred()` - the next non-blocking
    call does that right after queueing its own uploads - or by `result()` / `ready()`.  Work queued on the copy stream
    waits for the solver kernel; with the driver's default of 8 hardware queues the upload stream may share a queue with
    the copy stream, and uploads queued behind that wait would start only after the kernel (measured:
    profiles/r2_mpc_e2e_probe.txt)."""
    main, cs = torch.cuda.current_stream(), side_streams()[1]
    ready = torch.cuda.Event()
    ready.record(main)
    host = not (kind.torch and kind.cuda)
    outs, box = [], {}
    for t_soa, _ in items:
        for t in _tensors((t_soa,)):
            t.record_stream(cs)

    def start():
        if "done" in box:
            return
        with torch.cuda.stream(cs):
            cs.wait_event(ready)
            for t_soa, pad in items:
                t = bt.unpack_soa(t_soa)
                if not kind.batched:
                    t = t[0]
                if host:
                    rows = t.shape[-2] if pad is None else max(pad, t.shape[-2])
                    h = pinned_empty(t.shape[:-2] + (rows, t.shape[-1]), t.dtype)
                    if rows > t.shape[-2]:
                        h[..., t.shape[-2]:, :] = 0.0
                    h[..., :t.shape[-2], :].copy_(t, non_blocking=True)
                    outs.append((h, t))
                else:
                    if pad is not None and pad > t.shape[-2]:
                        tp = t.new_zeros(*t.shape[:-2], pad, t.shape[-1])
                        tp[..., :t.shape[-2], :] = t
                        t = tp
                    t.record_stream(main)
                    outs.append((t, None))
            box["done"] = torch.cuda.Event()
            box["done"].record(cs)
        if not host:
            main.wait_event(box["done"])

    def build():
        return [(h if kind.torch else h.numpy()) if host else h for h, _ in outs]

    if defer and host:
        _deferred.append(start)
        return Pending(lambda: box["done"], build)
    start()
    return Pending(box["done"], build)


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
