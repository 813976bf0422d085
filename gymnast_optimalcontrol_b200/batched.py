"""Device-level API: torch CUDA tensors in the structure-of-arrays layout of include/acro_abi.h.

Every function here is a thin wrapper that allocates outputs with torch and calls one C-ABI
entry point of libacro_b200.so on torch's current stream.

    point batches   plain tensors, problem index last:  state (4, B), input (2, B), scalars (B,)
    time-indexed    ``Traj`` objects: a tensor of shape (ntiles, T, C, 32) in the tiled layout
                    A[tile][t][c][lane] of include/acro_abi.h plus the true batch size B
                    (X: C=4, U: C=2, K: C=8 = 2x4 row-major, S: C=2)

torch is used for device memory and streams only.  There is no CPU path: without a CUDA
device these functions raise.
"""
import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import _abi
from ._abi import AcroNewtonOpts, AcroParams, AcroRef, AcroWeights, call

F64 = torch.float64

# dynamics.py:15-61
PARAM_SETS = {
    1: dict(m1=1.0, m2=1.0, l1=1.0, lc1=0.5, l2=1.0, lc2=0.5, I1=0.33, I2=0.33, g=9.81, f1=1.0, f2=1.0),
    2: dict(m1=2.0, m2=2.0, l1=1.5, lc1=0.75, l2=1.5, lc2=0.75, I1=1.5, I2=1.5, g=9.81, f1=1.0, f2=1.0),
    3: dict(m1=1.5, m2=1.5, l1=2.0, lc1=1.0, l2=2.0, lc2=1.0, I1=2.0, I2=2.0, g=9.81, f1=1.0, f2=1.0),
}
DT = 2e-2  # dynamics.py:173


def make_params(version=1, dt=DT, actuated_tau1=False, **overrides):
    d = dict(PARAM_SETS[version])
    d.update(overrides)
    p = AcroParams()
    for k, v in d.items():
        setattr(p, k, float(v))
    p.dt = float(dt)
    p.actuated_tau1 = 1 if actuated_tau1 else 0
    return p


DEFAULT_PARAMS = make_params()


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("gymnast_optimalcontrol_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Traj:
    """A time-indexed batch array in the device layout A[tile][t][c][lane] (tile = b // 32, lane = b % 32)."""
    __slots__ = ("data", "B")

    def __init__(self, data, B):
        if data.dim() != 4 or data.shape[3] != 32 or data.shape[0] != ntiles(B):
            raise ValueError("expected shape (%d, T, C, 32) for B = %d, got %r" % (ntiles(B), B, tuple(data.shape)))
        self.data, self.B = data, int(B)

    T = property(lambda self: self.data.shape[1])
    C = property(lambda self: self.data.shape[2])

    @staticmethod
    def empty(T, C, B):
        return Traj(_empty(ntiles(B), T, C, 32), B)

    @staticmethod
    def zeros(T, C, B):
        return Traj(torch.zeros(ntiles(B), T, C, 32, dtype=F64, device=device()), B)

    @staticmethod
    def from_batch_major(a):
        """(B, T, C) CUDA tensor -> Traj (acro_pack_soa)."""
        a = a.contiguous()
        Bn, T, Cn = a.shape
        out = Traj.empty(T, Cn, Bn)
        call("acro_pack_soa", Bn, T, Cn, _p(a), _p(out), _stream())
        return out

    def batch_major(self):
        """-> (B, T, C) CUDA tensor (acro_unpack_soa)."""
        out = _empty(self.B, self.T, self.C)
        call("acro_unpack_soa", self.B, self.T, self.C, _p(self), _p(out), _stream())
        return out

    def clone(self):
        return Traj(self.data.clone(), self.B)


def ntiles(B):
    return (int(B) + 31) // 32


def _p(t, dtype=F64):
    if t is None:
        return C.c_void_p(0)
    if isinstance(t, Traj):
        t = t.data
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise TypeError("expected a contiguous CUDA %s tensor, got %r" % (dtype, type(t) if not isinstance(t, torch.Tensor)
                                                                          else (t.device, t.dtype, t.is_contiguous())))
    return C.c_void_p(t.data_ptr())


def _empty(*shape, dtype=F64):
    return torch.empty(shape, dtype=dtype, device=device())


class Weights:
    """Q, R, Q_T (host, symmetric) with optional per-problem device overrides (16,B), (4,B), (16,B)."""

    def __init__(self, Q, R, QT=None, Q_b=None, R_b=None, QT_b=None):
        self.Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(4, 4)
        self.R = np.ascontiguousarray(R, dtype=np.float64).reshape(2, 2)
        self.QT = np.ascontiguousarray(self.Q if QT is None else QT, dtype=np.float64).reshape(4, 4)
        for name, M in (("Q", self.Q), ("R", self.R), ("Q_T", self.QT)):
            if not np.array_equal(M, M.T):
                raise ValueError("%s must be symmetric" % name)
        self.Q_b, self.R_b, self.QT_b = Q_b, R_b, QT_b
        s = AcroWeights()
        s.Q[:] = self.Q.ravel().tolist()
        s.R[:] = self.R.ravel().tolist()
        s.QT[:] = self.QT.ravel().tolist()
        s.Q_b, s.R_b, s.QT_b = _p(Q_b).value, _p(R_b).value, _p(QT_b).value
        self.struct = s

    @property
    def per_problem(self):
        return any(t is not None for t in (self.Q_b, self.R_b, self.QT_b))

    def ref(self):
        return C.byref(self.struct)


# trajectory_generation.py:16-18, trajectory_tracking.py:173-175, 38-39
def newton_weights():
    return Weights(np.diag([130.0, 30.0, 0.0001, 0.0001]), np.diag([1e-6, 1.5]), np.diag([130.0, 130.0, 1.0, 1.0]))


def lqr_weights():
    return Weights(np.diag([100.0, 100.0, 10.0, 10.0]), np.diag([1.0, 1.0]))


def mpc_weights():
    return Weights(np.diag([120.0, 100.0, 0.0001, 0.0001]), np.diag([1e-6, 10.0]))


class Ref:
    """Reference trajectory: shared tensors x (N,4), u (N-1,2), or per problem Traj x (C=4), u (C=2)."""

    def __init__(self, x, u):
        self.x, self.u = x, u
        self.per_problem = isinstance(x, Traj)
        if self.per_problem != isinstance(u, Traj):
            raise ValueError("x_ref and u_ref must both be shared or both per problem")
        self.N = x.T if self.per_problem else x.shape[0]
        nu_ = u.T if self.per_problem else u.shape[0]
        if nu_ != self.N - 1:
            raise ValueError("Incompatible dimensions: x_ref has %d states but u_ref has %d controls (expected %d)"
                             % (self.N, nu_, self.N - 1))
        self.B = x.B if self.per_problem else None
        s = AcroRef()
        s.x, s.u, s.per_problem = _p(x).value, _p(u).value, int(self.per_problem)
        self.struct = s

    def ref(self):
        return C.byref(self.struct)


def upload(a):
    """NumPy / CPU tensor -> contiguous CUDA float64 tensor (async when the source is pinned)."""
    if isinstance(a, torch.Tensor):
        if a.is_cuda:
            return a.to(F64).contiguous()
        return a.to(F64).contiguous().to(device(), non_blocking=a.is_pinned())
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device())


def make_ref(x_ref, u_ref):
    """Shared (N,4)/(N-1,2) or batch-major per-problem (B,N,4)/(B,N-1,2) host or device arrays -> Ref."""
    xd, ud = upload(x_ref), upload(u_ref)
    if xd.dim() == 3:
        return Ref(Traj.from_batch_major(xd), Traj.from_batch_major(ud))
    return Ref(xd, ud)


# ------------------------------------------------------------------------------------------
# layout helpers
# ------------------------------------------------------------------------------------------
def pack_soa(a):
    """batch-major CUDA tensor (B, T, C) -> Traj;  (B, C) -> plain (C, B)."""
    a = a.contiguous()
    if a.dim() == 3:
        return Traj.from_batch_major(a)
    Bn, Cn = a.shape
    out = _empty(Cn, Bn)
    call("acro_transpose", Bn, Cn, _p(a), _p(out), _stream())
    return out


def unpack_soa(a):
    """Traj -> (B, T, C);  plain (C, B) -> (B, C)."""
    if isinstance(a, Traj):
        return a.batch_major()
    a = a.contiguous()
    Cn, Bn = a.shape
    out = _empty(Bn, Cn)
    call("acro_transpose", Cn, Bn, _p(a), _p(out), _stream())
    return out


# ------------------------------------------------------------------------------------------
# D1-D3
# ------------------------------------------------------------------------------------------
PHYS_FIELDS = ("m1", "m2", "l1", "lc1", "l2", "lc2", "I1", "I2", "g", "f1", "f2")


def phys_params(rows, Bn=None):
    """Per-problem physical parameters for the *_pp entry points: a (B, 11) / (11, B) array-like in the order of
    PHYS_FIELDS, or a dict field -> (B,) values (missing fields from params_1) -> device tensor (11, B)."""
    if isinstance(rows, dict):
        n = Bn or max(np.size(v) for v in rows.values())
        rows = np.stack([np.broadcast_to(np.asarray(rows.get(f, getattr(DEFAULT_PARAMS, f)), dtype=np.float64), (n,))
                         for f in PHYS_FIELDS])
    if isinstance(rows, torch.Tensor):
        t = rows.to(device=device(), dtype=F64)
    else:
        t = upload(np.asarray(rows, dtype=np.float64))
    if t.shape[0] != len(PHYS_FIELDS):
        t = t.T
    if t.shape[0] != len(PHYS_FIELDS):
        raise ValueError("physical parameters: expected 11 values per problem (%s)" % ", ".join(PHYS_FIELDS))
    return t.contiguous()


def continuous_dynamics(x, u, params=DEFAULT_PARAMS, params_b=None):
    out = torch.empty_like(x)
    call("acro_continuous_dynamics_pp", C.byref(params), _p(params_b), x.shape[1], _p(x), _p(u), _p(out), _stream())
    return out


def rk4_step(x, u, params=DEFAULT_PARAMS, params_b=None):
    out = torch.empty_like(x)
    call("acro_rk4_step_pp", C.byref(params), _p(params_b), x.shape[1], _p(x), _p(u), _p(out), _stream())
    return out


def linearize(x, u, discrete=False, params=DEFAULT_PARAMS, params_b=None):
    """-> A (4,4,B), Bm (4,2,B)"""
    Bn = x.shape[1]
    A, Bm = _empty(4, 4, Bn), _empty(4, 2, Bn)
    call("acro_linearize_pp", C.byref(params), _p(params_b), Bn, _p(x), _p(u), _p(A), _p(Bm), int(discrete), _stream())
    return A, Bm


def equilibrium(u_target, theta_guess, params=DEFAULT_PARAMS, params_b=None, tol=1e-13, max_iter=50):
    """compute_equilibrium (tg:22-39) for a batch: u_target (2,B), theta_guess (2,B) -> theta (2,B), n_iter (B,) int32
    (negative: no convergence)."""
    Bn = u_target.shape[1]
    theta, n = _empty(2, Bn), _empty(Bn, dtype=torch.int32)
    call("acro_equilibrium", C.byref(params), _p(params_b), Bn, _p(u_target), _p(theta_guess), float(tol), int(max_iter),
         _p(theta), _p(n, torch.int32), _stream())
    return theta, n


# ------------------------------------------------------------------------------------------
# G1-G11
# ------------------------------------------------------------------------------------------
def rollout_open_loop(x0, U=None, N=None, params=DEFAULT_PARAMS, params_b=None):
    """x0 (4,B), U Traj (C=2) or None (zeros, N given) -> X Traj"""
    Bn = x0.shape[1]
    N = U.T + 1 if U is not None else N
    X = Traj.empty(N, 4, Bn)
    call("acro_rollout_open_loop_pp", C.byref(params), _p(params_b), Bn, N, _p(x0), _p(U), _p(X), _stream())
    return X


def total_cost(X, U, ref, w):
    cost = _empty(X.B)
    call("acro_total_cost", w.ref(), X.B, X.T, _p(X), _p(U), ref.ref(), _p(cost), _stream())
    return cost


def costate(X, U, ref, w, params=DEFAULT_PARAMS):
    lam = Traj.empty(X.T, 4, X.B)
    call("acro_costate", C.byref(params), w.ref(), X.B, X.T, _p(X), _p(U), ref.ref(), _p(lam), _stream())
    return lam


def riccati_affine(X, U, ref, w, params=DEFAULT_PARAMS):
    """-> K Traj (C=8), S Traj (C=2), delta_J (B,), sigma_norm (B,)"""
    N, Bn = X.T, X.B
    K, S, dJ, sn = Traj.empty(N - 1, 8, Bn), Traj.empty(N - 1, 2, Bn), _empty(Bn), _empty(Bn)
    call("acro_riccati_affine", C.byref(params), w.ref(), Bn, N, _p(X), _p(U), ref.ref(), _p(K), _p(S), _p(dJ), _p(sn),
         _stream())
    return K, S, dJ, sn


def closed_loop_rollout_cost(X, U, K, S, ref, w, gammas, store=False, params=DEFAULT_PARAMS):
    """gammas (G,) shared or (G,B) per problem -> cost (G,B) [, list of G Traj Xn, list of G Traj Un]"""
    N, Bn = X.T, X.B
    G = gammas.shape[0]
    cost = _empty(G, Bn)
    Xn = _empty(G, ntiles(Bn), N, 4, 32) if store else None
    Un = _empty(G, ntiles(Bn), N - 1, 2, 32) if store else None
    call("acro_closed_loop_rollout_cost", C.byref(params), w.ref(), Bn, N, _p(X), _p(U), _p(K), _p(S), ref.ref(), G,
         _p(gammas), int(gammas.dim() == 2), _p(Xn), _p(Un), _p(cost), _stream())
    if store:
        return cost, [Traj(Xn[g], Bn) for g in range(G)], [Traj(Un[g], Bn) for g in range(G)]
    return cost


def armijo_select(cost_k, delta_J, gammas, cost_cand, c=0.5):
    G, Bn = cost_cand.shape
    acc = _empty(Bn, dtype=torch.int32)
    call("acro_armijo_select", Bn, G, _p(cost_k), _p(delta_J), _p(gammas), int(gammas.dim() == 2), _p(cost_cand), float(c),
         _p(acc, torch.int32), _stream())
    return acc


@dataclass
class NewtonState:
    """Everything acro_newton_solve reads and writes; keep it to resume a solve."""
    X: Traj
    U: Traj
    K: Traj
    S: Traj
    cost: torch.Tensor
    delta_J: torch.Tensor
    sigma_norm: torch.Tensor
    gamma_acc: torch.Tensor
    iters: torch.Tensor
    status: torch.Tensor
    Xw: Traj
    Uw: Traj
    lin: Traj
    hist_cost: Optional[torch.Tensor] = None
    hist_sigma_norm: Optional[torch.Tensor] = None
    hist_gamma: Optional[torch.Tensor] = None
    hist_ntry: Optional[torch.Tensor] = None
    initialised: bool = False
    spec_ws: Optional[torch.Tensor] = None  # candidate trajectories of the speculative kernel (allocated on first use)

    def reset(self):
        """Make the state reusable for a new solve from x0 (same shapes): history back to NaN / 0, init on next call."""
        self.initialised = False
        if self.hist_cost is not None:
            for t in (self.hist_cost, self.hist_sigma_norm, self.hist_gamma):
                t.fill_(float("nan"))
            self.hist_ntry.zero_()


def newton_alloc(Bn, N, max_iters, history=True):
    st = NewtonState(
        X=Traj.empty(N, 4, Bn), U=Traj.empty(N - 1, 2, Bn), K=Traj.empty(N - 1, 8, Bn), S=Traj.empty(N - 1, 2, Bn),
        cost=_empty(Bn), delta_J=_empty(Bn), sigma_norm=_empty(Bn), gamma_acc=_empty(Bn),
        iters=_empty(Bn, dtype=torch.int32), status=_empty(Bn, dtype=torch.int32), Xw=Traj.empty(N, 4, Bn),
        Uw=Traj.empty(N - 1, 2, Bn), lin=Traj.empty(N - 1, 10, Bn))
    if Bn % 32:
        # the padding lanes of the last tile compute in lockstep and never store: give them defined operands
        for t in (st.X, st.U, st.K, st.S, st.Xw, st.Uw, st.lin):
            t.data[-1].zero_()
    if history:
        st.hist_cost = torch.full((max_iters + 1, Bn), float("nan"), dtype=F64, device=device())
        st.hist_sigma_norm = torch.full((max_iters, Bn), float("nan"), dtype=F64, device=device())
        st.hist_gamma = torch.full((max_iters, Bn), float("nan"), dtype=F64, device=device())
        st.hist_ntry = torch.zeros((max_iters, Bn), dtype=torch.int32, device=device())
    return st


# Named kernel variants of acro_newton_solve (AcroNewtonOpts.kernel / stage_steps / recompute_lin): what the automatic
# dispatch picks from, by batch size.  `kernel=` of newton_solve takes one of these names (or "auto").
KERNEL_VARIANTS = {
    "auto": dict(kernel="auto"),
    "duo": dict(kernel="duo"),                       # <= 148 tiles: two warps per tile, 16-step stages (8 with per-problem references)
    "duo4": dict(kernel="duo", stage_steps=4),       # 149-296 tiles: two blocks per SM
    "duo8": dict(kernel="duo", stage_steps=8),
    "ring": dict(kernel="ring"),                     # one warp per tile
    "ring16": dict(kernel="ring", stage_steps=16),
    "ring4": dict(kernel="ring", stage_steps=4),     # 297-592 tiles
    "ring2": dict(kernel="ring", stage_steps=2, recompute_lin=2),
    "ring-rl": dict(kernel="ring", stage_steps=2, recompute_lin=1),  # > 592 tiles: linearisation recomputed (304 B/step-iteration)
    "thread": dict(kernel="thread"),
    "ldg": dict(kernel="thread"),
    "spec": dict(kernel="spec"),                     # <= 148 tiles: duo + speculative parallel Armijo candidates
    "spec1": dict(kernel="spec", speculate=1),       # one candidate per round: the duo path inside the speculative kernel
    "spec3": dict(kernel="spec", speculate=3),       # fixed number of candidates per round
    "spec8": dict(kernel="spec", speculate=8),
}
# Process-wide default of newton_solve's kernel selection (a Python-side setting: the library reads no environment).
NEWTON_DEFAULTS = {"kernel": "auto"}
SPEC_AUTO_GAMMA = 0.5  # acro_kernels.cu: ACRO_SPEC_AUTO_GAMMA


def sm_count():
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count if torch.cuda.is_available() else 148


def newton_opts(max_iters, tol=1e-6, beta=0.7, c=0.5, gamma_0=1.0, chunk_iters=0, max_line_search=20, init=1,
                kernel=None, stage_steps=0, recompute_lin=0, speculate=0):
    if kernel is None:
        kernel = NEWTON_DEFAULTS["kernel"]
    if isinstance(kernel, str):
        v = KERNEL_VARIANTS[kernel]
        stage_steps = stage_steps or v.get("stage_steps", 0)
        recompute_lin = recompute_lin or v.get("recompute_lin", 0)
        speculate = speculate or v.get("speculate", 0)
        kernel = _abi.NEWTON_KERNELS[v["kernel"]]
    return AcroNewtonOpts(max_iters=int(max_iters), chunk_iters=int(chunk_iters), max_line_search=int(max_line_search),
                          init=int(init), tol=float(tol), beta=float(beta), c=float(c), gamma_0=float(gamma_0),
                          kernel=int(kernel), stage_steps=int(stage_steps), recompute_lin=int(recompute_lin),
                          speculate=int(speculate))


def newton_kernel_name(Bn, ref_per_problem=False, weights_per_problem=False, params_per_problem=False, **kw):
    """Name of the kernel newton_solve launches for a batch of Bn problems with these options (acro_newton_describe)."""
    kw.setdefault("gamma_0", 0.1)
    o = newton_opts(kw.pop("max_iters", 1), **kw)
    if o.kernel == 0 and o.gamma_0 >= SPEC_AUTO_GAMMA:
        o.spec_ws = 128  # (what newton_solve does: it provides the workspace when the automatic choice may need it)
    buf = C.create_string_buffer(128)
    call("acro_newton_describe", C.byref(o), int(Bn), int(ref_per_problem), int(weights_per_problem),
         int(params_per_problem), buf, 128)
    return buf.value.decode()


def newton_solve(x0, ref, max_iters, tol=1e-6, beta=0.7, c=0.5, gamma_0=1.0, w=None, params=DEFAULT_PARAMS,
                 state=None, chunk_iters=0, max_line_search=20, history=True, warm_start_U=None, params_b=None,
                 kernel=None, stage_steps=0, recompute_lin=0, speculate=0):
    """newton_Algorithm (trajectory_generation.py:298-398) for a batch, x0 (4,B).

    One kernel launch runs the whole loop for every problem.  Pass the returned state back
    (with chunk_iters) to continue a solve in pieces.  warm_start_U (Traj, C=2) replaces the reference's
    u = 0 initial guess (tg:311).  params_b (11,B): physical parameters per problem (phys_params).
    kernel / stage_steps / recompute_lin / speculate: AcroNewtonOpts kernel selection ("auto", "duo", "ring",
    "thread", "spec"); the default picks by batch size."""
    w = newton_weights() if w is None else w
    Bn = x0.shape[1]
    N = ref.N
    if state is None:
        state = newton_alloc(Bn, N, max_iters, history)
    init = 0 if state.initialised else 1
    if init and warm_start_U is not None:
        state.U.data.copy_(warm_start_U.data)
        init = 2
    o = newton_opts(max_iters, tol, beta, c, gamma_0, chunk_iters, max_line_search, init, kernel, stage_steps,
                    recompute_lin, speculate)
    s = state
    # the speculative kernel needs room for 8 candidate trajectories per problem: allocate it when that kernel is asked
    # for, or when the automatic choice may fall on it (one tile per SM, large initial step: see acro_abi.h)
    if o.kernel == _abi.NEWTON_KERNELS["spec"] or (o.kernel == 0 and float(gamma_0) >= SPEC_AUTO_GAMMA and params_b is None
                                                   and ntiles(Bn) <= sm_count() and not params.actuated_tau1):
        if s.spec_ws is None:
            s.spec_ws = torch.zeros(int(_abi.lib.acro_newton_spec_ws_doubles(Bn, N)), dtype=F64, device=device())
        o.spec_ws = s.spec_ws.data_ptr()
    call("acro_newton_solve_pp", C.byref(params), _p(params_b), w.ref(), C.byref(o), Bn, N, _p(x0), ref.ref(), _p(s.X), _p(s.U), _p(s.Xw),
         _p(s.Uw), _p(s.lin), _p(s.K), _p(s.S), _p(s.cost), _p(s.delta_J), _p(s.sigma_norm), _p(s.gamma_acc),
         _p(s.iters, torch.int32), _p(s.status, torch.int32), _p(s.hist_cost), _p(s.hist_sigma_norm), _p(s.hist_gamma),
         _p(s.hist_ntry, torch.int32), _stream())
    s.initialised = True
    return s


def stepsize_sweep(X, U, K, S, ref, w, steps, params=DEFAULT_PARAMS):
    """P base iterates x len(steps) step sizes -> cost (S_n, P)"""
    Sn = steps.shape[0]
    cost = _empty(Sn, X.B)
    call("acro_stepsize_sweep", C.byref(params), w.ref(), X.B, X.T, _p(X), _p(U), _p(K), _p(S), ref.ref(), Sn, _p(steps),
         _p(cost), _stream())
    return cost


# ------------------------------------------------------------------------------------------
# T1-T5
# ------------------------------------------------------------------------------------------
def lqr_gains(traj, w=None, params=DEFAULT_PARAMS):
    """traj shared -> K tensor (N-1, 8) [= (N-1,2,4) row-major]; per problem -> K Traj (C=8)"""
    w = lqr_weights() if w is None else w
    N = traj.N
    if traj.per_problem:
        Bn = traj.B
        K = Traj.empty(N - 1, 8, Bn)
    else:
        Bn = 1
        K = _empty(N - 1, 8)
    call("acro_lqr_gains", C.byref(params), w.ref(), Bn, N, traj.ref(), _p(K), _stream())
    return K


def lqr_track(traj, K, x0, params=DEFAULT_PARAMS, params_b=None):
    """-> Xt Traj (C=4), Ut Traj (C=2).  params_b (11,B): every plant its own physical parameters (phys_params)."""
    N, Bn = traj.N, x0.shape[1]
    Xt, Ut = Traj.empty(N, 4, Bn), Traj.empty(N - 1, 2, Bn)
    call("acro_lqr_track_pp", C.byref(params), _p(params_b), Bn, N, traj.ref(), _p(K), _p(x0), _p(Xt), _p(Ut), _stream())
    return Xt, Ut


def p_inf(A, Bm, w, max_iter=1000, tol=1e-6):
    """A (4,4,B), Bm (4,2,B) -> P (4,4,B), n_iter (B,) int32 (negative: not converged)"""
    Bn = A.shape[-1]
    P, n = _empty(4, 4, Bn), _empty(Bn, dtype=torch.int32)
    call("acro_p_inf", w.ref(), Bn, _p(A), _p(Bm), int(max_iter), float(tol), _p(P), _p(n, torch.int32), _stream())
    return P, n


def mpc_solve(x0, A_w, B_w, QT, w, T_pred, trajectories=True):
    """x0 (4,B), A_w Traj (T_pred-1, C=16), B_w Traj (T_pred-1, C=8), QT (4,4,B)
    -> U0 (2,B), X_opt Traj (T_pred, 4), U_opt Traj (T_pred, 2), gains Traj (T_pred-1, 8)"""
    Bn = x0.shape[1]
    U0 = _empty(2, Bn)
    Xo = Traj.empty(T_pred, 4, Bn) if trajectories else None
    Uo = Traj.empty(T_pred, 2, Bn) if trajectories else None
    Kws = Traj.empty(max(T_pred - 1, 1), 8, Bn) if trajectories else None
    call("acro_mpc_solve", w.ref(), Bn, int(T_pred), _p(x0), _p(A_w), _p(B_w), _p(QT), _p(U0), _p(Xo), _p(Uo), _p(Kws),
         _stream())
    return U0, Xo, Uo, Kws


def mpc_track(x0, ref, QT_inf, T=None, T_pred=75, w=None, x_f=(np.pi, 0.0, 0.0, 0.0), u_f=(0.0, 0.0),
              params=DEFAULT_PARAMS, params_b=None):
    """solve_mpc_tracking for a batch.  QT_inf (4,4) shared or (4,4,B).  -> Xr Traj, Ur Traj, K0 or None, n_solves
    params_b (11,B): every problem its own physical parameters (phys_params; needs a per-problem reference)."""
    w = mpc_weights() if w is None else w
    N, Bn = ref.N, x0.shape[1]
    T = N if T is None else T
    qt_pp = QT_inf.dim() == 3
    pp = ref.per_problem or w.per_problem or qt_pp
    Xr, Ur = Traj.empty(T, 4, Bn), Traj.empty(T - 1, 2, Bn)
    xf = (C.c_double * 4)(*[float(v) for v in x_f])
    uf = (C.c_double * 2)(*[float(v) for v in u_f])
    ns = C.c_int64(0)
    if params_b is not None:
        if not ref.per_problem:
            raise ValueError("mpc_track: per-problem physical parameters need a per-problem reference (B, N, 4)")
        lin = Traj.empty(N - 1, 10, Bn)
        call("acro_mpc_track_pp", C.byref(params), _p(params_b), w.ref(), Bn, N, int(T), int(T_pred), ref.ref(), xf, uf,
             _p(QT_inf), int(qt_pp), _p(x0), _p(lin), _p(Xr), _p(Ur), C.byref(ns), _stream())
        return Xr, Ur, None, int(ns.value)
    K0 = None if pp else _empty(T - 1, 8)
    lin = Traj.empty(N - 1, 10, Bn) if pp else _empty(N - 1, 10)
    call("acro_mpc_track", C.byref(params), w.ref(), Bn, N, int(T), int(T_pred), ref.ref(), xf, uf, _p(QT_inf), int(qt_pp),
         _p(x0), _p(K0), _p(lin), _p(Xr), _p(Ur), C.byref(ns), _stream())
    return Xr, Ur, K0, int(ns.value)


def mpc_track_box(x0, ref, QT_inf, tau_max=18.0, T=None, T_pred=75, w=None, x_f=(np.pi, 0.0, 0.0, 0.0), u_f=(0.0, 0.0),
                  params=DEFAULT_PARAMS, max_iter=0, params_b=None):
    """solve_mpc_tracking with the input box -tau_max <= u + u_ref <= tau_max of tt:87-91, 112-114 switched on; every
    step is solved exactly (active set on Riccati sweeps).  -> Xr Traj, Ur Traj, info {n_sweeps (B,), n_active (T-1,B),
    status (B,)}.  params_b (11,B): every problem its own physical parameters (needs a per-problem reference)."""
    from ._abi import lib
    w = mpc_weights() if w is None else w
    N, Bn = ref.N, x0.shape[1]
    T = N if T is None else T
    qt_pp = QT_inf.dim() == 3
    Xr, Ur = Traj.empty(T, 4, Bn), Traj.empty(T - 1, 2, Bn)
    lin = Traj.empty(N - 1, 10, Bn) if ref.per_problem else _empty(N - 1, 10)
    ws = _empty(int(lib.acro_mpc_box_ws_doubles(Bn, int(T), int(T_pred))))
    ns, st = _empty(Bn, dtype=torch.int32), _empty(Bn, dtype=torch.int32)
    na = _empty(T - 1, Bn, dtype=torch.int32)
    xf = (C.c_double * 4)(*[float(v) for v in x_f])
    uf = (C.c_double * 2)(*[float(v) for v in u_f])
    if params_b is not None:
        if not ref.per_problem:
            raise ValueError("mpc_track_box: per-problem physical parameters need a per-problem reference (B, N, 4)")
        call("acro_mpc_track_box_pp", C.byref(params), _p(params_b), w.ref(), Bn, N, int(T), int(T_pred), ref.ref(), xf, uf,
             _p(QT_inf), int(qt_pp), _p(x0), float(tau_max), int(max_iter), _p(lin), _p(ws), _p(Xr), _p(Ur),
             _p(ns, torch.int32), _p(na, torch.int32), _p(st, torch.int32), _stream())
        return Xr, Ur, {"n_sweeps": ns, "n_active": na, "status": st}
    call("acro_mpc_track_box", C.byref(params), w.ref(), Bn, N, int(T), int(T_pred), ref.ref(), xf, uf, _p(QT_inf), int(qt_pp),
         _p(x0), float(tau_max), int(max_iter), _p(lin), _p(ws), _p(Xr), _p(Ur), _p(ns, torch.int32), _p(na, torch.int32),
         _p(st, torch.int32), _stream())
    return Xr, Ur, {"n_sweeps": ns, "n_active": na, "status": st}


# ------------------------------------------------------------------------------------------
# stand-alone pieces of the Newton iteration (for the drop-in functions that expose them)
# ------------------------------------------------------------------------------------------
def cost_derivatives(x, x_ref, u, u_ref, w, terminal=False):
    """Batch of points (4,B)/(2,B) -> l (B,), grad_x (4,B), grad_u (2,B) or None   (tg:89-114)"""
    Bn = x.shape[1]
    l, gx = _empty(Bn), _empty(4, Bn)
    gu = None if terminal else _empty(2, Bn)
    call("acro_cost_derivatives", w.ref(), Bn, _p(x), _p(x_ref), _p(u), _p(u_ref), int(terminal), _p(l), _p(gx), _p(gu),
         _stream())
    return l, gx, gu


def discretize(Ac, Bc, dt=DT):
    """(4,4,B), (4,2,B) -> Ad, Bd   (tg:161-164)"""
    Ad, Bd = torch.empty_like(Ac), torch.empty_like(Bc)
    call("acro_discretize", Ac.shape[-1], _p(Ac), _p(Bc), float(dt), _p(Ad), _p(Bd), _stream())
    return Ad, Bd


def stage_lists(X, U, ref, w, params=DEFAULT_PARAMS):
    """-> A Traj (C=16), Bm Traj (C=8), q Traj (C=4), r Traj (C=2), q_T (4,B)   (tg:166-181)"""
    N, Bn = X.T, X.B
    A, Bm, q, r = Traj.empty(N - 1, 16, Bn), Traj.empty(N - 1, 8, Bn), Traj.empty(N - 1, 4, Bn), Traj.empty(N - 1, 2, Bn)
    qT = _empty(4, Bn)
    call("acro_stage_lists", C.byref(params), w.ref(), Bn, N, _p(X), _p(U), ref.ref(), _p(A), _p(Bm), _p(q), _p(r), _p(qT),
         _stream())
    return A, Bm, q, r, qT


def riccati_lists(A, Bm, Q, R, q, r, Q_T, q_T, S_cross=None):
    """Dense lists as Traj (C = 16, 8, 16, 4, 4, 2), Q_T (16,B), q_T (4,B) -> K Traj, S Traj, delta_J   (tg:183-216)"""
    T, Bn = A.T, A.B
    K, S, dJ = Traj.empty(T, 8, Bn), Traj.empty(T, 2, Bn), _empty(Bn)
    call("acro_riccati_lists", Bn, T, _p(A), _p(Bm), _p(Q), _p(R), _p(S_cross), _p(q), _p(r), _p(Q_T), _p(q_T), _p(K),
         _p(S), _p(dJ), _stream())
    return K, S, dJ
