"""Drop-in for the reference's ``trajectory_tracking.py`` (LQR and receding-horizon MPC tracking).

The MPC quadratic programme of ``solver_mpc`` (trajectory_tracking.py:73-140) has only equality
constraints (``test_constraints = False``, tt:87), so it is an LQ problem: the kernels solve it exactly by a
(T_pred-1)-step Riccati sweep instead of CasADi/IPOPT.  A solver failure cannot occur; where the reference
would ``exit()`` (tt:131-133) this module raises.
"""
import numpy as np
import torch

from . import _io
from . import batched as bt
from .trajectory_generation import *  # noqa: F401,F403  (tt:3)
from .trajectory_generation import _ref, _weights, active_params, dt, ns, nu, nx

Q_reg = np.diag([100.0, 100.0, 10.0, 10.0])  # tt:173
R_Reg = np.diag([1.0, 1.0])                  # tt:174
Q_mpc = np.diag([120.0, 100.0, 0.0001, 0.0001])  # tt:38
R_mpc = np.diag([1e-6, 10.0])                    # tt:39
x_f = np.array([np.pi, 0, 0, 0])  # tt:33
u_f = np.array([0, 0])            # tt:34


def solve_LQR_tracking(x_opt, u_opt):
    """trajectory_tracking.py:170-203.  One trajectory -> list of N-1 gains (2,4); batch (B,N,4) -> (B,N-1,2,4)."""
    traj = _ref(x_opt, u_opt)
    K = bt.lqr_gains(traj, bt.Weights(Q_reg, R_Reg), active_params())
    if traj.per_problem:
        kind = _io.Kind(x_opt, True)
        return _io.out(K, kind, tail=(2, 4), key="Kreg")
    Kh = K.cpu().numpy().reshape(-1, 2, 4)
    return list(Kh)


def simulate_tracking(x_opt, u_opt, K_reg, x0_perturbed, *, params_b=None):
    """trajectory_tracking.py:206-216.  x0_perturbed (4,) or (B,4); trajectory and gains shared or per problem.
    params_b (B, 11): every plant its own physical parameters (tracking under model mismatch)."""
    x0, kind = _io.state_in(x0_perturbed, nx)
    traj = _ref(x_opt, u_opt)
    if traj.per_problem:
        Kd, _ = _io.traj_in(K_reg, 8)
    else:
        Kd = bt.upload(np.asarray([np.asarray(k, dtype=np.float64).reshape(8) for k in K_reg])
                       if isinstance(K_reg, (list, tuple)) else K_reg).reshape(-1, 8).contiguous()
    pb = None if params_b is None else bt.phys_params(params_b, x0.shape[1])
    Xt, Ut = bt.lqr_track(traj, Kd, x0, active_params(), pb)
    return _io.out(Xt, kind, key="xt"), _io.out(Ut, kind, key="ut")


def LQR_tracking(x_ref, u_ref, t_ref, x0_perturbed=None, *, params_b=None, block=True):
    """trajectory_tracking.py:219-249.  params_b: the gains are those of the active (nominal) model, the plants differ.
    block=False: returns a handle (`.result()` -> the tuple); uploads and copies back run on side streams."""
    if x0_perturbed is None:
        x0_perturbed = x_ref[0].copy() if not isinstance(x_ref, torch.Tensor) else x_ref[0].clone()
    with _io.upload_scope(not block, x_ref, u_ref, x0_perturbed, params_b) as up:
        traj = _ref(x_ref, u_ref)
        x0, kind = _io.state_in(x0_perturbed, nx)
        pb = None if params_b is None else bt.phys_params(params_b, x0.shape[1])
        up.keep(traj, x0, pb)
    if not block:
        _io.flush_deferred()
    K = bt.lqr_gains(traj, bt.Weights(Q_reg, R_Reg), active_params())
    Xt, Ut = bt.lqr_track(traj, K, x0, active_params(), pb)
    if not block:
        pend = _io.out_async([(Xt, None), (Ut, None)], kind, defer=True)
        return _io.Pending(pend._event, lambda: tuple(pend.result()))
    return _io.out(Xt, kind, key="xt"), _io.out(Ut, kind, key="ut")


def compute_P_inf(A, B, Q, R):
    """trajectory_tracking.py:144-165"""
    Ad = bt.upload(np.asarray(A, dtype=np.float64).reshape(4, 4, 1))
    Bd = bt.upload(np.asarray(B, dtype=np.float64).reshape(4, 2, 1))
    P, n = bt.p_inf(Ad, Bd, bt.Weights(Q, R))
    if int(n[0]) < 0:
        print("P_inf did not converge!!!")
    return P.cpu().numpy()[:, :, 0]


def solver_mpc(x0, A_list, B_list, Q, R, Q_T, T_pred, u_ref=None):
    """trajectory_tracking.py:73-140 -> (U0 (2,), X_opt (T_pred,4), U_opt (T_pred,2)).  u_ref is unused, as in
    the reference when test_constraints is False (tt:87)."""
    H = int(T_pred)
    A = np.asarray([np.asarray(a, dtype=np.float64) for a in A_list[:H - 1]]).reshape(1, H - 1, 16)
    Bm = np.asarray([np.asarray(b, dtype=np.float64) for b in B_list[:H - 1]]).reshape(1, H - 1, 8)
    x0d, kind = _io.state_in(np.asarray(x0, dtype=np.float64).reshape(4), nx)
    U0, Xo, Uo, _ = bt.mpc_solve(x0d, bt.Traj.from_batch_major(bt.upload(A)), bt.Traj.from_batch_major(bt.upload(Bm)),
                                 bt.upload(np.asarray(Q_T, dtype=np.float64).reshape(4, 4, 1)), bt.Weights(Q, R), H)
    return U0.cpu().numpy()[:, 0], Xo.batch_major()[0].cpu().numpy(), Uo.batch_major()[0].cpu().numpy()


def _pad_time(a, n):
    """Zero-pad the time axis (second to last) to n rows: the reference allocates x_real / u_real with the shapes of
    x_ref / u_ref (tt:27-28) and fills the first T / T-1 rows."""
    t = a.shape[-2]
    if t >= n:
        return a
    if isinstance(a, torch.Tensor):
        out = a.new_zeros(*a.shape[:-2], n, a.shape[-1])
    else:
        out = np.zeros(a.shape[:-2] + (n, a.shape[-1]))
    out[..., :t, :] = a
    return out


_p_inf_cache = {}


def _terminal_weight(p, w):
    """P_inf of the linearisation about the final equilibrium (tt:33-40), computed on the device once per
    (parameters, weights) and kept there: the convergence check reads a flag back, which must not sit in the way of
    back-to-back non-blocking calls."""
    key = (bytes(p), torch.cuda.current_device(),
           None if w.per_problem else (w.Q.tobytes(), w.R.tobytes(), np.asarray(x_f, dtype=np.float64).tobytes(),
                                       np.asarray(u_f, dtype=np.float64).tobytes()))
    hit = _p_inf_cache.get(key) if key[2] is not None else None
    if hit is not None:
        return hit
    xf = bt.upload(np.asarray(x_f, dtype=np.float64).reshape(4, 1))
    uf = bt.upload(np.asarray(u_f, dtype=np.float64).reshape(2, 1))
    A_f, B_f = bt.linearize(xf, uf, True, p)
    P, n = bt.p_inf(A_f, B_f, w)
    if int(n[0]) < 0:
        print("P_inf did not converge!!!")
    P = P[:, :, 0].contiguous()
    if key[2] is not None:
        if len(_p_inf_cache) > 64:
            _p_inf_cache.clear()
        _p_inf_cache[key] = P
    return P


def solve_mpc_tracking(x0, x_ref, u_ref, T, *, T_pred=75, Q=None, R=None, return_info=False, tau_max=None, block=True,
                       params_b=None):
    """trajectory_tracking.py:8-69.  x0 (4,) or (B,4); reference shared (N,4) or per problem (B,N,4).

    tau_max: switches on the input box the reference keeps behind `test_constraints` (tt:87-91, 112-114; 18 there):
    -tau_max <= u + u_ref <= tau_max at every step of every horizon, each QP solved exactly.
    block=False (without return_info): returns a handle at once; `.result()` gives `(x_real, u_real)`.  The uploads of
    this call run on an upload stream and the copies back on a copy stream, so with several calls in flight the
    transfers of one batch overlap the kernel of another (pinned host inputs must stay untouched until then).
    params_b (B, 11): every problem its own physical parameters (domain randomisation): each linearises the reference
    with its own model, has its own terminal weight P_inf and steps its own plant (also with tau_max)."""
    w = bt.Weights(Q_mpc if Q is None else Q, R_mpc if R is None else R)
    p = active_params()
    piped = (not block) and not return_info
    with _io.upload_scope(piped, x0, x_ref, u_ref) as up:
        x0d, kind = _io.state_in(x0, nx)
        ref = _ref(x_ref, u_ref)
        up.keep(x0d, ref)
    if piped:
        _io.flush_deferred()  # copy-back of the previous call: queued after this call's uploads, before its kernel
    if params_b is not None:
        Bn = x0d.shape[1]
        pb = bt.phys_params(params_b, Bn)
        if not ref.per_problem:  # every problem linearises the reference with its own model: per-problem layout
            ref = bt.Ref(bt.Traj.from_batch_major(ref.x.unsqueeze(0).expand(Bn, -1, -1).contiguous()),
                         bt.Traj.from_batch_major(ref.u.unsqueeze(0).expand(Bn, -1, -1).contiguous()))
        xf = bt.upload(np.repeat(np.asarray(x_f, dtype=np.float64).reshape(4, 1), Bn, 1))
        uf = bt.upload(np.repeat(np.asarray(u_f, dtype=np.float64).reshape(2, 1), Bn, 1))
        A_f, B_f = bt.linearize(xf, uf, True, p, pb)
        Pb, n = bt.p_inf(A_f, B_f, w)
        if not piped and int(n.min()) < 0:
            print("P_inf did not converge!!!")
        if tau_max is not None:
            Xr, Ur, info = bt.mpc_track_box(x0d, ref, Pb, tau_max=float(tau_max), T=int(T), T_pred=int(T_pred), w=w, x_f=x_f,
                                            u_f=u_f, params=p, params_b=pb)
            if piped:
                return _MpcPending(_io.out_async([(Xr, ref.N), (Ur, ref.N - 1)], kind, defer=True), info["status"])
            if int(info["status"].max()) != 0:
                print("Attention! mpc solver: active-set iteration limit reached for %d problem(s)" % int((info["status"] != 0).sum()))
            xr, ur = _pad_time(_io.out(Xr, kind, key="xr"), ref.N), _pad_time(_io.out(Ur, kind, key="ur"), ref.N - 1)
            if return_info:
                return xr, ur, dict(n_solves=(int(T) - 1) * Bn, n_sweeps=info["n_sweeps"].cpu().numpy(),
                                    n_active=info["n_active"].cpu().numpy().T, status=info["status"].cpu().numpy(),
                                    P_inf=Pb.cpu().numpy())
            return xr, ur
        Xr, Ur, _, n_solves = bt.mpc_track(x0d, ref, Pb, T=int(T), T_pred=int(T_pred), w=w, x_f=x_f, u_f=u_f, params=p,
                                           params_b=pb)
        if piped:
            return _MpcPending(_io.out_async([(Xr, ref.N), (Ur, ref.N - 1)], kind, defer=True), None)
        xr, ur = _pad_time(_io.out(Xr, kind, key="xr"), ref.N), _pad_time(_io.out(Ur, kind, key="ur"), ref.N - 1)
        if return_info:
            return xr, ur, dict(n_solves=n_solves, K0=None, P_inf=Pb.cpu().numpy())
        return xr, ur
    P = _terminal_weight(p, w)
    if tau_max is not None:
        Xr, Ur, info = bt.mpc_track_box(x0d, ref, P, tau_max=float(tau_max), T=int(T),
                                        T_pred=int(T_pred), w=w, x_f=x_f, u_f=u_f, params=p)
        if piped:
            return _MpcPending(_io.out_async([(Xr, ref.N), (Ur, ref.N - 1)], kind, defer=True), info["status"])
        if int(info["status"].max()) != 0:
            print("Attention! mpc solver: active-set iteration limit reached for %d problem(s)" % int((info["status"] != 0).sum()))
        xr, ur = _pad_time(_io.out(Xr, kind, key="xr"), ref.N), _pad_time(_io.out(Ur, kind, key="ur"), ref.N - 1)
        if return_info:
            return xr, ur, dict(n_solves=(int(T) - 1) * x0d.shape[1], n_sweeps=info["n_sweeps"].cpu().numpy(),
                                n_active=info["n_active"].cpu().numpy().T, status=info["status"].cpu().numpy(),
                                P_inf=P.cpu().numpy())
        return xr, ur
    Xr, Ur, K0, n_solves = bt.mpc_track(x0d, ref, P, T=int(T), T_pred=int(T_pred), w=w, x_f=x_f, u_f=u_f, params=p)
    if piped:
        return _MpcPending(_io.out_async([(Xr, ref.N), (Ur, ref.N - 1)], kind, defer=True), None)
    xr, ur = _pad_time(_io.out(Xr, kind, key="xr"), ref.N), _pad_time(_io.out(Ur, kind, key="ur"), ref.N - 1)
    if return_info:
        return xr, ur, dict(n_solves=n_solves, K0=None if K0 is None else K0.cpu().numpy().reshape(-1, 2, 4),
                            P_inf=P.cpu().numpy())
    return xr, ur


class _MpcPending:
    """Handle of a non-blocking solve_mpc_tracking: result() -> (x_real, u_real)."""

    def __init__(self, pend, status):
        self._pend, self._status = pend, status

    def ready(self):
        return self._pend.ready()

    def result(self):
        xr, ur = self._pend.result()
        if self._status is not None:
            bad = int((self._status != 0).sum())
            self._status = None
            if bad:
                print("Attention! mpc solver: active-set iteration limit reached for %d problem(s)" % bad)
        return xr, ur
