"""Drop-in for the reference's ``trajectory_generation.py`` (Newton / Riccati / Armijo).

Same function names, argument order and defaults as the reference.  Every function accepts what the
reference accepts (one problem, NumPy) and also a leading batch axis and torch tensors; all arithmetic
runs in libacro_b200.so.  Module-level ``Q, R, Q_T`` (trajectory_generation.py:16-18) are read at call time,
so rebinding them changes the weights as it does in the reference; the batched entry points also take
``Q=, R=, Q_T=`` keyword arguments (arrays of shape (4,4) or per problem (B,4,4)).
"""
import numpy as np
import torch

from . import _io
from . import batched as bt
from .dynamics import *  # noqa: F401,F403  (the reference re-exports dynamics the same way, tg:3)
from .dynamics import active_params, dt, ni, ns

T = 10.0
N = int(T / dt) + 1
nu = 2
nx = 4

Q = np.diag([130.0, 30.0, 0.0001, 0.0001])
R = np.diag([1e-6, 1.5])
Q_T = np.diag([130, 130.0, 1.0, 1.0])


def _weights(Qm=None, Rm=None, QTm=None):
    """Shared (4,4)/(2,2) weights or per-problem stacks (B,4,4)/(B,2,2)."""
    Qm = Q if Qm is None else Qm
    Rm = R if Rm is None else Rm
    QTm = Q_T if QTm is None else QTm
    arrs = [np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float64) for a in (Qm, Rm, QTm)]
    if all(a.ndim == 2 for a in arrs):
        return bt.Weights(*arrs)
    Bn = max(a.shape[0] for a in arrs if a.ndim == 3)
    full = [np.broadcast_to(a, (Bn,) + a.shape[-2:]) for a in arrs]
    dev = [bt.upload(np.ascontiguousarray(a.reshape(Bn, -1).T)) for a in full]
    return bt.Weights(full[0][0], full[1][0], full[2][0], Q_b=dev[0], R_b=dev[1], QT_b=dev[2])


def _ref(x_ref, u_ref):
    """Shared (N,4)/(N-1,2) or per-problem (B,N,4)/(B,N-1,2) reference."""
    return bt.make_ref(x_ref, u_ref)


def compute_equilibrium(u_target, theta_guess, version=1, tol=1e-14, max_iter=50):
    """Equilibrium of the gravity vector, G(theta1, theta2) = u_target (trajectory_generation.py:22-39).

    The reference calls SciPy's MINPACK `hybr` root finder; the system is two smooth equations in two unknowns,
        g1 sin(th1) + g2 sin(th1 + th2) = u1,      g2 sin(th1 + th2) = u2,
    so a Newton iteration from the same initial guess is used here (host side: this builds an input of the hot
    path once per problem family).  Raises RuntimeError like the reference if it does not converge."""
    if (isinstance(u_target, torch.Tensor) and u_target.dim() == 2) or (not isinstance(u_target, torch.Tensor)
                                                                         and np.ndim(u_target) == 2):
        return compute_equilibrium_batch(u_target, theta_guess, version=version, tol=max(tol, 1e-13), max_iter=max_iter)
    p = bt.PARAM_SETS[version]
    g1 = p["g"] * (p["lc1"] * p["m1"] + p["m2"] * p["l1"])
    g2 = p["g"] * p["m2"] * p["lc2"]
    u = np.asarray(u_target, dtype=float).reshape(-1)
    th = np.asarray(theta_guess, dtype=float).reshape(2).copy()
    for _ in range(max_iter):
        s1, c1, s12, c12 = np.sin(th[0]), np.cos(th[0]), np.sin(th[0] + th[1]), np.cos(th[0] + th[1])
        res = np.array([g1 * s1 + g2 * s12 - u[0], g2 * s12 - u[1]])
        if np.max(np.abs(res)) < tol:
            return np.array([th[0], th[1], 0.0, 0.0]), u.copy()
        J = np.array([[g1 * c1 + g2 * c12, g2 * c12], [g2 * c12, g2 * c12]])
        th = th - np.linalg.solve(J, res)
    raise RuntimeError("Root finder failed: Newton iteration on G(theta) = u did not converge")


def compute_equilibrium_batch(u_target, theta_guess, version=1, tol=1e-13, max_iter=50, params_b=None):
    """compute_equilibrium for B targets at once on the device (acro_equilibrium): u_target (B,2), theta_guess (B,2) or
    (2,) -> x_e (B,4), u_e (B,2), same kind of array as u_target.  Raises RuntimeError like the reference (tg:33-34) if
    any problem does not converge.  params_b (B,11): every problem its own physical parameters."""
    ut, kind = _io.state_in(u_target, nu)
    Bn = ut.shape[1]
    tg_ = theta_guess.detach().cpu().numpy() if isinstance(theta_guess, torch.Tensor) else np.asarray(theta_guess, dtype=np.float64)
    th0 = bt.upload(np.ascontiguousarray(np.broadcast_to(tg_, (Bn, 2)).T))
    pb = None if params_b is None else bt.phys_params(params_b, Bn)
    theta, n = bt.equilibrium(ut, th0, params=bt.make_params(version, dt), params_b=pb, tol=tol, max_iter=max_iter)
    if bool((n < 0).any()):
        raise RuntimeError("Root finder failed: Newton iteration on G(theta) = u did not converge for %d of %d targets"
                           % (int((n < 0).sum()), Bn))
    x_e = torch.cat([theta, torch.zeros_like(theta)], dim=0)  # [theta1, theta2, 0, 0]
    return _io.out(x_e, kind), _io.out(ut, kind)


def define_reference_piecewise(T, x_e1, x_e2, u_e1, u_e2):
    """Two constant segments, x_e1 for t < T/2 and x_e2 after (trajectory_generation.py:41-58).  Host-side
    construction of an input; not a compute kernel.  Batched: x_e1, x_e2 (B,4) and u_e1, u_e2 (B,2) give x_ref (B,N,4),
    u_ref (B,N,2)."""
    n = int(T / dt) + 1
    t_ref = np.linspace(0.0, T, n)
    first = (t_ref < T / 2.0)[:, None]
    a = [np.asarray(v.detach().cpu() if isinstance(v, torch.Tensor) else v, dtype=float) for v in (x_e1, x_e2, u_e1, u_e2)]
    x_ref = np.where(first, a[0][..., None, :], a[1][..., None, :])
    u_ref = np.where(first, a[2][..., None, :], a[3][..., None, :])
    return t_ref, x_ref, u_ref


def get_fully_actuated_ref(path="trajectories_npz/fully_actuated_trajectory.npz"):
    """Load and rescale the fully-actuated reference (trajectory_generation.py:511-518): tau1 := 0, torques x 2."""
    data = np.load(path)
    u_ref = np.zeros(data["u"].shape)
    u_ref[:, 1] = data["u"][:, 1]
    return data["x"], np.multiply(u_ref, 2), data["time"]


# ---------------------------------------------------------------------------------------------------------
def simulate_open_loop(x0, u_traj, *, params_b=None):
    """trajectory_generation.py:74-87.  params_b (B, 11): physical parameters per problem."""
    x, kind = _io.state_in(x0, nx)
    U, ku = _io.traj_in(u_traj, nu)
    if U.B != x.shape[1]:
        raise ValueError("batch sizes of x0 and u_traj differ")
    kind.batched = kind.batched or ku.batched
    pb = None if params_b is None else bt.phys_params(params_b, x.shape[1])
    return _io.out(bt.rollout_open_loop(x, U, params=active_params(), params_b=pb), kind)


def derivatives_Cost(x, x_ref, u, u_ref, Q, R, Q_T=None, terminal=False):
    """trajectory_generation.py:89-114.  One point or a batch of points."""
    xd, kind = _io.state_in(x, nx)
    xr, _ = _io.state_in(x_ref, nx)
    xr = xr.expand_as(xd).contiguous()
    if terminal:
        w = _weights(np.zeros((4, 4)), np.zeros((2, 2)), Q_T)
        l, gx, _ = bt.cost_derivatives(xd, xr, None, None, w, terminal=True)
        return _io.vec_out(l, kind), _io.out(gx, kind), 2 * np.asarray(Q_T)
    ud, _ = _io.state_in(u, nu)
    ur, _ = _io.state_in(u_ref, nu)
    w = _weights(Q, R, np.zeros((4, 4)))
    l, gx, gu = bt.cost_derivatives(xd, xr, ud.expand(nu, xd.shape[1]).contiguous(), ur.expand(nu, xd.shape[1]).contiguous(), w)
    return _io.vec_out(l, kind), _io.out(gx, kind), _io.out(gu, kind), 2 * np.asarray(Q), 2 * np.asarray(R)


def compute_costate_trajectory(x_traj, u_traj, x_ref, u_ref):
    """trajectory_generation.py:138-159.  Returns a list of N costates (one problem) or an array (B,N,4)."""
    X, kind = _io.traj_in(x_traj, nx)
    U, _ = _io.traj_in(u_traj, nu)
    lam = _io.out(bt.costate(X, U, _ref(x_ref, u_ref), _weights(), active_params()), kind)
    return lam if kind.batched else list(lam)


def discretize_linearization(Ac, Bc, dt):
    """trajectory_generation.py:161-164"""
    A = np.asarray(Ac.detach().cpu() if isinstance(Ac, torch.Tensor) else Ac, dtype=np.float64)
    batched = A.ndim == 3
    Ad_, kind = _io.state_in(A.reshape(-1, 16) if batched else A.reshape(16), 16)
    Bd_, _ = _io.state_in(np.asarray(Bc.detach().cpu() if isinstance(Bc, torch.Tensor) else Bc, dtype=np.float64)
                          .reshape((-1, 8) if batched else (8,)), 8)
    kind = _io.Kind(Ac, batched)
    Ad, Bd = bt.discretize(Ad_, Bd_, dt)
    return _io.out(Ad, kind, tail=(4, 4), key="Ad"), _io.out(Bd, kind, tail=(4, 2), key="Bd")


def build_stage_lists(x_traj, u_traj, x_ref, u_ref, lambda_seq=None):
    """trajectory_generation.py:166-181.  lambda_seq is accepted and ignored exactly as the reference does
    (stage_blocks_and_affine never reads it, tg:116-129).  One problem -> the reference's nine lists/arrays."""
    X, kind = _io.traj_in(x_traj, nx)
    U, _ = _io.traj_in(u_traj, nu)
    A, Bm, q, r, qT = bt.stage_lists(X, U, _ref(x_ref, u_ref), _weights(), active_params())
    A_, B_ = _io.out(A, kind, tail=(4, 4), key="A"), _io.out(Bm, kind, tail=(4, 2), key="B")
    q_, r_, qT_ = _io.out(q, kind, key="q"), _io.out(r, kind, key="r"), _io.out(qT, kind, key="qT")
    n = X.T - 1
    Q_t, R_t, S_t, QTb = 2 * np.asarray(Q), 2 * np.asarray(R), np.zeros((nu, nx)), 2 * np.asarray(Q_T)
    if kind.batched:
        return A_, B_, Q_t, R_t, S_t, q_, r_, QTb, qT_
    return (list(A_), list(B_), [Q_t] * n, [R_t] * n, [S_t] * n, list(q_), list(r_), QTb, qT_)


def calculate_K_and_sigma(A_list, B_list, Q_list, R_list, S_list, q_list, r_list, Q_T_block, q_T):
    """trajectory_generation.py:183-216 on caller-supplied lists (one problem)."""
    Tn = len(A_list)

    def dev(lst, shape):
        a = np.asarray([np.asarray(v, dtype=np.float64) for v in lst]).reshape(1, Tn, shape)
        return bt.Traj.from_batch_major(bt.upload(a))

    K, S, dJ = bt.riccati_lists(dev(A_list, 16), dev(B_list, 8), dev(Q_list, 16), dev(R_list, 4), dev(q_list, 4),
                                dev(r_list, 2), bt.upload(np.asarray(Q_T_block, dtype=np.float64).reshape(16, 1)),
                                bt.upload(np.asarray(q_T, dtype=np.float64).reshape(4, 1)), S_cross=dev(S_list, 8))
    Kh = K.batch_major()[0].cpu().numpy().reshape(Tn, 2, 4)
    Sh = S.batch_major()[0].cpu().numpy()
    return list(Kh), list(Sh), float(dJ[0].item())


def forward_closed_loop_update(x_traj, u_traj, K, sigma, gamma=1.0):
    """trajectory_generation.py:218-229"""
    X, kind = _io.traj_in(x_traj, nx)
    U, _ = _io.traj_in(u_traj, nu)
    Kd, _ = _io.traj_in(K, 8)
    Sd, _ = _io.traj_in(sigma, nu)
    g = np.atleast_1d(np.asarray(gamma.detach().cpu() if isinstance(gamma, torch.Tensor) else gamma, dtype=np.float64))
    gam = bt.upload(np.broadcast_to(g, (X.B,)).reshape(1, -1).copy())
    ref = bt.Ref(X.clone(), U.clone())  # the cost the kernel also produces is not part of this call: any reference will do
    _, Xn, Un = bt.closed_loop_rollout_cost(X, U, Kd, Sd, ref, _weights(), gam, store=True, params=active_params())
    return _io.out(Xn[0], kind, key="xn"), _io.out(Un[0], kind, key="un")


def total_cost(x_traj, u_traj, x_ref, u_ref, Q, R, Q_T):
    """trajectory_generation.py:231-252"""
    X, kind = _io.traj_in(x_traj, nx)
    U, _ = _io.traj_in(u_traj, nu)
    return _io.vec_out(bt.total_cost(X, U, _ref(x_ref, u_ref), _weights(Q, R, Q_T)), kind)


def armijo_sweep(x_traj, u_traj, K, sigma, x_ref, u_ref, stepsizes_tested=(), n_steps=200):
    """The numeric part of plot_armijo_line_search (trajectory_generation.py:257-264): step sizes
    linspace(0, max(1.25, 1.3 max tested), 200) and the cost along the search direction for each."""
    max_step = max(1.25, max(stepsizes_tested) * 1.3 if len(stepsizes_tested) else 1.25)
    steps = np.linspace(0, max_step, n_steps)
    X, kind = _io.traj_in(x_traj, nx)
    U, _ = _io.traj_in(u_traj, nu)
    Kd, _ = _io.traj_in(K, 8)
    Sd, _ = _io.traj_in(sigma, nu)
    cost = bt.stepsize_sweep(X, U, Kd, Sd, _ref(x_ref, u_ref), _weights(), bt.upload(steps), params=active_params())
    c = cost.cpu().numpy()  # (S, B)
    return steps, (c.T if kind.batched else c[:, 0])


def plot_armijo_line_search(*args, **kwargs):
    """Plotting is out of scope (no matplotlib on the hot path); ``armijo_sweep`` returns the plotted numbers."""
    return None


class _NewtonSlot:
    """Device state + batch-major staging of one in-flight batched solve (see _NewtonPipeline)."""

    def __init__(self, Bn, N, max_iters):
        self.key = (Bn, N, max_iters)
        self.state = bt.newton_alloc(Bn, N, max_iters, history=True)
        self.copied = None  # event: the D2H copies out of this slot have completed


class _NewtonPipeline:
    """Double-buffered solver state and a copy stream, so that the device-to-host copies of one solve overlap the
    kernel of the next one (trajectory_generation.newton_Algorithm(..., block=False)).

    Solve i runs on the caller's stream into slot i mod 2; its results are un-tiled and copied to fresh pinned host
    tensors on the copy stream, which waits for the solve's event; solve i+1 (other slot) starts at once.  A slot is
    reused only after its copies have completed (stream-side wait, no host synchronisation)."""

    def __init__(self, depth=2):
        self.depth, self.slots, self.n, self.copy_stream = depth, {}, 0, None

    def slot(self, Bn, N, max_iters):
        dev_i = torch.cuda.current_device()
        if self.copy_stream is None or self.copy_stream.device.index != dev_i:
            self.copy_stream = torch.cuda.Stream()
            self.slots = {}
        i = self.n % self.depth
        self.n += 1
        sl = self.slots.get(i)
        if sl is None or sl.key != (Bn, N, max_iters):
            sl = self.slots[i] = _NewtonSlot(Bn, N, max_iters)
        if sl.copied is not None:
            torch.cuda.current_stream().wait_event(sl.copied)
        return sl


_pipeline = _NewtonPipeline()


PendingNewton = _io.Pending  # handle of a non-blocking batched solve (result(), ready())


def newton_Algorithm(x0, x_ref, u_ref, max_iters, tol=1e-6, beta=0.7, c=0.5, gamma_0=1, plot_armijo_iters=10, *,
                     Q=None, R=None, Q_T=None, return_history=False, history_stride=1, verbose=True,
                     return_state=False, params_b=None, return_gains=True, block=True, kernel=None):
    """Regularised Newton method with Armijo line search (trajectory_generation.py:298-398).

    One problem: returns ``(x_traj, u_traj, K, sigma, history)`` with the reference's types (K and sigma
    are lists of N-1 arrays; history has 'cost', 'sigma_norm', and - only with return_history=True, because
    it needs one launch per stored iterate - 'x_trajs' and 'sigmas').
    Batch (x0 of shape (B,4)): arrays with a leading batch axis; history['cost'] is (B, iters+1) padded with
    NaN past each problem's last iteration, plus 'iters', 'status', 'n_try', 'gamma'.
    params_b (B, 11): every problem its own physical parameters (domain randomisation).
    return_gains=False: K and sigma (half of the bytes a batched solve returns) stay on the device and come back as
    ``batched.Traj`` handles instead of host arrays.
    block=False (batched, without return_history): returns a ``PendingNewton`` at once; its ``result()`` gives the
    tuple.  The device-to-host copies of this solve then overlap the kernel of the next call.
    kernel: a name of ``batched.KERNEL_VARIANTS`` (default: chosen by batch size).
    """
    u_ref_n = u_ref.shape[0] if u_ref.ndim == 2 else u_ref.shape[1]
    x_ref_n = x_ref.shape[0] if x_ref.ndim == 2 else x_ref.shape[1]
    if u_ref_n == x_ref_n:
        if verbose:
            print(f" u_ref has same length as x_ref ({u_ref_n}). Using first N-1 controls.")
        u_ref = u_ref[:-1] if u_ref.ndim == 2 else u_ref[:, :-1]
        u_ref_n -= 1
    if u_ref_n != x_ref_n - 1:
        raise ValueError(f"Incompatible dimensions: x_ref has {x_ref_n} states but u_ref has {u_ref_n} controls "
                         f"(expected {x_ref_n - 1})")
    x0d, kind = _io.state_in(x0, nx)
    ref = _ref(x_ref, u_ref)
    w = _weights(Q, R, Q_T)
    kw = dict(max_iters=max_iters, tol=tol, beta=beta, c=c, gamma_0=gamma_0, w=w, params=active_params(), kernel=kernel)
    if params_b is not None:  # physical parameters per problem (B, 11): see batched.phys_params
        kw["params_b"] = bt.phys_params(params_b, x0d.shape[1])
    x_trajs, sigmas = [], []

    if return_history:
        # history['x_trajs'][0] is the initial open-loop rollout (tg:322-327); one launch per stored iterate after that
        x_trajs.append(_io.out(bt.rollout_open_loop(x0d, None, N=ref.N, params=active_params(),
                                                    params_b=kw.get("params_b")), kind))
        st, done = None, 0
        while True:
            st = bt.newton_solve(x0d, ref, state=st, chunk_iters=history_stride, **kw)
            done += history_stride
            # the reference logs sigma in every iteration (tg:341-342) and the iterate only after an ACCEPTED step
            # (tg:387-388): a line-search failure breaks out before that (tg:367-369)
            sigmas.append(_io.out(st.S, kind))
            if not bool((st.status == 3).all()):
                x_trajs.append(_io.out(st.X, kind))
            if bool((st.status != 0).all()) or done >= max_iters:
                break
        pend = _finish_newton(st, kind, verbose, return_gains, x_trajs, sigmas, return_state, None)
        return pend.result() if block else pend
    sl = _pipeline.slot(x0d.shape[1], ref.N, int(max_iters)) if kind.batched else None
    if sl is not None:
        sl.state.reset()
    st = bt.newton_solve(x0d, ref, state=None if sl is None else sl.state, **kw)
    pend = _finish_newton(st, kind, verbose, return_gains, x_trajs, sigmas, return_state, sl)
    return pend if (not block and kind.batched) else pend.result()



def _finish_newton(st, kind, verbose, return_gains, x_trajs, sigmas, return_state, slot):
    """Queue the un-tiling and the device-to-host copies of a finished (queued) solve on the copy stream; the returned
    handle builds the reference-shaped tuple once they are done."""
    host = not (kind.torch and kind.cuda)
    solved = torch.cuda.Event()
    solved.record()
    cs = _pipeline.copy_stream if slot is not None else None
    outs = {}
    ctx = torch.cuda.stream(cs) if cs is not None else _Null()
    with ctx:
        if cs is not None:
            cs.wait_event(solved)

        def fetch(name, dev_t):
            if host:
                h = _io.pinned_empty(dev_t.shape, dev_t.dtype)
                h.copy_(dev_t, non_blocking=True)
                outs[name] = (h, dev_t)  # keep the device staging alive until the copy has run
            else:
                outs[name] = (dev_t, None)

        def traj(name, t, tail=None):
            a = t.batch_major()
            if tail is not None:
                a = a.reshape(*a.shape[:-1], *tail)
            fetch(name, a if kind.batched else a[0])

        traj("x", st.X)
        traj("u", st.U)
        if return_gains:
            traj("K", st.K, (2, 4))
            traj("S", st.S)
        for name in ("iters", "status", "hist_cost", "hist_sigma_norm", "hist_ntry", "hist_gamma"):
            fetch(name, getattr(st, name))
        done = torch.cuda.Event()
        done.record()
    if slot is not None:
        slot.copied = done

    def build():
        g = {k: (v[0] if kind.torch else v[0].numpy()) if host else v[0] for k, v in outs.items()}
        np_of = (lambda v: v.numpy()) if host else (lambda v: v.cpu().numpy())
        iters, status = np_of(outs["iters"][0]), np_of(outs["status"][0])
        n_it = int(iters.max()) if iters.size else 0
        hc = np_of(outs["hist_cost"][0])[:n_it + 1].T  # (B, n_it+1)
        hs = np_of(outs["hist_sigma_norm"][0])[:n_it].T
        hn = np_of(outs["hist_ntry"][0])[:n_it].T
        hg = np_of(outs["hist_gamma"][0])[:n_it].T
        if verbose:
            if not kind.batched:  # the reference's progress line, every 10th iteration (tg:391-392)
                for k in range(0, len(hc[0]) - 1, 10):
                    if np.isfinite(hc[0][k + 1]):
                        print(f"Iter {k}: Cost={hc[0][k + 1]:.2f}, diff_cost={hc[0][k] - hc[0][k + 1]:.2e}, ")
            for b in np.where(status == 3)[0][:8]:
                print(f"Iteration {iters[b] - 1}: Line search failed to find sufficient decrease.")
            for b in np.where(status == 1)[0][:8]:
                print(f"Converged at iteration {iters[b] - 1}!")
        K, sigma = (g["K"], g["S"]) if return_gains else (st.K, st.S)
        if kind.batched:
            history = {"cost": hc, "sigma_norm": hs, "iters": iters, "status": status, "n_try": hn, "gamma": hg,
                       "x_trajs": x_trajs, "sigmas": sigmas}
        else:
            # accepted steps: every iteration but a final one whose line search failed (tg:367-369)
            n_acc = max(int(iters[0]) - (1 if int(status[0]) == 3 else 0), 0)  # history['cost'] always holds the initial cost (tg:322-327)
            history = {"cost": list(hc[0][:n_acc + 1]), "sigma_norm": list(hs[0][:int(iters[0])]), "x_trajs": x_trajs,
                       "sigmas": sigmas, "iters": int(iters[0]), "status": int(status[0]),
                       "n_try": list(hn[0][:int(iters[0])]), "gamma": list(hg[0][:int(iters[0])])}
            if return_gains:
                K, sigma = list(K), list(sigma)
        if return_state:
            return g["x"], g["u"], K, sigma, history, st
        return g["x"], g["u"], K, sigma, history

    return PendingNewton(done, build)


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
