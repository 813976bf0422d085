"""Drop-in for the reference's ``dynamics.py`` - same names, arguments and return shapes.

``dynamics(xx, uu)``, ``continuous_dynamics(xx, uu)`` and ``Calculate_A_B_matrixes(x_t, u_t)`` take one
state (4,) / input (2,) like the reference, or a batch (B,4)/(B,2), as NumPy arrays or torch tensors, and
run on the GPU through libacro_b200.so.  There is no SymPy model here: the kernels evaluate the closed
form of the same equations (include/acro_abi.h, csrc/acro_device.cuh).
"""
import numpy as np

from . import _io
from . import batched as bt

# dynamics.py:173-175
dt = 2e-2
ns = 4
ni = 2

# dynamics.py:15-61 (plain-string keys instead of SymPy symbols)
params_1 = dict(bt.PARAM_SETS[1])
params_2 = dict(bt.PARAM_SETS[2])
params_3 = dict(bt.PARAM_SETS[3])

_active = {"params": bt.make_params(1, dt)}


def set_params(version_num):
    """dynamics.py:117-144.  As in the reference, calling this does NOT change what ``dynamics`` and
    ``Calculate_A_B_matrixes`` evaluate (they stay on parameter set 1, dynamics.py:147): it returns the
    selected parameter set.  Use ``use_params`` to switch the model the kernels run."""
    if version_num not in (1, 2, 3):
        print("Invalid parameter version number, setting the default one.")
        version_num = 1
    return bt.make_params(version_num, dt)


def use_params(version_num=1, actuated_tau1=False, **overrides):
    """Extension: make parameter set ``version_num`` (optionally with overrides, optionally the fully-actuated
    plant of fully_actuated_ref_gen.py:20-73) the model used by this module's functions."""
    _active["params"] = bt.make_params(version_num, dt, actuated_tau1, **overrides)
    return _active["params"]


def active_params():
    return _active["params"]


def _xu(xx, uu):
    x, kind = _io.state_in(xx, ns)
    u, ku = _io.state_in(uu, ni)
    if u.shape[1] != x.shape[1]:
        if u.shape[1] == 1:
            u = u.expand(ni, x.shape[1]).contiguous()
        else:
            raise ValueError("batch sizes of xx and uu differ")
    return x, u, kind


def _pb(params_b, Bn):
    """Keyword-only extension of the functions below: physical parameters per problem, (B, 11) in the order
    m1, m2, l1, lc1, l2, lc2, I1, I2, g, f1, f2, or a dict of per-problem arrays (see batched.phys_params)."""
    if params_b is None:
        return None
    t = bt.phys_params(params_b, Bn)
    if t.shape[1] != Bn:
        raise ValueError("params_b: %d parameter sets for %d problems" % (t.shape[1], Bn))
    return t


def dynamics(xx, uu, *, params_b=None):
    """One RK4 step of the acrobot (dynamics.py:177-195)."""
    x, u, kind = _xu(xx, uu)
    return _io.out(bt.rk4_step(x, u, _active["params"], _pb(params_b, x.shape[1])), kind)


def continuous_dynamics(xx, uu, *, params_b=None):
    """x_dot = f(x, u) (dynamics.py:197-213)."""
    x, u, kind = _xu(xx, uu)
    return _io.out(bt.continuous_dynamics(x, u, _active["params"], _pb(params_b, x.shape[1])), kind)


def Calculate_A_B_matrixes(x_t, u_t, *, params_b=None):
    """Continuous Jacobians A_c (4,4), B_c (4,2) (dynamics.py:217-226); batched input -> (B,4,4), (B,4,2)."""
    x, u, kind = _xu(x_t, u_t)
    A, Bm = bt.linearize(x, u, False, _active["params"], _pb(params_b, x.shape[1]))
    Bn = x.shape[1]
    return (_io.out(A.reshape(16, Bn), kind, tail=(4, 4), key="A"), _io.out(Bm.reshape(8, Bn), kind, tail=(4, 2), key="B"))
