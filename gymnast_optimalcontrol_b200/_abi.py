"""ctypes binding of include/acro_abi.h (libacro_b200.so).

This is the binding a maintainer of the reference would add.  There is no
CPU fallback: if the library is missing, importing this module raises.
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
# (ACRO_B200_LIB: development hook for A/B runs of two builds of the library; the default is the in-tree build)
LIB_PATH = os.environ.get("ACRO_B200_LIB") or os.path.join(PKG, "libacro_b200.so")

OK, E_INVALID, E_CUDA = 0, -1, -2
RUNNING, CONVERGED, MAX_ITERS, LINE_SEARCH_FAILED = 0, 1, 2, 3
NEWTON_KERNELS = {"auto": 0, "duo": 1, "ring": 2, "thread": 3, "ldg": 3, "spec": 4}


class AcroParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("m1", "m2", "l1", "lc1", "l2", "lc2", "I1", "I2", "g", "f1", "f2", "dt")] + [
        ("actuated_tau1", C.c_int32), ("reserved", C.c_int32)]


class AcroWeights(C.Structure):
    _fields_ = [("Q", C.c_double * 16), ("R", C.c_double * 4), ("QT", C.c_double * 16),
                ("Q_b", C.c_void_p), ("R_b", C.c_void_p), ("QT_b", C.c_void_p)]


class AcroRef(C.Structure):
    _fields_ = [("x", C.c_void_p), ("u", C.c_void_p), ("per_problem", C.c_int32), ("reserved", C.c_int32)]


class AcroNewtonOpts(C.Structure):
    _fields_ = [("max_iters", C.c_int32), ("chunk_iters", C.c_int32), ("max_line_search", C.c_int32),
                ("init", C.c_int32), ("tol", C.c_double), ("beta", C.c_double), ("c", C.c_double),
                ("gamma_0", C.c_double), ("kernel", C.c_int32), ("stage_steps", C.c_int32),
                ("recompute_lin", C.c_int32), ("speculate", C.c_int32), ("spec_ws", C.c_void_p)]


P = C.c_void_p  # device pointer / stream
PP, PW, PR, PO = C.POINTER(AcroParams), C.POINTER(AcroWeights), C.POINTER(AcroRef), C.POINTER(AcroNewtonOpts)
I64, I32, F64 = C.c_int64, C.c_int, C.c_double

# name -> argtypes; every function returns int except the three queries
SIGNATURES = {
    "acro_continuous_dynamics": [PP, I64, P, P, P, P],
    "acro_rk4_step": [PP, I64, P, P, P, P],
    "acro_linearize": [PP, I64, P, P, P, P, I32, P],
    "acro_rollout_open_loop": [PP, I64, I32, P, P, P, P],
    "acro_equilibrium": [PP, P, I64, P, P, F64, I32, P, P, P],
    "acro_continuous_dynamics_pp": [PP, P, I64, P, P, P, P],
    "acro_rk4_step_pp": [PP, P, I64, P, P, P, P],
    "acro_linearize_pp": [PP, P, I64, P, P, P, P, I32, P],
    "acro_rollout_open_loop_pp": [PP, P, I64, I32, P, P, P, P],
    "acro_lqr_track_pp": [PP, P, I64, I32, PR, P, P, P, P, P],
    "acro_total_cost": [PW, I64, I32, P, P, PR, P, P],
    "acro_costate": [PP, PW, I64, I32, P, P, PR, P, P],
    "acro_cost_derivatives": [PW, I64, P, P, P, P, I32, P, P, P, P],
    "acro_discretize": [I64, P, P, F64, P, P, P],
    "acro_stage_lists": [PP, PW, I64, I32, P, P, PR, P, P, P, P, P, P],
    "acro_riccati_lists": [I64, I32] + [P] * 12 + [P],
    "acro_riccati_affine": [PP, PW, I64, I32, P, P, PR, P, P, P, P, P],
    "acro_closed_loop_rollout_cost": [PP, PW, I64, I32, P, P, P, P, PR, I32, P, I32, P, P, P, P],
    "acro_armijo_select": [I64, I32, P, P, P, I32, P, F64, P, P],
    "acro_newton_solve": [PP, PW, PO, I64, I32, P, PR] + [P] * 17 + [P],
    "acro_newton_solve_pp": [PP, P, PW, PO, I64, I32, P, PR] + [P] * 17 + [P],
    "acro_newton_describe": [PO, I64, I32, I32, I32, C.c_char_p, I32],
    "acro_stepsize_sweep": [PP, PW, I64, I32, P, P, P, P, PR, I32, P, P, P],
    "acro_lqr_gains": [PP, PW, I64, I32, PR, P, P],
    "acro_lqr_track": [PP, I64, I32, PR, P, P, P, P, P],
    "acro_p_inf": [PW, I64, P, P, I32, F64, P, P, P],
    "acro_mpc_solve": [PW, I64, I32, P, P, P, P, P, P, P, P, P],
    "acro_mpc_track": [PP, PW, I64, I32, I32, I32, PR, C.POINTER(C.c_double), C.POINTER(C.c_double), P, I32, P, P, P,
                       P, P, C.POINTER(C.c_int64), P],
    "acro_mpc_track_pp": [PP, P, PW, I64, I32, I32, I32, PR, C.POINTER(C.c_double), C.POINTER(C.c_double), P, I32, P, P,
                          P, P, C.POINTER(C.c_int64), P],
    "acro_mpc_track_box": [PP, PW, I64, I32, I32, I32, PR, C.POINTER(C.c_double), C.POINTER(C.c_double), P, I32, P, F64, I32,
                           P, P, P, P, P, P, P, P],
    "acro_mpc_track_box_pp": [PP, P, PW, I64, I32, I32, I32, PR, C.POINTER(C.c_double), C.POINTER(C.c_double), P, I32, P, F64,
                              I32, P, P, P, P, P, P, P, P],
    "acro_bench_fp64_peak": [I32, I32, I32, P, P],
    "acro_bench_fp64_chain": [I32, I32, I32, I32, I32, P, P, P],
    "acro_transpose": [I64, I64, P, P, P],
    "acro_pack_soa": [I64, I32, I32, P, P, P],
    "acro_unpack_soa": [I64, I32, I32, P, P, P],
}
QUERIES = {"acro_version": C.c_char_p, "acro_last_error_string": C.c_char_p, "acro_launch_count": C.c_int64}
SIZES = {"acro_mpc_box_ws_doubles": ([I64, I32, I32], C.c_int64), "acro_newton_spec_ws_doubles": ([I64, I32], C.c_int64)}


class AcroError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libacro_b200.so is not built (%s).  Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`python -m gymnast_optimalcontrol_b200._build`.  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = C.c_int
    for name, res in QUERIES.items():
        f = getattr(lib, name)
        f.argtypes = []
        f.restype = res
    for name, (args, res) in SIZES.items():
        f = getattr(lib, name)
        f.argtypes = args
        f.restype = res
    return lib


lib = _load()


def check(rc, what):
    if rc != 0:
        raise AcroError("%s failed (%d): %s" % (what, rc, lib.acro_last_error_string().decode()))


def call(name, *args):
    check(getattr(lib, name)(*args), name)


def version():
    return lib.acro_version().decode()


def launch_count():
    return int(lib.acro_launch_count())
