"""CPU-only checks: the C-ABI library loads and exports every symbol include/acro_abi.h declares, the ctypes
binding covers them, and the host-side logic of the drop-in modules (argument checks, error behaviour)
mirrors the reference.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "acro_abi.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(acro_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from gymnast_optimalcontrol_b200 import _abi
    lib = ctypes.CDLL(_abi.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libacro_b200.so does not export %s" % n
    bound = set(_abi.SIGNATURES) | set(_abi.QUERIES) | set(_abi.SIZES)
    assert set(names) == bound, (set(names) ^ bound)
    assert "sm_100a" in _abi.version()
    assert _abi.launch_count() == 0


def test_struct_layouts_match_header(tmp_path):
    """sizeof / offsetof of every struct of include/acro_abi.h, as gcc lays them out, against the ctypes mirror."""
    import subprocess
    from gymnast_optimalcontrol_b200 import _abi
    structs = {"AcroParams": _abi.AcroParams, "AcroWeights": _abi.AcroWeights, "AcroRef": _abi.AcroRef,
               "AcroNewtonOpts": _abi.AcroNewtonOpts}
    lines = []
    for name, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (name, name))
        for f, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (name, f, name, f))
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "%s"\nint main(void){%s return 0;}\n' % (HEADER, "".join(lines))
    c, exe = tmp_path / "layout.c", tmp_path / "layout"
    c.write_text(src)
    subprocess.run(["gcc", "-o", str(exe), str(c)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, cls in structs.items():
        assert int(got[name]) == ctypes.sizeof(cls), name
        for f, _ in cls._fields_:
            assert int(got["%s.%s" % (name, f)]) == getattr(cls, f).offset, (name, f)
    assert ctypes.sizeof(_abi.AcroNewtonOpts) == 16 + 32 + 16 + 8


def test_invalid_arguments_return_error_codes_without_a_gpu():
    from gymnast_optimalcontrol_b200 import _abi
    p = _abi.AcroParams()
    rc = _abi.lib.acro_rk4_step(ctypes.byref(p), 0, None, None, None, None)
    assert rc == _abi.E_INVALID
    assert b"acro_rk4_step" in _abi.lib.acro_last_error_string()
    with pytest.raises(_abi.AcroError):
        _abi.call("acro_pack_soa", 0, 1, 4, None, None, None)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from gymnast_optimalcontrol_b200 import dynamics
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dynamics.dynamics(np.zeros(4), np.zeros(2))


def test_newton_argument_checks_mirror_reference():
    """ValueError on incompatible u_ref length before anything else happens (trajectory_generation.py:305-306)."""
    from gymnast_optimalcontrol_b200 import trajectory_generation as tg
    with pytest.raises(ValueError, match="Incompatible dimensions"):
        tg.newton_Algorithm(np.zeros(4), np.zeros((501, 4)), np.zeros((400, 2)), 3)
    assert tg.N == 501 and tg.nx == 4 and tg.nu == 2 and tg.dt == 2e-2
    assert np.array_equal(tg.Q, np.diag([130.0, 30.0, 0.0001, 0.0001]))
    assert np.array_equal(tg.R, np.diag([1e-6, 1.5])) and np.array_equal(tg.Q_T, np.diag([130, 130.0, 1.0, 1.0]))


def test_weights_must_be_symmetric():
    from gymnast_optimalcontrol_b200 import batched as bt
    Q = np.eye(4)
    Q[0, 1] = 1.0
    with pytest.raises(ValueError, match="symmetric"):
        bt.Weights(Q, np.eye(2))


def test_reference_builders_match_reference_shapes():
    from gymnast_optimalcontrol_b200 import trajectory_generation as tg
    t, x, u = tg.define_reference_piecewise(10.0, [0, 0, 0, 0], [0.1, -0.1, 0, 0], [0, 0], [0.5, 0.5])
    assert x.shape == (501, 4) and u.shape == (501, 2) and t.shape == (501,)
    assert np.all(x[:250] == 0) and np.all(x[250:, 0] == 0.1) and np.all(u[250:] == 0.5)
    g = np.load(os.path.join(ROOT, "tests", "golden", "newton_task1.npz"))
    t2, x2, u2 = tg.define_reference_piecewise(10.0, g["x_e1"], g["x_e2"], g["u_e1"], g["u_e2"])
    assert np.array_equal(x2, g["x_ref"]) and np.array_equal(u2, g["u_ref"]) and np.array_equal(t2, g["t_ref"])
    xr, ur, tr = tg.get_fully_actuated_ref(os.path.join(ROOT, "tests", "golden", "fully_actuated_trajectory.npz"))
    g2 = np.load(os.path.join(ROOT, "tests", "golden", "newton_task2.npz"))
    assert np.array_equal(xr, g2["x_ref"]) and np.array_equal(ur, g2["u_ref"])


def test_compute_equilibrium_matches_reference_root_finder():
    """main.py:33-38 (task_1): the two equilibria the reference obtains with scipy.optimize.root(hybr)."""
    from gymnast_optimalcontrol_b200 import trajectory_generation as tg
    g = np.load(os.path.join(ROOT, "tests", "golden", "newton_task1.npz"))
    x_e1, u_e1 = tg.compute_equilibrium(np.array([0.0, 0.0]), (0.1, -0.1))
    x_e2, u_e2 = tg.compute_equilibrium(np.array([0.5, 0.5]), (0.35, -0.35))
    assert np.max(np.abs(x_e1 - g["x_e1"])) < 1e-10 and np.max(np.abs(x_e2 - g["x_e2"])) < 1e-10
    assert np.array_equal(u_e1, g["u_e1"]) and np.array_equal(u_e2, g["u_e2"])
    with pytest.raises(RuntimeError, match="Root finder failed"):
        tg.compute_equilibrium(np.array([1e3, 1e3]), (0.1, -0.1))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gymnast_optimalcontrol_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), "%s mentions the oracle" % f


def test_headline_kernels_carry_no_yield():
    """ptxas decides by heuristics of its own whether to put a YIELD at the head of the step loops of the
    warp-specialised Newton kernel; when it does, every time step of the recurrence warp pays about 50 cycles (-10 %,
    measured: DESIGN.md 4.1 and 7).  The variants the dispatch uses for config 2 (shared / per-problem references,
    <= 148 and <= 296 tiles) must stay free of it: a harmless-looking source change can flip the decision."""
    import shutil
    import subprocess
    from gymnast_optimalcontrol_b200 import _abi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _abi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    counts, name = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            counts[name] = 0
        elif name and " YIELD " in line:
            counts[name] += 1
    used = ["k_newton_duoILb0ELb0ELi16E", "k_newton_duoILb0ELb0ELi4E", "k_newton_duoILb1ELb1ELi8E", "k_newton_duoILb1ELb1ELi4E",
            "k_newton_duoILb1ELb0ELi16E"]
    for tag in used:
        hits = [n for n in counts if tag in n]
        assert hits, "kernel variant %s not found in the library" % tag
        for n in hits:
            assert counts[n] == 0, "%s: %d YIELD instructions" % (n, counts[n])
    # k_newton_spec (blocks of eight warps): ptxas puts a YIELD at the head of every loop that contains an inlined
    # mbarrier phase check, so its step loops wait behind a call (mbar_wait_call, NI variants of the duo loops).  Every
    # loop that does real FP64 work per trip (the chain / trailer / candidate steps) must be free of YIELD; the spin loop
    # inside mbar_wait_call and the cold paths may keep theirs.
    import re
    spec = [n for n in counts if "k_newton_specILb0ELb0ELi16E" in n]
    assert spec
    body, cur = [], None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
        elif cur == spec[0] and re.search(r"/\*[0-9a-f]{4,6}\*/", line):
            body.append(line)
    addr = lambda l: int(re.search(r"/\*([0-9a-f]{4,6})\*/", l).group(1), 16)
    loops = []
    for l in body:
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", l)
        if m and int(m.group(1), 16) < addr(l):
            loops.append((int(m.group(1), 16), addr(l)))
    hot = 0
    for lo, hi in loops:
        ins = [l for l in body if lo <= addr(l) <= hi]
        fp64 = sum(1 for l in ins if re.search(r"\bD(FMA|MUL|ADD)\b", l))
        reads_ring = any(re.search(r"\bLDS\b", l) for l in ins)  # (the commit loop gathers from global memory: not a ring consumer)
        if 150 <= fp64 and len(ins) < 800 and reads_ring:  # a step loop (forward chain / candidate / trailer), not an enclosing pass loop
            hot += 1
            assert not any(" YIELD " in l for l in ins), "k_newton_spec: YIELD in the step loop at 0x%x" % lo
    assert hot >= 4


def test_staged_reference_is_the_unmodified_reference():
    """oracle/build_ref.py stages the reference for the GPU box's CPU arm: byte-for-byte copies (sha256 of source and
    copy recorded), git-ignored, and ref_import finds the tree."""
    from oracle import build_ref, ref_import
    if os.path.isfile("/root/reference/dynamics.py"):
        man = build_ref.build()
        assert sorted(man) == sorted(build_ref.FILES)
        for rel, h in man.items():
            assert build_ref.sha256(os.path.join("/root/reference", rel)) == h
    if os.path.isdir(build_ref.DST):
        assert build_ref.check()
    ign = open(os.path.join(ROOT, ".gitignore")).read()
    assert "oracle/_ref/" in ign
    assert "oracle/_ref" not in open(os.path.join(ROOT, ".gpurunignore")).read()
    assert ref_import.REF_ROOT in ("/root/reference", build_ref.DST) or "ACRO_REFERENCE_ROOT" in os.environ


def test_kernel_variants_and_plan():
    """Kernel selection is an AcroNewtonOpts field (no environment variable): acro_newton_describe names what
    acro_newton_solve would launch; bad combinations are refused."""
    from gymnast_optimalcontrol_b200 import _abi
    from gymnast_optimalcontrol_b200 import batched as bt
    assert bt.newton_kernel_name(4096) == "acro::k_newton_duo<false,false,16>"
    assert bt.newton_kernel_name(4096, ref_per_problem=True) == "acro::k_newton_duo<true,true,8>"
    assert bt.newton_kernel_name(8192) == "acro::k_newton_duo<false,false,4>"
    assert bt.newton_kernel_name(16384) == "acro::k_newton_ring<false,false,4,false>"
    assert bt.newton_kernel_name(65536) == "acro::k_newton_ring<false,false,2,true>"
    assert bt.newton_kernel_name(65536, kernel="ring2") == "acro::k_newton_ring<false,false,2,false>"
    assert bt.newton_kernel_name(64, kernel="ldg") == "acro::k_newton<false,false,false>"
    assert bt.newton_kernel_name(64, params_per_problem=True) == "acro::k_newton_duo<true,false,16,true>"
    assert bt.newton_kernel_name(64, params_per_problem=True, kernel="ldg") == "acro::k_newton<false,false,true>"
    # a large initial step size back-tracks: the automatic choice is the speculative kernel (one tile per SM only)
    assert bt.newton_kernel_name(4096, gamma_0=1.0) == "acro::k_newton_spec<false,false,16>"
    assert bt.newton_kernel_name(4096, gamma_0=0.49) == "acro::k_newton_duo<false,false,16>"
    assert bt.newton_kernel_name(8192, gamma_0=1.0) == "acro::k_newton_duo<false,false,4>"
    with pytest.raises(_abi.AcroError):
        bt.newton_kernel_name(64, kernel="ring", stage_steps=8)
    with pytest.raises(_abi.AcroError):
        bt.newton_kernel_name(64, kernel="duo", ref_per_problem=True, stage_steps=16)
    src = open(os.path.join(ROOT, "gymnast_optimalcontrol_b200", "csrc", "acro_kernels.cu")).read()
    assert "getenv" not in src


def test_pinned_pool_never_hands_out_memory_that_is_still_referenced(monkeypatch):
    """Results live in pinned host blocks that are recycled only when every tensor / NumPy view of them is gone
    (successive results must not alias; round 2: no page-locking in the steady state)."""
    import torch
    from gymnast_optimalcontrol_b200 import _io
    orig = torch.empty
    monkeypatch.setattr(torch, "empty", lambda *a, **k: orig(*a, **{kk: v for kk, v in k.items() if kk != "pin_memory"}))
    pool = _io.PinnedPool()
    a = pool.empty((4, 5))
    pa = a.data_ptr()
    b = pool.empty((4, 5))
    assert b.data_ptr() != pa
    del a
    c = pool.empty((4, 5))
    assert c.data_ptr() == pa  # recycled once dropped
    row, arr = c[1], c.numpy()
    del c
    assert pool.empty((4, 5)).data_ptr() != pa  # a view and a NumPy array still point into the block
    del row
    assert pool.empty((4, 5)).data_ptr() != pa
    del arr
    assert pool.empty((4, 5)).data_ptr() == pa
    assert pool.empty((3,), torch.int32).dtype == torch.int32
