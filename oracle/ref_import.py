"""Loader for the UNMODIFIED reference modules (test infrastructure only).

This file is part of the oracle: only ``tests/``, ``tests/golden/make_golden.py``
and ``bench.py``'s cpu_baseline leg may import it.  It imports the reference's
``dynamics``, ``trajectory_generation`` and ``trajectory_tracking`` modules from
``/root/reference`` without changing them:

* ``matplotlib`` (imported at trajectory_generation.py:2) and ``casadi``
  (trajectory_tracking.py:2) are not installed in this image, so empty stub
  modules are put in ``sys.modules`` first;
* ``plot_armijo_line_search`` (trajectory_generation.py:254) is replaced by a
  no-op after import, otherwise ``plt.figure`` fails inside ``newton_Algorithm``
  (trajectory_generation.py:372-380);
* the working directory is switched to the reference root while a reference
  function that opens relative npz paths runs (trajectory_generation.py:513).

``/root/reference`` exists only in the build container.  On the GPU box the
loader falls back to ``oracle/_ref`` (the byte-for-byte copy made by
``oracle/build_ref.py``, git-ignored, shipped by gpurun) so that bench.py can time
the real reference there; when neither is present ``available()`` is False and
the committed fixtures under ``tests/golden/`` stand in for it.
"""
import contextlib
import os
import sys
import types

def _find_root():
    """$ACRO_REFERENCE_ROOT, else the read-only tree of the build container, else the byte-for-byte copy that
    oracle/build_ref.py ships to the GPU box (oracle/_ref, git-ignored)."""
    env = os.environ.get("ACRO_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isfile("/root/reference/dynamics.py"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


REF_ROOT = _find_root()


def available():
    if not os.path.isfile(os.path.join(REF_ROOT, "dynamics.py")):
        return False
    try:
        import sympy  # noqa: F401  (dynamics.py builds its model symbolically at import time)
    except Exception:
        return False
    return True


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__["__stub__"] = True
    sys.modules[name] = m
    return m


@contextlib.contextmanager
def in_ref_dir():
    old = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        yield
    finally:
        os.chdir(old)


_cache = {}


def load():
    """Return (dynamics, trajectory_generation, trajectory_tracking) reference modules."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _stub("matplotlib")
        plt = _stub("matplotlib.pyplot")
        anim = _stub("matplotlib.animation")
        mpl.pyplot = plt
        mpl.animation = anim
        anim.FuncAnimation = object
    try:
        import casadi  # noqa: F401
    except Exception:
        _stub("casadi")
    # the reference modules are imported under private names so they never shadow
    # the product's drop-in modules of the same names
    saved = {k: sys.modules.pop(k, None) for k in ("dynamics", "trajectory_generation", "trajectory_tracking")}
    sys.path.insert(0, REF_ROOT)
    try:
        with in_ref_dir():
            import dynamics as rd
            import trajectory_generation as rtg
            import trajectory_tracking as rtt
    finally:
        sys.path.remove(REF_ROOT)
        for k in ("dynamics", "trajectory_generation", "trajectory_tracking"):
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
    rtg.plot_armijo_line_search = lambda *a, **k: None
    rtt.plot_armijo_line_search = rtg.plot_armijo_line_search
    _cache["mods"] = (rd, rtg, rtt)
    return _cache["mods"]
