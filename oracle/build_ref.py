"""Recipe that makes the UNMODIFIED reference travel to the GPU box (test / benchmark infrastructure only).

The reference is pure Python (dynamics.py + trajectory_generation.py + trajectory_tracking.py, SymPy + NumPy), so
"building" it means copying those three files and the two trajectory files they load, byte for byte, from
``/root/reference`` into ``oracle/_ref/``.  That directory is git-ignored (reference sources never enter the
repository) but not gpurun-ignored, so `bench.py --impl reference` and the `cpu_baseline` leg can time the real
reference on the GPU box's host cores (``oracle/ref_import.py`` finds it there).  ``__graft_entry__.build()`` runs
this when ``/root/reference`` is present; on the GPU box the copied files are used as they are.

    python oracle/build_ref.py            # copy + write MANIFEST.json (sha256 of source and copy)
    python oracle/build_ref.py --check    # verify the copies against the manifest
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("ACRO_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
FILES = ["dynamics.py", "trajectory_generation.py", "trajectory_tracking.py",
         os.path.join("trajectories_npz", "fully_actuated_trajectory.npz"),
         os.path.join("trajectories_npz", "acrobot_optimal_trajectory.npz")]


def sha256(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(src=SRC, dst=DST):
    """Copy the reference files; returns the manifest.  No-op (returns None) when the source tree is absent."""
    if not os.path.isfile(os.path.join(src, "dynamics.py")):
        return None
    manifest = {}
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = sha256(d)
        assert manifest[rel] == sha256(s)
    with open(os.path.join(dst, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1)
    return manifest


def check(dst=DST):
    """True if oracle/_ref holds exactly the files the manifest describes (i.e. the unmodified reference)."""
    path = os.path.join(dst, "MANIFEST.json")
    if not os.path.isfile(path):
        return False
    man = json.load(open(path))["sha256"]
    return sorted(man) == sorted(FILES) and all(os.path.isfile(os.path.join(dst, r)) and sha256(os.path.join(dst, r)) == h
                                                for r, h in man.items())


if __name__ == "__main__":
    if "--check" in sys.argv:
        print("oracle/_ref ok" if check() else "oracle/_ref missing or modified")
        sys.exit(0 if check() else 1)
    m = build()
    print("reference tree not found at %s" % SRC if m is None else "copied %d files to %s" % (len(m), DST))
