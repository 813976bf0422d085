"""gymnast_optimalcontrol_b200 - the batched acrobot optimal-control hot path on NVIDIA B200.

Drop-in modules with the reference's function names and signatures:

    from gymnast_optimalcontrol_b200 import dynamics, trajectory_generation, trajectory_tracking

Device-level API on torch CUDA tensors (structure-of-arrays): ``gymnast_optimalcontrol_b200.batched``.
All compute goes through the C ABI of libacro_b200.so (include/acro_abi.h); importing the package
fails if that library has not been built, and calling it fails without a CUDA device.
"""
from . import _abi  # noqa: F401  (raises ImportError if the CUDA library is missing)
from . import batched  # noqa: F401

__version__ = "0.1.0"
