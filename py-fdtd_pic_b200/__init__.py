"""pyfdtd_b200 -- B200-native (sm_100a) implementation of the Py-FDTD_PIC time-stepping hot path.

The package directory is ``py-fdtd_pic_b200/`` (not an importable identifier); ``pyfdtd_b200.py`` at
the repository root loads it under the name ``pyfdtd_b200``.  Sub-modules carry the reference's own
module names so reference scripts port by changing the import prefix only:

    from pyfdtd_b200 import MasterController as MC, Environment_Setup as envDef, Solver_Engine as SE

Compute runs in ``libpyfdtd_b200.so`` (hand-written CUDA, C-ABI in include/pyfdtd_b200.h); there is
no CPU fallback.
"""
__version__ = "0.1.0"

from . import _native  # noqa: F401  (does not load the library until first use)
