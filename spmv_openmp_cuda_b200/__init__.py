"""B200-native SpMV engine (fp64 CSR / ELLPACK), a drop-in for the GPU path of
andreadiiorio/SpMV_openMP_CUDA.  The compute path is hand-written sm_100a CUDA behind the C ABI of
include/spmv_b200.h; this package is the host-side mirror of the reference's interface.
There is no CPU fallback."""
from . import capi, engine, synth  # noqa: F401
from .capi import SpmvB200Error  # noqa: F401
from .engine import *  # noqa: F401,F403

__version__ = "0.1.0"
