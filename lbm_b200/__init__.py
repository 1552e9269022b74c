"""lbm_b200 — B200-native (sm_100a) replacement for lbm-wgpu's per-timestep lattice update.

The product is `libblbm.so` (CUDA kernels + the C ABI of include/blbm.h).  This package is the thin
host-side mirror of the reference's `pub struct LBM` (lbm-wgpu/src/lbm.rs:32-98) over that ABI via
ctypes, used by the tests, the bench and Python callers.  There is no CPU path: importing works
anywhere, but constructing an `LBM` without the built library or without a GPU raises.
"""
from .lbm import (LBM, SlabGroup, BlbmError, SummaryStat, ColorMap, Kernel, POP_NAMES, library_path, load_library,
                  omega_from_viscosity, rasterize_line)

__all__ = ["LBM", "SlabGroup", "BlbmError", "SummaryStat", "ColorMap", "Kernel", "POP_NAMES", "library_path",
           "load_library", "omega_from_viscosity", "rasterize_line"]
