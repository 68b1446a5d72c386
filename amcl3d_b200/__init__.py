"""amcl3d_b200 -- B200 (sm_100a) implementation of amcl3d's measurement-update hot path.

Layout
  csrc/   hand-written CUDA kernels + the extern "C" ABI declared in include/amcl3d_cuda.h
  host/   C++ classes with the reference's API (Grid3d, ParticleFilter, PointCloudTools) over that ABI
  capi.py ctypes binding of the C-ABI (what tests and bench.py drive)
  synth.py seeded synthetic maps / clouds / particle sets for the BASELINE.json configurations
  build.py nvcc / g++ recipes (in-tree .so files)

There is no CPU fallback: creating a Context without a CUDA device raises.
"""
from .capi import Amcl3dCudaError, Context, Filter, Grid, load_library  # noqa: F401

__all__ = ["Amcl3dCudaError", "Context", "Filter", "Grid", "load_library"]
