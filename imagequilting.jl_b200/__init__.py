"""B200-native overlap-distance search for image quilting: Python mirror of the
ImageQuilting.jl v1.3.1 API (iqsim, voxelreuse) on top of libiqb200.so (sm_100a CUDA
kernels behind the C ABI of include/iqb200.h).  No CPU fallback."""
from .api import iqsim, voxelreuse, voxelreuse_sweep, IQ, SearchContext, graphcut, geometry  # noqa: F401
from . import _lib  # noqa: F401

__all__ = ["iqsim", "voxelreuse", "voxelreuse_sweep", "IQ", "SearchContext", "graphcut", "geometry"]
