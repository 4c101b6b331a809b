"""petibm_b200 -- a B200-native (sm_100a) pressure-Poisson linear solver behind PetIBM's LinSolver
interface.  The product is libb200ls.so (hand-written CUDA + C ABI, include/b200ls.h); this package is
the Python host mirror of the reference's plugin interface and the multi-GPU wiring."""
from ._lib import B200Error, LIB_PATH  # noqa: F401
from .linsolver import LinSolverB200, LinSolverBase, Mat, createLinSolver  # noqa: F401
from .mesh import Grid, axis_from_subdomains, slab_range  # noqa: F401

__all__ = ["B200Error", "LinSolverB200", "LinSolverBase", "Mat", "createLinSolver", "Grid",
           "axis_from_subdomains", "slab_range", "LIB_PATH"]
