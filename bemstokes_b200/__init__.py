"""bemstokes_b200 — B200-native (sm_100a) implementation of the BEMStokes hot path: FP64 collocation assembly
of the Stokes single/double-layer matrices and the GMRES / direct solve of the monolithic rigid-body system.

The compute lives in libbemstokes_b200.so (hand-written CUDA, C-ABI in include/bemstokes_b200.h); this package
is the host-side mirror of the reference's BEMProblem / StokesKernel / DirectPreconditioner interface."""
from . import _lib  # raises ImportError loudly when the CUDA library is missing: there is no CPU fallback
from .mesh import QuadMesh, read_mesh, read_inp, read_msh, cubesphere, to_q2
from .problem import (BEMProblem, DirectPreconditioner, SolverControl, StokesKernel, FreeSurfaceStokesKernel,
                      NoSlipWallStokesKernel, DeviceMatrix)
from ._lib import BemStokesError

__all__ = ["BEMProblem", "DirectPreconditioner", "SolverControl", "StokesKernel", "FreeSurfaceStokesKernel",
           "NoSlipWallStokesKernel", "DeviceMatrix", "QuadMesh", "read_mesh", "read_inp", "read_msh", "cubesphere",
           "to_q2", "BemStokesError"]
