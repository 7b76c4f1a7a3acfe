"""cmc_fluid_solver_b200 - B200-native implicit (ADI) time-stepping path of cmc-fluid-solver.

Host-side mirror of the reference's solver interface (FluidSolver3D::Solver3D / AdiSolver3D,
reference src/FluidSolver3D/Solver3D.h:24-49, AdiSolver3D.h:52-61) over the C ABI of
``include/cmc_adi.h`` (``libcmcadi.so``: hand-written sm_100a CUDA kernels).  There is no CPU
fallback: creating a solver without the CUDA library or without a GPU raises.
"""
from .cases import Case  # noqa: F401
from .solver import AdiSolver2D, AdiSolver3D, CmcError, DivergedError, lib_path, load_library  # noqa: F401

__all__ = ["AdiSolver2D", "AdiSolver3D", "Case", "CmcError", "DivergedError", "lib_path", "load_library"]
