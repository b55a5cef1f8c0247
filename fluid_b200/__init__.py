"""fluid_b200 -- B200 (sm_100a) implementation of the per-step hot path of
TheFellow/fluid's Go package pkg/fluid, behind that package's own API.

``Fluid`` mirrors ``fluid.Fluid`` (pkg/fluid/fluid.go:11-40) method for method;
all arithmetic runs in libfluidb200.so (hand-written CUDA, C ABI in
include/fluidb200.h).  The library is opened on first use (``New`` / any ``lib``
call) and that fails loudly if it is absent: there is no CPU path.
"""
from . import edits, presets  # noqa: F401
from ._lib import FluidError, SOLVER_EXACT, SOLVER_REDBLACK, SOLVER_REDBLACK_PRESSURE  # noqa: F401
from .fluid import Fluid, New, PARTICLE_DTYPE, ScalarField, VectorField  # noqa: F401
