"""meshfreetrixi.jl_b200 -- B200-native `rhs!` hot path for MeshfreeTrixi-style RBF-FD solvers.

The directory name is not a valid Python identifier; import it through the root-level shim `mft_b200`
(`import mft_b200`), which loads this package under that name.

Contents: csrc/ (sm_100a CUDA kernels + the C ABI of include/mft_b200.h), _lib.py (ctypes binding), api.py (host-side
mirror of the reference's interface for the path), setup_ops.py / cloud.py (setup-time code), partition.py
(space-filling-curve partition + halo plan for multi-GPU), vtk.py (VTK snapshot writer + SolutionSavingCallback).
"""
from . import _lib, cloud, setup_ops, vtk  # noqa: F401
from ._lib import MftError, load  # noqa: F401
from .api import *  # noqa: F401,F403
from .vtk import SolutionSavingCallback, trixi2vtk  # noqa: F401
from .api import (BoundaryConditionDirichlet, BoundaryConditionDoNothing, CompressibleEulerEquations2D,  # noqa: F401
                  HistoryCallback, LinearScalarAdvectionEquation2D, Point2D, PointCloudBasis, PointCloudDomain,
                  PointCloudSolver, PolyharmonicSpline, HybridGaussianPHS, RBF, RBFFDEngineCUDA, SemidiscretizationHyperbolic,
                  SourceHyperviscosityFlyer, SourceHyperviscosityTominec, SourceResidualViscosityTominec,
                  SourceTerms, SourceUpwindViscosityTominec, SourceIGR, cg_, SSPRK33, SSPRK43, PIController, ParallelPointCloudDomain,
                  boundary_condition_slip_wall, solve_adaptive, PositivityPreservingLimiterZhangShu, density, pressure,
                  calc_boundary_flux_, calc_fluxes_, compute_coefficients, rhs_, semidiscretize, solve)
