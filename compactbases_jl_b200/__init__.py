"""compactbases_jl_b200 — B200 (sm_100a) implementation of the data-parallel hot path of
CompactBases.jl behind the reference's quasi-array API.

The product is ``libcompactbases_cuda.so`` (C ABI, ``include/compactbases_cuda.h``); this
package is the host-side mirror of the reference interface for that path (Julia is not
available in this environment; INTEGRATION.md holds the Julia ``ccall`` glue).  There is no
CPU fallback: every compute call needs the CUDA library and a device.
"""
from . import _lib
from ._lib import DimensionMismatch, CudaError
from .knot_sets import ArbitraryKnotSet, ExpKnotSet, LinearKnotSet
from .quadrature import gausslegendre, gausslobatto, num_quadrature_points
from .lazy import CoulombDerivative, Derivative, QuasiDiagonal, materialize
from .operators import (BandCholesky, BandLU, BandedMatrix, BlockSkylineMatrix, ColMajor, Diagonal, FEDVRMatrix, ImplicitDerivative, LinearOperator, ShiftAndInvert,
                        SymTridiagonal, Tridiagonal, as_banded, axpby, batched_dot, factorize, inner_product, mul, norm)
from .bsplines import BSpline, CoulombPotential, RestrictedBSpline, SparseMatrixCSC, diffop, diffop_host, hamiltonian_bands
from .fedvr import FEDVR, RestrictedFEDVR
from .finite_differences import (FiniteDifferences, ImplicitFiniteDifferences, RestrictedFiniteDifferences,
                                 StaggeredFiniteDifferences)
from .densities import Density, DiagonalOperator, FunctionProduct
from .complex_ops import (ComplexColMajor, ComplexDiagonalOperator, ComplexOperator, ComplexScaledFEDVR, cmul, complex_density, complex_ldiv, complex_linear_operator,
                          complex_potential)

__all__ = [
    "ArbitraryKnotSet", "ExpKnotSet", "LinearKnotSet", "BSpline", "RestrictedBSpline", "SparseMatrixCSC",
    "CoulombPotential", "diffop", "diffop_host", "hamiltonian_bands", "FEDVR", "RestrictedFEDVR", "FiniteDifferences", "StaggeredFiniteDifferences",
    "RestrictedFiniteDifferences", "Derivative", "QuasiDiagonal", "CoulombDerivative", "materialize",
    "BandedMatrix", "BlockSkylineMatrix", "BandCholesky", "LinearOperator", "Tridiagonal", "SymTridiagonal", "Diagonal", "FEDVRMatrix", "ColMajor", "mul",
    "Density", "FunctionProduct", "complex_density", "ComplexDiagonalOperator", "DiagonalOperator", "ShiftAndInvert", "BandLU", "factorize", "batched_dot", "inner_product", "norm", "ImplicitFiniteDifferences", "ImplicitDerivative", "as_banded", "axpby", "DimensionMismatch", "CudaError", "gausslegendre", "gausslobatto",
    "num_quadrature_points",
]
