"""mol_b200 — B200-native backend for MethodOfLines.jl's hot path (RHS evaluation + explicit RK).

Import as `mol_b200` (see /_mol_import.py).  Public surface mirrors the reference's:
PDESystem / MOLFiniteDifference / discretize / solve, with the compute in libmol_cuda.so.
"""
from .interface import (Eq, Equation, Differential, Interval, PDESystem, ifelse,
                        UpwindScheme, WENOScheme, MOLFiniteDifference,
                        CudaStencilDiscretization, center_align, edge_align)
from .lowering import StencilLoweringError, lower
from .problem import (discretize, symbolic_discretize, solve, ODEProblem, ODESolution,
                      Tsit5, SSPRK33, Euler, RK4)
from . import capi
