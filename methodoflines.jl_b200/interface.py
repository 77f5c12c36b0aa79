"""Host-side mirror of the reference's user-facing types (symbolic layer = SymPy).

Mirrors, name for name, what a MethodOfLines.jl user writes
(/root/reference/docs/src/tutorials/brusselator.md:47-104):

    PDESystem(eqs, bcs, domains, ivs, dvs, ps)          ModelingToolkit.PDESystem
    Differential(x)            (Differential(x)**2)      Symbolics.Differential
    Interval(x, a, b)          x ∈ Interval(a, b)        DomainSets.Interval
    MOLFiniteDifference(dxs, t; approx_order, advection_scheme, grid_align,
                        discretization_strategy, **kwargs)
                               src/interface/MOLFiniteDifference.jl:31-76
    UpwindScheme(order), WENOScheme(epsilon)             src/interface/scheme_types.jl:8-17,
                                                         src/discretization/schemes/WENO/WENO.jl:83-89
    CudaStencilDiscretization  the NEW strategy type this backend adds next to
                               Scalarized/ArrayDiscretization (src/interface/disc_strategy_types.jl:3)

Julia's `lhs ~ rhs` is spelled `Eq(lhs, rhs)` here (no binary `~` in Python).
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field
from typing import Any, Dict, List, Sequence

import sympy as sp


class Equation:
    """`lhs ~ rhs` (never auto-evaluated, unlike sympy.Eq)."""

    def __init__(self, lhs, rhs):
        self.lhs = sp.sympify(lhs)
        self.rhs = sp.sympify(rhs)

    def __repr__(self):
        return f"{self.lhs} ~ {self.rhs}"


def Eq(lhs, rhs) -> Equation:
    return Equation(lhs, rhs)


class Differential:
    """Symbolics.Differential: D = Differential(x); D(u(t,x)); (D**2)(u(t,x))."""

    def __init__(self, var, order: int = 1):
        self.var = var
        self.order = int(order)

    def __pow__(self, n):
        return Differential(self.var, self.order * int(n))

    def __call__(self, expr):
        expr = sp.sympify(expr)
        if not expr.atoms(sp.core.function.AppliedUndef):
            # a known expression of the independent variables: differentiate now (Symbolics.expand_derivatives, as the
            # reference's tests do for variable coefficients, test/Diffusion/MOL_1D_Linear_Diffusion.jl:139-141)
            return sp.diff(expr, self.var, self.order)
        return sp.Derivative(expr, (self.var, self.order))


@dataclass
class Interval:
    """`var ∈ Interval(lo, hi)`."""
    var: Any
    lo: float
    hi: float


def ifelse(cond, a, b):
    """Symbolics `ifelse` / boolean-times-number idiom of the reference tests
    (test/Brusselator/brusselator_eq.jl:19)."""
    return sp.Piecewise((sp.sympify(a), cond), (sp.sympify(b), True))


class PDESystem:
    """ModelingToolkit.PDESystem(eqs, bcs, domain, ivs, dvs, ps)."""

    def __init__(self, eqs, bcs, domains, ivs, dvs, ps=None, name="pdesys"):
        self.eqs: List[Equation] = list(eqs) if isinstance(eqs, (list, tuple)) else [eqs]
        self.bcs: List[Equation] = list(bcs) if isinstance(bcs, (list, tuple)) else [bcs]
        self.domains: List[Interval] = list(domains)
        self.ivs = list(ivs)
        self.dvs = list(dvs)
        if ps is None:
            ps = []
        if isinstance(ps, dict):
            ps = list(ps.items())
        self.ps = [(p, float(v)) for p, v in ps]       # [(symbol, default value)]
        self.name = name


# ---- schemes (src/interface/scheme_types.jl) -------------------------------------------------
@dataclass
class UpwindScheme:
    order: int = 1


@dataclass
class WENOScheme:
    epsilon: float = 1.0e-6


# ---- grid alignment (src/interface/grid_types.jl:1-33) ----------------------------------------
class CenterAlignedGrid:
    pass


class EdgeAlignedGrid:
    pass


center_align = CenterAlignedGrid()
edge_align = EdgeAlignedGrid()


# ---- discretization strategies (src/interface/disc_strategy_types.jl) -------------------------
class AbstractDiscretizationStrategy:
    pass


@dataclass
class CudaStencilDiscretization(AbstractDiscretizationStrategy):
    """The new strategy: lower each discretised PDE to a stencil program executed by
    libmol_cuda.so.  `strict=True` raises on unsupported patterns (cf.
    StrictArrayDiscretization, src/array_discretization.jl:42-59)."""
    strict: bool = True
    device: int = 0


class MOLFiniteDifference:
    """src/interface/MOLFiniteDifference.jl:46-76 (same keyword names and defaults)."""

    def __init__(self, dxs, time=None, *, approx_order=2, advection_scheme=None,
                 grid_align=center_align, discretization_strategy=None,
                 should_transform=True, **kwargs):
        if approx_order % 2 != 0:
            warnings.warn(f"Discretization approx_order must be even, rounding up to {approx_order + 1}")
        assert approx_order >= 1, "approx_order must be at least 1"
        self.dxs: Dict[Any, Any] = dict(dxs)
        self.time = time
        self.approx_order = int(approx_order)
        self.advection_scheme = advection_scheme if advection_scheme is not None else UpwindScheme()
        self.grid_align = grid_align
        self.should_transform = should_transform
        self.disc_strategy = (discretization_strategy if discretization_strategy is not None
                              else CudaStencilDiscretization())
        self.kwargs = kwargs
