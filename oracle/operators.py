"""Oracle restatement of the reference's DerivativeOperator tables.

Test infrastructure only (see oracle/__init__).  Each constructor cites the
reference function it follows; field names mirror
/root/reference/src/discretization/derivative_operator.jl:1-15.
All indices stored here are 0-based Python lists of rows; callers translate the
reference's 1-based node numbers.
"""
from dataclasses import dataclass, field
from typing import Any, List

import numpy as np

from .fornberg import calculate_weights


@dataclass
class DerivativeOperator:
    derivative_order: int
    approximation_order: int
    dx: Any                      # float (uniform) or np.ndarray of spacings (non-uniform)
    stencil_length: int
    stencil_coefs: Any           # np.ndarray (uniform) or list of np.ndarray per node (NU)
    boundary_stencil_length: int
    boundary_point_count: int
    low_boundary_coefs: List[np.ndarray]
    high_boundary_coefs: List[np.ndarray]
    offside: int = 0
    uniform: bool = True


def _isuniform(dx):
    return np.isscalar(dx)


def centered(d, p, dx):
    """CompleteCenteredDifference — centered_diff_weights.jl:4-75 (uniform), :77-153 (NU)."""
    assert p > 1
    L = d + p - 1 + (d + p) % 2
    bsl = d + p
    bpc = L // 2
    if _isuniform(dx):
        dx = float(dx)
        scale = 1.0 / dx ** d
        dummy = list(range(-(L // 2), L // 2 + 1))
        coefs = scale * calculate_weights(d, 0.0, dummy)
        lbx = list(range(0, bsl))
        low = [scale * calculate_weights(d, float(x0), lbx) for x0 in lbx[:bpc]]
        high = [w[::-1] * (-1.0) ** d for w in low]
        return DerivativeOperator(d, p, dx, L, coefs, bsl, bpc, low, high, 0, True)
    x = np.asarray(dx, dtype=float)          # NU: the "dx" argument is the node vector
    n = len(x)
    dxs = np.array([x[i + 1] - x[i] for i in range(n - 1)])
    low_x = np.concatenate([[0.0], np.cumsum(dxs[: bsl - 1])])
    high_x = np.cumsum(dxs[n - 1 - bsl:])
    # interior_x = (bpc+1):(len-bpc) (1-based)
    coefs = [calculate_weights(d, x[i], x[i - bpc: i + bpc + 1]) for i in range(bpc, n - bpc)]
    low = [calculate_weights(d, low_x[i], low_x) for i in range(bpc)]
    high = [calculate_weights(d, high_x[len(high_x) - 1 - i], high_x) for i in range(bpc)]
    return DerivativeOperator(d, p, dxs, L, coefs, bsl, bpc, low, high, 0, False)


def upwind(d, p, dx, offside=0):
    """CompleteUpwindDifference — upwind_diff_weights.jl:4-91 (uniform), :93-171 (NU)."""
    assert offside > -1
    L = d + p
    bsl = d + p
    low_bpc = offside
    high_bpc = L - 1 - offside
    if _isuniform(dx):
        dx = float(dx)
        scale = 1.0 / dx ** d
        dummy = [0.0 - offside + k for k in range(L)]
        coefs = scale * calculate_weights(d, 0.0, dummy)
        lbx = [float(k) for k in range(bsl)]
        low = [scale * calculate_weights(d, float(x0), lbx) for x0 in range(low_bpc)]
        hbx = [-float(k) for k in range(bsl)]
        hscale = (-1.0 / dx) ** d
        _high = [hscale * calculate_weights(d, -float(k), hbx) for k in range(high_bpc)]
        high = _high[::-1]
        return DerivativeOperator(d, p, dx, L, coefs, bsl, high_bpc, low, high, offside, True)
    x = np.asarray(dx, dtype=float)
    n = len(x)
    assert offside <= L - 1
    dxs = np.array([x[i + 1] - x[i] for i in range(n - 1)])
    low_x = x[:bsl]
    high_x = x[n - bsl:]
    # i in (low_bpc+1):(n-high_bpc), taps x[i-offside : i+L-1-offside] (1-based)
    coefs = [calculate_weights(d, x[i], x[i - offside: i + L - offside])
             for i in range(low_bpc, n - high_bpc)]
    low = [calculate_weights(d, x0, low_x) for x0 in x[:low_bpc]]
    high = [calculate_weights(d, x0, high_x) for x0 in x[n - high_bpc:]]
    # NB reference quirk (upwind_diff_weights.jl:154): struct field offside is reset to 0
    return DerivativeOperator(d, p, dxs, L, coefs, bsl, high_bpc, low, high, 0, False)


def half_centered(d, p, dx):
    """CompleteHalfCenteredDifference — half_offset_weights.jl:4-70 (uniform), :72-150 (NU)."""
    assert p > 1
    L = p + 2 * (d // 2) + (p % 2)
    bsl = d + p
    endpoint = L // 2
    bpc = L // 2
    if _isuniform(dx):
        dx = float(dx)
        scale = 1.0 / dx ** d
        dummy = list(range(1 - endpoint, endpoint + 1))
        coefs = scale * calculate_weights(d, 0.5, dummy)
        lbx = list(range(1, bsl + 1))
        low = [scale * calculate_weights(d, 1.5 + k, lbx) for k in range(bpc)]
        high = [w[::-1] * (-1.0) ** d for w in low]
        return DerivativeOperator(d, p, dx, L, coefs, bsl, bpc, low, high, 0, True)
    x = np.asarray(dx, dtype=float)
    n = len(x)
    hx = np.array([(x[i] + x[i + 1]) / 2 for i in range(n - 1)])
    dxs = np.array([x[i + 1] - x[i] for i in range(n - 1)])
    low_x = x[:bsl]
    high_x = x[n - bsl:]
    # i in (endpoint+1):(n-endpoint): taps x[i-endpoint+1 : i+endpoint] at hx[i] (1-based)
    coefs = [calculate_weights(d, hx[i], x[i - endpoint + 1: i + endpoint + 1])
             for i in range(endpoint, n - endpoint)]
    low = [calculate_weights(d, hx[k], low_x) for k in range(bpc)]
    high = [calculate_weights(d, hx[len(hx) - 1 - k], high_x) for k in range(bpc)]
    return DerivativeOperator(d, p, dxs, L, coefs, bsl, bpc, low, high, 0, False)


def _insert(w, pos, val):
    return np.concatenate([w[:pos], [val], w[pos:]])


def extrapolator(p, dx):
    """BoundaryInterpolatorExtrapolator — extrapolation_weights.jl:1-72 (uniform), :74-175 (NU)."""
    assert p > 1
    L = p - 1 + p % 2
    bsl = p
    bpc = L // 2
    if _isuniform(dx):
        dx = float(dx)
        dummy = list(range(-(L // 2), L // 2 + 1))
        rem = [v for v in dummy if v != 0]
        coefs = _insert(calculate_weights(0, 0.0, rem), dummy.index(0), 0.0)
        lbx = list(range(bsl))
        low = []
        for i, x0 in enumerate(lbx[:bpc]):
            rem = [v for v in lbx if v != x0]
            low.append(_insert(calculate_weights(0, float(x0), rem), i, 0.0))
        high = [w[::-1] for w in low]
        return DerivativeOperator(0, p, dx, L, coefs, bsl, bpc, low, high, 0, True)
    x = np.asarray(dx, dtype=float)
    n = len(x)
    endpoint = bpc
    midpoint = L // 2 + L % 2          # 1-based insert position
    dxs = np.array([x[i + 1] - x[i] for i in range(n - 1)])
    low_x = np.concatenate([[0.0], np.cumsum(dxs[: bsl - 1])])
    high_x = np.cumsum(dxs[n - 1 - bsl:])
    coefs = []
    for i in range(endpoint, n - endpoint):
        loc = list(x[i - endpoint: i + endpoint + 1])
        rem = [v for v in loc if v != x[i]]
        coefs.append(_insert(calculate_weights(0, x[i], rem), midpoint - 1, 0.0))
    low, high = [], []
    for i in range(bpc):
        rem = [v for v in low_x if v != low_x[i]]
        low.append(_insert(calculate_weights(0, low_x[i], rem), i, 0.0))
    for i in range(bpc):
        x0 = high_x[len(high_x) - 1 - i]
        rem = [v for v in high_x if v != x0]
        high.append(_insert(calculate_weights(0, x0, rem), len(high_x) - i - 1, 0.0))
    return DerivativeOperator(0, p, dxs, L, coefs, bsl, bpc, low, high, 0, False)


@dataclass
class DifferentialDiscretizer:
    """construct_differential_discretizer — differential_discretizer.jl:14-109 (one spatial var)."""
    approx_order: int
    map: dict = field(default_factory=dict)          # d -> centered op
    windpos: dict = field(default_factory=dict)      # d -> forward op  (windmap[1], offside 0)
    windneg: dict = field(default_factory=dict)      # d -> backward op (windmap[2], offside d+p-1)
    half_inner: dict = field(default_factory=dict)   # d -> half-centered op (halfoffsetmap[1])
    half_outer: Any = None                           # halfoffsetmap[2][Dx]
    interp: Any = None                               # interpmap[x]
    boundary: Any = None                             # boundary[x]
    orders: list = field(default_factory=list)


def differential_discretizer(grid, dx, orders, approx_order, upwind_order, weno):
    """grid: node vector; dx: float if uniform else None.
    `weno` True when advection_scheme is a FunctionalScheme (order-1 derivs skip upwind tables,
    differential_discretizer.jl:63-67)."""
    uniform = dx is not None
    g = float(dx) if uniform else np.asarray(grid, dtype=float)
    D = DifferentialDiscretizer(approx_order, orders=sorted(orders))
    _orders = sorted(set(list(orders) + [1, 2]))
    if uniform:
        D.half_outer = half_centered(1, approx_order, g)
    else:
        hx = np.array([(g[i + 1] + g[i]) / 2 for i in range(len(g) - 1)])
        D.half_outer = half_centered(1, approx_order, hx)
    for d in _orders:
        D.map[d] = centered(d, approx_order, g)
        D.half_inner[d] = half_centered(d, approx_order, g)
    odd = [d for d in orders if d % 2 == 1]
    if weno:
        odd = [d for d in odd if d != 1]
    for d in odd:
        D.windpos[d] = upwind(d, upwind_order, g, 0)
        D.windneg[d] = upwind(d, upwind_order, g, d + upwind_order - 1)
    D.interp = half_centered(0, max(4, approx_order), g)
    try:
        D.boundary = extrapolator(max(6, approx_order), g)
    except Exception:           # grids smaller than the extrapolation stencil
        D.boundary = None
    return D
