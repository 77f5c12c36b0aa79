"""Oracle restatement of the reference's WENO5 kernels (vectorised over points).

Test infrastructure only (see oracle/__init__).
  weno_f_uniform            <- schemes/WENO/WENO.jl:6-57
  _fornberg3_weights        <- schemes/WENO/nonuniform_weno.jl:5-46
  _substencil_beta_r        <- nonuniform_weno.jl:54-73
  _weno_target_geometry     <- nonuniform_weno.jl:76-81
  _weno_ideal_d0d2          <- nonuniform_weno.jl:84-117
  weno_f_nonuniform_core    <- nonuniform_weno.jl:120-163
`u` is a sequence of 5 arrays (or scalars) u[0..4]; `x` likewise for the NU kernel.
"""
import numpy as np


def weno_f_uniform(u, eps, dx):
    u_m2, u_m1, u_0, u_p1, u_p2 = u
    gm1, gm2, gm3 = 1 / 10, 3 / 5, 3 / 10
    b1 = 13 * (u_0 - 2 * u_p1 + u_p2) ** 2 / 12 + (3 * u_0 - 4 * u_p1 + u_p2) ** 2 / 4
    b2 = 13 * (u_m1 - 2 * u_0 + u_p1) ** 2 / 12 + (u_m1 - u_p1) ** 2 / 4
    b3 = 13 * (u_m2 - 2 * u_m1 + u_0) ** 2 / 12 + (u_m2 - 4 * u_m1 + 3 * u_0) ** 2 / 4
    om1 = gm1 / (eps + b1) ** 2
    om2 = gm2 / (eps + b2) ** 2
    om3 = gm3 / (eps + b3) ** 2
    den = om1 + om2 + om3
    wm1, wm2, wm3 = om1 / den, om2 / den, om3 / den
    gp1, gp2, gp3 = 3 / 10, 3 / 5, 1 / 10
    op1 = gp1 / (eps + b1) ** 2
    op2 = gp2 / (eps + b2) ** 2
    op3 = gp3 / (eps + b3) ** 2
    denp = op1 + op2 + op3
    wp1, wp2, wp3 = op1 / denp, op2 / denp, op3 / denp
    hm1 = (11 * u_0 - 7 * u_p1 + 2 * u_p2) / 6
    hm2 = (5 * u_0 - u_p1 + 2 * u_m1) / 6
    hm3 = (2 * u_0 + 5 * u_m1 - u_m2) / 6
    hp1 = (2 * u_0 + 5 * u_p1 - u_p2) / 6
    hp2 = (5 * u_0 + 2 * u_p1 - u_m1) / 6
    hp3 = (11 * u_0 - 7 * u_m1 + 2 * u_m2) / 6
    hp = wp1 * hp1 + wp2 * hp2 + wp3 * hp3
    hm = wm1 * hm1 + wm2 * hm2 + wm3 * hm3
    return (hp - hm) / dx


def _fornberg3_weights(a0, a1, a2, xt):
    one = np.ones_like(xt * 1.0)
    zero = np.zeros_like(xt * 1.0)
    n1m0, n1m1, n1m2 = one, zero, zero
    c1 = one
    c2 = a1 - a0
    r1 = c1 / c2
    tA = a1 - xt
    a1m0 = (tA * n1m0) / c2
    a1m1 = (tA * n1m1 - n1m0) / c2
    a1m2 = (tA * n1m2 - 2 * n1m1) / c2
    sB = a0 - xt
    a2m0 = r1 * (-(sB) * n1m0)
    a2m1 = r1 * (n1m0 - sB * n1m1)
    a2m2 = r1 * (2 * n1m1 - sB * n1m2)
    c1 = c2
    n1m0, n1m1, n1m2 = a1m0, a1m1, a1m2
    n2m0, n2m1, n2m2 = a2m0, a2m1, a2m2
    c2 = (a2 - a0) * (a2 - a1)
    r2 = c1 / c2
    c3a = a2 - a0
    c3b = a2 - a1
    tA2 = a2 - xt
    b1m0 = (tA2 * n1m0) / c3a
    b1m1 = (tA2 * n1m1 - n1m0) / c3a
    b1m2 = (tA2 * n1m2 - 2 * n1m1) / c3a
    b2m0 = (tA2 * n2m0) / c3b
    b2m1 = (tA2 * n2m1 - n2m0) / c3b
    b2m2 = (tA2 * n2m2 - 2 * n2m1) / c3b
    sB2 = a1 - xt
    b3m0 = r2 * (-(sB2) * n2m0)
    b3m1 = r2 * (n2m0 - sB2 * n2m1)
    b3m2 = r2 * (2 * n2m1 - sB2 * n2m2)
    return (b1m0, b2m0, b3m0), (b1m1, b2m1, b3m1), (b1m2, b2m2, b3m2)


def _dot3(w, a, b, c):
    return w[0] * a + w[1] * b + w[2] * c


def _substencil_beta_r(al, ua, ub, uc, xi, xL, xM, xph, Dx):
    _, m1i, _ = _fornberg3_weights(*al, xi)
    _, m1L, _ = _fornberg3_weights(*al, xL)
    _, m1M, m2M = _fornberg3_weights(*al, xM)
    _, m1R, _ = _fornberg3_weights(*al, xph)
    r = _dot3(m1i, ua, ub, uc)
    pL = _dot3(m1L, ua, ub, uc)
    pM = _dot3(m1M, ua, ub, uc)
    pR = _dot3(m1R, ua, ub, uc)
    pp = _dot3(m2M, ua, ub, uc)
    I1 = (Dx / 6) * (pL ** 2 + 4 * pM ** 2 + pR ** 2)
    I2 = Dx * pp ** 2
    val = Dx * I1 + Dx ** 3 * I2
    return np.maximum(val, 0.0), r


def _target_geometry(T, x1, x2, x3, x4, x5):
    if T == 1:
        return x1, x1, (x1 + x2) / 2
    if T == 2:
        return x2, (x1 + x2) / 2, (x2 + x3) / 2
    if T == 3:
        return x3, (x2 + x3) / 2, (x3 + x4) / 2
    if T == 4:
        return x4, (x3 + x4) / 2, (x4 + x5) / 2
    return x5, (x4 + x5) / 2, x5


def _ideal_d0d2(T, x1, x2, x3, x4, x5):
    if T == 3:
        d0 = ((x3 - x4) * (x3 - x5)) / ((x1 - x4) * (x1 - x5))
        d2 = ((x3 - x1) * (x3 - x2)) / ((x5 - x1) * (x5 - x2))
    elif T == 1:
        d0 = ((2 * x1 - x2 - x3) * (x1 - x4) * (x1 - x5) + (x1 - x3) * (x1 - x5) * (x1 - x2)
              + (x1 - x3) * (x1 - x4) * (x1 - x2)) / ((2 * x1 - x2 - x3) * (x1 - x4) * (x1 - x5))
        d2 = ((x1 - x3) * (x1 - x4) * (x1 - x2)) / ((-x1 + x5) * (2 * x1 - x3 - x4) * (-x2 + x5))
    elif T == 2:
        d0 = ((x2 - x4) * (x2 - x5)) / ((x1 - x4) * (x1 - x5))
        d2 = ((-x1 + x2) * (x2 - x3) * (x2 - x4)) / ((-x1 + x5) * (2 * x2 - x3 - x4) * (-x2 + x5))
    elif T == 4:
        d0 = ((-x2 + x4) * (-x3 + x4) * (x4 - x5)) / ((x1 - x4) * (x1 - x5) * (-x2 - x3 + 2 * x4))
        d2 = ((-x1 + x4) * (-x2 + x4)) / ((-x1 + x5) * (-x2 + x5))
    else:
        d0 = ((-x2 + x5) * (-x3 + x5) * (-x4 + x5)) / ((x1 - x4) * (x1 - x5) * (-x2 - x3 + 2 * x5))
        d2 = ((-x1 - x4 + 2 * x5) * (-x2 + x5) * (-x3 + x5)
              + (-x1 + x5) * (-x2 - x3 + 2 * x5) * (-x4 + x5)) / ((-x1 + x5) * (-x2 + x5) * (-x3 - x4 + 2 * x5))
    return d0, d2


def weno_f_nonuniform_core(u, eps, x, T=3):
    x1, x2, x3, x4, x5 = [np.asarray(v, dtype=float) for v in x]
    u1, u2, u3, u4, u5 = u
    theta = 3.0
    half = 0.5
    xi, xL, xph = _target_geometry(T, x1, x2, x3, x4, x5)
    Dx = xph - xL
    xM = (xL + xph) / 2
    b0, r0 = _substencil_beta_r((x1, x2, x3), u1, u2, u3, xi, xL, xM, xph, Dx)
    b1, r1 = _substencil_beta_r((x2, x3, x4), u2, u3, u4, xi, xL, xM, xph, Dx)
    b2, r2 = _substencil_beta_r((x3, x4, x5), u3, u4, u5, xi, xL, xM, xph, Dx)
    d0, d2 = _ideal_d0d2(T, x1, x2, x3, x4, x5)
    d1 = 1.0 - d0 - d2
    dp0 = half * (d0 + theta * np.abs(d0))
    dp1 = half * (d1 + theta * np.abs(d1))
    dp2 = half * (d2 + theta * np.abs(d2))
    dm0, dm1, dm2 = dp0 - d0, dp1 - d1, dp2 - d2
    sp = dp0 + dp1 + dp2
    sm = dm0 + dm1 + dm2
    ap0 = (dp0 / sp) / (eps + b0) ** 2
    ap1 = (dp1 / sp) / (eps + b1) ** 2
    ap2 = (dp2 / sp) / (eps + b2) ** 2
    s_p = ap0 + ap1 + ap2
    wp0, wp1, wp2 = ap0 / s_p, ap1 / s_p, ap2 / s_p
    am0 = (dm0 / sm) / (eps + b0) ** 2
    am1 = (dm1 / sm) / (eps + b1) ** 2
    am2 = (dm2 / sm) / (eps + b2) ** 2
    s_m = am0 + am1 + am2
    wm0, wm1, wm2 = am0 / s_m, am1 / s_m, am2 / s_m
    Rp = wp0 * r0 + wp1 * r1 + wp2 * r2
    Rm = wm0 * r0 + wm1 * r1 + wm2 * r2
    return sp * Rp - sm * Rm
