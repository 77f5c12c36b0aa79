"""Non-uniform WENO5: the product's arithmetic against 50-digit evaluation of nonuniform_weno.jl:120-163.

The kernels do not repeat the reference's operation order (12 three-point Fornberg solves at cell midpoints per
point): the u-independent part is built at plan time and the u-dependent part works on divided differences over exact
node spacings (kernels/mol_device.cuh, csrc/mol_parse.cpp weno_nu_tables).  Both are evaluations of the same formula;
on a clustered grid the reference's own arithmetic carries a rounding error of relative size eps * |x| / h (midpoints
(x_a + x_b)/2 and differences against them), which is what bounds the agreement between the two.  This test pins
that reading: per node, the emulated kernel is closer to the exact value than the restated reference is."""
import mpmath as mp
import numpy as np
import pytest

import _mol_import  # noqa: F401
import mol_b200
from mol_b200 import capi
import problems as examples
from oracle import weno as oweno
from cuda_emu import EmuKernel


def exact_weno(u, x, eps, T):
    """nonuniform_weno.jl:120-163 in 50-digit arithmetic (same formula: Simpson smoothness indicators on the quadratic
    sub-stencil interpolants, closed-form-equivalent ideal weights, theta = 3 splitting)."""
    mp.mp.dps = 50
    u = [mp.mpf(float(v)) for v in u]
    x = [mp.mpf(float(v)) for v in x]
    eps = mp.mpf(eps)
    k = T - 1
    xi = x[k]
    xL = x[0] if k == 0 else (x[k - 1] + x[k]) / 2
    xR = x[4] if k == 4 else (x[k] + x[k + 1]) / 2
    dx, xM = xR - xL, (xL + xR) / 2

    def lag1(s, j, xt):
        den = mp.mpf(1)
        for l in range(len(s)):
            if l != j:
                den *= s[j] - s[l]
        num = mp.mpf(0)
        for m in range(len(s)):
            if m == j:
                continue
            p = mp.mpf(1)
            for l in range(len(s)):
                if l not in (j, m):
                    p *= xt - s[l]
            num += p
        return num / den

    r, beta = [], []
    for kk in range(3):
        s, us = x[kk:kk + 3], u[kk:kk + 3]
        d1 = lambda xt: sum(lag1(s, j, xt) * us[j] for j in range(3))
        pp = sum(2 * us[j] / ((s[j] - s[(j + 1) % 3]) * (s[j] - s[(j + 2) % 3])) for j in range(3))
        r.append(d1(xi))
        I1 = dx / 6 * (d1(xL) ** 2 + 4 * d1(xM) ** 2 + d1(xR) ** 2)
        beta.append(max(dx * I1 + dx ** 3 * (dx * pp ** 2), mp.mpf(0)))
    d0 = lag1(x, 0, xi) / lag1(x[0:3], 0, xi)
    d2 = lag1(x, 4, xi) / lag1(x[2:5], 2, xi)
    d = [d0, 1 - d0 - d2, d2]
    dp = [(v + 3 * abs(v)) / 2 for v in d]
    dm = [a - b for a, b in zip(dp, d)]
    sp, sm = sum(dp), sum(dm)
    ap = [(dp[i] / sp) / (eps + beta[i]) ** 2 for i in range(3)]
    am = [(dm[i] / sm) / (eps + beta[i]) ** 2 for i in range(3)]
    Rp = sum(ap[i] / sum(ap) * r[i] for i in range(3))
    Rm = sum(am[i] / sum(am) * r[i] for i in range(3))
    return sp * Rp - sm * Rm


@pytest.mark.parametrize("T", [1, 2, 3, 4, 5])
def test_exact_formula_agrees_with_the_restated_reference_on_a_benign_stencil(T):
    """Guards the 50-digit restatement itself: on an O(1)-spaced stencil the reference arithmetic is accurate to ~1e-15,
    including the wall targets whose ideal weights the reference writes as closed forms (nonuniform_weno.jl:76-117)."""
    x = np.array([0.0, 0.11, 0.23, 0.42, 0.55])          # benchmark/weno/suite.jl:20
    u = np.array([1.3, 2.1, 1.7, 0.4, 0.9])
    want = float(oweno.weno_f_nonuniform_core(u, 1e-6, x, T))
    assert abs(float(exact_weno(u, x, 1e-6, T)) - want) <= 1e-13 * abs(want)


def test_nu_weno_is_closer_to_exact_than_the_reference_arithmetic():
    n = 400
    xi = np.linspace(0.0, 1.0, n)
    grid = 50.0 + 2.0 * (xi + 0.12 * np.sin(2 * np.pi * xi) / (2 * np.pi))          # |x| / h ~ 1e4, smooth 0.88..1.12 stretching
    sys_, disc = examples.advection_dirichlet_nu(grid, v=1.0, scheme=mol_b200.WENOScheme())
    prog = mol_b200.symbolic_discretize(sys_, disc)
    plan = capi.Plan(prog.text, device=-1)
    from oracle.discretize import OracleProblem
    orc = OracleProblem(sys_, disc)
    rng = np.random.default_rng(1)
    u = orc.u0 + 0.05 * rng.standard_normal(orc.nstate)
    got = EmuKernel(plan, prog).rhs([u], [1.0], 0.0)
    ref = orc.rhs(u, 0.0)
    full = np.asarray(orc.full_state(u, 0.0)[0]).reshape(-1)
    lo = prog.ilo[0][0]
    ex = np.zeros(orc.nstate)
    for k in range(orc.nstate):
        node = lo + k                                   # 1-based
        if node <= 2:
            s0, T = 1, node
        elif node > n - 2:
            s0, T = n - 4, 5 - (n - node)
        else:
            s0, T = node - 2, 3
        ex[k] = -float(exact_weno(full[s0 - 1:s0 + 4], grid[s0 - 1:s0 + 4], 1e-6, T))
    scale = np.max(np.abs(ex))
    err_kernel = np.max(np.abs(got - ex)) / scale
    err_oracle = np.max(np.abs(ref - ex)) / scale
    assert err_kernel <= 2e-14, err_kernel                              # a few ulps of the largest term
    assert err_oracle >= 5 * err_kernel, (err_oracle, err_kernel)       # the reference arithmetic is the noisier one
    # ... so the distance between the two is the reference's rounding: a few eps * |x| / h (1.1e-12 on this grid, which
    # is why the parity bar of the non-uniform WENO cases is 1e-12 of max |du| on O(1) domains and not tighter)
    cond = np.finfo(float).eps * np.max(np.abs(grid)) / np.min(np.diff(grid))
    assert np.max(np.abs(got - ref)) / scale <= 4 * cond, (np.max(np.abs(got - ref)) / scale, cond)
    plan.close()
