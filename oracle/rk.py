"""Oracle explicit Runge-Kutta integrators (NumPy).  Test infrastructure only.

The reference delegates time stepping to OrdinaryDiffEq.jl (third-party, compat ">= 7",
Project.toml:33; not vendored).  Restated from the published methods:
  * Tsit5  — Ch. Tsitouras, "Runge-Kutta pairs of order 5(4) satisfying only the first column
             simplifying assumption", Comput. Math. Appl. 62 (2011); FSAL, 7 stages.
  * error norm / PI controller — OrdinaryDiffEq defaults: EEst =
             sqrt(mean((utilde/(abstol+max(|u|,|u+|)*reltol))^2)), beta1=7/50, beta2=2/25,
             gamma=0.9, qmin=0.2, qmax=10, qoldinit=1e-4.
  * saveat — dense output inside the covering step, never a clipped step (OrdinaryDiffEq's behaviour; t1 alone is a
             stop time): Tsit5's free 4th-order interpolant, cubic Hermite for the methods without one.
  * SSPRK33 (Shu-Osher), Euler, RK4 with fixed dt as used by the reference tests
             (test/Convection_WENO/MOL_1D_Linear_Convection_WENO.jl:45, test/Convection/...:45).
Acceptance (north_star): final-time agreement within abstol/reltol, not identical step sequences.
"""
import numpy as np

TSIT5_C = [0.0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0]
TSIT5_A = [
    [],
    [0.161],
    [-0.008480655492356989, 0.335480655492357],
    [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
    [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
    [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
    [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774],
]
TSIT5_BTILDE = [-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995,
                -0.1447110071732629, 0.5823571654525552, -0.45808210592918697, 0.015151515151515152]


def _norm(x):
    return float(np.sqrt(np.mean(x * x))) if x.size else 0.0


def initial_dt(f, u0, t0, tdir, abstol, reltol, order=5):
    """Hairer-Norsett-Wanner starting step as used by OrdinaryDiffEq (2 RHS calls)."""
    sk = abstol + np.abs(u0) * reltol
    d0 = _norm(u0 / sk)
    f0 = f(u0, t0)
    d1 = _norm(f0 / sk)
    dt0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * d0 / d1
    u1 = u0 + tdir * dt0 * f0
    f1 = f(u1, t0 + tdir * dt0)
    d2 = _norm((f1 - f0) / sk) / dt0
    m = max(d1, d2)
    dt1 = max(1e-6, dt0 * 1e-3) if m <= 1e-15 else 10.0 ** (-(2 + np.log10(m)) / order)
    return min(100 * dt0, dt1)


def tsit5_dense_weights(th):
    """b_i(theta) of the free 4th-order Tsit5 interpolant (Tsitouras 2011; OrdinaryDiffEq's dense output for Tsit5)."""
    th2 = th * th
    return [-1.0530884977290216 * th * (th - 1.3299890189751412) * (th2 - 1.4364028541716351 * th + 0.7139816917074209),
            0.1017 * th2 * (th2 - 2.1966568338249754 * th + 1.2949852507374631),
            2.490627285651252793 * th2 * (th2 - 2.38535645472061657 * th + 1.57803468208092486),
            -16.54810288924490272 * (th - 1.21712927295533244) * (th - 0.61620406037800089) * th2,
            47.37952196281928122 * (th - 1.203071208372362603) * (th - 0.658047292653547382) * th2,
            -34.87065786149660974 * (th - 1.2) * (th - 0.666666666666666667) * th2,
            2.5 * (th - 1.0) * (th - 0.6) * th2]


def _save_list(saveat, t0, t1):
    if saveat is None:
        return [float(t1)]
    if np.isscalar(saveat):
        return [float(v) for v in np.arange(t0, t1 + 0.5 * saveat, saveat) if v <= t1 + 1e-12]
    return [float(v) for v in np.atleast_1d(saveat)]


def solve_tsit5(f, u0, tspan, abstol=1e-6, reltol=1e-3, saveat=None, dt=None, maxiters=10 ** 6):
    """Adaptive Tsit5 as OrdinaryDiffEq runs it: t1 is a stop time, save points are not -- states at `saveat` come from
    the method's dense output inside the step that covers them, so saveat does not change the step sequence.
    The PI controller's "steady" dead band is [1, 1] for explicit algorithms, i.e. empty (6/5 belongs to the implicit ones)."""
    t0, t1 = tspan
    u = np.array(u0, dtype=float)
    t = float(t0)
    ttol = 1e-14 * max(1.0, abs(t0), abs(t1))
    saves = _save_list(saveat, t0, t1)
    ts, us = [], []
    while saves and saves[0] <= t0 + ttol:
        ts.append(saves[0]); us.append(u.copy()); saves = saves[1:]
    dt = initial_dt(f, u, t, 1.0, abstol, reltol) if dt is None else dt
    qold = 1e-4
    k = [None] * 7
    k[0] = f(u, t)
    stats = dict(nf=3, naccept=0, nreject=0)
    it = 0
    while t < t1 and it < maxiters:
        it += 1
        dtu = min(dt, t1 - t)
        for s in range(1, 7):
            tmp = u.copy()
            for j in range(s):
                tmp += dtu * TSIT5_A[s][j] * k[j]
            if s < 6:
                k[s] = f(tmp, t + TSIT5_C[s] * dtu)
        unew = tmp
        k[6] = f(unew, t + dtu)
        stats["nf"] += 6
        utilde = dtu * sum(bt * kk for bt, kk in zip(TSIT5_BTILDE, k))
        EEst = _norm(utilde / (abstol + np.maximum(np.abs(u), np.abs(unew)) * reltol))
        if EEst <= 1.0:
            q = max(1 / 10.0, min(1 / 0.2, EEst ** (7 / 50) / qold ** (2 / 25) / 0.9)) if EEst > 0 else 1 / 10.0
            qold = max(EEst, 1e-4)
            tnew = t1 if abs((t + dtu) - t1) <= ttol else t + dtu
            while saves and saves[0] <= tnew + ttol:
                if abs(saves[0] - tnew) <= ttol:
                    us.append(unew.copy())
                else:
                    b = tsit5_dense_weights((saves[0] - t) / dtu)
                    us.append(u + dtu * sum(bi * ki for bi, ki in zip(b, k)))
                ts.append(saves[0]); saves = saves[1:]
            clipped = dtu < dt
            t = tnew
            u = unew
            k[0] = k[6]
            stats["naccept"] += 1
            if not clipped or t < t1:
                dt = dtu / q
        else:
            stats["nreject"] += 1
            dt = dtu / min(1 / 0.2, EEst ** (7 / 50) / 0.9)
    return np.array(ts), us, stats


def _step_fixed(f, u, t, dt, alg):
    """One fixed step; returns (u1, f(u, t))."""
    f0 = f(u, t)
    if alg == "euler":
        return u + dt * f0, f0
    if alg == "ssprk33":
        u1 = u + dt * f0
        u2 = 0.75 * u + 0.25 * (u1 + dt * f(u1, t + dt))
        return u / 3 + (2 / 3) * (u2 + dt * f(u2, t + dt / 2)), f0
    if alg == "rk4":
        k2 = f(u + dt / 2 * f0, t + dt / 2)
        k3 = f(u + dt / 2 * k2, t + dt / 2); k4 = f(u + dt * k3, t + dt)
        return u + dt / 6 * (f0 + 2 * k2 + 2 * k3 + k4), f0
    raise ValueError(alg)


def solve_fixed(f, u0, tspan, dt, alg="ssprk33", saveat=None):
    """Fixed-step integration; the last step is shortened to land on t1.  saveat=None saves every step; otherwise the
    states at the save points, by dense output inside the covering step (Tsit5: its interpolant; the others: cubic
    Hermite between the step's end points, as OrdinaryDiffEq does for methods without their own interpolant)."""
    t0, t1 = tspan
    u = np.array(u0, dtype=float)
    t = float(t0)
    ttol = 1e-14 * max(1.0, abs(t0), abs(t1))
    nsteps = max(0, int(np.ceil((t1 - t0) / dt - 1e-9)))
    every = saveat is None
    saves = [] if every else _save_list(saveat, t0, t1)
    ts, us = [], []
    if every:
        ts.append(t); us.append(u.copy())
    while saves and saves[0] <= t0 + ttol:
        ts.append(saves[0]); us.append(u.copy()); saves = saves[1:]
    for n in range(nsteps):
        tnew = t1 if n == nsteps - 1 else t0 + (n + 1) * dt
        h = tnew - t
        if alg == "tsit5":        # fixed-step Tsit5 (adaptive=false): the 5th-order solution of the pair
            ks = [f(u, t)]
            for s_ in range(1, 6):
                us_ = u + h * sum(a * k for a, k in zip(TSIT5_A[s_], ks))
                ks.append(f(us_, t + TSIT5_C[s_] * h))
            u1 = u + h * sum(a * k for a, k in zip(TSIT5_A[6], ks))
            inside = [s for s in saves if s < tnew - ttol]
            if inside:
                ks.append(f(u1, tnew))
                for sv in inside:
                    b = tsit5_dense_weights((sv - t) / h)
                    ts.append(sv); us.append(u + h * sum(bi * ki for bi, ki in zip(b, ks)))
                saves = saves[len(inside):]
        else:
            u1, f0 = _step_fixed(f, u, t, h, alg)
            inside = [s for s in saves if s < tnew - ttol]
            if inside:
                f1 = f(u1, tnew)
                for sv in inside:
                    th = (sv - t) / h
                    w = th * (th - 1.0)
                    ts.append(sv)
                    us.append((1 - th) * u + th * u1 + w * ((1 - 2 * th) * (u1 - u) + (th - 1) * h * f0 + th * h * f1))
                saves = saves[len(inside):]
        u, t = u1, tnew
        if every:
            ts.append(t); us.append(u.copy())
        while saves and saves[0] <= t + ttol:
            ts.append(saves[0]); us.append(u.copy()); saves = saves[1:]
    return np.array(ts), us
