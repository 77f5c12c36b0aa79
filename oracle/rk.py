"""Oracle explicit Runge-Kutta integrators (NumPy).  Test infrastructure only.

The reference delegates time stepping to OrdinaryDiffEq.jl (third-party, compat ">= 7",
Project.toml:33; not vendored).  Restated from the published methods:
  * Tsit5  — Ch. Tsitouras, "Runge-Kutta pairs of order 5(4) satisfying only the first column
             simplifying assumption", Comput. Math. Appl. 62 (2011); FSAL, 7 stages.
  * error norm / PI controller — OrdinaryDiffEq defaults: EEst =
             sqrt(mean((utilde/(abstol+max(|u|,|u+|)*reltol))^2)), beta1=7/50, beta2=2/25,
             gamma=0.9, qmin=0.2, qmax=10, qoldinit=1e-4.
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


def solve_tsit5(f, u0, tspan, abstol=1e-6, reltol=1e-3, saveat=None, dt=None, maxiters=10 ** 6):
    """Adaptive Tsit5; save points are hit exactly by clipping steps (last bullet)."""
    t0, t1 = tspan
    u = np.array(u0, dtype=float)
    t = float(t0)
    saves = [float(t1)] if saveat is None else [s for s in np.atleast_1d(saveat)]
    if np.isscalar(saveat) and saveat is not None:
        saves = list(np.arange(t0, t1 + 0.5 * saveat, saveat))
    ts, us = [], []
    if saves and abs(saves[0] - t0) < 1e-14:
        ts.append(t0); us.append(u.copy()); saves = saves[1:]
    dt = initial_dt(f, u, t, 1.0, abstol, reltol) if dt is None else dt
    qold = 1e-4
    k = [None] * 7
    k[0] = f(u, t)
    stats = dict(nf=3, naccept=0, nreject=0)
    it = 0
    while saves and it < maxiters:
        it += 1
        target = saves[0]
        dtu = min(dt, target - t)
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
            t = t + dtu
            u = unew
            k[0] = k[6]
            stats["naccept"] += 1
            if dtu == dt or t < target:
                dt = dtu / q
            if abs(t - target) <= 1e-14 * max(1.0, abs(target)):
                t = target
                ts.append(t); us.append(u.copy()); saves = saves[1:]
        else:
            stats["nreject"] += 1
            dt = dtu / min(1 / 0.2, EEst ** (7 / 50) / 0.9)
    return np.array(ts), us, stats


def solve_fixed(f, u0, tspan, dt, alg="ssprk33", saveat=None):
    t0, t1 = tspan
    u = np.array(u0, dtype=float)
    t = float(t0)
    nsteps = int(round((t1 - t0) / dt))
    save_every = None
    ts, us = [t], [u.copy()]
    for n in range(nsteps):
        if alg == "euler":
            u = u + dt * f(u, t)
        elif alg == "ssprk33":
            u1 = u + dt * f(u, t)
            u2 = 0.75 * u + 0.25 * (u1 + dt * f(u1, t + dt))
            u = u / 3 + (2 / 3) * (u2 + dt * f(u2, t + dt / 2))
        elif alg == "rk4":
            k1 = f(u, t); k2 = f(u + dt / 2 * k1, t + dt / 2)
            k3 = f(u + dt / 2 * k2, t + dt / 2); k4 = f(u + dt * k3, t + dt)
            u = u + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        elif alg == "tsit5":        # fixed-step Tsit5 (adaptive=false): the 5th-order solution of the pair
            ks = [f(u, t)]
            for s_ in range(1, 6):
                us_ = u + dt * sum(a * k for a, k in zip(TSIT5_A[s_], ks))
                ks.append(f(us_, t + TSIT5_C[s_] * dt))
            u = u + dt * sum(a * k for a, k in zip(TSIT5_A[6], ks))
        else:
            raise ValueError(alg)
        t = t0 + (n + 1) * dt
        ts.append(t); us.append(u.copy())
    return np.array(ts), us
