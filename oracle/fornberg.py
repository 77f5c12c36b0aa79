"""Oracle restatement of Fornberg (1988) finite-difference weights.

Follows /root/reference/src/discretization/schemes/fornberg_calculate_weights.jl:20-67
operation-for-operation (same loop nest, same update order, same "sum-to-zero"
fix on element N÷2+1 at :62-65) so the float64 results are bit-identical to
the reference for the same inputs.  Test infrastructure only (see oracle/__init__).
"""
import numpy as np


def calculate_weights(order, x0, x):
    x = [float(v) for v in x]
    x0 = float(x0)
    N = len(x)
    assert order < N, "Not enough points for the requested order."
    M = order
    c1 = 1.0
    c4 = x[0] - x0
    C = np.zeros((N, M + 1))
    C[0, 0] = 1.0
    for i in range(1, N):
        mn = min(i, M)
        c2 = 1.0
        c5 = c4
        c4 = x[i] - x0
        for j in range(0, i):
            c3 = x[i] - x[j]
            c2 *= c3
            if j == i - 1:
                for s in range(mn, 0, -1):
                    C[i, s] = c1 * (s * C[i - 1, s - 1] - c5 * C[i - 1, s]) / c2
                C[i, 0] = -c1 * c5 * C[i - 1, 0] / c2
            for s in range(mn, 0, -1):
                C[j, s] = (c4 * C[j, s] - s * C[j, s - 1]) / c3
            C[j, 0] = c4 * C[j, 0] / c3
        c1 = c2
    w = C[:, M].copy()
    if order != 0:
        # Julia's sum() on a Vector{Float64} of length < 16 is a plain left fold.
        s = 0.0
        for v in w:
            s += v
        w[N // 2] -= s
    return w
