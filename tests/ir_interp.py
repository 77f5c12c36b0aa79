"""Reference interpreter of the stencil-program IR in NumPy/Python (test infrastructure, CPU only).

Executes the text IR that `lowering.lower()` emits with the semantics of the table-driven CUDA kernel
(csrc/kernels/mol_device.cuh: mol_node / mol_lin_g / mol_weno_g / ghost rules; csrc/mol_codegen.cpp, GENERIC mode), one
node at a time.  Comparing its du with the oracle's checks the *content* of a stencil program -- tables, ghost rules,
equations -- on a machine without a GPU; the GPU parity tests then only have to establish kernel == IR semantics."""
import math

import numpy as np

from oracle import weno as oweno


def _f(tok):
    return float.fromhex(tok) if "0x" in tok else float(tok)


class IRProgram:
    def __init__(self, text):
        self.tabs, self.wtabs, self.fns, self.ghosts, self.eqs = {}, {}, {}, {}, {}
        self.grid, self.uniform, self.dx = {}, {}, {}
        self.ilo, self.ihi, self.per, self.params = {}, {}, {}, {}
        for line in text.split("\n"):
            tk = line.split()
            if not tk:
                continue
            k = tk[0]
            if k == "ndim":
                self.nd = int(tk[1])
            elif k == "nvar":
                self.nv = int(tk[1])
            elif k == "param":
                self.params[int(tk[1])] = _f(tk[3])
            elif k == "grid":
                j = int(tk[1]); self.uniform[j] = tk[3] == "U"; self.dx[j] = _f(tk[4])
            elif k == "coords":
                self.grid[int(tk[1])] = np.array([_f(t) for t in tk[2:]])
            elif k == "interior":
                v = int(tk[1]); b = list(map(int, tk[2:]))
                self.ilo[v], self.ihi[v] = b[:self.nd], b[self.nd:]
            elif k == "periodic":
                self.per[int(tk[1])] = [int(t) != 0 for t in tk[2:]]
            elif k == "tab":
                self.tabs[int(tk[1])] = dict(L=int(tk[2]), nrows=int(tk[3]), first=int(tk[4]), rows={})
            elif k == "core":
                T = self.tabs[int(tk[1])]; lo, hi, off = int(tk[2]), int(tk[3]), int(tk[4])
                w = [_f(t) for t in tk[5:]]
                for idx in range(lo, hi + 1):
                    T["rows"][idx] = (idx + off, w)
            elif k == "row":
                T = self.tabs[int(tk[1])]
                T["rows"][int(tk[2])] = (int(tk[3]), [_f(t) for t in tk[5:5 + int(tk[4])]])
            elif k == "wtab":
                self.wtabs[int(tk[1])] = dict(first=int(tk[3]), rows={})
            elif k == "wcore":
                T = self.wtabs[int(tk[1])]
                for idx in range(int(tk[2]), int(tk[3]) + 1):
                    T["rows"][idx] = (idx - 2, 3)
            elif k == "wrow":
                self.wtabs[int(tk[1])]["rows"][int(tk[2])] = (int(tk[3]), int(tk[4]))
            elif k == "fn":
                self.fns[int(tk[1])] = tk[3:]
            elif k == "ghost":
                v, d, node, nt = int(tk[1]), int(tk[2]), int(tk[3]), int(tk[4])
                taps = [(int(tk[5 + 3 * q]), int(tk[6 + 3 * q]), _f(tk[7 + 3 * q])) for q in range(nt)]
                pos = 5 + 3 * nt
                self.ghosts[(v, d, node)] = (taps, tk[pos + 1:pos + 1 + int(tk[pos])])
            elif k == "ghostx":     # tap coefficients are expressions: ghostx v dim node ntaps nG G.. (var node nc coef..)*
                v, d, node, nt, ng = (int(x) for x in tk[1:6])
                pos = 6
                expr = tk[pos:pos + ng]
                pos += ng
                taps = []
                for _ in range(nt):
                    w_, nd_, nc = int(tk[pos]), int(tk[pos + 1]), int(tk[pos + 2])
                    taps.append((w_, nd_, tk[pos + 3:pos + 3 + nc]))
                    pos += 3 + nc
                self.ghosts[(v, d, node)] = (taps, expr)
            elif k == "eq":
                self.eqs[int(tk[1])] = tk[3:]
        self.n = [len(self.grid[j]) for j in range(self.nd)]
        self.ext = {v: [self.ihi[v][j] - self.ilo[v][j] + 1 for j in range(self.nd)] for v in range(self.nv)}
        self.off, o = {}, 0
        for v in range(self.nv):
            self.off[v] = o
            o += int(np.prod(self.ext[v]))
        self.nstate = o

    # ---- mol_node: state, periodic wrap or ghost rule, one dimension at a time ----------------------------------
    def node(self, v, idx):
        idx = list(idx)
        for d in range(self.nd):
            if idx[d] < self.ilo[v][d] or idx[d] > self.ihi[v][d]:
                if self.per[v][d]:
                    idx[d] += (self.n[d] - 1) if idx[d] <= 1 else -(self.n[d] - 1)
                else:
                    rule = self.ghosts.get((v, d, idx[d]))
                    if rule is None:
                        return 0.0
                    taps, expr = rule
                    r = self.rpn(expr, idx, mode="ghost")
                    for (w_, nd_, a) in taps:
                        j2 = list(idx); j2[d] = nd_
                        if isinstance(a, list):
                            a = self.rpn(a, idx, mode="ghost")
                        r += a * self.node(w_, j2)
                    return r
        flat, stride = self.off[v], 1
        for d in range(self.nd):
            flat += (idx[d] - self.ilo[v][d]) * stride
            stride *= self.ext[v][d]
        return float(self.u[flat])

    def lin(self, tab, var, dim, row_idx, idx):
        start, w = self.tabs[tab]["rows"][row_idx]
        acc = 0.0
        for k, wk in enumerate(w):
            j2 = list(idx); j2[dim] = start + k
            acc += wk * self.node(var, j2)
        return acc

    def mixed(self, tx, ty, var, dx, dy, idx):
        sx, wx = self.tabs[tx]["rows"][idx[dx]]
        sy, wy = self.tabs[ty]["rows"][idx[dy]]
        acc = 0.0
        for kx, a in enumerate(wx):
            for ky, b in enumerate(wy):
                j2 = list(idx); j2[dx] = sx + kx; j2[dy] = sy + ky
                for d in range(self.nd):              # periodic dimensions wrap first
                    if self.per[var][d] and not (self.ilo[var][d] <= j2[d] <= self.ihi[var][d]):
                        j2[d] += (self.n[d] - 1) if j2[d] <= 1 else -(self.n[d] - 1)
                outside = sum(1 for d in range(self.nd)
                              if not self.per[var][d] and not (self.ilo[var][d] <= j2[d] <= self.ihi[var][d]))
                if outside < 2:                       # corner nodes are 0 (generate_bc_eqs.jl:396-416)
                    acc += a * b * self.node(var, j2)
        return acc

    def lin_coord(self, tab, var, dim, row_idx):
        start, w = self.tabs[tab]["rows"][row_idx]
        n, acc = self.n[dim], 0.0
        for k, wk in enumerate(w):
            j = start + k
            if self.per[var][dim]:
                j = j + n - 1 if j <= 1 else (j - (n - 1) if j > n else j)
            acc += wk * self.grid[dim][j - 1]
        return acc

    def weno(self, wid, var, dim, eps, dx, idx):
        start, T = self.wtabs[wid]["rows"][idx[dim]]
        n = self.n[dim]
        u, x = [], []
        for k in range(5):
            raw = start + k
            j2 = list(idx); j2[dim] = raw
            u.append(self.node(var, j2))
            if dx == 0.0:
                j, shift = raw, 0.0
                if self.per[var][dim]:
                    period = self.grid[dim][n - 1] - self.grid[dim][0]
                    if j <= 1:
                        j += n - 1; shift = -period
                    elif j > n:
                        j -= n - 1; shift = period
                x.append(self.grid[dim][j - 1] + shift)
        if dx != 0.0:
            return float(oweno.weno_f_uniform(u, eps, dx))
        return float(oweno.weno_f_nonuniform_core(u, eps, x, T))

    def nll(self, var, dim, fn, itab, dtab, otab, idx):
        ms, wo = self.tabs[otab]["rows"][idx[dim]]
        r = 0.0
        for k, wk in enumerate(wo):
            if wk == 0.0:          # zero-weight terms drop out (the reference simplifies 0 * expr symbolically)
                continue
            m = ms + k
            uh = [self.lin(itab, v, dim, m, idx) for v in range(self.nv)]
            xh = self.lin_coord(itab, var, dim, m)
            dh = self.lin(dtab, var, dim, m, idx)
            r += wk * self.rpn(self.fns[fn], idx, mode="fn", uh=uh, xh=(dim, xh)) * dh
        return r

    UN = {"sqrt": math.sqrt, "exp": math.exp, "log": math.log, "sin": math.sin, "cos": math.cos, "tan": math.tan,
          "sinh": math.sinh, "cosh": math.cosh, "tanh": math.tanh, "abs": abs, "asin": math.asin, "acos": math.acos,
          "atan": math.atan, "erf": math.erf}

    def rpn(self, toks, idx, mode="eq", uh=None, xh=None):
        st = []
        for tk in toks:
            f = tk.split(":")
            op = f[0]
            if op == "c":
                st.append(_f(f[1]))
            elif op == "p":
                st.append(self.p[int(f[1])])
            elif op == "t":
                st.append(self.t)
            elif op == "x":
                j = int(f[1])
                st.append(xh[1] if (mode == "fn" and j == xh[0]) else float(self.grid[j][idx[j] - 1]))
            elif op == "u":
                st.append(uh[int(f[1])] if mode == "fn" else self.node(int(f[1]), idx))
            elif op == "s":                             # a variable of t alone, read at its single node
                st.append(self.node(int(f[1]), list(self.ilo[int(f[1])])))
            elif op == "L":
                st.append(self.lin(int(f[1]), int(f[2]), int(f[3]), idx[int(f[3])], idx))
            elif op == "W":
                st.append(self.weno(int(f[1]), int(f[2]), int(f[3]), _f(f[4]), _f(f[5]), idx))
            elif op == "M":
                st.append(self.mixed(int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]), idx))
            elif op == "N":
                st.append(self.nll(int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5]), int(f[6]), idx))
            elif op == "neg":
                st.append(-st.pop())
            elif op == "sign":
                a = st.pop(); st.append(1.0 if a > 0 else (-1.0 if a < 0 else 0.0))
            elif op in self.UN:
                st.append(self.UN[op](st.pop()))
            elif op in ("+", "-", "*", "/"):
                b = st.pop(); a = st.pop()
                st.append(a + b if op == "+" else a - b if op == "-" else a * b if op == "*" else a / b)
            elif op in ("pow", "min", "max"):
                b = st.pop(); a = st.pop()
                st.append(a ** b if op == "pow" else (min(a, b) if op == "min" else max(a, b)))
            elif op == "powi":
                a, n = st.pop(), int(f[1])
                r = 1.0
                for _ in range(abs(n)):
                    r *= a
                st.append(1.0 / r if n < 0 else r)
            elif op in ("gt", "ge", "lt", "le", "eq", "ne"):
                b = st.pop(); a = st.pop()
                st.append({"gt": a > b, "ge": a >= b, "lt": a < b, "le": a <= b, "eq": a == b, "ne": a != b}[op])
            elif op in ("and", "or"):
                b = bool(st.pop()); a = bool(st.pop())
                st.append((a and b) if op == "and" else (a or b))
            elif op == "not":
                st.append(not bool(st.pop()))
            elif op == "sel":
                b = st.pop(); a = st.pop(); c = st.pop()
                st.append(float(a) if c else float(b))
            else:
                raise ValueError(f"unknown RPN token {tk}")
        assert len(st) == 1
        return float(st[0])

    def rhs(self, u, t, p=None):
        self.u, self.t = np.asarray(u, dtype=float), float(t)
        self.p = [self.params[k] for k in sorted(self.params)] if p is None else list(p)
        du = np.zeros(self.nstate)
        for v in range(self.nv):
            ranges = [range(self.ilo[v][d], self.ihi[v][d] + 1) for d in range(self.nd)]
            flat = self.off[v]
            for idx in np.ndindex(*[len(r) for r in reversed(ranges)]):        # last dimension slowest
                node = [ranges[d][idx[self.nd - 1 - d]] for d in range(self.nd)]
                du[flat] = self.rpn(self.eqs[v], node)
                flat += 1
        return du
