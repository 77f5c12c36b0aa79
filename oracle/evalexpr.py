"""Recursive NumPy evaluator for SymPy expressions (oracle only; test infrastructure).

`evaluate(expr, env)` with env: {Symbol or applied function -> scalar / ndarray}.  Handles the
node types the reference's PDE/BC expressions use (arithmetic, powers, elementary functions,
relationals, And/Or/Not, Piecewise == Symbolics `ifelse`).  Piecewise evaluates every branch
and selects with np.where, exactly like the reference's generated `ifelse`.
"""
import numpy as np
import sympy as sp

_FUNCS = {
    sp.exp: np.exp, sp.log: np.log, sp.sin: np.sin, sp.cos: np.cos, sp.tan: np.tan,
    sp.sinh: np.sinh, sp.cosh: np.cosh, sp.tanh: np.tanh, sp.Abs: np.abs, sp.sign: np.sign,
    sp.asin: np.arcsin, sp.acos: np.arccos, sp.atan: np.arctan, sp.erf: None,
}


def evaluate(e, env):
    if e in env:
        return env[e]
    if e.is_Number or isinstance(e, sp.NumberSymbol):
        return float(e)
    if e is sp.true:
        return True
    if e is sp.false:
        return False
    if isinstance(e, sp.Add):
        args = [evaluate(a, env) for a in e.args]
        out = args[0]
        for a in args[1:]:
            out = out + a
        return out
    if isinstance(e, sp.Mul):
        args = [evaluate(a, env) for a in e.args]
        out = args[0]
        for a in args[1:]:
            out = out * a
        return out
    if isinstance(e, sp.Pow):
        b, x = e.args
        bv = evaluate(b, env)
        if x.is_Integer:
            n = int(x)
            if n == 2:
                return bv * bv
            if n == -1:
                return 1.0 / bv
            if n < 0:
                return 1.0 / np.power(bv, -n)
            return np.power(bv, n)
        if x == sp.Rational(1, 2):
            return np.sqrt(bv)
        return np.power(bv, evaluate(x, env))
    if isinstance(e, sp.Piecewise):
        out = None
        for val, cond in reversed(e.args):
            v = evaluate(val, env)
            if cond is sp.true:
                out = v
            else:
                c = evaluate(cond, env)
                out = np.where(c, v, np.nan if out is None else out)
        return out
    if isinstance(e, sp.And):
        out = True
        for a in e.args:
            out = np.logical_and(out, evaluate(a, env))
        return out
    if isinstance(e, sp.Or):
        out = False
        for a in e.args:
            out = np.logical_or(out, evaluate(a, env))
        return out
    if isinstance(e, sp.Not):
        return np.logical_not(evaluate(e.args[0], env))
    if isinstance(e, sp.core.relational.Relational):
        a, b = evaluate(e.lhs, env), evaluate(e.rhs, env)
        return {sp.Gt: np.greater, sp.Ge: np.greater_equal, sp.Lt: np.less, sp.Le: np.less_equal,
                sp.Eq: np.equal, sp.Ne: np.not_equal}[type(e)](a, b)
    if isinstance(e, (sp.Max, sp.Min)):
        f = np.maximum if isinstance(e, sp.Max) else np.minimum
        args = [evaluate(a, env) for a in e.args]
        out = args[0]
        for a in args[1:]:
            out = f(out, a)
        return out
    if isinstance(e, sp.Function) and type(e) in _FUNCS and _FUNCS[type(e)] is not None:
        return _FUNCS[type(e)](evaluate(e.args[0], env))
    if isinstance(e, sp.erf):
        from scipy.special import erf
        return erf(evaluate(e.args[0], env))
    raise NotImplementedError(f"oracle evaluator: unsupported node {type(e).__name__}: {e}")


def evaluate_abs(e, env):
    """Magnitude evaluation: the same tree as `evaluate`, but sums add |terms| and products multiply
    |factors|, so the result bounds what floating-point rounding of `evaluate(e, env)` scales with.
    Conditions (relationals, And/Or/Not) are evaluated exactly; leaves bound to arrays in `env` are
    taken as they are (callers pass magnitudes for stencil sums) and wrapped in abs."""
    if e in env:
        return np.abs(env[e])
    if e.is_Number or isinstance(e, sp.NumberSymbol):
        return abs(float(e))
    if isinstance(e, sp.Add):
        out = 0.0
        for a in e.args:
            out = out + evaluate_abs(a, env)
        return out
    if isinstance(e, sp.Mul):
        out = 1.0
        for a in e.args:
            out = out * evaluate_abs(a, env)
        return out
    if isinstance(e, sp.Piecewise):
        out = None
        for val, cond in reversed(e.args):
            v = evaluate_abs(val, env)
            if cond is sp.true:
                out = v
            else:
                out = np.where(evaluate(cond, env), v, 0.0 if out is None else out)
        return out
    if isinstance(e, sp.Pow) and e.args[1].is_Integer and int(e.args[1]) > 0:
        return np.power(evaluate_abs(e.args[0], env), int(e.args[1]))
    return np.abs(evaluate(e, env))
