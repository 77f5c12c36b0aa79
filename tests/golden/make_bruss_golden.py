"""Generates tests/golden/bruss_code_n4.json from the reference's literal RHS dump.

Run in the build container only (needs /root/reference):  python tests/golden/make_bruss_golden.py
Source artifact: /root/reference/docs/src/generated/bruss_code.md:48-113 — the Julia code MTK
generated for the 2-D Brusselator at dx = dy = 1/4 (32 unknowns).  The prefix-call text
`(+)((*)(160.0, var"u[2, 3](t)"), ...)` is parsed and evaluated in float64 with exactly the
nesting (= evaluation order) of the dump; inputs are seeded random vectors; outputs are stored.
The fixture therefore pins per-evaluation `du` values produced by the REFERENCE's generated code.
"""
import json
import os
import re

import numpy as np

SRC = "/root/reference/docs/src/generated/bruss_code.md"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bruss_code_n4.json")

TOK = re.compile(r'\s*(var"[^"]*"|\(\+\)|\(\*\)|\(\^\)|\(-\)|\(/\)|[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?|[(),])')


def tokenize(s):
    pos, out = 0, []
    while pos < len(s):
        m = TOK.match(s, pos)
        if not m:
            if s[pos:].strip() == "":
                break
            raise ValueError(f"bad token at {s[pos:pos + 40]!r}")
        out.append(m.group(1))
        pos = m.end()
    return out


def parse(tokens, i=0):
    tok = tokens[i]
    if tok in ("(+)", "(*)", "(^)", "(-)", "(/)"):
        assert tokens[i + 1] == "("
        args, i = [], i + 2
        while True:
            a, i = parse(tokens, i)
            args.append(a)
            if tokens[i] == ",":
                i += 1
            elif tokens[i] == ")":
                i += 1
                break
        return (tok[1], args), i
    if tok.startswith('var"'):
        return ("var", tok[4:-1]), i + 1
    return ("num", float(tok)), i + 1


def evaluate(node, env):
    kind, val = node
    if kind == "num":
        return val
    if kind == "var":
        return env[val]
    a = [evaluate(x, env) for x in val]
    if kind == "+":
        r = a[0]
        for x in a[1:]:
            r = r + x
        return r
    if kind == "*":
        r = a[0]
        for x in a[1:]:
            r = r * x
        return r
    if kind == "^":
        assert a[1] == 2.0
        return a[0] * a[0]          # Julia lowers x^2 to x*x (Base.literal_pow)
    if kind == "-":
        return -a[0] if len(a) == 1 else a[0] - a[1]
    if kind == "/":
        return a[0] / a[1]
    raise ValueError(kind)


def main():
    lines = open(SRC, encoding="utf-8").read().splitlines()
    # in-place function: bindings `var"u[2, 2](t)" = @inbounds(ˍ₋arg1[1])` then `ˍ₋out[k] = expr`
    start = next(i for i, l in enumerate(lines) if "function (ˍ₋out" in l)
    binds, outs = {}, {}
    for l in lines[start:]:
        if l.startswith("```"):
            break
        m = re.match(r'\s*(var"[^"]*") = @inbounds\(ˍ₋arg1\[(\d+)\]\)', l)
        if m:
            binds[m.group(1)[4:-1]] = int(m.group(2)) - 1
        m = re.match(r'\s*ˍ₋out\[(\d+)\] = (.*)$', l)
        if m:
            text = re.sub(r'\[Imgur\]\(.*\)\s*$', '', m.group(2))   # stray markdown link on line 102 of the doc
            node, _ = parse(tokenize(text))
            outs[int(m.group(1)) - 1] = node
    assert len(binds) == 32 and sorted(outs) == list(range(32)), (len(binds), len(outs))
    names = [None] * 32
    for k, v in binds.items():
        names[v] = k
    cases = []
    for seed in (0, 1, 2):
        rng = np.random.default_rng(seed)
        u = rng.uniform(0.0, 3.0, 32)
        env = {names[i]: float(u[i]) for i in range(32)}
        du = [evaluate(outs[k], env) for k in range(32)]
        cases.append({"seed": seed, "u": u.tolist(), "du": du})
    ones = {n: 1.0 for n in names}
    cases.append({"seed": "ones", "u": [1.0] * 32, "du": [evaluate(outs[k], ones) for k in range(32)]})
    json.dump({"source": "docs/src/generated/bruss_code.md:48-113", "dx": 0.25, "unknowns": names,
               "cases": cases}, open(OUT, "w"), indent=1)
    print("wrote", OUT, "ones->", cases[-1]["du"][0], cases[-1]["du"][16])


if __name__ == "__main__":
    main()
