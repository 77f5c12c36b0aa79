"""Diagnostic: instruction mix of one kernel variant of an example problem (compile-only plan, cuobjdump -sass)."""
import os, sys, re, subprocess, tempfile, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa
import _mol_import  # noqa
import numpy as np
import mol_b200
from mol_b200 import capi
import problems as examples
cases = {"weno1d": lambda: examples.advection_1d_periodic(dx=2.0 / 4096, scheme=mol_b200.WENOScheme()),
         "weno2d": lambda: examples.advection_2d_periodic(256, scheme=mol_b200.WENOScheme()),
         "bruss": lambda: examples.brusselator_2d(256),
         "burgers_nu": lambda: examples.burgers_2d(grid_x=0.5 * (1 + np.tanh(2.0 * np.linspace(-1, 1, 257)) / np.tanh(2.0)), grid_y=np.linspace(0, 1, 257) ** 1.3)}
name, key = sys.argv[1], sys.argv[2]
plan = capi.Plan(mol_b200.symbolic_discretize(*cases[name]()).text, device=-1)
cb = plan.cubin(key)
f = tempfile.NamedTemporaryFile(suffix=".cubin", delete=False); f.write(cb); f.close()
sass = subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", f.name], capture_output=True, text=True).stdout
os.unlink(f.name)
ops = collections.Counter()
for l in sass.splitlines():
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops[m.group(1).split(".")[0]] += 1
tot = sum(ops.values())
fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print(name, key, [l.split()[0] for l in res.splitlines() if "REG" in l], "total", tot, "fp64", fp64, dict(ops.most_common(14)))
