import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, sys
import _mol_import, subprocess, tempfile, os
import mol_b200
from mol_b200 import capi
import problems as examples
plan = capi.Plan(mol_b200.symbolic_discretize(*examples.brusselator_2d(4096)).text, device=-1)
for k in ["tiled_nin1","tiled_nin2","tiled_nin3","tiled_nin4","tiled_nin5","tiled_nin6_pre","tiled_nin1_fin"]:
    cb = plan.cubin(k)
    f = tempfile.NamedTemporaryFile(suffix=".cubin", delete=False); f.write(cb); f.close()
    out = subprocess.run(["cuobjdump","-res-usage",f.name],capture_output=True,text=True).stdout
    sass = subprocess.run(["cuobjdump","-sass",f.name],capture_output=True,text=True).stdout
    n = sum(1 for l in sass.splitlines() if l.strip().startswith("/*") and ";" in l)
    print(k, [l.split()[0] for l in out.splitlines() if "REG" in l], "sass lines", n)
    os.unlink(f.name)
