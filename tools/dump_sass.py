"""Diagnostic: write the SASS of one kernel variant of the Brusselator plan (compile-only) to a file."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa  (loads PyTorch's NVRTC first, as bench.py does)
import _mol_import  # noqa
import subprocess, tempfile
import mol_b200
from mol_b200 import capi
import problems as examples
key, outp = sys.argv[1], sys.argv[2]
plan = capi.Plan(mol_b200.symbolic_discretize(*examples.brusselator_2d(4096)).text, device=-1)
cb = plan.cubin(key)
f = tempfile.NamedTemporaryFile(suffix=".cubin", delete=False); f.write(cb); f.close()
open(outp, "w").write(subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True).stdout)
os.unlink(f.name)
