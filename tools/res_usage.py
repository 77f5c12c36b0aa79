"""Resource usage (registers / stack / spills) of every tiled-kernel variant of a stencil program, compile-only (no GPU).
usage: python tools/res_usage.py [example] [N]"""
import os, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import _mol_import  # noqa
import mol_b200
import problems as examples

name = sys.argv[1] if len(sys.argv) > 1 else "brusselator_2d"
arg = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
kw = {"n": arg} if name == "diffusion_reaction_3d" else {}
sys_, disc = getattr(examples, name)(**kw) if kw else getattr(examples, name)(arg)
prog = mol_b200.symbolic_discretize(sys_, disc)
plan = mol_b200.capi.Plan(prog.text, device=-1)
keys = ["tiled_nin1_tma"] + [f"tiled_nin{k}" for k in range(2, 6)] + ["tiled_nin6_pre", "tiled_nin1_fin_tma", "generic_nin1", "generic_nin6_pre", "generic_nin1_fin"]
for k in keys:
    try:
        cubin = plan.cubin(k)
    except Exception as e:
        print(k, "->", e)
        continue
    with tempfile.NamedTemporaryFile(suffix=".cubin") as f:
        f.write(cubin)
        f.flush()
        ru = subprocess.run(["cuobjdump", "-res-usage", f.name], capture_output=True, text=True).stdout
        sass = subprocess.run(["cuobjdump", "-sass", f.name], capture_output=True, text=True).stdout
    line = [l.strip() for l in ru.splitlines() if l.strip().startswith("REG")]
    nl = sum(1 for l in sass.splitlines() if "LDL" in l or "STL" in l)
    print(f"{k:18s} {line[0] if line else '?'}  local-mem ops: {nl}")
