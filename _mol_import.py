"""Makes the package directory `methodoflines.jl_b200/` importable as `mol_b200`.

The directory name is fixed by the project layout and is not a valid Python identifier
(it contains a dot), so it is registered under an alias via importlib.
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "methodoflines.jl_b200")


def load():
    if "mol_b200" in sys.modules:
        return sys.modules["mol_b200"]
    spec = importlib.util.spec_from_file_location(
        "mol_b200", os.path.join(PKG_DIR, "__init__.py"), submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["mol_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# the reference's test problems restated as fixtures (tests/problems.py) are shared by tests, tools, smoke and bench
if os.path.join(ROOT, "tests") not in sys.path:
    sys.path.insert(1, os.path.join(ROOT, "tests"))
load()
