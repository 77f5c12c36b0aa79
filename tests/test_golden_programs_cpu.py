"""tests/golden/program_*.txt: the stencil programs of the two reference problems the Julia serializer
(julia/MOLCudaStencil.jl, julia/test_molcuda.jl) is pinned on.  They must stay what the Python twin emits, parse in the
library, and -- for the Brusselator -- compile to kernels (device = -1) whose table-driven form reproduces the
reference's literal RHS through the emulator."""
import json
import os

import numpy as np

import _mol_import  # noqa: F401
import mol_b200
from mol_b200 import capi
import problems as examples

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_golden_programs_are_what_the_lowering_emits():
    for name, mk in (("bruss_n4", lambda: examples.brusselator_2d(4)), ("heat1d", lambda: examples.heat_1d_dirichlet(dx=0.01))):
        want = open(os.path.join(GOLD, f"program_{name}.txt")).read()
        assert mol_b200.symbolic_discretize(*mk()).text == want, name


def test_golden_brusselator_program_reproduces_the_reference_literal_rhs():
    from cuda_emu import EmuKernel
    text = open(os.path.join(GOLD, "program_bruss_n4.txt")).read()
    plan = capi.Plan(text, device=-1)
    prog = mol_b200.symbolic_discretize(*examples.brusselator_2d(4))
    G = json.load(open(os.path.join(GOLD, "bruss_code_n4.json")))
    emu = EmuKernel(plan, prog)
    for case in G["cases"]:
        ref = np.array(case["du"])
        got = emu.rhs([np.array(case["u"])], [1.0], 0.0)          # the dump has no time argument: forcing off
        assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref))
    plan.close()


def test_program_in_c99_shortest_hex_and_other_table_order_parses_to_the_same_tables():
    """A serializer in another language will not print Python's 13-digit hex floats nor number its tables the same way:
    the parser reads any C99 hex float, and table ids are arbitrary labels."""
    import re
    text = open(os.path.join(GOLD, "program_heat1d.txt")).read()

    short = re.sub(r"-?0x[0-9a-f.]+p[+-]\d+", lambda m: "%a" % float.fromhex(m.group(0)), text)      # C99 %a: shortest digits
    assert short != text
    short = short.replace("tab 0 ", "tab 7 ").replace("core 0 ", "core 7 ").replace("score 0 ", "score 7 ").replace("L:0:", "L:7:")
    a, b = capi.Plan(text, device=-1), capi.Plan(short, device=-1)
    wa, sa = a.tables()
    wb, sb = b.tables()
    assert np.array_equal(wa, wb) and np.array_equal(sa, sb)
    a.close(); b.close()
