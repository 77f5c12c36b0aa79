"""CPU-side checks of libmol_cuda.so: it loads, exports the whole C ABI of include/mol_cuda.h,
its host functions agree with the reference KATs, and the flagship stencil program compiles to an
sm_100a cubin through NVRTC without a GPU.  No compute entry point is called here."""
import ctypes
import os
import re

import numpy as np
import pytest

import mol_b200
from mol_b200 import capi
import problems as examples

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mol_cuda.h")).read()
    names = set(re.findall(r"\b(mol_[a-z0-9_]+)\s*\(", hdr))
    names -= {"mol_plan", "mol_rk"}
    assert len(names) >= 20
    lib = ctypes.CDLL(capi.LIB_PATH)
    for n in sorted(names):
        assert hasattr(lib, n), f"libmol_cuda.so does not export {n}"


def test_fd_weights_known_answers():
    # test/Components/MOLfornberg_weights.jl:8-30
    assert np.array_equal(capi.fd_weights(2, 0.0, [-1, 0, 1.0]), [1, -2, 1])
    assert np.array_equal(capi.fd_weights(1, 0.0, [-1.0, 1.0]), [-0.5, 0.5])
    assert np.array_equal(capi.fd_weights(1, 1.0, [0, 1]), [-1, 1])
    assert np.array_equal(capi.fd_weights(3, 0.0, [0, 1, 2, 3, 4, 5]), [-17 / 4, 71 / 4, -59 / 2, 49 / 2, -41 / 4, 7 / 4])


def test_fd_weights_bit_identical_to_oracle():
    from oracle.fornberg import calculate_weights
    rng = np.random.default_rng(3)
    for _ in range(50):
        n = int(rng.integers(2, 9))
        x = np.sort(rng.uniform(-2, 2, n))
        d = int(rng.integers(0, n))
        x0 = float(rng.uniform(-2, 2))
        assert np.array_equal(capi.fd_weights(d, x0, x), calculate_weights(d, x0, x))


def test_compile_only_plan_and_tma_in_sass(tmp_path):
    sys_, disc = examples.brusselator_2d(64)
    prog = mol_b200.symbolic_discretize(sys_, disc)
    assert prog.nstate == 2 * 64 * 64 and prog.corebox == ([2, 2], [65, 65])
    plan = capi.Plan(prog.text, device=-1)
    assert plan.state_len == prog.nstate
    src = plan.generated_source()
    assert "cp.async.bulk.tensor.2d" in src and "mol_eq_tile<1>" in src
    cubin = plan.cubin("tiled_nin1_tma")
    assert cubin[:4] == b"\x7fELF"
    p = tmp_path / "k.cubin"
    p.write_bytes(cubin)
    import shutil
    import subprocess
    if shutil.which("cuobjdump"):
        sass = subprocess.run(["cuobjdump", "-sass", str(p)], capture_output=True, text=True).stdout
        assert "UTMALDG" in sass and "DFMA" in sass      # TMA tile loads + FP64 FMAs in the hot kernel
    plan.close()


def test_compute_without_gpu_fails_loudly():
    sys_, disc = examples.heat_1d_dirichlet(dx=0.1)
    prog = mol_b200.symbolic_discretize(sys_, disc)
    plan = capi.Plan(prog.text, device=-1)
    with pytest.raises(capi.MolError) as e:
        plan.rhs(0, 0, 0.0)
    assert e.value.code in (-5, -6)
    plan.close()


def test_malformed_program_is_rejected():
    with pytest.raises(capi.MolError) as e:
        capi.Plan("MOLPROG 1\nndim 9\nend\n", device=-1)
    assert e.value.code == -1


def _sass(cubin, tmp_path, name):
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    p = tmp_path / (name + ".cubin")
    p.write_bytes(cubin)
    ru = subprocess.run(["cuobjdump", "-res-usage", str(p)], capture_output=True, text=True).stdout
    sass = subprocess.run(["cuobjdump", "-sass", str(p)], capture_output=True, text=True).stdout
    return ru, sass


def test_every_staging_flavour_compiles_for_sm100a(tmp_path):
    """The three staging flavours of the tiled kernel and both fused Runge-Kutta epilogues build (NVRTC, no GPU) and
    carry the instructions that define them."""
    # TMA: even row pitch, one input (256 nodes per row: tiles away from the edge exist, so the 128-bit loader is live)
    plan = capi.Plan(mol_b200.symbolic_discretize(*examples.brusselator_2d(256)).text, device=-1)
    for key, must in (("tiled_nin1_tma", "UTMALDG"), ("tiled_nin1_fin_tma", "UTMALDG"), ("tiled_nin3", "LDG.E.128"),
                      ("tiled_nin6_pre", "LDG.E.128"), ("generic_nin1", "DFMA"), ("generic_nin6_pre", "DFMA"),
                      ("generic_nin1_fin", "DFMA")):
        ru, sass = _sass(plan.cubin(key), tmp_path, key)
        assert must in sass, (key, must)
        assert "LDL" not in sass or key.endswith("pre"), f"{key}: local-memory traffic in the kernel"
    # the hot kernel keeps the tile-box fields in the constant bank (no local-stack copy of the kernel parameter)
    ru, sass = _sass(plan.cubin("tiled_nin1_tma"), tmp_path, "hot")
    assert "STACK:0" in ru and "LDL" not in sass and "STL" not in sass
    plan.close()
    # cp.async: odd row pitch (Dirichlet/Neumann on 2^k + 1 nodes) and 1-D programs
    plan = capi.Plan(mol_b200.symbolic_discretize(*examples.burgers_2d(nx=65, ny=65)).text, device=-1)
    ru, sass = _sass(plan.cubin("tiled_nin1"), tmp_path, "cpa2d")
    assert "LDGSTS" in sass and "UTMALDG" not in sass
    plan.close()
    plan = capi.Plan(mol_b200.symbolic_discretize(*examples.heat_1d_dirichlet(dx=0.001)).text, device=-1)
    ru, sass = _sass(plan.cubin("tiled_nin1"), tmp_path, "cpa1d")
    assert "LDGSTS" in sass
    plan.close()
    # 3-D: z-marching ring of planes, one TMA load per plane
    plan = capi.Plan(mol_b200.symbolic_discretize(*examples.diffusion_reaction_3d(n=64)).text, device=-1)
    src = plan.generated_source()
    assert "#define MOL_ZMARCH 1" in src and "#define MOL_RING 5" in src
    ru, sass = _sass(plan.cubin("tiled_nin1_tma"), tmp_path, "zmarch")
    assert "UTMALDG.3D" in sass
    plan.close()


def test_nonuniform_grid_takes_the_tiled_path_with_table_weights():
    """Non-uniform axes: interior rows share their taps and differ in their weights only (IR directive `score`); the
    tiled kernel reads them from the table; identical tables of u and v are shared."""
    gx = 0.5 * (1 + np.tanh(2.0 * np.linspace(-1, 1, 41)) / np.tanh(2.0))
    gy = np.linspace(0, 1, 37) ** 1.3
    prog = mol_b200.symbolic_discretize(*examples.burgers_2d(grid_x=gx, grid_y=gy))
    lines = prog.text.split("\n")
    assert prog.corebox == ([2, 2], [40, 36])
    assert sum(l.startswith("tab ") for l in lines) == 6          # 12 operators, u and v share theirs
    assert sum(l.startswith("score ") for l in lines) == 6 and not any(l.startswith("core ") for l in lines)
    plan = capi.Plan(prog.text, device=-1)
    src = plan.generated_source()
    assert "#define MOL_HAVE_TILE 1" in src and "c.tabw +" in src.split("mol_eq_tile<0>")[-1]
    # ... packed into one record per node and dimension, staged in shared memory with the tile (x records field-major);
    # such programs run 64 x 32 tiles, two stages, two CTAs of 256 threads per SM
    assert "#define MOL_WRS0 8" in src and "#define MOL_WRS1 8" in src and "MOL_WX(0, " in src and "MOL_WY(0, " in src
    assert "#define MOL_TY 32" in src and "#define MOL_NTHREADS 256" in src and "#define MOL_STAGES 2" in src
    plan.close()
    # non-uniform WENO5 tiles too: centre-target rows form the core, the kernel reads the per-interval geometry arrays
    # the library builds at plan time (no Fornberg recurrence in device code any more)
    prog = mol_b200.symbolic_discretize(*examples.advection_1d_periodic(dx=examples.stretched_grid(0, 2, 64),
                                                                        scheme=mol_b200.WENOScheme()))
    assert prog.corebox == ([2], [64])
    plan = capi.Plan(prog.text, device=-1)
    src = plan.generated_source()
    assert "mol_weno5_nu_core<double, true, false>" in src.split("mol_eq_tile<0>")[-1] and "mol_fornberg3" not in src
    plan.close()


def test_uniform_nodes_match_exact_rational_ranges():
    """Grid nodes of a:dx:b are rounded once per node from exact rationals (Julia's range arithmetic)."""
    from fractions import Fraction
    from mol_b200 import lowering
    for a, dx, n in ((0.0, 1 / 4096, 4097), (0.0, 0.01, 101), (-1.0, 0.05, 41), (0.3, 1 / 3, 10)):
        ra, rd = lowering._rationalize(a), lowering._rationalize(dx)
        assert np.array_equal(lowering.uniform_nodes(a, dx, n), np.array([float(ra + k * rd) for k in range(n)]))


def test_cubin_cache_roundtrip(tmp_path, monkeypatch):
    """MOL_CUBIN_CACHE (opt-in): a variant compiled once is read back from disk bit for bit; without the variable
    nothing is written."""
    import time
    prog = mol_b200.symbolic_discretize(*examples.brusselator_2d(64))
    monkeypatch.delenv("MOL_CUBIN_CACHE", raising=False)
    plan = capi.Plan(prog.text, device=-1)
    fresh = {k: plan.cubin(k) for k in ("tiled_nin1_tma", "tiled_nin6_pre", "generic_nin1")}
    plan.close()
    assert not list(tmp_path.iterdir())
    monkeypatch.setenv("MOL_CUBIN_CACHE", str(tmp_path))
    plan = capi.Plan(prog.text, device=-1)
    first = {k: plan.cubin(k) for k in fresh}
    plan.close()
    files = sorted(p.name for p in tmp_path.iterdir())
    assert len(files) >= 3 and all(f.endswith(".molcubin") for f in files)
    t0 = time.perf_counter()
    plan = capi.Plan(prog.text, device=-1)
    cached = {k: plan.cubin(k) for k in fresh}
    plan.close()
    dt = time.perf_counter() - t0
    assert sorted(p.name for p in tmp_path.iterdir()) == files          # nothing recompiled
    for k in fresh:
        assert first[k] == fresh[k] == cached[k], k                      # NVRTC is deterministic; the cache is exact
    assert dt < 1.0, dt
    # a corrupt entry is ignored and replaced
    victim = tmp_path / files[0]
    victim.write_bytes(b"garbage")
    plan = capi.Plan(prog.text, device=-1)
    again = {k: plan.cubin(k) for k in fresh}
    plan.close()
    assert again == fresh and victim.read_bytes()[:9] == b"MOLCUBIN1"


def test_fd_weights_rows_equals_row_by_row_calls():
    """mol_fd_weights_rows (the per-node loops of the non-uniform tables in one call) is bit-identical to one
    mol_fd_weights call per row, and an order-2 non-uniform lowering of 2^16 nodes stays a matter of seconds."""
    import time
    import mol_b200
    import problems as examples
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(0.0, 1.0, 64))
    win = np.lib.stride_tricks.sliding_window_view(x, 5)
    for d in (0, 1, 2, 3):
        W = capi.fd_weights_rows(d, x[2:-2], win)
        for r in range(win.shape[0]):
            assert np.array_equal(W[r], capi.fd_weights(d, x[2 + r], win[r]))
    with pytest.raises(capi.MolError):
        capi.fd_weights_rows(5, x[2:-2], win)                       # not enough points for the order: same error per row
    t0 = time.time()
    prog = mol_b200.symbolic_discretize(*examples.heat_1d_dirichlet_pi(examples.jittered_grid(0.0, float(np.pi), 1 << 16, amp=1e-7)))
    assert prog.nstate == (1 << 16) - 2 and time.time() - t0 < 30.0


def test_precompile_builds_every_variant_of_an_integrator():
    """mol_plan_precompile (called by mol_rk_init): all kernel variants of a Tsit5 step -- stage loaders nin = 1..5, the
    PRE and FIN epilogues; tiled where the plan tiles, table-driven where it has a frame or does not tile -- are compiled
    in one call on several host threads and are then served from the plan without another NVRTC run."""
    import time
    stages = ["nin2", "nin3", "nin4", "nin5", "nin6_pre", "nin1_fin"]
    for mk, kind in ((lambda: examples.burgers_2d(nx=64, ny=64), "tiled"),
                     (lambda: examples.diffusion_two_domains(l=40), "generic")):
        prog = mol_b200.symbolic_discretize(*mk())
        assert (prog.corebox is not None) == (kind == "tiled")
        plan = capi.Plan(prog.text, device=-1)
        plan.precompile("tsit5")
        t0 = time.time()
        for st in stages:
            assert plan.cubin(f"{kind}_{st}")[:4] == b"\x7fELF"
        assert time.time() - t0 < 1.0                               # no compilation happened in this loop (6 x >= 0.3 s otherwise)
        plan.precompile("ssprk33")                                  # a subset: nothing left to do
        with pytest.raises(capi.MolError):
            capi.check(capi.lib().mol_plan_precompile(plan.handle, 99))
        plan.close()
