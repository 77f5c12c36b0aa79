"""CPU-side checks of libmol_cuda.so: it loads, exports the whole C ABI of include/mol_cuda.h,
its host functions agree with the reference KATs, and the flagship stencil program compiles to an
sm_100a cubin through NVRTC without a GPU.  No compute entry point is called here."""
import ctypes
import os
import re

import numpy as np
import pytest

import mol_b200
from mol_b200 import capi, examples

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
