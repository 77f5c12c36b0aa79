"""CPU oracle for the MethodOfLines.jl hot path (RHS evaluation + explicit RK).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import anything from here.  The product package
(``methodoflines.jl_b200``) never imports ``oracle``.

It restates, in NumPy, the semantics of the reference's scalarized
discretization (``src/scalar_discretization.jl:1-64``) on *full-grid* arrays
(every grid node incl. boundary nodes), which is a deliberately different
architecture from the GPU path (interior-only state + on-the-fly ghost rules).

Parity status: the reference (Julia) cannot run in this image, so the oracle is
pinned against the reference's own known-answer tests and literal artifacts:
  * Fornberg weights   test/Components/MOLfornberg_weights.jl:8-30
  * stencil tables     test/shared/finite_diff_schemes.jl:23-30
  * periodic wrap      test/Components/utils_test.jl:129-138
  * grids / interiors  test/Components/DiscreteSpace.jl:42-46,89-93; test/Components/weno_boundary_integration.jl:54-86
  * literal RHS dump   docs/src/generated/bruss_code.md:82-113  (32 outputs)
  * independent loop   test/Brusselator/brusselator_eq.jl:79-105 (`brusselator_2d_loop`, N = 32, forcing off and on)
  * interface charts   test/Components/weno_interface_coords.jl:118-166 (bcoord across a two-domain interface)
  * WENO kernel        test/Components/weno_nonuniform_core.jl, weno_nonuniform_boundary.jl
  * solution level     the reference's own acceptance criteria, at its tolerances: test/Diffusion Tests 03 (both grid
                       alignments, integral conserved to 1e-9), 05, 07; test/Diffusion_NU Tests 00 (orders 2, 4), 04;
                       test/2D_Diffusion Test 00; test/Burgers (upwind, WENO); test/Nonlinear_Diffusion Test 01a;
                       test/Convection Test 00; test/Nonlinear_Diffusion_NU Tests 01a/01b
                       (tests/test_zz_reference_acceptance.py); test/Diffusion Test 14 (two domains), test/Convection_NU
                       interface / periodic non-uniform upwind tests, test/Convection_WENO two-domain convergence order
                       (tests/test_interface_cpu.py)
  * measured output    test/Convection_WENO/MOL_1D_WENO_NU_Convergence.jl:95-128 records what the reference's own run
                       measured ("Calibration: EOC ≈ 3.85 / ≈ 2.45, err_neg/err_pos ≈ 1.002, err ≈ 9.4e-6"): the oracle's
                       SSPRK33 solves give 3.846 / 2.445 / 1.0022 / 9.405e-6 (tests/test_zz_reference_acceptance.py)
(see tests/test_oracle_kats.py and tests/golden/).

Looped C restatements for the sizes the Python oracle cannot reach (oracle/bruss_ref.c: config 2 and its slab form for
the per-rank parity check inside bench.py; oracle/configs_ref.c: config 5 on a slab of z planes, config 3 with the
oracle's own row tables) are validated against this package at small sizes (tests/test_cref_cpu.py) and then used as
checkers at benchmark size (tests/test_gpu_bench_size_parity.py, bench.py).  oracle/cref.py builds them with a
content + host-CPU stamp, and times them as the CPU baseline (kind "port": the reference is pure Julia).

Per-evaluation du at other sizes / schemes is not pinned by any reference test: there the oracle is the
reference's semantics as restated here, and is itself cross-checked by two independent executions of the lowering's
stencil program (tests/ir_interp.py, tests/cuda_emu).
"""
