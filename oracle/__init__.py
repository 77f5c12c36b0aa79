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
  * literal RHS dump   docs/src/generated/bruss_code.md:82-113  (32 outputs)
  * WENO kernel        test/Components/weno_nonuniform_core.jl, weno_nonuniform_boundary.jl
(see tests/test_oracle_*.py and tests/golden/).
"""
