# test_molcuda.jl — what a maintainer runs after wiring julia/MOLCuda.jl + julia/MOLCudaStencil.jl into MethodOfLines.jl
# (needs Julia, CUDA.jl, a B200 and LIBMOL_CUDA pointing at libmol_cuda.so; not runnable in this repository's image).
#
#   julia --project=. julia/test_molcuda.jl /path/to/repo
#
# Pins:
#  1. the reference's own literal RHS (docs/src/generated/bruss_code.md, tests/golden/bruss_code_n4.json): the ODEProblem
#     built by CudaStencilDiscretization evaluates to the recorded du on the recorded u (1e-12 of max |du|);
#  2. the same problem through ScalarizedDiscretization gives the same du (the drop-in claim, on identical inputs);
#  3. the 1-D heat problem (config 1) solves to the analytic solution within the reference's acceptance (0.01);
#  4. the serializer's directive counts equal those of the Python twin's program (tests/golden/program_*.txt);
#  5. an unsupported system (WENO advection) falls back to ScalarizedDiscretization with an @info, and raises under
#     strict = true.
using MethodOfLines, ModelingToolkit, DomainSets, OrdinaryDiffEq, CUDA, JSON, Test

repo = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..")
golden = JSON.parsefile(joinpath(repo, "tests", "golden", "bruss_code_n4.json"))

function brusselator_system(N; strategy = ScalarizedDiscretization())
    @parameters x y t
    @variables u(..) v(..)
    Dt = Differential(t); Dxx = Differential(x)^2; Dyy = Differential(y)^2
    brusselator_f(x, y, t) = (((x - 0.3)^2 + (y - 0.6)^2) <= 0.1^2) * (t >= 1.1) * 5.0
    α = 10.0
    eqs = [Dt(u(x, y, t)) ~ 1.0 + v(x, y, t) * u(x, y, t)^2 - 4.4 * u(x, y, t) + α * (Dxx(u(x, y, t)) + Dyy(u(x, y, t))) + brusselator_f(x, y, t),
           Dt(v(x, y, t)) ~ 3.4 * u(x, y, t) - v(x, y, t) * u(x, y, t)^2 + α * (Dxx(v(x, y, t)) + Dyy(v(x, y, t)))]
    bcs = [u(x, y, 0) ~ 22 * (y * (1 - y))^(3 / 2), u(0, y, t) ~ u(1, y, t), u(x, 0, t) ~ u(x, 1, t),
           v(x, y, 0) ~ 27 * (x * (1 - x))^(3 / 2), v(0, y, t) ~ v(1, y, t), v(x, 0, t) ~ v(x, 1, t)]
    domains = [x ∈ Interval(0.0, 1.0), y ∈ Interval(0.0, 1.0), t ∈ Interval(0.0, 11.5)]
    @named pdesys = PDESystem(eqs, bcs, domains, [x, y, t], [u(x, y, t), v(x, y, t)])
    pdesys, MOLFiniteDifference([x => 1 / N, y => 1 / N], t; approx_order = 2, discretization_strategy = strategy)
end
brusselator(N; strategy) = discretize(brusselator_system(N; strategy = strategy)...)

@testset "libmol_cuda behind discretize" begin
    prob = brusselator(4; strategy = MethodOfLines.CudaStencilDiscretization())
    ref = brusselator(4; strategy = ScalarizedDiscretization())
    for case in golden["cases"]
        u = CuArray(Float64.(case["u"])); du = similar(u)
        prob.f(du, u, prob.p, 0.0)                         # the dump has no forcing (t < 1.1)
        want = Float64.(case["du"])
        @test maximum(abs.(Array(du) .- want)) <= 1e-12 * maximum(abs.(want))
        # the scalarized problem orders its unknowns as `unknowns(simpsys)`; the documented dump is x fastest, u then v,
        # which is the layout of the stencil program, so only the literal golden is compared element-wise
        duref = similar(ref.u0); ref.f(duref, Float64.(case["u"]), ref.p, 0.0)
        @test sort(duref) ≈ sort(want) rtol = 1e-12
    end

    @parameters t x
    @variables w(..)
    heat = PDESystem([Differential(t)(w(t, x)) ~ Differential(x)^2(w(t, x))],
                     [w(0, x) ~ cos(x), w(t, 0) ~ exp(-t), w(t, 1) ~ exp(-t) * cos(1)],
                     [t ∈ Interval(0.0, 1.0), x ∈ Interval(0.0, 1.0)], [t, x], [w(t, x)]; name = :heat)
    hp = discretize(heat, MOLFiniteDifference([x => 0.01], t; discretization_strategy = MethodOfLines.CudaStencilDiscretization()))
    sol = solve(hp, Tsit5(), saveat = 0.2)                       # OrdinaryDiffEq drives mol_rhs, u resident in HBM
    xs = 0.01:0.01:0.99
    for (k, tk) in enumerate(sol.t)
        @test maximum(abs.(Array(sol.u[k]) .- exp(-tk) .* cos.(xs))) <= 0.01
    end

    for name in ("bruss_n4", "heat1d")
        text = read(joinpath(repo, "tests", "golden", "program_$name.txt"), String)
        mine = name == "bruss_n4" ? first(MethodOfLines.stencil_program(brusselator_system(4)...)) :
               first(MethodOfLines.stencil_program(heat, MOLFiniteDifference([x => 0.01], t)))
        count(k, s) = length(collect(eachmatch(Regex("^" * k * " ", "m"), s)))
        for k in ("tab", "core", "row", "ghost", "eq", "interior", "periodic", "corebox")
            @test count(k, mine) == count(k, text)
        end
    end

    @parameters t x
    @variables q(..)
    adv = PDESystem([Differential(t)(q(t, x)) ~ -Differential(x)(q(t, x))], [q(0, x) ~ sinpi(x), q(t, 0) ~ q(t, 2)],
                    [t ∈ Interval(0.0, 1.0), x ∈ Interval(0.0, 2.0)], [t, x], [q(t, x)]; name = :adv)
    weno = MOLFiniteDifference([x => 0.02], t; advection_scheme = WENOScheme(), discretization_strategy = MethodOfLines.CudaStencilDiscretization())
    @test_logs (:info, r"falling back to ScalarizedDiscretization") match_mode = :any discretize(adv, weno)
    strict = MOLFiniteDifference([x => 0.02], t; advection_scheme = WENOScheme(),
                                 discretization_strategy = MethodOfLines.CudaStencilDiscretization(strict = true))
    @test_throws MethodOfLines.StencilUnsupported discretize(adv, strict)
end
