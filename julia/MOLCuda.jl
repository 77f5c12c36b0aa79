# MOLCuda.jl — the reference-side binding of libmol_cuda.so (include/mol_cuda.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia (SURVEY §0-4).  This file is the
# stub a MethodOfLines.jl maintainer adds; it binds exactly the entry points the Python `ctypes`
# driver (methodoflines.jl_b200/capi.py) binds and that tests/ exercise on the B200.
#
# Seams used (paths in SciML/MethodOfLines.jl):
#   * a new strategy type next to ScalarizedDiscretization/ArrayDiscretization
#       src/interface/disc_strategy_types.jl:3
#   * whitelist it in interface_errors                      src/MOL_discretization.jl:14-22
#   * SciMLBase.discretize override returning a hand-built ODEProblem
#       (precedent: src/discretization/staggered_discretize.jl:1-29)
#   * PDEBase.discretize_equation! receives interiormap / bcmap / derivweights / DiscreteSpace
#       (src/scalar_discretization.jl:1-5, src/array_discretization.jl:61-65): everything the stencil
#       program needs.  The serializer is the Julia twin of methodoflines.jl_b200/lowering.py.
module MOLCuda

using CUDA                      # owner of device memory (CuArray) and streams; no kernels come from CUDA.jl
using SparseArrays
import SciMLBase

const libmol = get(ENV, "LIBMOL_CUDA", "libmol_cuda.so")

struct MolError <: Exception
    code::Cint
    msg::String
end
const MOL_E_UNSUPPORTED = Cint(-2)      # pattern outside the kernel set: the caller falls back (julia/MOLCudaStencil.jl)
last_error() = unsafe_string(ccall((:mol_last_error, libmol), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : throw(MolError(rc, last_error()))

# ---- a1: Fornberg weights (fornberg_calculate_weights.jl:20-67) --------------------------------------
function fd_weights(order::Integer, x0::Float64, x::Vector{Float64})
    w = similar(x)
    check(ccall((:mol_fd_weights, libmol), Cint, (Cint, Cdouble, Ptr{Cdouble}, Cint, Ptr{Cdouble}),
                order, x0, x, length(x), w))
    w
end

# one row per node (the per-node loops of the non-uniform tables, centered_diff_weights.jl:94-103): x is n x nrows
# (one window per column, column-major = the library's row-major rows)
function fd_weights_rows(order::Integer, x0::Vector{Float64}, x::Matrix{Float64})
    w = similar(x)
    check(ccall((:mol_fd_weights_rows, libmol), Cint, (Cint, Int64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                order, size(x, 2), size(x, 1), x0, x, w))
    w
end

# ---- plan ----------------------------------------------------------------------------------------------
mutable struct Plan
    h::Ptr{Cvoid}
    function Plan(program::String, device::Integer = CUDA.deviceid(CUDA.device()))
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:mol_plan_create, libmol), Cint, (Cstring, Csize_t, Cint, Ptr{Ptr{Cvoid}}),
                    program, sizeof(program), device, out))
        p = new(out[])
        finalizer(p -> ccall((:mol_plan_destroy, libmol), Cint, (Ptr{Cvoid},), p.h), p)
        p
    end
end
state_len(p::Plan) = Int(ccall((:mol_plan_state_len, libmol), Csize_t, (Ptr{Cvoid},), p.h))
# every kernel variant of one integrator, compiled on several host threads (rk_solve! does this by itself)
precompile!(p::Plan, alg::Symbol) = check(ccall((:mol_plan_precompile, libmol), Cint, (Ptr{Cvoid}, Cint), p.h, ALG[alg]))

# ---- a19: the RHS contract f!(du, u, p, t) (SURVEY §8b): in place, no allocation, u not retained ---------
struct GpuRHS
    plan::Plan
end
function (f::GpuRHS)(du::CuVector{Float64}, u::CuVector{Float64}, p, t)
    ph = p isa AbstractVector{Float64} && !isempty(p) ? pointer(p) : Ptr{Cdouble}(C_NULL)
    check(ccall((:mol_rhs, libmol), Cint,
                (Ptr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cvoid}),
                f.plan.h, pointer(du), pointer(u), ph, Float64(t), CUDA.stream().handle))
    nothing
end

# ---- a20: explicit RK on the device (Tsit5 / SSPRK33 / RK4 / Euler) -------------------------------------
const ALG = Dict(:Euler => 1, :SSPRK33 => 2, :RK4 => 3, :Tsit5 => 4)
struct SolveStats
    t_final::Cdouble; dt_last::Cdouble; nf::Int64; naccept::Int64; nreject::Int64; retcode::Cint
end
function rk_solve!(plan::Plan, u::CuVector{Float64}, tspan, alg::Symbol; abstol = 1e-6, reltol = 1e-3,
                   dt = 0.0, adaptive = alg == :Tsit5, saveat = Float64[], maxiters = 10^6)
    rk = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:mol_rk_init, libmol), Cint, (Ptr{Cvoid}, Cint, Cdouble, Cdouble, Ptr{Ptr{Cvoid}}),
                plan.h, ALG[alg], abstol, reltol, rk))
    save = CUDA.zeros(Float64, length(saveat) * state_len(plan))
    st = Ref{SolveStats}()
    try
        check(ccall((:mol_rk_solve, libmol), Cint,
                    (Ptr{Cvoid}, CuPtr{Cdouble}, Cdouble, Cdouble, Cdouble, Cint, Ptr{Cdouble}, Cint,
                     CuPtr{Cdouble}, Int64, Ptr{SolveStats}, Ptr{Cvoid}),
                    rk[], pointer(u), tspan[1], tspan[2], dt, adaptive, saveat, length(saveat),
                    pointer(save), maxiters, st, CUDA.stream().handle))
    finally
        ccall((:mol_rk_destroy, libmol), Cint, (Ptr{Cvoid},), rk[])
    end
    st[], reshape(save, state_len(plan), :)
end

# ---- solution unpacking on the device: sol[u(t,x)] (src/interface/solution/timedep.jl:30-72) --------------------
# `states` = the saved flat unknown vectors (state_len x nt, e.g. the second return value of rk_solve!);
# returns nodes(1) x nodes(2) x .. x nvar x nt with boundary nodes rebuilt from the boundary conditions and
# invalid corner nodes 0 -- what PDETimeSeriesSolution assembles on the host from `observed`.
function grid_shape(plan::Plan, ndim::Integer)
    n = zeros(Int64, ndim)
    ccall((:mol_plan_grid_len, libmol), Int64, (Ptr{Cvoid}, Ptr{Int64}), plan.h, n)
    Tuple(n)
end
nvar(plan::Plan) = Int(ccall((:mol_plan_nvar, libmol), Cint, (Ptr{Cvoid},), plan.h))
function unpack(plan::Plan, states::CuMatrix{Float64}, ts::Vector{Float64}, ndim::Integer; p = nothing)
    shape = grid_shape(plan, ndim)
    full = CUDA.zeros(Float64, shape..., nvar(plan), length(ts))
    ph = p === nothing ? Ptr{Cdouble}(C_NULL) : pointer(Vector{Float64}(p))
    check(ccall((:mol_unpack, libmol), Cint,
                (Ptr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cdouble}, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
                plan.h, pointer(full), pointer(states), length(ts), ts, ph, CUDA.stream().handle))
    full
end

# ---- Jacobian-vector product jv = J(u, p, t) v (forward-mode differentiation of the generated equations) ------------
# usable as `ODEFunction(...; jvp = ...)` / a `JacobianOperator` for matrix-free Newton-Krylov (TRBDF2, FBDF, KenCarp4)
function jvp!(jv::CuVector{Float64}, plan::Plan, u::CuVector{Float64}, v::CuVector{Float64}, p, t)
    ph = p === nothing ? Ptr{Cdouble}(C_NULL) : pointer(Vector{Float64}(p))
    check(ccall((:mol_jvp, libmol), Cint,
                (Ptr{Cvoid}, CuPtr{Cdouble}, CuPtr{Cdouble}, CuPtr{Cdouble}, Ptr{Cdouble}, Cdouble, Ptr{Cvoid}),
                plan.h, pointer(jv), pointer(u), pointer(v), ph, Float64(t), CUDA.stream().handle))
    jv
end

# ---- Jacobian sparsity pattern (jac_prototype for implicit solvers), read off the stencil program -------------------
function jac_sparsity(plan::Plan)
    nnz = Ref{Int64}(0)
    check(ccall((:mol_plan_jac_sparsity, libmol), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
                plan.h, C_NULL, C_NULL, nnz))
    n = state_len(plan)
    colptr = zeros(Int64, n + 1); rowval = zeros(Int64, nnz[])
    check(ccall((:mol_plan_jac_sparsity, libmol), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}),
                plan.h, colptr, rowval, nnz))
    SparseArrays.SparseMatrixCSC(n, n, colptr .+ 1, rowval .+ 1, ones(Float64, nnz[]))
end

# ---- the strategy, the stencil-program serializer and the `discretize` override live in julia/MOLCudaStencil.jl
# (included into MethodOfLines.jl itself: it walks the package's internal objects).

end # module
