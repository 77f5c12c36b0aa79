# MOLCudaStencil.jl — the Julia side of the drop-in: the discretization strategy, the stencil-program serializer and the
# `discretize` override that put libmol_cuda.so behind `discretize(pdesys, MOLFiniteDifference(...))`.
#
# This file is meant to be `include`d into MethodOfLines.jl next to src/array_discretization.jl (it uses the package's
# internal helpers by their unqualified names) after `include("julia/MOLCuda.jl")` (the ccall binding).  Julia is not
# installed in the build image, so nothing here has been executed; every internal it touches is cited with the reference
# file:line it was read from, and julia/test_molcuda.jl pins it on the reference's own literal RHS
# (docs/src/generated/bruss_code.md, committed as tests/golden/bruss_code_n4.json) and on the program text the Python
# twin emits for the same problems (tests/golden/program_*.txt).
#
# What is serialised (everything else raises StencilUnsupported, which the driver turns into the reference's own
# fallback policy, src/array_discretization.jl:25-59,112-136):
#   grids      center-aligned, uniform or non-uniform (s.grid / s.dxs, discretize_vars.jl:219-251)
#   terms      even-order centred derivatives of any approximation order (derivweights.map, centered_difference.jl:5-57),
#              odd-order derivatives under UpwindScheme with the reference's winding patterns
#              `*(a.., Dx^d(u), b..)`, `/(*(a.., Dx^d(u), b..), c)` and the bare derivative (upwind_difference.jl:200-310),
#              any pointwise expression of the dependent variables, t, the coordinates and the parameters
#   boundaries periodic (same-variable interface), Dirichlet / Neumann / Robin of any derivative order, affine in the
#              boundary value, with data depending on t, parameters and the coordinates along the face
#              (generate_bc_eqs.jl:238-328)
# The directive grammar is DESIGN.md §2 / csrc/mol_parse.cpp; the Python twin is methodoflines.jl_b200/lowering.py.

"""
    CudaStencilDiscretization(; strict = false)

Discretization strategy that lowers each PDE to a stencil program for libmol_cuda.so instead of scalarizing it.
`strict = false` (default): a system outside the kernel set is handed back to `ScalarizedDiscretization()` with an
`@info` line, the way `ArrayDiscretization` falls back to the pointwise path; `strict = true` raises instead
(`StrictArrayDiscretization`'s policy).  Add the type to the whitelist in `interface_errors`
(src/MOL_discretization.jl:14-22).
"""
struct CudaStencilDiscretization <: AbstractDiscretizationStrategy
    strict::Bool
end
CudaStencilDiscretization(; strict = false) = CudaStencilDiscretization(strict)

struct StencilUnsupported <: Exception
    msg::String
end
Base.showerror(io::IO, e::StencilUnsupported) = print(io, "CudaStencilDiscretization: ", e.msg)

# ---- number formatting: C99 hex floats, the exact-round-trip format mol_parse.cpp reads with strtod -------------------
# (written like Python's float.hex so that tests/golden/program_*.txt can be diffed token by token)
function hexfloat(x::Float64)
    x == 0 && return signbit(x) ? "-0x0.0p+0" : "0x0.0p+0"
    isfinite(x) || throw(StencilUnsupported("non-finite number in the stencil program"))
    bits = reinterpret(UInt64, abs(x))
    e = Int((bits >> 52) & 0x7ff)
    m = bits & 0x000fffffffffffff
    lead = e == 0 ? 0 : 1
    ex = e == 0 ? -1022 : e - 1023
    string(x < 0 ? "-" : "", "0x", lead, ".", string(m, base = 16, pad = 13), "p", ex >= 0 ? "+" : "-", abs(ex))
end
hexfloat(x::Real) = hexfloat(Float64(x))

# ---- the program under construction --------------------------------------------------------------------------------
mutable struct StencilProgram
    lines::Vector{String}            # tab / core / row / score / ghost / eq directives in emission order
    tabkeys::Dict{Any, Int}          # (L, lo, hi, rows) -> table id: tables are shared by content
    ntab::Int
    params::Vector{Any}
    xs::Vector{Any}                  # spatial independent variables in layout order
    time::Any
    depvars::Vector{Any}             # dependent variables in state order
    cores::Dict{Int, Tuple{Int, Int}}     # table id -> node range on which its row is translation invariant
    tabdims::Dict{Int, Vector{Int}}       # table id -> dimensions (0-based) along which equations apply it
end

varindex(P::StencilProgram, u) = findfirst(v -> isequal(operation(safe_unwrap(v)), operation(safe_unwrap(u))), P.depvars) - 1
dimindex(P::StencilProgram, x) = findfirst(y -> isequal(safe_unwrap(y), safe_unwrap(x)), P.xs) - 1

# one table of stencil rows: `rowfn(i)` returns (first tap node, weights) for node i of lo:hi.  Tables are shared by
# CONTENT (the Brusselator's four Laplacians -- u and v, x and y -- are one table, as in the Python twin's output).
# Rows that share taps-relative-to-the-node and weights over a contiguous range become the literal `core` row (uniform
# grids; padded with zeros to the row length L, which the parser strips again); rows that share only their taps form a
# `score` range (non-uniform grids: per-node weights stay in the table and the tiled kernel reads the node's own row).
function emit_table!(P::StencilProgram, L::Int, lo::Int, hi::Int, rowfn, dim::Int)
    rows = [rowfn(i) for i in lo:hi]
    key = (L, lo, hi, rows)
    if haskey(P.tabkeys, key)
        return P.tabkeys[key]
    end
    id = P.ntab
    P.ntab += 1
    P.tabkeys[key] = id
    push!(P.lines, "tab $id $L $(hi - lo + 1) $lo")
    offs = [r[1] - i for (r, i) in zip(rows, lo:hi)]
    function longest(pred)          # longest run of consecutive rows that satisfy pred pairwise
        best = (1, 0); a = 1
        for k in 2:(length(rows) + 1)
            if k > length(rows) || !pred(k - 1, k)
                (k - 1 - a) > (best[2] - best[1]) && (best = (a, k - 1))
                a = k
            end
        end
        best
    end
    same(k1, k2) = offs[k1] == offs[k2] && rows[k1][2] == rows[k2][2]
    shape(k1, k2) = offs[k1] == offs[k2] && length(rows[k1][2]) == length(rows[k2][2])
    explicit = trues(length(rows))
    ca, cb = longest(same)
    if cb > ca
        w = vcat(rows[ca][2], zeros(L - length(rows[ca][2])))
        push!(P.lines, "core $id $(lo + ca - 1) $(lo + cb - 1) $(offs[ca]) " * join(hexfloat.(w), " "))
        explicit[ca:cb] .= false
        P.cores[id] = (lo + ca - 1, lo + cb - 1)
    end
    sa, sb = longest(shape)
    if sb > sa
        push!(P.lines, "score $id $(lo + sa - 1) $(lo + sb - 1) $(offs[sa]) $(length(rows[sa][2]))")
        haskey(P.cores, id) || (P.cores[id] = (lo + sa - 1, lo + sb - 1))
    end
    for (k, i) in enumerate(lo:hi)
        explicit[k] || continue
        start, w = rows[k]
        push!(P.lines, "row $id $i $start $(length(w)) " * join(hexfloat.(w), " "))
    end
    return id
end
# a table can serve several dimensions (equal grids); the core box needs its core range per use
notedim!(P::StencilProgram, id::Int, dim::Int) = push!(get!(P.tabdims, id, Int[]), dim)

# ---- row selection: the reference's own branches, per node ------------------------------------------------------------
# central_difference_weights_and_stencil (centered_difference.jl:5-57); `periodic`: both ends are interfaces
function centered_row(D, i::Int, n::Int, periodic::Bool)
    bpc, bsl, L = D.boundary_point_count, D.boundary_stencil_length, D.stencil_length
    uniform = D.dx isa Number
    if i <= bpc && !periodic
        return 1, collect(D.low_boundary_coefs[i])
    elseif i > n - bpc && !periodic
        return n - bsl + 1, collect(D.high_boundary_coefs[n - i + 1])
    end
    w = uniform ? collect(D.stencil_coefs) : collect(D.stencil_coefs[i - bpc])
    return i + first(half_range(L)), w              # taps past a periodic seam are wrapped by the kernels (bwrap)
end

# _upwind_difference (upwind_difference.jl:1-28 uniform, :131-162 non-uniform)
function upwind_row(D, i::Int, n::Int, ispositive::Bool, periodic::Bool)
    L, bsl = D.stencil_length, D.boundary_stencil_length
    uniform = D.dx isa Number
    if !ispositive
        if i > n - D.boundary_point_count && !periodic
            return n - bsl + 1, collect(D.high_boundary_coefs[n - i + 1])
        end
        return i, uniform ? collect(D.stencil_coefs) : collect(D.stencil_coefs[i])
    end
    if i <= D.offside && !periodic
        return 1, collect(D.low_boundary_coefs[i])
    end
    return i - L + 1, uniform ? collect(D.stencil_coefs) : collect(D.stencil_coefs[i - D.offside])
end

# ---- expressions -> RPN tokens (the `arrayify` of this strategy, array_discretization.jl:640-670) ------------------------
struct RpnContext
    P::StencilProgram
    rules::Vector{Pair{Any, Vector{String}}}     # ordered: first match wins, as in ArrayifyContext
    fieldsok::Bool
end

const RPN_UNARY = Dict{Any, String}(
    sqrt => "sqrt", exp => "exp", log => "log", sin => "sin", cos => "cos", tan => "tan", sinh => "sinh",
    cosh => "cosh", tanh => "tanh", abs => "abs", asin => "asin", acos => "acos", atan => "atan", sign => "sign")
const RPN_CMP = Dict{Any, String}((>) => "gt", (>=) => "ge", (<) => "lt", (<=) => "le", (==) => "eq", (!=) => "ne")

function rpnify(expr, ctx::RpnContext)::Vector{String}
    expr = safe_unwrap(expr)
    for (k, v) in ctx.rules
        isequal(expr, k) && return v
    end
    v = unwrap_const(expr)
    v isa Bool && return ["c:" * hexfloat(v ? 1.0 : 0.0)]
    v isa Number && return ["c:" * hexfloat(Float64(v))]
    P = ctx.P
    if !iscall(expr)                      # a symbol: time, a coordinate or a parameter
        isequal(expr, safe_unwrap(P.time)) && return ["t"]
        j = findfirst(y -> isequal(safe_unwrap(y), expr), P.xs)
        j === nothing || return ["x:$(j - 1)"]
        k = findfirst(p -> isequal(safe_unwrap(p), expr), P.params)
        k === nothing || return ["p:$(k - 1)"]
        throw(StencilUnsupported("unknown symbol $expr"))
    end
    op = operation(expr)
    args = arguments(expr)
    if op isa Differential
        throw(StencilUnsupported("derivative term without a scheme here: $expr"))
    elseif any(u -> isequal(op, operation(safe_unwrap(u))), P.depvars)
        ctx.fieldsok || throw(StencilUnsupported("field value inside boundary data: $expr"))
        any(a -> unwrap_const(safe_unwrap(a)) isa Number, args) &&
            throw(StencilUnsupported("boundary value $expr in an interior equation"))
        return ["u:$(varindex(P, expr))"]
    elseif op === (+) || op === (*)
        toks = rpnify(args[1], ctx)
        for a in args[2:end]
            append!(toks, rpnify(a, ctx))
            push!(toks, op === (+) ? "+" : "*")
        end
        return toks
    elseif op === (-)
        length(args) == 1 && return vcat(rpnify(args[1], ctx), ["neg"])
        return vcat(rpnify(args[1], ctx), rpnify(args[2], ctx), ["-"])
    elseif op === (/)
        return vcat(rpnify(args[1], ctx), rpnify(args[2], ctx), ["/"])
    elseif op === (^)
        e = unwrap_const(safe_unwrap(args[2]))
        if e isa Integer || (e isa Real && isinteger(e))
            return vcat(rpnify(args[1], ctx), ["powi:$(Int(e))"])
        end
        return vcat(rpnify(args[1], ctx), rpnify(args[2], ctx), ["pow"])
    elseif op === ifelse
        return vcat(rpnify(args[1], ctx), rpnify(args[2], ctx), rpnify(args[3], ctx), ["sel"])
    elseif op === (&) || op === (|)
        return vcat(rpnify(args[1], ctx), rpnify(args[2], ctx), [op === (&) ? "and" : "or"])
    elseif op === (!)
        return vcat(rpnify(args[1], ctx), ["not"])
    elseif op === min || op === max
        return vcat(rpnify(args[1], ctx), rpnify(args[2], ctx), [op === min ? "min" : "max"])
    elseif haskey(RPN_CMP, op)
        return vcat(rpnify(args[1], ctx), rpnify(args[2], ctx), [RPN_CMP[op]])
    elseif haskey(RPN_UNARY, op)
        return vcat(rpnify(args[1], ctx), [RPN_UNARY[op]])
    end
    throw(StencilUnsupported("unhandled operation $op in $expr"))
end

# ---- one PDE -> tables + `eq` (called with exactly what discretize_equation! receives, array_discretization.jl:61-65) ----
function periodic_dims(s, u, bcmap)
    map(ivs(u, s)) do x
        bs = filter_interfaces(bcmap[operation(u)][x])
        isempty(bs) && return false
        all(haslowerupper(bs, x)) || throw(StencilUnsupported("interface boundary at one end of $x only"))
        for b in bs
            (isequal(b.x, x) && isequal(b.x2, x) && isequal(depvar(b.u, s), depvar(u, s)) &&
                isequal(depvar(b.u2, s), depvar(u, s))) ||
                throw(StencilUnsupported("interface boundary $(b.eq) joins different variables"))
        end
        true
    end
end

function stencil_equation!(P::StencilProgram, pde, interiormap, eqvar, bcmap, depvars, s, derivweights, indexmap,
        discretization)
    get_grid_type(s) <: CenterAlignedGrid || throw(StencilUnsupported("only center-aligned grids are serialised"))
    derivweights.advection_scheme isa UpwindScheme || throw(StencilUnsupported("only UpwindScheme advection is serialised"))
    args = ivs(eqvar, s)
    for u in depvars
        isequal(ivs(u, s), args) || throw(StencilUnsupported("variables of differing dimensionality"))
    end
    # special schemes (mixed derivatives, nonlinear / spherical Laplacian, integrals, callbacks): probe the reference's
    # own rule generators at an interior point exactly as discretize_equation_array_form does (:196-219)
    interior = interiormap.I[pde]
    length(interior) == 0 && throw(StencilUnsupported("equation without spatial extent"))
    II0 = first(interior)
    terms = split_terms(pde, s.x̄)
    special = vcat(
        vec(generate_mixed_rules(II0, s, depvars, derivweights, bcmap, indexmap, terms)),
        vec(generate_nonlinlap_rules(II0, s, depvars, derivweights, bcmap, indexmap, terms)),
        vec(generate_spherical_diffusion_rules(II0, s, depvars, derivweights, bcmap, indexmap, split_additive_terms(pde))),
        vec(generate_euler_integration_rules(II0, s, depvars, indexmap, terms)),
        vec(generate_whole_domain_integration_rules(II0, s, depvars, indexmap, terms)),
        vec(generate_cb_rules(II0, s, depvars, derivweights, bcmap, indexmap, terms)))
    for r in special
        (subsmatch(pde.lhs, r) || subsmatch(pde.rhs, r)) && throw(StencilUnsupported("unsupported pattern $(r.first)"))
    end

    v = varindex(P, eqvar)
    lo, hi = Tuple(first(interior)), Tuple(last(interior))
    pdeorders = Dict(x => d_orders(x, [pde]) for x in args)
    rules = Pair{Any, Vector{String}}[]
    # centred rules for even orders (array_cartesian_rules, :556-571)
    for u in depvars, x in ivs(depvar(u, s), s)
        per = periodic_dims(s, depvar(u, s), bcmap)[indexmap[x]]
        j, n, uv = dimindex(P, x), length(s, x), varindex(P, u)
        for d in filter(iseven, pdeorders[x])
            D = derivweights.map[Differential(x)^d]
            id = emit_table!(P, max(D.stencil_length, D.boundary_stencil_length), lo[j + 1], hi[j + 1],
                             i -> centered_row(D, i, n, per), j)
            notedim!(P, id, j)
            push!(rules, safe_unwrap((Differential(x)^d)(u)) => ["L:$id:$uv:$j"])
        end
    end
    base = RpnContext(P, copy(rules), true)
    # winding rules for odd orders (array_winding_rules, :609-660): ifelse(coef > 0, coef * backward, coef * forward)
    function winding(coef, u, x, d)
        per = periodic_dims(s, depvar(u, s), bcmap)[indexmap[x]]
        j, n, uv = dimindex(P, x), length(s, x), varindex(P, u)
        Dm, Dp = derivweights.windmap[2][Differential(x)^d], derivweights.windmap[1][Differential(x)^d]
        idm = emit_table!(P, max(Dm.stencil_length, Dm.boundary_stencil_length), lo[j + 1], hi[j + 1],
                          i -> upwind_row(Dm, i, n, true, per), j)
        notedim!(P, idm, j)
        coef === nothing && return ["L:$idm:$uv:$j"]          # bare derivative: positive winding (:289-301)
        idp = emit_table!(P, max(Dp.stencil_length, Dp.boundary_stencil_length), lo[j + 1], hi[j + 1],
                          i -> upwind_row(Dp, i, n, false, per), j)
        notedim!(P, idp, j)
        c = rpnify(coef, base)
        return vcat(c, ["c:" * hexfloat(0.0), "gt"], c, ["L:$idm:$uv:$j", "*"], c, ["L:$idp:$uv:$j", "*"], ["sel"])
    end
    windrules = Pair{Any, Vector{String}}[]
    for u in depvars, x in ivs(depvar(u, s), s), d in filter(isodd, pdeorders[x])
        r1 = @rule *(~~a, $(Differential(x)^d)(u), ~~b) => winding(*(~a..., ~b...), u, x, d)
        r2 = @rule /(*(~~a, $(Differential(x)^d)(u), ~~b), ~c) => winding(*(~a..., ~b...) / ~c, u, x, d)
        for t in terms, r in (r1, r2)
            w = r(t)
            w === nothing || push!(windrules, safe_unwrap(t) => w)
        end
        push!(windrules, safe_unwrap((Differential(x)^d)(u)) => winding(nothing, u, x, d))
    end
    ctx = RpnContext(P, vcat(windrules, rules), true)
    # cardinalised residual lhs - rhs ~ 0 with D_t(u) on the left: du = -(residual without the time derivative) / its factor
    Dt = Differential(P.time)
    resid = pde.lhs - pde.rhs
    dtu = Dt(eqvar)
    a, _, islin = Symbolics.linear_expansion(resid, dtu)
    islin || throw(StencilUnsupported("equation is not linear in the time derivative of $eqvar"))
    # the spatial part by substitution, not from the expansion: the winding rules are keyed by whole terms of the
    # equation (split_terms), which a re-assembled expression need not preserve
    b = substitute(resid, Dict(dtu => 0))
    toks = vcat(rpnify(b, ctx), ["neg"])
    isequal(unwrap_const(safe_unwrap(a)), 1) || (toks = vcat(toks, rpnify(a, ctx), ["/"]))
    push!(P.lines, "eq $v $(length(toks)) " * join(toks, " "))
    return nothing
end

# ---- boundary conditions -> ghost rules (generate_bc_eqs.jl:238-328: u(t, x_b) -> the edge node, Dx^d u(t, x_b) -> the
# one-sided row of the centred operator at the edge node), solved for the edge node: u[node] = g + sum_k a_k u[tap_k] ------
function stencil_ghosts!(P::StencilProgram, eqvar, bcmap, s, derivweights)
    u = depvar(eqvar, s)
    v = varindex(P, u)
    for x in ivs(u, s), b in bcmap[operation(u)][x]
        b isa AbstractTruncatingBoundary || continue            # periodic seams are wrapped, not ghosted
        j, n = dimindex(P, x), length(s, x)
        node = isupper(b) ? n : 1
        ξ = Symbolics.variable(:molghost)
        taps = Dict{Int, Any}()
        sub = Dict{Any, Any}(safe_unwrap(b.u) => ξ)
        for d in derivweights.orders[x]
            D = derivweights.map[Differential(x)^d]
            start, w = centered_row(D, node, n, false)
            acc = 0
            for (k, wk) in enumerate(w)
                tp = start + k - 1
                sym = tp == node ? ξ : get!(taps, tp, Symbolics.variable(:moltap, tp))
                acc += wk * sym
            end
            sub[safe_unwrap((Differential(x)^d)(b.u))] = acc
        end
        resid = substitute(b.eq.lhs - b.eq.rhs, sub)
        a, rest, islin = Symbolics.linear_expansion(resid, ξ)
        islin || throw(StencilUnsupported("boundary condition $(b.eq) is not affine in the boundary value"))
        coefs = Pair{Int, Float64}[]
        for (tp, sym) in sort(collect(taps), by = first)
            c, rest, lin = Symbolics.linear_expansion(rest, sym)
            lin || throw(StencilUnsupported("boundary condition $(b.eq) is not affine in the field"))
            cv = unwrap_const(safe_unwrap(-c / a))
            cv isa Number || throw(StencilUnsupported("expression-valued Robin coefficients (ghostx) are not serialised yet"))
            push!(coefs, tp => Float64(cv))
        end
        g = rpnify(-rest / a, RpnContext(P, Pair{Any, Vector{String}}[], false))
        push!(P.lines, "ghost $v $j $node $(length(coefs)) " * join(["$v $tp $(hexfloat(c))" for (tp, c) in coefs], " ") *
                       " $(length(g)) " * join(g, " "))
    end
    return nothing
end

# ---- the whole program ------------------------------------------------------------------------------------------------
function stencil_program_text(P::StencilProgram, s, interiormap, pdes, bcmap, p_defaults)
    io = IOBuffer()
    println(io, "MOLPROG 1")
    println(io, "ndim $(length(P.xs))")
    println(io, "nvar $(length(P.depvars))")
    println(io, "nparam $(length(P.params))")
    for (k, p) in enumerate(P.params)
        println(io, "param $(k - 1) $p $(hexfloat(p_defaults[k]))")
    end
    for (j, x) in enumerate(P.xs)
        g = collect(Float64, s.grid[x])
        uniform = s.dxs[x] isa Number
        println(io, "grid $(j - 1) $(length(g)) $(uniform ? "U" : "N") $(hexfloat(uniform ? Float64(s.dxs[x]) : 0.0))")
        println(io, "coords $(j - 1) " * join(hexfloat.(g), " "))
    end
    los = Dict{Int, Any}(); his = Dict{Int, Any}()
    for pde in pdes
        u = interiormap.var[pde]
        v = varindex(P, u)
        I = interiormap.I[pde]
        los[v], his[v] = Tuple(first(I)), Tuple(last(I))
        println(io, "var $v $(operation(safe_unwrap(u)))")
        println(io, "interior $v " * join(vcat(collect(los[v]), collect(his[v])), " "))
        println(io, "periodic $v " * join(Int.(periodic_dims(s, depvar(u, s), bcmap)), " "))
    end
    foreach(l -> println(io, l), P.lines)
    # core box: where every node-indexed table is on its core (or shape-core) row (array_bands, :368-420)
    if all(v -> los[v] == los[0] && his[v] == his[0], keys(los))
        clo, chi = collect(los[0]), collect(his[0])
        ok = true
        for (id, dims) in P.tabdims, j in dims
            haskey(P.cores, id) || (ok = false; break)
            a, b = P.cores[id]
            clo[j + 1], chi[j + 1] = max(clo[j + 1], a), min(chi[j + 1], b)
        end
        ok && all(chi .>= clo) && println(io, "corebox " * join(vcat(clo, chi), " "))
    end
    println(io, "end")
    return String(take!(io))
end

"""
    stencil_program(pdesys, discretization) -> (program::String, u0::Vector{Float64}, tspan, p::Vector{Float64})

Replays the front half of `PDEBase.symbolic_discretize` exactly as the package's own `get_discrete` does
(src/MOL_discretization.jl:102-143: cardinalize, `VariableMap`, `parse_bcs`, `check_boundarymap`,
`construct_discrete_space`), continues with the two hooks MethodOfLines.jl implements for the rest of that pipeline
(`construct_differential_discretizer` differential_discretizer.jl:14, `construct_var_equation_mapping`
interior_map.jl:53) and, per equation, serialises instead of scalarizing.
"""
function stencil_program(pdesys::PDESystem, discretization::MOLFiniteDifference)
    t = get_time(discretization)
    t === nothing && throw(StencilUnsupported("steady-state problems are not an explicit-RK path"))
    PDEBase.cardinalize_eqs!(pdesys)
    v = VariableMap(pdesys, discretization)
    PDEBase.interface_errors(pdesys, v, discretization)
    bcorders = Dict(map(x -> x => d_orders(x, get_bcs(pdesys)), all_ivs(v)))
    boundarymap = PDEBase.parse_bcs(get_bcs(pdesys), v, bcorders)
    PDEBase.check_boundarymap(boundarymap, v, discretization)
    should_transform(pdesys, discretization, boundarymap) &&
        throw(StencilUnsupported("system needs the auxiliary-variable transformation (nonlinear Laplacian with mixed terms)"))
    pdes = get_eqs(pdesys)
    s = PDEBase.construct_discrete_space(v, discretization)
    derivweights = PDEBase.construct_differential_discretizer(pdesys, s, discretization, bcorders)
    interiormap = PDEBase.construct_var_equation_mapping(pdes, boundarymap, s, discretization)
    # parameters: `pdesys.ps` as the tutorials write it, `[p => value, ...]` (pde_system_transformation.jl:37 reads the same field)
    pspec = pdesys.ps isa AbstractVector ? pdesys.ps : []
    ps = [safe_unwrap(p isa Pair ? first(p) : p) for p in pspec]
    pvals = Float64[p isa Pair ? Float64(unwrap_const(safe_unwrap(last(p)))) : 0.0 for p in pspec]
    # state order = the order in which equations were matched to variables, variable-major (SURVEY a19)
    depvars_in_order = [interiormap.var[pde] for pde in pdes]
    P = StencilProgram(String[], Dict{Any, Int}(), 0, ps, collect(s.x̄), t, depvars_in_order, Dict{Int, Tuple{Int, Int}}(),
                       Dict{Int, Vector{Int}}())
    for pde in pdes
        eqvar = interiormap.var[pde]
        depvars = collect(filter(u -> !any(x -> unwrap_const(safe_unwrap(x)) isa Number, arguments(u)),
                                 get_depvars(pde.lhs, s.vars.depvar_ops) ∪ get_depvars(pde.rhs, s.vars.depvar_ops)))
        args = ivs(eqvar, s)
        indexmap = Dict([args[i] => i for i in 1:length(args)])
        stencil_equation!(P, pde, interiormap, eqvar, boundarymap, depvars, s, derivweights, indexmap, discretization)
        stencil_ghosts!(P, eqvar, boundarymap, s, derivweights)
    end
    text = stencil_program_text(P, s, interiormap, pdes, boundarymap, pvals)
    # u0: the initial conditions (boundaries in t) evaluated on the interior nodes, x fastest, variable-major
    u0 = Float64[]
    tspan = (Float64(v.intervals[t][1]), Float64(v.intervals[t][2]))
    for pde in pdes
        u = interiormap.var[pde]
        ic = only(boundarymap[operation(u)][t])
        f = Symbolics.build_function(ic.eq.rhs, ivs(u, s)...; expression = Val{false})
        for II in interiormap.I[pde]
            push!(u0, Float64(f((s.grid[x][II[k]] for (k, x) in enumerate(ivs(u, s)))...)))
        end
    end
    return text, u0, tspan, pvals
end

# ---- the override (precedent: staggered_discretize.jl:1-29) ------------------------------------------------------------
function SciMLBase.discretize(pdesys::PDESystem, discretization::MOLFiniteDifference{G, CudaStencilDiscretization};
        kwargs...) where {G}
    strat = discretization.disc_strategy
    try
        text, u0, tspan, p = stencil_program(pdesys, discretization)
        plan = MOLCuda.Plan(text)                       # mol_plan_create: MOL_E_UNSUPPORTED -> MolError below
        f = SciMLBase.ODEFunction{true}(MOLCuda.GpuRHS(plan))
        return SciMLBase.ODEProblem(f, CUDA.CuArray(u0), tspan, p; kwargs...)
    catch e
        unsupported = e isa StencilUnsupported || (e isa MOLCuda.MolError && e.code == MOLCuda.MOL_E_UNSUPPORTED)
        (unsupported && !strat.strict) || rethrow(e)
        # the reference's policy (array_discretization.jl:112-136): never turn a system the scalar path can discretize
        # into an error -- hand it back to the pointwise strategy, on the CPU, and say so
        @info "CudaStencilDiscretization: falling back to ScalarizedDiscretization" reason = sprint(showerror, e)
        scalar = MOLFiniteDifference(discretization.dxs, discretization.time;
            approx_order = discretization.approx_order, advection_scheme = discretization.advection_scheme,
            grid_align = discretization.grid_align, discretization_strategy = ScalarizedDiscretization(),
            discretization.kwargs...)
        return SciMLBase.discretize(pdesys, scalar; kwargs...)
    end
end
