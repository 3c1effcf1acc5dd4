# B200DiffEqFlux.jl -- the Julia side of the drop-in boundary (include/b200node.h).
#
# WRITE-ONLY in this repository: the build image has no Julia toolchain, so this file is the binding a
# DiffEqFlux maintainer would add; the Python ctypes layer (diffeqflux.jl_b200/layers.py) exercises
# the same C ABI with the same argument conventions and is what the tests run.
#
# It re-declares the hot-path layer types of DiffEqFlux (src/neural_de.jl:35-61,372-412, src/ffjord.jl:52-245,
# src/multiple_shooting.jl:36-94,231-243) so that
#     node = NeuralODE(model, tspan, Tsit5(); saveat, sensealg);  sol, st = node(u0, ps, st)
#     Zygote.gradient(sum ∘ first ∘ node, u0, ps, st)
# keep working, with ONE ccall per solve and ONE ccall per pullback.  Every layer has its ChainRules `rrule`
# (Zygote never differentiates through a ccall).  Unsupported configurations (non-Tsit5 solver, non-Dense model,
# Float64, other sensealgs, FFJORD with a basedist or AutoForwardDiff) throw: there is no CPU fallback.
module B200DiffEqFlux

using ChainRulesCore, Lux, LuxCore, ComponentArrays, SciMLBase, Libdl, Random

export NeuralODE, AugmentedNDELayer, FFJORD, FFJORDDistribution, DimMover, B200RHS, multiple_shoot, group_ranges,
       train_step!, AdamState

const LIB = Ref{Ptr{Cvoid}}(C_NULL)
lib() = (LIB[] == C_NULL && (LIB[] = Libdl.dlopen(get(ENV, "B200NODE_LIB", "libb200node.so"))); LIB[])
sym(s) = Libdl.dlsym(lib(), s)

# ---- mirrors of the C structs (include/b200node.h) -------------------------------------------------
Base.@kwdef mutable struct B200Opts
    abstol::Cfloat = 1.0f-6
    reltol::Cfloat = 1.0f-3
    dt0::Cfloat = 0.0f0
    maxiters::Cint = 100000
    save_start::Cint = 1
    save_end::Cint = 1
    sensealg::Cint = 0          # -1 none, 0 Interpolating, 1 Backsolve, 2 Quadrature, 3 Gauss
    ptr_kind::Cint = 0          # 0 host (Array), 1 device (CuArray pointer)
    tile::Cint = 0
    max_dense_steps::Cint = 0
    seed::UInt64 = 0
    noise_per_call::Cint = 1
    b_offset::Cint = 0
    save_everystep::Cint = 0
    path::Cint = 0              # 0 auto, 1 FFMA only, 2 require tcgen05
    ffjord_exact_trace::Cint = 0    # FFJORD monte_carlo = false
    ffjord_regularize::Cint = 0     # FFJORD regularize = true (two more state rows)
end

mutable struct B200Stats
    retcode::Cint; naccept::Cint; nreject::Cint; nf::Cint
    t_final::Cfloat; dt_last::Cfloat; tie_dist::Cfloat; kernel_ms::Cfloat
    n_out::Cint; grid::Cint; tile::Cint; engine::Cint
    reserved::NTuple{4, Cint}
    B200Stats() = new(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, (0, 0, 0, 0))
end

const CTX = Ref{Ptr{Cvoid}}(C_NULL)
function ctx()
    if CTX[] == C_NULL
        out = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall(sym(:b200_ctx_create), Cint, (Cint, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), 0, C_NULL, out)
        rc == 0 || error("b200_ctx_create failed ($rc): no usable CUDA device; there is no CPU fallback")
        CTX[] = out[]
    end
    return CTX[]
end
check(rc) = rc < 0 ? error(unsafe_string(ccall(sym(:b200_last_error), Cstring, (Ptr{Cvoid},), ctx()))) : rc

# the opaque forward record: freed when the last Julia reference (solution / pullback closure) is collected
mutable struct Tape
    h::Ptr{Cvoid}
    function Tape(h)
        t = new(h)
        finalizer(x -> (x.h != C_NULL && ccall(sym(:b200_tape_free), Cvoid, (Ptr{Cvoid},), x.h); x.h = C_NULL), t)
    end
end

const ACT = IdDict{Any, Cint}(identity => 0, tanh => 1, Lux.relu => 2, Lux.sigmoid => 3, Lux.softplus => 4)
const MODELS = Dict{Tuple{Vector{Cint}, Vector{Cint}, Cint}, Ptr{Cvoid}}()     # one b200_model per architecture (never re-created)

# Chain([x -> x .^ 3,] Dense...) -> b200_model (anything else is outside the B200 path)
function model_handle(model::Lux.Chain)
    layers = collect(values(model.layers))
    pre = Cint(0)
    if !isempty(layers) && layers[1] isa Lux.WrappedFunction      # docs/src/examples/neural_ode.md:32 `x -> x .^ 3`
        layers[1].func([2.0f0]) == [8.0f0] || error("B200 path: the only supported pre-layer is x -> x .^ 3")
        pre = Cint(1); layers = layers[2:end]
    end
    all(l -> l isa Lux.Dense, layers) || error("B200 path: model must be a Chain of Dense layers")
    dims = Cint[layers[1].in_dims; [l.out_dims for l in layers]]
    acts = Cint[get(() -> error("B200 path: unsupported activation $(l.activation)"), ACT, l.activation) for l in layers]
    return get!(MODELS, (dims, acts, pre)) do
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall(sym(:b200_model_create), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cint}, Cint, Ptr{Ptr{Cvoid}}),
                    ctx(), length(layers), dims, acts, pre, out))
        out[]
    end
end

sense_code(sa) = sa === nothing ? Cint(0) :
    (n = nameof(typeof(sa)); n == :InterpolatingAdjoint ? Cint(0) : n == :BacksolveAdjoint ? Cint(1) :
     n == :QuadratureAdjoint ? Cint(2) : n == :GaussAdjoint ? Cint(3) : error("B200 path: unsupported sensealg $(typeof(sa))"))

# ---- NeuralODE (src/neural_de.jl:35-61) -------------------------------------------------------------
struct NeuralODE{M <: LuxCore.AbstractLuxLayer, T, A, K} <: LuxCore.AbstractLuxWrapperLayer{:model}
    model::M
    tspan::T
    args::A
    kwargs::K
end
NeuralODE(model, tspan, args...; kwargs...) = NeuralODE(model, tspan, args, kwargs)

struct B200Solution{T, U}          # what callers touch of an ODESolution: .u, .t, .retcode, Array(sol), sol[end]
    t::Vector{T}
    u::Vector{U}
    retcode::SciMLBase.ReturnCode.T
    stats::B200Stats
end
Base.Array(s::B200Solution) = cat(s.u...; dims = ndims(s.u[1]) + 1)
Base.last(s::B200Solution) = last(s.u)
Base.first(s::B200Solution) = first(s.u)
Base.getindex(s::B200Solution, i) = s.u[i]
Base.lastindex(s::B200Solution) = lastindex(s.u)

const RETCODES = (SciMLBase.ReturnCode.Success, SciMLBase.ReturnCode.MaxIters,
                  SciMLBase.ReturnCode.DtLessThanMin, SciMLBase.ReturnCode.Unstable)

function split_saveat(saveat, t0::Float32, t1::Float32, kw)
    h(s) = t1 >= t0 ? Float32(abs(s)) : -Float32(abs(s))
    sv = saveat === nothing ? Float32[] : saveat isa Number ? collect(t0:h(saveat):t1) : Float32.(collect(saveat))
    everystep = get(kw, :save_everystep, isempty(sv))
    ss = get(kw, :save_start, everystep || isempty(sv) || saveat isa Number || t0 in sv)
    se = get(kw, :save_end, everystep || isempty(sv) || saveat isa Number || t1 in sv)
    lo, hi = minmax(t0, t1)
    return sort(filter(t -> lo < t < hi, sv); rev = t1 < t0), ss, se, everystep && isempty(sv)
end

function forward(n::NeuralODE, x::AbstractVecOrMat{Float32}, p::AbstractVector{Float32}; want_tape::Bool, augment::Int = 0)
    length(n.args) == 1 && nameof(typeof(n.args[1])) == :Tsit5 || error("B200 path: only Tsit5() is supported")
    kw = n.kwargs
    t0, t1 = Float32.(n.tspan)
    interior, ss, se, everystep = split_saveat(get(kw, :saveat, nothing), t0, t1, kw)
    # src/neural_de.jl:57: InterpolatingAdjoint(autojacvec = ZygoteVJP()) unless the caller passes a sensealg
    sense = sense_code(get(kw, :sensealg, nothing))
    everystep && sense == 1 && error("B200 path: save_everystep with BacksolveAdjoint needs saveat")
    o = B200Opts(; abstol = get(kw, :abstol, 1.0f-6), reltol = get(kw, :reltol, 1.0f-3), save_start = ss, save_end = se,
                 sensealg = want_tape ? sense : (everystep ? Cint(0) : Cint(-1)), save_everystep = everystep,
                 maxiters = get(kw, :maxiters, 100000))
    d_in, B = size(x, 1), size(x, 2)
    D = d_in + augment
    n_out = everystep ? 0 : ss + length(interior) + se
    out_u = Array{Float32}(undef, D, B, n_out)
    out_t = Vector{Float32}(undef, n_out)
    st = B200Stats()
    tape = Ref{Ptr{Cvoid}}(C_NULL)
    need_tape = want_tape || everystep            # the accepted steps of a save_everystep solve live on the tape
    rc = check(ccall(sym(:b200_node_forward), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Cint, Cfloat, Cfloat, Ptr{Cfloat}, Cint,
                      Ref{B200Opts}, Ptr{Cfloat}, Ptr{Cfloat}, Ref{B200Stats}, Ptr{Ptr{Cvoid}}),
                     ctx(), model_handle(n.model), p, x, d_in, augment, B, t0, t1, interior, length(interior),
                     o, out_u, out_t, st, need_tape ? tape : C_NULL))
    tp = need_tape ? Tape(tape[]) : nothing
    if everystep
        ns = ccall(sym(:b200_tape_nsteps), Cint, (Ptr{Cvoid},), tp.h)
        out_t = Vector{Float32}(undef, ns + 1)
        out_u = Array{Float32}(undef, D, B, ns + 1)
        check(ccall(sym(:b200_tape_steps), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Cint), ctx(), tp.h, out_t, out_u, 0))
        keep = (ss ? 1 : 2):(se || ns == 0 ? ns + 1 : ns)
        out_t, out_u = out_t[keep], out_u[:, :, keep]
    else
        out_t, out_u = out_t[1:st.n_out], out_u[:, :, 1:st.n_out]      # an aborted solve saved fewer points
    end
    u = [x isa AbstractVector ? out_u[:, 1, i] : out_u[:, :, i] for i in axes(out_u, 3)]
    return B200Solution(out_t, u, RETCODES[rc + 1], st), (want_tape ? tp : nothing), o
end

# cotangent of `sol` -> dense (D, B, T) array (Zygote hands back a Tangent with a `u` field, or a plain array for Array(sol))
function cotangent_array(Δsol, sol)
    dL = zeros(Float32, size(sol.u[1], 1), size(sol.u[1], 2), length(sol.u))
    du = Δsol isa AbstractArray ? [selectdim(Δsol, ndims(Δsol), i) for i in axes(Δsol, ndims(Δsol))] : Δsol.u
    for i in eachindex(sol.u)
        du[i] isa AbstractZero || (selectdim(dL, 3, i) .= reshape(du[i], size(dL, 1), size(dL, 2)))
    end
    return dL
end

function adjoint(model, tape::Tape, o, dL, d_in, B, nparams)
    du0 = Matrix{Float32}(undef, d_in, B)
    dp = Vector{Float32}(undef, nparams)
    check(ccall(sym(:b200_node_adjoint), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Ref{B200Opts}, Ptr{Cfloat}, Ptr{Cfloat}, Ref{B200Stats}),
                ctx(), model_handle(model), tape.h, dL, o, du0, dp, B200Stats()))
    return du0, dp
end

# `(n::NeuralODE)(x, p, st)` (src/neural_de.jl:47-61)
(n::NeuralODE)(x, p, st) = (forward(n, x, collect(p); want_tape = false)[1], st)

# the rrule Zygote finds instead of SciMLSensitivity's (src/DiffEqFlux.jl:17-22)
function ChainRulesCore.rrule(n::NeuralODE, x, p, st; augment::Int = 0)
    pv = collect(p)
    sol, tape, o = forward(n, x, pv; want_tape = true, augment)
    function node_pullback(Δ)
        du0, dp = adjoint(n.model, tape, o, cotangent_array(first(Δ), sol), size(x, 1), size(x, 2), length(pv))
        return NoTangent(), (x isa AbstractVector ? vec(du0) : du0), ComponentArray(dp, getaxes(p)), NoTangent()
    end
    return (sol, st), node_pullback
end

# ---- AugmentedNDELayer (src/neural_de.jl:372-386): the zero padding is fused into the kernel ----------
struct AugmentedNDELayer{N <: NeuralODE} <: LuxCore.AbstractLuxWrapperLayer{:nde}
    nde::N
    adim::Int
end
(a::AugmentedNDELayer)(x, p, st) = (forward(a.nde, x, collect(p); want_tape = false, augment = a.adim)[1], st)
function ChainRulesCore.rrule(a::AugmentedNDELayer, x, p, st)       # no gradient reaches the zero rows
    y, back = ChainRulesCore.rrule(a.nde, x, p, st; augment = a.adim)
    return y, Δ -> back(Δ)
end

# ---- DimMover (src/neural_de.jl:397-412) ---------------------------------------------------------------
struct DimMover <: LuxCore.AbstractLuxLayer
    from::Int
    to::Int
end
DimMover(; from = -2, to = -1) = (@assert from !== 0 && to !== 0; DimMover(from, to))
function (dm::DimMover)(x::Array{Float32}, ps, st)
    out = Vector{Float32}(undef, length(x))
    nd = Ref{Cint}(0); od = zeros(Cint, 4)
    check(ccall(sym(:b200_dimmover), Cint, (Ptr{Cvoid}, Ptr{Cfloat}, Cint, Ptr{Cint}, Cint, Cint, Ptr{Cfloat}, Ptr{Cint}, Ptr{Cint}, Cint),
                ctx(), x, ndims(x), Cint[size(x)...], dm.from, dm.to, out, nd, od, 0))
    return reshape(out, Int.(od[1:nd[]])...), st
end

# ---- FFJORD (src/ffjord.jl:52-245): monte_carlo / regularize flags of `st` as in the reference -----------
struct FFJORD{M, T, A, K} <: LuxCore.AbstractLuxWrapperLayer{:model}
    model::M
    tspan::T
    input_dims
    args::A
    kwargs::K
end
FFJORD(model, tspan, input_dims, args...; ad = nothing, basedist = nothing, kwargs...) =
    (basedist === nothing || error("B200 path: basedist must be nothing"); FFJORD(model, tspan, input_dims, args, kwargs))
LuxCore.initialstates(rng, n::FFJORD) = (; model = LuxCore.initialstates(rng, n.model), regularize = false, monte_carlo = true)

function ffjord_forward(n::FFJORD, x::AbstractMatrix{Float32}, pv, st, tspan; want_tape::Bool, seed = rand(UInt64))
    B = size(x, 2)
    logpx = Vector{Float32}(undef, B)
    z = Matrix{Float32}(undef, size(x, 1) + (st.regularize ? 3 : 1), B)
    o = B200Opts(; abstol = get(n.kwargs, :abstol, 1.0f-6), reltol = get(n.kwargs, :reltol, 1.0f-3),
                 sensealg = want_tape ? 0 : -1, seed,             # InterpolatingAdjoint is hard-wired at src/ffjord.jl:143
                 ffjord_exact_trace = st.monte_carlo ? 0 : 1, ffjord_regularize = st.regularize ? 1 : 0)
    tape = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall(sym(:b200_ffjord_forward), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cfloat, Cfloat, Ref{B200Opts}, Ptr{Cfloat}, Ptr{Cfloat},
                 Ref{B200Stats}, Ptr{Ptr{Cvoid}}),
                ctx(), model_handle(n.model), pv, x, B, Float32(tspan[1]), Float32(tspan[2]), o, logpx, z, B200Stats(),
                want_tape ? tape : C_NULL))
    return logpx, z, (want_tape ? Tape(tape[]) : nothing), o
end
# (delta_logp, lambda_1, lambda_2) out of the end state, the selections of src/ffjord.jl:161-169 verbatim
function ffjord_outputs(z, regularize::Bool)
    L = size(z, 1)
    i1 = L - (regularize ? 2 : 0)
    Δlogp = z[i1:i1, :]
    regularize || return Δlogp, Δlogp, Δlogp
    return Δlogp, z[:, (L - 1):(L - 1)], z[:, L:L]
end
function (n::FFJORD)(x::AbstractMatrix{Float32}, ps, st)
    logpx, z, _, _ = ffjord_forward(n, x, collect(ps), st, n.tspan; want_tape = false)
    _, λ1, λ2 = ffjord_outputs(z, st.regularize)
    return (reshape(logpx, 1, :), λ1, λ2), st
end
function ChainRulesCore.rrule(n::FFJORD, x::AbstractMatrix{Float32}, ps, st)
    pv = collect(ps)
    logpx, z, tape, o = ffjord_forward(n, x, pv, st, n.tspan; want_tape = true)
    _, λ1, λ2 = ffjord_outputs(z, st.regularize)
    function ffjord_pullback(Δ)
        Δout = first(Δ)
        Δ1 = Δout[1]
        g = Δ1 isa AbstractArray ? Float32.(vec(Δ1)) : zeros(Float32, size(x, 2))         # cotangent of logpx
        gz = zeros(Float32, size(z))                                                        # cotangent of the end state
        if st.regularize
            L = size(z, 1)
            Δout[2] isa AbstractArray && (gz[:, L - 1] .+= vec(Δout[2]))
            Δout[3] isa AbstractArray && (gz[:, L] .+= vec(Δout[3]))
        end
        dx = similar(x); dp = similar(pv)
        check(ccall(sym(:b200_ffjord_adjoint_state), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Ref{B200Opts}, Ptr{Cfloat}, Ptr{Cfloat},
                     Ref{B200Stats}),
                    ctx(), model_handle(n.model), tape.h, g, gz, o, dx, dp, B200Stats()))
        return NoTangent(), dx, ComponentArray(dp, getaxes(ps)), NoTangent()
    end
    return ((reshape(logpx, 1, :), λ1, λ2), st), ffjord_pullback
end
# `__backward_ffjord` (src/ffjord.jl:183-214): samples = reverse-time flow of z ~ N(0, I)
function backward_ffjord(n::FFJORD, n_samples::Int, ps, st, rng = Random.default_rng())
    zs = randn(rng, Float32, n.input_dims..., n_samples)
    _, z, _, _ = ffjord_forward(n, reshape(zs, :, n_samples), collect(ps), st, reverse(n.tspan); want_tape = false)
    return z[1:(end - (st.regularize ? 3 : 1)), :]
end
struct FFJORDDistribution{F <: FFJORD, P, S}
    model::F
    ps::P
    st::S
end
logpdf(d::FFJORDDistribution, x::AbstractVecOrMat) = vec(first(first(d.model(reshape(x, size(x, 1), :), d.ps, d.st))))
Base.rand(rng::AbstractRNG, d::FFJORDDistribution, n::Int) = backward_ffjord(d.model, n, d.ps, d.st, rng)

# ---- multiple_shoot (src/multiple_shooting.jl:36-94, 231-243) -----------------------------------------
function group_ranges(datasize::Integer, groupsize::Integer)
    2 <= groupsize <= datasize || throw(DomainError(groupsize, "datasize must be positive and groupsize must to be within [2, datasize]"))
    G = ccall(sym(:b200_group_ranges), Cint, (Cint, Cint, Ptr{Cint}, Ptr{Cint}, Cint), datasize, groupsize, C_NULL, C_NULL, 0)
    f = Vector{Cint}(undef, G); l = Vector{Cint}(undef, G)
    ccall(sym(:b200_group_ranges), Cint, (Cint, Cint, Ptr{Cint}, Ptr{Cint}, Cint), datasize, groupsize, f, l, G)
    return [(f[i] + 1):(l[i] + 1) for i in 1:G]
end

# The reference takes `prob::ODEProblem` whose RHS closes over the network (docs/src/examples/multiple_shooting.md).  A C
# ABI cannot look inside an arbitrary closure, so the RHS must be this functor: `ODEProblem(B200RHS(nn, st), u0, tspan, p)`.
struct B200RHS{M, S}
    model::M
    st::S
end
(f::B200RHS)(u, p, t) = first(f.model(u, p, f.st))

_default_continuity_loss(û_end, u_0) = sum(abs, û_end - u_0)          # src/multiple_shooting.jl:243

# the grouped solves + their pullback: preds (D, B, rows) as a differentiable function of (p, ode_data)
function mshoot_solve(p, ode_data::Array{Float32, 3}, tsteps, model, ranges, kw; want_tape::Bool)
    D, B, T = size(ode_data)
    G = length(ranges)
    first0 = Cint[first(r) - 1 for r in ranges]; last0 = Cint[last(r) - 1 for r in ranges]
    preds = Array{Float32}(undef, D, B, sum(length, ranges))
    o = B200Opts(; abstol = get(kw, :abstol, 1.0f-6), reltol = get(kw, :reltol, 1.0f-3),
                 sensealg = want_tape ? sense_code(get(kw, :sensealg, nothing)) : -1)
    tapes = fill(C_NULL, G)
    rc = check(ccall(sym(:b200_mshoot_forward), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Cint, Ptr{Cint}, Ptr{Cint},
                      Ref{B200Opts}, Ptr{Cfloat}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
                     ctx(), model_handle(model), collect(p), ode_data, Float32.(tsteps), T, B, G, first0, last0, o, preds,
                     C_NULL, want_tape ? tapes : C_NULL))
    return preds, rc, [Tape(t) for t in tapes if t != C_NULL], o
end
function ChainRulesCore.rrule(::typeof(mshoot_solve), p, ode_data, tsteps, model, ranges, kw; want_tape::Bool)
    preds, rc, tapes, o = mshoot_solve(p, ode_data, tsteps, model, ranges, kw; want_tape = true)
    function mshoot_pullback(Δ)
        dpreds = Float32.(collect(first(Δ)))
        dp = Vector{Float32}(undef, length(p))
        ddata_g = Array{Float32}(undef, size(ode_data, 1), size(ode_data, 2), length(ranges))     # [G][B][D] blocks
        check(ccall(sym(:b200_mshoot_adjoint), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}, Cint, Ptr{Cfloat}, Ref{B200Opts}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cvoid}),
                    ctx(), model_handle(model), [t.h for t in tapes], length(tapes), dpreds, o, dp, ddata_g, C_NULL))
        ddata = zero(ode_data)                                     # block g belongs to time index first(ranges[g])
        for (g, r) in enumerate(ranges)
            ddata[:, :, first(r)] .+= ddata_g[:, :, g]
        end
        return NoTangent(), ComponentArray(dp, getaxes(p)), ddata, NoTangent(), NoTangent(), NoTangent(), NoTangent()
    end
    return (preds, rc, tapes, o), mshoot_pullback
end

function multiple_shoot(p, ode_data, tsteps, prob::SciMLBase.ODEProblem, loss_function::F, continuity_loss::C, solver,
                        group_size::Integer; continuity_term::Real = 100, kwargs...) where {F, C}
    rhs = prob.f.f
    rhs isa B200RHS || error("B200 path: build the problem as ODEProblem(B200RHS(model, st), u0, tspan, p)")
    nameof(typeof(solver)) == :Tsit5 || error("B200 path: only Tsit5() is supported")
    datasize = size(ode_data, ndims(ode_data))
    if group_size < 2 || group_size > datasize
        throw(DomainError(group_size, "group_size can't be < 2 or > number of data points"))
    end
    data3 = ndims(ode_data) == 2 ? reshape(Float32.(ode_data), size(ode_data, 1), 1, :) : Float32.(ode_data)
    ranges = group_ranges(datasize, group_size)
    preds, rc, _, _ = mshoot_solve(p, data3, tsteps, rhs.model, ranges, (; prob.kwargs..., kwargs...); want_tape = false)
    group_predictions, row = Array{Float32}[], 0
    for r in ranges
        g = preds[:, :, (row + 1):(row + length(r))]; row += length(r)
        push!(group_predictions, ndims(ode_data) == 2 ? g[:, 1, :] : g)
    end
    rc == 0 || return Inf, group_predictions                      # src/multiple_shooting.jl:66-67
    loss = zero(eltype(preds))
    N = ndims(ode_data)
    for (i, rg) in enumerate(ranges)
        u = selectdim(ode_data, N, rg)
        û = group_predictions[i]
        loss += loss_function(u, û)
        i > 1 && (loss += continuity_term * continuity_loss(selectdim(group_predictions[i - 1], N, size(group_predictions[i - 1], N)), selectdim(u, N, 1)))
    end
    return loss, group_predictions
end
multiple_shoot(p, ode_data, tsteps, prob::SciMLBase.ODEProblem, loss_function::F, solver, group_size::Integer; kwargs...) where {F} =
    multiple_shoot(p, ode_data, tsteps, prob, loss_function, _default_continuity_loss, solver, group_size; kwargs...)

# ---- one training iteration in one call (docs/src/examples/neural_ode.md:36-44,64-70) -------------------
Base.@kwdef mutable struct AdamState
    lr::Float32 = 0.05f0
    beta::NTuple{2, Float32} = (0.9f0, 0.999f0)
    eps::Float32 = 1.0f-8
    step::Int = 0
    m::Vector{Float32} = Float32[]
    v::Vector{Float32} = Float32[]
end
"`train_step!(node, u0, p, ode_data, opt)`: loss = sum(abs2, ode_data .- Array(node(u0, p, st)[1])), its gradient and the
Adam update of `p` in place; returns the loss before the update.  (With CuArrays and ptr_kind = 1 nothing leaves the device.)"
function train_step!(n::NeuralODE, u0::AbstractVecOrMat{Float32}, p::Vector{Float32}, ode_data::Array{Float32}, opt::AdamState)
    kw = n.kwargs
    t0, t1 = Float32.(n.tspan)
    interior, ss, se, everystep = split_saveat(get(kw, :saveat, nothing), t0, t1, kw)
    everystep && error("B200 path: train_step! needs saveat")
    isempty(opt.m) && (opt.m = zeros(Float32, length(p)); opt.v = zeros(Float32, length(p)))
    opt.step += 1
    o = B200Opts(; abstol = get(kw, :abstol, 1.0f-6), reltol = get(kw, :reltol, 1.0f-3), save_start = ss, save_end = se,
                 sensealg = sense_code(get(kw, :sensealg, nothing)))
    loss = Ref{Cfloat}(0)
    rc = check(ccall(sym(:b200_node_train_step), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cfloat, Cfloat, Cfloat, Cfloat, Ptr{Cfloat},
                      Cint, Cint, Cint, Cfloat, Cfloat, Ptr{Cfloat}, Cint, Ptr{Cfloat}, Ref{B200Opts}, Ptr{Cfloat}, Ref{B200Stats}, Ref{B200Stats}),
                     ctx(), model_handle(n.model), p, opt.m, opt.v, opt.step, opt.lr, opt.beta[1], opt.beta[2], opt.eps, u0,
                     size(u0, 1), 0, size(u0, 2), t0, t1, interior, length(interior), ode_data, o, loss, B200Stats(), B200Stats()))
    return rc == 0 ? loss[] : Inf32
end

end # module
