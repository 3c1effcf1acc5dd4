# B200DiffEqFlux.jl -- the Julia side of the drop-in boundary (include/b200node.h).
#
# WRITE-ONLY in this repository: the build image has no Julia toolchain, so this file is the binding a
# DiffEqFlux maintainer would add; the Python ctypes layer (diffeqflux.jl_b200/layers.py) exercises
# the same C ABI with the same argument conventions and is what the tests run.
#
# It re-declares the hot-path layer types of DiffEqFlux (src/neural_de.jl:35-61, src/ffjord.jl:52-178,
# src/neural_de.jl:372-386, src/multiple_shooting.jl:36-94) so that
#     node = NeuralODE(model, tspan, Tsit5(); saveat, sensealg);  sol, st = node(u0, ps, st)
#     Zygote.gradient(sum ∘ first ∘ node, u0, ps, st)
# keep working, with ONE ccall per solve and ONE ccall per pullback.  Unsupported configurations
# (non-Tsit5 solver, non-Dense model, Float64, other sensealgs) throw: there is no CPU fallback.
module B200DiffEqFlux

using ChainRulesCore, Lux, LuxCore, ComponentArrays, SciMLBase, Libdl

export NeuralODE, AugmentedNDELayer, FFJORD, multiple_shoot, group_ranges

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
    sensealg::Cint = 0          # 0 InterpolatingAdjoint, 1 BacksolveAdjoint, -1 none
    ptr_kind::Cint = 0          # 0 host (Array), 1 device (CuArray pointer)
    tile::Cint = 0
    max_dense_steps::Cint = 0
    seed::UInt64 = 0
    noise_per_call::Cint = 1
    b_offset::Cint = 0
    save_everystep::Cint = 0
    path::Cint = 0              # 0 auto, 1 FFMA only, 2 require tcgen05
    reserved::NTuple{2, Cint} = (0, 0)
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

const ACT = IdDict{Any, Cint}(identity => 0, tanh => 1, Lux.relu => 2, Lux.sigmoid => 3, Lux.softplus => 4)

# Chain(Dense...) -> b200_model (anything else is outside the B200 path)
function model_handle(model::Lux.Chain)
    layers = collect(values(model.layers))
    all(l -> l isa Lux.Dense, layers) || error("B200 path: model must be a Chain of Dense layers")
    dims = Cint[layers[1].in_dims; [l.out_dims for l in layers]]
    acts = Cint[get(() -> error("B200 path: unsupported activation $(l.activation)"), ACT, l.activation) for l in layers]
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall(sym(:b200_model_create), Cint, (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cint}, Cint, Ptr{Ptr{Cvoid}}),
                ctx(), length(layers), dims, acts, 0, out))
    return out[]
end

# ---- NeuralODE (src/neural_de.jl:35-61) -------------------------------------------------------------
struct NeuralODE{M <: LuxCore.AbstractLuxLayer, T, A, K} <: LuxCore.AbstractLuxWrapperLayer{:model}
    model::M
    tspan::T
    args::A
    kwargs::K
end
NeuralODE(model, tspan, args...; kwargs...) = NeuralODE(model, tspan, args, kwargs)

struct B200Solution{T, U}          # what callers touch of an ODESolution: .u, .t, .retcode, Array(sol)
    t::Vector{T}
    u::Vector{U}
    retcode::SciMLBase.ReturnCode.T
    stats::B200Stats
end
Base.Array(s::B200Solution) = cat(s.u...; dims = ndims(s.u[1]) + 1)
Base.last(s::B200Solution) = last(s.u)
Base.getindex(s::B200Solution, i) = s.u[i]

const RETCODES = (SciMLBase.ReturnCode.Success, SciMLBase.ReturnCode.MaxIters,
                  SciMLBase.ReturnCode.DtLessThanMin, SciMLBase.ReturnCode.Unstable)

function split_saveat(saveat, t0::Float32, t1::Float32, kw)
    sv = saveat === nothing ? Float32[] : saveat isa Number ? collect(t0:Float32(abs(saveat)):t1) : Float32.(collect(saveat))
    everystep = get(kw, :save_everystep, isempty(sv))
    ss = get(kw, :save_start, everystep || isempty(sv) || saveat isa Number || t0 in sv)
    se = get(kw, :save_end, everystep || isempty(sv) || saveat isa Number || t1 in sv)
    return sort(filter(t -> t0 < t < t1, sv)), ss, se, everystep && isempty(sv)
end

function forward(n::NeuralODE, x::AbstractVecOrMat{Float32}, p::AbstractVector{Float32}; want_tape::Bool, augment::Int = 0)
    length(n.args) == 1 && nameof(typeof(n.args[1])) == :Tsit5 || error("B200 path: only Tsit5() is supported")
    kw = n.kwargs
    t0, t1 = Float32.(n.tspan)
    interior, ss, se, everystep = split_saveat(get(kw, :saveat, nothing), t0, t1, kw)
    sa = get(kw, :sensealg, nothing)
    sense = sa === nothing ? 0 : nameof(typeof(sa)) == :InterpolatingAdjoint ? 0 :
            nameof(typeof(sa)) == :BacksolveAdjoint ? 1 : error("B200 path: unsupported sensealg $(typeof(sa))")
    o = B200Opts(; abstol = get(kw, :abstol, 1.0f-6), reltol = get(kw, :reltol, 1.0f-3), save_start = ss, save_end = se,
                 sensealg = want_tape ? sense : -1, save_everystep = everystep, maxiters = get(kw, :maxiters, 100000))
    d_in, B = size(x, 1), size(x, 2)
    D = d_in + augment
    n_out = ss + length(interior) + se
    out_u = Array{Float32}(undef, D, B, n_out)
    out_t = Vector{Float32}(undef, n_out)
    st = B200Stats()
    tape = Ref{Ptr{Cvoid}}(C_NULL)
    rc = check(ccall(sym(:b200_node_forward), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Cint, Cfloat, Cfloat, Ptr{Cfloat}, Cint,
                      Ref{B200Opts}, Ptr{Cfloat}, Ptr{Cfloat}, Ref{B200Stats}, Ptr{Ptr{Cvoid}}),
                     ctx(), model_handle(n.model), p, x, d_in, augment, B, t0, t1, interior, length(interior),
                     o, out_u, out_t, st, want_tape ? tape : C_NULL))
    u = [x isa AbstractVector ? out_u[:, 1, i] : out_u[:, :, i] for i in 1:n_out]
    return B200Solution(out_t, u, RETCODES[rc + 1], st), tape[], o
end

# `(n::NeuralODE)(x, p, st)` (src/neural_de.jl:47-61)
(n::NeuralODE)(x, p, st) = (forward(n, x, collect(p); want_tape = false)[1], st)

# the rrule Zygote finds instead of SciMLSensitivity's (src/DiffEqFlux.jl:17-22)
function ChainRulesCore.rrule(n::NeuralODE, x, p, st)
    sol, tape, o = forward(n, x, collect(p); want_tape = true)
    function node_pullback(Δ)
        dsol = first(Δ)
        dL = Array{Float32}(undef, size(sol.u[1])..., length(sol.u))
        for i in eachindex(sol.u)
            selectdim(dL, ndims(dL), i) .= dsol.u[i]
        end
        du0 = similar(x)
        dp = similar(collect(p))
        st2 = B200Stats()
        check(ccall(sym(:b200_node_adjoint), Cint,
                    (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Ref{B200Opts}, Ptr{Cfloat}, Ptr{Cfloat}, Ref{B200Stats}),
                    ctx(), model_handle(n.model), tape, dL, o, du0, dp, st2))
        ccall(sym(:b200_tape_free), Cvoid, (Ptr{Cvoid},), tape)
        return NoTangent(), du0, ComponentArray(dp, getaxes(p)), NoTangent()
    end
    return (sol, st), node_pullback
end

# ---- AugmentedNDELayer (src/neural_de.jl:372-386): the zero padding is fused into the kernel ----------
struct AugmentedNDELayer{N <: NeuralODE} <: LuxCore.AbstractLuxWrapperLayer{:nde}
    nde::N
    adim::Int
end
(a::AugmentedNDELayer)(x, p, st) = (forward(a.nde, x, collect(p); want_tape = false, augment = a.adim)[1], st)

# ---- FFJORD (src/ffjord.jl:52-178), monte_carlo = true / regularize = false branch -------------------
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

function (n::FFJORD)(x::AbstractMatrix{Float32}, ps, st)
    (st.regularize || !st.monte_carlo) && error("B200 path: only regularize=false, monte_carlo=true")
    B = size(x, 2)
    logpx = Vector{Float32}(undef, B)
    z = Matrix{Float32}(undef, size(x, 1) + 1, B)
    o = B200Opts(; abstol = get(n.kwargs, :abstol, 1.0f-6), reltol = get(n.kwargs, :reltol, 1.0f-3),
                 sensealg = -1, seed = rand(UInt64))
    s = B200Stats()
    check(ccall(sym(:b200_ffjord_forward), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cfloat, Cfloat, Ref{B200Opts}, Ptr{Cfloat}, Ptr{Cfloat},
                 Ref{B200Stats}, Ptr{Ptr{Cvoid}}),
                ctx(), model_handle(n.model), collect(ps), x, B, Float32(n.tspan[1]), Float32(n.tspan[2]), o, logpx, z, s, C_NULL))
    Δlogp = z[end:end, :]
    return (reshape(logpx, 1, B), Δlogp, Δlogp), st
end
# (the rrule of FFJORD calls b200_ffjord_adjoint exactly as NeuralODE's calls b200_node_adjoint)

# ---- multiple_shoot (src/multiple_shooting.jl:36-94, 231-239) -----------------------------------------
function group_ranges(datasize::Integer, groupsize::Integer)
    2 <= groupsize <= datasize || throw(DomainError(groupsize, "datasize must be positive and groupsize must to be within [2, datasize]"))
    G = ccall(sym(:b200_group_ranges), Cint, (Cint, Cint, Ptr{Cint}, Ptr{Cint}, Cint), datasize, groupsize, C_NULL, C_NULL, 0)
    f = Vector{Cint}(undef, G); l = Vector{Cint}(undef, G)
    ccall(sym(:b200_group_ranges), Cint, (Cint, Cint, Ptr{Cint}, Ptr{Cint}, Cint), datasize, groupsize, f, l, G)
    return [(f[i] + 1):(l[i] + 1) for i in 1:G]
end

function multiple_shoot(p, ode_data::Array{Float32, 3}, tsteps, model::Lux.Chain, loss_function, continuity_loss, ::Any,
                        group_size::Integer; continuity_term::Real = 100, abstol = 1.0f-6, reltol = 1.0f-3)
    D, B, T = size(ode_data)
    ranges = group_ranges(T, group_size)
    G = length(ranges)
    first0 = Cint[first(r) - 1 for r in ranges]; last0 = Cint[last(r) - 1 for r in ranges]
    rows = sum(length, ranges)
    preds = Array{Float32}(undef, D, B, rows)
    stats = [B200Stats() for _ in 1:G]
    o = B200Opts(; abstol, reltol, sensealg = -1)
    rc = check(ccall(sym(:b200_mshoot_forward), Cint,
                     (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cfloat}, Ptr{Cfloat}, Ptr{Cfloat}, Cint, Cint, Cint, Ptr{Cint}, Ptr{Cint},
                      Ref{B200Opts}, Ptr{Cfloat}, Ptr{Cvoid}, Ptr{Ptr{Cvoid}}),
                     ctx(), model_handle(model), collect(p), ode_data, Float32.(tsteps), T, B, G, first0, last0, o, preds,
                     C_NULL, C_NULL))
    group_predictions, row = Array{Float32, 3}[], 0
    for r in ranges
        push!(group_predictions, preds[:, :, (row + 1):(row + length(r))]); row += length(r)
    end
    rc == 0 || return Inf, group_predictions                      # src/multiple_shooting.jl:66-67
    loss = 0.0f0
    for (i, rg) in enumerate(ranges)
        u = ode_data[:, :, rg]
        loss += loss_function(u, group_predictions[i])
        i > 1 && (loss += continuity_term * continuity_loss(group_predictions[i - 1][:, :, end], u[:, :, 1]))
    end
    return loss, group_predictions
end

end # module
