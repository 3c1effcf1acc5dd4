# run_reference.jl -- times the UNMODIFIED reference stack (DiffEqFlux + OrdinaryDiffEq + SciMLSensitivity + Lux +
# Zygote) on the host cores for BASELINE.json configs[1] (and configs[0] with `spiral`), on the same inputs as
# bench.py.  WRITE-ONLY in the build image: there is no Julia toolchain and no network there, so this script has
# never been executed; it exists so that the reference CPU number (and the reference's own step counts, final
# state and gradients -- the vectors that would PIN the oracle) can be produced wherever a Julia depot exists.
#
#   python tools/dump_inputs.py /tmp/b200_inputs          # writes p.f32, u0.f32, meta.txt (same generator as bench.py)
#   julia --project=/path/to/env -t auto baseline/run_reference.jl /tmp/b200_inputs [nsteps]
#
# Output: wall time per forward+gradient, sol.stats (nf / naccept / nreject) of the forward solve, and
# final_state.f32 / du0.f32 / dp.f32 / stats.txt next to the inputs (Float32, column-major (D, B) = the C ABI's [B][D]).
# `python tools/compare_reference_dump.py DIR` then compares them with tests/golden/headline.npz: one command pins the
# oracle (and with it the CUDA path's parity target) to the reference.
using DiffEqFlux, OrdinaryDiffEq, SciMLSensitivity, Lux, ComponentArrays, Zygote, LinearAlgebra, Random

dir = ARGS[1]
nsteps = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 3
meta = Dict(split(l, '=')[1] => split(l, '=')[2] for l in eachline(joinpath(dir, "meta.txt")))
dims = parse.(Int, split(meta["dims"], ','))
B = parse(Int, meta["batch"])
tol = parse(Float32, meta["tol"])
BLAS.set_num_threads(Sys.CPU_THREADS)

readf32(name, n) = (a = Vector{Float32}(undef, n); read!(joinpath(dir, name), a); a)
P = sum(dims[i] * dims[i + 1] + dims[i + 1] for i in 1:length(dims) - 1)
p_flat = readf32("p.f32", P)                       # [vec(W1) (out,in) column-major; b1; vec(W2); b2; ...]
u0 = reshape(readf32("u0.f32", dims[1] * B), dims[1], B)

layers = [Dense(dims[i] => dims[i + 1], i < length(dims) - 1 ? tanh : identity) for i in 1:length(dims) - 1]
model = Chain(layers...)
ps, st = Lux.setup(Xoshiro(0), model)
ps = ComponentArray(ps)
@assert length(ps) == P
ps .= p_flat                                       # ComponentArray flat order == the C ABI's parameter order

node = NeuralODE(model, (0.0f0, 1.0f0), Tsit5(); save_everystep = false, save_start = false,
                 abstol = tol, reltol = tol, sensealg = BacksolveAdjoint(; autojacvec = ZygoteVJP()))
loss(x, p) = sum(last(first(node(x, p, st)).u))

sol = first(node(u0, ps, st))                      # warm-up (compilation) + stats
println("forward stats: nf=", sol.stats.nf, " naccept=", sol.stats.naccept, " nreject=", sol.stats.nreject)
g = Zygote.gradient(loss, u0, ps)                  # warm-up of the pullback
times = Float64[]
for _ in 1:nsteps
    t = @elapsed begin
        Zygote.gradient(loss, u0, ps)
    end
    push!(times, t)
end
nf_fwd = sol.stats.nf
println("wall per forward+gradient [s]: ", times, "  (median ", sort(times)[cld(nsteps, 2)], ")")
println("threads: julia ", Threads.nthreads(), ", BLAS ", BLAS.get_num_threads())
println("RHS evals/s (forward count only is exact; adjoint nf is not exposed by the pullback): >= ",
        B * nf_fwd / sort(times)[cld(nsteps, 2)])
# step statistics of the adjoint solve: SciMLSensitivity does not return the backward solution through the pullback, so the
# same adjoint problem is solved once more explicitly (adjoint_sensitivities on the dense-free Backsolve formulation)
adj_stats = try
    _, _, adjsol = SciMLSensitivity.adjoint_sensitivities(
        solve(ODEProblem{false}((u, p, t) -> first(model(u, p, st)), u0, (0.0f0, 1.0f0), ps), Tsit5(); save_everystep = false,
              save_start = false, abstol = tol, reltol = tol), Tsit5();
        t = [1.0f0], dgdu_discrete = (out, u, p, t, i) -> (out .= 1), sensealg = BacksolveAdjoint(; autojacvec = ZygoteVJP()),
        abstol = tol, reltol = tol, return_adjoint_solution = true)      # keyword available in recent SciMLSensitivity
    adjsol.stats
catch err
    @warn "adjoint step statistics not available" err
    nothing
end
open(joinpath(dir, "stats.txt"), "w") do io
    println(io, "fwd_naccept=", sol.stats.naccept); println(io, "fwd_nreject=", sol.stats.nreject); println(io, "fwd_nf=", sol.stats.nf)
    if adj_stats !== nothing
        println(io, "adj_naccept=", adj_stats.naccept); println(io, "adj_nreject=", adj_stats.nreject); println(io, "adj_nf=", adj_stats.nf)
    end
end
write(joinpath(dir, "final_state.f32"), Float32.(vec(last(sol.u))))
write(joinpath(dir, "du0.f32"), Float32.(vec(g[1])))
write(joinpath(dir, "dp.f32"), Float32.(collect(g[2])))
