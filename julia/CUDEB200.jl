# CUDEB200.jl — Julia host side of the B200 cUDE hot path (thin `ccall` shim over libcude_b200.so).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image, so this file is
# reviewed against include/cude_b200.h only.  The tested host mirror is the Python package
# conditional_ude_b200 (same names, same argument meaning).
#
# Drop-in usage inside the reference repository (after `include("src/parameter-estimation.jl")`):
#
#     include("CUDEB200.jl"); using .CUDEB200
#     include("cude_overrides.jl")      # re-defines the reference's own `loss` / `likelihood_profile` methods (same signatures)
#                                       # on top of this module: `train`, `train_with_sigma`, `evaluate_model` and the scripts
#                                       # then run UNMODIFIED, AutoForwardDiff() included (see cude_overrides.jl)
#
# or explicitly:
#     pop  = CUDEB200.Population(models, timepoints, cpeptide_data)   # from the reference's own model vector
#     optf = CUDEB200.optimization_function(pop)          # replaces OptimizationFunction(loss, AutoForwardDiff())
#     prob = OptimizationProblem(optf, θ0, nothing)       # θ0::ComponentArray(neural=…, conditional=…)
#     mctx = CUDEB200.Context([0, 1, 2, 3, 4, 5, 6, 7])   # all GPUs of the box from this one Julia session
#     mpop = CUDEB200.MultiPopulation(models, timepoints, cpeptide_data; ctx=mctx, shard=:individuals)
module CUDEB200

using ComponentArrays: ComponentArray
using SciMLBase: OptimizationFunction

const libcude = get(ENV, "CUDE_B200_LIB", "libcude_b200.so")

struct CudeNet
    n_in::Cint
    depth::Cint
    width::Cint
end

struct CudeOpts
    abstol::Cdouble
    reltol::Cdouble
    maxiters::Cint
    precision::Cint
    block::Cint
    balance::Cint     # 0: automatic (<= 8192 trajectories: warp per trajectory; >= 32768 individuals: two-kernel gradient; else fused kernel); 1: history regrouping; 2 / 3 / 4: force two-kernel / fused / warp-per-trajectory (include/cude_b200.h)
    split::Cint       # 2: split gradient pipeline instead of the fused adjoint kernel (include/cude_b200.h)
end
CudeOpts(; abstol=1e-6, reltol=1e-3, maxiters=1_000_000, precision=0, balance=0, split=0) =
    CudeOpts(abstol, reltol, maxiters, precision, 0, balance, split)

check(rc::Cint, ctx=C_NULL) = rc == 0 ? nothing :
    error("cude_b200 error $rc: " * unsafe_string(ccall((:cude_last_error, libcude), Cstring, (Ptr{Cvoid},), ctx)))

mutable struct Context
    handle::Ptr{Cvoid}
    function Context(device::Integer=0)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:cude_ctx_create, libcude), Cint, (Cint, Ref{Ptr{Cvoid}}), device, h))
        ctx = new(h[])
        finalizer(c -> ccall((:cude_ctx_destroy, libcude), Cint, (Ptr{Cvoid},), c.handle), ctx)
        ctx
    end
end

const default_context = Ref{Union{Nothing,Context}}(nothing)
context() = (default_context[] === nothing && (default_context[] = Context(0)); default_context[])

"""
    MultiContext(devices)   /   Context(devices::AbstractVector)

All (or some) GPUs of the box driven from this one Julia session (`cude_mctx_create`): one context, stream and host
worker thread per device inside the library; NCCL (loaded by the library at run time) all-reduces the per-start sums of
individual-sharded populations.  `MultiContext()` takes every visible device.
"""
mutable struct MultiContext
    handle::Ptr{Cvoid}
    n_gpus::Int
    function MultiContext(devices::AbstractVector{<:Integer}=Int[])
        h = Ref{Ptr{Cvoid}}(C_NULL)
        ids = Cint.(devices)
        rc = ccall((:cude_mctx_create, libcude), Cint, (Cint, Ptr{Cint}, Ref{Ptr{Cvoid}}), length(ids), isempty(ids) ? C_NULL : ids, h)
        rc == 0 || error("cude_b200 error $rc: " * unsafe_string(ccall((:cude_mlast_error, libcude), Cstring, (Ptr{Cvoid},), C_NULL)))
        m = new(h[], Int(ccall((:cude_mctx_size, libcude), Cint, (Ptr{Cvoid},), h[])))
        finalizer(c -> ccall((:cude_mctx_destroy, libcude), Cint, (Ptr{Cvoid},), c.handle), m)
        m
    end
end
Context(devices::AbstractVector{<:Integer}) = MultiContext(devices)
mcheck(rc::Cint, m::MultiContext) = rc == 0 ? nothing :
    error("cude_b200 error $rc: " * unsafe_string(ccall((:cude_mlast_error, libcude), Cstring, (Ptr{Cvoid},), m.handle)))

"""
Constants of one reference model object, read out of the closures the reference's own constructor built
(src/c-peptide-models.jl:170-194 / :196-220): `model.problem.f.f` is `combined!` (`combine`, :108-114) capturing
`kinetics!` — the `ode!` closure of `van_cauter_model` (:56-64) with `k0, k1, k2, c_peptide_0` — and `production`, which
captures the `LinearInterpolation` `glucose` (`.t` knots, `.u` values) and, for the covariate model, `age`.
"""
function model_constants(model)
    comb = model.problem.f.f
    kin = getfield(comb, Symbol("kinetics!"))
    prod = getfield(comb, :production)
    g = getfield(prod, :glucose)
    age = :age in fieldnames(typeof(prod)) ? Float64(getfield(prod, :age)) : nothing
    (k0 = Float64(kin.k0), k1 = Float64(kin.k1), k2 = Float64(kin.k2), c0 = Float64(kin.c_peptide_0),
     knot_t = collect(Float64, g.t), knot_g = collect(Float64, g.u), covariate = age)
end

"""
Device-resident image of a vector of `CPeptideConditionalUDEModel` (src/c-peptide-models.jl:170-194) and
its data.  The constants are read back out of the reference's own model objects: `problem.u0`,
`problem.tspan`; the kinetic parameters are recomputed with `van_cauter_parameters(age, t2dm)`, so the
constructor additionally needs `ages` and `t2dm` (the reference's ODEProblem closure does not expose them).
"""
mutable struct Population
    handle::Ptr{Cvoid}
    ctx::Context
    n::Int
    net::CudeNet
    nparams::Int
end

function Population(glucose::AbstractMatrix, glucose_timepoints::AbstractVector, ages::AbstractVector,
                    t2dm::AbstractVector{Bool}, timepoints::AbstractVector, cpeptide::AbstractMatrix;
                    net::CudeNet=CudeNet(2, 2, 4), covariate=nothing, ctx::Context=context())
    n, K, M = size(glucose, 1), length(glucose_timepoints), length(timepoints)
    # row-major [n x K] for the C side == column-major [K x n] here
    knot_t = repeat(Float64.(glucose_timepoints), 1, n)
    knot_g = permutedims(Float64.(glucose))
    obs_t = repeat(Float64.(timepoints), 1, n)
    obs_y = permutedims(Float64.(cpeptide))
    kin = zeros(4, n)
    for i in 1:n
        k0 = Ref(0.0); k1 = Ref(0.0); k2 = Ref(0.0)
        ccall((:cude_van_cauter_parameters, libcude), Cvoid, (Cdouble, Cint, Ref{Cdouble}, Ref{Cdouble}, Ref{Cdouble}),
              ages[i], t2dm[i], k0, k1, k2)
        kin[:, i] .= (k0[], k1[], k2[], cpeptide[i, 1])        # c0 = cpeptide_data[1], c-peptide-models.jl:174
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    cov = covariate === nothing ? C_NULL : Float64.(covariate)      # ccall roots the array for the call
    check(ccall((:cude_population_create, libcude), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble},
                 Ptr{Cdouble}, Ptr{Cdouble}, Ref{Ptr{Cvoid}}),
                ctx.handle, n, K, fill(Cint(K), n), knot_t, knot_g, M, fill(Cint(M), n), obs_t, obs_y, kin, cov, h), ctx.handle)
    P = Int(ccall((:cude_net_nparams, libcude), Cint, (Ref{CudeNet},), net))
    pop = Population(h[], ctx, n, net, P)
    finalizer(x -> ccall((:cude_population_destroy, libcude), Cint, (Ptr{Cvoid},), x.handle), pop)
    pop
end

"""
    Population(models, timepoints, cpeptide_data; ctx)

From the reference's own vector of `CPeptideConditionalUDEModel` (or covariate models) and the `(timepoints, cpeptide_data)`
of the loss tuple (:126): nothing else is needed, the model constants come from `model_constants`.
"""
function _pack(models::AbstractVector, timepoints::AbstractVector, cpeptide::AbstractVecOrMat)
    cs = [model_constants(m) for m in models]
    n = length(models)
    K = maximum(length(c.knot_t) for c in cs); M = length(timepoints)
    knot_t = zeros(K, n); knot_g = zeros(K, n); nk = Vector{Cint}(undef, n)
    for (i, c) in enumerate(cs)
        nk[i] = length(c.knot_t)
        knot_t[1:nk[i], i] .= c.knot_t; knot_g[1:nk[i], i] .= c.knot_g
    end
    Y = cpeptide isa AbstractVector ? reshape(Float64.(cpeptide), 1, :) : Float64.(cpeptide)       # [n x M]
    kin = [getfield(c, f) for f in (:k0, :k1, :k2, :c0), c in cs]                                   # [4 x n]
    cov = cs[1].covariate === nothing ? nothing : Float64[c.covariate for c in cs]
    net = CudeNet(cov === nothing ? 2 : 3, 2, 4)                                                    # chain(4, 2, tanh; input_dims)
    (; n, K, M, nk, knot_t, knot_g, obs_t = repeat(Float64.(timepoints), 1, n), obs_y = permutedims(Y), kin, cov, net)
end

function Population(models::AbstractVector, timepoints::AbstractVector, cpeptide::AbstractVecOrMat; ctx::Context=context())
    p = _pack(models, timepoints, cpeptide)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:cude_population_create, libcude), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble},
                 Ptr{Cdouble}, Ptr{Cdouble}, Ref{Ptr{Cvoid}}),
                ctx.handle, p.n, p.K, p.nk, p.knot_t, p.knot_g, p.M, fill(Cint(p.M), p.n), p.obs_t, p.obs_y, p.kin,
                p.cov === nothing ? C_NULL : p.cov, h), ctx.handle)
    pop = Population(h[], ctx, p.n, p.net, Int(ccall((:cude_net_nparams, libcude), Cint, (Ref{CudeNet},), p.net)))
    finalizer(x -> ccall((:cude_population_destroy, libcude), Cint, (Ptr{Cvoid},), x.handle), pop)
    pop
end

"""
Population on a `MultiContext`.  `shard = :starts`: every device holds the whole population and a call's starts are split
over the devices (screening :362-366, selected starts :374-376, beta-only fits :272-288, profiles): no communication.
`shard = :individuals`: the individuals (loop of :126-140) are split and every call ends with the library's NCCL
all-reduce of the per-start sums.  Same `loss` / `loss_grad` methods and results as `Population`.
"""
mutable struct MultiPopulation
    handle::Ptr{Cvoid}
    ctx::MultiContext
    n::Int
    net::CudeNet
    nparams::Int
end
function MultiPopulation(models::AbstractVector, timepoints::AbstractVector, cpeptide::AbstractVecOrMat;
                         ctx::MultiContext=MultiContext(), shard::Symbol=:starts)
    p = _pack(models, timepoints, cpeptide)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    mode = shard === :individuals ? 1 : 0          # CUDE_SHARD_INDIVIDUALS / CUDE_SHARD_STARTS
    mcheck(ccall((:cude_mpopulation_create, libcude), Cint,
                 (Ptr{Cvoid}, Cint, Cint, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble},
                  Ptr{Cdouble}, Ptr{Cdouble}, Ref{Ptr{Cvoid}}),
                 ctx.handle, mode, p.n, p.K, p.nk, p.knot_t, p.knot_g, p.M, fill(Cint(p.M), p.n), p.obs_t, p.obs_y, p.kin,
                 p.cov === nothing ? C_NULL : p.cov, h), ctx)
    mp = MultiPopulation(h[], ctx, p.n, p.net, Int(ccall((:cude_net_nparams, libcude), Cint, (Ref{CudeNet},), p.net)))
    finalizer(x -> ccall((:cude_mpopulation_destroy, libcude), Cint, (Ptr{Cvoid},), x.handle), mp)
    mp
end

function loss(pop::MultiPopulation, neural::AbstractVecOrMat{Float64}, cond::AbstractMatrix{Float64}; opts=CudeOpts())
    S = size(cond, 2)
    out = Vector{Float64}(undef, S)
    stride = ndims(neural) == 1 ? 0 : size(neural, 1)
    mcheck(ccall((:cude_mloss, libcude), Cint,
                 (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CudeNet}, Ref{CudeOpts}, Cint, Ptr{Cdouble}, Clonglong, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                 pop.ctx.handle, pop.handle, pop.net, opts, S, neural, stride, cond, C_NULL, out), pop.ctx)
    out
end

function loss_grad(pop::MultiPopulation, neural::AbstractVecOrMat{Float64}, cond::AbstractMatrix{Float64}; opts=CudeOpts(),
                   mean::Bool=true, neural_grad::Bool=true)
    S = size(cond, 2)
    l = Vector{Float64}(undef, S)
    gn = neural_grad ? Matrix{Float64}(undef, pop.nparams, S) : nothing
    gc = Matrix{Float64}(undef, pop.n, S)
    stride = ndims(neural) == 1 ? 0 : size(neural, 1)
    mcheck(ccall((:cude_mloss_grad, libcude), Cint,
                 (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CudeNet}, Ref{CudeOpts}, Cint, Ptr{Cdouble}, Clonglong, Ptr{Cdouble}, Cint,
                  Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                 pop.ctx.handle, pop.handle, pop.net, opts, S, neural, stride, cond, mean ? 1 : 0,
                 C_NULL, l, gn === nothing ? C_NULL : gn, gc), pop.ctx)
    l, gn, gc
end

# one device population per (model vector, data) the reference passes around: built on first use, then re-used by every
# `loss(θ, (models, t, Y))` call of an optimisation (the reference re-creates nothing between calls either)
const _pop_cache = IdDict{Any,Any}()
function cached_population(models, timepoints, cpeptide)
    get!(_pop_cache, models) do
        (Population(models isa AbstractVector ? models : [models], timepoints, cpeptide), copy(cpeptide))
    end[1]
end

"""
    loss(pop, neural, cond; opts) -> Vector{Float64}

Batched loss: `neural` is `P` (shared network: fixed-NN / profile case) or `P × S`; `cond` is `N × S`
(column s = start s).  Returns the S population losses (mean over individuals, `Inf` on solver failure) —
the batched form of src/parameter-estimation.jl:126-140, :362-366 and src/likelihood-profiles.jl:11-14.
"""
function loss(pop::Population, neural::AbstractVecOrMat{Float64}, cond::AbstractMatrix{Float64}; opts=CudeOpts(),
              sse::Union{Nothing,Matrix{Float64}}=nothing)
    S = size(cond, 2)
    out = Vector{Float64}(undef, S)
    stride = ndims(neural) == 1 ? 0 : size(neural, 1)
    check(ccall((:cude_loss, libcude), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CudeNet}, Ref{CudeOpts}, Cint, Ptr{Cdouble}, Clonglong, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                pop.ctx.handle, pop.handle, pop.net, opts, S, neural, stride, cond, sse === nothing ? C_NULL : sse, out),
          pop.ctx.handle)
    out
end

"""
    loss_grad(pop, neural, cond; opts, mean=true) -> (loss[S], g_neural[P×S], g_cond[N×S])

Value and gradient — what `OptimizationFunction(loss, AutoForwardDiff())` (:231, :281, :299, :370) computes
with 8 chunked dual-number re-solves, here one adjoint pass per trajectory.
"""
function loss_grad(pop::Population, neural::AbstractVecOrMat{Float64}, cond::AbstractMatrix{Float64}; opts=CudeOpts(),
                   mean::Bool=true, neural_grad::Bool=true)
    S = size(cond, 2)
    l = Vector{Float64}(undef, S)
    gn = neural_grad ? Matrix{Float64}(undef, pop.nparams, S) : nothing
    gc = Matrix{Float64}(undef, pop.n, S)
    stride = ndims(neural) == 1 ? 0 : size(neural, 1)
    check(ccall((:cude_loss_grad, libcude), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CudeNet}, Ref{CudeOpts}, Cint, Ptr{Cdouble}, Clonglong, Ptr{Cdouble}, Cint,
                 Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                pop.ctx.handle, pop.handle, pop.net, opts, S, neural, stride, cond, mean ? 1 : 0,
                C_NULL, l, gn === nothing ? C_NULL : gn, gc), pop.ctx.handle)
    l, gn, gc
end

"""
Sharded-population form (one rank's block of the individuals, `:126-140` split over ranks): unscaled per-start sums
`sums[P+1, S] = {Σ_i sse_i, Σ_i ∂sse_i/∂neural}` of this shard and `g_cond = cond_scale · ∂sse_i/∂cond` (pass
`1/N_global`); all-reduce `sums` over the ranks and divide by `N_global`.  A start whose `sums[1, s]` is not finite failed.
"""
function loss_grad_sums(pop::Population, neural::AbstractVecOrMat{Float64}, cond::AbstractMatrix{Float64}, cond_scale::Float64;
                        opts=CudeOpts())
    S = size(cond, 2)
    sums = Matrix{Float64}(undef, pop.nparams + 1, S)
    gc = Matrix{Float64}(undef, pop.n, S)
    stride = ndims(neural) == 1 ? 0 : size(neural, 1)
    check(ccall((:cude_loss_grad_sums, libcude), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CudeNet}, Ref{CudeOpts}, Cint, Ptr{Cdouble}, Clonglong, Ptr{Cdouble}, Cdouble,
                 Ptr{Cdouble}, Ptr{Cdouble}),
                pop.ctx.handle, pop.handle, pop.net, opts, S, neural, stride, cond, cond_scale, sums, gc), pop.ctx.handle)
    sums, gc
end

# --- the reference's own entry points, same tuple shapes (src/parameter-estimation.jl:56, :93, :126) ---------------
# `p` carries a Population in place of the model (vector): (pop, timepoints, cpeptide_data[, nn]).
loss(θ, (pop, _, _)::Tuple{Population,Any,Any}) =
    loss(pop, collect(Float64, θ.neural), reshape(collect(Float64, θ.conditional), :, 1))[1]
loss(β, (pop, _, _, nn)::Tuple{Population,Any,Any,AbstractVector}) =
    loss(pop, collect(Float64, nn), fill(Float64(first(β)), pop.n, 1))[1]
loss_sigma(θ, p::Tuple{Population,Any,Any,AbstractVector}) =
    (n = length(p[2]); (n / 2) * log(θ.sigma^2) + loss(θ.ode, p) / (2 * θ.sigma^2))

"""
`OptimizationFunction` with the analytic gradient from the GPU — drop-in for
`OptimizationFunction(loss, AutoForwardDiff())` in `train` (src/parameter-estimation.jl:370).
"""
function optimization_function(pop::Population; opts=CudeOpts())
    f(θ, _) = loss(pop, collect(Float64, θ.neural), reshape(collect(Float64, θ.conditional), :, 1); opts)[1]
    function g!(G, θ, _)
        _, gn, gc = loss_grad(pop, collect(Float64, θ.neural), reshape(collect(Float64, θ.conditional), :, 1); opts)
        G.neural .= vec(gn); G.conditional .= vec(gc)
        nothing
    end
    OptimizationFunction(f; grad=g!)
end

"""
    likelihood_profile(β, nn, pop_one, lower, upper, sigma; steps=1000)

src/likelihood-profiles.jl:4-17 with the whole grid evaluated as one launch (`pop_one` holds one individual).
"""
function likelihood_profile(β, nn, pop::Population, lower_bound, upper_bound, sigma; steps=1000, opts=CudeOpts())
    parameter_values = range(lower_bound, stop=upper_bound, length=steps)
    cond = reshape(vcat(Float64(first(β)), collect(parameter_values)), 1, :)
    sse = loss(pop, collect(Float64, nn), cond; opts)
    sse[2:end] ./ (2 * sigma^2), sse[1] / (2 * sigma^2), parameter_values
end

"""
    simulate(pop, neural, cond; opts) -> yhat[M × N × S]

`solve(model.problem, p=θ, saveat=timepoints, save_idxs=1)` (src/parameter-estimation.jl:59) for every (individual, start):
the model's plasma c-peptide at the observation times (`NaN` beyond an individual's last observation and for failed solves) —
what the scripts plot against the data (c-peptide/02-conditional.jl:170).  `pop.max_obs` rows.
"""
function simulate(pop::Population, neural::AbstractVecOrMat{Float64}, cond::AbstractMatrix{Float64}, max_obs::Integer; opts=CudeOpts())
    S = size(cond, 2)
    yhat = Array{Float64,3}(undef, max_obs, pop.n, S)
    stride = ndims(neural) == 1 ? 0 : size(neural, 1)
    check(ccall((:cude_simulate, libcude), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CudeNet}, Ref{CudeOpts}, Cint, Ptr{Cdouble}, Clonglong, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                pop.ctx.handle, pop.handle, pop.net, opts, S, neural, stride, cond, yhat, C_NULL), pop.ctx.handle)
    yhat
end

# --- device-resident `_optimize` (src/parameter-estimation.jl:170-183) for all selected starts of `train` (:374-376) -------------
mutable struct CudeTrainOpts
    adam_iters::Cint
    adam_lr::Cdouble
    adam_beta1::Cdouble
    adam_beta2::Cdouble
    adam_eps::Cdouble
    lbfgs_iters::Cint
    lbfgs_m::Cint
    g_tol::Cdouble
    c1::Cdouble
    rho_hi::Cdouble
    rho_lo::Cdouble
    ls_maxiter::Cint
    check_every::Cint
    CudeTrainOpts() = new()
end
function train_opts(; adam_iters=1000, adam_lr=1e-2, lbfgs_iters=1000)
    t = CudeTrainOpts()
    ccall((:cude_train_default_opts, libcude), Cvoid, (Ref{CudeTrainOpts},), t)
    t.adam_iters = adam_iters; t.adam_lr = adam_lr; t.lbfgs_iters = lbfgs_iters
    t
end

"""
    train_starts!(pop, neural, cond; adam_iters, adam_lr, lbfgs_iters, opts) -> (objective[S], lbfgs_iterations[S], status[S], evaluations)

Adam, then L-BFGS with BackTracking, for the `S` columns of `neural[P × S]` / `cond[N × S]` in lock-step on the device
(`cude_train`); the matrices hold the solutions on return.  Replaces the per-start `_optimize` calls of `train` (:374-376).
"""
function train_starts!(pop::Population, neural::Matrix{Float64}, cond::Matrix{Float64}; adam_iters=1000, adam_lr=1e-2,
                       lbfgs_iters=1000, opts=CudeOpts())
    S = size(cond, 2)
    size(neural) == (pop.nparams, S) && size(cond, 1) == pop.n || error("train_starts!: neural must be P × S and cond N × S")
    obj = Vector{Float64}(undef, S); iters = zeros(Cint, S); status = zeros(Cint, S); evals = Ref{Cint}(0)
    t = train_opts(; adam_iters, adam_lr, lbfgs_iters)
    check(ccall((:cude_train, libcude), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Ref{CudeNet}, Ref{CudeOpts}, Ref{CudeTrainOpts}, Cint, Ptr{Cdouble}, Ptr{Cdouble},
                 Ptr{Cdouble}, Ptr{Cint}, Ptr{Cint}, Ref{Cint}),
                pop.ctx.handle, pop.handle, pop.net, opts, t, S, neural, cond, obj, iters, status, evals), pop.ctx.handle)
    obj, Int.(iters), Int.(status), Int(evals[])
end

# --- the suppression example (suppression/src/suppression_model.jl) ------------------------------------------------------------------
"""
Device image of the suppression example's data: `data[3, n_obs, n_ind]` as `group_data` in suppression/suppression.jl:18-24
(u0 of individual i = `data[:, 1, i]`, :99-104), `timepoints`, `p_true = [p1, p2, p3]`; `scale` defaults to the mean over
individuals of the per-state maxima (:125).
"""
mutable struct SuppressionPopulation
    handle::Ptr{Cvoid}
    ctx::Context
    n::Int
    depth::Int
    width::Int
    nparams::Int
end
function SuppressionPopulation(data::Array{Float64,3}, timepoints::AbstractVector; p_true=[0.4, 0.9, 0.3], scale=nothing,
                               depth::Integer=5, width::Integer=3, ctx::Context=context())
    size(data, 1) == 3 || error("SuppressionPopulation: data must be 3 × n_obs × n_ind")
    t = collect(Float64, timepoints)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    sc = scale === nothing ? C_NULL : collect(Float64, scale)
    check(ccall((:cude_sup_population_create, libcude), Cint,
                (Ptr{Cvoid}, Cint, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Ref{Ptr{Cvoid}}),
                ctx.handle, size(data, 3), size(data, 2), t, data, collect(Float64, p_true), sc, t[1], t[end], h), ctx.handle)
    P = width * 5 + (depth - 1) * width * (width + 1) + width + 1        # chain(4 -> depth × width -> 1)
    pop = SuppressionPopulation(h[], ctx, size(data, 3), depth, width, P)
    finalizer(x -> ccall((:cude_sup_population_destroy, libcude), Cint, (Ptr{Cvoid},), x.handle), pop)
    pop
end

"""
    suppression_loss_grad(pop, neural, theta, lambda; opts, grad=true) -> (loss[S], g_neural[P × S], g_theta[N × S])

`suppression_loss(p, (prob, data, timepoints, λ))` (suppression_model.jl:117-130) and its `AutoForwardDiff()` gradient
(:155-156) for `S` parameter sets at once: `neural` is `P` (shared) or `P × S`, `theta` is `N × S`.
"""
function suppression_loss_grad(pop::SuppressionPopulation, neural::AbstractVecOrMat{Float64}, theta::AbstractMatrix{Float64},
                               lambda::Real; opts=CudeOpts(), grad::Bool=true)
    S = size(theta, 2)
    l = Vector{Float64}(undef, S)
    gn = grad ? Matrix{Float64}(undef, pop.nparams, S) : nothing
    gt = grad ? Matrix{Float64}(undef, pop.n, S) : nothing
    stride = ndims(neural) == 1 ? 0 : size(neural, 1)
    check(ccall((:cude_sup_loss_grad, libcude), Cint,
                (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Cint, Ref{CudeOpts}, Cint, Ptr{Cdouble}, Clonglong, Ptr{Cdouble}, Cdouble,
                 Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                pop.ctx.handle, pop.handle, pop.depth, pop.width, opts, S, neural, stride, theta, Float64(lambda),
                C_NULL, l, gn === nothing ? C_NULL : gn, gt === nothing ? C_NULL : gt), pop.ctx.handle)
    l, gn, gt
end
# the reference's own signature: p = ComponentArray(neural = …, theta = …), args = (pop, data, timepoints, λ)
suppression_loss(p, (pop, _, _, λ)::Tuple{SuppressionPopulation,Any,Any,Real}) =
    suppression_loss_grad(pop, collect(Float64, p.neural), reshape(collect(Float64, p.theta), :, 1), λ; grad=false)[1][1]

end # module
