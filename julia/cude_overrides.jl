# cude_overrides.jl — include AFTER the reference's own sources and CUDEB200.jl:
#
#     include("src/parameter-estimation.jl"); include("src/likelihood-profiles.jl")
#     include("CUDEB200.jl"); using .CUDEB200
#     include("cude_overrides.jl")
#
# It re-defines, with IDENTICAL signatures, the methods of the reference that sit on the hot path
#   loss(θ, (model, timepoints, cpeptide_data))                                src/parameter-estimation.jl:56
#   loss(θ, (model, timepoints, cpeptide_data, neural_network_parameters))     :93
#   loss(θ, (models, timepoints, cpeptide_data))                               :126
#   likelihood_profile(β, nn, model, timepoints, cpeptide_data, lb, ub, σ)     src/likelihood-profiles.jl:4
#   suppression_loss(p, (prob, individual_data, timepoints, λ))                suppression/src/suppression_model.jl:117
# so that `train` (:211, :272, :340), `train_with_sigma` (:290), `evaluate_model` (:406) and the scripts run unmodified on
# the GPU.  The reference differentiates with `OptimizationFunction(loss, AutoForwardDiff())` (:231, :281, :299, :370):
# ForwardDiff calls `loss` with dual numbers, which cannot cross a `ccall`.  The methods below therefore accept duals,
# evaluate value AND gradient of the primal point on the GPU (one adjoint pass), and return the dual
# value + Σ_i g_i · partials(θ_i): exactly what ForwardDiff would have propagated through the discrete solve with the step
# sequence frozen (DESIGN.md section 2).  ForwardDiff's chunks (8 sweeps for 94 parameters) hit a one-entry cache.
#
# NOT EXECUTED IN THIS REPOSITORY (no Julia in the build image): reviewed against the reference's signatures only.
using ForwardDiff
using ForwardDiff: Dual, Partials, value, partials
using ComponentArrays: ComponentArray

_primal(x) = x isa AbstractArray ? Float64.(value.(x)) : Float64(value(x))
_ndual(::Type{Dual{T,V,N}}) where {T,V,N} = N
# dual result from a primal value, the gradient entries and the matching dual inputs
function _lift(val::Float64, grads::AbstractVector{Float64}, inputs::AbstractVector{D}) where {D<:Dual}
    N = _ndual(D)
    Dual{ForwardDiff.tagtype(D)}(val, Partials(ntuple(k -> sum(grads[i] * partials(inputs[i], k) for i in eachindex(inputs)), N)))
end

const _last_grad = Ref{Any}(nothing)       # (key, loss, g_neural, g_cond) of the last gradient evaluation
function _cached_loss_grad(pop, neural::Vector{Float64}, cond::Vector{Float64}; neural_grad=true, mean=true)
    key = (objectid(pop), neural, cond, neural_grad, mean)
    c = _last_grad[]
    if c === nothing || c[1] != key
        l, gn, gc = CUDEB200.loss_grad(pop, neural_grad ? neural : neural, reshape(cond, :, 1); mean, neural_grad)
        c = (key, l[1], gn === nothing ? Float64[] : vec(gn), vec(gc))
        _last_grad[] = c
    end
    c[2], c[3], c[4]
end

# --- :126 population loss ------------------------------------------------------------------------------------------------
function loss(θ, (models, timepoints, cpeptide_data)::Tuple{AbstractVector{CPeptideConditionalUDEModel}, AbstractVector{T}, AbstractVecOrMat{T}}) where T <: Real
    pop = CUDEB200.cached_population(models, timepoints, cpeptide_data)
    neural, cond = vec(_primal(θ.neural)), vec(_primal(θ.conditional))
    if eltype(θ) <: Dual
        l, gn, gc = _cached_loss_grad(pop, neural, cond)
        return _lift(l, vcat(gn, gc), vcat(vec(collect(θ.neural)), vec(collect(θ.conditional))))
    end
    CUDEB200.loss(pop, neural, reshape(cond, :, 1))[1]
end

# --- :56 one individual, network + beta in θ -----------------------------------------------------------------------------
function loss(θ, (model, timepoints, cpeptide_data)::Tuple{M, AbstractVector{T}, AbstractVector{T}}) where T <: Real where M <: CPeptideModel
    loss(θ, ([model], timepoints, reshape(cpeptide_data, 1, :)))      # mean over one individual = its SSE
end

# --- :93 one individual, network fixed, θ = beta (scalar or 1-vector) ----------------------------------------------------
function loss(θ, (model, timepoints, cpeptide_data, neural_network_parameters)::Tuple{M, AbstractVector{T}, AbstractVector{T}, AbstractVector{T}}) where T <: Real where M <: CPeptideModel
    pop = CUDEB200.cached_population(model, timepoints, reshape(cpeptide_data, 1, :))
    nn = collect(Float64, neural_network_parameters)
    β = first(θ)
    if β isa Dual
        l, _, gc = _cached_loss_grad(pop, nn, [Float64(value(β))]; neural_grad=false, mean=false)
        return _lift(l, gc, [β])
    end
    CUDEB200.loss(pop, nn, fill(Float64(β), 1, 1))[1]
end

# --- likelihood profile: the whole grid in one launch --------------------------------------------------------------------
function likelihood_profile(β, neural_network_parameters, model, timepoints, cpeptide_data, lower_bound, upper_bound, sigma; steps=1000)
    pop = CUDEB200.cached_population(model, timepoints, reshape(cpeptide_data, 1, :))
    CUDEB200.likelihood_profile(β, neural_network_parameters, pop, lower_bound, upper_bound, sigma; steps)
end

# --- suppression example: suppression/src/suppression_model.jl:117 --------------------------------------------------------
# Include after suppression/src/suppression_model.jl.  `prob` carries `ude_lsup!(du,u,p,t)`, a global method that closes over
# nothing readable (suppression.jl:19), so the true kinetic parameters and the network shape of the script are stated here;
# set them before the first call if your script differs (suppression.jl:17-19: depth 5, width 3, p_true = [0.4, 0.9, 0.3]).
const SUPPRESSION_P_TRUE = Ref([0.4, 0.9, 0.3])
const SUPPRESSION_NET = Ref((depth = 5, width = 3))
const _sup_pop_cache = IdDict{Any,Any}()
function _suppression_population(individual_data, timepoints)
    get!(_sup_pop_cache, individual_data) do
        CUDEB200.SuppressionPopulation(Array{Float64,3}(individual_data), collect(Float64, timepoints);
                                       p_true = SUPPRESSION_P_TRUE[], depth = SUPPRESSION_NET[].depth, width = SUPPRESSION_NET[].width)
    end
end
function suppression_loss(p, (prob, individual_data, timepoints, λ))
    pop = _suppression_population(individual_data, timepoints)
    neural, theta = vec(_primal(p.neural)), vec(_primal(p.theta))
    if eltype(p) <: Dual                     # OptimizationFunction(suppression_loss, AutoForwardDiff()), :155-156
        l, gn, gt = CUDEB200.suppression_loss_grad(pop, neural, reshape(theta, :, 1), λ)
        # the gradient in the order of the dual inputs gathered below: theta, then neural
        return _lift(l[1], vcat(vec(gt), vec(gn)), vcat(vec(collect(p.theta)), vec(collect(p.neural))))
    end
    CUDEB200.suppression_loss_grad(pop, neural, reshape(theta, :, 1), λ; grad=false)[1][1]
end
