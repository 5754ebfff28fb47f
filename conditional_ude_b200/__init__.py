"""conditional_ude_b200 — B200-native hot path of Computational-Biology-TUe/conditional-ude:
the batched forward solve and gradient of the c-peptide conditional-UDE loss.

Host side mirrors the reference's constructors and loss entry points; all numerics run in the
hand-written sm_100a CUDA kernels of csrc/ through the C ABI in include/cude_b200.h.
There is no CPU fallback.
"""
from .models import (Chain, chain, softplus, van_cauter_parameters, CPeptideConditionalUDEModel, CPeptideUDEModel,
                     CPeptideConditionalCovariateUDEModel, pack_models)
from .population import (Context, Population, SolverOptions, default_context, MultiContext, MultiPopulation,
                         comm_unique_id, pick_device)
from .losses import loss, loss_sigma, loss_and_gradient, ComponentVector
from .profiles import likelihood_profile, likelihood_profile_population, find_confidence_intervals
from .estimation import (initial_parameters, train, train_with_sigma, evaluate_model, stratified_split, argmedian,
                         OptimizationSolution)

from .suppression import (SuppressionPopulation, neural_network_model, suppression_loss, fit_suppression_model,
                          validate_suppression_model)
from .dataprep import prepare_ohashi, prepare_fujita, split_like_reference
from .saem import (SAEM, mcmc_step, individual_log_likelihood, total_nll, map_objective, compute_individual_maps,
                   update_population_parameters)

__all__ = [
    "SuppressionPopulation", "neural_network_model", "suppression_loss", "fit_suppression_model", "validate_suppression_model",
    "Chain", "chain", "softplus", "van_cauter_parameters", "CPeptideConditionalUDEModel", "CPeptideUDEModel",
    "CPeptideConditionalCovariateUDEModel", "pack_models", "Context", "Population", "SolverOptions",
    "default_context", "MultiContext", "MultiPopulation", "comm_unique_id", "pick_device", "loss", "loss_sigma", "loss_and_gradient", "ComponentVector",
    "likelihood_profile", "likelihood_profile_population", "find_confidence_intervals",
    "initial_parameters", "train", "train_with_sigma", "evaluate_model", "stratified_split", "argmedian",
    "OptimizationSolution",
    "SAEM", "mcmc_step", "individual_log_likelihood", "total_nll", "map_objective", "compute_individual_maps",
    "update_population_parameters",
    "prepare_ohashi", "prepare_fujita", "split_like_reference",
]
