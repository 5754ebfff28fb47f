"""Likelihood profiles (src/likelihood-profiles.jl), with the grid evaluated as ONE GPU batch.

  likelihood_profile(beta, nn, model, t, y, lb, ub, sigma; steps)   :4-17
  likelihood_profile_population(...)   batched over individuals (the loop at 02-conditional.jl:371-377)
  find_confidence_intervals(loss_values, loss_minimum, parameter_values; target)   :34-59
"""
import numpy as np

from .population import Population, cached_population

CHISQ1_095 = 3.841458820694124  # quantile(Chisq(1), 0.95)


def likelihood_profile(beta, neural_network_parameters, model, timepoints, cpeptide_data, lower_bound, upper_bound=None,
                       sigma=None, steps=1000, opts=None):
    """Returns (nll_values[steps], nll_minimum, parameter_values[steps]) like the reference: the loss at
    `beta` and on range(lower_bound, upper_bound, length=steps), each scaled by 1/(2 sigma^2).
    The reference's second method (:19-32), likelihood_profile(beta, loss_function, args, lb, ub, sigma; steps), is
    dispatched on a callable second argument; with this package's `loss` and a fixed-network 4-tuple as `args` the grid
    still runs as one GPU batch, any other callable is evaluated point by point."""
    if callable(neural_network_parameters):
        loss_function, args = neural_network_parameters, model
        lower_bound, upper_bound, sigma = timepoints, cpeptide_data, lower_bound      # positional shift of that method
        from .losses import loss as _loss
        if loss_function is _loss and len(args) == 4:
            return likelihood_profile(beta, args[3], args[0], args[1], args[2], lower_bound, upper_bound, sigma, steps=steps, opts=opts)
        parameter_values = np.linspace(lower_bound, upper_bound, steps)
        scale = 1.0 / (2.0 * sigma ** 2)
        return (np.array([scale * loss_function(b, args) for b in parameter_values]), scale * loss_function(beta, args),
                parameter_values)
    pop = cached_population([model], np.asarray(timepoints, dtype=np.float64),
                            np.asarray(cpeptide_data, dtype=np.float64).reshape(1, -1))
    parameter_values = np.linspace(lower_bound, upper_bound, steps)
    b0 = float(np.asarray(beta, dtype=np.float64).reshape(-1)[0])
    cond = np.concatenate([[b0], parameter_values]).reshape(-1, 1)
    sse = pop.loss(np.asarray(neural_network_parameters, dtype=np.float64), cond, opts)
    scale = 1.0 / (2.0 * sigma ** 2)
    return scale * sse[1:], scale * sse[0], parameter_values


def likelihood_profile_population(betas, neural_network_parameters, population, lower_bounds, upper_bounds, sigmas,
                                  steps=1000, opts=None, distributed=False, group=None):
    """All individuals' profiles in one launch: grid[s, i] = lower_i + s (upper_i - lower_i)/(steps-1).
    Returns (nll[steps x N], nll_minimum[N], parameter_values[steps x N]).
    distributed=True: the grid points are split over the torch.distributed ranks (no communication on the data path;
    the sse rows are all-gathered) and every rank returns the full profile."""
    from .estimation import _Comm
    if not (isinstance(population, Population) or (hasattr(population, "loss") and hasattr(population, "n_ind"))):
        raise TypeError("population must be a Population")
    comm = _Comm(distributed, group)
    n = population.n_ind
    lb = np.broadcast_to(np.asarray(lower_bounds, dtype=np.float64), (n,))
    ub = np.broadcast_to(np.asarray(upper_bounds, dtype=np.float64), (n,))
    sig = np.broadcast_to(np.asarray(sigmas, dtype=np.float64), (n,))
    grid = np.linspace(lb, ub, steps)                       # [steps x N]
    cond = np.vstack([np.asarray(betas, dtype=np.float64).reshape(1, n), grid])
    lo, hi = comm.bounds(steps + 1)
    nn = np.asarray(neural_network_parameters, dtype=np.float64)
    sse = population.loss(nn, cond[lo:hi], opts, return_sse=True)[1] if hi > lo else np.empty((0, n))
    sse = comm.allgather_rows(sse, steps + 1)
    scale = 1.0 / (2.0 * sig ** 2)
    return sse[1:] * scale, sse[0] * scale, grid


def find_confidence_intervals(loss_values, loss_minimum, parameter_values, target="cantelli95"):
    """src/likelihood-profiles.jl:34-59 (host-side bookkeeping, identical thresholds)."""
    if target == "cantelli95":
        threshold = loss_minimum + 7.16
    elif target == "cantelli90":
        threshold = loss_minimum + 5.24
    else:
        if target != "raue95":
            print(f"Unknown target {target}. Using raue 95%")
        threshold = loss_minimum + CHISQ1_095
    loss_values = np.asarray(loss_values)
    idx = np.flatnonzero(loss_values <= threshold)
    if idx.size == 0:
        raise ValueError("no profile value below the threshold")   # Julia: minimum of an empty collection throws
    lo, hi = idx.min(), idx.max()
    lo_v = -np.inf if lo == 0 else parameter_values[lo]
    hi_v = np.inf if hi == len(parameter_values) - 1 else parameter_values[hi]
    return lo_v, hi_v
