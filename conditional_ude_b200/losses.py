"""Reference-named loss entry points (src/parameter-estimation.jl:56-140), evaluated on the GPU.

The Julia reference dispatches `loss(theta, p::Tuple)` on the tuple shape; the same three shapes
are accepted here:

  loss(theta, (model, timepoints, cpeptide_data))                 :56-68   theta.neural, theta.conditional
  loss(beta,  (model, timepoints, cpeptide_data, nn_parameters))  :93-99   beta scalar or 1-vector
  loss(theta, (models, timepoints, cpeptide_matrix))              :126-140 mean over individuals

`theta` may be a dict, a ComponentVector (below) or any object with `.neural` / `.conditional`.
Every call goes through the C ABI to the CUDA kernels; there is no CPU path.
"""
import math

import numpy as np

from .models import CPeptideConditionalUDEModel, CPeptideUDEModel, embed_ude_parameters, extract_ude_gradient
from .population import cached_population, SolverOptions


class ComponentVector(dict):
    """Minimal stand-in for ComponentArrays.ComponentArray: named blocks with attribute access."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _get(theta, name):
    if isinstance(theta, dict):
        return theta[name]
    return getattr(theta, name)


def _scalar(x):
    a = np.asarray(x, dtype=np.float64).reshape(-1)
    if a.size != 1:
        raise ValueError("exactly one conditional parameter per individual is supported")
    return float(a[0])


def _single(model, timepoints, cpeptide_data, neural, beta, opts, grad=False):
    pop = cached_population([model], np.asarray(timepoints, dtype=np.float64),
                            np.asarray(cpeptide_data, dtype=np.float64).reshape(1, -1))
    cond = np.array([[beta]])
    if not grad:
        return float(pop.loss(neural, cond, opts)[0])
    l, gn, gc = pop.loss_grad(neural, cond, opts)
    return float(l[0]), gn[0], float(gc[0, 0])


def loss(theta, p, opts=None):
    """Sum-of-squared-errors loss; Inf when the solver fails (parameter-estimation.jl:61-64)."""
    if len(p) == 4:                                   # fixed-NN, beta-only (:93-99)
        model, timepoints, cpeptide_data, nn = p
        if not isinstance(model, CPeptideConditionalUDEModel):
            raise TypeError("loss(beta, (model, t, y, nn)) needs a CPeptideConditionalUDEModel")
        return _single(model, timepoints, cpeptide_data, np.asarray(nn, dtype=np.float64), _scalar(theta), opts)
    if len(p) != 3:
        raise TypeError("loss(theta, (model, timepoints, cpeptide_data[, neural_network_parameters]))")
    first, timepoints, cpeptide_data = p
    if isinstance(first, CPeptideUDEModel):             # non-conditional UDE: theta is the plain 1-input network vector
        return _single(first, timepoints, cpeptide_data, embed_ude_parameters(theta, first.ude_chain.width), 0.0, opts)
    if isinstance(first, CPeptideConditionalUDEModel):  # single individual (:56-68)
        return _single(first, timepoints, cpeptide_data, np.asarray(_get(theta, "neural"), dtype=np.float64),
                       _scalar(_get(theta, "conditional")), opts)
    models = list(first)                                # population mean (:126-140)
    pop = cached_population(models, timepoints, cpeptide_data)
    cond = np.asarray(_get(theta, "conditional"), dtype=np.float64).reshape(1, -1)
    return float(pop.loss(np.asarray(_get(theta, "neural"), dtype=np.float64), cond, opts)[0])


def loss_and_gradient(theta, p, opts=None):
    """Value and gradient of `loss` — what `OptimizationFunction(loss, AutoForwardDiff())`
    (:231,:281,:299,:370) hands the optimiser.  Returns (loss, ComponentVector(neural=, conditional=))
    for the 3-tuples and (loss, dbeta) for the fixed-NN 4-tuple."""
    if len(p) == 4:
        model, timepoints, cpeptide_data, nn = p
        l, _, gb = _single(model, timepoints, cpeptide_data, np.asarray(nn, dtype=np.float64), _scalar(theta), opts, grad=True)
        return l, gb
    first, timepoints, cpeptide_data = p
    if isinstance(first, CPeptideUDEModel):
        w = first.ude_chain.width
        l, gn, _ = _single(first, timepoints, cpeptide_data, embed_ude_parameters(theta, w), 0.0, opts, grad=True)
        return l, extract_ude_gradient(gn, w)
    if isinstance(first, CPeptideConditionalUDEModel):
        l, gn, gb = _single(first, timepoints, cpeptide_data, np.asarray(_get(theta, "neural"), dtype=np.float64),
                            _scalar(_get(theta, "conditional")), opts, grad=True)
        return l, ComponentVector(neural=gn, conditional=np.array([gb]))
    pop = cached_population(list(first), timepoints, cpeptide_data)
    cond = np.asarray(_get(theta, "conditional"), dtype=np.float64).reshape(1, -1)
    l, gn, gc = pop.loss_grad(np.asarray(_get(theta, "neural"), dtype=np.float64), cond, opts)
    return float(l[0]), ComponentVector(neural=gn[0], conditional=gc[0])


def loss_sigma(theta, p, opts=None):
    """(n/2) log sigma^2 + SSE/(2 sigma^2); parameter-estimation.jl:70-75 (theta.sigma + model
    parameters) and :101-109 (theta.ode = [beta], theta.sigma, fixed NN)."""
    sigma = float(np.asarray(_get(theta, "sigma")).reshape(-1)[0])
    if len(p) == 4:
        error = loss(_get(theta, "ode"), p, opts)
    else:
        error = loss(theta, p, opts)
    n = len(p[1])
    return (n / 2) * math.log(sigma ** 2) + (1 / (2 * sigma ** 2)) * error


__all__ = ["loss", "loss_sigma", "loss_and_gradient", "ComponentVector", "SolverOptions"]
