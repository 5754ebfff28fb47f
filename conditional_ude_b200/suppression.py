"""Host mirror of the suppression example (suppression/src/suppression_model.jl) on the second CUDA kernel variant.

  neural_network_model(depth, width; input_dims)          :78-86   (suppression.jl:18 builds depth 5, width 3, input_dims 4)
  suppression_loss(p, (prob, individual_data, timepoints, lambda))   :117-130   p.theta, p.neural
  fit_suppression_model / validate_suppression_model      :132-226  (select best initials, Adam then L-BFGS)

The reference solves the individuals with `EnsembleThreads()` and differentiates with ForwardDiff; here every
(individual, parameter set) pair is one GPU thread of `cude_sup_kernel`, value and exact discrete-adjoint gradient in
one launch.  All trajectory arithmetic is in csrc/cude_sup_kernel.cuh; this file marshals and runs the (batched, host)
optimiser loops of estimation.py.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from .estimation import OptimizationSolution, adam_batched, lbfgs_batched
from .losses import ComponentVector
from .models import Chain
from .population import SolverOptions, default_context, _dptr

P_TRUE = (0.4, 0.9, 0.3)        # suppression.jl:19


def neural_network_model(depth, width, input_dims=2):
    """(:78-86) `depth` tanh layers of `width`, softplus output.  NB the argument order (depth, width) — the reference
    script passes hidden_layer_dim = 5 as the *depth* (SURVEY.md 3.8)."""
    return Chain(int(input_dims), int(width), int(depth))


class SuppressionPopulation:
    """Device image of `individual_data` [3 x n_obs x n_ind] + the time grid: u0_i = data[:,1,i], tspan = (t[1], t[end])."""

    def __init__(self, individual_data, timepoints, p_true=P_TRUE, network=None, scale=None, tspan=None, ctx=None):
        self.ctx = ctx or default_context()
        self._lib = self.ctx._lib
        data = np.asarray(individual_data, dtype=np.float64)
        if data.ndim != 3 or data.shape[0] != 3:
            raise ValueError("individual_data must be [3 x n_obs x n_ind]")
        self.n_obs, self.n_ind = data.shape[1], data.shape[2]
        self.network = network or neural_network_model(5, 3, input_dims=4)
        if self.network.input_dims != 4:
            raise ValueError("the suppression network takes [u1; u2; u3; exp(theta)]")
        self.n_params = self.network.n_params
        t = np.ascontiguousarray(timepoints, dtype=np.float64)
        t0, tend = (float(t[0]), float(t[-1])) if tspan is None else map(float, tspan)
        dj = np.ascontiguousarray(data.transpose(2, 1, 0))          # Julia column-major 3 x n_obs x n_ind
        pt = np.ascontiguousarray(p_true, dtype=np.float64)
        sc = None if scale is None else np.ascontiguousarray(scale, dtype=np.float64)
        h = C.c_void_p()
        _lib.check(self._lib.cude_sup_population_create(self.ctx.handle, self.n_ind, self.n_obs, _dptr(t), _dptr(dj), _dptr(pt),
                                                        _dptr(sc), t0, tend, C.byref(h)), self.ctx.handle)
        self._h = h
        self._fin = weakref.finalize(self, self._lib.cude_sup_population_destroy, h)

    def _call(self, neural, theta, lam, grad, opts, return_sse):
        neural = np.ascontiguousarray(neural, dtype=np.float64)
        theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(-1, self.n_ind))
        S, P = theta.shape[0], self.n_params
        if neural.ndim == 1:
            if neural.size != P:
                raise ValueError(f"neural must have {P} entries")
            stride = 0
        else:
            if neural.shape != (S, P):
                raise ValueError(f"neural must be [{S} x {P}] or [{P}]")
            stride = P
        o = (opts or SolverOptions()).c()
        loss = np.empty(S)
        gn = np.empty((S, P)) if grad else None
        gt = np.empty((S, self.n_ind)) if grad else None
        sse = np.empty((S, self.n_ind)) if return_sse else None
        _lib.check(self._lib.cude_sup_loss_grad(self.ctx.handle, self._h, self.network.depth, self.network.width, C.byref(o), S,
                                                _dptr(neural), stride, _dptr(theta), float(lam), _dptr(sse), _dptr(loss),
                                                _dptr(gn), _dptr(gt)), self.ctx.handle)
        out = (loss, gn, gt) if grad else (loss,)
        return out + (sse,) if return_sse else (out if grad else loss)

    def loss(self, neural, theta, lam=0.0, opts=None, return_sse=False):
        """suppression_loss for S parameter sets: neural [P] (shared) or [S x P], theta [S x N]."""
        return self._call(neural, theta, lam, False, opts, return_sse)

    def loss_grad(self, neural, theta, lam=0.0, opts=None, return_sse=False):
        """(loss[S], g_neural[S x P], g_theta[S x N]) of suppression_loss (ridge term included)."""
        return self._call(neural, theta, lam, True, opts, return_sse)


def _get(p, name):
    return p[name] if isinstance(p, dict) else getattr(p, name)


def suppression_loss(p, args, opts=None):
    """suppression_loss(p, (prob, individual_data, timepoints, lambda)) — :117-130.  `prob` may be a
    SuppressionPopulation built from the same data (re-used) or anything else (ignored: the problem is fixed by the
    data, the time grid and p_true)."""
    prob, individual_data, timepoints, lam = args
    pop = prob if isinstance(prob, SuppressionPopulation) else SuppressionPopulation(individual_data, timepoints)
    return float(pop.loss(np.asarray(_get(p, "neural"), dtype=np.float64), np.asarray(_get(p, "theta"), dtype=np.float64)[None], lam, opts)[0])


def fit_suppression_model(p_init, prob, data, timepoints, lam, select_best_n=1, adam_iters=2000, lbfgs_iters=2000,
                          adam_lr=1e-3, opts=None):
    """:132-178 — screen all initials (one launch), keep the best n, Adam (Optimisers.Adam() default eta = 1e-3) then
    L-BFGS on every selected start in lock-step.  Returns (optsols, loss_traces)."""
    pop = prob if isinstance(prob, SuppressionPopulation) else SuppressionPopulation(data, timepoints)
    P, N = pop.n_params, pop.n_ind
    neural0 = np.stack([np.asarray(_get(p, "neural"), dtype=np.float64) for p in p_init])
    theta0 = np.stack([np.asarray(_get(p, "theta"), dtype=np.float64) for p in p_init])
    initial_losses = pop.loss(neural0, theta0, lam, opts)
    best = np.argsort(initial_losses, kind="stable")[:max(1, select_best_n)]
    print(f"Selected best {best.size} initials")
    x0 = np.concatenate([neural0[best], theta0[best]], axis=1)
    traces = [[] for _ in best]

    def f(x):
        return pop.loss(x[:, :P], x[:, P:], lam, opts)

    def fg(x):
        l, gn, gt = pop.loss_grad(x[:, :P], x[:, P:], lam, opts)
        for k, v in enumerate(l):
            traces[k].append(float(v))
        return l, np.concatenate([gn, gt], axis=1)

    x1, _ = adam_batched(fg, x0, lr=adam_lr, maxiters=adam_iters)
    x2, fx, iters, conv = lbfgs_batched(f, fg, x1, maxiters=lbfgs_iters)
    sols = []
    for k in range(x2.shape[0]):
        if not np.isfinite(fx[k]):
            print("Optimization failed")
            continue
        sols.append(OptimizationSolution(ComponentVector(neural=x2[k, :P].copy(), theta=x2[k, P:].copy()), fx[k], iters[k], conv[k]))
    return sols, traces


def validate_suppression_model(p_init, prob, data, timepoints, network_params, lbfgs_iters=2000, opts=None):
    """:180-226 — theta-only fit on new individuals with the network fixed (lambda = 0): best initial, then L-BFGS.
    Returns (theta, objective)."""
    pop = prob if isinstance(prob, SuppressionPopulation) else SuppressionPopulation(data, timepoints)
    nn = np.asarray(network_params, dtype=np.float64)
    th0 = np.stack([np.asarray(p, dtype=np.float64) for p in p_init])
    x0 = th0[int(np.argmin(pop.loss(nn, th0, 0.0, opts)))][None]

    def f(x):
        return pop.loss(nn, x, 0.0, opts)

    def fg(x):
        l, _, gt = pop.loss_grad(nn, x, 0.0, opts)
        return l, gt

    x, fx, _, _ = lbfgs_batched(f, fg, x0, maxiters=lbfgs_iters)
    if not np.isfinite(fx[0]):
        print("Optimization failed")
        return x0[0], np.inf
    return x[0], float(fx[0])
