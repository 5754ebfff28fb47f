"""Reference-named estimation drivers on top of the GPU loss/gradient.

  initial_parameters(chain, n; rng) / initial_parameters(n_models, lb, ub, n, rng)   src/parameter-estimation.jl:22-24, 36-38
  train(models, t, Y, rng; ...)            multi-start cUDE training: LHS + Glorot guesses, screening, top-k,
                                           Adam then L-BFGS                                   :340-386, _optimize :170-183
  train(models, t, Y, nn; ...)             beta-only fits with fixed network                  :272-288, :159-168
  train_with_sigma(models, t, Y, nn; ...)  beta + sigma                                       :290-307
  evaluate_model(models, t, Y, nns, betas) candidate-network x validation-individual matrix   :406-433
  stratified_split, argmedian              src/utils.jl:15-31, :43-45

The reference runs one optimiser at a time on one CPU thread, each objective call re-solving every ODE with
dual numbers.  Here every optimiser iteration is ONE batched GPU call: all starts of the multi-start training (or
all individuals of a beta-only fit) advance in lock-step; the optimisers themselves are small vectorised numpy
loops over [n_problems x n_dims] arrays (host bookkeeping, no trajectory arithmetic).

Parity note.  Optimisers.Adam / Optim.LBFGS / LineSearches.BackTracking / Fminbox are third-party Julia packages that
cannot be run here; the restatements below follow their published algorithms (Adam beta=(0.9,0.999), eps=1e-8, best
iterate kept; L-BFGS m=10 two-loop recursion, BackTracking c1=1e-4, rho in [0.1,0.5] with quadratic/cubic
interpolation, g_tol=1e-8).
Box constraints: `fminbox_batched` restates Optim.Fminbox's log-barrier method (the default for the bounded beta fits);
projection is kept as an option.  Training steps 1 and 2 of the multi-start training run on the device (cude_train).
Julia's RNG streams (StableRNG, QuasiMonteCarlo, SimpleChains.init_params) are not reproducible: sampling uses numpy.
The end-to-end check is on *results*: re-fitting beta on the Ohashi train split with the stored weights recovers the
stored betas (tests/test_estimation_gpu.py).
"""
import numpy as np

from .losses import ComponentVector
from .population import Population, SolverOptions


# ----------------------------------------------------------------------------- sampling
def initial_parameters(*args, rng=None):
    """initial_parameters(chain, n_initials; rng) -> list of Glorot-normal weight vectors (:22-24), or
    initial_parameters(n_models, lhs_lb, lhs_ub, n_initials, rng) -> [n_models x n_initials] Latin hypercube (:36-38)."""
    if len(args) == 2:
        chain, n_initials = args
        rng = rng or np.random.default_rng()
        return list(chain.init_params_batch(rng, int(n_initials)))       # rows of one [n x P] draw
    n_models, lb, ub, n_initials = args[:4]
    rng = args[4] if len(args) > 4 else (rng or np.random.default_rng())
    # LatinHypercubeSample: one point per stratum in every dimension, strata permuted independently
    u = (np.stack([rng.permutation(n_initials) for _ in range(n_models)]) + rng.random((n_models, n_initials))) / n_initials
    return lb + (ub - lb) * u


def stratified_split(rng, types, f_train):
    """src/utils.jl:15-31 — (training_indices, testing_indices), 0-based, training indices sorted."""
    types = np.asarray(types)
    train = []
    for ty in dict.fromkeys(types.tolist()):                       # unique(), first-occurrence order
        idx = np.flatnonzero(types == ty)
        train.extend(rng.choice(idx, size=int(round(f_train * idx.size)), replace=False).tolist())
    train = np.sort(np.array(train, dtype=np.int64))
    test = np.setdiff1d(np.arange(types.size), train)
    return train, test


def argmedian(x):
    x = np.asarray(x)
    return int(np.argmin(np.abs(x - np.median(x))))


# ----------------------------------------------------------------------------- batched optimisers
class OptimizationSolution:
    """The fields the reference scripts read from Optimization.jl's solution: `.u`, `.objective`."""

    def __init__(self, u, objective, iterations=0, converged=False):
        self.u, self.objective, self.iterations, self.converged = u, float(objective), int(iterations), bool(converged)

    def __repr__(self):
        return f"OptimizationSolution(objective={self.objective:.6g}, iterations={self.iterations}, converged={self.converged})"


def adam_batched(fg, x0, lr=1e-2, maxiters=1000, beta=(0.9, 0.999), eps=1e-8, callback=None):
    """Optimisers.Adam on S independent problems in lock-step.  fg(x[S,D]) -> (f[S], g[S,D]).
    Returns the best iterate of every problem (as Optimization.jl's Optimisers wrapper does)."""
    x = np.array(x0, dtype=np.float64)
    m = np.zeros_like(x)
    v = np.zeros_like(x)
    best_x, best_f = x.copy(), np.full(x.shape[0], np.inf)
    b1t, b2t = 1.0, 1.0
    for it in range(maxiters):
        f, g = fg(x)
        better = f < best_f
        best_x[better], best_f[better] = x[better], f[better]
        ok = np.isfinite(f) & np.all(np.isfinite(g), axis=1)
        g = np.where(ok[:, None], g, 0.0)
        b1t *= beta[0]
        b2t *= beta[1]
        m = beta[0] * m + (1 - beta[0]) * g
        v = beta[1] * v + (1 - beta[1]) * g * g
        x = x - lr * (m / (1 - b1t)) / (np.sqrt(v / (1 - b2t)) + eps) * ok[:, None]
        if callback is not None and callback(it, best_f):
            break
    f, _ = fg(x)
    better = f < best_f
    best_x[better], best_f[better] = x[better], f[better]
    return best_x, best_f


def lbfgs_batched(f, fg, x0, maxiters=1000, m=10, g_tol=1e-8, lb=None, ub=None, c1=1e-4, rho_hi=0.5, rho_lo=0.1,
                  ls_maxiter=50):
    """Optim.LBFGS(m=10, linesearch=BackTracking(order=3)) on S independent problems in lock-step.
    f(x[S,D]) -> f[S];  fg(x[S,D]) -> (f[S], g[S,D]).  Bounds by projection.  Returns (x, fx, iterations, converged)."""
    x = np.array(x0, dtype=np.float64)
    S, D = x.shape
    proj = (lambda z: np.clip(z, lb, ub)) if (lb is not None or ub is not None) else (lambda z: z)
    x = proj(x)
    fx, g = fg(x)
    sh, yh, rho = np.zeros((m, S, D)), np.zeros((m, S, D)), np.zeros((m, S))
    nh = np.zeros(S, dtype=int)          # history length per problem
    head = 0
    iters = np.zeros(S, dtype=int)

    def pgrad(x_, g_):
        """projected gradient for the convergence test under bounds"""
        if lb is None and ub is None:
            return g_
        pg = g_.copy()
        if lb is not None:
            pg = np.where((x_ <= lb) & (g_ > 0), 0.0, pg)
        if ub is not None:
            pg = np.where((x_ >= ub) & (g_ < 0), 0.0, pg)
        return pg

    conv = np.abs(pgrad(x, g)).max(axis=1) <= g_tol
    active = np.isfinite(fx) & ~conv
    for it in range(maxiters):
        if not active.any():
            break
        # ---- two-loop recursion (history slots newest-first order) ----
        q = pgrad(x, g).copy()
        order = [(head - 1 - k) % m for k in range(m)]
        alphas = np.zeros((m, S))
        for k, slot in enumerate(order):
            use = nh > k
            a = np.where(use, rho[slot] * np.einsum("sd,sd->s", sh[slot], q), 0.0)
            alphas[k] = a
            q -= a[:, None] * yh[slot]
        newest = order[0]
        yy = np.einsum("sd,sd->s", yh[newest], yh[newest])
        gamma = np.where((nh > 0) & (yy > 0), np.einsum("sd,sd->s", sh[newest], yh[newest]) / np.where(yy > 0, yy, 1.0), 1.0)
        r = gamma[:, None] * q
        for k in reversed(range(m)):
            slot = order[k]
            use = nh > k
            b = np.where(use, rho[slot] * np.einsum("sd,sd->s", yh[slot], r), 0.0)
            r += np.where(use, alphas[k] - b, 0.0)[:, None] * sh[slot]
        d = -r
        dphi0 = np.einsum("sd,sd->s", g, d)
        bad = ~(dphi0 < 0)                          # not a descent direction: steepest descent, drop the history
        d[bad] = -pgrad(x, g)[bad]
        nh[bad] = 0
        dphi0 = np.einsum("sd,sd->s", g, d)
        # ---- backtracking line search (LineSearches.BackTracking, order 3), alpha0 = 1 ----
        alpha = np.ones(S)
        alpha_prev = np.ones(S)
        f_prev = fx.copy()
        searching = active.copy()
        xt, ft = x.copy(), fx.copy()
        for ls in range(ls_maxiter):
            trial = proj(x + alpha[:, None] * d)
            ftry = f(np.where(searching[:, None], trial, x))
            ok = searching & np.isfinite(ftry) & (ftry <= fx + c1 * alpha * dphi0)
            xt[ok], ft[ok] = trial[ok], ftry[ok]
            searching &= ~ok
            if not searching.any():
                break
            nonfin = searching & ~np.isfinite(ftry)
            # interpolation: quadratic on the first backtrack, cubic afterwards
            with np.errstate(all="ignore"):
                a_quad = -(dphi0 * alpha ** 2) / (2.0 * (ftry - fx - dphi0 * alpha))
                div = 1.0 / (alpha_prev ** 2 * alpha ** 2 * (alpha - alpha_prev))
                t1 = ftry - fx - dphi0 * alpha
                t2 = f_prev - fx - dphi0 * alpha_prev
                ca = (alpha_prev ** 2 * t1 - alpha ** 2 * t2) * div
                cb = (-alpha_prev ** 3 * t1 + alpha ** 3 * t2) * div
                disc = cb * cb - 3.0 * ca * dphi0
                a_cub = np.where(np.abs(ca) < 1e-300, -dphi0 / (2.0 * cb), (-cb + np.sqrt(np.maximum(disc, 0.0))) / (3.0 * ca))
            a_new = a_quad if ls == 0 else a_cub
            a_new = np.where(np.isfinite(a_new), a_new, alpha * rho_hi)
            a_new = np.clip(a_new, alpha * rho_lo, alpha * rho_hi)
            a_new = np.where(nonfin, alpha * 0.5, a_new)
            alpha_prev, f_prev = np.where(searching, alpha, alpha_prev), np.where(searching, ftry, f_prev)
            alpha = np.where(searching, a_new, alpha)
        moved = active & ~searching
        stuck = active & searching                  # line search failed: stop this problem (Optim: terminates)
        fnew, gnew = fg(np.where(moved[:, None], xt, x))
        lost = moved & ~(np.isfinite(fnew) & np.all(np.isfinite(gnew), axis=1))   # value ok but gradient pass failed there
        moved &= ~lost
        stuck |= lost
        s_vec = np.where(moved[:, None], xt - x, 0.0)
        y_vec = np.where(moved[:, None], gnew - g, 0.0)
        ys = np.einsum("sd,sd->s", y_vec, s_vec)
        store = moved & (ys > 1e-300)
        sh[head] = np.where(store[:, None], s_vec, 0.0)
        yh[head] = np.where(store[:, None], y_vec, 0.0)
        rho[head] = np.where(store, 1.0 / np.where(store, ys, 1.0), 0.0)
        head = (head + 1) % m
        nh = np.where(store, np.minimum(nh + 1, m), np.where(moved, 0, nh))
        no_change = moved & (np.abs(xt - x).max(axis=1) == 0.0)
        x = np.where(moved[:, None], xt, x)
        fx = np.where(moved, fnew, fx)
        g = np.where(moved[:, None], gnew, g)
        iters += active
        conv |= moved & (np.abs(pgrad(x, g)).max(axis=1) <= g_tol)
        active &= ~(conv | stuck | no_change) & np.isfinite(fx)
    return x, fx, iters, conv


def fminbox_batched(f, fg, x0, lb, ub, maxiters=1000, mufactor=1e-3, outer_iterations=8, g_tol=1e-8, x_tol=1e-10):
    """Optim.Fminbox(LBFGS(linesearch = BackTracking())) on S independent problems in lock-step — what Optimization.jl runs
    when `lb` / `ub` are given (reference src/parameter-estimation.jl:159-168).  Logarithmic-barrier method: the inner
    L-BFGS minimises f(x) + mu * sum(-log(x - lb) - log(ub - x)) over the open box (trial points outside the box have an
    infinite objective, so the backtracking line search keeps the iterates inside), mu starts at
    mufactor * |grad f|_1 / |grad barrier|_1 (Optim's mu0 = :auto) and is multiplied by mufactor after every outer
    iteration; it stops when the projected gradient of f is below g_tol, the iterate stops moving, or after
    `outer_iterations` (mu has then shrunk by 1e-24).  Infinite bounds carry no barrier.  A start on the boundary is moved
    inside by 1 % of the box like Optim does.  Returns (x, f(x), inner iterations, converged)."""
    x = np.array(x0, dtype=np.float64)
    S, D = x.shape
    lbv = np.broadcast_to(np.asarray(-np.inf if lb is None else lb, dtype=np.float64), (D,)).copy()
    ubv = np.broadcast_to(np.asarray(np.inf if ub is None else ub, dtype=np.float64), (D,)).copy()
    fl, fu = np.isfinite(lbv), np.isfinite(ubv)
    width = np.where(fl & fu, ubv - lbv, 1.0)
    x = np.where(fl & (x <= lbv), lbv + 0.01 * width, x)
    x = np.where(fu & (x >= ubv), ubv - 0.01 * width, x)

    def barrier(z):
        with np.errstate(all="ignore"):
            b = np.where(fl, -np.log(z - lbv), 0.0).sum(axis=1) + np.where(fu, -np.log(ubv - z), 0.0).sum(axis=1)
            gb = np.where(fl, -1.0 / (z - lbv), 0.0) + np.where(fu, 1.0 / (ubv - z), 0.0)
        inside = np.all((~fl | (z > lbv)) & (~fu | (z < ubv)), axis=1)
        return np.where(inside, b, np.inf), np.where(inside[:, None], gb, 0.0), inside

    f0, g0 = fg(x)
    _, gb0, _ = barrier(x)
    n1f, n1b = np.abs(g0).sum(axis=1), np.abs(gb0).sum(axis=1)
    mu = np.where((n1f > 0) & (n1b > 0), mufactor * n1f / np.where(n1b > 0, n1b, 1.0), 1e-3 * mufactor)
    iters = np.zeros(S, dtype=int)
    conv = np.zeros(S, dtype=bool)
    for _ in range(outer_iterations):
        def fb(z):
            b, _, inside = barrier(z)
            val = f(np.where(inside[:, None], z, x))
            return np.where(inside, val + mu * b, np.inf)

        def fgb(z):
            b, gb, inside = barrier(z)
            val, grad = fg(np.where(inside[:, None], z, x))
            return np.where(inside, val + mu * b, np.inf), np.where(inside[:, None], grad + mu[:, None] * gb, 0.0)

        xn, _, it, _ = lbfgs_batched(fb, fgb, x, maxiters=maxiters, g_tol=g_tol)
        iters += it
        moved = np.abs(xn - x).max(axis=1)
        x = xn
        fx, g = fg(x)
        pg = np.where((fl & (x - lbv <= 1e-8 * width) & (g > 0)) | (fu & (ubv - x <= 1e-8 * width) & (g < 0)), 0.0, g)
        conv = (np.abs(pg).max(axis=1) <= g_tol) | (moved <= x_tol)
        mu = mu * mufactor
        if conv.all():
            break
    return x, f(x), iters, conv


# ----------------------------------------------------------------------------- beta-only estimation
def _as_population(models, timepoints, cpeptide_data, ctx=None):
    if isinstance(models, Population) or (hasattr(models, "loss_grad") and hasattr(models, "n_params")):
        return models            # already a device population (or a test double with the same interface)
    return Population(list(models), np.asarray(timepoints, dtype=np.float64), np.asarray(cpeptide_data, dtype=np.float64), ctx=ctx)


def train_conditional(models, timepoints, cpeptide_data, neural_network_parameters, initial_beta=-2.0,
                      lbfgs_lower_bound=-4.0, lbfgs_upper_bound=1.0, lbfgs_iterations=1000, opts=None, bounds="fminbox"):
    """train(models, timepoints, cpeptide_data, neural_network_parameters; initial_beta, bounds) — :272-288:
    per-individual beta with the network fixed.  All individuals are fitted simultaneously: one GPU call per
    optimiser iteration (shared network => flat trajectory indexing).  bounds="fminbox" (default): the reference's
    Fminbox(LBFGS) barrier method; "projection": projected L-BFGS (round 1; same solutions on the reference's data, see
    tests/test_estimation_gpu.py::test_bounded_beta_fits_fminbox_and_projection_agree)."""
    pop = _as_population(models, timepoints, cpeptide_data)
    nn = np.asarray(neural_network_parameters, dtype=np.float64)
    n = pop.n_ind
    lb = None if np.isneginf(lbfgs_lower_bound) else float(lbfgs_lower_bound)
    ub = None if np.isposinf(lbfgs_upper_bound) else float(lbfgs_upper_bound)

    def f(x):
        return pop.loss(nn, x.reshape(1, n), opts, return_sse=True)[1][0]

    def fg(x):
        _, _, gc, sse = pop.loss_grad(nn, x.reshape(1, n), opts, neural_grad=False, mean=False, return_sse=True)
        return sse[0], gc[0].reshape(n, 1)

    x0 = np.broadcast_to(np.asarray(initial_beta, dtype=np.float64), (n,)).reshape(n, 1).copy()
    if bounds == "fminbox" and (lb is not None or ub is not None):
        x, fx, iters, conv = fminbox_batched(f, fg, x0, lb, ub, maxiters=lbfgs_iterations)
    else:
        x, fx, iters, conv = lbfgs_batched(f, fg, x0, maxiters=lbfgs_iterations, lb=lb, ub=ub)
    return [OptimizationSolution(np.array([x[i, 0]]), fx[i], iters[i], conv[i]) for i in range(n)]


def train_with_sigma(models, timepoints, cpeptide_data, neural_network_parameters, initial_beta=-2.0,
                     lbfgs_lower_bound=-4.0, lbfgs_upper_bound=1.0, lbfgs_iterations=1000, opts=None, bounds="fminbox"):
    """:290-307 — theta = (ode=[beta], sigma=1.0), objective (n/2) log sigma^2 + sse/(2 sigma^2), bounds on beta only."""
    pop = _as_population(models, timepoints, cpeptide_data)
    nn = np.asarray(neural_network_parameters, dtype=np.float64)
    n = pop.n_ind
    n_t = np.asarray(timepoints).size if not isinstance(timepoints, (list, tuple)) or np.ndim(timepoints[0]) == 0 else None
    nobs = float(n_t) if n_t is not None else np.array([len(t) for t in timepoints], dtype=np.float64)
    lbv = np.array([lbfgs_lower_bound, -np.inf])
    ubv = np.array([lbfgs_upper_bound, np.inf])

    def f(x):
        sse = pop.loss(nn, x[:, 0].reshape(1, n), opts, return_sse=True)[1][0]
        s2 = x[:, 1] ** 2
        return (nobs / 2) * np.log(s2) + sse / (2 * s2)

    def fg(x):
        _, _, gc, sse = pop.loss_grad(nn, x[:, 0].reshape(1, n), opts, neural_grad=False, mean=False, return_sse=True)
        sse, gb, sig = sse[0], gc[0], x[:, 1]
        s2 = sig ** 2
        val = (nobs / 2) * np.log(s2) + sse / (2 * s2)
        return val, np.stack([gb / (2 * s2), nobs / sig - sse / sig ** 3], axis=1)

    x0 = np.stack([np.broadcast_to(np.asarray(initial_beta, dtype=np.float64), (n,)), np.ones(n)], axis=1)
    if bounds == "fminbox":
        x, fx, iters, conv = fminbox_batched(f, fg, x0, lbv, ubv, maxiters=lbfgs_iterations)
    else:
        x, fx, iters, conv = lbfgs_batched(f, fg, x0, maxiters=lbfgs_iterations, lb=lbv, ub=ubv)
    return [OptimizationSolution(ComponentVector(ode=np.array([x[i, 0]]), sigma=float(x[i, 1])), fx[i], iters[i], conv[i])
            for i in range(n)]


# ----------------------------------------------------------------------------- multi-start cUDE training
class _Comm:
    """Start-sharding helper for the communication-free configurations (multi-start screening / training, profiles):
    every rank works on a contiguous slice of the starts and the (tiny) per-start results are all-gathered.
    distributed=False (or an uninitialised torch.distributed) is the single-process case."""

    def __init__(self, distributed=False, group=None):
        self.rank, self.world, self.group, self.dist = 0, 1, group, None
        if distributed:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                self.dist, self.rank, self.world = dist, dist.get_rank(group), dist.get_world_size(group)

    def bounds(self, n, rank=None):
        r = self.rank if rank is None else rank
        return r * n // self.world, (r + 1) * n // self.world

    def allgather_rows(self, local, n_total):
        """local = rows [lo:hi) of an [n_total x ...] float64 array -> the full array on every rank."""
        local = np.ascontiguousarray(local, dtype=np.float64)
        if self.world == 1:
            return local
        import torch
        full = np.zeros((n_total,) + local.shape[1:], dtype=np.float64)
        lo, hi = self.bounds(n_total)
        full[lo:hi] = local
        t = torch.from_numpy(full)
        if self.dist.get_backend(self.group) == "nccl":          # NCCL reduces device tensors only
            t = t.cuda()
        self.dist.all_reduce(t, group=self.group)                # disjoint slices: the sum is the concatenation
        return t.cpu().numpy()


def train(models, timepoints, cpeptide_data, rng_or_nn, initial_guesses=None, selected_initials=None,
          lhs_lower_bound=-2.0, lhs_upper_bound=0.0, n_conditional_parameters=1, number_of_iterations_adam=1000,
          number_of_iterations_lbfgs=1000, learning_rate_adam=1e-2, opts=None, distributed=False, group=None,
          device_optimizer=True, **kw):
    """The `train` methods of the reference, dispatched like Julia on the first and 4th argument:
      train(model::CPeptideUDEModel, t, y, rng; ...)  non-conditional UDE on one individual (:211-247; 10 000 / 10)
      train(models, t, Y, rng::Generator; ...)  full cUDE training (:340-386; 25 000 guesses / 25 selected)
      train(models, t, Y, nn::vector; initial_beta, lbfgs_lower_bound, ...)  beta-only (:272-288).
    distributed=True (one process per GPU under torch.distributed; BASELINE config "starts sharded over 8 x B200"):
    every rank draws the same initial guesses from an identically seeded `rng`, screens its slice of them, the losses
    are all-gathered, the globally best `selected_initials` starts are split over the ranks and optimised there, and
    every rank returns all solutions in selection order.  No communication on the data path.
    device_optimizer=True (default, on a device Population): training steps 1 and 2 (:170-183) run in the library's
    device-resident lock-step optimisers (cude_train: parameters, Adam moments and L-BFGS history stay in HBM, the host only
    enqueues kernels); False keeps the host numpy optimisers of this module (same algorithms, one GPU call per evaluation)."""
    from .models import CPeptideUDEModel
    if isinstance(models, CPeptideUDEModel):           # train(model::CPeptideUDEModel, t, y, rng; ...) :211-247, its own defaults
        return train_ude(models, timepoints, cpeptide_data, rng_or_nn,
                         initial_guesses=10_000 if initial_guesses is None else initial_guesses,
                         selected_initials=10 if selected_initials is None else selected_initials,
                         number_of_iterations_adam=number_of_iterations_adam,
                         number_of_iterations_lbfgs=number_of_iterations_lbfgs, learning_rate_adam=learning_rate_adam, opts=opts)
    initial_guesses = 25_000 if initial_guesses is None else initial_guesses
    selected_initials = 25 if selected_initials is None else selected_initials
    if not isinstance(rng_or_nn, np.random.Generator):
        return train_conditional(models, timepoints, cpeptide_data, rng_or_nn, opts=opts, **kw)
    if n_conditional_parameters != 1:
        raise NotImplementedError("one conditional parameter per individual")
    rng = rng_or_nn
    comm = _Comm(distributed, group)
    pop = _as_population(models, timepoints, cpeptide_data)
    n, P = pop.n_ind, pop.n_params
    # sample initial parameters (:350-357)
    neural0 = np.stack(initial_parameters(pop.chain, initial_guesses, rng=rng))
    cond0 = initial_parameters(n, lhs_lower_bound, lhs_upper_bound, initial_guesses, rng).T        # [guesses x n]
    # preselect (:360-366): ONE loss-only launch over this rank's guesses
    lo, hi = comm.bounds(initial_guesses)
    losses_initial = comm.allgather_rows(pop.loss(neural0[lo:hi], cond0[lo:hi], opts) if hi > lo else np.empty(0), initial_guesses)
    pick = np.argsort(losses_initial, kind="stable")[:selected_initials]                            # partialsortperm
    x0 = np.concatenate([neural0[pick], cond0[pick]], axis=1)

    def f(x):
        return pop.loss(x[:, :P], x[:, P:], opts)

    def fg(x):
        l, gn, gc = pop.loss_grad(x[:, :P], x[:, P:], opts)
        return l, np.concatenate([gn, gc], axis=1)

    # training step 1 (Adam) and 2 (LBFGS), :170-183 — this rank's selected starts in lock-step
    k = x0.shape[0]
    lo, hi = comm.bounds(k)
    res = np.full((hi - lo, P + n + 3), np.nan)
    if hi > lo and device_optimizer and type(pop) is Population:
        nn2, cc2, fx, iters, status, _ = pop.train_starts(x0[lo:hi, :P], x0[lo:hi, P:], adam_iters=number_of_iterations_adam,
                                                          lr=learning_rate_adam, lbfgs_iters=number_of_iterations_lbfgs, opts=opts)
        res = np.concatenate([nn2, cc2, fx[:, None], iters.astype(np.float64)[:, None], (status == 1).astype(np.float64)[:, None]], axis=1)
    elif hi > lo:
        x1, _ = adam_batched(fg, x0[lo:hi], lr=learning_rate_adam, maxiters=number_of_iterations_adam)
        x2, fx, iters, conv = lbfgs_batched(f, fg, x1, maxiters=number_of_iterations_lbfgs)
        res = np.concatenate([x2, fx[:, None], np.asarray(iters, dtype=np.float64)[:, None],
                              np.asarray(conv, dtype=np.float64)[:, None]], axis=1)
    # Inf objectives travel as a flag (the gather is a sum over ranks)
    bad = ~np.isfinite(res[:, P + n])
    res = np.where(np.isfinite(res), res, 0.0)
    res = comm.allgather_rows(np.concatenate([res, bad[:, None].astype(np.float64)], axis=1), k)
    sols = []
    for s_ in range(k):
        if res[s_, -1] != 0.0:
            print("Optimization failed... Skipping")                                                # :378-380
            continue
        sols.append(OptimizationSolution(ComponentVector(neural=res[s_, :P].copy(), conditional=res[s_, P:P + n].copy()),
                                         res[s_, P + n], int(res[s_, P + n + 1]), bool(res[s_, P + n + 2])))
    return sols


def train_ude(model, timepoints, cpeptide_data, rng, initial_guesses=10_000, selected_initials=10,
              number_of_iterations_adam=1000, number_of_iterations_lbfgs=1000, learning_rate_adam=1e-2, opts=None,
              population=None):
    """train(model::CPeptideUDEModel, timepoints, cpeptide_data, rng; ...) — src/parameter-estimation.jl:211-247: the
    non-conditional UDE on one (mean) individual.  All initial guesses are screened in one launch (a population of one
    individual x `initial_guesses` networks), the best `selected_initials` are trained in lock-step (Adam, then L-BFGS).
    Solutions carry the plain 1-input network vector in `.u` (the reference's `optsol.u`)."""
    from .models import embed_ude_parameters, extract_ude_gradient
    pop = population if population is not None else _as_population(
        [model], timepoints, np.asarray(cpeptide_data, dtype=np.float64).reshape(1, -1))   # `population`: test double
    w = model.ude_chain.width
    p0 = np.stack(initial_parameters(model.ude_chain, initial_guesses, rng=rng))

    def f(x):
        return pop.loss(embed_ude_parameters(x, w), np.zeros((x.shape[0], 1)), opts)

    def fg(x):
        l, gn, _ = pop.loss_grad(embed_ude_parameters(x, w), np.zeros((x.shape[0], 1)), opts)
        return l, extract_ude_gradient(gn, w)

    losses_initial = f(p0)
    pick = np.argsort(losses_initial, kind="stable")[:selected_initials]
    x1, _ = adam_batched(fg, p0[pick], lr=learning_rate_adam, maxiters=number_of_iterations_adam)
    x2, fx, iters, conv = lbfgs_batched(f, fg, x1, maxiters=number_of_iterations_lbfgs)
    sols = []
    for s_ in range(x2.shape[0]):
        if not np.isfinite(fx[s_]):
            print("Optimization failed... Skipping")                                                # :239-241
            continue
        sols.append(OptimizationSolution(x2[s_].copy(), fx[s_], iters[s_], conv[s_]))
    return sols


def evaluate_model(models, timepoints, cpeptide_data, neural_network_parameters, betas_train, opts=None):
    """:406-433 — for every candidate network: fit beta for every validation individual (unbounded, from the mean of
    that network's training betas) and collect the objectives.  Returns [n_individuals x n_models] like `hcat`.
    All candidate networks are fitted at once: problems = (network, individual) pairs, one launch per iteration."""
    pop = _as_population(models, timepoints, cpeptide_data)
    nns = np.stack([np.asarray(p, dtype=np.float64) for p in neural_network_parameters])
    S, n = nns.shape[0], pop.n_ind
    init = np.array([np.mean(b) for b in betas_train], dtype=np.float64)

    def f(x):
        return pop.loss(nns, x.reshape(S, n), opts, return_sse=True)[1].reshape(-1)

    def fg(x):
        _, _, gc, sse = pop.loss_grad(nns, x.reshape(S, n), opts, neural_grad=False, mean=False, return_sse=True)
        return sse.reshape(-1), gc.reshape(-1, 1)

    x0 = np.repeat(init, n).reshape(-1, 1)
    x, fx, _, _ = lbfgs_batched(f, fg, x0, maxiters=1000)
    return fx.reshape(S, n).T
