"""SAEM for the c-peptide cUDE (reference src/saem.jl) on top of the batched GPU loss.

The reference runs, per iteration, a Metropolis–Hastings step per individual (two solves each, `mcmc_step` :86-108), a
log-likelihood evaluation per individual (:185) and five Adam / L-BFGS iterations on (network weights, sigma) with the
total negative log-likelihood as objective (`update_population_parameters` :118-131) — all of it serial loops over
individuals.  Within one iteration the individuals are independent (the network, sigma, Omega and the prior are fixed
while they are visited), so here every such loop is ONE batched call on a device-resident `Population`:

  * an MH sweep = `population.loss(neural, [current; proposed])` — two "starts" of the shared network (flat indexing);
  * the population update = `population.loss_grad(neural, current)` per optimiser iteration: d total_nll / d neural =
    sum_i d sse_i / d neural / (2 sigma^2), d / d sigma analytic.

Names, arguments and defaults follow src/saem.jl; randomness comes from a numpy Generator (Julia's global RNG stream is
not reproducible here).  The SAEM right-hand side subtracts glucose(0) instead of glucose(timepoints[1]) (saem.jl:25);
the two coincide when the first time point is 0 (Ohashi, the only data set the reference runs SAEM on) and this module
requires that.
"""
import numpy as np

from .estimation import _as_population, adam_batched, lbfgs_batched

__all__ = ["individual_log_likelihood", "total_nll", "map_objective", "compute_individual_maps", "mcmc_step",
           "update_population_parameters", "SAEM"]

_LOG_SQRT_2PI = 0.5 * np.log(2.0 * np.pi)


def _normal_logpdf(x, mu, sd):
    """logpdf(Normal(mu, sd), x) — Distributions.jl parametrises Normal by its standard deviation (saem.jl:73, :91)."""
    z = (np.asarray(x, dtype=np.float64) - mu) / sd
    return -0.5 * z * z - np.log(sd) - _LOG_SQRT_2PI


def _n_obs(pop):
    n = getattr(pop, "n_obs", None)
    if n is None:
        raise ValueError("population must expose n_obs (observations per individual)")
    return np.asarray(n, dtype=np.float64)


def individual_log_likelihood(p_individuals, p_neural, population, sigma, opts=None):
    """saem.jl:55-67 for all individuals at once: -(n_i/2) log sigma^2 - sse_i / (2 sigma^2); -Inf where the solve failed.
    p_individuals: [N] or [S x N] (S parameter sets per individual) -> same shape."""
    pop = _as_population(population, None, None)
    p = np.asarray(p_individuals, dtype=np.float64)
    cond = p.reshape(-1, pop.n_ind)
    sse = pop.loss(np.asarray(p_neural, dtype=np.float64), cond, opts, return_sse=True)[1]
    ll = -(0.5 * _n_obs(pop))[None, :] * np.log(sigma ** 2) - sse / (2.0 * sigma ** 2)
    ll = np.where(np.isfinite(sse), ll, -np.inf)
    return ll.reshape(p.shape)


def total_nll(p_individuals, p_neural, population, sigma, opts=None):
    """saem.jl:110-116."""
    return float(-np.sum(individual_log_likelihood(p_individuals, p_neural, population, sigma, opts)))


def map_objective(p_individuals, p_neural, population, sigma, Omega, prior_individual=0.0, opts=None):
    """saem.jl:71-75, vectorised over individuals: -(log-likelihood + log prior)."""
    ll = individual_log_likelihood(p_individuals, p_neural, population, sigma, opts)
    return -(ll + _normal_logpdf(p_individuals, prior_individual, Omega))


def compute_individual_maps(p_individuals, p_neural, population, sigma, Omega, prior_individual=0.0, opts=None, maxiters=1000):
    """saem.jl:77-84: per-individual MAP estimate by L-BFGS — all individuals in lock-step, one beta-only gradient call
    (forward-sensitivity kernel) per iteration."""
    pop = _as_population(population, None, None)
    nn = np.asarray(p_neural, dtype=np.float64)
    n, nobs = pop.n_ind, _n_obs(pop)

    def f(x):
        return map_objective(x[:, 0], nn, pop, sigma, Omega, prior_individual, opts)

    def fg(x):
        b = x[:, 0]
        _, _, gc, sse = pop.loss_grad(nn, b.reshape(1, n), opts, neural_grad=False, mean=False, return_sse=True)
        val = (0.5 * nobs) * np.log(sigma ** 2) + sse[0] / (2.0 * sigma ** 2) - _normal_logpdf(b, prior_individual, Omega)
        g = gc[0] / (2.0 * sigma ** 2) + (b - prior_individual) / Omega ** 2
        return np.where(np.isfinite(sse[0]), val, np.inf), g.reshape(n, 1)

    x, _, _, _ = lbfgs_batched(f, fg, np.asarray(p_individuals, dtype=np.float64).reshape(n, 1).copy(), maxiters=maxiters)
    return x[:, 0]


def mcmc_step(p_individuals, p_neural, population, sigma, Omega, proposal_std, rng, prior_individual=0.0, temperature=1.0,
              opts=None):
    """saem.jl:86-108 for all individuals at once: random-walk proposal, prior ratio, tempered likelihood ratio, accept.
    Returns (new parameters [N], accepted [N] bool).  One loss call: current and proposed parameters as two starts."""
    p = np.asarray(p_individuals, dtype=np.float64)
    p_new = p + rng.standard_normal(p.shape) * proposal_std
    prior_ratio = _normal_logpdf(p_new, prior_individual, Omega) - _normal_logpdf(p, prior_individual, Omega)
    ll = individual_log_likelihood(np.stack([p, p_new]), p_neural, population, sigma, opts)
    with np.errstate(invalid="ignore"):
        likelihood_ratio = ll[1] / temperature - ll[0] / temperature
        accepted = np.log(rng.random(p.shape)) < (prior_ratio + likelihood_ratio)      # NaN (Inf - Inf) compares false
    return np.where(accepted, p_new, p), accepted


def update_population_parameters(p_individuals, p_neural, population, sigma, Omega=None, use_LBFGS=False, opts=None,
                                 maxiters=5):
    """saem.jl:118-131: `maxiters` iterations of Adam(1e-2) (or L-BFGS with backtracking) on (neural, sigma) with
    total_nll as objective.  Returns (neural, sigma)."""
    pop = _as_population(population, None, None)
    n, P = pop.n_ind, pop.n_params
    cond = np.asarray(p_individuals, dtype=np.float64).reshape(1, n)
    ntot = float(np.sum(_n_obs(pop)))

    def fg(x):
        nn, sg = x[0, :P], x[0, P]
        l, gn, _ = pop.loss_grad(nn, cond, opts, neural_grad=True, mean=False)     # l = sum_i sse_i, gn = sum_i d sse_i
        if not np.isfinite(l[0]):
            return np.array([np.inf]), np.zeros((1, P + 1))
        val = 0.5 * ntot * np.log(sg ** 2) + l[0] / (2.0 * sg ** 2)
        g = np.concatenate([gn[0] / (2.0 * sg ** 2), [ntot / sg - l[0] / sg ** 3]])
        return np.array([val]), g[None, :]

    def f(x):
        return np.array([total_nll(cond[0], x[0, :P], pop, x[0, P], opts)])

    x0 = np.concatenate([np.asarray(p_neural, dtype=np.float64), [float(sigma)]])[None, :]
    if use_LBFGS:
        x, _, _, _ = lbfgs_batched(f, fg, x0, maxiters=maxiters)
    else:
        x, _ = adam_batched(fg, x0, lr=1e-2, maxiters=maxiters)
    return x[0, :P].copy(), float(x[0, P])


def SAEM(individuals, initial_neural_params, network=None, sigma=1.0, prior_eta=0.0, prior_Omega=1.0, iterations=500,
         n_burnin_iterations=100, proposal_std=0.1, proposal_std_bounds=(1e-3, 1.0), alpha=0.7, n_mcmc_steps=1,
         initial_mcmc_steps=None, target_acceptance_rate=0.25, initial_temperature=10.0, temperature_decay=0.05,
         Omega_learning_rate=0.04, rng=None, opts=None, callback=None):
    """saem.jl:134-237.  `individuals`: a Population (or an object with the same loss / loss_grad interface) holding the
    individuals' glucose, c-peptide and kinetics; `network` is implied by it.  Returns a dict with the reference's fields
    (p_neural, p_individuals, Omega, sigma, eta, total_nll_values, acceptance_rates)."""
    pop = _as_population(individuals, None, None)
    if np.any(np.asarray(getattr(pop, "t_first", 0.0)) != 0.0):
        raise ValueError("SAEM: the reference's right-hand side subtracts glucose(0) (saem.jl:25); this implementation "
                         "covers populations whose first time point is 0 (Ohashi)")
    rng = rng or np.random.default_rng()
    if initial_mcmc_steps is None:
        initial_mcmc_steps = n_mcmc_steps
    n = pop.n_ind
    p_individuals = np.full(n, float(prior_eta))
    p_neural = np.asarray(initial_neural_params, dtype=np.float64).copy()
    Omega, eta, sigma = float(prior_Omega), float(prior_eta), float(sigma)
    total_nll_values, acceptance_rates = [], []
    for iteration in range(1, iterations + 1):
        burnin = iteration <= n_burnin_iterations
        gamma = 1.0 if burnin else 1.0 / (iteration - n_burnin_iterations) ** alpha
        steps = initial_mcmc_steps if burnin else n_mcmc_steps
        temperature = max(1.0, initial_temperature * np.exp(-temperature_decay * iteration))
        acceptance_count = 0
        for _ in range(steps):                                  # every individual takes the step at once (:176-184)
            p_new, acc = mcmc_step(p_individuals, p_neural, pop, sigma, Omega, proposal_std, rng, prior_individual=eta,
                                   temperature=temperature, opts=opts)
            acceptance_count += int(acc.sum())
            p_individuals = (1.0 - gamma) * p_individuals + gamma * p_new          # stochastic update
        loglikelihood = float(np.sum(individual_log_likelihood(p_individuals, p_neural, pop, sigma, opts)))     # :185
        # population parameters (:189-198): Adam during burn-in; afterwards the reference calls the same function with
        # its default use_LBFGS = false as well
        p_neural_new, sigma = update_population_parameters(p_individuals, p_neural, pop, sigma, Omega, use_LBFGS=False, opts=opts)
        p_neural = (1.0 - gamma) * p_neural + gamma * p_neural_new
        Omega = (1.0 - Omega_learning_rate) * Omega + Omega_learning_rate * float(np.var(p_individuals, ddof=1))
        eta = (1.0 - Omega_learning_rate) * eta + Omega_learning_rate * float(np.mean(p_individuals))
        acceptance_rate = acceptance_count / (n * steps)
        total_nll_values.append(-loglikelihood)
        acceptance_rates.append(acceptance_rate)
        log_proposal_std = np.log(proposal_std) + gamma * (acceptance_rate - target_acceptance_rate)
        if not burnin:
            proposal_std = float(np.clip(np.exp(log_proposal_std), proposal_std_bounds[0], proposal_std_bounds[1]))
        if callback is not None:
            callback(iteration, -loglikelihood, acceptance_rate, proposal_std, sigma)
    return dict(p_neural=p_neural, p_individuals=p_individuals, Omega=Omega, sigma=sigma, eta=eta,
                total_nll_values=np.array(total_nll_values), acceptance_rates=np.array(acceptance_rates))
