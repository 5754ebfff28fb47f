#!/usr/bin/env python
"""SAEM (src/saem.jl, the settings of c-peptide/06-saem.jl:76-94: 82 training individuals, 180 iterations, 25 MCMC steps per
iteration) through the batched GPU loss: wall-clock on one B200."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conditional_ude_b200 as cu
from helpers import ohashi_models

fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
models, t, c = ohashi_models(fx, "train")
nn = fx["cude_neural"][int(fx["cude_best_model_index"]) - 1]
pop = cu.Population(models, t, c, ctx=cu.Context(0))
kw = dict(sigma=0.5, prior_eta=-1.0, prior_Omega=1.0, n_burnin_iterations=80, proposal_std=0.8, proposal_std_bounds=(1e-3, 10.0),
          alpha=0.7, n_mcmc_steps=25, initial_mcmc_steps=25, target_acceptance_rate=0.35, initial_temperature=2.0,
          temperature_decay=0.2, Omega_learning_rate=0.04)
cu.SAEM(pop, nn, iterations=2, rng=np.random.default_rng(0), **kw)
t0 = time.perf_counter()
r = cu.SAEM(pop, nn, iterations=180, rng=np.random.default_rng(1), **kw)
dt = time.perf_counter() - t0
solves = 180 * (25 * 2 + 1 + 6 * 1) * len(models)
print(json.dumps({"individuals": len(models), "iterations": 180, "mcmc_steps_per_iteration": 25, "seconds": dt,
                  "ode_solves": solves, "nll_first_last": [float(r["total_nll_values"][0]), float(r["total_nll_values"][-1])],
                  "acceptance_rate_last": float(r["acceptance_rates"][-1]), "sigma": r["sigma"], "Omega": r["Omega"]}))
