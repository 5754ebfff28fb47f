"""Known-answer checks against the reference's stored result artifacts.

source_data/cude_neural_parameters.jld2 holds 25 trained weight sets, 25 x 57 betas and
best_model_index = 14 (c-peptide/02-conditional.jl:44-50).  beta_i only enters individual i's term of
the population loss (parameter-estimation.jl:130-131), so the stored betas must be stationary points of
loss_i under the stored weights — a check that involves the kinetics, the interpolant, the input order
[dG; beta], the weight layout and the solver at once.
"""
import numpy as np

import conditional_ude_b200 as cu
from oracle import oracle
from helpers import train57, ohashi_models


def test_train_split_class_counts(fx):
    # stratified_split(rng, types, 0.7), src/utils.jl:15-31: round(0.7 * {36, 34, 12}) = {25, 24, 8}
    types = fx["ohashi_train_types"][fx["train_split_idx"]]
    assert (np.sum(types == "T2DM"), np.sum(types == "NGT"), np.sum(types == "IGT")) == (25, 24, 8)
    assert np.all(np.diff(fx["train_split_idx"]) > 0)          # stratified_split sorts the indices


def test_stored_betas_are_stationary_points(fx):
    models, t, c, nn, betas = train57(fx)
    op = oracle.OraclePopulation(cu.pack_models(models, t, c))
    g = op.eval(nn, betas, grad_mode=0)
    dbeta = np.abs(g["g_cond"][0])
    assert dbeta.max() < 0.3 and dbeta.mean() < 0.06          # max 0.25, mean 0.043
    # away from the optimum the same derivative is O(1): the check is not vacuous
    g_off = op.eval(nn, betas + 0.5, grad_mode=0)
    assert np.abs(g_off["g_cond"][0]).mean() > 10 * dbeta.mean()
    # each stored beta is a local minimiser along its own axis
    lo = op.eval(nn, betas - 0.05)["sse"][0]
    hi = op.eval(nn, betas + 0.05)["sse"][0]
    at = g["sse"][0]
    assert np.mean((at <= lo + 1e-3) & (at <= hi + 1e-3)) > 0.9


def test_stored_weights_are_a_population_optimum(fx):
    models, t, c, nn, betas = train57(fx)
    op = oracle.OraclePopulation(cu.pack_models(models, t, c))
    p = op.population_loss(nn, betas[None], with_grad=True)
    assert abs(p["loss"][0] - 0.428) < 1e-3                    # mean train SSE 0.428
    assert np.linalg.norm(p["g_neural"][0]) < 0.2              # 0.13 (central differences)
    per_ind = op.eval(nn, betas, grad_mode=0)["g_neural"][0]
    assert np.abs(per_ind).max() > 10 * np.abs(p["g_neural"][0]).max()   # the terms cancel at the optimum


def test_best_model_beats_a_wrong_alignment(fx):
    """With the betas assigned to the wrong individuals the loss is far worse: the inferred split is real."""
    models, t, c, nn, betas = train57(fx)
    op = oracle.OraclePopulation(cu.pack_models(models, t, c))
    good = op.population_loss(nn, betas[None])["loss"][0]
    bad = op.population_loss(nn, np.roll(betas, 7)[None])["loss"][0]
    assert bad > 3 * good


def test_covariate_artifact(fx):
    """cude_covariate_neural_parameters_2.jld2 (07-covariate-inclusion.jl:59-65): 3-input network,
    stored betas of the best model (2) are near-stationary on the same split."""
    models, t, c = ohashi_models(fx, "train", covariate=True)
    idx = fx["train_split_idx"]
    pk = cu.pack_models([models[i] for i in idx], t, c[idx])
    best = int(fx["cov_best_model_index"]) - 1
    op = oracle.OraclePopulation(pk)
    g = op.eval(fx["cov_neural"][best], fx["cov_betas"][best], grad_mode=0)
    off = op.eval(fx["cov_neural"][best], fx["cov_betas"][best] + 0.5, grad_mode=0)
    assert np.abs(g["g_cond"]).mean() < 0.2 * np.abs(off["g_cond"]).mean()
    assert g["sse"].mean() < 1.0


def test_all_fifty_stored_optima_are_stationary(fx):
    """Both stored training runs (`cude_neural_parameters.jld2`, `cude_neural_parameters_sigma.jld2`): 2 x 25 networks,
    each with its own 57 fitted betas.  Every one of the 50 (network, betas) pairs the reference's optimiser stopped at
    must be a near-stationary point of the restated loss: |d loss_i/d beta_i| small (>= 30x smaller than half a unit
    away), and the per-individual network gradients cancelling in the population mean.  A wrong kinetic constant,
    interpolant, input order or weight layout would break all fifty at once."""
    models, t, c, nn, betas = train57(fx)
    op = oracle.OraclePopulation(cu.pack_models(models, t, c))
    for name in ("cude", "cude_sigma"):
        W, B = fx[name + "_neural"], fx[name + "_betas"]
        r = op.eval(W, B, grad_mode=0)                                   # 25 starts x 57 individuals
        off = op.eval(W, B + 0.5, grad_mode=0)
        db, dboff = np.abs(r["g_cond"]).mean(axis=1), np.abs(off["g_cond"]).mean(axis=1)
        assert db.max() < 0.5 and np.median(db) < 0.08 and (dboff / db).min() > 25
        pop_grad = np.linalg.norm(r["g_neural"].sum(axis=1) / 57, axis=1)
        largest_term = np.abs(r["g_neural"]).max(axis=(1, 2))
        assert (pop_grad / largest_term).max() < 0.12
        loss = r["sse"].mean(axis=1)
        assert 0.25 < loss.min() and loss.max() < 0.5                    # the runs' objectives: 0.28 - 0.45
    assert int(fx["cude_sigma_best_model_index"]) == 2
