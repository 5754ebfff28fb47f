#!/usr/bin/env python
"""Wall-clock of BASELINE.json's configurations 1-4 through the public host API (host buffers in, host buffers out,
synchronous), plus the suppression example's screening batch.  One JSON object on stdout.
  python profiles/configs_timing.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conditional_ude_b200 as cu
from helpers import train57, mixed_population, ohashi_models

fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
sup = dict(np.load(os.path.join(ROOT, "tests", "golden", "suppression_fixtures.npz")))
ctx = cu.Context(0)
nn = fx["cude_neural"][int(fx["cude_best_model_index"]) - 1]
out = {}


def timeit(fn, reps=5):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


# config 1: 57 training individuals, stored weights and betas, loss + full gradient (latency)
models, t, c, nn57, betas = train57(fx)
pop = cu.Population(models, t, c, ctx=ctx)
s = timeit(lambda: pop.loss_grad(nn57, betas.reshape(1, -1)), 20)
out["config1_loss_grad_57x1"] = {"trajectories": 57, "seconds": s, "note": "latency of one call (one trajectory per lane, 2 warps)"}
# config 2: beta-only, 137 individuals x 1000 starts, loss + d/dbeta
models, ts, ys = mixed_population(fx)
pop = cu.Population(packed=cu.pack_models(models, ts, ys), ctx=ctx)
cond = np.random.default_rng(0).uniform(-4.0, 1.0, size=(1000, 137))
s = timeit(lambda: pop.loss_grad(nn, cond, neural_grad=False, mean=False))
out["config2_beta_only_137x1000"] = {"trajectories": 137000, "seconds": s, "evals_per_s": 137000 / s}
# config 3: screening 57 x 25000 loss only, then 25 selected starts with gradients
models, t, c, _, _ = train57(fx)
pop = cu.Population(models, t, c, ctx=ctx)
rng = np.random.default_rng(1)
neural = np.stack(cu.initial_parameters(pop.chain, 25_000, rng=rng))
cond = cu.initial_parameters(57, -2.0, 0.0, 25_000, rng).T
s = timeit(lambda: pop.loss(neural, cond))
out["config3_screening_57x25000_loss_only"] = {"trajectories": 57 * 25000, "seconds": s, "evals_per_s": 57 * 25000 / s}
s = timeit(lambda: pop.loss_grad(neural[:25], cond[:25]), 20)
out["config3_selected_57x25_loss_grad"] = {"trajectories": 57 * 25, "seconds": s, "evals_per_s": 57 * 25 / s}
# config 4: profiles, 117 individuals x 1000 and x 10000 grid points, loss only
m1, t, c1 = ohashi_models(fx, "train"); m2, _, c2 = ohashi_models(fx, "test")
models, c = m1 + m2, np.vstack([c1, c2])
pop = cu.Population(models, t, c, ctx=ctx)
bhat = np.full(117, -1.0)
for steps in (1000, 10000):
    s = timeit(lambda: cu.likelihood_profile_population(bhat, nn, pop, bhat - 10.0, bhat + 15.0, 0.1, steps=steps), 3)
    out["config4_profiles_117x%d" % steps] = {"trajectories": 117 * steps, "seconds": s, "evals_per_s": 117 * steps / s}
# suppression example: 37 individuals x 10000 initial networks (suppression.jl:11,39), loss only and loss + gradient of 25
data, tp = sup["group_data"], sup["timepoints"]
spop = cu.SuppressionPopulation(data, tp, ctx=ctx)
r = np.random.default_rng(2)
nns = sup["neural_0p01"][r.integers(0, 25, 10000)] + 0.05 * r.standard_normal((10000, 67))
th = r.uniform(-1, 1, (10000, 37))
s = timeit(lambda: spop.loss(nns, th, lam=0.01), 3)
out["suppression_37x10000_loss_only"] = {"trajectories": 370000, "seconds": s, "evals_per_s": 370000 / s}
s = timeit(lambda: spop.loss_grad(nns[:1000], th[:1000], lam=0.01), 3)
out["suppression_37x1000_loss_grad"] = {"trajectories": 37000, "seconds": s, "evals_per_s": 37000 / s}
print(json.dumps(out, indent=1))
