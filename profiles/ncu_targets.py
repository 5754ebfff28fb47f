#!/usr/bin/env python
"""One call of each kernel the bench's secondary block measures, for `ncu --set full` captures:
  python profiles/ncu_targets.py fused|loss_flat|bsens|sup_loss|sup_grad|screening"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import conditional_ude_b200 as cu
what = sys.argv[1]
ctx = cu.Context(0)
fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
nn = bench.stored_network()
if what == "fused":
    n, S = 1_000_000, 8
    pop = cu.Population(packed=bench.synthetic_population(n, 1000, bench.simulate_gpu(ctx)), ctx=ctx)
    neural, cond = bench.synthetic_starts(n, S, 11, 2000)
    pop.loss_grad(neural, cond, mean=False)
elif what == "loss_flat":
    m, t, y = bench._fixture_models(fx, cu, ["train", "test"])
    pop = cu.Population(m, fx["ohashi_timepoints"], np.stack(y), ctx=ctx)
    pop.loss(nn, np.linspace(np.full(117, -11.0), np.full(117, 14.0), 10000))
elif what == "bsens":
    m, t, y = bench._fixture_models(fx, cu, ["train", "test", "fujita"])
    pop = cu.Population(packed=cu.pack_models(m, t, y), ctx=ctx)
    pop.loss_grad(nn, np.random.default_rng(0).uniform(-4.0, 1.0, size=(1000, len(m))), neural_grad=False, mean=False)
elif what == "screening":
    m, t, y = bench._fixture_models(fx, cu, ["train"])
    idx = fx["train_split_idx"]
    pop = cu.Population([m[i] for i in idx], fx["ohashi_timepoints"], np.stack([y[i] for i in idx]), ctx=ctx)
    rng = np.random.default_rng(1)
    pop.loss(np.stack(cu.initial_parameters(pop.chain, 25_000, rng=rng)), cu.initial_parameters(57, -2.0, 0.0, 25_000, rng).T)
else:
    sup = dict(np.load(os.path.join(ROOT, "tests", "golden", "suppression_fixtures.npz")))
    spop = cu.SuppressionPopulation(sup["group_data"], sup["timepoints"], ctx=ctx)
    r = np.random.default_rng(2)
    nns = sup["neural_0p01"][r.integers(0, 25, 10000)] + 0.05 * r.standard_normal((10000, 67))
    th = r.uniform(-1, 1, (10000, 37))
    (spop.loss if what == "sup_loss" else spop.loss_grad)(nns, th, lam=0.01)
print(ctx.stats())
