#!/usr/bin/env python
"""Where `train` (reference settings: 25 000 guesses, 25 selected, 1000 Adam + <= 1000 L-BFGS) spends its time: screening,
the device-resident Adam phase, the L-BFGS phase.  python profiles/train_phases.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conditional_ude_b200 as cu
from helpers import train57
fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
models, t, c, nn, betas = train57(fx)
ctx = cu.Context(0)
pop = cu.Population(models, t, c, ctx=ctx)
rng = np.random.default_rng(1)
out = {}
def timed(f, n=3):
    f(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), r
G = 25000
neural0 = np.stack(cu.initial_parameters(pop.chain, G, rng=rng)); cond0 = cu.initial_parameters(57, -2.0, 0.0, G, rng).T
out["screening_25000_s"], l = timed(lambda: pop.loss(neural0, cond0))
best = np.argsort(l)[:25]
nb, cb = neural0[best], cond0[best]
for name, o in (("fused", cu.SolverOptions(balance=3)), ("warp", cu.SolverOptions(balance=4))):
    out["adam_1000_%s_s" % name], r = timed(lambda: pop.train_starts(nb, cb, adam_iters=1000, lbfgs_iters=0, opts=o))
    out["adam_100_%s_s" % name], _ = timed(lambda: pop.train_starts(nb, cb, adam_iters=100, lbfgs_iters=0, opts=o))
    out["lbfgs_after_adam_%s_s" % name], r2 = timed(lambda: pop.train_starts(r[0], r[1], adam_iters=0, lbfgs_iters=1000, opts=o))
    out["lbfgs_iterations_%s" % name] = float(np.mean(r2[3])); out["evaluations_lbfgs_%s" % name] = int(r2[5])
out["train_total_s"], _ = timed(lambda: cu.train(pop, t, c, np.random.default_rng(1)), 2)
print(json.dumps(out, indent=1))
