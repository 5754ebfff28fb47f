#!/usr/bin/env python
"""Latency of one loss+gradient evaluation of the 25 selected starts (57 individuals x 25, config 3) through the device
API, fused kernel vs warp-per-trajectory latency kernel vs split pipeline: kernel ms (events) and wall time per call in a back-to-back loop."""
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
S = int(sys.argv[1]) if len(sys.argv) > 1 else 25
neural = nn[None] + 0.1 * rng.standard_normal((S, 37)); cond = np.tile(betas, (S, 1)) + 0.2 * rng.standard_normal((S, 57))
out = {}
for name, o in (("fused", cu.SolverOptions(balance=3)), ("warp_per_trajectory", cu.SolverOptions(balance=4)), ("split", cu.SolverOptions(balance=3, split=2))):
    for _ in range(5): pop.loss_grad(neural, cond, opts=o)
    k = []
    t0 = time.perf_counter()
    for _ in range(200):
        pop.loss_grad(neural, cond, opts=o); k.append(ctx.stats()["kernel_ms"])
    out[name] = {"kernel_ms_median": float(np.median(k)), "wall_ms_per_call": (time.perf_counter() - t0) / 200 * 1e3, "launches": ctx.stats()["launches"]}
print(json.dumps(out))
