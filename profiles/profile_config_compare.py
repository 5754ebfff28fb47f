#!/usr/bin/env python
"""config 4 (profiles, flat loss-only) timing for a library given by CUDE_B200_LIB: 117 x 10000 grid points, 5 repeats."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conditional_ude_b200 as cu
from helpers import ohashi_models
fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
nn = fx["cude_neural"][int(fx["cude_best_model_index"]) - 1]
m1, t, c1 = ohashi_models(fx, "train"); m2, _, c2 = ohashi_models(fx, "test")
pop = cu.Population(m1 + m2, t, np.vstack([c1, c2]), ctx=cu.Context(0))
bhat = np.full(117, -1.0)
f = lambda: cu.likelihood_profile_population(bhat, nn, pop, bhat - 10.0, bhat + 15.0, 0.1, steps=10000)
f(); ts = []
for _ in range(7):
    t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
print(json.dumps({"lib": os.environ.get("CUDE_B200_LIB", "main"), "ms_median": float(np.median(ts)) * 1e3, "ms_min": min(ts) * 1e3,
                  "kernel_ms": pop.ctx.stats()["kernel_ms"]}))
