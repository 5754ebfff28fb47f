#!/usr/bin/env python
"""Wall time of the host-driven workflows next to the kernel time inside them, and cProfile's top entries: beta-only fits of all
137 individuals (train_conditional, Fminbox), likelihood profiles of the whole population, SAEM's Metropolis steps."""
import cProfile, io, json, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conditional_ude_b200 as cu
from helpers import mixed_population
fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
models, ts, ys = mixed_population(fx)
import bench
nn = bench.stored_network()
ctx = cu.Context(0)
pop = cu.Population(packed=cu.pack_models(models, ts, ys), ctx=ctx)
def run(name, f):
    f()
    t0 = time.perf_counter(); pr = cProfile.Profile(); pr.enable(); r = f(); pr.disable(); dt = time.perf_counter() - t0
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(12)
    print("==== %s: %.3f s" % (name, dt)); print("\n".join(s.getvalue().splitlines()[4:22]))
    return r
run("train_conditional (137 individuals, Fminbox)", lambda: cu.train(pop, None, None, nn))
run("likelihood_profile_population (137 x 1000)", lambda: cu.likelihood_profile_population(np.full(137, -1.0), nn, pop, np.full(137, -4.0), np.full(137, 1.0), np.full(137, 0.1), steps=1000))
