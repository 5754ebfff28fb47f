#!/usr/bin/env python
"""`train` at the reference's settings (src/parameter-estimation.jl:340-348: 25 000 initial guesses, 25 selected, Adam 1000
iterations at 1e-2, L-BFGS 1000 iterations) on the 57-individual Ohashi training split: device-resident optimisers
(cude_train) against the round-1 host loop (numpy optimisers, one GPU call per evaluation).
  python profiles/train_timing.py [adam_iters] [lbfgs_iters]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conditional_ude_b200 as cu
from helpers import train57
fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
models, t, c, nn, betas = train57(fx)
na = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
pop = cu.Population(models, t, c, ctx=cu.Context(0))
out = {"settings": {"individuals": 57, "initial_guesses": 25000, "selected_initials": 25, "adam_iters": na, "lbfgs_iters": nl}}
for name, dev in (("device_optimisers", True), ("host_optimisers_round1", False)):
    t0 = time.perf_counter()
    sols = cu.train(pop, t, c, np.random.default_rng(1), number_of_iterations_adam=na, number_of_iterations_lbfgs=nl, device_optimizer=dev)
    dt = time.perf_counter() - t0
    obj = np.sort([s.objective for s in sols])
    out[name] = {"seconds": dt, "n_solutions": len(sols), "best": float(obj[0]), "median": float(np.median(obj)), "worst": float(obj[-1]),
                 "lbfgs_iterations_mean": float(np.mean([s.iterations for s in sols]))}
out["speedup"] = out["host_optimisers_round1"]["seconds"] / out["device_optimisers"]["seconds"]
print(json.dumps(out))
