#!/usr/bin/env python
"""Small end-to-end case running every kernel instantiation once (written for compute-sanitizer, which is closed on
this pool; the host-compiled kernel source passes the emulation tests under -fsanitize=address,undefined instead).
  compute-sanitizer --tool memcheck python profiles/sanitizer_case.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conditional_ude_b200 as cu
import bench
from helpers import mixed_population, random_starts

fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
sup = dict(np.load(os.path.join(ROOT, "tests", "golden", "suppression_fixtures.npz")))
ctx = cu.Context(0)
models, ts, ys = mixed_population(fx)                    # ragged: 5 and 14 knots
pk = cu.pack_models(models, ts, ys)
pop = cu.Population(packed=pk, ctx=ctx)
rng = np.random.default_rng(0)
neural, cond = random_starts(rng, pk["chain"], len(models), 3)
l = pop.loss(neural, cond)                               # loss only, tile mode
l2 = pop.loss(neural[0], cond)                           # loss only, flat mode
g = pop.loss_grad(neural, cond)                          # adjoint: automatic = warp per trajectory (small batch)
tight = dict(abstol=1e-10, reltol=1e-7)                  # beyond 64 / 32 recorded steps: the fused-kernel fallbacks
for bal in (2, 3, 4):                                    # two-kernel gradient, fused kernel, warp per trajectory
    pop.loss_grad(neural, cond, opts=cu.SolverOptions(balance=bal))
    pop.loss_grad(neural, cond, opts=cu.SolverOptions(balance=bal, **tight))
pop.loss_grad(neural, cond, opts=cu.SolverOptions(balance=2, precision=2))   # FP32 adjoint kernel
pop.loss_grad(neural, cond, opts=cu.SolverOptions(balance=3, split=2))       # split pipeline
pop.simulate(neural, cond)
b = pop.loss_grad(neural[0], cond, neural_grad=False)    # forward sensitivity, flat
m = pop.loss_grad(neural, cond, opts=cu.SolverOptions(precision=1))   # mixed precision
# lane balancing + pipelined host call on a population large enough to enable both (kept small for the sanitizer)
n = 4200
pk2 = bench.synthetic_population(n, 3)
pop2 = cu.Population(packed=pk2, ctx=ctx)
ne, co = bench.synthetic_starts(n, 2, 11, 4)
ob = cu.SolverOptions(balance=1)
for _ in range(2):
    s_, gc_ = pop2.loss_grad_sums(ne, co, 1.0, ob)
spop = cu.SuppressionPopulation(sup["group_data"], sup["timepoints"], ctx=ctx)
sl = spop.loss_grad(sup["neural_0p01"][:2], np.zeros((2, 37)), lam=0.01)
print("ok", float(l[0]), float(g[0][0]), float(b[0][0]), float(m[0][0]), float(s_[0, 0]), float(sl[0][0]))
