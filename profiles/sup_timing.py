#!/usr/bin/env python
"""Suppression kernel timing (37 individuals x 10 000 starts, suppression.jl:11,39): kernel ms of loss and loss+gradient."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import conditional_ude_b200 as cu
sup = dict(np.load(os.path.join(ROOT, "tests", "golden", "suppression_fixtures.npz")))
ctx = cu.Context(0)
spop = cu.SuppressionPopulation(sup["group_data"], sup["timepoints"], ctx=ctx)
r = np.random.default_rng(2)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
o = cu.SolverOptions(block=int(sys.argv[2])) if len(sys.argv) > 2 else None      # optional: threads per block
nns = sup["neural_0p01"][r.integers(0, 25, S)] + 0.05 * r.standard_normal((S, 67))
th = r.uniform(-1, 1, (S, 37))
out = {}
of = cu.SolverOptions(block=int(sys.argv[2]), split=1) if len(sys.argv) > 2 else cu.SolverOptions(split=1)
for name, fn in (("loss", lambda: spop.loss(nns, th, lam=0.01, opts=o)), ("loss_grad", lambda: spop.loss_grad(nns, th, lam=0.01, opts=o)),
                 ("loss_grad_fused_kernel", lambda: spop.loss_grad(nns, th, lam=0.01, opts=of))):
    ms = []
    for _ in range(5):
        fn(); ms.append(ctx.stats()["kernel_ms"])
    out[name] = {"kernel_ms": round(float(np.median(ms)), 3), "evals_per_s": 37 * S / (float(np.median(ms)) * 1e-3)}
print(json.dumps(out))
