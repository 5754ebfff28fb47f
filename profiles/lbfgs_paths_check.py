import os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import conditional_ude_b200 as cu
from helpers import train57
fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
models, t, c, nn, betas = train57(fx)
ctx = cu.Context(0)
pop = cu.Population(models, t, c, ctx=ctx)
rng = np.random.default_rng(1)
G = 25000
neural0 = np.stack(cu.initial_parameters(pop.chain, G, rng=rng)); cond0 = cu.initial_parameters(57, -2.0, 0.0, G, rng).T
l = pop.loss(neural0, cond0)
best = np.argsort(l)[:25]
nb, cb = neural0[best], cond0[best]
r = pop.train_starts(nb, cb, adam_iters=1000, lbfgs_iters=0, opts=cu.SolverOptions(balance=3))
for name, o in (("fused", cu.SolverOptions(balance=3)), ("warp", cu.SolverOptions(balance=4))):
    r2 = pop.train_starts(r[0], r[1], adam_iters=0, lbfgs_iters=1000, opts=o)
    print(name, "evals", r2[5], "iters", r2[3].tolist(), "status", r2[4].tolist())
    print("  obj", np.round(np.sort(r2[2]), 5).tolist()[:8])
    f = pop.loss_grad(r[0], r[1], opts=o)
    print("  loss0", f[0][:3], np.abs(f[1]).max())
a = pop.loss_grad(r[0], r[1], opts=cu.SolverOptions(balance=3)); b = pop.loss_grad(r[0], r[1], opts=cu.SolverOptions(balance=4))
print("max rel diff loss %.2e gn %.2e gc %.2e" % (np.abs(a[0]-b[0]).max()/np.abs(a[0]).max(), np.abs(a[1]-b[1]).max()/np.abs(a[1]).max(), np.abs(a[2]-b[2]).max()/np.abs(a[2]).max()))
