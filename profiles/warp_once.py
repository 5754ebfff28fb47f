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
S = int(sys.argv[1])
neural = nn[None] + 0.1 * rng.standard_normal((S, 37)); cond = np.tile(betas, (S, 1)) + 0.2 * rng.standard_normal((S, 57))
for b in (3, 4):
    for _ in range(3): pop.loss_grad(neural, cond, opts=cu.SolverOptions(balance=b))
print(ctx.stats())
