import os, sys, time, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
import conditional_ude_b200 as cu
import bench
from helpers import mixed_population
fx = dict(np.load("/root/repo/tests/golden/cpeptide_fixtures.npz"))
nn = fx["cude_neural"][int(fx["cude_best_model_index"]) - 1]
ctx = cu.Context(0)
def timeit(fn, reps=5):
    fn(); ts=[]
    for _ in range(reps):
        t0=time.perf_counter(); fn(); ts.append(time.perf_counter()-t0)
    return float(np.median(ts))
models, ts, ys = mixed_population(fx)
pop = cu.Population(packed=cu.pack_models(models, ts, ys), ctx=ctx)
cond = np.random.default_rng(0).uniform(-4.0, 1.0, size=(1000, 137))
s1 = timeit(lambda: pop.loss_grad(nn, cond, neural_grad=False, mean=False))
# large beta-only batch: 200k synthetic individuals x 16 starts, shared network
n=200000
pk = bench.synthetic_population(n, 3)
pop2 = cu.Population(packed=pk, ctx=ctx)
c2 = np.random.default_rng(1).uniform(-2, 0, size=(16, n))
s2 = timeit(lambda: pop2.loss_grad(nn, c2, neural_grad=False, mean=False), 3)
st = ctx.stats()
print(json.dumps({"lib": os.environ.get("CUDE_B200_LIB","main"), "config2_ms": s1*1e3, "big_ms": s2*1e3, "big_kernel_ms": st["kernel_ms"], "big_evals_per_s_kernel": n*16/(st["kernel_ms"]*1e-3)}))
