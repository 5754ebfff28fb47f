#!/usr/bin/env python
"""Does lane balancing (cude_opts.balance) survive moving parameters?  Device-resident Adam on a synthetic population:
kernel time per iteration with the grouping refreshed every 8 iterations vs natural order.
  python profiles/adam_balance_experiment.py [individuals] [iterations]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
import conditional_ude_b200 as cu
from conditional_ude_b200.distributed import DevicePopulationShard

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 33
S = 64
dev = torch.device("cuda", 0)
out = {}
for bal in (0, 1):
    ctx = cu.Context(0)
    pop = cu.Population(packed=bench.synthetic_population(n, 1000, bench.simulate_gpu(ctx)), ctx=ctx)
    neural, cond = bench.synthetic_starts(n, S, 11, 2000)
    shard = DevicePopulationShard(pop, n, S, dev)
    split = int(os.environ.get("CUDE_EXP_SPLIT", "1"))
    with torch.cuda.stream(shard.stream):
        shard.neural.copy_(torch.from_numpy(neural)); shard.cond.copy_(torch.from_numpy(cond))
    opts = cu.SolverOptions(balance=bal, split=split)
    ms, losses = [], []
    for it in range(iters):
        shard.adam_step(lr=1e-2, opts=opts)
        st = ctx.stats()                     # synchronises
        ms.append(st["kernel_ms"])
        losses.append(float(shard.result()[0].mean()))
    out["balance_%d" % bal] = {"kernel_ms": [round(x, 2) for x in ms], "mean_loss_first_last": [losses[0], losses[-1]],
                               "evals_per_s_steady": n * S / (np.mean(ms[9:]) * 1e-3)}
out["gain"] = out["balance_1"]["evals_per_s_steady"] / out["balance_0"]["evals_per_s_steady"]
print(json.dumps(out))
