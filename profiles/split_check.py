#!/usr/bin/env python
"""Split gradient pipeline vs fused kernel: agreement and kernel time.  python profiles/split_check.py [individuals] [starts]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import conditional_ude_b200 as cu

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ctx = cu.Context(0)
pk = bench.synthetic_population(n, 1000, bench.simulate_gpu(ctx))
neural, cond = bench.synthetic_starts(n, S, 11, 2000)
pop = cu.Population(packed=pk, ctx=ctx)
out = {"individuals": n, "starts": S}
res = {}
for name, o in (("fused", cu.SolverOptions(split=1)), ("split", cu.SolverOptions(split=2)),
                ("fused_p2", cu.SolverOptions(split=1, precision=2)), ("split_p2", cu.SolverOptions(split=2, precision=2)),
                ("split_bal", cu.SolverOptions(split=2, balance=1)), ("exact", cu.SolverOptions(balance=2)),
                ("exact_p2", cu.SolverOptions(balance=2, precision=2))):
    ms = []
    for it in range(4):
        r = pop.loss_grad(neural, cond, opts=o, mean=False, return_sse=True)
        st = ctx.stats()
        ms.append(st["kernel_ms"])
    res[name] = r
    out[name] = {"kernel_ms": [round(x, 2) for x in ms], "evals_per_s": n * S / (min(ms[1:]) * 1e-3), "launches": st["launches"],
                 "n_acc": st["n_acc"], "n_fail": st["n_fail"]}
def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())
f = res["fused"]
for name in ("split", "split_bal", "split_p2", "exact", "exact_p2"):
    r = res[name]
    out[name]["vs_fused"] = {"sse_bitwise": bool(np.array_equal(r[3], f[3])), "loss_rel": rel(r[0], f[0]), "g_neural_rel": rel(r[1], f[1]),
                             "g_cond_rel": rel(r[2], f[2])}
out["split_p2"]["vs_fused_p2"] = {"g_neural_rel": rel(res["split_p2"][1], res["fused_p2"][1]), "g_cond_rel": rel(res["split_p2"][2], res["fused_p2"][2])}
print(json.dumps(out))
