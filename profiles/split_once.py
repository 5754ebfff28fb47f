#!/usr/bin/env python
"""One loss+gradient call (for ncu launch lists): python profiles/split_once.py individuals starts split [precision] [balance]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import conditional_ude_b200 as cu
n, S, split = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prec = int(sys.argv[4]) if len(sys.argv) > 4 else 0
bal = int(sys.argv[5]) if len(sys.argv) > 5 else 0
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
ctx = cu.Context(0)
pk = bench.synthetic_population(n, 1000, bench.simulate_gpu(ctx))
neural, cond = bench.synthetic_starts(n, S, 11, 2000)
pop = cu.Population(packed=pk, ctx=ctx)
for _ in range(reps):
    pop.loss_grad(neural, cond, opts=cu.SolverOptions(split=split, precision=prec, balance=bal), mean=False)
    print(ctx.stats())
