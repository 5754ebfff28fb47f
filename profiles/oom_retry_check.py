#!/usr/bin/env python
"""The two-kernel gradient sizes its groups of starts for the scratch budget decided at first use; when the device cannot give
that much any more (another allocation took it) the library halves the groups instead of failing.  A host-buffer call of
1 M individuals x 64 starts runs in 6 chunks of ~11 starts = 24 GB of step records each; a torch allocation that leaves 18 GB
forces two groups per chunk."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench, conditional_ude_b200 as cu
ctx = cu.Context(0)
n, S = 1_000_000, 64
pop = cu.Population(packed=bench.synthetic_population(n, 5, bench.simulate_gpu(ctx)), ctx=ctx)
neural, cond = bench.synthetic_starts(n, S, 11, 6)
small = pop.loss_grad(neural[:2], cond[:2], mean=False)            # fixes the budget (64 GB) while the device is empty
free, total = torch.cuda.mem_get_info(0)
hog = torch.empty(free - 18 * (1 << 30), dtype=torch.uint8, device="cuda:0")
free, total = torch.cuda.mem_get_info(0)
a = pop.loss_grad(neural, cond, mean=False)
la = ctx.stats()["launches"]
del hog; torch.cuda.empty_cache()
ctx2 = cu.Context(0)
b = cu.Population(packed=pop.packed if hasattr(pop, "packed") else bench.synthetic_population(n, 5, bench.simulate_gpu(ctx2)), ctx=ctx2).loss_grad(neural, cond, mean=False)
lb = ctx2.stats()["launches"]
print("free with hog %.1f GB; launches %d (squeezed) vs %d (free device); identical results: %s" %
      (free / 2**30, la, lb, all(np.array_equal(x, y) for x, y in zip(a, b))))
