#!/usr/bin/env python
"""BASELINE config 3 with the starts sharded over the ranks (no communication on the data path):
  torchrun --nproc-per-node N profiles/sharded_train_demo.py   (or plain python for one GPU)
57 Ohashi training individuals, 25 000 LHS/Glorot initial guesses screened, the best 24 trained (Adam 200 + L-BFGS 100
here instead of 1000 + 1000).  Rank 0 prints one JSON line: wall-clock of screening+training and the best objectives."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import conditional_ude_b200 as cu
from helpers import train57

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
fx = dict(np.load(os.path.join(ROOT, "tests", "golden", "cpeptide_fixtures.npz")))
models, t, c, nn, betas = train57(fx)
pop = cu.Population(models, t, c, ctx=cu.Context(local))
kw = dict(initial_guesses=25_000, selected_initials=24, number_of_iterations_adam=200, number_of_iterations_lbfgs=100)
cu.train(pop, t, c, np.random.default_rng(1), distributed=world > 1, **dict(kw, number_of_iterations_adam=2, number_of_iterations_lbfgs=2))
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
sols = cu.train(pop, t, c, np.random.default_rng(1), distributed=world > 1, **kw)
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
obj = sorted(s.objective for s in sols)
if rank == 0:
    os.write(1, (json.dumps({"n_gpus": world, "seconds": dt, "solutions": len(sols), "best_objectives": obj[:5],
                             "stored_reference_best_is": 0.428}) + "\n").encode())
if world > 1:
    dist.destroy_process_group()
