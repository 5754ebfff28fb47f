#!/usr/bin/env python
"""One rank of the process-per-GPU population step (run under torchrun by tests/test_multi_gpu.py and by
`gpurun --gpus N`): the individuals are sharded over the ranks, the per-start sums are all-reduced INSIDE the library
(cude_comm_init_rank + cude_loss_grad_sharded / cude_eval_dev + cude_allreduce_dev), and rank 0 compares with one GPU
evaluating the whole population.  Prints one JSON line on rank 0; exit code != 0 on disagreement."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import bench
import conditional_ude_b200 as cu
from conditional_ude_b200.distributed import DevicePopulationShard, shard_bounds, init_library_comm

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, S = int(os.environ.get("CUDE_TEST_N", "60000")), 16
pk = bench.synthetic_population(n, 77)                       # the same global population on every rank
neural, cond = bench.synthetic_starts(n, S, 11, 78)
lo, hi = shard_bounds(n, world, rank)
sub = {k: (v[lo:hi] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v) for k, v in pk.items()}
sub["n_ind"] = hi - lo

ctx = cu.Context()                                           # no device named: LOCAL_RANK decides (one GPU per rank)
assert ctx.device == local
pop = cu.Population(packed=sub, ctx=ctx)
if world == 1:
    ctx.comm_init(1, 0, cu.comm_unique_id())                 # one rank: still through NCCL (loaded at run time by the library)
assert init_library_comm(ctx) == world and ctx.comm_size == world and ctx.comm_rank == rank
# (1) host-buffer collective call
cond_loc = np.ascontiguousarray(cond[:, lo:hi])
loss_h, gn_h, gc_h = pop.loss_grad_sharded(neural, cond_loc, n)
lossonly_h, _, _ = pop.loss_grad_sharded(neural, cond_loc, n, loss_only=True)
# (2) device-resident step: eval kernel -> partial reduction -> cude_allreduce_dev on the same stream
shard = DevicePopulationShard(pop, n, S, dev)
assert shard.lib_comm == (world > 1)
with torch.cuda.stream(shard.stream):
    shard.neural.copy_(torch.from_numpy(neural)); shard.cond.copy_(torch.from_numpy(cond_loc))
shard.step(cu.SolverOptions())
loss_d, gn_d = shard.result()
gc_d = shard.g_cond.cpu().numpy()
shard.close()
ok = True
if rank == 0:
    full = cu.Population(packed=pk, ctx=cu.Context(local))
    l1, gn1, gc1 = full.loss_grad(neural, cond)
    out = {"n_gpus": world, "nccl": cu._lib.load().cude_nccl_version(),
           "host_call": {"loss_rel_err": float(np.abs(loss_h / l1 - 1).max()),
                         "g_neural_rel_err": float(np.abs(gn_h - gn1).max() / np.abs(gn1).max()),
                         "g_cond_bitwise": bool(np.array_equal(gc_h, gc1[:, lo:hi])),
                         "loss_only_equals_grad_call": bool(np.allclose(lossonly_h, loss_h, rtol=1e-14, atol=0))},
           "device_call": {"loss_rel_err": float(np.abs(loss_d / l1 - 1).max()),
                           "g_neural_rel_err": float(np.abs(gn_d - gn1).max() / np.abs(gn1).max()),
                           "g_cond_bitwise": bool(np.array_equal(gc_d, gc1[:, lo:hi]))}}
    os.write(1, (json.dumps(out) + "\n").encode())
    for k in ("host_call", "device_call"):
        ok &= out[k]["loss_rel_err"] < 1e-13 and out[k]["g_neural_rel_err"] < 1e-12 and out[k]["g_cond_bitwise"]
    ok &= out["host_call"]["loss_only_equals_grad_call"]
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
