#!/usr/bin/env python
"""Multi-GPU correctness: the all-reduced per-start sums of the individual-sharded population step (NCCL) against one
GPU evaluating the whole population.  torchrun --nproc-per-node N profiles/multi_gpu_check.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import bench
import conditional_ude_b200 as cu
from conditional_ude_b200.distributed import DevicePopulationShard, shard_bounds, finalize_sums

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, S = 60_000, 16
pk = bench.synthetic_population(n, 77)                       # the same global population on every rank
neural, cond = bench.synthetic_starts(n, S, 11, 78)
lo, hi = shard_bounds(n, world, rank)
sub = {k: (v[lo:hi] if isinstance(v, np.ndarray) and v.shape[:1] == (n,) else v) for k, v in pk.items()}
sub["n_ind"] = hi - lo
ctx = cu.Context(local)
shard = DevicePopulationShard(cu.Population(packed=sub, ctx=ctx), n, S, dev)
with torch.cuda.stream(shard.stream):
    shard.neural.copy_(torch.from_numpy(neural)); shard.cond.copy_(torch.from_numpy(np.ascontiguousarray(cond[:, lo:hi])))
shard.step(cu.SolverOptions())
loss, g = shard.result()
gc = shard.g_cond.cpu().numpy()
if rank == 0:
    full = cu.Population(packed=pk, ctx=ctx)
    l1, gn1, gc1 = full.loss_grad(neural, cond)
    out = {"n_gpus": world, "loss_rel_err": float(np.abs(loss / l1 - 1).max()),
           "g_neural_rel_err": float(np.abs(g - gn1).max() / np.abs(gn1).max()),
           "g_cond_bitwise_equal_on_rank0_shard": bool(np.array_equal(gc, gc1[:, lo:hi]))}
    os.write(1, (json.dumps(out) + "\n").encode())
    assert out["loss_rel_err"] < 1e-13 and out["g_neural_rel_err"] < 1e-12 and out["g_cond_bitwise_equal_on_rank0_shard"]
dist.barrier()
dist.destroy_process_group()
