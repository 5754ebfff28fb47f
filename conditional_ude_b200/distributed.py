"""Multi-GPU population training step: individuals sharded over ranks, one all-reduce.

The population loss of src/parameter-estimation.jl:126-140 is a mean over individuals; the
individuals (and their conditional parameters) are independent, so each rank owns a contiguous
block of them and the only exchange is the all-reduce of the per-start sums
{sum_i sse_i, sum_i d sse_i / d neural} — (P+1) x S doubles (19 KB for 64 starts).  Every other
workload (multi-start screening, beta-only fits, profiles) shards with no communication at all.

One process per GPU, `torch.distributed` (NCCL over NVLink on the GPU box, gloo in the CPU tests)
is the plumbing.  The kernel's second stage writes the sums straight into the tensor that is
all-reduced in place.
"""
import numpy as np


def shard_bounds(n_total, world, rank):
    """Contiguous block [lo, hi) of the individuals owned by `rank`."""
    return rank * n_total // world, (rank + 1) * n_total // world


def finalize_sums(sums, n_total):
    """sums[S, P+1] (global, after the all-reduce) -> (loss[S], g_neural[S, P]).
    A start with a failed trajectory has loss = Inf and a zero gradient."""
    sums = np.asarray(sums, dtype=np.float64)
    loss = sums[:, 0] / n_total
    ok = np.isfinite(sums[:, 0])
    g = np.where(ok[:, None], sums[:, 1:] / n_total, 0.0)
    return loss, g


def allreduce_sums(sums, group=None):
    """In-place sum over ranks of a torch tensor (device tensor -> NCCL, CPU tensor -> gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def sharded_loss_grad(local_sums_fn, neural, cond_local, n_total, group=None):
    """Host-level sharded step used by the CPU (gloo) tests and by hosts that keep buffers on the CPU.

    local_sums_fn(neural[S,P], cond_local[S,n_loc]) -> (sums[S,P+1], g_cond_sse[S,n_loc]) with the
    *unscaled* local sums and per-trajectory d sse / d cond.  Returns (loss[S], g_neural[S,P],
    g_cond_local[S,n_loc]) of the global mean loss."""
    import torch
    sums, gc = local_sums_fn(neural, cond_local)
    t = torch.from_numpy(np.ascontiguousarray(sums, dtype=np.float64))
    allreduce_sums(t, group)
    loss, g = finalize_sums(t.numpy(), n_total)
    ok = np.isfinite(loss)
    gc = np.where(ok[:, None], np.asarray(gc) / n_total, 0.0)
    return loss, g, gc


class DevicePopulationShard:
    """GPU-resident shard: device tensors for the starts, one fused launch sequence per step
    (loss+gradient kernel -> deterministic block-partial reduction into `sums` -> NCCL all-reduce of
    `sums` in place on the same stream)."""

    def __init__(self, population, n_total, n_starts, device, group=None):
        import torch
        self.torch = torch
        self.pop, self.n_total, self.S, self.group = population, int(n_total), int(n_starts), group
        self.P = population.n_params
        self.n_loc = population.n_ind
        f64 = dict(dtype=torch.float64, device=device)
        self.neural = torch.empty((self.S, self.P), **f64)
        self.cond = torch.empty((self.S, self.n_loc), **f64)
        self.sums = torch.zeros((self.S, self.P + 1), **f64)
        self.g_cond = torch.empty((self.S, self.n_loc), **f64)

    def step(self, opts=None, want_grad=True):
        """Asynchronous: after it returns the stream holds kernel + reduction + all-reduce."""
        self.pop.eval_dev(self.S, self.neural.data_ptr(), self.P, self.cond.data_ptr(), 3 if want_grad else 0,
                          1.0 / self.n_total, 0, self.sums.data_ptr(), self.g_cond.data_ptr() if want_grad else 0, opts)
        allreduce_sums(self.sums, self.group)
        return self.sums

    def result(self):
        """(loss[S], g_neural[S,P]) on the host; synchronises."""
        return finalize_sums(self.sums.cpu().numpy(), self.n_total)
