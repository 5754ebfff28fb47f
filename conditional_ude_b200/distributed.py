"""Multi-GPU population training step: individuals sharded over ranks, one all-reduce.

The population loss of src/parameter-estimation.jl:126-140 is a mean over individuals; the
individuals (and their conditional parameters) are independent, so each rank owns a contiguous
block of them and the only exchange is the all-reduce of the per-start sums
{sum_i sse_i, sum_i d sse_i / d neural} — (P+1) x S doubles (19 KB for 64 starts).  Every other
workload (multi-start screening, beta-only fits, profiles) shards with no communication at all.

One process per GPU.  On the GPU the exchange happens INSIDE the library: `init_library_comm` hands the context an
NCCL communicator (the unique id travels over whatever the host already has — here torch.distributed, in Julia
`Distributed` — and that is torch's only role), after which `cude_allreduce_dev` / `cude_loss_grad_sharded` sum the
rows in place on the buffer and stream the reduction kernel used.  The host-level helpers below (`sharded_loss_grad`,
gloo) cover hosts that keep their buffers on the CPU and the CPU tests.
"""
import numpy as np


def init_library_comm(ctx, group=None):
    """Give `ctx` (a Context on this rank's GPU) the library-side NCCL communicator of the torch.distributed group:
    rank 0 draws the unique id (cude_comm_get_unique_id), the 128 bytes are broadcast through torch.distributed, every
    rank joins (cude_comm_init_rank).  Returns the communicator size; a no-op (1) outside a distributed run."""
    import torch
    import torch.distributed as dist
    from .population import comm_unique_id
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1
    if ctx.comm_size > 1:
        return ctx.comm_size
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    on_gpu = dist.get_backend(group) == "nccl"
    t = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        t = torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8).clone()
    if on_gpu:
        t = t.cuda(ctx.device)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init(world, rank, bytes(t.cpu().numpy().tobytes()))
    return world


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

    def __init__(self, population, n_total, n_starts, device, group=None, stream=None):
        import torch
        from .population import _default_ctx
        self.torch = torch
        self.pop, self.n_total, self.S, self.group = population, int(n_total), int(n_starts), group
        self.P = population.n_params
        self.n_loc = population.n_ind
        if any(population.ctx is c for c in _default_ctx.values()):
            raise ValueError("DevicePopulationShard re-points its context's stream: give the shard's Population its own "
                             "Context(device), not the process-wide default_context()")
        # one explicit stream for the kernels, the torch ops on the shard's tensors and the all-reduce; the stream object is
        # kept alive for as long as the context points at it and the context's own stream is restored by close()
        cur = torch.cuda.current_stream(device)
        self.stream = stream or (cur if cur.cuda_stream != 0 else torch.cuda.Stream(device=device))
        population.ctx.set_stream(self.stream.cuda_stream)
        import weakref
        self._restore = weakref.finalize(self, population.ctx.set_stream, 0)
        # the exchange runs inside the library (NCCL communicator on the context) when the job has more than one rank
        self.lib_comm = init_library_comm(population.ctx, group) > 1
        f64 = dict(dtype=torch.float64, device=device)
        with torch.cuda.stream(self.stream):
            self.neural = torch.empty((self.S, self.P), **f64)
            self.cond = torch.empty((self.S, self.n_loc), **f64)
            self.sums = torch.zeros((self.S, self.P + 1), **f64)
            self.g_cond = torch.empty((self.S, self.n_loc), **f64)

    def step(self, opts=None, want_grad=True):
        """Asynchronous: after it returns the stream holds kernel(s) + reduction + all-reduce.  Default options: the
        reference's tolerances, automatic lane balance (cude_opts.balance = 0)."""
        if opts is None:
            from .population import SolverOptions
            opts = SolverOptions()
        with self.torch.cuda.stream(self.stream):
            self.pop.eval_dev(self.S, self.neural.data_ptr(), self.P, self.cond.data_ptr(), 3 if want_grad else 0,
                              1.0 / self.n_total, 0, self.sums.data_ptr(), self.g_cond.data_ptr() if want_grad else 0, opts)
            if self.lib_comm:      # cude_allreduce_dev: ncclAllReduce in place, same stream as the kernels
                self.pop.ctx.allreduce_dev(self.sums.data_ptr(), self.sums.numel())
        return self.sums

    def close(self):
        """Hand the context back its own stream (the shard's stream may then be released)."""
        self._restore()

    # ---- device-resident Adam (population-scale training: the N x S conditional parameters stay in HBM) ----
    def adam_init(self):
        t = self.torch
        self._adam_t = 0
        with t.cuda.stream(self.stream):
            self._m_n, self._v_n = t.zeros_like(self.neural), t.zeros_like(self.neural)
            self._m_c, self._v_c = t.zeros_like(self.cond), t.zeros_like(self.cond)

    def adam_step(self, lr=1e-2, beta=(0.9, 0.999), eps=1e-8, opts=None):
        """One training iteration of the population loss (parameter-estimation.jl:126-140 + Optimisers.Adam), all
        starts in lock-step, no host synchronisation: loss+gradient kernel, partial reduction, all-reduce of the
        [S x (P+1)] sums, Adam on the (replicated) network weights and on this rank's conditional parameters."""
        from . import _lib
        import ctypes as C
        if not hasattr(self, "_adam_t"):
            self.adam_init()
        self.step(opts)                       # sums[:, 0] = sum sse, sums[:, 1:] = sum d sse/d neural (global after all-reduce)
        self._adam_t += 1
        lib, h = self.pop.ctx._lib, self.pop.ctx.handle
        p = lambda x: C.c_void_p(x.data_ptr())
        flag = C.c_void_p(self.sums.data_ptr())
        # network weights: gradient rows live inside `sums` (stride P+1) -> contiguous copy, scaled by 1/N_total
        with self.torch.cuda.stream(self.stream):
            g_n = self.sums[:, 1:].contiguous()
        _lib.check(lib.cude_adam_dev(h, g_n.numel(), p(self.neural), p(g_n), p(self._m_n), p(self._v_n), lr, beta[0], beta[1], eps,
                                     self._adam_t, 1.0 / self.n_total, flag, self.P, self.P + 1), h)
        # conditional parameters of this shard: g_cond already carries the 1/N_total factor
        _lib.check(lib.cude_adam_dev(h, self.g_cond.numel(), p(self.cond), p(self.g_cond), p(self._m_c), p(self._v_c), lr, beta[0],
                                     beta[1], eps, self._adam_t, 1.0, flag, self.n_loc, self.P + 1), h)

    def result(self):
        """(loss[S], g_neural[S,P]) on the host; synchronises."""
        with self.torch.cuda.stream(self.stream):
            h = self.sums.cpu()
        self.stream.synchronize()
        return finalize_sums(h.numpy(), self.n_total)
