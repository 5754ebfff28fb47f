"""Device-resident population + batched loss / gradient calls through the C ABI.

`Population` is the batched seam the reference does not have: the reference loops
`for (i, model) in enumerate(models)` on one CPU thread (src/parameter-estimation.jl:129-138,
:224-228, :362-366; src/likelihood-profiles.jl:11-14); here every (individual, start) pair is one
GPU thread.  All arithmetic happens in conditional_ude_b200/csrc (CUDA); this file only marshals.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from .models import pack_models, Chain


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class SolverOptions:
    """abstol / reltol / maxiters of the reference's `solve` call (OrdinaryDiffEq defaults)."""

    def __init__(self, abstol=1e-6, reltol=1e-3, maxiters=1000000, block=0, precision=0, balance=0, split=0):
        self.abstol, self.reltol, self.maxiters, self.block, self.precision = abstol, reltol, maxiters, block, precision
        self.balance = balance      # 0 automatic (<= 8192 trajectories: warp per trajectory; >= 32768 individuals: two-kernel gradient; else fused), 1 history, 2 two-kernel, 3 fused, 4 warp per trajectory
        self.split = split          # gradient pipeline: 0 automatic, 1 fused kernel, 2 split pipeline (cude_b200.h)

    def c(self):
        return _lib.cude_opts(self.abstol, self.reltol, int(self.maxiters), int(self.precision), int(self.block),
                              int(self.balance), int(self.split))


class Context:
    """One CUDA device + stream.  Raises if no device is present (no CPU fallback)."""

    def __init__(self, device=None, _borrowed=None):
        self._lib = _lib.load()
        if _borrowed is not None:            # a device context owned by a MultiContext
            self._h, self.device = _borrowed
            return
        if device is None:
            device = auto_device()
        h = C.c_void_p()
        _lib.check(self._lib.cude_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        self._fin = weakref.finalize(self, self._lib.cude_ctx_destroy, h)

    # ---- communicator of a sharded population (one process per GPU) ----
    def comm_init(self, nranks, rank, unique_id):
        """Collective: join the NCCL communicator described by `unique_id` (128 bytes from `comm_unique_id()` on rank 0,
        broadcast by the host).  Afterwards the sharded entry points all-reduce inside the library."""
        buf = (C.c_char * _lib.CUDE_UNIQUE_ID_BYTES).from_buffer_copy(bytes(unique_id))
        _lib.check(self._lib.cude_comm_init_rank(self._h, int(nranks), int(rank), C.cast(buf, C.c_void_p)), self._h)

    @property
    def comm_size(self):
        return self._lib.cude_comm_size(self._h)

    @property
    def comm_rank(self):
        return self._lib.cude_comm_rank(self._h)

    def allreduce_dev(self, d_buf, count):
        """In-place sum of a device buffer of doubles over the communicator, asynchronous on the context's stream."""
        _lib.check(self._lib.cude_allreduce_dev(self._h, C.c_void_p(d_buf), int(count)), self._h)

    @property
    def handle(self):
        return self._h

    def sync(self):
        _lib.check(self._lib.cude_sync(self._h), self._h)

    def stats(self):
        st = _lib.cude_stats()
        _lib.check(self._lib.cude_get_stats(self._h, C.byref(st)), self._h)
        return {k: getattr(st, k) for k, _ in st._fields_}

    def stream(self):
        return self._lib.cude_ctx_stream(self._h)

    def set_stream(self, cuda_stream):
        """Launch on a caller-owned cudaStream_t (int address), e.g. torch.cuda.current_stream().cuda_stream."""
        _lib.check(self._lib.cude_ctx_set_stream(self._h, C.c_void_p(cuda_stream) if cuda_stream else None), self._h)

    def math_probe(self, which, x):
        """Evaluate one of the kernels' elementary functions on the device (tests)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        _lib.check(self._lib.cude_math_probe(self._h, int(which), x.size, _dptr(x), _dptr(y)), self._h)
        return y

    def fp64_peak_tflops(self, register_operands=False):
        """Measured DFMA peak; register_operands=True uses three distinct register operands per DFMA (diagnostic)."""
        v = C.c_double()
        fn = self._lib.cude_measure_fp64_peak_rrr if register_operands else self._lib.cude_measure_fp64_peak
        _lib.check(fn(self._h, C.byref(v)), self._h)
        return v.value


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 calls this; the host broadcasts it to the other ranks)."""
    lib = _lib.load()
    buf = C.create_string_buffer(_lib.CUDE_UNIQUE_ID_BYTES)
    _lib.check(lib.cude_comm_get_unique_id(C.cast(buf, C.c_void_p)))
    return buf.raw


def pick_device(n_devices, local_rank=None, torch_device=None, one_visible=False):
    """Device a process should use when the caller names none: under a one-process-per-GPU launcher (torchrun exports
    LOCAL_RANK without narrowing CUDA_VISIBLE_DEVICES) rank r takes device r; a launcher that narrows
    CUDA_VISIBLE_DEVICES to one device per rank (`one_visible`) leaves device 0; a host that already selected a device
    through torch keeps it; otherwise device 0.  Raises when two local ranks would share a device."""
    if local_rank is not None and one_visible:
        return 0
    if local_rank is not None:
        if n_devices > 0 and local_rank >= n_devices:
            raise RuntimeError(f"LOCAL_RANK {local_rank} but only {n_devices} visible CUDA device(s): ranks would share a GPU; "
                               "pass Context(device) explicitly if that is intended")
        return int(local_rank)
    if torch_device is not None:
        return int(torch_device)
    return 0


def auto_device():
    import os
    import sys
    lr = os.environ.get("LOCAL_RANK")
    td = None
    if "torch" in sys.modules:           # never import torch from here: it is plumbing for hosts that use it
        torch = sys.modules["torch"]
        try:
            if torch.cuda.is_available() and torch.cuda.is_initialized():
                td = torch.cuda.current_device()
        except Exception:
            td = None
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    one = vis is not None and len([x for x in vis.split(",") if x.strip()]) == 1
    return pick_device(_lib.load().cude_device_count(), int(lr) if lr is not None and lr.strip().isdigit() else None, td, one)


def _fp32_peaks(self):
    """(FP32 FMA peak in TFLOP/s, MUFU ex2 rate in 1e9 ops/s) measured on this context's device."""
    a, b = C.c_double(), C.c_double()
    _lib.check(self._lib.cude_measure_fp32_peak(self._h, C.byref(a), C.byref(b)), self._h)
    return a.value, b.value


Context.fp32_peaks = _fp32_peaks

_default_ctx = {}


def default_context(device=None):
    """Process-wide context of a device; device=None follows `auto_device()` (LOCAL_RANK under torchrun)."""
    if device is None:
        device = auto_device()
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class Population:
    """Vector of CPeptideConditionalUDEModel + observations, uploaded once to HBM."""

    def __init__(self, models=None, timepoints=None, cpeptide_data=None, ctx=None, packed=None):
        self.ctx = ctx or default_context()
        self._lib = self.ctx._lib
        pk = packed if packed is not None else pack_models(models, timepoints, cpeptide_data)
        self.chain = pk["chain"]
        self.n_ind = int(pk["n_ind"])
        self.net = _lib.cude_net(self.chain.input_dims, self.chain.depth, self.chain.width)
        self.n_params = self.chain.n_params
        arrs = {k: np.ascontiguousarray(pk[k], dtype=np.float64) for k in ("knot_t", "knot_g", "obs_t", "obs_y", "kin")}
        nk = np.ascontiguousarray(pk["n_knots"], dtype=np.int32)
        no = np.ascontiguousarray(pk["n_obs"], dtype=np.int32)
        self.n_obs = no.copy()                       # observations per individual (sigma-likelihoods, SAEM)
        self.max_obs = int(pk["max_obs"])
        self.t_first = arrs["knot_t"][:, 0].copy()   # first glucose time point of every individual
        cov = None if pk.get("cov") is None else np.ascontiguousarray(pk["cov"], dtype=np.float64)
        h = C.c_void_p()
        _lib.check(self._lib.cude_population_create(
            self.ctx.handle, self.n_ind, int(pk["max_knots"]), _iptr(nk), _dptr(arrs["knot_t"]), _dptr(arrs["knot_g"]),
            int(pk["max_obs"]), _iptr(no), _dptr(arrs["obs_t"]), _dptr(arrs["obs_y"]), _dptr(arrs["kin"]), _dptr(cov),
            C.byref(h)), self.ctx.handle)
        self._h = h
        self._fin = weakref.finalize(self, self._lib.cude_population_destroy, h)

    @property
    def handle(self):
        return self._h

    # ---- argument marshalling -------------------------------------------------------------
    def _prep(self, neural, cond):
        neural = np.asarray(neural, dtype=np.float64)
        cond = np.asarray(cond, dtype=np.float64)
        P, N = self.n_params, self.n_ind
        if cond.ndim == 1:
            cond = cond.reshape(1, N) if cond.size == N else None
        if cond is None or cond.ndim != 2 or cond.shape[1] != N:
            raise ValueError(f"cond must be [n_starts x {N}] (row s = start s)")
        S = cond.shape[0]
        if neural.ndim == 1:
            if neural.size != P:
                raise ValueError(f"neural must have {P} entries")
            stride = 0
        else:
            if neural.shape != (S, P):
                raise ValueError(f"neural must be [{S} x {P}] or [{P}]")
            stride = P
        # row-major [S x N] / [S x P] == the ABI's column-major [N x S] / [P x S]
        return np.ascontiguousarray(neural), stride, np.ascontiguousarray(cond), S

    def loss(self, neural, cond, opts=None, return_sse=False):
        """Loss-only evaluation of S starts.  neural: [P] (shared) or [S x P]; cond: [S x N].
        Returns loss[S] = mean_i sse (Inf when a trajectory failed), optionally sse[S x N]."""
        neural, stride, cond, S = self._prep(neural, cond)
        o = (opts or SolverOptions()).c()
        loss = np.empty(S)
        sse = np.empty((S, self.n_ind)) if return_sse else None
        _lib.check(self._lib.cude_loss(self.ctx.handle, self._h, C.byref(self.net), C.byref(o), S, _dptr(neural), stride,
                                       _dptr(cond), _dptr(sse), _dptr(loss)), self.ctx.handle)
        return (loss, sse) if return_sse else loss

    def simulate(self, neural, cond, opts=None):
        """Model prediction at every individual's observation times (`solve(...; saveat=timepoints, save_idxs=1)`,
        src/parameter-estimation.jl:59): yhat[S x N x max_obs], NaN where unobserved or where the solve failed."""
        neural, stride, cond, S = self._prep(neural, cond)
        o = (opts or SolverOptions()).c()
        yhat = np.empty((S, self.n_ind, int(self.n_obs.max()) if self.max_obs is None else self.max_obs))
        _lib.check(self._lib.cude_simulate(self.ctx.handle, self._h, C.byref(self.net), C.byref(o), S, _dptr(neural), stride,
                                           _dptr(cond), _dptr(yhat), None), self.ctx.handle)
        return yhat

    def loss_grad(self, neural, cond, opts=None, neural_grad=True, mean=True, return_sse=False):
        """Loss and gradient.  Returns (loss[S], g_neural[S x P] or None, g_cond[S x N])
        (+ sse[S x N] when return_sse).  mean=False gives sums / per-trajectory derivatives."""
        neural, stride, cond, S = self._prep(neural, cond)
        o = (opts or SolverOptions()).c()
        loss = np.empty(S)
        gn = np.empty((S, self.n_params)) if neural_grad else None
        gc = np.empty((S, self.n_ind))
        sse = np.empty((S, self.n_ind)) if return_sse else None
        _lib.check(self._lib.cude_loss_grad(self.ctx.handle, self._h, C.byref(self.net), C.byref(o), S, _dptr(neural), stride,
                                            _dptr(cond), 1 if mean else 0, _dptr(sse), _dptr(loss), _dptr(gn), _dptr(gc)),
                   self.ctx.handle)
        return (loss, gn, gc, sse) if return_sse else (loss, gn, gc)

    def loss_grad_sums(self, neural, cond, cond_scale=1.0, opts=None, out_sums=None, out_g_cond=None):
        """Sharded-population form with host buffers (cude_loss_grad_sums): returns (sums[S x (P+1)], g_cond[S x N]) with
        sums[s] = {sum_i sse_i, sum_i d sse_i/d neural} of this shard, unscaled, and g_cond = cond_scale * d sse_i/d cond.
        `out_*` let the caller supply (page-locked) result buffers; large batches are pipelined inside the call."""
        neural, stride, cond, S = self._prep(neural, cond)
        o = (opts or SolverOptions()).c()
        sums = out_sums if out_sums is not None else np.empty((S, self.n_params + 1))
        gc = out_g_cond if out_g_cond is not None else np.empty((S, self.n_ind))
        assert sums.flags.c_contiguous and sums.shape == (S, self.n_params + 1) and sums.dtype == np.float64
        assert gc.flags.c_contiguous and gc.shape == (S, self.n_ind) and gc.dtype == np.float64
        _lib.check(self._lib.cude_loss_grad_sums(self.ctx.handle, self._h, C.byref(self.net), C.byref(o), S, _dptr(neural), stride,
                                                 _dptr(cond), float(cond_scale), _dptr(sums), _dptr(gc)), self.ctx.handle)
        return sums, gc

    def train_starts(self, neural, cond, adam_iters=1000, lr=1e-2, lbfgs_iters=1000, opts=None, **kw):
        """`_optimize` (src/parameter-estimation.jl:170-183) for S starts in lock-step on the device (cude_train): Adam, then
        L-BFGS with BackTracking; parameters, moments and history stay in HBM.  neural[S x P], cond[S x N] -> (neural,
        cond, objective[S], lbfgs_iterations[S], status[S], evaluations)."""
        neural, stride, cond, S = self._prep(neural, cond)
        if stride == 0:
            raise ValueError("train_starts needs one network per start")
        neural, cond = neural.copy(), cond.copy()
        o = (opts or SolverOptions()).c()
        t = _lib.cude_train_opts()
        self._lib.cude_train_default_opts(C.byref(t))
        t.adam_iters, t.adam_lr, t.lbfgs_iters = int(adam_iters), float(lr), int(lbfgs_iters)
        for k, v in kw.items():
            setattr(t, k, v)
        obj = np.empty(S)
        iters = np.zeros(S, dtype=np.int32)
        status = np.zeros(S, dtype=np.int32)
        ev = C.c_int(0)
        _lib.check(self._lib.cude_train(self.ctx.handle, self._h, C.byref(self.net), C.byref(o), C.byref(t), S, _dptr(neural),
                                        _dptr(cond), _dptr(obj), _iptr(iters), _iptr(status), C.byref(ev)), self.ctx.handle)
        return neural, cond, obj, iters, status, ev.value

    def loss_grad_sharded(self, neural, cond, n_total, opts=None, neural_grad=True, mean=True, loss_only=False,
                          out_loss=None, out_g_neural=None, out_g_cond=None):
        """Collective population loss / gradient when this Population holds one rank's block of `n_total` individuals
        (cude_loss_sharded / cude_loss_grad_sharded): the per-start sums are all-reduced inside the library over the
        context's communicator.  cond: this rank's [S x n_ind] block.  Returns (loss[S], g_neural[S x P] or None,
        g_cond[S x n_ind] or None) with loss and g_neural global.  `out_*`: caller-supplied (page-locked) result arrays."""
        neural, stride, cond, S = self._prep(neural, cond)
        o = (opts or SolverOptions()).c()
        loss = out_loss if out_loss is not None else np.empty(S)
        if loss_only:
            _lib.check(self._lib.cude_loss_sharded(self.ctx.handle, self._h, C.byref(self.net), C.byref(o), S, _dptr(neural), stride,
                                                   _dptr(cond), 0, int(n_total), None, _dptr(loss)), self.ctx.handle)
            return loss, None, None
        gn = out_g_neural if out_g_neural is not None else (np.empty((S, self.n_params)) if neural_grad else None)
        gc = out_g_cond if out_g_cond is not None else np.empty((S, self.n_ind))
        for a, shp in ((loss, (S,)), (gn, (S, self.n_params)), (gc, (S, self.n_ind))):
            assert a is None or (a.flags.c_contiguous and a.shape == shp and a.dtype == np.float64)
        _lib.check(self._lib.cude_loss_grad_sharded(self.ctx.handle, self._h, C.byref(self.net), C.byref(o), S, _dptr(neural), stride,
                                                    _dptr(cond), 0, int(n_total), 1 if mean else 0, None, _dptr(loss), _dptr(gn),
                                                    _dptr(gc)), self.ctx.handle)
        return loss, gn, gc

    def eval_dev(self, n_starts, d_neural, neural_stride, d_cond, want_grad, cond_scale, d_sse, d_sums, d_g_cond, opts=None):
        """Asynchronous device-pointer call (ints are raw device addresses)."""
        o = (opts or SolverOptions()).c()
        _lib.check(self._lib.cude_eval_dev(self.ctx.handle, self._h, C.byref(self.net), C.byref(o), int(n_starts),
                                           C.c_void_p(d_neural), int(neural_stride), C.c_void_p(d_cond), int(want_grad),
                                           float(cond_scale), C.c_void_p(d_sse or 0) if d_sse else None,
                                           C.c_void_p(d_sums) if d_sums else None,
                                           C.c_void_p(d_g_cond) if d_g_cond else None), self.ctx.handle)


class MultiContext:
    """All (or some) GPUs of the box driven from this one process: one context + stream + host worker thread per
    device inside the library (cude_mctx_create)."""

    def __init__(self, devices=None):
        self._lib = _lib.load()
        h = C.c_void_p()
        if devices is None or isinstance(devices, int):
            n, ids = int(devices or 0), None
        else:
            ids = np.ascontiguousarray(list(devices), dtype=np.int32)
            n = ids.size
        _lib.mcheck(self._lib.cude_mctx_create(n, _iptr(ids) if ids is not None else None, C.byref(h)))
        self._h = h
        self.n_gpus = self._lib.cude_mctx_size(h)
        self._fin = weakref.finalize(self, self._lib.cude_mctx_destroy, h)

    @property
    def handle(self):
        return self._h

    def device_context(self, k):
        """The k-th device's Context (borrowed: owned by this MultiContext)."""
        p = self._lib.cude_mctx_ctx(self._h, int(k))
        if not p:
            raise IndexError(k)
        c = Context(_borrowed=(C.c_void_p(p), int(k)))
        c._owner = self
        return c

    def stats(self):
        st = _lib.cude_stats()
        _lib.mcheck(self._lib.cude_mget_stats(self._h, C.byref(st)), self._h)
        return {k: getattr(st, k) for k, _ in st._fields_}


class MultiPopulation(Population):
    """Population on a MultiContext.  shard="starts": every device holds the whole population and a call's starts are
    split over the devices (screening, multi-start training, beta-only fits, profiles: no communication).
    shard="individuals": the individuals are split and every call ends with the NCCL all-reduce of the per-start sums
    (population training).  Same `loss` / `loss_grad` signatures and results as `Population`."""

    def __init__(self, models=None, timepoints=None, cpeptide_data=None, mctx=None, shard="starts", packed=None):
        self.mctx = mctx or MultiContext()
        self.ctx = self.mctx
        self._lib = self.mctx._lib
        pk = packed if packed is not None else pack_models(models, timepoints, cpeptide_data)
        self.chain = pk["chain"]
        self.n_ind = int(pk["n_ind"])
        self.net = _lib.cude_net(self.chain.input_dims, self.chain.depth, self.chain.width)
        self.n_params = self.chain.n_params
        arrs = {k: np.ascontiguousarray(pk[k], dtype=np.float64) for k in ("knot_t", "knot_g", "obs_t", "obs_y", "kin")}
        nk = np.ascontiguousarray(pk["n_knots"], dtype=np.int32)
        no = np.ascontiguousarray(pk["n_obs"], dtype=np.int32)
        self.n_obs = no.copy()
        self.max_obs = int(pk["max_obs"])
        self.t_first = arrs["knot_t"][:, 0].copy()
        cov = None if pk.get("cov") is None else np.ascontiguousarray(pk["cov"], dtype=np.float64)
        self.shard = shard
        mode = {"starts": _lib.CUDE_SHARD_STARTS, "individuals": _lib.CUDE_SHARD_INDIVIDUALS}[shard]
        h = C.c_void_p()
        _lib.mcheck(self._lib.cude_mpopulation_create(
            self.mctx.handle, mode, self.n_ind, int(pk["max_knots"]), _iptr(nk), _dptr(arrs["knot_t"]), _dptr(arrs["knot_g"]),
            int(pk["max_obs"]), _iptr(no), _dptr(arrs["obs_t"]), _dptr(arrs["obs_y"]), _dptr(arrs["kin"]), _dptr(cov),
            C.byref(h)), self.mctx.handle)
        self._h = h
        self._fin = weakref.finalize(self, self._lib.cude_mpopulation_destroy, h)

    def loss(self, neural, cond, opts=None, return_sse=False):
        neural, stride, cond, S = self._prep(neural, cond)
        o = (opts or SolverOptions()).c()
        loss = np.empty(S)
        sse = np.empty((S, self.n_ind)) if return_sse else None
        _lib.mcheck(self._lib.cude_mloss(self.mctx.handle, self._h, C.byref(self.net), C.byref(o), S, _dptr(neural), stride,
                                         _dptr(cond), _dptr(sse), _dptr(loss)), self.mctx.handle)
        return (loss, sse) if return_sse else loss

    def loss_grad(self, neural, cond, opts=None, neural_grad=True, mean=True, return_sse=False):
        neural, stride, cond, S = self._prep(neural, cond)
        o = (opts or SolverOptions()).c()
        loss = np.empty(S)
        gn = np.empty((S, self.n_params)) if neural_grad else None
        gc = np.empty((S, self.n_ind))
        sse = np.empty((S, self.n_ind)) if return_sse else None
        _lib.mcheck(self._lib.cude_mloss_grad(self.mctx.handle, self._h, C.byref(self.net), C.byref(o), S, _dptr(neural), stride,
                                              _dptr(cond), 1 if mean else 0, _dptr(sse), _dptr(loss), _dptr(gn), _dptr(gc)),
                    self.mctx.handle)
        return (loss, gn, gc, sse) if return_sse else (loss, gn, gc)

    def loss_grad_sums(self, *a, **k):
        raise NotImplementedError("device-level sums belong to one device: use Population on MultiContext.device_context(k)")

    eval_dev = simulate = loss_grad_sharded = train_starts = loss_grad_sums


_pop_cache = {}


def cached_population(models, timepoints, cpeptide_data, ctx=None):
    """Population for a (models, timepoints, data) triple, cached on object identity + data bytes so
    that reference-style per-call `loss(theta, (models, t, Y))` does not re-upload every call."""
    ts = np.ascontiguousarray(timepoints, dtype=np.float64)
    ys = np.ascontiguousarray(cpeptide_data, dtype=np.float64)
    key = (tuple(id(m) for m in models), ts.tobytes(), ys.tobytes(), id(ctx))
    pop = _pop_cache.get(key)
    if pop is None:
        if len(_pop_cache) > 64:
            _pop_cache.clear()
        pop = Population(list(models), ts, ys, ctx=ctx)
        _pop_cache[key] = (pop, list(models))   # keep the models alive so ids stay unique
        return pop
    return pop[0]
