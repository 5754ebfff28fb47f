"""ctypes wrapper of oracle/libcude_oracle.so (the CPU restatement in cude_oracle.cpp).

TEST INFRASTRUCTURE ONLY: the checker for the CUDA path, and the timed CPU baseline.  The product
package (conditional_ude_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcude_oracle.so")
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)


class _Pop(C.Structure):
    _fields_ = [("n_ind", C.c_int), ("max_knots", C.c_int), ("max_obs", C.c_int),
                ("n_knots", _I), ("knot_t", _D), ("knot_g", _D),
                ("n_obs", _I), ("obs_t", _D), ("obs_y", _D), ("kin", _D), ("cov", _D)]


def build(force=False):
    src = os.path.join(_HERE, "cude_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libcude_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.cude_oracle_van_cauter.argtypes = [C.c_double, C.c_int, _D, _D, _D]
        L.cude_oracle_nparams.argtypes = [C.c_int] * 3
        L.cude_oracle_mlp.restype = C.c_double
        L.cude_oracle_mlp.argtypes = [C.c_int, C.c_int, C.c_int, _D, _D]
        L.cude_oracle_glucose.restype = C.c_double
        L.cude_oracle_glucose.argtypes = [C.c_int, _D, _D, C.c_double]
        L.cude_oracle_eval.argtypes = [C.POINTER(_Pop), C.c_int, C.c_int, C.c_int, C.c_int, _D, C.c_long, _D,
                                       C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _D, _D, _I, _D, _D]
        L.cude_oracle_population_loss.argtypes = [C.POINTER(_Pop), C.c_int, C.c_int, C.c_int, C.c_int, _D, C.c_long, _D,
                                                  C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _D, _D, _D,
                                                  C.POINTER(C.c_long)]
        L.cude_oracle_sup_eval.argtypes = [C.c_int, C.c_int, _D, _D, _D, _D, C.c_double, C.c_double, C.c_int, C.c_int,
                                           C.c_int, _D, C.c_long, _D, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int,
                                           _D, _D, _I, _D, _D]
        L.cude_oracle_eval_generic.argtypes = [C.POINTER(_Pop), C.c_int, C.c_int, C.c_int, C.c_int, _D, C.c_long, _D,
                                               C.c_double, C.c_double, C.c_int, _D, _I]
        L.cude_oracle_trace.argtypes = [C.POINTER(_Pop), C.c_int, C.c_int, C.c_int, C.c_int, _D, C.c_double, C.c_double,
                                        C.c_double, C.c_int, _D, C.c_int, _D]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(_D) if a is not None else None


class OraclePopulation:
    """Holds the row-major arrays produced by conditional_ude_b200.models.pack_models."""

    def __init__(self, packed):
        self.n_ind = int(packed["n_ind"])
        self.max_knots, self.max_obs = int(packed["max_knots"]), int(packed["max_obs"])
        ch = packed["chain"]
        self.n_in, self.depth, self.width = ch.input_dims, ch.depth, ch.width
        self.P = ch.n_params
        self._a = {k: np.ascontiguousarray(packed[k], dtype=np.float64) for k in ("knot_t", "knot_g", "obs_t", "obs_y", "kin")}
        self._nk = np.ascontiguousarray(packed["n_knots"], dtype=np.int32)
        self._no = np.ascontiguousarray(packed["n_obs"], dtype=np.int32)
        self._cov = None if packed.get("cov") is None else np.ascontiguousarray(packed["cov"], dtype=np.float64)
        self._c = _Pop(self.n_ind, self.max_knots, self.max_obs, self._nk.ctypes.data_as(_I), _dp(self._a["knot_t"]),
                       _dp(self._a["knot_g"]), self._no.ctypes.data_as(_I), _dp(self._a["obs_t"]), _dp(self._a["obs_y"]),
                       _dp(self._a["kin"]), _dp(self._cov))

    def _prep(self, neural, cond):
        neural = np.ascontiguousarray(neural, dtype=np.float64)
        cond = np.ascontiguousarray(np.asarray(cond, dtype=np.float64).reshape(-1, self.n_ind))
        S = cond.shape[0]
        stride = 0 if neural.ndim == 1 else self.P
        if neural.ndim == 2:
            assert neural.shape == (S, self.P)
        return neural, stride, cond, S

    def eval(self, neural, cond, abstol=1e-6, reltol=1e-3, maxiters=100000, grad_mode=-1, n_threads=0, want_yhat=False):
        """Per-trajectory evaluation.  Returns dict(sse[S,N], stats[S,N,4], g_neural[S,N,P], g_cond[S,N], yhat)."""
        neural, stride, cond, S = self._prep(neural, cond)
        N, P = self.n_ind, self.P
        sse = np.empty((S, N))
        stats = np.empty((S, N, 4), dtype=np.int32)
        yhat = np.full((S, N, self.max_obs), np.nan) if want_yhat else None
        gn = np.zeros((S, N, P)) if grad_mode >= 0 else None
        gc = np.zeros((S, N)) if grad_mode >= 0 else None
        rc = lib().cude_oracle_eval(C.byref(self._c), self.n_in, self.depth, self.width, S, _dp(neural), stride, _dp(cond),
                                    abstol, reltol, maxiters, grad_mode, n_threads, _dp(sse), _dp(yhat),
                                    stats.ctypes.data_as(_I), _dp(gn), _dp(gc))
        assert rc == 0
        return dict(sse=sse, stats=stats, g_neural=gn, g_cond=gc, yhat=yhat)

    def eval_generic(self, neural, cond, abstol=1e-6, reltol=1e-3, maxiters=100000):
        """The same loss through the generic D-state Tsit5 core (the one pinned by the suppression artifacts)."""
        neural, stride, cond, S = self._prep(neural, cond)
        sse = np.empty((S, self.n_ind))
        stats = np.empty((S, self.n_ind, 4), dtype=np.int32)
        rc = lib().cude_oracle_eval_generic(C.byref(self._c), self.n_in, self.depth, self.width, S, _dp(neural), stride,
                                            _dp(cond), abstol, reltol, maxiters, _dp(sse), stats.ctypes.data_as(_I))
        assert rc == 0
        return dict(sse=sse, stats=stats)

    def trace(self, i, neural, cond, abstol=1e-6, reltol=1e-3, maxiters=100000, cap=100000):
        """Step trace of individual i: array of rows (t, dt, EEst, accepted) and the sse."""
        neural = np.ascontiguousarray(neural, dtype=np.float64)
        rows = np.zeros((cap, 4))
        sse = C.c_double()
        n = lib().cude_oracle_trace(C.byref(self._c), self.n_in, self.depth, self.width, int(i), _dp(neural), float(cond),
                                    abstol, reltol, maxiters, _dp(rows), cap, C.byref(sse))
        return rows[:n], sse.value

    def population_loss(self, neural, cond, abstol=1e-6, reltol=1e-3, maxiters=100000, with_grad=False, n_threads=0):
        """loss[S] (mean over individuals, parameter-estimation.jl:126-140) (+ gradients) and step counters."""
        neural, stride, cond, S = self._prep(neural, cond)
        N, P = self.n_ind, self.P
        loss = np.empty(S)
        gn = np.zeros((S, P)) if with_grad else None
        gc = np.zeros((S, N)) if with_grad else None
        cnt = (C.c_long * 3)()
        rc = lib().cude_oracle_population_loss(C.byref(self._c), self.n_in, self.depth, self.width, S, _dp(neural), stride,
                                               _dp(cond), abstol, reltol, maxiters, int(with_grad), n_threads, _dp(loss),
                                               _dp(gn), _dp(gc), cnt)
        assert rc == 0
        return dict(loss=loss, g_neural=gn, g_cond=gc, n_acc=cnt[0], n_rej=cnt[1], n_rhs=cnt[2])


def van_cauter(age, t2dm):
    k = [C.c_double() for _ in range(3)]
    lib().cude_oracle_van_cauter(float(age), int(bool(t2dm)), *[C.byref(x) for x in k])
    return tuple(x.value for x in k)


def mlp(n_in, depth, width, p, x):
    p = np.ascontiguousarray(p, dtype=np.float64)
    x = np.ascontiguousarray(x, dtype=np.float64)
    return lib().cude_oracle_mlp(n_in, depth, width, _dp(p), _dp(x))


def glucose(kt, kg, tau):
    kt = np.ascontiguousarray(kt, dtype=np.float64)
    kg = np.ascontiguousarray(kg, dtype=np.float64)
    return lib().cude_oracle_glucose(kt.size, _dp(kt), _dp(kg), float(tau))


def max_threads():
    return lib().cude_oracle_max_threads()


def suppression_scale(data):
    """scale = mean(maximum(individual_data, dims=2), dims=3)[:]  (suppression_model.jl:125); data is [3, n_obs, n_ind]."""
    return np.asarray(data, dtype=np.float64).max(axis=1).mean(axis=1)


def sup_eval(data, timepoints, neural, theta, p_true=(0.4, 0.9, 0.3), depth=5, width=3, tspan=None, scale=None,
             abstol=1e-6, reltol=1e-3, maxiters=100000, with_grad=False, n_threads=0, want_yhat=False):
    """Suppression example, per-trajectory scaled SSE (+ gradient): data [3, n_obs, n_ind] (Julia layout),
    neural [P] or [S, P], theta [S, n_ind].  Returns dict(sse[S,N], stats, g_neural[S,N,P], g_theta[S,N], yhat[S,N,n_obs,3])."""
    data = np.asarray(data, dtype=np.float64)
    _, n_obs, n_ind = data.shape
    dj = np.ascontiguousarray(data.transpose(2, 1, 0))               # [i][k][state] == Julia column-major 3 x n_obs x n_ind
    ot = np.ascontiguousarray(timepoints, dtype=np.float64)
    sc = np.ascontiguousarray(suppression_scale(data) if scale is None else scale, dtype=np.float64)
    pt = np.ascontiguousarray(p_true, dtype=np.float64)
    t0, tend = (float(ot[0]), float(ot[-1])) if tspan is None else tspan
    neural = np.ascontiguousarray(neural, dtype=np.float64)
    theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(-1, n_ind))
    S = theta.shape[0]
    P = lib().cude_oracle_nparams(4, depth, width)
    stride = 0 if neural.ndim == 1 else P
    sse = np.empty((S, n_ind))
    stats = np.empty((S, n_ind, 4), dtype=np.int32)
    yhat = np.full((S, n_ind, n_obs, 3), np.nan) if want_yhat else None
    gn = np.zeros((S, n_ind, P)) if with_grad else None
    gt = np.zeros((S, n_ind)) if with_grad else None
    rc = lib().cude_oracle_sup_eval(n_ind, n_obs, _dp(ot), _dp(dj), _dp(pt), _dp(sc), t0, tend, depth, width, S, _dp(neural),
                                    stride, _dp(theta), abstol, reltol, maxiters, int(with_grad), n_threads, _dp(sse), _dp(yhat),
                                    stats.ctypes.data_as(_I), _dp(gn), _dp(gt))
    assert rc == 0
    return dict(sse=sse, stats=stats, g_neural=gn, g_theta=gt, yhat=yhat)
