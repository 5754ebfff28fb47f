"""ctypes wrapper for tests/emu/libcude_emu.so — the kernel source compiled for the host (test tool)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
# CUDE_EMU_SANITIZE=1: build with -fsanitize=address,undefined (run pytest with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
# ASAN_OPTIONS=detect_leaks=0) — the memory check of the kernel sources on this pool, where compute-sanitizer is closed
_SAN = os.environ.get("CUDE_EMU_SANITIZE") == "1"
LIB = os.path.join(_HERE, "libcude_emu_san.so" if _SAN else "libcude_emu.so")
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)


def build():
    src = [os.path.join(_HERE, "emu_kernel.cpp")] + [
        os.path.join(_HERE, "..", "..", "conditional_ude_b200", "csrc", f)
        for f in ("cude_kernels.cuh", "cude_math.cuh", "cude_sup_kernel.cuh", "cude_split.cuh")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas"] +
                              (["-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined"] if _SAN else []) +
                              ["-o", LIB, src[0]])
    return LIB


def _dp(a):
    return a.ctypes.data_as(_D) if a is not None else None


def emu_eval(packed, neural, cond, abstol=1e-6, reltol=1e-3, maxiters=100000, grad=True, flat=False, mixed=False):
    L = C.CDLL(build())
    L.emu_eval.argtypes = [C.c_int, C.c_int, _I, _D, _D, C.c_int, _I, _D, _D, _D, _D, C.c_int, C.c_int, _D, C.c_longlong, _D,
                           C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _D, _D, _D, C.POINTER(C.c_ulonglong), C.c_int]
    ch = packed["chain"]
    N, P = int(packed["n_ind"]), ch.n_params
    a = {k: np.ascontiguousarray(packed[k], dtype=np.float64) for k in ("knot_t", "knot_g", "obs_t", "obs_y", "kin")}
    nk = np.ascontiguousarray(packed["n_knots"], dtype=np.int32)
    no = np.ascontiguousarray(packed["n_obs"], dtype=np.int32)
    cov = None if packed.get("cov") is None else np.ascontiguousarray(packed["cov"], dtype=np.float64)
    neural = np.ascontiguousarray(neural, dtype=np.float64)
    cond = np.ascontiguousarray(np.asarray(cond, dtype=np.float64).reshape(-1, N))
    S = cond.shape[0]
    stride = 0 if neural.ndim == 1 else P
    sse = np.empty((S, N))
    gn = np.zeros((S, N, P))
    gc = np.zeros((S, N))
    cnt = (C.c_ulonglong * 3)()
    rc = L.emu_eval(N, int(packed["max_knots"]), nk.ctypes.data_as(_I), _dp(a["knot_t"]), _dp(a["knot_g"]),
                    int(packed["max_obs"]), no.ctypes.data_as(_I), _dp(a["obs_t"]), _dp(a["obs_y"]), _dp(a["kin"]), _dp(cov),
                    ch.input_dims, S, _dp(neural), stride, _dp(cond), abstol, reltol, maxiters, int(grad), int(flat),
                    _dp(sse), _dp(gn), _dp(gc), cnt, int(mixed))
    assert rc == 0
    return dict(sse=sse, g_neural=gn, g_cond=gc, n_acc=cnt[0], n_rej=cnt[1], n_fail=cnt[2])


def emu_eval_split(packed, neural, cond, abstol=1e-6, reltol=1e-3, maxiters=100000):
    """The split gradient pipeline (csrc/cude_split.cuh) through the host-compiled kernel sources, stage by stage.
    Returns dict(sse[S,N], sums[S,P+1] = {sum sse, d sum sse/d neural}, g_cond[S,N], n_overflow)."""
    L = C.CDLL(build())
    L.emu_eval_split.argtypes = [C.c_int, C.c_int, _I, _D, _D, C.c_int, _I, _D, _D, _D, C.c_int, _D, _D, C.c_double, C.c_double,
                                 C.c_int, _D, _D, _D, _I]
    ch = packed["chain"]
    N, P = int(packed["n_ind"]), ch.n_params
    a = {k: np.ascontiguousarray(packed[k], dtype=np.float64) for k in ("knot_t", "knot_g", "obs_t", "obs_y", "kin")}
    nk = np.ascontiguousarray(packed["n_knots"], dtype=np.int32)
    no = np.ascontiguousarray(packed["n_obs"], dtype=np.int32)
    neural = np.ascontiguousarray(neural, dtype=np.float64)
    cond = np.ascontiguousarray(np.asarray(cond, dtype=np.float64).reshape(-1, N))
    S = cond.shape[0]
    assert neural.shape == (S, P) and ch.input_dims == 2
    sse, sums, gc = np.empty((S, N)), np.zeros((S, P + 1)), np.zeros((S, N))
    novf = C.c_int(0)
    rc = L.emu_eval_split(N, int(packed["max_knots"]), nk.ctypes.data_as(_I), _dp(a["knot_t"]), _dp(a["knot_g"]),
                          int(packed["max_obs"]), no.ctypes.data_as(_I), _dp(a["obs_t"]), _dp(a["obs_y"]), _dp(a["kin"]),
                          S, _dp(neural), _dp(cond), abstol, reltol, maxiters, _dp(sse), _dp(sums), _dp(gc), C.byref(novf))
    assert rc == 0
    return dict(sse=sse, sums=sums, g_cond=gc, n_overflow=novf.value)


def emu_eval_exact(packed, neural, cond, abstol=1e-6, reltol=1e-3, maxiters=100000):
    """The two-kernel gradient (forward kernel with step records -> per-start sort by accepted steps -> cude_adjoint_kernel in
    sorted order) through the host-compiled kernel sources.  Same return value as emu_eval_split."""
    L = C.CDLL(build())
    L.emu_eval_exact.argtypes = [C.c_int, C.c_int, _I, _D, _D, C.c_int, _I, _D, _D, _D, C.c_int, _D, _D, C.c_double, C.c_double,
                                 C.c_int, _D, _D, _D, _I]
    ch = packed["chain"]
    N, P = int(packed["n_ind"]), ch.n_params
    a = {k: np.ascontiguousarray(packed[k], dtype=np.float64) for k in ("knot_t", "knot_g", "obs_t", "obs_y", "kin")}
    nk = np.ascontiguousarray(packed["n_knots"], dtype=np.int32)
    no = np.ascontiguousarray(packed["n_obs"], dtype=np.int32)
    neural = np.ascontiguousarray(neural, dtype=np.float64)
    cond = np.ascontiguousarray(np.asarray(cond, dtype=np.float64).reshape(-1, N))
    S = cond.shape[0]
    assert neural.shape == (S, P) and ch.input_dims == 2
    sse, sums, gc = np.empty((S, N)), np.zeros((S, P + 1)), np.zeros((S, N))
    novf = C.c_int(0)
    rc = L.emu_eval_exact(N, int(packed["max_knots"]), nk.ctypes.data_as(_I), _dp(a["knot_t"]), _dp(a["knot_g"]),
                          int(packed["max_obs"]), no.ctypes.data_as(_I), _dp(a["obs_t"]), _dp(a["obs_y"]), _dp(a["kin"]),
                          S, _dp(neural), _dp(cond), abstol, reltol, maxiters, _dp(sse), _dp(sums), _dp(gc), C.byref(novf))
    assert rc == 0, rc
    return dict(sse=sse, sums=sums, g_cond=gc, n_overflow=novf.value)


LIB_WARP = os.path.join(_HERE, "libcude_emu_warp_san.so" if _SAN else "libcude_emu_warp.so")


def build_warp():
    src = [os.path.join(_HERE, "emu_warp.cpp")] + [
        os.path.join(_HERE, "..", "..", "conditional_ude_b200", "csrc", f)
        for f in ("cude_kernels.cuh", "cude_math.cuh", "cude_split.cuh", "cude_warp.cuh")]
    if not os.path.exists(LIB_WARP) or os.path.getmtime(LIB_WARP) < max(os.path.getmtime(s) for s in src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas"] +
                              (["-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined"] if _SAN else []) +
                              ["-o", LIB_WARP, src[0]])
    return LIB_WARP


def emu_warp_eval(packed, neural, cond, abstol=1e-6, reltol=1e-3, maxiters=100000, grad=True, lanes=32):
    """The warp-per-trajectory kernel (csrc/cude_warp.cuh) with every CUDA thread as a host thread (tests/emu/emu_warp.cpp).
    Returns dict(sse[S,N], g_neural[S,N,P], g_cond[S,N], overflow[S,N] bool, n_acc, n_rej, n_fail)."""
    L = C.CDLL(build_warp())
    L.emu_warp_eval.argtypes = [C.c_int, C.c_int, _I, _D, _D, C.c_int, _I, _D, _D, _D, _D, C.c_int, C.c_int, _D, _D,
                                C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _D, _D, _D, _I, C.POINTER(C.c_ulonglong)]
    ch = packed["chain"]
    N, P = int(packed["n_ind"]), ch.n_params
    a = {k: np.ascontiguousarray(packed[k], dtype=np.float64) for k in ("knot_t", "knot_g", "obs_t", "obs_y", "kin")}
    nk = np.ascontiguousarray(packed["n_knots"], dtype=np.int32)
    no = np.ascontiguousarray(packed["n_obs"], dtype=np.int32)
    cov = None if packed.get("cov") is None else np.ascontiguousarray(packed["cov"], dtype=np.float64)
    neural = np.ascontiguousarray(neural, dtype=np.float64)
    cond = np.ascontiguousarray(np.asarray(cond, dtype=np.float64).reshape(-1, N))
    S = cond.shape[0]
    assert neural.shape == (S, P)
    sse, rows, gc = np.empty((S, N)), np.zeros((S, N, P + 1)), np.zeros((S, N))
    ovf = np.zeros((S, N), dtype=np.int32)
    cnt = (C.c_ulonglong * 3)()
    rc = L.emu_warp_eval(N, int(packed["max_knots"]), nk.ctypes.data_as(_I), _dp(a["knot_t"]), _dp(a["knot_g"]),
                         int(packed["max_obs"]), no.ctypes.data_as(_I), _dp(a["obs_t"]), _dp(a["obs_y"]), _dp(a["kin"]), _dp(cov),
                         ch.input_dims, S, _dp(neural), _dp(cond), abstol, reltol, maxiters, int(grad), int(lanes), _dp(sse), _dp(rows), _dp(gc),
                         ovf.ctypes.data_as(_I), cnt)
    assert rc == 0, rc
    return dict(sse=sse, row_sse=rows[:, :, 0], g_neural=rows[:, :, 1:], g_cond=gc, overflow=ovf < 0, n_acc=cnt[0], n_rej=cnt[1], n_fail=cnt[2])


LIB_TRAIN = os.path.join(_HERE, "libcude_emu_train_san.so" if _SAN else "libcude_emu_train.so")


def build_train():
    src = [os.path.join(_HERE, "emu_train.cpp"), os.path.join(_HERE, "emu_threads.h"),
           os.path.join(_HERE, "..", "..", "conditional_ude_b200", "csrc", "cude_train.cuh")]
    if not os.path.exists(LIB_TRAIN) or os.path.getmtime(LIB_TRAIN) < max(os.path.getmtime(s) for s in src):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-DCUDE_TRAIN_T=64"] +
                              (["-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fno-sanitize-recover=undefined"] if _SAN else []) +
                              ["-o", LIB_TRAIN, src[0]])
    return LIB_TRAIN


def emu_train(fg, x0, n_neural, adam_iters=0, adam_lr=1e-2, lbfgs_iters=0, lbfgs_m=10, g_tol=1e-8, c1=1e-4, rho_hi=0.5, rho_lo=0.1,
              ls_maxiter=50, check_every=16):
    """The device-resident optimisers of cude_train (csrc/cude_train.cuh) with every CUDA thread as a host thread and the
    library's host loop restated in tests/emu/emu_train.cpp.  fg(x[S, D]) -> (f[S], g[S, D]); the first n_neural columns play
    the network part.  Returns (x[S, D], objective[S], lbfgs_iterations[S], status[S], evaluations)."""
    L = C.CDLL(build_train())
    x = np.array(x0, dtype=np.float64, order="C")
    S, D = x.shape
    CB = C.CFUNCTYPE(None, _D, _D, _D)

    def cb(px, pf, pg):
        f, g = fg(np.ctypeslib.as_array(px, shape=(S, D)).copy())
        np.ctypeslib.as_array(pf, shape=(S,))[:] = f
        np.ctypeslib.as_array(pg, shape=(S, D))[:] = g

    L.emu_train.argtypes = [C.c_int, C.c_int, C.c_int, CB, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                            C.c_double, C.c_int, C.c_int, _D, _D, _I, _I, _I]
    obj = np.empty(S)
    iters = np.zeros(S, dtype=np.int32)
    status = np.zeros(S, dtype=np.int32)
    ev = C.c_int(0)
    rc = L.emu_train(S, int(n_neural), D - int(n_neural), CB(cb), int(adam_iters), float(adam_lr), int(lbfgs_iters), int(lbfgs_m),
                     g_tol, c1, rho_hi, rho_lo, int(ls_maxiter), int(check_every), _dp(x), _dp(obj), iters.ctypes.data_as(_I),
                     status.ctypes.data_as(_I), C.byref(ev))
    assert rc == 0
    return x, obj, iters, status, ev.value


def emu_sup_eval(data, timepoints, neural, theta, p_true=(0.4, 0.9, 0.3), scale=None, abstol=1e-6, reltol=1e-3,
                 maxiters=100000, grad=True):
    """Suppression variant through the host-compiled kernel source; same conventions as oracle.sup_eval."""
    L = C.CDLL(build())
    L.emu_sup_eval.argtypes = [C.c_int, C.c_int, _D, _D, _D, _D, C.c_double, C.c_double, C.c_int, _D, C.c_longlong, _D,
                               C.c_double, C.c_double, C.c_int, C.c_int, _D, _D, _D, C.POINTER(C.c_ulonglong)]
    data = np.asarray(data, dtype=np.float64)
    _, n_obs, n_ind = data.shape
    dj = np.ascontiguousarray(data.transpose(2, 1, 0))
    ot = np.ascontiguousarray(timepoints, dtype=np.float64)
    sc = np.ascontiguousarray(data.max(axis=1).mean(axis=1) if scale is None else scale, dtype=np.float64)
    pt = np.ascontiguousarray(p_true, dtype=np.float64)
    neural = np.ascontiguousarray(neural, dtype=np.float64)
    theta = np.ascontiguousarray(np.asarray(theta, dtype=np.float64).reshape(-1, n_ind))
    S, P = theta.shape[0], 67
    stride = 0 if neural.ndim == 1 else P
    sse = np.empty((S, n_ind)); gn = np.zeros((S, n_ind, P)); gt = np.zeros((S, n_ind))
    cnt = (C.c_ulonglong * 3)()
    rc = L.emu_sup_eval(n_ind, n_obs, _dp(ot), _dp(dj), _dp(pt), _dp(sc), float(ot[0]), float(ot[-1]), S, _dp(neural), stride,
                        _dp(theta), abstol, reltol, maxiters, int(grad), _dp(sse), _dp(gn), _dp(gt), cnt)
    assert rc == 0
    return dict(sse=sse, g_neural=gn, g_theta=gt, n_acc=cnt[0], n_rej=cnt[1], n_fail=cnt[2] & 0xffffffff, n_overflow=cnt[2] >> 32)
