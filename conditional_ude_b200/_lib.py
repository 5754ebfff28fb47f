"""ctypes binding of libcude_b200.so (C ABI declared in include/cude_b200.h).

There is no CPU fallback: if the CUDA library is missing, or no CUDA device is usable, every
compute entry point raises.  The library is built in-tree by `__graft_entry__.build()` or
`make -C conditional_ude_b200/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CUDE_B200_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = os.environ.get("CUDE_B200_LIB") or os.path.join(_HERE, "csrc", "libcude_b200.so")

CUDE_OK = 0
CUDE_EINVAL, CUDE_ENODEVICE, CUDE_ECUDA, CUDE_ENOMEM, CUDE_EUNSUPPORTED, CUDE_ENCCL = -1, -2, -3, -4, -5, -6
CUDE_SHARD_STARTS, CUDE_SHARD_INDIVIDUALS = 0, 1
CUDE_UNIQUE_ID_BYTES = 128
ABI_VERSION = 3


class CudeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cude_b200 error {code}: {msg}")
        self.code = code


class cude_net(C.Structure):
    _fields_ = [("n_in", C.c_int), ("depth", C.c_int), ("width", C.c_int)]


class cude_opts(C.Structure):
    _fields_ = [("abstol", C.c_double), ("reltol", C.c_double), ("maxiters", C.c_int),
                ("precision", C.c_int), ("block", C.c_int), ("balance", C.c_int), ("split", C.c_int)]


class cude_train_opts(C.Structure):
    _fields_ = [("adam_iters", C.c_int), ("adam_lr", C.c_double), ("adam_beta1", C.c_double), ("adam_beta2", C.c_double),
                ("adam_eps", C.c_double), ("lbfgs_iters", C.c_int), ("lbfgs_m", C.c_int), ("g_tol", C.c_double), ("c1", C.c_double),
                ("rho_hi", C.c_double), ("rho_lo", C.c_double), ("ls_maxiter", C.c_int), ("check_every", C.c_int)]


class cude_stats(C.Structure):
    _fields_ = [("n_traj", C.c_ulonglong), ("n_acc", C.c_ulonglong), ("n_rej", C.c_ulonglong),
                ("n_rhs", C.c_ulonglong), ("n_fail", C.c_ulonglong), ("kernel_ms", C.c_float),
                ("launches", C.c_int)]


# every symbol include/cude_b200.h declares: (restype, argtypes)
_P = C.c_void_p
_D = C.POINTER(C.c_double)
_I = C.POINTER(C.c_int)
SYMBOLS = {
    "cude_abi_version": (C.c_int, []),
    "cude_default_opts": (None, [C.POINTER(cude_opts)]),
    "cude_net_nparams": (C.c_int, [C.POINTER(cude_net)]),
    "cude_van_cauter_parameters": (None, [C.c_double, C.c_int, _D, _D, _D]),
    "cude_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "cude_ctx_destroy": (C.c_int, [_P]),
    "cude_last_error": (C.c_char_p, [_P]),
    "cude_sync": (C.c_int, [_P]),
    "cude_get_stats": (C.c_int, [_P, C.POINTER(cude_stats)]),
    "cude_ctx_stream": (_P, [_P]),
    "cude_ctx_set_stream": (C.c_int, [_P, _P]),
    "cude_population_create": (C.c_int, [_P, C.c_int, C.c_int, _I, _D, _D, C.c_int, _I, _D, _D, _D, _D, C.POINTER(_P)]),
    "cude_population_destroy": (C.c_int, [_P]),
    "cude_population_size": (C.c_int, [_P]),
    "cude_loss": (C.c_int, [_P, _P, C.POINTER(cude_net), C.POINTER(cude_opts), C.c_int, _D, C.c_longlong, _D, _D, _D]),
    "cude_simulate": (C.c_int, [_P, _P, C.POINTER(cude_net), C.POINTER(cude_opts), C.c_int, _D, C.c_longlong, _D, _D, _D]),
    "cude_loss_grad": (C.c_int, [_P, _P, C.POINTER(cude_net), C.POINTER(cude_opts), C.c_int, _D, C.c_longlong, _D,
                                 C.c_int, _D, _D, _D, _D]),
    "cude_loss_grad_sums": (C.c_int, [_P, _P, C.POINTER(cude_net), C.POINTER(cude_opts), C.c_int, _D, C.c_longlong, _D,
                                      C.c_double, _D, _D]),
    "cude_eval_dev": (C.c_int, [_P, _P, C.POINTER(cude_net), C.POINTER(cude_opts), C.c_int, _P, C.c_longlong, _P,
                                C.c_int, C.c_double, _P, _P, _P]),
    "cude_measure_fp64_peak": (C.c_int, [_P, _D]),
    "cude_measure_fp64_peak_rrr": (C.c_int, [_P, _D]),
    "cude_measure_fp32_peak": (C.c_int, [_P, _D, _D]),
    "cude_math_probe": (C.c_int, [_P, C.c_int, C.c_int, _D, _D]),
    "cude_train_default_opts": (None, [C.POINTER(cude_train_opts)]),
    "cude_train": (C.c_int, [_P, _P, C.POINTER(cude_net), C.POINTER(cude_opts), C.POINTER(cude_train_opts), C.c_int, _D, _D, _D, _I, _I, _I]),
    "cude_adam_dev": (C.c_int, [_P, C.c_longlong, _P, _P, _P, _P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                C.c_double, _P, C.c_longlong, C.c_longlong]),
    # multi-GPU: one process per GPU (communicator on a context) ...
    "cude_device_count": (C.c_int, []),
    "cude_nccl_version": (C.c_int, []),
    "cude_comm_get_unique_id": (C.c_int, [_P]),
    "cude_comm_init_rank": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "cude_comm_destroy": (C.c_int, [_P]),
    "cude_comm_size": (C.c_int, [_P]),
    "cude_comm_rank": (C.c_int, [_P]),
    "cude_allreduce_dev": (C.c_int, [_P, _P, C.c_longlong]),
    "cude_loss_sharded": (C.c_int, [_P, _P, C.POINTER(cude_net), C.POINTER(cude_opts), C.c_int, _D, C.c_longlong, _D, C.c_longlong,
                                    C.c_longlong, _D, _D]),
    "cude_loss_grad_sharded": (C.c_int, [_P, _P, C.POINTER(cude_net), C.POINTER(cude_opts), C.c_int, _D, C.c_longlong, _D, C.c_longlong,
                                         C.c_longlong, C.c_int, _D, _D, _D, _D]),
    # ... and one process driving all GPUs
    "cude_mctx_create": (C.c_int, [C.c_int, _I, C.POINTER(_P)]),
    "cude_mctx_destroy": (C.c_int, [_P]),
    "cude_mctx_size": (C.c_int, [_P]),
    "cude_mctx_ctx": (_P, [_P, C.c_int]),
    "cude_mlast_error": (C.c_char_p, [_P]),
    "cude_mpopulation_create": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _I, _D, _D, C.c_int, _I, _D, _D, _D, _D, C.POINTER(_P)]),
    "cude_mpopulation_destroy": (C.c_int, [_P]),
    "cude_mpopulation_size": (C.c_int, [_P]),
    "cude_mpopulation_mode": (C.c_int, [_P]),
    "cude_mloss": (C.c_int, [_P, _P, C.POINTER(cude_net), C.POINTER(cude_opts), C.c_int, _D, C.c_longlong, _D, _D, _D]),
    "cude_mloss_grad": (C.c_int, [_P, _P, C.POINTER(cude_net), C.POINTER(cude_opts), C.c_int, _D, C.c_longlong, _D,
                                  C.c_int, _D, _D, _D, _D]),
    "cude_mget_stats": (C.c_int, [_P, C.POINTER(cude_stats)]),
    "cude_sup_population_create": (C.c_int, [_P, C.c_int, C.c_int, _D, _D, _D, _D, C.c_double, C.c_double, C.POINTER(_P)]),
    "cude_sup_population_destroy": (C.c_int, [_P]),
    "cude_sup_loss_grad": (C.c_int, [_P, _P, C.c_int, C.c_int, C.POINTER(cude_opts), C.c_int, _D, C.c_longlong, _D, C.c_double,
                                     _D, _D, _D, _D]),
}

_lib = None


def load():
    """Load the CUDA library; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C conditional_ude_b200/csrc`.  conditional_ude_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.cude_abi_version() != ABI_VERSION:
        raise ImportError("libcude_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc, ctx=None):
    if rc != CUDE_OK:
        msg = load().cude_last_error(ctx)
        raise CudeError(rc, msg.decode() if msg else "?")


def mcheck(rc, mctx=None):
    if rc != CUDE_OK:
        msg = load().cude_mlast_error(mctx)
        raise CudeError(rc, msg.decode() if msg else "?")
