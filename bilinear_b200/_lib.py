"""ctypes binding of the C-ABI shared library (include/bilinear_b200.h).

This is the reference-side stub a maintainer would add under model/bilinear.py:
plain pointers and sizes, no torch types cross the boundary.  The product path has NO fallback: if
the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BILINEAR_B200_LIB: an alternative build of the same library (A/B measurements of compile-time variants)
LIB_PATH = os.environ.get("BILINEAR_B200_LIB") or os.path.join(_HERE, "lib", "libbilinear_b200.so")

PARAM_NUMEL = 4291632
NUM_BN = 5
PREC_BF16 = 1
PREC_FP32 = 3
PLAN_TRAIN = 0
PLAN_INFER = 1

c_float_p = C.POINTER(C.c_float)
c_i64_p = C.POINTER(C.c_int64)


class PlanDesc(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("precision", C.c_int32), ("kind", C.c_int32), ("reserved", C.c_int32),
        ("params", C.c_void_p), ("grads", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
        ("running_mean", C.c_void_p * NUM_BN), ("running_var", C.c_void_p * NUM_BN),
        ("num_batches_tracked", C.c_void_p * NUM_BN),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
    ]


class AdamArgs(C.Structure):
    _fields_ = [
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("step", C.c_int64), ("max_norm", C.c_float), ("grad_scale", C.c_float),
    ]


class TrainArgs(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("y", C.c_void_p), ("p_out", C.c_void_p), ("masks", C.c_void_p),
        ("seed", C.c_uint64), ("step", C.c_uint64), ("momentum", C.c_float * NUM_BN),
        ("loss_out", C.c_void_p), ("gnorm_out", C.c_void_p),
    ]


class CommDesc(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("epoch", C.c_uint32), ("reserved", C.c_uint32),
                ("buffers", C.c_void_p * 8)]


_SIGS = {
    "b200_version": (C.c_int, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_param_numel": (C.c_int64, []),
    "b200_launch_count": (C.c_int64, []),
    "b200_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "b200_plan_create": (C.c_int, [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]),
    "b200_plan_destroy": (None, [C.c_void_p]),
    "b200_refresh_weights": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200_forward_eval": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200_forward_train": (C.c_int, [C.c_void_p, C.POINTER(TrainArgs), C.c_void_p]),
    "b200_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200_train_step": (C.c_int, [C.c_void_p, C.POINTER(TrainArgs), C.POINTER(AdamArgs), C.c_int32, C.c_void_p]),
    "b200_train_step_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(TrainArgs), C.POINTER(AdamArgs),
                                       C.c_int32, C.c_void_p, C.c_void_p]),
    "b200_forward_eval_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "b200_comm_bytes": (C.c_int64, []),
    "b200_comm_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "b200_comm_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "b200_comm_close": (C.c_int, [C.c_void_p]),
    "b200_comm_free": (C.c_int, [C.c_void_p]),
    "b200_plan_set_comm": (C.c_int, [C.c_void_p, C.POINTER(CommDesc)]),
    "b200_train_step_dp": (C.c_int, [C.c_void_p, C.POINTER(TrainArgs), C.POINTER(AdamArgs), C.POINTER(C.c_uint32),
                                     C.c_void_p]),
    "b200_adam_step": (C.c_int, [C.c_void_p, C.POINTER(AdamArgs), C.c_void_p, C.c_void_p]),
    "b200_mpjpe": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 4),
    "b200_pmpjpe": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 4),
    "b200_mpjpe_pmpjpe": (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_void_p, C.c_int32] + [C.c_void_p] * 6),
    "b200_debug_keep_bits": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "b200_debug_phase_ns": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "b200_debug_gemm": (C.c_int, [C.c_void_p] * 3 + [C.c_int32] * 7 + [C.c_void_p, C.c_int64, C.c_void_p]),
    "b200_debug_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
}

EXPORTS = tuple(_SIGS)
_lib = None


def lib():
    """Load (once) and return the shared library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "bilinear_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C bilinear_b200/csrc`; there is no CPU / PyTorch fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise RuntimeError("bilinear_b200: " + lib().b200_last_error().decode("utf-8", "replace"))


def ptr(t):
    """device (or host) address of a torch tensor, None -> NULL"""
    return None if t is None else C.c_void_p(t.data_ptr())
