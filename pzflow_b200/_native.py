"""ctypes binding of libpzflow_b200.so (include/pzflow_b200.h).

There is deliberately no fallback: if the CUDA library is missing or no CUDA
device is visible, every evaluation entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PZFLOW_B200_LIB: another build of the same library (kernel experiments, A/B runs); default: the in-tree build
LIB_PATH = os.environ.get("PZFLOW_B200_LIB") or os.path.join(_HERE, "libpzflow_b200.so")

PZB_OK = 0
PZB_ERR_INVALID = -1
PZB_ERR_UNSUPPORTED = -2
PZB_ERR_CUDA = -3
PZB_ERR_NO_DEVICE = -4

STAGE_SHIFT_BOUNDS = 1
STAGE_STANDARD_SCALER = 2
STAGE_ROLL = 3
STAGE_NSC = 4
STAGE_CHAIN_BEGIN = 5
STAGE_CHAIN_END = 6
STAGE_REVERSE = 7
STAGE_SCALE = 8
STAGE_INV_SOFTPLUS = 9
STAGE_COLOR_TRANSFORM = 10
STAGE_SHUFFLE = 11
STAGE_DEQUANTIZE = 12

LATENT_UNIFORM = 1
LATENT_CENTBETA13 = 2
LATENT_NORMAL = 3
LATENT_CENTBETA = 4
LATENT_TDIST = 5
LATENT_JOINT = 6

ENGINE_AUTO = 0
ENGINE_SIMT = 1
ENGINE_TC = 2
ENGINE_WIDE = 3

c_float_p = C.POINTER(C.c_float)


class StageDesc(C.Structure):
    """pzb_stage_desc."""

    _fields_ = [
        ("kind", C.c_int32),
        ("shift", C.c_int32),
        ("vec_a", c_float_p),
        ("vec_b", c_float_p),
        ("n_vec", C.c_int32),
        ("B", C.c_float),
        ("K", C.c_int32),
        ("hidden_layers", C.c_int32),
        ("hidden_dim", C.c_int32),
        ("transformed_dim", C.c_int32),
        ("n_conditions", C.c_int32),
        ("periodic", C.c_int32),
        ("weights", C.POINTER(c_float_p)),
    ]


class LatentDesc(C.Structure):
    """pzb_latent_desc."""

    _fields_ = [("kind", C.c_int32), ("dim", C.c_int32), ("B", C.c_float), ("params", c_float_p),
                ("n_params", C.c_int32)]


# every symbol include/pzflow_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "pzb_abi_version": (C.c_int, []),
    "pzb_last_error": (C.c_char_p, []),
    "pzb_device_count": (C.c_int, []),
    "pzb_plan_create": (C.c_int, [C.POINTER(StageDesc), C.c_int32, C.c_int32, C.c_int32, C.c_float,
                                  C.c_int32, C.POINTER(C.c_void_p)]),
    "pzb_plan_set_latent": (C.c_int, [C.c_void_p, C.POINTER(LatentDesc), C.c_int32]),
    "pzb_plan_destroy": (None, [C.c_void_p]),
    "pzb_plan_input_dim": (C.c_int, [C.c_void_p]),
    "pzb_plan_n_conditions": (C.c_int, [C.c_void_p]),
    "pzb_plan_n_spline_dims": (C.c_int, [C.c_void_p]),
    "pzb_plan_flops_per_row": (C.c_double, [C.c_void_p]),
    "pzb_plan_tc_supported": (C.c_int, [C.c_void_p]),
    "pzb_plan_wide_supported": (C.c_int, [C.c_void_p]),
    "pzb_plan_overflow_rows": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "pzb_plan_set_engine": (C.c_int, [C.c_void_p, C.c_int32]),
    "pzb_plan_engine": (C.c_int, [C.c_void_p]),
    "pzb_launch_count": (C.c_int64, []),
    "pzb_log_prob": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                               C.c_void_p]),
    "pzb_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "pzb_inverse": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "pzb_posterior": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "pzb_ensemble_logmeanexp": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "pzb_ensemble_posterior_mean": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p,
                                              C.c_int32, C.c_void_p, C.c_void_p]),
    "pzb_normalize_pdfs": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_void_p]),
    "pzb_sample_uniform": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_uint64, C.c_int64, C.c_void_p]),
    "pzb_sample_uniform_threefry": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_uint32, C.c_uint32,
                                              C.c_int64, C.c_int64, C.c_int32, C.c_void_p]),
    "pzb_rational_quadratic_spline": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                                C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                                C.c_void_p]),
    "pzb_set_host_threads": (C.c_int, [C.c_int32]),
    "pzb_set_error_stream": (C.c_int, [C.c_int32, C.c_int64]),
    "pzb_sample_normal_threefry": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_uint32, C.c_uint32, C.c_int64, C.c_int64,
                                             C.c_int32, C.c_void_p]),
    "pzb_host_digest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "pzb_sample_centbeta": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_float, C.c_uint64,
                                      C.c_int64, C.c_void_p]),
    "pzb_sample_normal": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_int64, C.c_void_p]),
    "pzb_sample_tdist": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_float, C.c_uint64, C.c_int64, C.c_void_p]),
    "pzb_log_prob_err": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                   C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pzb_posterior_err": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                    C.c_void_p, C.c_int32, C.c_int32, C.c_uint64, C.c_int64, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "pzb_posterior_marg": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                     C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_uint64, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pzb_error_eps": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, C.c_void_p]),
    "pzb_spline_train_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float,
                                           C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pzb_spline_train_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                            C.c_int32, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pzb_plan_train_supported": (C.c_int, [C.c_void_p]),
    "pzb_plan_set_train_layout": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64]),
    "pzb_train_workspace_floats": (C.c_int64, [C.c_void_p, C.c_int64]),
    "pzb_train_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64,
                                 C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "pzb_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float,
                                C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "pzb_log_prob_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pzb_log_prob_host_columns": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.c_void_p]),
    "pzb_inverse_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_void_p]),
    "pzb_inverse_to_host_columns": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pzb_ensemble_log_prob_host_columns": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p,
                                                     C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "pzb_ensemble_posterior_host": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                              C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "pzb_log_prob_host_columns_f32": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                            C.c_void_p]),
    "pzb_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pzb_inverse_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "pzb_posterior_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
}

_lib = None


class NativeError(RuntimeError):
    """CUDA / library failure that has no reference-side exception type."""


def lib() -> C.CDLL:
    """Load libpzflow_b200.so (built in-tree by pzflow_b200._build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} is missing: build it with `python -m pzflow_b200._build` "
                "(nvcc, sm_100a). pzflow_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.pzb_abi_version() != 1:
            raise NativeError("libpzflow_b200.so ABI version mismatch; rebuild it")
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().pzb_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a pzb_status to the reference's exception types."""
    if rc == PZB_OK:
        return
    msg = last_error()
    if rc == PZB_ERR_INVALID:
        raise ValueError(msg)
    if rc == PZB_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise NativeError(msg)
