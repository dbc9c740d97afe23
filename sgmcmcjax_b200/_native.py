"""ctypes binding of libsgmcmc_b200.so (include/sgmcmc_b200.h).

There is no CPU fallback: if the library is missing or a CUDA device is not
available, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build

MAX_LEAVES = 8
ABI_VERSION = 4

FAMILY_GAUSSIAN_MEAN, FAMILY_LOGREG, FAMILY_MLP, FAMILY_PMF = 0, 1, 2, 3
EST_STANDARD, EST_CV, EST_SVRG = 0, 1, 2
(DIFF_SGLD, DIFF_PSGLD, DIFF_SGLDADAM, DIFF_SGHMC, DIFF_BAOAB, DIFF_SGNHT, DIFF_BADODAB, DIFF_ADAM_OPT) = range(8)
FLAG_BADODABCV_REFERENCE_BUG = 1
SHARD_UPDATE2, SHARD_STEP = 1, 2

EXPORTS = [
    "sgmcmc_last_error", "sgmcmc_abi_version", "sgmcmc_launch_count", "sgmcmc_create", "sgmcmc_destroy", "sgmcmc_launch_shape", "sgmcmc_set_data",
    "sgmcmc_set_centering", "sgmcmc_fullbatch_grad", "sgmcmc_init", "sgmcmc_run", "sgmcmc_run_thinned", "sgmcmc_optimize", "sgmcmc_kernel_step",
    "sgmcmc_set_row_shard", "sgmcmc_sum_width", "sgmcmc_set_centering_grad", "sgmcmc_shard_init", "sgmcmc_shard_step",
    "sgmcmc_minibatch_indices", "sgmcmc_normal", "sgmcmc_uniform", "sgmcmc_bernoulli_rows", "sgmcmc_split", "sgmcmc_ksd_imq_partial", "sgmcmc_ksd_workspace_bytes", "sgmcmc_ksd_imq_tc",
]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("family", C.c_int32), ("estimator", C.c_int32), ("diffusion", C.c_int32),
        ("n_data", C.c_int64), ("batch", C.c_int32), ("n_leaves", C.c_int32),
        ("leaf_size", C.c_int32 * MAX_LEAVES), ("data_dim", C.c_int32), ("leapfrog", C.c_int32),
        ("update_rate", C.c_int32), ("flags", C.c_int32),
        ("lik_coef", C.c_float * MAX_LEAVES), ("prior_coef", C.c_float * MAX_LEAVES), ("hyper", C.c_float * 4),
        ("dims", C.c_int32 * 8),
    ]


class State(C.Structure):
    _fields_ = [
        ("n_chains", C.c_int32), ("reserved", C.c_int32),
        ("theta", C.c_void_p), ("aux1", C.c_void_p), ("aux2", C.c_void_p), ("alpha", C.c_void_p),
        ("grad", C.c_void_p), ("svrg_center", C.c_void_p), ("svrg_grad", C.c_void_p), ("key", C.c_void_p),
    ]


class NativeError(RuntimeError):
    pass


_lib = None


def lib_path() -> str:
    # SGMCMC_B200_LIB: load an alternative build of the same ABI (kernel-variant experiments)
    return os.environ.get("SGMCMC_B200_LIB") or _build.LIB_PATH


def load(build_if_missing: bool = True):
    """Load (building first if the .so is absent/stale and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if build_if_missing and path == _build.LIB_PATH and _build.is_stale():
        try:
            _build.build()                        # serialised across processes by a file lock (torchrun ranks)
        except Exception as e:
            if not os.path.exists(path):
                raise NativeError(f"libsgmcmc_b200.so is missing and could not be built: {e}") from e
            if not _build.toolchain_missing(e):   # nvcc ran and failed: a stale library would hide the broken sources
                raise NativeError(f"libsgmcmc_b200.so is older than its sources and the rebuild failed: {e}") from e
            import warnings

            warnings.warn(f"libsgmcmc_b200.so is older than its sources and cannot be rebuilt here ({e}); using it as is")
    if not os.path.exists(path):
        raise NativeError(f"{path} not found; run `python -m sgmcmcjax_b200._build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.sgmcmc_last_error.restype = C.c_char_p
    lib.sgmcmc_last_error.argtypes = []
    lib.sgmcmc_abi_version.restype = i32
    lib.sgmcmc_launch_count.restype = i64
    lib.sgmcmc_launch_count.argtypes = []
    sigs = {
        "sgmcmc_create": [C.POINTER(Config), C.POINTER(vp)],
        "sgmcmc_destroy": [vp],
        "sgmcmc_launch_shape": [vp, i32, C.POINTER(i32), C.POINTER(i32)],
        "sgmcmc_set_data": [vp, vp, vp, vp],
        "sgmcmc_set_centering": [vp, vp, vp],
        "sgmcmc_fullbatch_grad": [vp, vp, i32, vp, vp],
        "sgmcmc_init": [vp, C.POINTER(State), vp, i32, vp],
        "sgmcmc_run": [vp, C.POINTER(State), i32, i32, vp, i32, vp, i64, vp],
        "sgmcmc_run_thinned": [vp, C.POINTER(State), i32, i32, vp, i32, i32, vp, i64, vp, vp],
        "sgmcmc_kernel_step": [vp, C.POINTER(State), i32, vp, vp, vp],
        "sgmcmc_optimize": [vp, C.POINTER(State), i32, i32, vp, vp, vp],
        "sgmcmc_set_row_shard": [vp, i64, i64],
        "sgmcmc_set_centering_grad": [vp, vp, vp, vp],
        "sgmcmc_sum_width": [vp],
        "sgmcmc_shard_init": [vp, C.POINTER(State), vp, i32, vp, vp],
        "sgmcmc_shard_step": [vp, C.POINTER(State), i32, vp, vp, i32, vp, vp, vp, vp],
        "sgmcmc_minibatch_indices": [C.POINTER(C.c_uint32), i64, i32, vp, vp],
        "sgmcmc_normal": [C.POINTER(C.c_uint32), i32, vp, vp],
        "sgmcmc_split": [C.POINTER(C.c_uint32), i32, vp, vp],
        "sgmcmc_uniform": [C.POINTER(C.c_uint32), i32, vp, vp],
        "sgmcmc_bernoulli_rows": [C.POINTER(C.c_uint32), i32, vp, vp, vp],
        "sgmcmc_ksd_imq_partial": [vp, vp, i32, i32, i32, i32, vp, vp],
        "sgmcmc_ksd_imq_tc": [vp, vp, i32, i32, i32, i32, vp, vp, vp],
    }
    lib.sgmcmc_ksd_workspace_bytes.restype = i64
    lib.sgmcmc_ksd_workspace_bytes.argtypes = [i32, i32]
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = i32
    if lib.sgmcmc_abi_version() != ABI_VERSION:
        raise NativeError("libsgmcmc_b200.so ABI version mismatch; rebuild it")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().sgmcmc_last_error().decode("utf-8", "replace")
        raise NativeError(f"libsgmcmc_b200 error {rc}: {msg}")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise NativeError("sgmcmcjax_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch
