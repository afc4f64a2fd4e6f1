"""ctypes binding of libcokrig_b200.so (the C ABI declared in include/cokrig.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded this module raises,
and every compute wrapper in ``cokrig_b200.ops`` raises when no CUDA device is present.
"""
from __future__ import annotations

import ctypes
import os

import torch  # noqa: F401  (first: libcokrig_b200.so links the shared CUDA runtime, which torch brings along)
from ctypes import POINTER, c_char_p, c_double, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COKRIG_B200_LIB", os.path.join(_HERE, "libcokrig_b200.so"))

CK_OK = 0
CK_ERR_ARG = -1
CK_ERR_CUDA = -2
CK_ERR_UNSUPPORTED = -3
METRIC_EUCLID = 0
METRIC_HAVERSINE = 1


class CokrigError(RuntimeError):
    """A C-ABI call returned a CK_ERR_* status."""


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python sif-xco2-cokriging_b200/build.py` "
            "(or __graft_entry__.build()). cokrig_b200 has no CPU fallback.")
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_dp = c_void_p  # device pointers are passed as integers (torch.Tensor.data_ptr())
_hp = POINTER(c_double)  # host double arrays

# name -> (restype, argtypes); mirrors include/cokrig.h one to one
SIGNATURES = {
    "ck_version": (c_int, []),
    "ck_last_error": (c_char_p, []),
    "ck_launch_count": (ctypes.c_longlong, []),
    "ck_matern_eval": (c_int, [_dp, c_int64, c_double, c_double, c_double, c_double, _dp, c_void_p]),
    "ck_distance_block": (c_int, [_dp, c_int64, _dp, c_int64, c_int, _dp, c_int64, c_void_p]),
    "ck_matern_block": (c_int, [_dp, c_int64, _dp, c_int64, c_int, c_double, c_double, c_double, c_double,
                                _dp, c_int64, _dp, c_int64, c_int, c_void_p]),
    "ck_joint_cov": (c_int, [_dp, c_int64, _dp, c_int64, _hp, c_int, c_int, _dp, c_int64, c_void_p]),
    "ck_cross_cov": (c_int, [_dp, c_int64, _dp, c_int64, _dp, c_int64, _hp, c_int, c_int, c_int, _dp, c_int64,
                             c_void_p]),
    "ck_potrf_workspace_bytes": (c_size_t, [c_int64]),
    "ck_potrf": (c_int, [_dp, c_int64, c_int64, _dp, _dp, c_void_p]),
    "ck_trsm_lower": (c_int, [_dp, c_int64, c_int64, _dp, _dp, c_int64, c_int64, c_void_p]),
    "ck_potrs_predict": (c_int, [_dp, c_int64, c_int64, _dp, _dp, c_int64, c_int64, _dp, c_double, _dp, _dp,
                                 c_void_p]),
    "ck_logdet": (c_int, [_dp, c_int64, c_int64, _dp, c_void_p]),
    "ck_nll": (c_int, [_dp, c_int64, _dp, c_int64, _hp, c_int, c_int, _dp, _dp, c_int64, _dp, _dp, _dp, _dp,
                       c_void_p]),
    "ck_vario_minmax_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ck_vario_tile_rows": (c_int64, [c_int64]),
    "ck_vario_tile_cols": (c_int64, [c_int64]),
    "ck_vario_tile_shape": (c_int, [POINTER(c_int), POINTER(c_int)]),
    "ck_vario_minmax": (c_int, [_dp, c_int64, _dp, c_int64, c_int, c_int, c_double, c_int64, c_int64, _dp, _dp, c_void_p]),
    "ck_vario_candidates": (c_int, [_dp, c_int64, _dp, c_int64, c_int, c_int, c_double, c_double, c_double, c_int64,
                                    c_int64, _dp, _dp, c_int64, _dp, c_void_p]),
    "ck_vario_bin_partials_offset": (c_size_t, [c_int]),
    "ck_vario_bin_tiles": (c_int, [_dp, _dp, c_int64, c_double, _dp, _dp, c_int64, c_double, c_int, c_int, c_int,
                                   c_double, _hp, c_int, c_int64, c_int64, _dp, c_int64, _dp, _dp, c_void_p]),
    "ck_vario_bin_reduce": (c_int, [c_int64, c_int64, c_int, _dp, _dp, _dp, c_void_p]),
    "ck_mg_local_tiles": (c_int64, [c_int64, c_int64, c_int64]),
    "ck_mg_assemble": (c_int, [_dp, c_int64, _dp, c_int64, _dp, c_int64, _dp, _hp, c_int, c_int, c_int, c_int64,
                               c_int, c_int, c_int, c_int, _dp, c_int64, c_void_p]),
    "ck_mg_unique_id": (c_int, [c_void_p]),
    "ck_mg_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int64, c_void_p]),
    "ck_mg_destroy": (c_int, [c_void_p]),
    "ck_mg_grid": (c_int, [c_void_p, POINTER(c_int)]),
    "ck_mg_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int64]),
    "ck_mg_joint_cov": (c_int, [c_void_p, _dp, c_int64, _dp, c_int64, _dp, c_int64, _dp, _hp, c_int, c_int, c_int, _dp, c_size_t,
                                c_void_p]),
    "ck_mg_potrf": (c_int, [c_void_p, c_void_p]),
    "ck_mg_potrs_predict": (c_int, [c_void_p, _dp, _dp, _dp, c_void_p]),
    "ck_mg_logdet": (c_int, [c_void_p, _dp, c_void_p]),
    "ck_mg_times_ms": (c_int, [c_void_p, _hp]),
    "ck_mg_local_factor": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64), POINTER(c_int64)]),
    "ck_gemm_nt": (c_int, [_dp, c_int64, _dp, c_int64, _dp, c_int64, c_int64, c_int64, c_int64, c_int, c_void_p]),
    "ck_mg_update": (c_int, [_dp, c_int64, _dp, c_int64, _dp, c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                             c_int64, c_int64, c_int64, c_void_p]),
    "ck_row_dots": (c_int, [_dp, c_int64, c_int64, c_int64, _dp, c_int64, _dp, _dp, c_void_p]),
    "ck_oz_configure": (c_int, [c_int, c_int64]),
    "ck_potf2_debug_buffer": (c_int, [_dp]),
    "ck_oz_active": (c_int, [c_int64]),
    "ck_oz_debug_buffer": (c_int, [_dp]),
    "ck_oz_slices_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "ck_oz_scales_len": (c_int64, [c_int64]),
    "ck_oz_split": (c_int, [_dp, c_int64, c_int64, c_int64, _dp, _dp, _dp, c_void_p]),
    "ck_oz_mg_update": (c_int, [_dp, _dp, c_int64, _dp, _dp, c_int64, c_int64, _dp, c_int64, c_int64, c_int64, c_int64, c_int64,
                                c_int64, c_int, c_void_p]),
    "ck_oz_tile_order": (c_int64, [c_int64, c_int64, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, POINTER(c_int), c_int64,
                                   POINTER(c_int64), POINTER(c_int64)]),
    "ck_oz_gemm": (c_int, [_dp, _dp, c_int64, _dp, _dp, c_int64, c_int64, _dp, c_int64, c_int, c_int, c_void_p]),
    "ck_vario_bin_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "ck_vario_bin": (c_int, [_dp, _dp, c_int64, c_double, _dp, _dp, c_int64, c_double, c_int, c_int, c_int,
                             c_double, _hp, c_int, _dp, _dp, _dp, c_int64, _dp, _dp, c_void_p]),
    "ck_local_predict_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "ck_xcor_lags": (c_int, [_dp, _dp, c_int64, c_int, POINTER(c_int), c_int, c_int, c_int, _dp, _dp, _dp, c_void_p]),
    "ck_local_debug_buffer": (c_int, [_dp]),
    "ck_local_count": (c_int, [_dp, c_int64, _dp, c_int64, _dp, c_int64, c_int, c_int, c_int, c_double, c_int,
                               _dp, _dp, c_void_p]),
    "ck_local_predict": (c_int, [_dp, _dp, c_int64, _dp, _dp, c_int64, _dp, c_int64, _hp, _hp, c_int, c_int, c_int,
                                 c_double, c_int, c_double, _dp, c_int64, _dp, _dp, c_int64, _dp, _dp, _dp, _dp, c_void_p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = library/header mismatch: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    msg = lib.ck_last_error()
    return msg.decode() if msg else ""


def check(rc: int, what: str = "") -> None:
    if rc != CK_OK:
        raise CokrigError(f"{what or 'cokrig_b200'} failed (status {rc}): {last_error()}")
