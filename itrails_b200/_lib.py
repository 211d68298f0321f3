"""ctypes binding of libitrails_b200.so (the C ABI in include/itrails_b200.h)."""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libitrails_b200.so")


class ItrailsError(RuntimeError):
    """A call into libitrails_b200 failed (message from itr_last_error)."""

    def __init__(self, code, msg):
        super().__init__(f"[itr_status {code}] {msg}")
        self.code = code


class ItrailsCudaError(ItrailsError):
    """The CUDA library or device is missing/unusable.  There is no CPU fallback."""


ITR_ERR_ARG, ITR_ERR_STATE, ITR_ERR_CUDA, ITR_ERR_NOMEM, ITR_ERR_UNSUPPORTED, ITR_ERR_IO = -1, -2, -3, -4, -5, -6
PHASES = {"loglik": 0, "viterbi_fwd": 1, "viterbi_trace": 2, "post_fwd": 3, "post_bwd": 4,
          "model": 5, "emit_table": 6, "post_combine": 7, "post_total": 8}

_c_ctx = ctypes.c_void_p
_dp = ctypes.POINTER(ctypes.c_double)
_u8p = ctypes.POINTER(ctypes.c_uint8)
_u16p = ctypes.POINTER(ctypes.c_uint16)
_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)

# name -> (restype, argtypes); every symbol declared in include/itrails_b200.h
SIGNATURES = {
    "itr_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_c_ctx)]),
    "itr_destroy": (None, [_c_ctx]),
    "itr_last_error": (ctypes.c_char_p, [_c_ctx]),
    "itr_version": (ctypes.c_int, []),
    "itr_load_blocks": (ctypes.c_int, [_c_ctx, _u16p, _i64p, ctypes.c_int64]),
    "itr_load_blocks_i64": (ctypes.c_int, [_c_ctx, _i64p, _i64p, ctypes.c_int64]),
    "itr_set_model": (ctypes.c_int, [_c_ctx, ctypes.c_int, ctypes.c_int, _dp, _dp, _dp]),
    "itr_build_model": (ctypes.c_int, [_c_ctx, ctypes.c_int, _dp, ctypes.c_int, ctypes.c_int,
                                       _dp, _dp, _dp, _dp, _dp, _i32p]),
    "itr_num_states": (ctypes.c_int, [ctypes.c_int, ctypes.c_int]),
    "itr_plan_info": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, _i32p, _i32p, _i64p, _i64p, _i32p]),
    "itr_loglik": (ctypes.c_int, [_c_ctx, _dp, _dp]),
    "itr_viterbi": (ctypes.c_int, [_c_ctx, _dp, _dp, _dp, _u8p]),
    "itr_viterbi_fetch": (ctypes.c_int, [_c_ctx, _u8p]),
    "itr_viterbi_fetch_range": (ctypes.c_int, [_c_ctx, ctypes.c_int64, ctypes.c_int64, _u8p]),
    "itr_posterior": (ctypes.c_int, [_c_ctx, _dp]),
    "itr_posterior_fetch": (ctypes.c_int, [_c_ctx, _dp]),
    "itr_posterior_fetch_range": (ctypes.c_int, [_c_ctx, ctypes.c_int64, ctypes.c_int64, _dp]),
    "itr_posterior_write_csv": (ctypes.c_int, [_c_ctx, ctypes.c_char_p, _i64p, ctypes.c_int]),
    "itr_posterior_write_csv_ex": (ctypes.c_int, [_c_ctx, ctypes.c_char_p, _i64p, _i64p, ctypes.c_int, _i64p, ctypes.c_int]),
    "itr_posterior_stream": (ctypes.c_int, [_c_ctx, _dp, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "itr_csv_posterior_host": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_int64, _i64p, _i64p, _dp, ctypes.c_int]),
    "itr_csv_posterior_host_ex": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_int64, _i64p, _i64p, _dp, _i64p,
                                                 ctypes.c_int, _i64p, ctypes.c_int]),
    "itr_csv_format_double": (ctypes.c_int, [ctypes.c_double, ctypes.c_char_p, ctypes.c_int]),
    "itr_set_async": (ctypes.c_int, [_c_ctx, ctypes.c_int]),
    "itr_sync": (ctypes.c_int, [_c_ctx]),
    "itr_maf_read": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_char_p, ctypes.c_int,
                                    ctypes.POINTER(ctypes.c_void_p), ctypes.c_char_p, ctypes.c_int]),
    "itr_maf_free": (None, [ctypes.c_void_p]),
    "itr_maf_num_blocks": (ctypes.c_int64, [ctypes.c_void_p]),
    "itr_maf_num_columns": (ctypes.c_int64, [ctypes.c_void_p]),
    "itr_maf_export": (ctypes.c_int, [ctypes.c_void_p, _u16p, _i64p, ctypes.c_int]),
    "itr_maf_symbols": (_u16p, [ctypes.c_void_p]),
    "itr_maf_offsets": (_i64p, [ctypes.c_void_p]),
    "itr_maf_num_coord_blocks": (ctypes.c_int64, [ctypes.c_void_p]),
    "itr_maf_coordinates": (_i64p, [ctypes.c_void_p]),
    "itr_maf_coord_offsets": (_i64p, [ctypes.c_void_p]),
    "itr_phase_ms": (ctypes.c_double, [_c_ctx, ctypes.c_int]),
    "itr_launch_count": (ctypes.c_int64, [_c_ctx]),
    "itr_lockstep_launch_count": (ctypes.c_int64, [_c_ctx]),
    "itr_total_columns": (ctypes.c_int64, [_c_ctx]),
    "itr_num_blocks": (ctypes.c_int64, [_c_ctx]),
    "itr_device_info": (ctypes.c_int, [_c_ctx, ctypes.c_char_p, ctypes.c_int,
                                       ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                       ctypes.POINTER(ctypes.c_int)]),
}

# int sink(void *user, int64_t col0, int64_t n_cols, const double *rows)
ROWS_SINK = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, _dp)

_lib = None


def load():
    """Load the shared library; raises ItrailsCudaError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ItrailsCudaError(
            ITR_ERR_CUDA,
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C itrails_b200/csrc`.  itrails_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def as_ptr(arr, ctype):
    return arr.ctypes.data_as(ctypes.POINTER(ctype)) if arr is not None else None


def check(lib, ctx, rc):
    if rc == 0:
        return
    msg = lib.itr_last_error(ctx)
    msg = msg.decode("utf-8", "replace") if msg else ""
    if rc == ITR_ERR_ARG:
        raise ValueError(msg)
    if rc == ITR_ERR_CUDA:
        raise ItrailsCudaError(rc, msg)
    if rc == ITR_ERR_NOMEM:
        raise MemoryError(msg)
    if rc == ITR_ERR_IO:
        raise OSError(msg or "file could not be written")
    raise ItrailsError(rc, msg)


def c_f64(x):
    return np.ascontiguousarray(x, dtype=np.float64)
