# -*- coding: utf-8 -*-
"""ctypes binding of the C ABI declared in include/fastproject_b200.h.

The product path has NO fallback: if the shared library is missing or a call
fails, an exception is raised.  Build it with ``python __graft_entry__.py``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libfastproject_b200.so")

c_i64 = ctypes.c_int64
c_f64 = ctypes.c_double
c_ptr = ctypes.c_void_p
c_int = ctypes.c_int
c_size = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/fastproject_b200.h declares
SIGNATURES = {
    "fp_abi_version": (c_int, []),
    "fp_version": (ctypes.c_char_p, []),
    "fp_last_error": (ctypes.c_char_p, []),
    "fp_launch_count": (ctypes.c_uint64, []),
    "fp_match_labels": (c_int, [ctypes.c_char_p, c_i64, c_i64, ctypes.c_char_p, c_i64, c_i64, c_ptr]),
    "fp_gene_detected_mean": (c_int, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr]),
    "fp_score_workspace_bytes": (c_size, [c_i64, c_i64, c_i64]),
    "fp_score_signatures": (c_int, [c_int, c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64,
                                    c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_ptr, c_ptr, c_i64, c_ptr, c_size, c_ptr]),
    "fp_rank_workspace_bytes": (c_size, [c_i64, c_i64, c_i64]),
    "fp_rank_columns": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_size, c_ptr]),
    "fp_rank_columns_ex": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr, c_size, c_int, c_ptr]),
    "fp_normalize_rows_workspace_bytes": (c_size, [c_i64]),
    "fp_normalize_rows": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_size, c_ptr]),
    "fp_normalize_cols": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64, c_ptr]),
    "fp_consistency_workspace_bytes": (c_size, [c_i64, c_i64, c_int]),
    "fp_consistency": (c_int, [c_ptr, c_ptr, c_f64, c_ptr, c_i64, c_i64, c_i64, c_ptr, c_i64,
                               c_int, c_int, c_ptr, c_size, c_ptr]),
    "fp_consistency_status": (c_int, [c_ptr, c_i64, c_i64, c_int, c_ptr]),
    "fp_values_pack_bytes": (c_size, [c_i64, c_i64]),
    "fp_values_pack": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_size, c_ptr]),
    "fp_consistency_packed": (c_int, [c_ptr, c_ptr, c_f64, c_ptr, c_i64, c_i64, c_ptr, c_i64, c_int, c_ptr, c_size,
                                      c_ptr]),
    "fp_compute_weights_workspace_bytes": (c_size, [c_i64]),
    "fp_compute_weights": (c_int, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i64, c_ptr, c_i64, c_ptr,
                                   c_size, c_ptr]),
    "fp_pull_host": (c_int, [c_ptr, c_ptr, c_size, c_ptr]),
    "fp_row_medians": (c_int, [c_ptr, c_i64, c_i64, c_i64, c_ptr, c_ptr]),
    "fp_background_pvalues": (c_int, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_i64,
                                      c_ptr, c_ptr, c_ptr, c_ptr]),
}

METHOD_CODES = {"naive": 0, "weighted_avg": 1, "imputed": 2, "only_nonzero": 3}
PRECISION_CODES = {"fp64": 0, "tc": 1}
OUT_ABS_ERROR, OUT_PREDICTION = 0, 1
OUT_ANY_CELL_ORDER = 0x200
RANK_AVERAGE, RANK_MIN = 0, 1
FP_EUNSUPPORTED = -4

_lib = None


class NativeError(RuntimeError):
    pass


def load():
    """Load the shared library (idempotent).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            "fastproject_b200: %s not found -- the CUDA extension is not built "
            "(run `python __graft_entry__.py`); there is no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_pack_ext = None


def pack_ext():
    """The CPython extension next to the shared library (csrc/pack_ext.c: signature dicts -> CSR
    arrays).  Raises if it has not been built."""
    global _pack_ext
    if _pack_ext is None:
        import glob
        import importlib.util
        found = sorted(glob.glob(os.path.join(os.path.dirname(LIB_PATH), "_fp_pack*.so")))
        if not found:
            raise NativeError("fastproject_b200: _fp_pack extension not found in %s -- run "
                              "`python __graft_entry__.py`" % os.path.dirname(LIB_PATH))
        spec = importlib.util.spec_from_file_location("_fp_pack", found[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _pack_ext = mod
    return _pack_ext


def call(name, *args):
    """Call an int-returning entry point; raise NativeError on a non-zero code."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        msg = lib.fp_last_error()
        raise NativeError("%s failed (%d): %s" % (name, rc, msg.decode() if msg else ""))


def query(name, *args):
    return getattr(load(), name)(*args)


def launch_count():
    return int(load().fp_launch_count())
