"""ctypes binding of libcoda_b200.so (the C ABI in include/coda_b200.h).

There is no CPU fallback: importing works anywhere (so the symbol table can be checked
without a GPU), but every compute entry point needs an sm_100 device and ``require_device``
fails loudly without one.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

_LIB = None

OK = 0
FLAG_NONFINITE_INPUT = 0x01
FLAG_RANGE_INPUT = 0x02
FLAG_NONFINITE_TABLE = 0x04
FLAG_NONFINITE_PI = 0x08
FLAG_NONFINITE_PBEST = 0x10
FLAG_NONFINITE_EIG = 0x20
FLAG_NAMES = {
    FLAG_NONFINITE_INPUT: "preds", FLAG_RANGE_INPUT: "preds range", FLAG_NONFINITE_TABLE: "pdf/cdf/integrand",
    FLAG_NONFINITE_PI: "pi_hat_xi", FLAG_NONFINITE_PBEST: "Pbest", FLAG_NONFINITE_EIG: "Pbest(beta) normalized",
}

p, i32, i64, f64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t

# name -> (restype, argtypes); mirrors include/coda_b200.h one to one
SIGNATURES = {
    "coda_b200_last_error": (C.c_char_p, []),
    "coda_b200_version": (i32, []),
    "coda_b200_sm_count": (i32, []),
    "coda_b200_device_check": (i32, []),
    "coda_b200_scan_slab": (i32, [p, i32, i64, i32, p, p, p, p, p, p]),
    "coda_b200_confusion_accum": (i32, [p, p, i32, i64, i32, i32, p, p]),
    "coda_b200_confusion_sorted": (i32, [p, p, p, i32, i64, i32, i32, p, p]),
    "coda_b200_init_dirichlets": (i32, [p, i32, i32, i32, f64, f64, i32, p, p]),
    "coda_b200_pi_full": (i32, [p, p, i32, i64, i32, p, p]),
    "coda_b200_pi_reduce": (i32, [p, i64, i32, i32, p, p, p, p]),
    "coda_b200_label_row": (i32, [p, i32, i64, p, p, p, p]),
    "coda_b200_label_apply": (i32, [p, i32, i32, p, p, f64, p]),
    "coda_b200_pi_rank1": (i32, [p, p, p, p, i32, i64, i32, p, p, f64, i32, p, p, p, p, i32, p]),
    "coda_b200_shadow_build": (i32, [p, i32, i64, i32, p, i32, p, p]),
    "coda_b200_set_l2_fetch_granularity": (i32, [i32]),
    "coda_b200_tables_scratch_bytes": (sz, [i32, i32]),
    "coda_b200_beta_tables": (i32, [p, p, i32, i32, i32, f64, i32, i32, p, p, p, p, p, p, p, p, p, p]),
    "coda_b200_pair_rows_tc": (i32, [p, i32, i32, p, p, p, p, p, p, i32, p, p, p, p, p, p]),
    "coda_b200_mixture": (i32, [p, p, i32, i32, p, p, p, p, p, p]),
    "coda_b200_pair_count": (i32, [p, i32, i64, i32, p, p, p]),
    "coda_b200_pair_fill": (i32, [p, i32, i64, i32, p, p, p, p, p, p, p, p, p]),
    "coda_b200_pair_rows": (i32, [p, i32, i32, p, p, p, p, p, p, p, i32, p, p, p, p, p, p]),
    "coda_b200_pair_gain": (i32, [p, p, i64, i32, p, p, p, p, p, p, i32, i32, p]),
    "coda_b200_eig_blocks": (i32, [i64]),
    "coda_b200_eig_points": (i32, [p, i64, i32, p, p, p, p, p, p, p, i64, p, i32, p, p, p, p]),
    "coda_b200_ell_build": (i32, [p, p, p, i64, i32, p, p]),
    "coda_b200_select_merge": (i32, [p, i32, p, p]),
    "coda_b200_ties": (i32, [p, i64, p, p, i64, p, i32, p, p, p, p]),
    "coda_b200_device_pick": (i32, [p, p, p, i64, i64, p, p, p, p, i64, p]),
}


class NativeError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """dlopen the in-tree library (building it with nvcc first if it is not there)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if build_if_missing and not os.environ.get("CODA_B200_NO_BUILD"):
        try:
            _build.build()            # no-op when the in-tree .so matches the sources
        except Exception as e:        # e.g. no nvcc on this box: use the shipped .so if there is one
            if not os.path.exists(path):
                raise NativeError(f"cannot build {path}: {e}") from e
    if not os.path.exists(path):
        raise NativeError(f"{path} is missing; run `python -m coda_b200.build`")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError here == header/library drift
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def last_error() -> str:
    return load().coda_b200_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc != OK:
        raise NativeError(f"coda_b200 {what} failed (rc={rc}): {last_error()}")


def require_device() -> None:
    """Fail loudly if the CUDA path cannot run (no silent CPU route exists)."""
    check(load().coda_b200_device_check(), "device_check")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
