"""ctypes binding of libcovfn_b200.so -- the C ABI declared in include/covfn_b200.h.

The library is built in-tree (csrc/Makefile -> lib/libcovfn_b200.so).  There is no CPU fallback: if the
shared library is missing, loading raises; if no CUDA device is usable, every compute call raises CudaError.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libcovfn_b200.so")

CF_OK = 0
CF_ERR_BAD_ARGUMENT, CF_ERR_DIMENSION, CF_ERR_UNSUPPORTED, CF_ERR_DOMAIN = -1, -2, -3, -4
CF_ERR_CUDA, CF_ERR_NCCL, CF_ERR_NONFINITE, CF_ERR_INTERNAL = -5, -6, -7, -8
CF_F32, CF_F64 = 0, 1


class KNode(C.Structure):
    """cf_knode_t"""
    _fields_ = [("op", C.c_int32), ("iparam", C.c_int32), ("fparam", C.c_double)]


class CovFnError(RuntimeError):
    pass


class DimensionMismatch(CovFnError, ValueError):
    """Julia DimensionMismatch (reference src/util.jl:9,41)"""


class DomainError(CovFnError, ValueError):
    """Julia DomainError (reference src/stationary.jl:19,47,124)"""


class UnsupportedKernel(CovFnError, NotImplementedError):
    """kernel tree not lowerable to the device; in Julia the shim falls through to the reference method"""


class CudaError(CovFnError):
    pass


_ERRORS = {
    CF_ERR_BAD_ARGUMENT: CovFnError,
    CF_ERR_DIMENSION: DimensionMismatch,
    CF_ERR_UNSUPPORTED: UnsupportedKernel,
    CF_ERR_DOMAIN: DomainError,
    CF_ERR_CUDA: CudaError,
    CF_ERR_NCCL: CudaError,
    CF_ERR_NONFINITE: DomainError,
    CF_ERR_INTERNAL: CovFnError,
}

# every symbol include/covfn_b200.h declares: (name, restype, argtypes)
_i64, _int, _dbl, _vp = C.c_int64, C.c_int, C.c_double, C.c_void_p
SYMBOLS = {
    "cf_version": (_int, []),
    "cf_last_error": (C.c_char_p, []),
    "cf_device_count": (_int, []),
    "cf_init": (_int, [_int, C.POINTER(_int)]),
    "cf_gramian_create": (_int, [C.POINTER(_vp), C.POINTER(KNode), _int, _int, _int, _i64, _vp, _i64, _i64, _vp, _i64]),
    "cf_gramian_destroy": (_int, [_vp]),
    "cf_gramian_size": (_int, [_vp, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_int), C.POINTER(_int)]),
    "cf_gramian_set_row_range": (_int, [_vp, _i64, _i64]),
    "cf_gramian_set_option": (_int, [_vp, _int, _int]),
    "cf_gramian_mul": (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _dbl, _dbl]),
    "cf_gramian_mul_device": (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _dbl, _dbl, _vp]),
    "cf_gramian_matrix": (_int, [_vp, _vp, _i64]),
    "cf_gramian_getindex": (_int, [_vp, _i64, _i64, C.POINTER(_dbl)]),
    "cf_gradient_mul": (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _dbl, _dbl]),
    "cf_gradient_mul_device": (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _dbl, _dbl, _vp]),
    "cf_value_gradient_mul": (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _dbl, _dbl]),
    "cf_value_gradient_mul_device": (_int, [_vp, _vp, _i64, _vp, _i64, _i64, _dbl, _dbl, _vp]),
    "cf_cg_solve": (_int, [_vp, _dbl, _vp, _vp, _dbl, _int, _int, C.POINTER(_int), C.POINTER(_dbl)]),
    "cf_last_timing": (_int, [_vp, C.POINTER(C.c_float), C.POINTER(_int)]),
    "cf_peak_probe": (_int, [_int, _int, C.POINTER(_dbl), C.POINTER(C.c_float)]),
    "cf_cg_timing": (_int, [C.c_void_p, C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_int)]),
    "cf_comm_unique_id": (_int, [C.c_void_p, _int]),
    "cf_comm_init": (_int, [C.c_void_p, _int, _int]),
    "cf_comm_destroy": (_int, []),
    "cf_comm_info": (_int, [C.POINTER(_int), C.POINTER(_int), C.POINTER(_int)]),
    "cf_gramian_mul_collective_device": (_int, [C.c_void_p, C.c_void_p, C.c_void_p, _dbl, _dbl, C.c_void_p]),
    "cf_comm_allgather_rows": (_int, [C.c_void_p, C.c_int64, C.c_int64, _int, C.c_void_p]),
    "cf_jit_stats": (_int, [C.POINTER(_int), C.POINTER(_int), C.POINTER(_int), C.POINTER(_dbl)]),
    "cf_jit_check": (_int, [C.POINTER(KNode), _int, _int, _int, C.c_char_p, _int]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CovFnError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C covariancefunctions.jl_b200/csrc).  This package has no CPU fallback."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status == CF_OK:
        return
    msg = lib().cf_last_error().decode("utf-8", "replace")
    raise _ERRORS.get(status, CovFnError)(msg)


def device_count() -> int:
    return lib().cf_device_count()


def init(devices) -> None:
    devices = list(devices)
    arr = (C.c_int * len(devices))(*devices)
    check(lib().cf_init(len(devices), arr))
