"""ctypes front-end of the CPU oracle (oracle/covfn_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, ``__graft_entry__.smoke()`` and the cpu_baseline / ``--impl reference`` legs of bench.py
may import this module.  The product never does.

A kernel *program* here is a plain list of ``(op, iparam, fparam)`` tuples in postfix order, with the
op numbering of include/covfn_b200.h (``cf_op``) -- the oracle does not import the product's kernel
classes.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

OP_EQ, OP_EXP, OP_RQ, OP_MATERNP, OP_DOT, OP_CONST, OP_SUM, OP_PROD, OP_POW, OP_LENGTHSCALE, OP_ARDSCALE, OP_ARD = range(1, 13)


class _KNode(C.Structure):
    _fields_ = [("op", C.c_int32), ("iparam", C.c_int32), ("fparam", C.c_double)]


def build(force: bool = False) -> str:
    """Compile liboracle.so with the Makefile next to this file (gcc, OpenMP)."""
    src = os.path.join(_HERE, "covfn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_getindex.restype = C.c_double
        _lib.orc_truth_getindex.restype = C.c_double
        _lib.orc_cg_solve.restype = C.c_int
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _prog(program):
    arr = (_KNode * len(program))()
    for t, (op, ip, fp) in enumerate(program):
        arr[t].op, arr[t].iparam, arr[t].fparam = int(op), int(ip), float(fp)
    return arr, C.c_int(len(program))


def _pts(X, dtype):
    """points as (n, d) C-contiguous == d x n column-major with ldx = d"""
    X = np.ascontiguousarray(X, dtype=dtype)
    if X.ndim == 1:
        X = X.reshape(-1, 1)
    return X


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i64(v):
    return C.c_int64(int(v))


def num_threads() -> int:
    return lib().orc_num_threads()


def set_num_threads(t: int) -> None:
    lib().orc_set_num_threads(C.c_int(t))


def getindex(program, X, i, Y, j, dtype=np.float64) -> float:
    X, Y = _pts(X, dtype), _pts(Y, dtype)
    pa, nn = _prog(program)
    return lib().orc_getindex(pa, nn, C.c_int(dtype == np.float64), C.c_int(X.shape[1]), _p(X), _i64(X.shape[1]),
                              _i64(i), _p(Y), _i64(Y.shape[1]), _i64(j))


def matrix(program, X, Y=None, dtype=np.float64):
    X = _pts(X, dtype)
    Y = X if Y is None else _pts(Y, dtype)
    n, m, d = X.shape[0], Y.shape[0], X.shape[1]
    M = np.empty((m, n), dtype=dtype)  # column-major n x m
    pa, nn = _prog(program)
    lib().orc_matrix(pa, nn, C.c_int(dtype == np.float64), C.c_int(d), _i64(n), _p(X), _i64(d), _i64(m), _p(Y),
                     _i64(d), _p(M), _i64(n))
    return M.T


def mul_vec(program, X, a, Y=None, alpha=1.0, beta=0.0, y0=None, rows=None, dtype=np.float64):
    """mul!(y, gramian(k, X, Y), a, alpha, beta) restated (reference src/gramian.jl:78-87)."""
    X = _pts(X, dtype)
    Y = X if Y is None else _pts(Y, dtype)
    n, m, d = X.shape[0], Y.shape[0], X.shape[1]
    i0, i1 = (0, n) if rows is None else rows
    a = np.ascontiguousarray(a, dtype=dtype)
    assert a.shape == (m,)
    y = np.zeros(i1 - i0, dtype=dtype) if y0 is None else np.array(y0, dtype=dtype, copy=True)
    pa, nn = _prog(program)
    lib().orc_gramian_mul_vec(pa, nn, C.c_int(dtype == np.float64), C.c_int(d), _i64(n), _p(X), _i64(d), _i64(m),
                              _p(Y), _i64(d), _i64(i0), _i64(i1), _p(y), _p(a), C.c_double(alpha), C.c_double(beta))
    return y


def mul_mat(program, X, A, Y=None, alpha=1.0, beta=0.0, B0=None, rows=None, dtype=np.float64, fused=False):
    """mul!(B, gramian(k, X, Y), A, alpha, beta) restated (reference src/gramian.jl:89-99).
    A is (m, p); returns (rows, p).  fused=True evaluates each entry once (baseline timing only)."""
    X = _pts(X, dtype)
    Y = X if Y is None else _pts(Y, dtype)
    n, m, d = X.shape[0], Y.shape[0], X.shape[1]
    i0, i1 = (0, n) if rows is None else rows
    A = np.asfortranarray(A, dtype=dtype)
    p = A.shape[1]
    B = np.zeros((i1 - i0, p), dtype=dtype, order="F") if B0 is None else np.array(B0, dtype=dtype, order="F", copy=True)
    pa, nn = _prog(program)
    fn = lib().orc_gramian_mul_mat_fused if fused else lib().orc_gramian_mul_mat
    fn(pa, nn, C.c_int(dtype == np.float64), C.c_int(d), _i64(n), _p(X), _i64(d), _i64(m), _p(Y), _i64(d), _i64(i0),
       _i64(i1), _p(B), _i64(B.shape[0]), _p(A), _i64(A.shape[0]), _i64(p), C.c_double(alpha), C.c_double(beta))
    return B


def value_derivative_laplacian(program, r2: float):
    out = (C.c_double * 3)()
    pa, nn = _prog(program)
    lib().orc_value_derivative_laplacian(pa, nn, C.c_double(r2), out)
    return out[0], out[1], out[2]


def gradient_mul(program, X, a, Y=None, alpha=1.0, beta=0.0, y0=None, rows=None):
    """blockmul! with the isotropic gradient element (reference src/gramian.jl:241-253, src/gradient.jl:86-92).
    a is flat (m*d,), returns flat (rows*d,)."""
    X = _pts(X, np.float64)
    Y = X if Y is None else _pts(Y, np.float64)
    n, m, d = X.shape[0], Y.shape[0], X.shape[1]
    i0, i1 = (0, n) if rows is None else rows
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.shape == (m * d,)
    y = np.zeros((i1 - i0) * d) if y0 is None else np.array(y0, dtype=np.float64, copy=True)
    pa, nn = _prog(program)
    lib().orc_gradient_mul(pa, nn, C.c_int(d), _i64(n), _p(X), _i64(d), _i64(m), _p(Y), _i64(d), _i64(i0), _i64(i1),
                           _p(y), _p(a), C.c_double(alpha), C.c_double(beta))
    return y


def gradient_matrix(program, X, Y=None):
    X = _pts(X, np.float64)
    Y = X if Y is None else _pts(Y, np.float64)
    n, m, d = X.shape[0], Y.shape[0], X.shape[1]
    M = np.empty((m * d, n * d))  # column-major (n d) x (m d)
    pa, nn = _prog(program)
    lib().orc_gradient_matrix(pa, nn, C.c_int(d), _i64(n), _p(X), _i64(d), _i64(m), _p(Y), _i64(d), _p(M), _i64(n * d))
    return M.T


def cg_solve(program, X, b, sigma2, x0=None, reltol=0.0, maxiter=0, gradient=False):
    """(sigma2 I + K) \\ b restated (reference src/lazy_linear_algebra.jl:126-144 + IterativeSolvers 0.9.2 cg!).
    Returns (x, iterations, residual norm, residual history)."""
    X = _pts(X, np.float64)
    n, d = X.shape
    N = n * d if gradient else n
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros(N) if x0 is None else np.array(x0, dtype=np.float64, copy=True)
    hist = np.zeros(maxiter if maxiter > 0 else N)
    res = C.c_double(0)
    pa, nn = _prog(program)
    it = lib().orc_cg_solve(pa, nn, C.c_int(d), _i64(n), _p(X), _i64(d), C.c_double(sigma2), _p(x), _p(b),
                            C.c_double(reltol), C.c_int(maxiter), C.c_int(int(gradient)), C.byref(res), _p(hist))
    return x, it, res.value, hist[:it]


def truth_mul_vec(program, X, a, Y=None, alpha=1.0, rows=None):
    """long-double evaluation of the same product (exact leaves, 64-bit-mantissa accumulation)."""
    X = _pts(X, np.float64)
    Y = X if Y is None else _pts(Y, np.float64)
    n, m, d = X.shape[0], Y.shape[0], X.shape[1]
    i0, i1 = (0, n) if rows is None else rows
    a = np.ascontiguousarray(a, dtype=np.float64)
    y = np.zeros(i1 - i0)
    pa, nn = _prog(program)
    lib().orc_truth_mul_vec(pa, nn, C.c_int(d), _i64(n), _p(X), _i64(d), _i64(m), _p(Y), _i64(d), _i64(i0), _i64(i1),
                            _p(y), _p(a), C.c_double(alpha), C.c_double(0.0))
    return y


def truth_getindex(program, x, y) -> float:
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    y = np.ascontiguousarray(y, dtype=np.float64).ravel()
    pa, nn = _prog(program)
    return lib().orc_truth_getindex(pa, nn, C.c_int(x.size), _p(x), _p(y))


def derivative_mul(program, X, a, Y=None, trait="isotropic", value_gradient=False, alpha=1.0, beta=0.0, y0=None, rows=None):
    """blockmul! for GradientKernel / ValueGradientKernel with the isotropic or dot-product element
    (reference src/gramian.jl:241-253, src/gradient.jl:86-92, 109-115, 442-463).  Flat vectors, blocks of d (+1)."""
    X = _pts(X, np.float64)
    Y = X if Y is None else _pts(Y, np.float64)
    n, m, d = X.shape[0], Y.shape[0], X.shape[1]
    bs = d + (1 if value_gradient else 0)
    i0, i1 = (0, n) if rows is None else rows
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.shape == (m * bs,)
    y = np.zeros((i1 - i0) * bs) if y0 is None else np.array(y0, dtype=np.float64, copy=True)
    pa, nn = _prog(program)
    lib().orc_derivative_mul(pa, nn, C.c_int(d), C.c_int(0 if trait == "isotropic" else 1), C.c_int(int(value_gradient)),
                             _i64(n), _p(X), _i64(d), _i64(m), _p(Y), _i64(d), _i64(i0), _i64(i1), _p(y), _p(a),
                             C.c_double(alpha), C.c_double(beta))
    return y


def derivative_matrix(program, X, Y=None, trait="isotropic", value_gradient=False):
    X = _pts(X, np.float64)
    Y = X if Y is None else _pts(Y, np.float64)
    n, m, d = X.shape[0], Y.shape[0], X.shape[1]
    bs = d + (1 if value_gradient else 0)
    M = np.zeros((m * bs, n * bs))  # column-major (n bs) x (m bs)
    pa, nn = _prog(program)
    lib().orc_derivative_matrix(pa, nn, C.c_int(d), C.c_int(0 if trait == "isotropic" else 1), C.c_int(int(value_gradient)),
                                _i64(n), _p(X), _i64(d), _i64(m), _p(Y), _i64(d), _p(M), _i64(n * bs))
    return M.T
