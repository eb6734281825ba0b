"""Lazy Gramian: the host-side mirror of reference src/gramian.jl for the multiply path.

``gramian(k, x[, y])`` stays lazy (the reference's O(1) constructor, src/gramian.jl:18-21,144-159); the first use
creates the device handle through the C ABI (points are copied to the GPU(s) once).  ``mul_(y, G, x, alpha, beta)``
is the reference's ``mul!`` (src/gramian.jl:78-99, 241-257) and dispatches on the shapes exactly like it:
vector -> MVM, matrix -> multi-RHS, GradientKernel Gramian -> block MVM on the flat (n d) vector.
Everything numerical happens in libcovfn_b200.so; there is no fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import CF_F32, CF_F64, DimensionMismatch, KNode, UnsupportedKernel, check, lib
from .kernels import AbstractKernel, Dot, DotProductInput, GradientKernel, IsotropicInput


def _points(x, dtype=None):
    """Accepts what gramian() accepts (src/gramian.jl:144-155): a d x n matrix whose COLUMNS are points, a vector of
    vectors, or a vector of scalars (d = 1).  Returns an (n, d) C-contiguous array == d x n column-major, ld = d."""
    if isinstance(x, np.ndarray) and x.ndim == 2:
        pts = np.ascontiguousarray(x.T)  # columns are points
    else:
        seq = list(x)
        if len(seq) and np.ndim(seq[0]) == 0:
            pts = np.asarray(seq).reshape(-1, 1)
        else:
            lens = {len(v) for v in seq}
            if len(lens) > 1:
                raise DimensionMismatch(f"points do not all have the same length: {sorted(lens)}")
            pts = np.asarray(seq)
            if pts.ndim == 1:
                pts = pts.reshape(len(seq), -1)
    if dtype is None:
        dtype = np.float32 if pts.dtype == np.float32 else np.float64
    return np.ascontiguousarray(pts, dtype=dtype)


class Gramian:
    """Gramian{T,K,U,V} (src/gramian.jl:10-14): K[i, j] = k(x[i], y[j]), never instantiated."""

    def __init__(self, k, x, y=None, _rows_are_points=False):
        if not isinstance(k, (AbstractKernel, GradientKernel)):
            raise TypeError("Gramian(k, x, y): k must be a kernel")
        self.k = k
        self._symmetric = y is None or y is x
        if _rows_are_points:  # internal: already (n, d) point arrays
            self.x = x
            self.y = x if self._symmetric else y
        else:
            self.x = _points(x)
            self.y = self.x if self._symmetric else _points(y)
            if self.y.dtype != self.x.dtype:  # promote_type(eltype(x), eltype(y)) (src/gramian.jl:30-33): Float32 with Float64 is Float64
                self.x, self.y = self.x.astype(np.float64), self.y.astype(np.float64)
        if self.x.shape[1] != self.y.shape[1]:
            raise DimensionMismatch(
                f"inputs have to have the same length: {self.x.shape[1]}, {self.y.shape[1]}")  # src/util.jl:41
        self.dtype = self.x.dtype  # gramian_eltype: promote(eltype(k), coordinate types) (src/gramian.jl:30-33)
        self.is_gradient = isinstance(k, GradientKernel)
        self.block = (self.x.shape[1] + k.block_extra) if self.is_gradient else 1  # BlockFactorization block size
        self._handle = None
        self._rows = None

    # ---- shape -------------------------------------------------------------------------------------------------
    @property
    def d(self):
        return self.x.shape[1]

    @property
    def shape(self):
        b = self.block  # BlockFactorization size (d n) x (d m) (test/gradient.jl:35), (d+1) for ValueGradientKernel
        return (self.x.shape[0] * b, self.y.shape[0] * b)

    def size(self, i=None):
        return self.shape if i is None else self.shape[i - 1]

    @property
    def eltype(self):
        return self.dtype

    def issymmetric(self):  # src/gramian.jl:132
        return self._symmetric or (self.x.shape == self.y.shape and np.array_equal(self.x, self.y))

    @property
    def T(self):  # adjoint / transpose (src/gramian.jl:116-117)
        return Gramian(self.k, self.y, None if self._symmetric else self.x, _rows_are_points=True)

    # ---- device handle -----------------------------------------------------------------------------------------
    def handle(self):
        if self._handle is None:
            prog = self.k.program()
            arr = (KNode * len(prog))()
            for t, (op, ip, fp) in enumerate(prog):
                arr[t].op, arr[t].iparam, arr[t].fparam = op, ip, fp
            h = C.c_void_p()
            n, d = self.x.shape
            m = self.y.shape[0]
            check(lib().cf_gramian_create(
                C.byref(h), arr, len(prog), CF_F64 if self.dtype == np.float64 else CF_F32, d, n,
                self.x.ctypes.data_as(C.c_void_p), d, m,
                None if self._symmetric else self.y.ctypes.data_as(C.c_void_p), d))
            self._handle = h
            if self._rows is not None:
                check(lib().cf_gramian_set_row_range(h, self._rows[0], self._rows[1]))
        return self._handle

    def set_row_range(self, row_begin: int, row_end: int):
        """Restrict the rows (points, for gradient Gramians) this object computes: contiguous row-block sharding."""
        n = self.x.shape[0]
        if not (0 <= row_begin <= row_end <= n):
            raise DimensionMismatch(f"row range [{row_begin}, {row_end}) is not inside [0, {n})")
        self._rows = (int(row_begin), int(row_end))
        if self._handle is not None:
            check(lib().cf_gramian_set_row_range(self._handle, *self._rows))
        return self

    def set_symmetric(self, on: bool = True):
        """opt in to the symmetric variant (CF_OPT_SYMMETRIC): each unordered pair evaluated once; see include/covfn_b200.h"""
        check(lib().cf_gramian_set_option(self.handle(), 1, int(bool(on))))
        return self

    @property
    def row_range(self):
        return self._rows if self._rows is not None else (0, self.x.shape[0])

    def close(self):
        if self._handle is not None:
            lib().cf_gramian_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- indexing / dense ----------------------------------------------------------------------------------------
    def __getitem__(self, ij):  # getindex(G, i, j) (src/gramian.jl:37-40); 0-based here
        if self.is_gradient:
            raise UnsupportedKernel("scalar indexing of a gradient Gramian is not on the device path")
        i, j = ij
        n, m = self.shape
        if not (isinstance(i, (int, np.integer)) and isinstance(j, (int, np.integer))):
            raise TypeError("only scalar indexing G[i, j] is lowered; use Matrix(G) for blocks")
        if not (0 <= i < n and 0 <= j < m):
            raise IndexError(f"BoundsError: attempt to access {n}x{m} Gramian at index [{i}, {j}]")
        out = C.c_double()
        check(lib().cf_gramian_getindex(self.handle(), int(i), int(j), C.byref(out)))
        return self.dtype.type(out.value)

    def Matrix(self):
        """Matrix(G) / Matrix!(M, G) (src/gramian.jl:102-114)."""
        if self.is_gradient:
            raise UnsupportedKernel("dense instantiation of a gradient Gramian is not on the device path")
        r0, r1 = self.row_range
        n, m = r1 - r0, self.y.shape[0]
        M = np.empty((m, max(n, 1)), dtype=self.dtype)  # column-major n x m
        check(lib().cf_gramian_matrix(self.handle(), M.ctypes.data_as(C.c_void_p), max(n, 1)))
        return M.T[:n, :]

    # ---- products -------------------------------------------------------------------------------------------------
    def __matmul__(self, a):  # *(G, a) / *(G, A) (src/gramian.jl:66-75)
        a = np.asarray(a)
        r0, r1 = self.row_range
        rows = (r1 - r0) * self.block
        # T = promote_type(eltype(G), eltype(a)) (src/gramian.jl:67,72).  The device computes in the Gramian's eltype: a Float64 vector
        # against a Float32 Gramian is multiplied in Float32 and the result returned as Float64 (the reference evaluates the entries
        # in Float32 too and only accumulates in Float64 -- a summation-order-class difference, inside the Float32 tolerance)
        dt = np.promote_types(self.dtype, a.dtype) if a.dtype.kind == "f" else self.dtype
        if a.ndim == 1:
            b = np.zeros(rows, dtype=self.dtype)
        else:
            b = np.zeros((rows, a.shape[1]), dtype=self.dtype, order="F")
        b = mul_(b, self, a)
        return b if dt == self.dtype else b.astype(dt)

    __mul__ = __matmul__

    def __add__(self, other):  # D + G / G + D stays lazy (src/gramian.jl:55-60)
        if isinstance(other, Diagonal):
            return LazyMatrixSum(self, other)
        return NotImplemented

    __radd__ = __add__

    def last_timing(self):
        ms, launches = C.c_float(), C.c_int()
        check(lib().cf_last_timing(self.handle(), C.byref(ms), C.byref(launches)))
        return ms.value, launches.value

    # ---- device-resident multiply (torch tensors or raw pointers) ------------------------------------------------
    def mul_device(self, y_ptr: int, x_ptr: int, nrhs: int = 1, ldy: int = 0, ldx: int = 0, alpha=1.0, beta=0.0, stream: int = 0):
        fn = lib().cf_gramian_mul_device
        if self.is_gradient:
            fn = lib().cf_value_gradient_mul_device if self.k.block_extra else lib().cf_gradient_mul_device
        check(fn(self.handle(), C.c_void_p(y_ptr), ldy, C.c_void_p(x_ptr), ldx, nrhs, float(alpha), float(beta),
                 C.c_void_p(stream) if stream else None))


def mul_collective_device(G, y_full_ptr: int, x_ptr: int, alpha=1.0, beta=0.0, stream: int = 0):
    """cf_gramian_mul_collective_device: the complete product on every rank of the library's communicator (distributed.comm_init_from_torch)"""
    check(lib().cf_gramian_mul_collective_device(G.handle(), C.c_void_p(y_full_ptr), C.c_void_p(x_ptr), float(alpha), float(beta),
                                                  C.c_void_p(stream) if stream else None))


def gramian(k, x=None, y=None):
    """gramian(k, x[, y]) (src/gramian.jl:144-159).  GradientKernel -> lazy block Gramian (src/gramian.jl:120-123)."""
    if not isinstance(k, (AbstractKernel, GradientKernel)):
        # gramian(x, y) = Gramian(Dot(), x, y), gramian(x) = gramian(x, x) (src/gramian.jl:23,150-151)
        if x is None:
            return Gramian(Dot(), k)
        return Gramian(Dot(), k, x)
    return Gramian(k, x, y)


def _as_vec(a, dtype, what):
    a = np.asarray(a)
    if a.dtype != dtype:
        a = a.astype(dtype)
    return a


def mul_(y, G, x, alpha=1, beta=0):
    """mul!(y, G, x, α=1, β=0): y <- α G x + β y; β == 0 overwrites y, NaNs included (src/gramian.jl:78-99, 241-257).
    y is modified in place (it must be a writable array of the Gramian's eltype; vectors contiguous, matrices
    column-major) and returned."""
    if isinstance(G, LazyMatrixSum):
        return G.mul_(y, x, alpha, beta)
    if not isinstance(G, Gramian):
        raise TypeError("mul_(y, G, x): G must be a Gramian")
    x = np.asarray(x)
    if not isinstance(y, np.ndarray):
        raise TypeError("y must be a numpy array (modified in place)")
    if y.ndim != x.ndim or y.ndim not in (1, 2):
        raise DimensionMismatch("y and x must both be vectors or both be matrices")
    r0, r1 = G.row_range
    blk = G.block
    rows, cols = (r1 - r0) * blk, G.shape[1]
    if y.shape[0] != rows or x.shape[0] != cols:
        raise DimensionMismatch(
            f"Gramian block is {rows}x{cols}, y has {y.shape[0]} rows, x has {x.shape[0]} rows")
    nrhs = 1 if y.ndim == 1 else y.shape[1]
    if x.ndim == 2 and x.shape[1] != nrhs:
        raise DimensionMismatch(f"y has {nrhs} columns, x has {x.shape[1]}")
    dt = G.dtype
    if y.dtype != dt:
        raise TypeError(f"y has dtype {y.dtype}, the Gramian's eltype is {dt}")
    xc = np.asarray(x, dtype=dt)
    if y.ndim == 1:
        xc = np.ascontiguousarray(xc)
        ywork = y if y.flags.c_contiguous else np.ascontiguousarray(y)
        ldy = ldx = 0
    else:
        xc = np.asfortranarray(xc)
        ywork = y if y.flags.f_contiguous else np.asfortranarray(y)
        ldy, ldx = max(rows, 1), max(cols, 1)
    fn = lib().cf_gramian_mul
    if G.is_gradient:
        fn = lib().cf_value_gradient_mul if G.k.block_extra else lib().cf_gradient_mul
        if G.k.input_trait() not in (IsotropicInput(), DotProductInput()):
            raise UnsupportedKernel("derivative-kernel MVMs are lowered for IsotropicInput and DotProductInput kernels only "
                                    "(src/gradient.jl:83-115); GenericInput kernels use the reference's dense fallback")
    check(fn(G.handle(), ywork.ctypes.data_as(C.c_void_p), ldy, xc.ctypes.data_as(C.c_void_p), ldx, nrhs,
             float(alpha), float(beta)))
    if ywork is not y:
        y[...] = ywork
    return y


class Diagonal:
    """sigma2 * I(n): the only diagonal the lowered solver needs (reference: Diagonal + Gramian -> LazyMatrixSum)."""

    def __init__(self, sigma2: float, n: int):
        self.sigma2, self.n = float(sigma2), int(n)

    @property
    def shape(self):
        return (self.n, self.n)


def I(n: int):
    return Diagonal(1.0, n)


def _scale_diag(c, D):
    return Diagonal(c * D.sigma2, D.n)


Diagonal.__rmul__ = lambda self, c: _scale_diag(c, self)
Diagonal.__mul__ = lambda self, c: _scale_diag(c, self)


class LazyMatrixSum:
    """LazyMatrixSum(D, G) (src/lazy_linear_algebra.jl:91-133) restricted to sigma2*I + Gramian, the config-5 operator."""

    def __init__(self, G: Gramian, D: Diagonal):
        if G.shape[0] != G.shape[1] or D.n != G.shape[0]:
            raise DimensionMismatch("LazyMatrixSum: sizes differ")  # src/lazy_linear_algebra.jl:95
        self.G, self.D = G, D

    @property
    def shape(self):
        return self.G.shape

    def mul_(self, y, x, alpha=1, beta=0):  # src/lazy_linear_algebra.jl:126-133
        x = np.asarray(x, dtype=self.G.dtype)
        if beta == 0:
            y[...] = 0
        else:
            y *= beta
        r0, r1 = self.G.row_range  # a row-restricted Gramian (multi-process sharding) owns rows [r0, r1) of the sum as well
        b = self.G.block
        y += alpha * self.D.sigma2 * x[r0 * b:r1 * b]
        return mul_(y, self.G, x, alpha, 1)

    def __matmul__(self, x):
        x = np.asarray(x)
        y = np.zeros(x.shape, dtype=self.G.dtype, order="F" if x.ndim == 2 else "C")
        return self.mul_(y, x)

    def solve(self, b, x0=None, reltol=0.0, maxiter=0):
        """A \\ b -> ldiv! -> cg! (src/lazy_linear_algebra.jl:135-144), run entirely on the device(s).
        Returns (x, iterations, residual norm)."""
        G = self.G
        b = np.ascontiguousarray(b, dtype=G.dtype)
        N = G.shape[0]
        if b.shape != (N,):
            raise DimensionMismatch(f"b has length {b.shape}, operator is {N}x{N}")
        x = np.zeros(N, dtype=G.dtype) if x0 is None else np.array(x0, dtype=G.dtype, copy=True)
        iters, res = C.c_int(), C.c_double()
        check(lib().cf_cg_solve(G.handle(), self.D.sigma2, x.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
                                float(reltol), int(maxiter), (1 + G.k.block_extra) if G.is_gradient else 0, C.byref(iters), C.byref(res)))
        return x, iters.value, res.value

    def cg_timing(self):
        """(total ms, operator-product ms, all-gather ms, products) of the last solve (cf_cg_timing)"""
        t, m, ga, k = C.c_double(), C.c_double(), C.c_double(), C.c_int()
        check(lib().cf_cg_timing(self.G.handle(), C.byref(t), C.byref(m), C.byref(ga), C.byref(k)))
        return t.value, m.value, ga.value, k.value


def peak_probe(kind: str = "dfma", iters: int = 1 << 16):
    """Measured pipe peak on the current device: lane-instructions per second (bench.py's roofline denominator)."""
    code = {"dfma": 0, "ffma": 1, "mufu": 2}[kind]
    ops, ms = C.c_double(), C.c_float()
    check(lib().cf_peak_probe(code, iters, C.byref(ops), C.byref(ms)))
    return ops.value, ms.value


def jit_stats():
    """Counters of the run-time specialisation of composite kernel programs (csrc/cf_jit.h): kernels compiled, cache hits,
    failures (each falls back to the ahead-of-time interpreter kernel) and seconds spent compiling, since process start."""
    c, h, f, t = C.c_int(), C.c_int(), C.c_int(), C.c_double()
    check(lib().cf_jit_stats(C.byref(c), C.byref(h), C.byref(f), C.byref(t)))
    return {"compiled": c.value, "cache_hits": h.value, "failures": f.value, "compile_seconds": t.value}


JIT_KERNELS = {"mvm": 0, "mm_dmma": 1, "mvm_dmma": 2, "mm_tf32": 3, "mvm_tf32": 4, "grad_dmma": 5, "mm_tf32_legacy": 6, "mvm_tc5": 7}


def jit_check(kernel, d: int, which: str = "mvm"):
    """Compile (only) the run-time specialisation of one device kernel for `kernel`'s program -- works without a GPU (NVRTC
    cross-compiles for sm_100a).  Returns the NVRTC log; raises if the generated code does not build."""
    prog = kernel.program()
    arr = (KNode * len(prog))()
    for i, (op, ip, fp) in enumerate(prog):
        arr[i].op, arr[i].iparam, arr[i].fparam = op, ip, fp
    buf = C.create_string_buffer(1 << 16)
    check(lib().cf_jit_check(arr, len(prog), int(d), JIT_KERNELS[which], buf, len(buf)))
    return buf.value.decode(errors="replace")
