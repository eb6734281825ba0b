"""Parity of the MaternP form of the scaled-domain value kernel (csrc/gram_mvm_eq.cuh, FAST = 2, and its symmetric variant): a single
MaternP atom with p >= 1 on well-scaled Float64 points, r2 from the norm expansion, 6-instruction square root, clamp-free exp.
Reference semantics: mul!(y, G, x, alpha, beta) with k = MaternP(p) (src/gramian.jl:78-87, src/stationary.jl:134-158).  Tolerance 1e-12
against the oracle (which restates both branches of the reference evaluation), and agreement with the direct-difference kernel K1."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL64 = 1e-12


def _scalar(fn):
    os.environ["COVFN_MVM_SCALAR"] = "1"
    try:
        return fn()
    finally:
        del os.environ["COVFN_MVM_SCALAR"]


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("p", [1, 2, 3, 5])
def test_matern_fast_vs_oracle_and_direct_kernel(cf, O, d, p):
    rng = np.random.default_rng(7000 + 10 * d + p)
    n, m = 1500, 1111  # ragged against the 128-column tile and the row tiles
    X, Y = rng.standard_normal((n, d)), rng.standard_normal((m, d))
    a = rng.standard_normal(m)
    # (with l = 0.7 the host's cancellation bound may or may not admit the expansion, depending on d and p: parity must hold either way)
    for k, surely_fast in ((cf.MaternP(p), True), (2.5 * cf.Lengthscale(cf.MaternP(p), 0.7), False), (cf.Lengthscale(cf.MaternP(p), 3.0), True)):
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        b = G @ a
        ref = O.mul_vec(k.program(), X, a, Y=Y)
        assert relerr(b, ref) < TOL64, (d, p)
        bs = _scalar(lambda: G @ a)
        assert relerr(bs, ref) < TOL64
        if surely_fast:
            assert not np.array_equal(b, bs), "expected the norm-expansion kernel (different rounding)"
        assert relerr(b, bs) < 1e-13
        y0 = rng.standard_normal(n)
        y = y0.copy()
        cf.mul_(y, G, a, -0.7, 1.9)
        assert relerr(y, O.mul_vec(k.program(), X, a, Y=Y, alpha=-0.7, beta=1.9, y0=y0)) < TOL64


def test_matern_fast_coincident_and_nearly_coincident_points(cf, O):
    # r2 from the norm expansion is 0 or +-1e-16 here; the value must be 1 - O(r2) exactly as with direct differences
    rng = np.random.default_rng(71)
    n, d = 600, 3
    X = rng.standard_normal((n, d))
    X[100:200] = X[0:100]                       # exact duplicates
    X[200:300] = X[0:100] + 1e-9 * rng.standard_normal((100, d))  # r2 ~ 1e-18: far below the rounding error of the expansion
    X[300:400] = X[0:100] + 1e-5 * rng.standard_normal((100, d))
    a = rng.standard_normal(n)
    for p in (1, 2, 3):
        k = cf.MaternP(p)
        G = cf.gramian(k, X.T.copy())
        b = G @ a
        assert np.all(np.isfinite(b))
        assert relerr(b, O.mul_vec(k.program(), X, a)) < TOL64
        M = G.Matrix()
        assert abs(M[0, 100] - 1.0) < 1e-15 and abs(M[5, 5] - 1.0) < 1e-15


def test_matern_fast_falls_back_when_the_points_are_ill_scaled(cf, O):
    rng = np.random.default_rng(72)
    n, d = 500, 3
    a = rng.standard_normal(n)
    k = cf.MaternP(2)
    # far from the origin: the cancellation bound of the expansion fails -> direct differences (bit-identical to the scalar path)
    X = rng.standard_normal((n, d)) + 1.0e4
    G = cf.gramian(k, X.T.copy())
    b = G @ a
    assert relerr(b, O.mul_vec(k.program(), X, a)) < TOL64
    assert np.array_equal(b, _scalar(lambda: G @ a))
    # p = 0 (Exp) is not differentiable in r2 at 0: never on the expansion
    X = rng.standard_normal((n, d))
    Ge = cf.gramian(cf.MaternP(0), X.T.copy())
    be = Ge @ a
    assert np.array_equal(be, _scalar(lambda: Ge @ a))
    # a short length scale: exponent of the farthest pair beyond the clamp-free range -> direct differences
    Gs = cf.gramian(cf.Lengthscale(k, 0.01), X.T.copy())
    bs = Gs @ a
    assert relerr(bs, O.mul_vec(cf.Lengthscale(k, 0.01).program(), X, a)) < TOL64


@pytest.mark.parametrize("d,n", [(3, 5000), (8, 4000), (6, 3000)])
def test_matern_fast_symmetric_variant(cf, O, d, n):
    rng = np.random.default_rng(73 + d)
    X = rng.standard_normal((n, d)) / (1.0 if d <= 4 else np.sqrt(d))
    a = rng.standard_normal(n)
    for k in (cf.MaternP(2), cf.Lengthscale(cf.MaternP(1), 0.8)):
        G = cf.gramian(k, X.T.copy())
        G.set_symmetric(True)
        b1 = G @ a
        assert np.array_equal(b1, G @ a), "bit-reproducible"
        ref = O.mul_vec(k.program(), X, a)
        assert relerr(b1, ref) < TOL64
        G.set_symmetric(False)
        assert relerr(G @ a, ref) < TOL64
        bs = _scalar(lambda: (G.set_symmetric(True), G @ a)[1])
        assert relerr(bs, ref) < TOL64
        assert not np.array_equal(b1, bs)


def test_matern_fast_row_range_and_cg(cf, O):
    rng = np.random.default_rng(74)
    n, d = 2600, 3
    X = rng.standard_normal((n, d)) * 2
    a = rng.standard_normal(n)
    k = cf.MaternP(2)
    G = cf.gramian(k, X.T.copy())
    full = G @ a
    G.set_row_range(300, 1901)
    assert relerr(G @ a, full[300:1901]) < 1e-14
    G.set_row_range(0, n)
    y = rng.standard_normal(n)
    x, iters, res = (0.5 * cf.I(n) + G).solve(y)
    Kd = O.matrix(k.program(), X) + 0.5 * np.eye(n)
    assert np.linalg.norm(Kd @ x - y) < 1e-6 * np.linalg.norm(y)
