"""Parity of the Float64 tensor-core value MVM (csrc/gram_mvm_dmma.cuh, padded D >= 8, well-scaled points) against the oracle
and against the scalar kernel K1 it replaces.  Reference semantics: mul!(y::AbstractVector, G::Gramian, x::AbstractVector,
alpha, beta), src/gramian.jl:78-87; shapes follow test/gramian.jl:56-72 (rectangular Gramian, lazy vs dense product)."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL64 = 1e-12  # relative 2-norm, BASELINE.json north_star


def _kernels(cf):
    return {
        "eq": cf.EQ(),
        "eq_ls": 1.7 * cf.Lengthscale(cf.EQ(), 0.6),
        "matern2": cf.MaternP(2),
        "matern5": cf.MaternP(5),
        "rq2": cf.RQ(2),
        "rq_real": cf.RQ(1.3),
        "config3": 0.5 * cf.RQ(2) + cf.Dot() ** 2,
        "poly": (cf.Dot() + 1.0) ** 3,
        "exp": cf.Exp(),  # exp(-sqrt(r2)): must stay on the direct-difference kernel
    }


def _scalar(fn):
    os.environ["COVFN_MVM_SCALAR"] = "1"
    try:
        return fn()
    finally:
        del os.environ["COVFN_MVM_SCALAR"]


@pytest.mark.parametrize("d", [8, 11, 16, 24, 32])
def test_mvm_dmma_dims_ragged_rectangular(cf, O, d):
    rng = np.random.default_rng(200 + d)
    n, m = 333, 1061  # neither a multiple of the 128-row / 32-column tiles
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    a = rng.standard_normal(m)
    for name, k in _kernels(cf).items():
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        b = G @ a
        ref = O.mul_vec(k.program(), X, a, Y=Y)
        assert relerr(b, ref) < TOL64, (d, name)
        bs = _scalar(lambda: G @ a)
        assert relerr(b, bs) < 1e-13, (d, name)
        if name == "exp":
            assert np.array_equal(b, bs), "exp(-sqrt(r2)) programs must not use the norm expansion"
        elif name == "eq":
            assert not np.array_equal(b, bs), "expected the tensor-core kernel (different summation order than K1)"


def test_mvm_dmma_alpha_beta_rows_and_unaligned_weights(cf, O):
    rng = np.random.default_rng(5)
    n, d = 3000, 16  # several column chunks per row tile -> partial buffer + reduction
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    k = cf.MaternP(2)
    G = cf.gramian(k, X.T.copy())
    a = rng.standard_normal(n)
    b0 = rng.standard_normal(n)
    b = b0.copy()
    cf.mul_(b, G, a, 0.3, -1.1)
    assert relerr(b, O.mul_vec(k.program(), X, a, alpha=0.3, beta=-1.1, y0=b0)) < TOL64
    bn = np.full(n, np.nan)
    cf.mul_(bn, G, a, 1.0, 0.0)  # beta == 0 overwrites NaN (src/gramian.jl:80)
    full = O.mul_vec(k.program(), X, a)
    assert np.isfinite(bn).all() and relerr(bn, full) < TOL64
    # weights at an address that is 8- but not 16-byte aligned: no TMA, cooperative loads
    buf = np.zeros(n + 1)
    a_un = buf[1:]
    a_un[:] = a
    assert relerr(G @ a_un, full) < TOL64
    # row block of the operator
    G.set_row_range(700, 2300)
    part = G @ a
    assert part.shape == (1600,) and relerr(part, full[700:2300]) < 1e-14


def test_mvm_dmma_duplicates_far_points_and_determinism(cf, O):
    rng = np.random.default_rng(8)
    n, d = 1024, 32
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    X[1::2] = X[0::2]          # exact duplicates: r2 must clamp to 0, not go negative
    X[100] = 0.3               # an outlier (|x|^2 = 2.9) that still passes the scale check of the norm expansion
    a = rng.standard_normal(n)
    for k in (cf.EQ(), cf.MaternP(1), cf.MaternP(2), cf.RQ(2)):
        G = cf.gramian(k, X.T.copy())
        b1 = G @ a
        b2 = G @ a
        assert np.array_equal(b1, b2)  # run-to-run bit-identical
        assert relerr(b1, O.mul_vec(k.program(), X, a)) < TOL64


def test_cg_on_tensor_core_operator_matches_scalar(cf, O):
    # BASELINE config 5's operator (MaternP(2), d = 8) at small n: the CG solution through K1d equals the one through K1
    rng = np.random.default_rng(12)
    n, d = 2048, 8
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    y = rng.standard_normal(n)
    k = cf.MaternP(2)
    G = cf.gramian(k, X.T.copy())
    x1, it1, res1 = (1e-2 * cf.I(n) + G).solve(y, reltol=1e-10)
    x2, it2, res2 = _scalar(lambda: (1e-2 * cf.I(n) + cf.gramian(k, X.T.copy())).solve(y, reltol=1e-10))
    assert relerr(x1, x2) < 1e-8 and abs(it1 - it2) <= max(3, it2 // 20)
    K = O.matrix(k.program(), X, X)
    assert relerr(K @ x1 + 1e-2 * x1, y) < 1e-8


def test_short_length_scales_disable_the_norm_expansion(cf, O):
    """|dk/dr2| grows like 1/l^2: with l = 0.02 an absolute error of 1e-15 in r2 would be 1e-12 in k, so the library must keep
    direct differences (bit-identical to the scalar kernel) -- and stay within tolerance either way"""
    rng = np.random.default_rng(19)
    n, d = 600, 16
    X = rng.standard_normal((n, d)) / np.sqrt(d) * 0.05  # clustered points so that k is not all zeros at l = 0.02
    a = rng.standard_normal(n)
    A = rng.standard_normal((n, 3))
    for k in (cf.Lengthscale(cf.EQ(), 0.02), cf.Lengthscale(cf.MaternP(2), 0.02), cf.Lengthscale(cf.RQ(2), 0.02)):
        G = cf.gramian(k, (X / 0.05).T.copy())  # unit-scale points, short length scale
        b = G @ a
        assert np.array_equal(b, _scalar(lambda: G @ a))
        assert relerr(b, O.mul_vec(k.program(), X / 0.05, a)) < TOL64
        G2 = cf.gramian(k, X.T.copy())          # points scaled with the length scale: the expansion is safe again
        assert relerr(G2 @ a, O.mul_vec(k.program(), X, a)) < TOL64
        assert relerr(G2 @ A, O.mul_mat(k.program(), X, A)) < TOL64


@pytest.mark.parametrize("shape", [(1, 1), (1, 40), (5, 3), (129, 33)])
def test_tensor_core_kernels_tiny_shapes(cf, O, shape):
    """single point, fewer columns than one tile, one row past a tile boundary: every tensor-core kernel family
    (value MVM, multi-RHS, gradient, value-gradient; Float64 and Float32)"""
    n, m = shape
    d = 16
    rng = np.random.default_rng(1000 + 7 * n + m)
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    a = rng.standard_normal(m)
    A = rng.standard_normal((m, 3))
    k = cf.MaternP(2)
    os.environ["COVFN_GRAD_DMMA"] = "1"
    try:
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        assert relerr(G @ a, O.mul_vec(k.program(), X, a, Y=Y)) < TOL64
        assert relerr(G @ A, O.mul_mat(k.program(), X, A, Y=Y)) < TOL64
        ag = rng.standard_normal(m * d)
        Gg = cf.gramian(cf.GradientKernel(k), X.T.copy(), Y.T.copy())
        assert relerr(Gg @ ag, O.gradient_mul(k.program(), X, ag, Y=Y)) < TOL64
        av = rng.standard_normal(m * (d + 1))
        Gv = cf.gramian(cf.ValueGradientKernel(k), X.T.copy(), Y.T.copy())
        assert relerr(Gv @ av, O.derivative_mul(k.program(), X, av, Y=Y, trait="isotropic", value_gradient=True)) < TOL64
    finally:
        del os.environ["COVFN_GRAD_DMMA"]
    Xf, Yf = X.astype(np.float32), Y.astype(np.float32)
    Gf = cf.gramian(k, Xf.T.copy(), Yf.T.copy())
    t64 = O.mul_vec(k.program(), Xf.astype(np.float64), a.astype(np.float32).astype(np.float64), Y=Yf.astype(np.float64))
    assert relerr((Gf @ a.astype(np.float32)).astype(np.float64), t64) < 1e-5
    T64 = O.mul_mat(k.program(), Xf.astype(np.float64), A.astype(np.float32).astype(np.float64), Y=Yf.astype(np.float64))
    assert relerr((Gf @ A.astype(np.float32)).astype(np.float64), T64) < 1e-5
