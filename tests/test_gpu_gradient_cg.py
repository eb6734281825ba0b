"""GPU parity for the isotropic GradientKernel MVM (reference src/gramian.jl:241-253, src/gradient.jl:86-92,
tests test/gradient.jl:16-63) and the on-device CG solve (src/lazy_linear_algebra.jl:126-144)."""
import zlib

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def iso_kernels(cf):
    return {
        "EQ": cf.EQ(),
        "MaternP(2)": cf.MaternP(2),
        "MaternP(3)": cf.MaternP(3),
        "MaternP(1)": cf.MaternP(1),
        "RQ(2)": cf.RQ(2),
        "RQ(1.5)": cf.RQ(1.5),
        "0.7*EQ": 0.7 * cf.EQ(),
        "Lengthscale(EQ,0.6)": cf.Lengthscale(cf.EQ(), 0.6),
        "Lengthscale(MaternP(2),2.0)": cf.Lengthscale(cf.MaternP(2), 2.0),
        "EQ+RQ(2)": cf.EQ() + cf.RQ(2),
        "EQ*MaternP(2)": cf.EQ() * cf.MaternP(2),
    }


@pytest.mark.parametrize("name", list(iso_kernels(__import__("covfn_b200")).keys()))
@pytest.mark.parametrize("d", [1, 2, 5, 16])
def test_gradient_mvm_vs_oracle(cf, O, name, d):
    k = iso_kernels(cf)[name]
    rng = np.random.default_rng(zlib.crc32(f"{name}-{d}".encode()))
    n, m = 70, 131
    X = rng.standard_normal((n, d)) / np.sqrt(d)  # test/gradient.jl:18
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    a = rng.standard_normal(m * d)
    G = cf.gramian(cf.GradientKernel(k), X.T.copy(), Y.T.copy())
    assert G.shape == (n * d, m * d)  # test/gradient.jl:35
    b = G @ a
    ref = O.gradient_mul(k.program(), X, a, Y=Y)
    assert relerr(b, ref) < 1e-12
    # 5-argument mul! with random alpha, beta against the dense matrix (test/gradient.jl:47-52)
    alpha, beta = rng.standard_normal(2)
    y0 = rng.standard_normal(n * d)
    y = y0.copy()
    cf.mul_(y, G, a, alpha, beta)
    M = O.gradient_matrix(k.program(), X, Y)
    assert relerr(y, alpha * (M @ a) + beta * y0) < 1e-12


def test_gradient_symmetric_diagonal_blocks(cf, O):
    # x === y: diagonal blocks take k'(0), k''(0) -- the MaternP Taylor branch (src/stationary.jl:139-146)
    rng = np.random.default_rng(21)
    n, d = 90, 5
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    a = rng.standard_normal(n * d)
    for k in (cf.EQ(), cf.MaternP(1), cf.MaternP(2), cf.MaternP(3), cf.RQ(2)):
        G = cf.gramian(cf.GradientKernel(k), X.T.copy())
        assert relerr(G @ a, O.gradient_mul(k.program(), X, a)) < 1e-12, repr(k)
    # Exp = MaternP(0) is not differentiable at r = 0 and has no Taylor branch (taylor_bound = eps^(1/0) = 0, src/stationary.jl:137):
    # the reference's ForwardDiff derivatives at r2 = 0 are -Inf / Inf and its diagonal blocks -2 (k' a + 2 k'' r (r.a)) are NaN
    # (Inf * 0, src/gradient.jl:86-92) -- so the whole product is NaN.  Same here and in the oracle; off the diagonal it is finite.
    for k in (cf.Exp(), cf.MaternP(0)):
        G = cf.gramian(cf.GradientKernel(k), X.T.copy())
        ref = O.gradient_mul(k.program(), X, a)
        got = G @ a
        assert np.all(np.isnan(ref)) and np.all(np.isnan(got)), repr(k)
        Y = rng.standard_normal((70, d)) / np.sqrt(d)
        ay = rng.standard_normal(70 * d)
        Gxy = cf.gramian(cf.GradientKernel(k), X.T.copy(), Y.T.copy())
        assert relerr(Gxy @ ay, O.gradient_mul(k.program(), X, ay, Y=Y)) < 1e-12, repr(k)


def test_gradient_config4_shape(cf, O):
    # BASELINE config 4 shape at reduced n: GradientKernel(EQ), d = 16
    rng = np.random.default_rng(22)
    n, d = 2048, 16
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    a = rng.standard_normal(n * d)
    k = cf.EQ()
    G = cf.gramian(cf.GradientKernel(k), X.T.copy())
    b = G @ a
    rows = (100, 356)
    ref = O.gradient_mul(k.program(), X, a, rows=rows)
    assert relerr(b[rows[0] * d:rows[1] * d], ref) < 1e-12


def test_gradient_rejects_generic_input(cf):
    # EQ + Dot has the GenericInput trait (src/properties.jl:57-58): the reference uses a dense ForwardDiff fallback
    X = np.random.default_rng(0).standard_normal((3, 10))
    G = cf.gramian(cf.GradientKernel(cf.EQ() + cf.Dot()), X)
    with pytest.raises(cf.UnsupportedKernel):
        G @ np.ones(30)


DOT_KERNELS = ["Dot^3", "Dot", "Poly(3,0.5)", "Dot^2+2*Dot^3"]


def _dot_kernel(cf, name):
    return {"Dot^3": cf.Dot() ** 3, "Dot": cf.Dot(), "Poly(3,0.5)": cf.Poly(3, 0.5),
            "Dot^2+2*Dot^3": cf.Dot() ** 2 + 2 * cf.Dot() ** 3}[name]


@pytest.mark.parametrize("name", DOT_KERNELS)
@pytest.mark.parametrize("d", [1, 5, 16])
def test_dot_product_gradient_mvm(cf, O, name, d):
    # DotProductGradientKernelElement (src/gradient.jl:107-115), kernels of test/gradient.jl:21
    k = _dot_kernel(cf, name)
    rng = np.random.default_rng(zlib.crc32(f"dot-{name}-{d}".encode()))
    n, m = 40, 67
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    a = rng.standard_normal(m * d)
    G = cf.gramian(cf.GradientKernel(k), X.T.copy(), Y.T.copy())
    assert relerr(G @ a, O.derivative_mul(k.program(), X, a, Y=Y, trait="dot")) < 1e-12
    alpha, beta = rng.standard_normal(2)
    y0 = rng.standard_normal(n * d)
    y = y0.copy()
    cf.mul_(y, G, a, alpha, beta)
    M = O.derivative_matrix(k.program(), X, Y, trait="dot")
    assert relerr(y, alpha * (M @ a) + beta * y0) < 1e-12  # test/gradient.jl:47-52


@pytest.mark.parametrize("name", ["MaternP(3)", "EQ", "RQ(1.0)", "Dot^3", "0.5*EQ+MaternP(2)"])
@pytest.mark.parametrize("d", [2, 5, 16])
def test_value_gradient_kernel_mvm(cf, O, name, d):
    # ValueGradientKernel (src/gradient.jl:400-474; test/gradient.jl:87-137): (d+1) x (d+1) blocks
    k = {"MaternP(3)": cf.MaternP(3), "EQ": cf.EQ(), "RQ(1.0)": cf.RQ(1.0), "Dot^3": cf.Dot() ** 3,
         "0.5*EQ+MaternP(2)": 0.5 * cf.EQ() + cf.MaternP(2)}[name]
    trait = "dot" if name.startswith("Dot") else "isotropic"
    rng = np.random.default_rng(zlib.crc32(f"vg-{name}-{d}".encode()))
    n = 37
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    a, b = rng.standard_normal((d + 1) * n), rng.standard_normal((d + 1) * n)
    G = cf.gramian(cf.ValueGradientKernel(k), X.T.copy())
    assert G.shape == ((d + 1) * n, (d + 1) * n)  # test/gradient.jl:103
    MK = O.derivative_matrix(k.program(), X, trait=trait, value_gradient=True)
    assert np.abs(MK - MK.T).max() < 1e4 * np.finfo(float).eps  # test/gradient.jl:96
    alpha, beta = rng.standard_normal(2)
    Kab = b.copy()
    cf.mul_(Kab, G, a, alpha, beta)
    assert relerr(Kab, alpha * (MK @ a) + beta * b) < 1e-12  # test/gradient.jl:119-123
    assert relerr(G @ a, O.derivative_mul(k.program(), X, a, trait=trait, value_gradient=True)) < 1e-12
    # rectangular
    Y = rng.standard_normal((53, d)) / np.sqrt(d)
    a2 = rng.standard_normal(53 * (d + 1))
    G2 = cf.gramian(cf.ValueGradientKernel(k), X.T.copy(), Y.T.copy())
    assert relerr(G2 @ a2, O.derivative_mul(k.program(), X, a2, Y=Y, trait=trait, value_gradient=True)) < 1e-12


def test_value_gradient_solve(cf, O):
    # test/gradient.jl:127-136: K \ (K a) residual < 1e-6 for EQ and RQ(1.)
    rng = np.random.default_rng(77)
    n, d = 12, 3
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    for k in (cf.EQ(), cf.RQ(1.0)):
        G = cf.gramian(cf.ValueGradientKernel(k), X.T.copy())
        a = rng.standard_normal(n * (d + 1))
        Ka = G @ a
        x, iters, res = (1e-10 * cf.I(n * (d + 1)) + G).solve(Ka, reltol=1e-10, maxiter=5000)
        assert np.linalg.norm((G @ x) - Ka) / np.linalg.norm(Ka) < 1e-6


def test_cg_solve(cf, O):
    # config 5 at reduced size: (K + sigma2 I) \ y with MaternP(2), d = 8
    rng = np.random.default_rng(31)
    n, d = 1500, 8
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    y = rng.standard_normal(n)
    k = cf.MaternP(2)
    sigma2 = 1e-2
    G = cf.gramian(k, X.T.copy())
    A = sigma2 * cf.I(n) + G
    assert isinstance(A, cf.LazyMatrixSum)  # test/gramian.jl:51-53
    x, iters, res = A.solve(y)
    xo, ito, reso, hist = O.cg_solve(k.program(), X, y, sigma2)
    assert abs(iters - ito) <= max(3, 0.05 * ito)  # same algorithm; rounding shifts the stopping iteration slightly
    assert relerr(x, xo) < 1e-6  # CG amplifies rounding by the condition number; both satisfy the residual bound
    r = y - (A @ x)
    assert np.linalg.norm(r) / np.linalg.norm(y) < 1e-6  # test/gradient.jl:62
    # operator product itself
    v = rng.standard_normal(n)
    assert relerr(A @ v, sigma2 * v + O.mul_vec(k.program(), X, v)) < 1e-12


def test_cg_gradient_operator(cf, O):
    rng = np.random.default_rng(32)
    n, d = 60, 4
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    k = cf.MaternP(3)
    G = cf.gramian(cf.GradientKernel(k), X.T.copy())
    a = rng.standard_normal(n * d)
    Ka = G @ a
    A = 1e-8 * cf.I(n * d) + G
    x, iters, res = A.solve(Ka, reltol=1e-10, maxiter=2000)
    assert np.linalg.norm((G @ x) - Ka) / np.linalg.norm(Ka) < 1e-6  # test/gradient.jl:56-63
