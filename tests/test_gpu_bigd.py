"""Point dimension d > 32 (tiled contraction kernels, csrc/bigd.cuh): value MVM, multi-RHS, gradient MVM (isotropic and
dot product), dense instantiation and CG, against the oracle.  Includes the README's gradient example shape
(d = n = 1024, README.md:231-245) at reduced n for the oracle and at full size through properties."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("d", [33, 48, 100, 257])
def test_value_mvm_bigd(cf, O, d):
    rng = np.random.default_rng(d)
    n, m = 130, 203
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    a = rng.standard_normal(m)
    for k in (cf.EQ(), cf.MaternP(2), cf.RQ(2), 0.5 * cf.RQ(2) + cf.Dot() ** 2, cf.Poly(3, 1.0), cf.Lengthscale(cf.EQ(), 0.5)):
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        assert relerr(G @ a, O.mul_vec(k.program(), X, a, Y=Y)) < 1e-12, repr(k)
        y0 = rng.standard_normal(n)
        y = y0.copy()
        cf.mul_(y, G, a, 0.7, -1.3)
        assert relerr(y, O.mul_vec(k.program(), X, a, Y=Y, alpha=0.7, beta=-1.3, y0=y0)) < 1e-12
    G = cf.gramian(cf.EQ(), X.T.copy(), Y.T.copy())
    A = rng.standard_normal((m, 3))
    assert relerr(G @ A, O.mul_mat(cf.EQ().program(), X, A, Y=Y)) < 1e-12
    assert relerr(G.Matrix(), O.matrix(cf.EQ().program(), X, Y)) < 1e-13
    assert abs(G[3, 5] - cf.EQ()(X[3], Y[5])) < 1e-14


@pytest.mark.parametrize("d", [40, 128, 300])
def test_gradient_mvm_bigd(cf, O, d):
    rng = np.random.default_rng(1000 + d)
    n, m = 70, 90
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    a = rng.standard_normal(m * d)
    for k, trait in ((cf.EQ(), "isotropic"), (cf.MaternP(2), "isotropic"), (0.5 * cf.EQ() + cf.RQ(2), "isotropic"),
                     (cf.Dot() ** 3, "dot")):
        G = cf.gramian(cf.GradientKernel(k), X.T.copy(), Y.T.copy())
        assert G.shape == (n * d, m * d)
        assert relerr(G @ a, O.derivative_mul(k.program(), X, a, Y=Y, trait=trait)) < 1e-12, repr(k)
        y0 = rng.standard_normal(n * d)
        y = y0.copy()
        cf.mul_(y, G, a, -0.4, 1.7)
        assert relerr(y, O.derivative_mul(k.program(), X, a, Y=Y, trait=trait, alpha=-0.4, beta=1.7, y0=y0)) < 1e-12


@pytest.mark.parametrize("d", [33, 70, 129])
def test_value_gradient_mvm_bigd(cf, O, d):
    # ValueGradientKernel ((d + 1)-blocks, entry 0 = value: src/gradient.jl:400-474) beyond 32 dimensions: bigd_jet_vg_kernel
    rng = np.random.default_rng(3000 + d)
    n, m = 90, 133
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    a = rng.standard_normal(m * (d + 1))
    for k, trait in ((cf.EQ(), "isotropic"), (cf.MaternP(2), "isotropic"), (0.5 * cf.EQ() + cf.RQ(2), "isotropic"), (cf.Dot() ** 3, "dot")):
        G = cf.gramian(cf.ValueGradientKernel(k), X.T.copy(), Y.T.copy())
        assert G.shape == (n * (d + 1), m * (d + 1))
        ref = O.derivative_mul(k.program(), X, a, Y=Y, trait=trait, value_gradient=True)
        assert relerr(G @ a, ref) < 1e-12, repr(k)
        y0 = rng.standard_normal(n * (d + 1))
        y = y0.copy()
        cf.mul_(y, G, a, -0.4, 1.7)
        assert relerr(y, -0.4 * ref + 1.7 * y0) < 1e-12
    # a square system solved on the device: (sigma^2 I + K_vg) x = b
    Xs = rng.standard_normal((40, d)) * 1.5 / np.sqrt(d)
    Gs = cf.gramian(cf.ValueGradientKernel(cf.EQ()), Xs.T.copy())
    N = 40 * (d + 1)
    b = rng.standard_normal(N)
    x, iters, res = (0.3 * cf.I(N) + Gs).solve(b, reltol=1e-10)
    assert relerr(Gs @ x + 0.3 * x, b) < 1e-8


def test_readme_gradient_example_shape(cf, O):
    # README.md:231-245: GradientKernel(MaternP(2)), d = n = 1024, a (n d) x (n d) = 1,048,576^2 operator
    n = d = 1024
    rng = np.random.default_rng(5)
    X = rng.standard_normal((n, d)) / np.sqrt(d)  # scaled so that the kernel is not numerically the identity
    a, c = rng.standard_normal(n * d), rng.standard_normal(n * d)
    k = cf.MaternP(2)
    G = cf.gramian(cf.GradientKernel(k), X.T)
    Ga, Gc = G @ a, G @ c
    rows = (17, 25)
    ref = O.derivative_mul(k.program(), X, a, rows=rows)
    assert relerr(Ga[rows[0] * d:rows[1] * d], ref) < 1e-12
    assert abs(float(c @ Ga) - float(a @ Gc)) <= 1e-11 * np.linalg.norm(c) * np.linalg.norm(Ga)
    # solve (README.md:255-258): G \ a by CG on the device
    x, iters, res = (1e-8 * cf.I(n * d) + G).solve(Ga, reltol=1e-9, maxiter=200)
    assert np.linalg.norm((G @ x) - Ga) / np.linalg.norm(Ga) < 1e-6


def test_float32_bigd_runs_on_the_float64_shadow(cf, O):
    # d > 32 kernels are Float64; a Float32 Gramian converts on the fly (tests/test_gpu_f32_operators.py has the full coverage)
    rng = np.random.default_rng(5)
    X = (rng.standard_normal((50, 40)) / np.sqrt(40)).astype(np.float32)
    a = rng.standard_normal(50).astype(np.float32)
    b = cf.gramian(cf.EQ(), X.T.copy()) @ a
    assert b.dtype == np.float32 and relerr(b, O.mul_vec(cf.EQ().program(), X, a, dtype=np.float32)) < 1e-5
