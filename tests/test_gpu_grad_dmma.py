"""Parity of the Float64 tensor-core isotropic gradient-kernel MVM (csrc/grad_mvm_dmma.cuh, padded D in {8, 12, 16, 24, 32},
well-scaled points) against the oracle and against the scalar kernel K5 it replaces.  Reference semantics:
blockmul!(y, G::Gramian, x, alpha, beta) src/gramian.jl:241-253 with mul!(b, ::IsotropicGradientKernelElement, a, alpha, beta)
src/gradient.jl:86-92; the shapes follow test/gradient.jl:29-52 (lazy operator vs dense matrix, 5-argument mul!)."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL64 = 1e-12


@pytest.fixture(autouse=True)
def _force_tensor_core_kernel():
    # the library only selects the tensor-core kernel above 2^22 blocks per call; the parity tests use small sizes
    os.environ["COVFN_GRAD_DMMA"] = "1"
    yield
    del os.environ["COVFN_GRAD_DMMA"]


def _scalar(fn):
    os.environ["COVFN_GRAD_SCALAR"] = "1"
    try:
        return fn()
    finally:
        del os.environ["COVFN_GRAD_SCALAR"]


def _kernels(cf):
    return {
        "eq": cf.EQ(),
        "eq_ls": 0.7 * cf.Lengthscale(cf.EQ(), 1.3),
        "matern2": cf.MaternP(2),
        "matern3": cf.MaternP(3),
        "rq2": cf.RQ(2),
        "rq_real": cf.RQ(1.7),
        "eq_plus_rq": cf.EQ() + 0.5 * cf.RQ(2),
        "eq_times_matern": cf.EQ() * cf.MaternP(2),
        "matern1": cf.MaternP(1),  # k'' has a 1/sqrt(r2) term: must stay on the direct-difference kernel
    }


@pytest.mark.parametrize("d", [8, 11, 12, 16, 24, 32])
def test_grad_dmma_dims_ragged_rectangular(cf, O, d):
    rng = np.random.default_rng(300 + d)
    n, m = 203, 301  # neither a multiple of the 128-row / 32-column tiles
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    a = rng.standard_normal(m * d)
    for name, k in _kernels(cf).items():
        G = cf.gramian(cf.GradientKernel(k), X.T.copy(), Y.T.copy())
        b = G @ a
        ref = O.gradient_mul(k.program(), X, a, Y=Y)
        assert relerr(b, ref) < TOL64, (d, name)
        bs = _scalar(lambda: G @ a)
        assert relerr(b, bs) < 1e-13, (d, name)
        if name == "matern1":
            assert np.array_equal(b, bs), "MaternP(1) gradient operators must not use the norm expansion"
        elif name == "eq":
            assert not np.array_equal(b, bs), "expected the tensor-core kernel"


def test_grad_dmma_alpha_beta_rows_duplicates(cf, O):
    rng = np.random.default_rng(33)
    n, d = 1500, 16  # several column chunks -> partial buffer + reduction
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    X[1::2] = X[0::2]  # exact duplicates: r2 and r.a must vanish exactly, no spurious 1e-16 / 1e-16
    k = cf.MaternP(2)
    G = cf.gramian(cf.GradientKernel(k), X.T.copy())
    a = rng.standard_normal(n * d)
    b0 = rng.standard_normal(n * d)
    b = b0.copy()
    cf.mul_(b, G, a, 0.3, -1.1)
    assert relerr(b, O.gradient_mul(k.program(), X, a, alpha=0.3, beta=-1.1, y0=b0)) < TOL64
    full = G @ a
    assert np.array_equal(full, G @ a)  # run-to-run bit-identical
    G.set_row_range(100, 900)
    part = G @ a
    assert part.shape == (800 * d,) and relerr(part, full[100 * d:900 * d]) < 1e-14


def test_grad_dmma_dense_operator_and_symmetry(cf, O):
    # test/gradient.jl:35-45: the lazy operator equals the dense (n d) x (n d) matrix, which is symmetric
    rng = np.random.default_rng(35)
    n, d = 40, 8
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    k = cf.EQ()
    G = cf.gramian(cf.GradientKernel(k), X.T.copy())
    M = O.gradient_matrix(k.program(), X)
    E = np.eye(n * d)
    cols = np.stack([G @ E[:, j] for j in range(0, n * d, 37)], axis=1)
    assert relerr(cols, M[:, ::37]) < TOL64
    u, v = rng.standard_normal(n * d), rng.standard_normal(n * d)
    assert abs(u @ (G @ v) - v @ (G @ u)) < 1e-12 * np.linalg.norm(u) * np.linalg.norm(v) * np.linalg.norm(M, 2)


@pytest.mark.parametrize("d", [8, 10, 16, 32])
def test_value_gradient_dmma(cf, O, d):
    """ValueGradientKernel blocks (d+1) x (d+1) (reference src/gradient.jl:400-474) on the tensor-core kernel"""
    rng = np.random.default_rng(400 + d)
    n, m = 150, 333
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    a = rng.standard_normal(m * (d + 1))
    for name, k in _kernels(cf).items():
        G = cf.gramian(cf.ValueGradientKernel(k), X.T.copy(), Y.T.copy())
        b = G @ a
        ref = O.derivative_mul(k.program(), X, a, Y=Y, trait="isotropic", value_gradient=True)
        assert relerr(b, ref) < TOL64, (d, name)
        bs = _scalar(lambda: G @ a)
        assert relerr(b, bs) < 1e-13, (d, name)
        if name == "eq":
            assert not np.array_equal(b, bs), "expected the tensor-core kernel"
    # alpha / beta and a symmetric operator with duplicates
    Xs = X.copy()
    Xs[1::2] = Xs[0::2]
    k = cf.MaternP(2)
    G = cf.gramian(cf.ValueGradientKernel(k), Xs.T.copy())
    v = rng.standard_normal(n * (d + 1))
    b0 = rng.standard_normal(n * (d + 1))
    b = b0.copy()
    cf.mul_(b, G, v, -0.7, 0.4)
    assert relerr(b, O.derivative_mul(k.program(), Xs, v, trait="isotropic", value_gradient=True, alpha=-0.7, beta=0.4, y0=b0)) < TOL64


@pytest.mark.parametrize("d", [8, 12, 16, 32])
def test_dot_product_programs_on_tensor_cores(cf, O, d):
    """DotProductInput kernels (reference src/gradient.jl:109-115): t = x.y and s = x.a_g are GEMMs as they are, nothing cancels,
    so the tensor-core kernel is used whatever the scale of the points"""
    rng = np.random.default_rng(700 + d)
    n, m = 170, 290
    X = rng.standard_normal((n, d))  # deliberately not scaled by 1/sqrt(d)
    Y = rng.standard_normal((m, d))
    for name, k in {"dot2": cf.Dot() ** 2, "poly3": (cf.Dot() + 1.0) ** 3, "line_plus_dot2": 0.5 * cf.Dot() ** 2 + cf.Dot() + 0.3}.items():
        for vg in (False, True):
            K = cf.ValueGradientKernel(k) if vg else cf.GradientKernel(k)
            bs = d + (1 if vg else 0)
            a = rng.standard_normal(m * bs)
            G = cf.gramian(K, X.T.copy(), Y.T.copy())
            b = G @ a
            ref = O.derivative_mul(k.program(), X, a, Y=Y, trait="dot", value_gradient=vg)
            assert relerr(b, ref) < TOL64, (d, name, vg)
            bsc = _scalar(lambda: G @ a)
            assert relerr(b, bsc) < 1e-13, (d, name, vg)
            if name == "dot2" and not vg:
                assert not np.array_equal(b, bsc), "expected the tensor-core kernel"


def test_runtime_specialised_jets(cf, O):
    """COVFN_JIT=1: composite derivative programs get their product-rule jets generated (csrc/cf_jit.h part 3) and compiled
    into the tensor-core gradient kernel; specialised == interpreted == oracle, GradientKernel and ValueGradientKernel"""
    rng = np.random.default_rng(801)
    n, m, d = 210, 330, 16
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    programs = {
        "eq_plus_rq": cf.EQ() + 0.5 * cf.RQ(2),
        "eq_times_matern3": 0.7 * cf.EQ() * cf.MaternP(3),
        "square_of_sum": (cf.Lengthscale(cf.EQ(), 1.4) + cf.RQ(1)) ** 2,
        "rq_real_plus_matern": cf.RQ(1.7) + cf.MaternP(2),
    }
    before = cf.jit_stats()
    for name, k in programs.items():
        for vg in (False, True):
            K = cf.ValueGradientKernel(k) if vg else cf.GradientKernel(k)
            bs = d + (1 if vg else 0)
            a = rng.standard_normal(m * bs)
            G = cf.gramian(K, X.T.copy(), Y.T.copy())
            os.environ["COVFN_JIT"] = "1"
            try:
                bj = G @ a
            finally:
                os.environ["COVFN_JIT"] = "0"
            try:
                bi = G @ a
            finally:
                del os.environ["COVFN_JIT"]
            ref = O.derivative_mul(k.program(), X, a, Y=Y, trait="isotropic", value_gradient=vg)
            assert relerr(bj, ref) < TOL64, (name, vg)
            assert relerr(bj, bi) < 1e-13, (name, vg)
    after = cf.jit_stats()
    assert after["failures"] == before["failures"]
    assert after["compiled"] - before["compiled"] >= len(programs)
