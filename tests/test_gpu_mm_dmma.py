"""Parity of the Float64 tensor-core (DMMA) multi-RHS kernel, csrc/gram_mm_dmma.cuh, against the oracle and against the
scalar kernel it replaces.  Reference semantics: mul!(B::AbstractMatrix, G::Gramian, A::AbstractMatrix, alpha, beta),
src/gramian.jl:89-99; shapes follow test/gramian.jl:65-72 (rectangular Gramian, several right-hand sides)."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL64 = 1e-12  # relative 2-norm, BASELINE.json north_star


def _kernels(cf):
    return {
        "config3": 0.5 * cf.RQ(2) + cf.Dot() ** 2,
        "eq": cf.EQ(),
        "matern_ls": cf.Lengthscale(cf.MaternP(2), 0.7),
        "exp_times_rq": cf.Exp() * cf.RQ(1.5),
        "poly_sum": 0.3 * cf.EQ() + (cf.Dot() + 1.0) ** 3 + 0.25,
    }


@pytest.mark.parametrize("d", [8, 12, 16, 24, 32])
def test_dmma_dims_ragged_rectangular(cf, O, d):
    # d = 8, 16, 24, 32 pad the point copies to stride d + 4, d = 12 keeps stride 12; n, m are not multiples of the tile sizes
    rng = np.random.default_rng(100 + d)
    n, m, p = 301, 517, 5
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    A = rng.standard_normal((m, p))
    for name, k in _kernels(cf).items():
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        B = G @ A
        ref = O.mul_mat(k.program(), X, A, Y=Y)
        assert relerr(B, ref) < TOL64, (d, name)


def test_dmma_matches_scalar_kernel_and_alpha_beta(cf, O):
    rng = np.random.default_rng(7)
    n, d, p = 700, 32, 70  # two passes of <= 64 columns, last one ragged
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    A = rng.standard_normal((n, p))
    k = 0.5 * cf.RQ(2) + cf.Dot() ** 2
    G = cf.gramian(k, X.T.copy())
    B0 = rng.standard_normal((n, p))
    ref = O.mul_mat(k.program(), X, A, alpha=0.3, beta=-1.1, B0=B0)
    B1 = np.asfortranarray(B0.copy())
    cf.mul_(B1, G, A, 0.3, -1.1)
    assert relerr(B1, ref) < TOL64
    os.environ["COVFN_MM_SCALAR"] = "1"
    try:
        B2 = np.asfortranarray(B0.copy())
        cf.mul_(B2, G, A, 0.3, -1.1)
    finally:
        del os.environ["COVFN_MM_SCALAR"]
    assert relerr(B2, ref) < TOL64
    assert relerr(B1, B2) < 1e-13  # same result up to summation order (tensor-core k-blocking vs sequential FMA chain)
    # beta = 0 must overwrite NaN-filled output (src/gramian.jl:90)
    B3 = np.asfortranarray(np.full((n, p), np.nan))
    cf.mul_(B3, G, A, 1.0, 0.0)
    assert np.isfinite(B3).all() and relerr(B3, O.mul_mat(k.program(), X, A)) < TOL64


def test_dmma_coincident_and_far_points(cf, O):
    # r^2 = |x|^2 + |y|^2 - 2 x.y can round to a tiny negative number for coincident points: it is clamped to 0, and the
    # result must stay within tolerance for kernels with a kink at r = 0 (Exp, MaternP(0))
    rng = np.random.default_rng(9)
    n, d, p = 256, 16, 4
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    X[1::2] = X[0::2]  # every point duplicated
    A = rng.standard_normal((n, p))
    for k in (cf.EQ(), cf.MaternP(2), cf.RQ(2)):
        G = cf.gramian(k, X.T.copy())
        assert relerr(G @ A, O.mul_mat(k.program(), X, A)) < TOL64


def test_dmma_row_range(cf, O):
    rng = np.random.default_rng(13)
    n, d, p = 1000, 24, 9
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    A = rng.standard_normal((n, p))
    k = cf.EQ() + 0.1 * cf.Dot()
    G = cf.gramian(k, X.T.copy())
    full = G @ A
    G.set_row_range(130, 777)
    part = G @ A
    assert part.shape == (647, p)
    assert relerr(part, full[130:777]) < 1e-14
    assert relerr(full, O.mul_mat(k.program(), X, A)) < TOL64


def test_ill_scaled_points_use_direct_differences(cf, O):
    # a large common offset makes the norm expansion cancel catastrophically; the library must detect it and keep the
    # (x - y)^2 form (scalar kernel), still within tolerance
    rng = np.random.default_rng(21)
    n, d, p = 300, 16, 3
    X = rng.standard_normal((n, d)) / np.sqrt(d) + 1.0e6
    A = rng.standard_normal((n, p))
    k = cf.EQ()
    G = cf.gramian(k, X.T.copy())
    assert relerr(G @ A, O.mul_mat(k.program(), X, A)) < TOL64


def test_runtime_specialised_program_matches_interpreter(cf, O):
    """COVFN_JIT=1 re-compiles the DMMA kernel with the program structure as compile-time constants (csrc/cf_jit.h).  The
    specialised and the interpreted evaluation must agree with the oracle and with each other for every structure."""
    rng = np.random.default_rng(31)
    n, m, d, p = 300, 260, 16, 7
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    Y = rng.standard_normal((m, d)) / np.sqrt(d)
    A = rng.standard_normal((m, p))
    ks = dict(_kernels(cf))
    ks["matern5_x_line"] = cf.MaternP(5) * (cf.Dot() + 0.5) + 2.0 * cf.RQ(3)
    ks["rq_real"] = cf.RQ(0.7) + cf.Exp()
    ks["square_of_sum"] = (cf.EQ() + cf.RQ(1)) ** 2
    before = cf.jit_stats()
    os.environ["COVFN_JIT"] = "1"
    try:
        for name, k in ks.items():
            G = cf.gramian(k, X.T.copy(), Y.T.copy())
            Bj = G @ A
            os.environ["COVFN_JIT"] = "0"
            Bi = G @ A
            os.environ["COVFN_JIT"] = "1"
            ref = O.mul_mat(k.program(), X, A, Y=Y)
            assert relerr(Bj, ref) < TOL64, name
            assert relerr(Bj, Bi) < 1e-13, name
        # same structure, different hyper-parameters: served from the cache, no new compilation
        mid = cf.jit_stats()
        k2 = 0.25 * cf.Lengthscale(cf.RQ(2), 1.7) + 3.0 * cf.Dot() ** 2
        G = cf.gramian(k2, X.T.copy(), Y.T.copy())
        assert relerr(G @ A, O.mul_mat(k2.program(), X, A, Y=Y)) < TOL64
        after = cf.jit_stats()
    finally:
        del os.environ["COVFN_JIT"]
    assert mid["failures"] == before["failures"], "run-time specialisation fell back to the interpreter"
    assert mid["compiled"] - before["compiled"] >= 1
    assert after["compiled"] == mid["compiled"] and after["cache_hits"] > mid["cache_hits"]


def test_sqrt_kernels_on_symmetric_gramians_keep_direct_differences(cf, O):
    """Exp = exp(-sqrt(r2)) has an infinite derivative in r2 at 0, so the norm expansion (error ~1e-16 in r2 on the diagonal of a
    symmetric Gramian) would cost 1e-8 in k: such programs must not use it.  MaternP(p >= 1) is smooth in r2 and may."""
    rng = np.random.default_rng(17)
    n, d, p = 400, 16, 6
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    X[7] = X[5] + 1e-9 * rng.standard_normal(d)  # a near-duplicate: r2 ~ 1e-17, far below the rounding error of the norm expansion
    A = rng.standard_normal((n, p))
    for k in (cf.Exp(), cf.MaternP(0) * cf.RQ(2), 0.3 * cf.Exp() + cf.EQ(), cf.MaternP(1), cf.MaternP(3)):
        G = cf.gramian(k, X.T.copy())
        B = G @ A
        ref = O.mul_mat(k.program(), X, A)
        assert relerr(B, ref) < TOL64
        # entry-wise: K_55 must be k(0) and K_75 = k(x_7, x_5) to rounding, not k(1e-8)
        a1 = np.zeros(n); a1[5] = 1.0
        col = G @ np.stack([a1, a1], axis=1)
        assert abs(col[5, 0] - k(X[5], X[5])) < 1e-13
        assert abs(col[7, 0] - k(X[7], X[5])) < 1e-13
