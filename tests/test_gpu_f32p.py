"""Parity of the packed-FP32 Float32 value MVM (csrc/gram_mvm_f32p.cuh: single isotropic atoms at padded D <= 8, direct differences,
FADD2 / FMUL2 / FFMA2 over column pairs) against the Float64 truth, the Float32 oracle and the scalar Float32 kernel.
Reference semantics: mul!(y::AbstractVector, G::Gramian{Float32}, x::AbstractVector, alpha, beta), src/gramian.jl:78-87.  Tolerance 1e-5."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL32 = 1e-5


def _scalar(fn):
    os.environ["COVFN_MVM_SCALAR"] = "1"
    try:
        return fn()
    finally:
        del os.environ["COVFN_MVM_SCALAR"]


def _kernels(cf):
    return {
        "eq": cf.EQ(),
        "eq_l": 2.5 * cf.Lengthscale(cf.EQ(), 0.6),
        "exp": cf.Exp(),
        "matern1": cf.MaternP(1),
        "matern2": cf.MaternP(2),
        "matern3_l": cf.Lengthscale(cf.MaternP(3), 1.7),
        "rq2": cf.RQ(2),
        "rq1_l": cf.Lengthscale(cf.RQ(1), 0.8),
    }


@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7, 8])
def test_packed_dims_ragged_rectangular(cf, O, d):
    rng = np.random.default_rng(900 + d)
    n, m = 1100, 1333  # ragged against the 128-column tile and the 1024-row tile; odd m: the last packed pair is half padding
    X = (rng.standard_normal((n, d)) * 1.5).astype(np.float32)  # (spread: the well-scaled check fails at d = 8, so d = 8 stays here too)
    Y = (rng.standard_normal((m, d)) * 1.5).astype(np.float32)
    a = rng.standard_normal(m).astype(np.float32)
    for name, k in _kernels(cf).items():
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        b = G @ a
        assert b.dtype == np.float32
        truth = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64), Y=Y.astype(np.float64))
        assert relerr(b.astype(np.float64), truth) < TOL32, (d, name)
        assert relerr(b, O.mul_vec(k.program(), X, a, Y=Y, dtype=np.float32)) < TOL32, (d, name)
        bs = _scalar(lambda: G @ a)
        assert relerr(bs.astype(np.float64), truth) < TOL32, (d, name)
        assert not np.array_equal(b, bs), "expected the packed kernel (different summation order)"


def test_packed_alpha_beta_unaligned_row_range_and_long_sums(cf, O):
    rng = np.random.default_rng(77)
    n, d = 9000, 3
    X = rng.standard_normal((n, d)).astype(np.float32)
    a = np.abs(rng.standard_normal(n)).astype(np.float32)  # same-sign terms: a summation bias would show
    k = cf.EQ()
    G = cf.gramian(k, X.T.copy())
    b0 = rng.standard_normal(n).astype(np.float32)
    b = np.full(n, np.nan, dtype=np.float32)
    cf.mul_(b, G, a, 1.0, 0.0)  # beta == 0 overwrites NaNs (src/gramian.jl:80)
    truth = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64))
    assert relerr(b.astype(np.float64), truth) < 2e-6
    b = b0.copy()
    cf.mul_(b, G, a, 0.3, -1.1)
    truth2 = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64), alpha=0.3, beta=-1.1, y0=b0.astype(np.float64))
    assert relerr(b.astype(np.float64), truth2) < 2e-6
    full = G @ a
    assert np.array_equal(full, G @ a)  # run-to-run bit-identical
    buf = np.zeros(n + 1, dtype=np.float32)
    a_un = buf[1:]
    a_un[:] = a
    assert relerr(G @ a_un, full) < 1e-6
    G.set_row_range(700, 2300)
    part = G @ a
    assert part.shape == (1600,) and relerr(part, full[700:2300]) < 1e-6


def test_packed_direct_differences_survive_a_large_offset(cf, O):
    # points far from the origin: r2 from norms would lose everything in Float32, direct differences do not
    rng = np.random.default_rng(78)
    n, d = 700, 3
    X = (rng.standard_normal((n, d)) + 300.0).astype(np.float32)
    a = rng.standard_normal(n).astype(np.float32)
    for k in (cf.EQ(), cf.MaternP(2)):
        G = cf.gramian(k, X.T.copy())
        truth = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64))
        assert relerr((G @ a).astype(np.float64), truth) < TOL32


def test_packed_multi_rhs_free_function_and_entries_unchanged(cf, O):
    # the neighbours of the path keep working on a handle that has built the transposed copy: multi-RHS product and entries
    rng = np.random.default_rng(79)
    n, d = 400, 2
    X = rng.standard_normal((n, d)).astype(np.float32)
    a = rng.standard_normal(n).astype(np.float32)
    A = np.asfortranarray(rng.standard_normal((n, 3)).astype(np.float32))
    k = cf.MaternP(2)
    G = cf.gramian(k, X.T.copy())
    b = G @ a
    B = G @ A
    truth = O.mul_mat(k.program(), X.astype(np.float64), A.astype(np.float64))
    assert relerr(B.astype(np.float64), truth) < TOL32
    assert abs(G[3, 7] - O.matrix(k.program(), X.astype(np.float64))[3, 7]) < 1e-6
    assert relerr(b, G @ a) == 0.0
