"""Parity of the Float32 multi-RHS kernel on the tensor cores in 3xTF32 split precision (csrc/gram_mm_tf32.cuh) against the
Float32 oracle and the Float64 truth.  Reference semantics: mul!(B::AbstractMatrix, G::Gramian{Float32}, A, alpha, beta),
src/gramian.jl:89-99.  Tolerance: relative 2-norm 1e-5 (BASELINE.json north_star, Float32)."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL32 = 1e-5


def _kernels(cf):
    return {
        "config3": 0.5 * cf.RQ(2) + cf.Dot() ** 2,
        "eq": cf.EQ(),
        "matern2_ls": cf.Lengthscale(cf.MaternP(2), 0.8),
        "rq_real_plus_const": cf.RQ(1.5) + 0.25,
        "eq_times_poly": cf.EQ() * (cf.Dot() + 1.0) ** 2,
    }


@pytest.mark.parametrize("d", [8, 12, 16, 24, 32])
def test_tf32_dims_ragged_rectangular(cf, O, d):
    rng = np.random.default_rng(500 + d)
    n, m, p = 301, 517, 5
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Y = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    A = rng.standard_normal((m, p)).astype(np.float32)
    for name, k in _kernels(cf).items():
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        B = G @ A
        assert B.dtype == np.float32
        truth = O.mul_mat(k.program(), X.astype(np.float64), A.astype(np.float64), Y=Y.astype(np.float64))
        assert relerr(B.astype(np.float64), truth) < TOL32, (d, name)
        ref32 = O.mul_mat(k.program(), X, A, Y=Y, dtype=np.float32)
        assert relerr(B, ref32) < TOL32, (d, name)


def test_tf32_long_sums_do_not_drift_and_alpha_beta(cf, O):
    # many column tiles: the tensor core truncates when it adds to the accumulator, so the kernel sums per-tile results with
    # round-to-nearest adds; the error must stay at the Float32 rounding level for a sum over 40000 columns
    rng = np.random.default_rng(61)
    n, m, d, p = 257, 40000, 16, 70  # two passes of <= 64 right-hand sides
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Y = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    A = np.abs(rng.standard_normal((m, p))).astype(np.float32)  # same-sign terms: a truncation bias would show
    k = cf.EQ()
    G = cf.gramian(k, X.T.copy(), Y.T.copy())
    B0 = rng.standard_normal((n, p)).astype(np.float32)
    B = np.asfortranarray(B0.copy())
    cf.mul_(B, G, A, 0.5, -2.0)
    truth = O.mul_mat(k.program(), X.astype(np.float64), A.astype(np.float64), Y=Y.astype(np.float64), alpha=0.5, beta=-2.0,
                      B0=B0.astype(np.float64))
    assert relerr(B.astype(np.float64), truth) < 2e-6
    os.environ["COVFN_MM_SCALAR"] = "1"
    try:
        Bs = np.asfortranarray(B0.copy())
        cf.mul_(Bs, G, A, 0.5, -2.0)
    finally:
        del os.environ["COVFN_MM_SCALAR"]
    assert relerr(Bs.astype(np.float64), truth) < TOL32
    assert not np.array_equal(B, Bs), "expected the tensor-core kernel"


def test_tf32_symmetric_row_range_and_nan_overwrite(cf, O):
    rng = np.random.default_rng(62)
    n, d, p = 1000, 32, 9
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    A = rng.standard_normal((n, p)).astype(np.float32)
    k = 0.5 * cf.RQ(2) + cf.Dot() ** 2
    G = cf.gramian(k, X.T.copy())
    B = np.asfortranarray(np.full((n, p), np.nan, dtype=np.float32))
    cf.mul_(B, G, A, 1.0, 0.0)  # beta == 0 overwrites (src/gramian.jl:90)
    truth = O.mul_mat(k.program(), X.astype(np.float64), A.astype(np.float64))
    assert np.isfinite(B).all() and relerr(B.astype(np.float64), truth) < TOL32
    G.set_row_range(130, 777)
    part = G @ A
    assert part.shape == (647, p) and relerr(part, B[130:777]) < 1e-6


def test_tf32_runtime_specialised_program(cf, O):
    """COVFN_JIT=1: the Float32 evaluator of the program is generated too (csrc/cf_jit.h, part 2 of cf_jit_shape.h)"""
    rng = np.random.default_rng(63)
    n, m, d, p = 300, 260, 16, 7
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Y = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    A = rng.standard_normal((m, p)).astype(np.float32)
    before = cf.jit_stats()
    os.environ["COVFN_JIT"] = "1"
    try:
        for name, k in _kernels(cf).items():
            G = cf.gramian(k, X.T.copy(), Y.T.copy())
            Bj = G @ A
            os.environ["COVFN_JIT"] = "0"
            Bi = G @ A
            os.environ["COVFN_JIT"] = "1"
            truth = O.mul_mat(k.program(), X.astype(np.float64), A.astype(np.float64), Y=Y.astype(np.float64))
            assert relerr(Bj.astype(np.float64), truth) < TOL32, name
            assert relerr(Bj, Bi) < 2e-6, name
    finally:
        del os.environ["COVFN_JIT"]
    after = cf.jit_stats()
    assert after["failures"] == before["failures"]
    assert after["compiled"] + after["cache_hits"] > before["compiled"] + before["cache_hits"]


def test_tcgen05_and_legacy_kernels_agree(cf, O):
    """The default Float32 multi-RHS kernel is the tcgen05 / TMEM one (csrc/gram_mm_tc5.cuh); COVFN_MM_LEGACY=1 selects the mma.sync
    kernel (csrc/gram_mm_tf32.cuh).  Both against the oracle, and against each other at the 3xTF32 error level."""
    import os
    rng = np.random.default_rng(77)
    n, m, d, p = 700, 1300, 16, 70  # ragged tiles, two passes of 64 right-hand sides
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Y = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    A = rng.standard_normal((m, p)).astype(np.float32)
    k = 0.5 * cf.RQ(2) + cf.Dot() ** 2
    ref = O.mul_mat(k.program(), X, A, Y=Y, dtype=np.float32)
    out = {}
    for legacy in ("0", "1"):
        os.environ["COVFN_MM_LEGACY"] = legacy
        try:
            out[legacy] = cf.gramian(k, X.T.copy(), Y.T.copy()) @ A
        finally:
            del os.environ["COVFN_MM_LEGACY"]
        assert relerr(out[legacy], ref) < 1e-5
    assert relerr(out["0"], out["1"]) < 2e-6
