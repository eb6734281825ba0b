"""Parity of the Float32 value MVM with the pair dot products on tcgen05 / TMEM in 3xTF32 (csrc/gram_mvm_tc5.cuh, padded D >= 8,
well-scaled points) against the Float64 truth, the scalar Float32 kernel and its mma.sync predecessor (csrc/gram_mvm_tf32.cuh).
Reference semantics: mul!(y::AbstractVector, G::Gramian{Float32}, x::AbstractVector, alpha, beta), src/gramian.jl:78-87.  Tolerance 1e-5."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL32 = 1e-5


def _with_env(name, fn):
    os.environ[name] = "1"
    try:
        return fn()
    finally:
        del os.environ[name]


def _kernels(cf):
    return {
        "eq": cf.EQ(),
        "eq_l": 1.5 * cf.Lengthscale(cf.EQ(), 0.8),
        "matern2": cf.MaternP(2),
        "matern1": cf.MaternP(1),
        "rq2": cf.RQ(2),
        "config3": 0.5 * cf.RQ(2) + cf.Dot() ** 2,
        "eq_times_matern": 1.5 * cf.EQ() * cf.MaternP(1),
    }


@pytest.mark.parametrize("d", [8, 11, 16, 24, 32])
def test_mvm_tc5_dims_ragged_rectangular(cf, O, d):
    rng = np.random.default_rng(1600 + d)
    n, m = 333, 1061
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Y = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    a = rng.standard_normal(m).astype(np.float32)
    for name, k in _kernels(cf).items():
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        b = G @ a
        assert b.dtype == np.float32
        truth = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64), Y=Y.astype(np.float64))
        assert relerr(b.astype(np.float64), truth) < TOL32, (d, name)
        legacy = _with_env("COVFN_MVM_LEGACY", lambda: G @ a)
        assert relerr(legacy.astype(np.float64), truth) < TOL32, (d, name)
        assert not np.array_equal(b, legacy), "expected two different kernels (tcgen05 and mma.sync)"
        assert relerr(b, legacy) < 3e-6, (d, name)


def test_mvm_tc5_many_tiles_per_cta_and_long_same_sign_sums(cf, O):
    # 300 row tiles x 3-4 column chunks of ~30 tiles each: the stage ring and the four TMEM buffers wrap several times
    rng = np.random.default_rng(1665)
    n, m, d = 38400 + 77, 6400 + 13, 16
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Y = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    a = np.abs(rng.standard_normal(m)).astype(np.float32)  # same-sign terms: a truncation bias would show
    for k in (cf.EQ(), cf.MaternP(2)):
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        b = G @ a
        truth = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64), Y=Y.astype(np.float64))
        assert relerr(b.astype(np.float64), truth) < 2e-6
        assert np.array_equal(b, G @ a)  # run-to-run bit-identical


def test_mvm_tc5_alpha_beta_unaligned_and_row_range(cf, O):
    rng = np.random.default_rng(1666)
    n, d = 6000, 32
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    a = rng.standard_normal(n).astype(np.float32)
    k = cf.EQ()
    G = cf.gramian(k, X.T.copy())
    b0 = rng.standard_normal(n).astype(np.float32)
    b = np.full(n, np.nan, dtype=np.float32)
    cf.mul_(b, G, a, 1.0, 0.0)  # beta == 0 overwrites NaNs (src/gramian.jl:80)
    full = b.copy()
    truth = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64))
    assert relerr(full.astype(np.float64), truth) < 2e-6
    b = b0.copy()
    cf.mul_(b, G, a, 0.3, -1.1)
    truth2 = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64), alpha=0.3, beta=-1.1, y0=b0.astype(np.float64))
    assert relerr(b.astype(np.float64), truth2) < 2e-6
    buf = np.zeros(n + 1, dtype=np.float32)
    a_un = buf[1:]
    a_un[:] = a
    assert relerr(G @ a_un, full) < 1e-6
    G.set_row_range(700, 2300)
    part = G @ a
    assert part.shape == (1600,) and relerr(part, full[700:2300]) < 1e-6


def test_mvm_tc5_runtime_specialised_program(cf, O):
    # composite program, product large enough for the run-time specialisation (COVFN_JIT=1 forces it): same kernel source, program compiled in
    rng = np.random.default_rng(1667)
    n, d = 3000, 16
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    a = rng.standard_normal(n).astype(np.float32)
    k = 0.5 * cf.RQ(2) + cf.Dot() ** 2
    G = cf.gramian(k, X.T.copy())
    truth = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64))
    plain = G @ a
    jit = _with_env("COVFN_JIT", lambda: cf.gramian(k, X.T.copy()) @ a)
    assert relerr(plain.astype(np.float64), truth) < TOL32
    assert relerr(jit.astype(np.float64), truth) < TOL32
