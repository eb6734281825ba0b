"""Parity of the Float32 value MVM with the distance GEMM in 3xTF32 on the tensor cores (csrc/gram_mvm_tf32.cuh, padded
D >= 8, well-scaled points) against the Float64 truth, the Float32 oracle and the scalar Float32 kernel.  Reference semantics:
mul!(y::AbstractVector, G::Gramian{Float32}, x::AbstractVector, alpha, beta), src/gramian.jl:78-87.  Tolerance 1e-5."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL32 = 1e-5


@pytest.fixture(autouse=True)
def _legacy_kernel():
    """This file covers the mma.sync kernel, which the tcgen05 one (tests/test_gpu_mvm_tc5.py) has replaced as the default."""
    os.environ["COVFN_MVM_LEGACY"] = "1"
    yield
    del os.environ["COVFN_MVM_LEGACY"]


def _scalar(fn):
    os.environ["COVFN_MVM_SCALAR"] = "1"
    try:
        return fn()
    finally:
        del os.environ["COVFN_MVM_SCALAR"]


def _kernels(cf):
    return {
        "eq": cf.EQ(),
        "matern2": cf.MaternP(2),
        "rq2": cf.RQ(2),
        "config3": 0.5 * cf.RQ(2) + cf.Dot() ** 2,
        "eq_times_matern": 1.5 * cf.EQ() * cf.MaternP(1),
        "exp": cf.Exp(),  # stays on the direct-difference kernel
    }


@pytest.mark.parametrize("d", [8, 11, 16, 24, 32])
def test_mvm_tf32_dims_ragged_rectangular(cf, O, d):
    rng = np.random.default_rng(600 + d)
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
        bs = _scalar(lambda: G @ a)
        assert relerr(bs.astype(np.float64), truth) < TOL32, (d, name)
        if name == "exp":  # both are direct-difference kernels (d = 8: the packed-FP32 one, tests/test_gpu_f32p.py)
            assert relerr(b, bs) < 2e-6
        elif name == "eq":
            assert not np.array_equal(b, bs), "expected the tensor-core kernel"


def test_mvm_tf32_alpha_beta_rows_unaligned_and_long_sums(cf, O):
    rng = np.random.default_rng(65)
    n, d = 6000, 16
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    a = np.abs(rng.standard_normal(n)).astype(np.float32)  # same-sign terms: a truncation bias would show
    k = cf.MaternP(2)
    G = cf.gramian(k, X.T.copy())
    b0 = rng.standard_normal(n).astype(np.float32)
    b = b0.copy()
    cf.mul_(b, G, a, 0.3, -1.1)
    truth = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64), alpha=0.3, beta=-1.1, y0=b0.astype(np.float64))
    assert relerr(b.astype(np.float64), truth) < 2e-6
    full = G @ a
    assert np.array_equal(full, G @ a)  # run-to-run bit-identical
    buf = np.zeros(n + 1, dtype=np.float32)
    a_un = buf[1:]
    a_un[:] = a  # 4-byte aligned only: no TMA for the weights
    assert relerr(G @ a_un, full) < 1e-6
    G.set_row_range(700, 2300)
    part = G @ a
    assert part.shape == (1600,) and relerr(part, full[700:2300]) < 1e-6
