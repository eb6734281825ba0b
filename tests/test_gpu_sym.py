"""Symmetric variant of the value MVM (csrc/gram_mvm_sym.cuh): each unordered pair of y === x evaluated once.  It is the default
for Float64 symmetric Gramians with n >= 32768; results must match the oracle like every other path AND be bit-reproducible
(single-writer partial sums, fixed-order combine: no floating-point atomics)."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,d,n", [("EQ", 3, 40000), ("EQ", 2, 33001), ("MaternP(2)", 3, 36000), ("RQ(2)", 4, 32768),
                                      ("0.5*EQ+MaternP(1)", 3, 33333), ("EQ_l0.5", 6, 34000), ("MaternP(2)", 8, 33000)])
def test_symmetric_matches_oracle_and_plain_path(cf, O, name, d, n):
    rng = np.random.default_rng(d * 1000 + n)
    k = {"EQ": cf.EQ(), "MaternP(2)": cf.MaternP(2), "RQ(2)": cf.RQ(2), "0.5*EQ+MaternP(1)": 0.5 * cf.EQ() + cf.MaternP(1),
         "EQ_l0.5": cf.Lengthscale(cf.EQ(), 0.5)}[name]
    X = rng.standard_normal((n, d)) / (1.0 if d <= 4 else np.sqrt(d))
    a = rng.standard_normal(n)
    G = cf.gramian(k, X.T.copy())
    G.set_symmetric(True)
    b1 = G @ a
    b2 = G @ a
    assert np.array_equal(b1, b2), "symmetric variant must be bit-reproducible"
    y0 = rng.standard_normal(n)
    y = y0.copy()
    cf.mul_(y, G, a, -0.5, 2.0)
    G.set_symmetric(False)
    bp = G @ a
    assert relerr(b1, bp) < 1e-13
    assert relerr(y, -0.5 * bp + 2.0 * y0) < 1e-13
    for rows in ((0, 64), (n // 2 - 32, n // 2 + 32), (n - 64, n)):
        ref = O.mul_vec(k.program(), X, a, rows=rows)
        assert relerr(b1[rows[0]:rows[1]], ref) < 1e-12


def test_symmetric_is_default_and_can_be_disabled(cf):
    rng = np.random.default_rng(5)
    n, d = 65536, 3
    X = rng.standard_normal((n, d))
    a = rng.standard_normal(n)
    G = cf.gramian(cf.EQ(), X.T.copy())
    b = G @ a
    t_default = G.last_timing()[0]
    G.set_symmetric(False)
    bp = G @ a
    t_plain = G.last_timing()[0]
    assert relerr(b, bp) < 1e-13
    assert t_default < 0.8 * t_plain, (t_default, t_plain)  # half the evaluations
    # a fresh process-wide switch: COVFN_SYMMETRIC=0 at create time
    os.environ["COVFN_SYMMETRIC"] = "0"
    try:
        G2 = cf.gramian(cf.EQ(), X.T.copy())
        b2 = G2 @ a
        assert np.array_equal(b2, bp)
    finally:
        del os.environ["COVFN_SYMMETRIC"]
