"""Accuracy of the device arithmetic (csrc/cf_math.cuh) measured in ulps against extended precision: the table-based FP64
exp (cf_exp_cv), the SFU-seeded sqrt and reciprocal, through single Gramian entries K[i, 0] = k((t_i - 0)^2) on 1-D points.
The reference uses Julia Base exp / sqrt / ^ (<= 1 ulp).  Dense instantiation (Matrix!, getindex) uses the two-step
reduction: <= 2 ulp for every argument.  The MVM hot loop uses the one-step reduction: <= (2 + 0.35 |arg|) ulp, i.e. an
absolute error <= 4e-17 relative to k = 1 (the same size as the rounding error the squared distance itself carries)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ulps(got, want_ld):
    want = want_ld.astype(np.float64)
    spacing = np.spacing(np.abs(want))
    return np.abs((got.astype(np.longdouble) - want_ld) / spacing.astype(np.longdouble)).astype(np.float64)


def entries(cf, k, t):
    """k(t_i^2) for every t_i via the dense instantiation kernel (same atom code as the MVM kernels)"""
    G = cf.gramian(k, t.reshape(1, -1), np.zeros((1, 1)))
    return G.Matrix()[:, 0]


def test_exp_eq_ulp(cf):
    rng = np.random.default_rng(0)
    t = np.concatenate([rng.uniform(0, 37.0, 40000), rng.uniform(0, 1e-3, 2000), [0.0, 1e-200, 37.2, 37.4]])
    r2 = (t * t).astype(np.longdouble)  # the device forms r2 = t*t in double; exp is then exact in long double
    want = np.exp(-r2 / 2)
    got = entries(cf, cf.EQ(), t)
    ok = want > 1e-300
    u = ulps(got[ok], want[ok])
    assert u.max() <= 2.0, u.max()
    assert u.mean() < 0.6
    assert np.all(got[~ok] < 1e-290)  # below the clamp: flushed towards zero, never garbage
    # length scales change the exponent constants (cf_lower.h fill_exp): same accuracy
    for l in (0.1, 3.7):
        want_l = np.exp(-r2 / (2 * np.longdouble(l) ** 2))
        got_l = entries(cf, cf.Lengthscale(cf.EQ(), l), t)
        okl = want_l > 1e-300
        assert ulps(got_l[okl], want_l[okl]).max() <= 2.5  # r2 / l^2 is folded into one rounded constant


def test_sqrt_exp_matern_ulp(cf):
    # exp(-sqrt(r2)): any double-precision evaluation rounds the square root first, and that half ulp in the argument
    # alone becomes up to 0.5 |arg| ulp in the result (the reference's exp(-sqrt(r2)) has it too); bound 2.5 + 0.75 |arg|
    rng = np.random.default_rng(1)
    t = np.concatenate([rng.uniform(0, 300.0, 30000), rng.uniform(0, 1e-6, 2000), [0.0]])
    r2 = (t * t).astype(np.longdouble)
    r = np.sqrt(r2)
    got = entries(cf, cf.Exp(), t)
    assert np.all(ulps(got, np.exp(-r)) <= 2.5 + 0.75 * r.astype(np.float64))
    # against the argument a correctly rounded double sqrt produces: <= 2.5 ulp except where the device sqrt is 1 ulp off
    r_d = np.sqrt((t * t)).astype(np.longdouble)
    assert np.quantile(ulps(got, np.exp(-r_d)), 0.99) <= 2.5
    s = np.sqrt(5 * r2)
    want = (1 + s + s * s / 3) * np.exp(-s)
    got = entries(cf, cf.MaternP(2), t)
    ok = want > 1e-290
    assert np.all(ulps(got[ok], want[ok]) <= 4.0 + 0.75 * s[ok].astype(np.float64))


def test_rcp_rq_ulp(cf):
    rng = np.random.default_rng(2)
    t = rng.uniform(0, 1e3, 30000)
    r2 = (t * t).astype(np.longdouble)
    got = entries(cf, cf.RQ(1), t)
    assert ulps(got, 1 / (1 + r2 / 2)).max() <= 2.0
    # alpha = 3: the rounding of the base (1 ulp) is tripled by the power in ANY double evaluation, plus two products and
    # the reciprocal
    got = entries(cf, cf.RQ(3), t)
    assert ulps(got, (1 + r2 / 6) ** -3).max() <= 6.5


def test_mvm_fast_exp_error_model(cf):
    # column 0 of K through the MVM kernel (a = e_0): the one-step reduction's error grows linearly with |argument|
    rng = np.random.default_rng(3)
    t = np.concatenate([[0.0], rng.uniform(0, 37.0, 20000)])
    G = cf.gramian(cf.EQ(), t.reshape(1, -1))
    e0 = np.zeros(t.size)
    e0[0] = 1.0
    col = G @ e0  # = exp(-(t_i - 0)^2 / 2)
    r2 = (t * t).astype(np.longdouble)
    want = np.exp(-r2 / 2)
    ok = want > 1e-300
    u = ulps(col[ok], want[ok])
    arg = (r2[ok] / 2).astype(np.float64)
    assert np.all(u <= 2.0 + 0.35 * arg)
    assert np.abs(col[ok].astype(np.longdouble) - want[ok]).max() <= 2.3e-16  # absolute error: at most one ulp of k = 1
