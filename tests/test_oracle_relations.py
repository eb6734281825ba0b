"""Pins the CPU oracle (oracle/covfn_oracle.c) on the relations the reference's own tests assert for this path.

The reference stores no golden vectors and seeds no RNG: its tests are relations between two
computations on the same random draw.  Each test below reproduces one of them on the oracle (file:line cited), plus
checks against an independent evaluator (the Python mirror's scalar formulas, mpmath, sympy, scipy's Bessel-K Matern).
"""
import math

import numpy as np
import pytest

from conftest import relerr

OP_EQ, OP_EXP, OP_RQ, OP_MATERNP, OP_DOT, OP_CONST, OP_SUM, OP_PROD, OP_POW, OP_LS = range(1, 11)


def all_kernels(cf):
    return [cf.EQ(), cf.Exp(), cf.RQ(2), cf.RQ(1.0), cf.RQ(0.7), cf.MaternP(0), cf.MaternP(1), cf.MaternP(2), cf.MaternP(3),
            cf.MaternP(8), cf.Dot(), cf.Dot() ** 3, cf.Poly(3, 0.5), 0.5 * cf.RQ(2) + cf.Dot() ** 2, cf.EQ() * cf.Exp(),
            cf.Lengthscale(cf.EQ(), 0.3), cf.Lengthscale(cf.MaternP(2), 2.5), (cf.EQ() + 1) ** 2, 3 * cf.EQ()]


def test_lazy_vs_dense_rectangular(cf, O):
    # test/gramian.jl:56-63 (n x 2n, vector) and :65-72 (p = 3 columns)
    rng = np.random.default_rng(1)
    n = 8
    for d in (1, 3):
        x, y = rng.standard_normal((n, d)), rng.standard_normal((2 * n, d))
        a, A = rng.standard_normal(2 * n), rng.standard_normal((2 * n, 3))
        for k in all_kernels(cf):
            M = O.matrix(k.program(), x, y)
            assert M.shape == (n, 2 * n)
            assert np.allclose(O.mul_vec(k.program(), x, a, Y=y), M @ a, rtol=1e-12, atol=1e-13)
            assert np.allclose(O.mul_mat(k.program(), x, A, Y=y), M @ A, rtol=1e-12, atol=1e-13)
            assert np.allclose(O.mul_mat(k.program(), x, A, Y=y, fused=True), M @ A, rtol=1e-12, atol=1e-13)


def test_entries_match_independent_evaluator(cf, O):
    # test/gramian.jl:17-21,75-80: G[i,j] ~ k(x[i], y[j]); the independent evaluator is the mirror's numpy formula
    rng = np.random.default_rng(2)
    x, y = rng.standard_normal((6, 3)), rng.standard_normal((7, 3))
    for k in all_kernels(cf):
        for i in range(6):
            for j in range(7):
                assert math.isclose(O.getindex(k.program(), x, i, y, j), k(x[i], y[j]), rel_tol=1e-13, abs_tol=1e-300), repr(k)
                assert math.isclose(O.getindex(k.program(), x, i, y, j), O.truth_getindex(k.program(), x[i], y[j]),
                                    rel_tol=1e-13, abs_tol=1e-300), repr(k)


def test_isotropic_two_argument_convention(cf, O):
    # test/stationary.jl:39-42: k(x1, x2) ~ k(r2)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((2, 4))
    r2 = float(np.sum((x[0] - x[1]) ** 2))
    for k in (cf.EQ(), cf.Exp(), cf.RQ(1.0), cf.MaternP(2), cf.Lengthscale(cf.EQ(), 0.4)):
        assert math.isclose(O.getindex(k.program(), x, 0, x, 1), k(r2), rel_tol=1e-14)
        assert math.isclose(O.value_derivative_laplacian(k.program(), r2)[0], k(r2), rel_tol=1e-14)


def naive_maternp(r2, p):
    # reference src/stationary.jl:162-169 (the control implementation the reference tests against)
    r = math.sqrt((2 * p + 1) * r2)
    val = sum((math.factorial(p + i) // (math.factorial(p - i) * math.factorial(i))) * (2 * r) ** (p - i) for i in range(p + 1))
    return val * math.exp(-r) / (math.factorial(2 * p) // math.factorial(p))


def test_maternp_against_control_implementation(cf, O):
    # test/stationary.jl:60-69: p = 0 and p = 2, 3 on r2 = 10^(1:16) * eps and at zero
    eps = np.finfo(float).eps
    r2s = [10.0 ** e * eps for e in range(1, 17)]
    for p in (0, 2, 3):
        prog = cf.MaternP(p).program()
        x0 = np.zeros((1, 1))
        assert O.getindex(prog, x0, 0, x0, 0) == pytest.approx(naive_maternp(0.0, p), rel=1e-15)
        for r2 in r2s:
            x = np.array([[0.0], [math.sqrt(r2)]])
            got = O.getindex(prog, x, 0, x, 1)
            assert got == pytest.approx(naive_maternp(float(x[1, 0] ** 2), p), rel=1.5e-8)  # Julia's default isapprox rtol
            assert got == pytest.approx(cf.MaternP(p)(float(x[1, 0] ** 2)), rel=1e-14)


def test_maternp_coefficients_and_taylor_derivatives(cf):
    # src/stationary.jl:184-191: p = 2 -> [12, 6]; :172-182 via symbolic differentiation (SymEngine there, sympy here)
    import sympy as sp

    from covfn_b200.kernels import MaternP_coefficients, MaternP_derivatives_at_zero

    assert MaternP_coefficients(2) == [12.0, 6.0]
    assert MaternP_derivatives_at_zero(2) == pytest.approx([-5 / 6, 25 / 12])
    s = sp.symbols("s", positive=True)  # s = r2
    for p in range(1, 7):
        r = sp.sqrt((2 * p + 1) * s)
        kr = sum(sp.Integer(math.factorial(p + i)) / (math.factorial(p - i) * math.factorial(i)) * (2 * r) ** (p - i)
                 for i in range(p + 1)) * sp.exp(-r) / sp.Integer(math.factorial(2 * p) // math.factorial(p))
        ser = sp.series(kr, s, 0, p + 1).removeO()
        want = [float(sp.factorial(i) * ser.coeff(s, i)) for i in range(1, p + 1)]
        assert MaternP_derivatives_at_zero(p) == pytest.approx(want, rel=1e-13)


def test_maternp_derivatives_against_bessel_matern(cf, O):
    # test/stationary.jl:70-82: dk/dr2 and d2k/dr2^2 of MaternP(p) vs the general Matern(z, p + 1/2) of
    # src/stationary.jl:98-114 (Bessel-K closed form, with ITS OWN Taylor branch below eps^(1/2) for nu > 2),
    # compared like the reference does: isapprox(vector, vector, atol = 1e-6 / 1e-5), i.e. in the 2-norm
    import mpmath as mp

    mp.mp.dps = 40
    eps = np.finfo(float).eps

    def matern_derivs(z, nu):
        bound = eps ** 0.5 if nu > 2 else (eps if nu > 1 else 0.0)
        if z < bound:  # y = 1 + nu/(2(1-nu)) z + nu^2/(8(2-3nu+nu^2)) z^2
            c1 = nu / (2 * (1 - nu))
            c2 = nu**2 / (8 * (2 - 3 * nu + nu**2)) if nu > 2 else 0.0
            return c1 + 2 * c2 * z, 2 * c2
        f = lambda t: 2 ** (1 - mp.mpf(nu)) / mp.gamma(nu) * mp.sqrt(2 * nu * t) ** nu * mp.besselk(nu, mp.sqrt(2 * nu * t))
        return float(mp.diff(f, mp.mpf(z), 1)), float(mp.diff(f, mp.mpf(z), 2))

    for p in (2, 3):
        prog = cf.MaternP(p).program()
        e1, e2 = [], []
        for e in range(1, 17):
            r2 = 10.0**e * eps
            _, k1, k2 = O.value_derivative_laplacian(prog, r2)
            d1, d2 = matern_derivs(r2, p + 0.5)
            e1.append(k1 - d1)
            e2.append(k2 - d2)
        assert np.linalg.norm(e1) < 1e-6
        assert np.linalg.norm(e2) < 1e-5


def test_dot_poly_and_algebra(cf, O):
    # test/mercer.jl:12-19 and test/algebra.jl:28-51
    rng = np.random.default_rng(4)
    x, y = rng.standard_normal((1, 3)), rng.standard_normal((1, 3))

    def ev(k):
        return O.getindex(k.program(), x, 0, y, 0)

    assert ev(cf.Dot() ** 3) == pytest.approx(ev(cf.Poly(3)), rel=1e-14)
    assert ev(cf.Dot()) == pytest.approx(float(x[0] @ y[0]), rel=1e-15)
    k1, k2, k3 = cf.EQ(), cf.RQ(1.0), cf.Dot()
    for ka, kb in ((k1, k2), (k1, k3)):
        assert ev(ka + kb) == pytest.approx(ev(ka) + ev(kb), rel=1e-15)
        assert ev(kb + ka) == pytest.approx(ev(ka) + ev(kb), rel=1e-15)
        assert ev(ka * kb) == pytest.approx(ev(ka) * ev(kb), rel=1e-15)
        assert ev(kb * ka) == pytest.approx(ev(ka) * ev(kb), rel=1e-15)
    a = math.exp(rng.standard_normal())
    assert ev(a * k1) == pytest.approx(a * ev(k1), rel=1e-15)
    assert ev(k1 * a) == pytest.approx(a * ev(k1), rel=1e-15)
    assert ev(a + k1) == pytest.approx(a + ev(k1), rel=1e-15)
    assert ev(k1 + a) == pytest.approx(a + ev(k1), rel=1e-15)
    for p in range(1, 5):
        for k in (k1, k2, k3):
            assert ev(k**p) == pytest.approx(ev(k) ** p, rel=1e-14)
    # README.md:78-87 composite (config 3): kernel(x,y) ~ smooth(x,y)/2 + line(x,y)^2
    assert ev(0.5 * cf.RQ(2) + cf.Dot() ** 2) == pytest.approx(ev(cf.RQ(2)) / 2 + ev(cf.Dot()) ** 2, rel=1e-15)


def test_gramians_are_covariances(cf, O):
    # test/stationary.jl:43-49,100-116: iscov(Sigma, tol) -- symmetric and eigenvalues >= -tol
    rng = np.random.default_rng(5)
    for d in (1, 2, 3):
        x = rng.standard_normal((16, d))
        for k in (cf.EQ(), cf.Exp(), cf.RQ(1.0), cf.MaternP(0), cf.MaternP(2), cf.MaternP(8), cf.Lengthscale(cf.EQ(), 0.5)):
            S = O.matrix(k.program(), x)
            assert np.array_equal(S, S.T)
            assert np.linalg.eigvalsh(S).min() > -1e-12


def test_lengthscale_semantics(cf, O):
    # test/stationary.jl:120-130: Lengthscale(k, l)(r) ~ k(|r|^2 / l^2)
    rng = np.random.default_rng(6)
    l = math.exp(rng.standard_normal())
    for d in (1, 2, 3):
        x = rng.standard_normal((2, d))
        r2 = float(np.sum((x[0] - x[1]) ** 2))
        for k in (cf.EQ(), cf.Exp(), cf.RQ(1.0), cf.MaternP(2)):
            assert O.getindex(cf.Lengthscale(k, l).program(), x, 0, x, 1) == pytest.approx(k(r2 / l**2), rel=1e-13)


def test_ard_semantics(cf, O):
    """test/stationary.jl:132-154: ARD(k, l) with l = 1 is k; n2(x - y) = sum((x - y)^2 / l); kl(x, y) ~ k(x .* w, y .* w) with
    w = sqrt(1 / l).  The oracle's ARD node restates enorm2(Diagonal(inv.(l)), difference(x, y)) (src/transformation.jl:38-45)."""
    rng = np.random.default_rng(66)
    for d in (2, 3, 5):
        x = rng.standard_normal((2, d))
        k = cf.EQ()
        assert O.getindex(cf.ARD(k, np.ones(d)).program(), x, 0, x, 1) == pytest.approx(O.getindex(k.program(), x, 0, x, 1), rel=1e-15)
        c = 2.0
        r2 = float(np.sum((x[0] - x[1]) ** 2))
        assert O.getindex(cf.ARD(k, c**2 * np.ones(d)).program(), x, 0, x, 1) == pytest.approx(k(r2 / c**2), rel=1e-14)
        l = np.exp(rng.standard_normal(d))
        w = np.sqrt(1 / l)
        for k in (cf.EQ(), cf.Exp(), cf.RQ(1.0), cf.MaternP(2), 0.5 * cf.EQ() + cf.MaternP(1) * cf.RQ(2)):
            kl = cf.ARD(k, l)
            got = O.getindex(kl.program(), x, 0, x, 1)
            assert got == pytest.approx(O.getindex(k.program(), x * w, 0, x * w, 1), rel=1e-13)
            assert got == pytest.approx(kl(x[0], x[1]), rel=1e-13)                       # the mirror's own formula
            assert got == pytest.approx(O.truth_getindex(kl.program(), x[0], x[1]), rel=1e-13)
        # a constant outside the ARD node, and a Lengthscale inside it
        kk = 3.0 * cf.ARD(cf.Lengthscale(cf.MaternP(2), 0.7), l)
        s2 = float(np.sum((x[0] - x[1]) ** 2 / l))
        assert O.getindex(kk.program(), x, 0, x, 1) == pytest.approx(3.0 * cf.MaternP(2)(s2 / 0.49), rel=1e-13)
    # lazy vs dense (test/gramian.jl:56-72) with the ARD kernel
    X = rng.standard_normal((40, 3))
    a = rng.standard_normal(40)
    kl = cf.ARD(cf.MaternP(2), [0.5, 2.0, 1.1])
    assert np.allclose(O.mul_vec(kl.program(), X, a), O.matrix(kl.program(), X) @ a, rtol=1e-13, atol=1e-14)


def mp_hessian_block(kfun, x, y):
    """d x d block  d/dx d/dy^T k(x, y) by high-precision central differences (the 'generic AD fallback' comparator,
    reference src/gradient.jl:27-42, test/gradient.jl:37-45)"""
    import mpmath as mp

    mp.mp.dps = 60
    d = len(x)
    h = mp.mpf(10) ** -15
    xs, ys = [mp.mpf(float(v)) for v in x], [mp.mpf(float(v)) for v in y]
    B = np.zeros((d, d))
    for c in range(d):
        for e in range(d):
            def f(sx, sy):
                xx, yy = list(xs), list(ys)
                xx[c] += sx
                yy[e] += sy
                return kfun(sum((a - b) ** 2 for a, b in zip(xx, yy)))
            B[c, e] = float((f(h, h) - f(h, -h) - f(-h, h) + f(-h, -h)) / (4 * h * h))
    return B


def test_gradient_gramian_structure(cf, O):
    # test/gradient.jl:29-52: size (d n)^2, symmetric to 1e4 eps, PSD, specialised == generic, 5-arg mul! vs dense
    import mpmath as mp

    rng = np.random.default_rng(7)
    n, d = 2, 5
    X = rng.standard_normal((n, d)) / math.sqrt(d)
    kernels = {
        "EQ": (cf.EQ(), lambda r2: mp.exp(-r2 / 2)),
        "MaternP(3)": (cf.MaternP(3), lambda r2: (lambda s: (1 + s + 2 * s**2 / 5 + s**3 / 15) * mp.exp(-s))(mp.sqrt(7 * r2))),
        "RQ(2)": (cf.RQ(2), lambda r2: (1 + r2 / 4) ** -2),
    }
    for name, (k, kmp) in kernels.items():
        M = O.gradient_matrix(k.program(), X)
        assert M.shape == (d * n, d * n)
        assert np.abs(M - M.T).max() < 1e4 * np.finfo(float).eps
        assert np.linalg.eigvalsh((M + M.T) / 2).min() >= -1e-12
        # block (0, 1) occupies rows 0:d, columns d:2d (BlockFactorization layout, test/gradient.jl:45)
        B = mp_hessian_block(kmp, X[0], X[1])
        assert np.allclose(M[0:d, d:2 * d], B, rtol=1e-9, atol=1e-11), name
        a, b = rng.standard_normal(d * n), rng.standard_normal(d * n)
        alpha, beta = rng.standard_normal(2)
        got = O.gradient_mul(k.program(), X, a, alpha=alpha, beta=beta, y0=b)
        assert np.allclose(got, alpha * (M @ a) + beta * b, rtol=1e-12, atol=1e-13)


def test_eq_gradient_element_woodbury(cf, O):
    # test/gradient.jl:66-70 with src/gradient.jl:95-105: W = (-2 k1) I + r (-4 k2) r'
    rng = np.random.default_rng(8)
    d = 5
    x, y, a = rng.standard_normal(d) / 2, rng.standard_normal(d) / 2, rng.standard_normal(d)
    r = x - y
    _, k1, k2 = O.value_derivative_laplacian(cf.EQ().program(), float(r @ r))
    W = -2 * k1 * np.eye(d) + np.outer(r, r) * (-4 * k2)
    Ga = O.gradient_mul(cf.EQ().program(), x[None, :], a, Y=y[None, :])
    assert np.allclose(W @ a, Ga, rtol=1e-13)
    # closed form: block = e (I - r r')
    e = math.exp(-float(r @ r) / 2)
    assert np.allclose(Ga, e * (a - r * (r @ a)), rtol=1e-13)


def test_cg_through_blockmul(cf, O):
    # test/gradient.jl:56-63: K \ (K a) residual < 1e-6
    rng = np.random.default_rng(9)
    n, d = 6, 3
    X = rng.standard_normal((n, d)) / math.sqrt(d)
    k = cf.MaternP(3)
    a = rng.standard_normal(n * d)
    Ka = O.gradient_mul(k.program(), X, a)
    xs, it, res, hist = O.cg_solve(k.program(), X, Ka, 0.0, gradient=True)
    assert np.linalg.norm(O.gradient_mul(k.program(), X, xs) - Ka) / np.linalg.norm(Ka) < 1e-6
    # scalar Gramian + sigma2 I (config 5 operator)
    y = rng.standard_normal(n)
    xs, it, res, hist = O.cg_solve(cf.MaternP(2).program(), X, y, 1e-2)
    M = O.matrix(cf.MaternP(2).program(), X) + 1e-2 * np.eye(n)
    assert np.allclose(M @ xs, y, rtol=1e-6, atol=1e-8)
    assert it <= n + 1 and len(hist) == it


def test_fast_loops_match_interpreter_bitwise(cf, O):
    rng = np.random.default_rng(10)
    X, a = rng.standard_normal((300, 3)), rng.standard_normal(300)
    for k in (cf.EQ(), cf.Exp(), cf.RQ(2), cf.RQ(1), cf.RQ(2.5), cf.MaternP(0), cf.MaternP(2), cf.MaternP(5)):
        O.lib().orc_set_force_interpreter(0)
        fast = O.mul_vec(k.program(), X, a, alpha=1.3, beta=0.4, y0=np.ones(300))
        O.lib().orc_set_force_interpreter(1)
        slow = O.mul_vec(k.program(), X, a, alpha=1.3, beta=0.4, y0=np.ones(300))
        O.lib().orc_set_force_interpreter(0)
        assert np.array_equal(fast, slow), repr(k)


def test_oracle_close_to_extended_precision_truth(cf, O):
    rng = np.random.default_rng(11)
    X, a = rng.standard_normal((1000, 3)), rng.standard_normal(1000)
    for k in all_kernels(cf):
        assert relerr(O.mul_vec(k.program(), X, a), O.truth_mul_vec(k.program(), X, a)) < 1e-14, repr(k)


def test_beta_zero_overwrites_nan(cf, O):
    # src/gramian.jl:80,90
    rng = np.random.default_rng(12)
    X, a = rng.standard_normal((20, 2)), rng.standard_normal(20)
    y = O.mul_vec(cf.EQ().program(), X, a, beta=0.0, y0=np.full(20, np.nan))
    assert np.all(np.isfinite(y))


def test_float32_promotion(cf, O):
    # src/gramian.jl:30-33 + Julia promotion: EQ on Float32 data stays Float32; MaternP / Constant{Float64} compute in Float64
    rng = np.random.default_rng(13)
    X = rng.standard_normal((50, 3)).astype(np.float32)
    a = rng.standard_normal(50).astype(np.float32)
    for k in (cf.EQ(), cf.RQ(2), cf.MaternP(2), 0.5 * cf.EQ()):
        b32 = O.mul_vec(k.program(), X, a, dtype=np.float32)
        b64 = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64))
        assert b32.dtype == np.float32
        assert relerr(b32, b64) < 5e-6
