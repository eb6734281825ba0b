"""BASELINE.json configurations at their FULL sizes, checked through size-independent properties plus oracle parity on
row subsets (the oracle finishes a few thousand rows in seconds).  Tolerance: relative 2-norm 1e-12 (north_star)."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _philox(seed):
    return np.random.Generator(np.random.Philox(seed))


def test_config2_eq_n2pow20(cf, O):
    # EQ, d = 3, n = 2^20, Float64
    n, d = 1 << 20, 3
    rng = _philox(0xC0F00002)
    X = rng.standard_normal((n, d))
    a, c = rng.standard_normal(n), rng.standard_normal(n)
    k = cf.EQ()
    G = cf.gramian(k, X.T)
    Ka, Kc = G @ a, G @ c
    # oracle parity on two row blocks (first rows and a block in the middle)
    for rows in ((0, 1024), (n // 2 + 77, n // 2 + 77 + 512)):
        ref = O.mul_vec(k.program(), X, a, rows=rows)
        assert relerr(Ka[rows[0]:rows[1]], ref) < 1e-12
    # linearity
    alpha, beta = 0.37, -1.9
    assert relerr(G @ (alpha * a + beta * c), alpha * Ka + beta * Kc) < 1e-12
    # symmetry of K: c'(K a) == a'(K c)
    s1, s2 = float(c @ Ka), float(a @ Kc)
    assert abs(s1 - s2) <= 1e-11 * (np.linalg.norm(c) * np.linalg.norm(Ka))
    # a row block computed alone (the multi-GPU shard of rank 3 of 8) equals the same rows of the full product; the column
    # chunking (hence the summation order) is chosen per launch shape, so the match is to rounding, not bit-for-bit
    r0, r1 = 3 * n // 8, 4 * n // 8
    Gs = cf.gramian(k, X.T).set_row_range(r0, r1)
    bs = Gs @ a
    assert relerr(bs, Ka[r0:r1]) < 1e-14
    assert np.array_equal(bs, Gs @ a)  # and a given launch shape is deterministic run to run
    # positive weights: every entry of K is in (0, 1], so 0 < (K 1)_i <= n and >= 1 (diagonal)
    ones = G @ np.ones(n)
    assert ones.min() >= 1.0 and ones.max() <= n


def test_config1_maternp2_n16384(cf, O):
    n, d = 16384, 3
    rng = _philox(0xC0F00001)
    X = rng.standard_normal((n, d))
    a = rng.standard_normal(n)
    k = cf.MaternP(2)
    G = cf.gramian(k, X.T)
    b = np.zeros(n)
    cf.mul_(b, G, a)
    assert relerr(b, O.mul_vec(k.program(), X, a)) < 1e-12  # all rows: the README example (README.md:26-38)


def test_config3_multirhs_d32_n262144(cf, O):
    n, d, p = 262144, 32, 64
    rng = _philox(0xC0F00003)
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    A = np.asfortranarray(rng.standard_normal((n, p)))
    k = 0.5 * cf.RQ(2) + cf.Dot() ** 2
    rows = (1000, 1000 + 4096)
    G = cf.gramian(k, X.T).set_row_range(*rows)  # a row block of the full-size operator (all 262144 columns)
    B = G @ A
    assert B.shape == (4096, p)
    # multi-RHS kernel == vector kernel column by column (different code paths: K4 vs K3)
    for j in (0, 17, 63):
        assert relerr(B[:, j], G @ np.ascontiguousarray(A[:, j])) < 1e-12
    # oracle (reference loop order, entry re-evaluated per column) on 64 rows x 2 columns
    sub = (rows[0], rows[0] + 64)
    ref = O.mul_mat(k.program(), X, A[:, :2], rows=sub)
    assert relerr(B[:64, :2], ref) < 1e-12


def test_config4_gradient_d16_n65536(cf, O):
    n, d = 65536, 16
    rng = _philox(0xC0F00004)
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    a, c = rng.standard_normal(n * d), rng.standard_normal(n * d)
    k = cf.EQ()
    G = cf.gramian(cf.GradientKernel(k), X.T)
    Ga, Gc = G @ a, G @ c
    rows = (5000, 5064)
    ref = O.gradient_mul(k.program(), X, a, rows=rows)
    assert relerr(Ga[rows[0] * d:rows[1] * d], ref) < 1e-12
    assert abs(float(c @ Ga) - float(a @ Gc)) <= 1e-11 * np.linalg.norm(c) * np.linalg.norm(Ga)  # symmetric operator
    assert relerr(G @ (2 * a - c), 2 * Ga - Gc) < 1e-12


def test_config5_cg_d8_n2pow19_bounded_iterations(cf, O):
    # (K + sigma2 I) \ y with MaternP(2), d = 8, n = 2^19: a bounded number of iterations at full size
    n, d = 1 << 19, 8
    rng = _philox(0xC0F00005)
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    y = rng.standard_normal(n)
    k, sigma2 = cf.MaternP(2), 1e-2
    G = cf.gramian(k, X.T)
    A = sigma2 * cf.I(n) + G
    x, iters, res = A.solve(y, maxiter=6)
    assert iters == 6
    true_res = np.linalg.norm(y - (A @ x))
    assert abs(true_res - res) <= 1e-8 * np.linalg.norm(y)  # the recurrence residual is the true residual
    # operator parity on a row block
    v = rng.standard_normal(n)
    Av = A @ v
    rows = (12345, 12345 + 512)
    assert relerr(Av[rows[0]:rows[1]], sigma2 * v[rows[0]:rows[1]] + O.mul_vec(k.program(), X, v, rows=rows)) < 1e-12


def test_symmetric_variant_matches_default(cf, O):
    # CF_OPT_SYMMETRIC: every unordered pair once (gram_mvm_sym.cuh); equal to the default path to rounding
    n, d = 1 << 17, 3
    rng = _philox(0xC0F00012)
    X = rng.standard_normal((n, d))
    a = rng.standard_normal(n)
    for k in (cf.EQ(), cf.MaternP(2), 0.5 * cf.RQ(2) + cf.EQ()):
        G = cf.gramian(k, X.T)
        b0 = G @ a
        G.set_symmetric(True)
        b1 = G @ a
        assert relerr(b1, b0) < 1e-13, repr(k)
        rows = (n - 300, n)  # the last rows: their result is almost entirely column sums
        assert relerr(b1[rows[0]:rows[1]], O.mul_vec(k.program(), X, a, rows=rows)) < 1e-12
        y0 = rng.standard_normal(n)
        y = y0.copy()
        cf.mul_(y, G, a, -0.5, 2.0)
        assert relerr(y, -0.5 * b0 + 2.0 * y0) < 1e-13
    # ragged n (not a multiple of any tile size)
    n2 = 70001
    G = cf.gramian(cf.EQ(), X[:n2].T)
    b0 = G @ a[:n2]
    assert relerr(G.set_symmetric(True) @ a[:n2], b0) < 1e-13
