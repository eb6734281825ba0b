"""Seeded random sweep over shapes, dimensions, precisions, kernel programs and operators, through the C ABI against the oracle.
Small sizes, many combinations: the point is dispatch coverage (scalar / tensor-core / d > 32 paths, ragged tiles, rectangular
Gramians, alpha / beta, row ranges), in the spirit of the reference's randomized tests (test/gramian.jl, test/gradient.jl)."""
import os

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _random_kernel(cf, rng, isotropic_only):
    iso = [lambda: cf.EQ(), lambda: cf.MaternP(int(rng.integers(2, 5))), lambda: cf.RQ(int(rng.integers(1, 4))),
           lambda: cf.RQ(float(rng.uniform(0.6, 2.5))), lambda: cf.Lengthscale(cf.EQ(), float(rng.uniform(0.6, 1.6))),
           lambda: cf.Lengthscale(cf.MaternP(2), float(rng.uniform(0.6, 1.6)))]
    extra = [lambda: cf.Exp(), lambda: cf.MaternP(1), lambda: cf.Dot() ** int(rng.integers(1, 4)),
             lambda: (cf.Dot() + float(rng.uniform(0.1, 1.0))) ** 2]
    pool = iso if isotropic_only else iso + extra
    k = pool[int(rng.integers(len(pool)))]()
    r = rng.uniform()
    if r < 0.3:
        k = float(rng.uniform(0.2, 2.0)) * k + pool[int(rng.integers(len(pool)))]()
    elif r < 0.5:
        k = k * iso[int(rng.integers(len(iso)))]()
    elif r < 0.6:
        k = k + float(rng.uniform(0.1, 1.0))
    return k


@pytest.mark.parametrize("seed", range(int(os.environ.get("COVFN_FUZZ_SEEDS", "6"))))
def test_random_sweep(cf, O, seed):
    rng = np.random.default_rng(9000 + seed)
    os.environ["COVFN_GRAD_DMMA"] = "1"  # let the small gradient cases reach the tensor-core kernels too
    try:
        for case in range(14):
            d = int(rng.choice([1, 2, 3, 5, 8, 11, 16, 24, 32, 40]))
            n, m = int(rng.integers(1, 420)), int(rng.integers(1, 520))
            f32 = d <= 32 and rng.uniform() < 0.3
            op = rng.choice(["vec", "mat", "grad", "vgrad"]) if not f32 and d <= 32 else rng.choice(["vec", "mat"])
            if d > 32:
                op = rng.choice(["vec", "grad"])
            scale = 1.0 / np.sqrt(d) if rng.uniform() < 0.8 else 3.0  # mostly well-scaled, sometimes not (direct differences)
            X = rng.standard_normal((n, d)) * scale
            sym = rng.uniform() < 0.3
            Y = X if sym else rng.standard_normal((m, d)) * scale
            mm = Y.shape[0]
            k = _random_kernel(cf, rng, isotropic_only=op in ("grad", "vgrad"))
            alpha, beta = (1.0, 0.0) if rng.uniform() < 0.5 else (float(rng.uniform(-2, 2)), float(rng.uniform(-2, 2)))
            tag = (seed, case, d, n, mm, op, "f32" if f32 else "f64", repr(k))
            prog = k.program()
            if f32:
                Xd, Yd = X.astype(np.float32), Y.astype(np.float32)
                G = cf.gramian(k, Xd.T.copy()) if sym else cf.gramian(k, Xd.T.copy(), Yd.T.copy())
                X64, Y64 = Xd.astype(np.float64), Yd.astype(np.float64)
                if op == "vec":
                    a = rng.standard_normal(mm).astype(np.float32)
                    b0 = rng.standard_normal(n).astype(np.float32)
                    b = b0.copy()
                    cf.mul_(b, G, a, alpha, beta)
                    ref = O.mul_vec(prog, X64, a.astype(np.float64), Y=None if sym else Y64, alpha=alpha, beta=beta, y0=b0.astype(np.float64))
                else:
                    p = int(rng.integers(2, 6))
                    a = rng.standard_normal((mm, p)).astype(np.float32)
                    b0 = rng.standard_normal((n, p)).astype(np.float32)
                    b = np.asfortranarray(b0.copy())
                    cf.mul_(b, G, a, alpha, beta)
                    ref = O.mul_mat(prog, X64, a.astype(np.float64), Y=None if sym else Y64, alpha=alpha, beta=beta, B0=b0.astype(np.float64))
                assert relerr(b.astype(np.float64), ref) < 2e-5, tag
                continue
            Yo = None if sym else Y
            if op == "vec":
                G = cf.gramian(k, X.T.copy()) if sym else cf.gramian(k, X.T.copy(), Y.T.copy())
                a, b0 = rng.standard_normal(mm), rng.standard_normal(n)
                b = b0.copy()
                cf.mul_(b, G, a, alpha, beta)
                ref = O.mul_vec(prog, X, a, Y=Yo, alpha=alpha, beta=beta, y0=b0)
            elif op == "mat":
                G = cf.gramian(k, X.T.copy()) if sym else cf.gramian(k, X.T.copy(), Y.T.copy())
                p = int(rng.integers(2, 6))
                a, b0 = rng.standard_normal((mm, p)), rng.standard_normal((n, p))
                b = np.asfortranarray(b0.copy())
                cf.mul_(b, G, a, alpha, beta)
                ref = O.mul_mat(prog, X, a, Y=Yo, alpha=alpha, beta=beta, B0=b0)
            else:
                vg = op == "vgrad"
                K = cf.ValueGradientKernel(k) if vg else cf.GradientKernel(k)
                G = cf.gramian(K, X.T.copy()) if sym else cf.gramian(K, X.T.copy(), Y.T.copy())
                bs = d + (1 if vg else 0)
                a, b0 = rng.standard_normal(mm * bs), rng.standard_normal(n * bs)
                b = b0.copy()
                cf.mul_(b, G, a, alpha, beta)
                ref = O.derivative_mul(prog, X, a, Y=Yo, trait="isotropic", value_gradient=vg, alpha=alpha, beta=beta, y0=b0)
            assert relerr(b, ref) < 1e-12, tag
    finally:
        del os.environ["COVFN_GRAD_DMMA"]
