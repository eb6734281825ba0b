"""Float32 Gramians with the operators that exist in Float64 only -- derivative kernels, conjugate gradients, d > 32 -- run on a
Float64 shadow of the points (csrc/capi.cu ensure_shadow64): inputs and outputs are Float32, as in the reference, which is generic in
the element type (src/gradient.jl:86-92, src/gramian.jl:229-238).  Tolerance 1e-5 against the oracle on the same Float32 data."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def test_float32_gradient_and_value_gradient_mvm(cf, O):
    rng = np.random.default_rng(91)
    n, m, d = 300, 211, 5
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Y = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    for k in (cf.EQ(), cf.MaternP(2), cf.Dot() ** 3):
        trait = "dotproduct" if isinstance(k, cf.Power) else "isotropic"
        a = rng.standard_normal(m * d).astype(np.float32)
        G = cf.gramian(cf.GradientKernel(k), X.T.copy(), Y.T.copy())
        assert G.eltype == np.float32
        b = G @ a
        assert b.dtype == np.float32
        ref = O.derivative_mul(k.program(), X.astype(np.float64), a.astype(np.float64), Y=Y.astype(np.float64), trait=trait)
        assert relerr(b, ref) < 1e-5
        y0 = rng.standard_normal(n * d).astype(np.float32)
        y = y0.copy()
        cf.mul_(y, G, a, 1.5, -1.0)
        assert relerr(y, 1.5 * ref - y0) < 1e-5
        av = rng.standard_normal(m * (d + 1)).astype(np.float32)
        V = cf.gramian(cf.ValueGradientKernel(k), X.T.copy(), Y.T.copy())
        refv = O.derivative_mul(k.program(), X.astype(np.float64), av.astype(np.float64), Y=Y.astype(np.float64), trait=trait, value_gradient=True)
        assert relerr(V @ av, refv) < 1e-5


def test_float32_cg_solve(cf, O):
    rng = np.random.default_rng(92)
    n, d = 600, 3
    X = (rng.standard_normal((n, d)) * 3).astype(np.float32)  # spread out: a well-conditioned system
    y = rng.standard_normal(n).astype(np.float32)
    k = cf.MaternP(2)
    A = 0.5 * cf.I(n) + cf.gramian(k, X.T.copy())
    x, iters, res = A.solve(y)
    assert x.dtype == np.float32 and iters > 0
    Kd = O.matrix(k.program(), X.astype(np.float64)) + 0.5 * np.eye(n)
    xs = np.linalg.solve(Kd, y.astype(np.float64))
    assert relerr(x, xs) < 2e-3  # reltol = sqrt(eps(Float32)) = 3.5e-4, as cg! would use for Float32
    assert np.linalg.norm(Kd @ x - y) < 1e-3 * np.linalg.norm(y)
    # gradient system
    Xg = (rng.standard_normal((60, 3)) * 2).astype(np.float32)
    Gg = cf.gramian(cf.GradientKernel(cf.EQ()), Xg.T.copy())
    rhs = (Gg @ rng.standard_normal(180).astype(np.float32))
    xg, it, r = (0.1 * cf.I(180) + Gg).solve(rhs)
    Md = O.gradient_matrix(cf.EQ().program(), Xg.astype(np.float64)) + 0.1 * np.eye(180)
    assert np.linalg.norm(Md @ xg - rhs) < 2e-3 * np.linalg.norm(rhs)


def test_float32_points_beyond_32_dimensions(cf, O):
    rng = np.random.default_rng(93)
    n, m, d = 257, 190, 40
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Y = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    a = rng.standard_normal(m).astype(np.float32)
    for k in (cf.EQ(), 0.5 * cf.RQ(2) + cf.Dot() ** 2):
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        b = G @ a
        assert b.dtype == np.float32
        assert relerr(b, O.mul_vec(k.program(), X, a, Y=Y, dtype=np.float32)) < 1e-5
    ag = rng.standard_normal(m * d).astype(np.float32)
    Gg = cf.gramian(cf.GradientKernel(cf.EQ()), X.T.copy(), Y.T.copy())
    ref = O.derivative_mul(cf.EQ().program(), X.astype(np.float64), ag.astype(np.float64), Y=Y.astype(np.float64))
    assert relerr(Gg @ ag, ref) < 1e-5


def test_float32_device_pointers_with_padded_blocks(cf, O):
    """cf_gradient_mul_device / cf_value_gradient_mul_device on a Float32 handle: the device-side conversion buffers must not alias the
    scratch the Float64 operator pads its input into (d = 5 is padded to 6; the value-gradient form always repacks)"""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(94)
    n, d = 700, 5
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    dev = torch.device("cuda", 0)
    for vg in (False, True):
        bs = d + (1 if vg else 0)
        a = rng.standard_normal(n * bs).astype(np.float32)
        K = cf.ValueGradientKernel(cf.EQ()) if vg else cf.GradientKernel(cf.EQ())
        G = cf.gramian(K, X.T.copy())
        a_dev = torch.from_numpy(a).to(dev)
        y0 = rng.standard_normal(n * bs).astype(np.float32)
        b_dev = torch.from_numpy(y0).to(dev)
        torch.cuda.synchronize()
        G.mul_device(b_dev.data_ptr(), a_dev.data_ptr(), alpha=1.5, beta=-0.5)
        torch.cuda.synchronize()
        ref = O.derivative_mul(cf.EQ().program(), X.astype(np.float64), a.astype(np.float64), value_gradient=vg)
        assert relerr(b_dev.cpu().numpy(), 1.5 * ref - 0.5 * y0) < 1e-5, vg
        assert relerr(a_dev.cpu().numpy(), a) == 0.0  # the input vector is untouched
