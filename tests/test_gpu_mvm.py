"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.

Tolerance from BASELINE.json north_star: relative 2-norm <= 1e-12 in Float64, <= 1e-5 in Float32.
Mirrors the reference's own tests: lazy product vs dense product on rectangular Gramians (test/gramian.jl:56-72),
entries vs k(x[i], y[j]) (test/gramian.jl:75-80), kernel algebra (test/algebra.jl:28-51), Dot/Poly (test/mercer.jl:12-19).
"""
import zlib

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu

TOL64 = 1e-12
TOL32 = 1e-5


def kernels(cf):
    return {
        "EQ": cf.EQ(),
        "Exp": cf.Exp(),
        "RQ(2)": cf.RQ(2),
        "RQ(1)": cf.RQ(1),
        "RQ(2.5)": cf.RQ(2.5),
        "MaternP(0)": cf.MaternP(0),
        "MaternP(1)": cf.MaternP(1),
        "MaternP(2)": cf.MaternP(2),
        "MaternP(3)": cf.MaternP(3),
        "MaternP(5)": cf.MaternP(5),
        "Dot": cf.Dot(),
        "Dot^3": cf.Dot() ** 3,
        "Poly(3,1)": cf.Poly(3, 1.0),
        "Line(0.5)": cf.Line(0.5),
        "0.5*EQ": 0.5 * cf.EQ(),
        "EQ+Exp": cf.EQ() + cf.Exp(),
        "EQ*RQ(2)": cf.EQ() * cf.RQ(2),
        "C3": 0.5 * cf.RQ(2) + cf.Dot() ** 2,
        "Lengthscale(EQ,0.7)": cf.Lengthscale(cf.EQ(), 0.7),
        "Lengthscale(MaternP(2),1.3)": cf.Lengthscale(cf.MaternP(2), 1.3),
        "(EQ+MaternP(1))^2": (cf.EQ() + cf.MaternP(1)) ** 2,
        "2*EQ+1": 2 * cf.EQ() + 1,
    }


@pytest.mark.parametrize("name", list(kernels(__import__("covfn_b200")).keys()))
@pytest.mark.parametrize("d", [1, 3, 5])
def test_mvm_vs_oracle_f64(cf, O, name, d):
    k = kernels(cf)[name]
    rng = np.random.default_rng(zlib.crc32(f"{name}-{d}".encode()))
    n, m = 257, 515  # rectangular, ragged against every tile size (test/gramian.jl:56-63)
    X = rng.standard_normal((n, d))
    Y = rng.standard_normal((m, d))
    a = rng.standard_normal(m)
    G = cf.gramian(k, X.T.copy(), Y.T.copy())
    assert G.shape == (n, m)
    b = G @ a
    ref = O.mul_vec(k.program(), X, a, Y=Y)
    assert relerr(b, ref) < TOL64
    # 5-argument form with NaN-filled y and beta = 0 (src/gramian.jl:80), then alpha/beta
    y = np.full(n, np.nan)
    cf.mul_(y, G, a, 1.0, 0.0)
    assert relerr(y, ref) < TOL64
    y0 = rng.standard_normal(n)
    y = y0.copy()
    cf.mul_(y, G, a, -0.7, 1.9)
    ref2 = O.mul_vec(k.program(), X, a, Y=Y, alpha=-0.7, beta=1.9, y0=y0)
    assert relerr(y, ref2) < TOL64


@pytest.mark.parametrize("d", [1, 2, 3, 4, 6, 7, 8, 11, 16, 20, 32])
def test_mvm_all_dims(cf, O, d):
    rng = np.random.default_rng(d)
    n = 300
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    a = rng.standard_normal(n)
    for k in (cf.EQ(), cf.MaternP(2), cf.RQ(2), 0.5 * cf.RQ(2) + cf.Dot() ** 2):
        G = cf.gramian(k, X.T.copy())
        assert relerr(G @ a, O.mul_vec(k.program(), X, a)) < TOL64


def test_symmetric_large_two_level(cf, O):
    # several column chunks + TMA ring wrap-around (n large enough for > 3 tiles per chunk)
    rng = np.random.default_rng(7)
    n, d = 16384, 3
    X = rng.standard_normal((n, d))
    a = rng.standard_normal(n)
    k = cf.MaternP(2)  # BASELINE config 1
    G = cf.gramian(k, X.T.copy())
    b = np.zeros(n)
    cf.mul_(b, G, a)
    rows = (0, 2048)
    ref = O.mul_vec(k.program(), X, a, rows=rows)
    assert relerr(b[rows[0]:rows[1]], ref) < TOL64
    truth = O.truth_mul_vec(k.program(), X, a, rows=rows)
    # both are rounding-close to the exact product; ours must not be worse than 4x the oracle's own error
    assert relerr(b[rows[0]:rows[1]], truth) < max(4 * relerr(ref, truth), 1e-14)


def test_entries_and_dense(cf, O):
    rng = np.random.default_rng(3)
    X = rng.standard_normal((37, 3))
    Y = rng.standard_normal((53, 3))
    for name, k in kernels(cf).items():
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        M = G.Matrix()
        Mo = O.matrix(k.program(), X, Y)
        assert M.shape == (37, 53)
        assert relerr(M, Mo) < 1e-13, name
        for (i, j) in [(0, 0), (5, 7), (36, 52)]:
            assert abs(G[i, j] - k(X[i], Y[j])) <= 1e-13 * max(1.0, abs(k(X[i], Y[j]))), name  # test/gramian.jl:75-80


def test_multi_rhs(cf, O):
    rng = np.random.default_rng(11)
    n, m, d, p = 130, 260, 3, 3  # test/gramian.jl:65-72
    X = rng.standard_normal((n, d))
    Y = rng.standard_normal((m, d))
    A = rng.standard_normal((m, p))
    for k in (cf.EQ(), 0.5 * cf.RQ(2) + cf.Dot() ** 2):
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        B = G @ A
        assert B.shape == (n, p)
        ref = O.mul_mat(k.program(), X, A, Y=Y)
        assert relerr(B, ref) < TOL64
        B0 = rng.standard_normal((n, p))
        B1 = np.asfortranarray(B0.copy())
        cf.mul_(B1, G, A, 0.3, -1.1)
        ref = O.mul_mat(k.program(), X, A, Y=Y, alpha=0.3, beta=-1.1, B0=B0)
        assert relerr(B1, ref) < TOL64
    # wide block: more than one 64-column pass, d = 32 (BASELINE config 3 shape, small n)
    n, d, p = 200, 32, 70
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    A = rng.standard_normal((n, p))
    k = 0.5 * cf.RQ(2) + cf.Dot() ** 2
    G = cf.gramian(k, X.T.copy())
    assert relerr(G @ A, O.mul_mat(k.program(), X, A)) < TOL64


def test_float32(cf, O):
    rng = np.random.default_rng(5)
    n, m, d = 400, 300, 3
    X = rng.standard_normal((n, d)).astype(np.float32)
    Y = rng.standard_normal((m, d)).astype(np.float32)
    a = rng.standard_normal(m).astype(np.float32)
    for k in (cf.EQ(), cf.Exp(), cf.RQ(2), cf.MaternP(2), cf.Dot() ** 2, 0.5 * cf.RQ(2) + cf.Dot() ** 2):
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        assert G.eltype == np.float32
        b = G @ a
        assert b.dtype == np.float32
        ref = O.mul_vec(k.program(), X, a, Y=Y, dtype=np.float32)
        assert relerr(b, ref) < TOL32, repr(k)


def test_edge_cases(cf, O):
    rng = np.random.default_rng(9)
    k = cf.EQ()
    # single point, single column
    X = rng.standard_normal((1, 3))
    G = cf.gramian(k, X.T.copy())
    assert relerr(G @ np.array([2.0]), np.array([2.0])) < 1e-15
    # duplicated points (r2 == 0 off the diagonal) for the sqrt-based kernels
    X = np.repeat(rng.standard_normal((5, 2)), 3, axis=0)
    a = rng.standard_normal(15)
    for kk in (cf.Exp(), cf.MaternP(2), cf.MaternP(1)):
        Gd = cf.gramian(kk, X.T.copy())
        assert relerr(Gd @ a, O.mul_vec(kk.program(), X, a)) < TOL64
    # far-apart points: exp underflow must give 0, not garbage
    X = np.array([[0.0, 0.0], [1e3, 0.0], [1e7, 1e7], [1e150, 0.0]])
    a = np.ones(4)
    for kk in (cf.EQ(), cf.Exp(), cf.MaternP(2)):
        Gd = cf.gramian(kk, X.T.copy())
        b = Gd @ a
        assert np.all(np.isfinite(b))
        assert relerr(b, O.mul_vec(kk.program(), X, a)) < TOL64
    # empty column set: y = beta*y
    G0 = cf.gramian(k, rng.standard_normal((3, 4)), rng.standard_normal((3, 0)))
    y = np.ones(4)
    cf.mul_(y, G0, np.zeros(0), 1.0, 2.0)
    assert np.allclose(y, 2.0)
    # NaN coordinate is rejected at create time
    Xn = rng.standard_normal((4, 3))
    Xn[2, 1] = np.nan
    with pytest.raises(cf.DomainError):
        cf.gramian(k, Xn.T.copy()) @ np.ones(4)
    # NaN weights propagate like the reference
    X = rng.standard_normal((8, 3))
    a = rng.standard_normal(8)
    a[3] = np.nan
    assert np.all(np.isnan(cf.gramian(k, X.T.copy()) @ a))
    # dimension mismatch
    with pytest.raises(cf.DimensionMismatch):
        cf.gramian(k, X.T.copy()) @ np.ones(9)


def test_row_range(cf, O):
    rng = np.random.default_rng(13)
    n, d = 1000, 3
    X = rng.standard_normal((n, d))
    a = rng.standard_normal(n)
    k = cf.MaternP(2)
    full = cf.gramian(k, X.T.copy()) @ a
    parts = []
    for r in range(3):
        G = cf.gramian(k, X.T.copy()).set_row_range(n * r // 3, n * (r + 1) // 3)
        parts.append(G @ a)
    assert np.array_equal(np.concatenate(parts), full)  # sharding does not change a single bit


def test_ard_lengthscales(cf, O):
    """ARD(k, l) = Normed(k, tau -> sum(tau^2 / l)) (src/transformation.jl:42-45, test/stationary.jl:132-154): the ARD node of the
    C ABI (metric applied to the points on the device) against the oracle's restatement of enorm2(Diagonal(inv.(l)), x - y)."""
    rng = np.random.default_rng(17)
    n, m = 260, 145
    for d in (2, 3, 5, 11):
        X, Y = rng.standard_normal((n, d)), rng.standard_normal((m, d))
        a = rng.standard_normal(m)
        A = rng.standard_normal((m, 3))
        l = np.exp(rng.standard_normal(d))
        for k in (cf.EQ(), cf.MaternP(2), cf.RQ(2), cf.Exp(), 0.5 * cf.EQ() + cf.MaternP(1) * cf.RQ(1.5)):
            kard = cf.ARD(k, l)
            prog = kard.program()
            G = cf.gramian(kard, X.T.copy(), Y.T.copy())
            assert relerr(G @ a, O.mul_vec(prog, X, a, Y=Y)) < TOL64
            assert relerr(G @ A, O.mul_mat(prog, X, A, Y=Y)) < TOL64
            assert relerr(G.Matrix(), O.matrix(prog, X, Y)) < 1e-13
            assert abs(G[3, 7] - kard(X[3], Y[7])) < 1e-13
            # symmetric case and a leading constant outside the ARD node
            Gs = cf.gramian(2.5 * kard, X.T.copy())
            assert relerr(Gs @ a[:1].repeat(n), O.mul_vec((2.5 * kard).program(), X, a[:1].repeat(n))) < TOL64
    # Float32 data
    X32 = rng.standard_normal((200, 3)).astype(np.float32)
    a32 = rng.standard_normal(200).astype(np.float32)
    k32 = cf.ARD(cf.EQ(), [0.5, 2.0, 1.3])
    assert relerr(cf.gramian(k32, X32.T.copy()) @ a32, O.mul_vec(k32.program(), X32, a32, dtype=np.float32)) < 1e-5
    assert isinstance(cf.ARD(cf.EQ(), 2.0), cf.Lengthscale)  # ARD(k, l::Real) = Lengthscale(k, l)
    # one metric per program; no derivative operators (Normed is a StationaryKernel)
    with pytest.raises(cf.DimensionMismatch):
        cf.gramian(cf.ARD(cf.EQ(), [1.0, 2.0]), X.T.copy()).handle()
    with pytest.raises(cf.UnsupportedKernel):
        cf.gramian(cf.ARD(cf.EQ(), l) + cf.MaternP(2), X.T.copy()).handle()
    with pytest.raises(cf.UnsupportedKernel):
        cf.gramian(cf.ARD(cf.EQ(), l) + cf.Dot(), X.T.copy()).handle()
    with pytest.raises(cf.UnsupportedKernel):
        cf.gramian(cf.ARD(cf.EQ(), l) * cf.ARD(cf.RQ(2), 2 * l), X.T.copy()).handle()
    Gsame = cf.gramian(cf.ARD(cf.EQ(), l) * cf.ARD(cf.RQ(2), l), X.T.copy())  # the same metric twice is one pre-scaling
    assert relerr(Gsame @ a[:1].repeat(n), O.mul_vec((cf.ARD(cf.EQ(), l) * cf.ARD(cf.RQ(2), l)).program(), X, a[:1].repeat(n))) < TOL64


def test_runtime_specialised_mvm_matches_interpreter(cf, O):
    """composite-kernel MVM with the program structure compiled in at run time (COVFN_JIT=1, csrc/cf_jit.h) against the
    interpreter kernel and the oracle, for row tiles of every R (d = 3: R = 4, d = 8: R = 2, d = 20: R = 1)"""
    import os
    rng = np.random.default_rng(41)
    for d in (3, 8, 20):
        n, m = 700, 1100
        X = rng.standard_normal((n, d)) / np.sqrt(d)
        Y = rng.standard_normal((m, d)) / np.sqrt(d)
        a = rng.standard_normal(m)
        for k in (0.5 * cf.EQ() + cf.MaternP(2) * cf.RQ(2), cf.Exp() + 0.1 * (cf.Dot() + 1.0) ** 2, (cf.EQ() + cf.RQ(1.5)) ** 2):
            G = cf.gramian(k, X.T.copy(), Y.T.copy())
            before = cf.jit_stats()
            os.environ["COVFN_JIT"] = "1"
            try:
                bj = G @ a
            finally:
                os.environ["COVFN_JIT"] = "0"
            try:
                bi = G @ a
            finally:
                del os.environ["COVFN_JIT"]
            after = cf.jit_stats()
            assert after["failures"] == before["failures"]
            assert after["compiled"] + after["cache_hits"] > before["compiled"] + before["cache_hits"]
            ref = O.mul_vec(k.program(), X, a, Y=Y)
            assert relerr(bj, ref) < TOL64 and relerr(bi, ref) < TOL64
            assert relerr(bj, bi) < 1e-13


def test_runtime_specialisation_unavailable_falls_back_to_interpreter_kernel():
    """without libnvrtc the library must keep using the ahead-of-time (interpreter) CUDA kernel and count a failure --
    run in a subprocess because the library probes NVRTC once per process"""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, covfn_b200 as cf\n"
        "from oracle import oracle as O\n"
        "rng = np.random.default_rng(3)\n"
        "X = rng.standard_normal((500, 3)); a = rng.standard_normal(500)\n"
        "k = 0.5 * cf.EQ() + cf.MaternP(2) * cf.RQ(2)\n"
        "b = cf.gramian(k, X.T.copy()) @ a\n"
        "ref = O.mul_vec(k.program(), X, a)\n"
        "s = cf.jit_stats()\n"
        "assert np.linalg.norm(b - ref) / np.linalg.norm(ref) < 1e-12\n"
        "assert s['compiled'] == 0 and s['failures'] >= 1, s\n"
        "print('fallback ok')\n")
    env = dict(os.environ, COVFN_JIT="1", COVFN_JIT_TEST_NO_NVRTC="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "fallback ok" in out.stdout, out.stderr[-2000:]


def test_handles_return_all_device_memory(cf):
    """create / multiply / destroy in a loop: every per-handle buffer (padded point copies of the tensor-core kernels, padded
    weights, partial sums, transposed right-hand sides) must go back to the pool -- free device memory stays flat"""
    import torch
    rng = np.random.default_rng(2)
    n, d = 30000, 16
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    a = rng.standard_normal(n)
    A = rng.standard_normal((n, 3))
    g = rng.standard_normal(n * d)

    def cycle():
        G = cf.gramian(cf.MaternP(2), X.T.copy())
        G @ a
        G @ A
        G.close()
        H = cf.gramian(cf.GradientKernel(cf.EQ()), X[:3000].T.copy())
        H @ g[:3000 * d]
        H.close()

    for _ in range(3):
        cycle()
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    for _ in range(12):
        cycle()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 8 << 20, f"device memory shrank by {(free0 - free1) / 2**20:.1f} MiB over 12 handle lifetimes"


def test_runtime_specialisation_disk_cache(tmp_path):
    """a second process finds the compiled cubin in $COVFN_JIT_CACHE: no compilation, one cache hit, same result"""
    import os
    import subprocess
    import sys
    code = (
        "import numpy as np, covfn_b200 as cf\n"
        "rng = np.random.default_rng(3)\n"
        "X = rng.standard_normal((500, 3)); a = rng.standard_normal(500)\n"
        "k = 0.25 * cf.EQ() + cf.MaternP(3) * cf.RQ(2)\n"
        "b = cf.gramian(k, X.T.copy()) @ a\n"
        "s = cf.jit_stats()\n"
        "print('STATS', s['compiled'], s['cache_hits'], s['failures'], repr(float(np.linalg.norm(b))))\n")
    env = dict(os.environ, COVFN_JIT="1", COVFN_JIT_CACHE=str(tmp_path))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for _ in range(2):
        out = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr[-2000:]
        outs.append([l for l in out.stdout.splitlines() if l.startswith("STATS")][0].split())
    assert outs[0][1:4] == ["1", "0", "0"], outs[0]      # first process compiles
    assert outs[1][1:4] == ["0", "1", "0"], outs[1]      # second process loads the cubin from disk
    assert outs[0][4] == outs[1][4]                       # bit-identical result
    assert any(f.endswith(".cubin") for f in os.listdir(tmp_path))
