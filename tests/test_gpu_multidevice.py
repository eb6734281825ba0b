"""Single-process multi-GPU mode (cf_init): one host thread drives several devices, rows sharded inside the library --
the form the Julia ccall shim uses.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


@pytest.fixture()
def two_gpus(cf):
    if cf.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cf.init([0, 1])
    yield
    cf.init([0])


def test_rows_sharded_over_two_devices(cf, O, two_gpus):
    rng = np.random.default_rng(41)
    n, m, d = 3001, 2000, 3  # ragged split
    X, Y = rng.standard_normal((n, d)), rng.standard_normal((m, d))
    a = rng.standard_normal(m)
    for k in (cf.EQ(), cf.MaternP(2), 0.5 * cf.RQ(2) + cf.Dot() ** 2):
        G = cf.gramian(k, X.T, Y.T)
        assert relerr(G @ a, O.mul_vec(k.program(), X, a, Y=Y)) < 1e-12
        y0 = rng.standard_normal(n)
        y = y0.copy()
        cf.mul_(y, G, a, 0.5, -2.0)
        assert relerr(y, O.mul_vec(k.program(), X, a, Y=Y, alpha=0.5, beta=-2.0, y0=y0)) < 1e-12
    A = rng.standard_normal((m, 5))
    G = cf.gramian(cf.EQ(), X.T, Y.T)
    assert relerr(G @ A, O.mul_mat(cf.EQ().program(), X, A, Y=Y)) < 1e-12
    Xg = rng.standard_normal((301, 4)) / 2
    ag = rng.standard_normal(301 * 4)
    Gg = cf.gramian(cf.GradientKernel(cf.MaternP(2)), Xg.T)
    assert relerr(Gg @ ag, O.gradient_mul(cf.MaternP(2).program(), Xg, ag)) < 1e-12
    M = cf.gramian(cf.EQ(), X[:100].T, Y[:70].T).Matrix()
    assert relerr(M, O.matrix(cf.EQ().program(), X[:100], Y[:70])) < 1e-13


def test_cg_over_two_devices(cf, O, two_gpus):
    rng = np.random.default_rng(42)
    n, d = 1201, 8
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    y = rng.standard_normal(n)
    k = cf.MaternP(2)
    A = 1e-2 * cf.I(n) + cf.gramian(k, X.T)
    x, iters, res = A.solve(y)
    xo, ito, reso, _ = O.cg_solve(k.program(), X, y, 1e-2)
    assert abs(iters - ito) <= max(3, 0.05 * ito)
    assert relerr(x, xo) < 1e-6


def test_cg_spmd_matches_single_device(cf, O):
    # the multi-device solve (fused peer-store all-gather, csrc/capi.cu cg_solve_spmd) against the single-device solve
    if cf.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rng = np.random.default_rng(43)
    n, d = 5003, 3
    X = rng.standard_normal((n, d))
    y = rng.standard_normal(n)
    k = cf.EQ()
    cf.init([0])
    x1, it1, res1 = (0.1 * cf.I(n) + cf.gramian(k, X.T)).solve(y)
    cf.init([0, 1])
    try:
        x2, it2, res2 = (0.1 * cf.I(n) + cf.gramian(k, X.T)).solve(y)
        Gg = cf.gramian(cf.GradientKernel(cf.MaternP(2)), X[:400].T)
        a = rng.standard_normal(400 * d)
        Ka = Gg @ a
        xs, its, ress = (1e-8 * cf.I(400 * d) + Gg).solve(Ka, reltol=1e-10, maxiter=3000)
    finally:
        cf.init([0])
    assert abs(it1 - it2) <= 2
    assert relerr(x2, x1) < 1e-8
    M = O.matrix(k.program(), X) + 0.1 * np.eye(n)
    assert np.linalg.norm(M @ x2 - y) / np.linalg.norm(y) < 1e-6
    assert np.linalg.norm((Gg @ xs) - Ka) / np.linalg.norm(Ka) < 1e-6


def test_runtime_specialised_and_tensor_core_kernels_on_two_devices(cf, O, two_gpus):
    """the run-time specialised kernels are loaded once (cuLibraryLoadData) and launched on both devices' streams; the
    tensor-core kernels keep one padded point copy per device"""
    import os
    rng = np.random.default_rng(77)
    n, d, p = 2500, 16, 5
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    a = rng.standard_normal(n)
    A = rng.standard_normal((n, p))
    k = 0.5 * cf.EQ() + cf.MaternP(2) * cf.RQ(2)
    before = cf.jit_stats()
    os.environ["COVFN_JIT"] = "1"
    try:
        G = cf.gramian(k, X.T.copy())
        b = G @ a
        B = G @ A
        X3 = rng.standard_normal((n, 3))
        G3 = cf.gramian(k, X3.T.copy())
        b3 = G3 @ a
    finally:
        del os.environ["COVFN_JIT"]
    after = cf.jit_stats()
    assert after["failures"] == before["failures"]
    assert relerr(b, O.mul_vec(k.program(), X, a)) < 1e-12
    assert relerr(B, O.mul_mat(k.program(), X, A)) < 1e-12
    assert relerr(b3, O.mul_vec(k.program(), X3, a)) < 1e-12


def test_float32_value_kernels_on_two_devices(cf, O, two_gpus):
    """the Float32 value kernels keep one transposed / canonical point copy per device: packed FP32 (d = 3), tcgen05 (d = 16), and the
    Float64 shadow of a Float32 handle (gradient operator, CG)"""
    rng = np.random.default_rng(78)
    n, m = 2901, 1777
    for d in (3, 16):
        X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
        Y = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
        a = rng.standard_normal(m).astype(np.float32)
        for k in (cf.EQ(), cf.MaternP(2)):
            G = cf.gramian(k, X.T.copy(), Y.T.copy())
            truth = O.mul_vec(k.program(), X.astype(np.float64), a.astype(np.float64), Y=Y.astype(np.float64))
            b = G @ a
            assert relerr(b.astype(np.float64), truth) < 1e-5, (d, k)
            assert np.array_equal(b, G @ a)
    Xg = (rng.standard_normal((400, 4)) / 2).astype(np.float32)
    ag = rng.standard_normal(400 * 4).astype(np.float32)
    Gg = cf.gramian(cf.GradientKernel(cf.EQ()), Xg.T.copy())
    assert relerr(Gg @ ag, O.gradient_mul(cf.EQ().program(), Xg.astype(np.float64), ag.astype(np.float64))) < 1e-5
    Xc = (rng.standard_normal((900, 3)) * 3).astype(np.float32)
    y = rng.standard_normal(900).astype(np.float32)
    x, iters, res = (0.5 * cf.I(900) + cf.gramian(cf.MaternP(2), Xc.T.copy())).solve(y)
    Kd = O.matrix(cf.MaternP(2).program(), Xc.astype(np.float64)) + 0.5 * np.eye(900)
    assert np.linalg.norm(Kd @ x - y) < 1e-3 * np.linalg.norm(y)


def test_symmetric_variant_over_two_devices(cf, O, two_gpus):
    """y === x on several devices of one process: every device evaluates the unordered pairs of its row tiles, the partial vectors
    are summed with peer loads in device order (csrc/capi.cu mul_host_sym_spmd); bit-reproducible."""
    rng = np.random.default_rng(44)
    n, d = 70001, 3
    X = rng.standard_normal((n, d))
    a = rng.standard_normal(n)
    for k in (cf.EQ(), cf.MaternP(2)):
        G = cf.gramian(k, X.T.copy())
        b = G @ a
        assert np.array_equal(b, G @ a)
        for rows in ((0, 50), (n // 2 - 25, n // 2 + 25), (n - 50, n)):
            assert relerr(b[rows[0]:rows[1]], O.mul_vec(k.program(), X, a, rows=rows)) < 1e-12
        y0 = rng.standard_normal(n)
        y = y0.copy()
        cf.mul_(y, G, a, -0.5, 2.0)
        assert relerr(y, -0.5 * b + 2.0 * y0) < 1e-13
        G.set_symmetric(False)
        assert relerr(G @ a, b) < 1e-13


def test_cg_spmd_symmetric_matches_plain(cf, two_gpus):
    rng = np.random.default_rng(45)
    n, d = 40000, 8
    X = rng.standard_normal((n, d)) / np.sqrt(d)
    y = rng.standard_normal(n)
    sols = {}
    for symm in (True, False):
        G = cf.gramian(cf.MaternP(2), X.T.copy())
        G.set_symmetric(symm)
        x, it, res = (1e-2 * cf.I(n) + G).solve(y, reltol=1e-300, maxiter=5)
        true_res = float(np.linalg.norm(y - (G @ x) - 1e-2 * x))
        assert it == 5 and abs(true_res - res) < 1e-8 * res
        sols[symm] = x
    assert relerr(sols[True], sols[False]) < 1e-8


def test_multi_process_comm_mode(cf):
    """one process per GPU under torchrun, the library's own NCCL communicator: bench_aux/comm_check.py"""
    import os
    import subprocess
    import sys

    if cf.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    run = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(root, "bench_aux", "comm_check.py")], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and "COMM_CHECK_OK" in run.stdout, run.stdout[-2000:] + run.stderr[-4000:]
