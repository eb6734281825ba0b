"""The Julia shim's ccall sequences, replayed by the C harness tests/abi/shim_sequence.c (gcc, dlopen of the product library)
and checked against the oracle: the executable stand-in for julia/CovarianceFunctionsB200.jl, which cannot run here (no Julia
in the image or on the GPU box).  Scenario tags S1..S10 are the ones the shim's comments carry."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, relerr

HARNESS_SRC = os.path.join(ROOT, "tests", "abi", "shim_sequence.c")
LIB = os.path.join(ROOT, "covariancefunctions.jl_b200", "lib", "libcovfn_b200.so")


def build_harness(tmp_path):
    exe = str(tmp_path / "shim_sequence")
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    subprocess.check_call([cc, "-O2", "-Wall", "-I" + os.path.join(ROOT, "include"), HARNESS_SRC, "-ldl", "-lm", "-o", exe])
    return exe


def test_harness_builds_and_binds_every_symbol(tmp_path):
    """CPU: the harness compiles against include/covfn_b200.h and resolves every symbol the shim ccalls (no compute)."""
    exe = build_harness(tmp_path)
    out = subprocess.run([exe, LIB, "--symbols"], capture_output=True, text=True)
    assert out.returncode == 0 and "symbols ok" in out.stdout, out.stderr


def read_records(path):
    recs = {}
    with open(path, "rb") as f:
        while True:
            nm = f.read(32)
            if len(nm) < 32:
                break
            (count,) = struct.unpack("<q", f.read(8))
            recs[nm.split(b"\0")[0].decode()] = np.frombuffer(f.read(8 * count), dtype=np.float64).copy()
    return recs


@pytest.mark.gpu
def test_shim_call_sequences_match_oracle(tmp_path, cf, O):
    exe = build_harness(tmp_path)
    rng = np.random.default_rng(2024)
    n, d, p = 384, 3, 5
    X = rng.standard_normal((n, d))
    Y2 = rng.standard_normal((n, d))
    a = rng.standard_normal(n)
    A = rng.standard_normal((n, p))
    ag = rng.standard_normal(n * d)
    avg = rng.standard_normal(n * (d + 1))
    rhs = rng.standard_normal(n)
    rhsg = rng.standard_normal(n * d)
    l = np.exp(rng.standard_normal(d))
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("<qqq", n, d, p))
        for arr in (X, Y2, a, np.asfortranarray(A).T, ag, avg, rhs, rhsg, l):
            f.write(np.ascontiguousarray(arr, dtype=np.float64).tobytes())
    run = subprocess.run([exe, LIB, fin, fout], capture_output=True, text=True)
    assert run.returncode == 0 and "shim sequences ok" in run.stdout, run.stderr + run.stdout
    R = read_records(fout)
    eq, m2 = cf.EQ().program(), cf.MaternP(2).program()
    # S1: two kernels on one x; same x, different y
    ref_eq, ref_m2 = O.mul_vec(eq, X, a), O.mul_vec(m2, X, a)
    assert relerr(R["S1_eq"], ref_eq) < 1e-12 and relerr(R["S1_matern2"], ref_m2) < 1e-12
    assert relerr(R["S1_eq_again"], ref_eq) < 1e-12
    assert relerr(ref_eq, ref_m2) > 1e-2  # the two products really differ
    assert relerr(R["S1_eq_xy"], O.mul_vec(eq, X, a[: n // 2], Y=Y2[: n // 2])) < 1e-12
    # S2: Float32 points under a Float64 Gramian, and a genuine Float32 Gramian
    X32 = X.astype(np.float32)
    half_rq = (0.5 * cf.RQ(2)).program()
    assert relerr(R["S2_f32pts_f64gram"], O.mul_vec(half_rq, X32.astype(np.float64), a)) < 1e-12
    assert relerr(R["S2_f32"], O.mul_vec(eq, X32, a.astype(np.float32), dtype=np.float32)) < 1e-5
    # S3: one result per length scale
    S3 = R["S3_lengthscales"].reshape(4, n)
    for it in range(4):
        assert relerr(S3[it], O.mul_vec(cf.Lengthscale(cf.EQ(), 0.5 + 0.25 * it).program(), X, a)) < 1e-12
    # S4: matrix mul! through a view (leading dimension n + 3), alpha = -0.5, beta = 2
    S4 = R["S4_matrix"].reshape(p, n + 3)
    assert np.all(S4[:, n:] == -77.0)  # rows beyond n untouched
    assert relerr(S4[:, :n].T, O.mul_mat(m2, X, A, alpha=-0.5, beta=2.0, B0=0.25 * A)) < 1e-12
    # S5: blockmul! for GradientKernel(EQ) (alpha 1.5, beta -1) and ValueGradientKernel(MaternP(2))
    assert relerr(R["S5_gradient"], O.derivative_mul(eq, X, ag, alpha=1.5, beta=-1.0, y0=0.5 * ag)) < 1e-12
    assert relerr(R["S5_value_gradient"], O.derivative_mul(m2, X, avg, value_gradient=True)) < 1e-12
    # S6 / S7: ldiv! (CG on the device) against the oracle's restatement of cg!
    xo, ito, reso, _ = O.cg_solve(m2, X, rhs, 1e-2)
    S6 = R["S6_ldiv_lazysum"]
    # (~270 iterations on an ill-conditioned system: the count moves by a few with the summation order)
    assert abs(int(S6[n]) - ito) <= 0.05 * ito + 2 and relerr(S6[:n], xo) < 1e-5
    K = O.matrix(m2, X) + 1e-2 * np.eye(n)
    assert np.linalg.norm(K @ S6[:n] - rhs) <= 2e-8 * np.linalg.norm(rhs) * np.linalg.cond(K) ** 0.5
    S7 = R["S7_ldiv_blockgramian"]
    Gd = O.gradient_matrix(eq, 4.0 * Y2)
    assert np.linalg.norm(Gd @ S7[: n * d] - rhsg) < 1e-6 * np.linalg.norm(rhsg)  # test/gradient.jl:56-63 residual check
    # S9: ARD node and the error codes the shim turns into exceptions
    assert relerr(R["S9_ard"], O.mul_vec(cf.ARD(cf.MaternP(2), l).program(), X, a)) < 1e-12
    assert list(R["S9_errors"]) == [-2.0, -2.0, 1.0]
    # S10: Float32 Gramians through blockmul! and ldiv! (Float64 arithmetic inside the library, Float32 vectors)
    X32s = (2.0 * Y2).astype(np.float32)
    ag32 = ag.astype(np.float32)
    assert relerr(R["S10_f32_gradient"], O.derivative_mul(eq, X32s.astype(np.float64), ag32.astype(np.float64))) < 1e-5
    S10 = R["S10_f32_ldiv"]
    K32 = O.matrix(m2, X32s.astype(np.float64)) + 0.5 * np.eye(n)
    rhs32 = rhs.astype(np.float32).astype(np.float64)
    assert int(S10[n]) > 0 and np.linalg.norm(K32 @ S10[:n] - rhs32) < 2e-3 * np.linalg.norm(rhs32)
    # S8: every visible device
    S8 = R["S8_init_all_devices"]
    assert int(S8[n]) == cf.device_count() or int(S8[n]) == 8
    assert relerr(S8[:n], ref_m2) < 1e-12
    S8l = R["S8_ldiv_all_devices"]
    assert abs(int(S8l[n]) - ito) <= 0.05 * ito + 2 and relerr(S8l[:n], xo) < 1e-5
