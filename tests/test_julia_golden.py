"""Golden vectors from the UNMODIFIED reference (julia/make_golden.jl): when tests/golden/julia_v1/manifest.json is present, both the
CPU oracle and the GPU library are compared with the reference's own outputs on the reference's own inputs (1e-12 Float64 /
1e-5 Float32).  Julia exists neither in the build image nor on the GPU box (profiles/r2_julia_probe.txt), so today these tests skip
-- they are the route by which parity becomes pinned the moment someone runs the generator.  The reader itself is exercised on a
self-made fixture in the same format (test_reader_round_trip)."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, relerr

GOLD = os.path.join(ROOT, "tests", "golden", "julia_v1")
HAVE = os.path.exists(os.path.join(GOLD, "manifest.json"))
need_golden = pytest.mark.skipif(not HAVE, reason="no Julia golden vectors (run julia/make_golden.jl with the reference installed)")


def load_case(case, base=GOLD):
    """arrays of one manifest entry; Julia arrays are column-major: shape [d, n] comes back as an (n, d) C-ordered numpy array"""
    out = {}
    for key, meta in case["arrays"].items():
        dt = np.float32 if meta["dtype"] == "f32" else np.float64
        raw = np.fromfile(os.path.join(base, meta["file"]), dtype=np.dtype(dt).newbyteorder("<"))
        out[key] = raw.astype(dt).reshape(tuple(reversed(meta["shape"])))
    return out


def kernel_of(cf, name, arrays):
    K = {"EQ": cf.EQ(), "Exp": cf.Exp(), "RQ(2)": cf.RQ(2), "RQ(1.5)": cf.RQ(1.5), "MaternP(0)": cf.MaternP(0), "MaternP(1)": cf.MaternP(1),
         "MaternP(2)": cf.MaternP(2), "MaternP(3)": cf.MaternP(3), "Dot^3": cf.Dot() ** 3, "1/2*RQ(2)+Dot()^2": 0.5 * cf.RQ(2) + cf.Dot() ** 2,
         "0.5*RQ(2)": 0.5 * cf.RQ(2), "Lengthscale(EQ,0.7)": cf.Lengthscale(cf.EQ(), 0.7),
         "2.5*Lengthscale(MaternP(2),1.3)": 2.5 * cf.Lengthscale(cf.MaternP(2), 1.3), "EQ+1/2*RQ(2)": cf.EQ() + 0.5 * cf.RQ(2),
         "1/2*EQ+MaternP(2)*RQ(2)": 0.5 * cf.EQ() + cf.MaternP(2) * cf.RQ(2)}
    if name.startswith("ARD("):
        return cf.ARD(K[name[4:-1]], arrays["l"].ravel())
    return K[name]


def cases():
    if not HAVE:
        return []
    return json.load(open(os.path.join(GOLD, "manifest.json")))["cases"]


def tol_of(arr):
    return 1e-5 if arr.dtype == np.float32 else 1e-12


def nan_aware_close(got, ref, tol):
    """the reference's NaN / Inf pattern must be reproduced; finite entries to tol"""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    fin = np.isfinite(ref)
    if not np.array_equal(fin, np.isfinite(got)):
        return False
    return relerr(got[fin], ref[fin]) < tol if fin.any() else True


@need_golden
@pytest.mark.parametrize("case", cases(), ids=lambda c: c["name"])
def test_oracle_against_julia(case, cf, O):
    A = load_case(case)
    k = kernel_of(cf, case["kernel"], A)
    prog, op = k.program(), case["op"]
    if op == "mul_vec":
        X = A["X"]
        Y = A.get("Y", X)
        dt = X.dtype.type
        out_dt = A["b"].dtype.type
        b = O.mul_vec(prog, X.astype(out_dt), A["a"].astype(out_dt), Y=None if Y is X else Y.astype(out_dt), dtype=out_dt)
        assert relerr(b, A["b"]) < tol_of(A["b"]), dt
        if "y" in A:
            y = O.mul_vec(prog, X, A["a"], Y=None if Y is X else Y, alpha=case["alpha"], beta=case["beta"], y0=A["y0"])
            assert relerr(y, A["y"]) < 1e-12
            n, m = min(64, X.shape[0]), min(64, Y.shape[0])
            assert relerr(O.matrix(prog, X[:n], Y[:m]), A["M"].T) < 1e-13
    elif op == "mul_mat":
        assert relerr(O.mul_mat(prog, A["X"], A["A"].T), A["B"].T) < 1e-12
    elif op == "ladder":
        for i, r2 in enumerate(A["r2"].ravel()):
            v, d1, d2 = O.value_derivative_laplacian(prog, float(r2))
            assert v == pytest.approx(A["k"].ravel()[i], rel=1e-13)
            for got, ref in ((d1, A["k1"].ravel()[i]), (d2, A["k2"].ravel()[i])):
                assert (np.isnan(ref) and np.isnan(got)) or (np.isinf(ref) and got == ref) or got == pytest.approx(ref, rel=1e-10, abs=1e-300)
    elif op in ("gradient_mul", "value_gradient_mul"):
        trait = "dotproduct" if case["kernel"].startswith("Dot") else "isotropic"
        vg = op == "value_gradient_mul"
        b = O.derivative_mul(prog, A["X"], A["a"].ravel(), trait=trait, value_gradient=vg)
        assert nan_aware_close(b, A["b"].ravel(), 1e-12)
        if "y" in A:
            y = O.derivative_mul(prog, A["X"], A["a"].ravel(), trait=trait, alpha=case["alpha"], beta=case["beta"], y0=A["y0"].ravel())
            assert relerr(y, A["y"].ravel()) < 1e-12
    elif op == "solve":
        x, it, res, _ = O.cg_solve(prog, A["X"], A["y"].ravel(), case["sigma2"])
        assert relerr(x, A["x"].ravel()) < 1e-6  # cg! stops at reltol sqrt(eps): iterate-level agreement, not 1e-12
    elif op == "gradient_solve":
        x, it, res, _ = O.cg_solve(prog, A["X"], A["rhs"].ravel(), 0.0, gradient=True)
        assert relerr(x, A["x"].ravel()) < 1e-5


@need_golden
@pytest.mark.gpu
@pytest.mark.parametrize("case", cases(), ids=lambda c: c["name"])
def test_gpu_against_julia(case, cf):
    A = load_case(case)
    k = kernel_of(cf, case["kernel"], A)
    op = case["op"]
    if op == "mul_vec":
        X = A["X"]
        Y = A.get("Y", X)
        out_dt = A["b"].dtype.type
        G = cf.gramian(k, X.astype(out_dt).T.copy(), None if Y is X else Y.astype(out_dt).T.copy())
        assert relerr(G @ A["a"].astype(out_dt), A["b"]) < tol_of(A["b"])
        if "y" in A:
            y = A["y0"].copy()
            cf.mul_(y, G, A["a"], case["alpha"], case["beta"])
            assert relerr(y, A["y"]) < 1e-12
            n, m = min(64, X.shape[0]), min(64, Y.shape[0])
            assert relerr(cf.gramian(k, X[:n].T.copy(), Y[:m].T.copy()).Matrix(), A["M"].T) < 1e-13
    elif op == "mul_mat":
        assert relerr(cf.gramian(k, A["X"].T.copy()) @ A["A"].T, A["B"].T) < 1e-12
    elif op in ("gradient_mul", "value_gradient_mul"):
        wrap = cf.ValueGradientKernel if op == "value_gradient_mul" else cf.GradientKernel
        G = cf.gramian(wrap(k), A["X"].T.copy())
        assert nan_aware_close(G @ A["a"].ravel(), A["b"].ravel(), 1e-12)
        if "y" in A:
            y = A["y0"].ravel().copy()
            cf.mul_(y, G, A["a"].ravel(), case["alpha"], case["beta"])
            assert relerr(y, A["y"].ravel()) < 1e-12
    elif op == "solve":
        x, it, res = (case["sigma2"] * cf.I(A["X"].shape[0]) + cf.gramian(k, A["X"].T.copy())).solve(A["y"].ravel())
        assert relerr(x, A["x"].ravel()) < 1e-6
    elif op == "gradient_solve":
        G = cf.gramian(cf.GradientKernel(k), A["X"].T.copy())
        x, it, res = (0.0 * cf.I(G.shape[0]) + G).solve(A["rhs"].ravel())
        assert relerr(x, A["x"].ravel()) < 1e-5
    elif op == "ladder":
        pytest.skip("scalar ladder: covered by the oracle comparison and tests/test_gpu_math.py")


def test_reader_round_trip(tmp_path, cf, O):
    """the manifest / raw-binary reader on a fixture written in the generator's format (column-major shapes, little-endian)"""
    rng = np.random.default_rng(0)
    X = rng.standard_normal((7, 3))  # n = 7 points, d = 3: Julia shape [3, 7]
    A = rng.standard_normal((7, 2))  # Julia n x p matrix, column-major
    X.astype("<f8").tofile(tmp_path / "t_X.bin")
    np.ascontiguousarray(A.T).astype("<f8").tofile(tmp_path / "t_A.bin")
    X.astype("<f4").tofile(tmp_path / "t_X32.bin")
    case = {"name": "t", "kernel": "EQ", "op": "mul_mat",
            "arrays": {"X": {"file": "t_X.bin", "dtype": "f64", "shape": [3, 7]}, "A": {"file": "t_A.bin", "dtype": "f64", "shape": [7, 2]},
                       "X32": {"file": "t_X32.bin", "dtype": "f32", "shape": [3, 7]}}}
    got = load_case(case, base=str(tmp_path))
    assert np.array_equal(got["X"], X) and np.array_equal(got["A"].T, A) and got["X32"].dtype == np.float32
    assert relerr(O.mul_mat(kernel_of(cf, "EQ", got).program(), got["X"], got["A"].T), O.matrix(cf.EQ().program(), X) @ A) < 1e-13
