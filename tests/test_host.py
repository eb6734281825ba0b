"""CPU-only checks of the boundary and the host logic: the C-ABI library loads and exports every symbol the header
declares, kernel programs lower (or fail) with the reference's error classes, traits follow src/properties.jl, the
mirror validates shapes before touching the device, and -- without a GPU -- every compute call fails loudly."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(cf):
    hdr = open(os.path.join(ROOT, "include", "covfn_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cf_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 17
    lib = C.CDLL(cf.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/covfn_b200.h but not exported"
    assert declared == set(cf.SYMBOLS), "Python binding table and header disagree"
    assert cf.lib().cf_version() == 100


def test_op_numbering_matches_oracle(cf):
    from covfn_b200 import kernels as K
    from oracle import oracle as O

    hdr = open(os.path.join(ROOT, "include", "covfn_b200.h")).read()
    for name in ("EQ", "EXP", "RQ", "MATERNP", "DOT", "CONST", "SUM", "PROD", "POW", "LENGTHSCALE"):
        v = int(re.search(rf"CF_OP_{name}\s*=\s*(\d+)", hdr).group(1))
        assert getattr(K, f"OP_{name}") == v == getattr(O, f"OP_{name}")


def _create(cf, prog, d=3, n=4, dtype=1):
    from covfn_b200._lib import KNode, check

    arr = (KNode * max(1, len(prog)))()
    for t, (op, ip, fp) in enumerate(prog):
        arr[t].op, arr[t].iparam, arr[t].fparam = op, ip, fp
    X = np.zeros((n, max(d, 1)))
    h = C.c_void_p()
    rc = cf.lib().cf_gramian_create(C.byref(h), arr, len(prog), dtype, d, n, X.ctypes.data_as(C.c_void_p), max(d, 1), n, None, 0)
    if rc == 0:
        cf.lib().cf_gramian_destroy(h)
    check(rc)


def test_create_error_classes(cf):
    # argument / lowering errors are raised before the device is touched, so they are checkable without a GPU
    with pytest.raises(cf.DomainError):
        _create(cf, [(3, 0, -1.0)])  # RQ alpha <= 0 (src/stationary.jl:47)
    with pytest.raises(cf.DomainError):
        _create(cf, [(4, -1, 0.0)])  # MaternP(-1) (src/stationary.jl:124, test/stationary.jl:86)
    with pytest.raises(cf.DomainError):
        _create(cf, [(1, 0, 0.0), (10, 0, 0.0)])  # Lengthscale l <= 0 (src/transformation.jl:10)
    with pytest.raises(cf.DomainError):
        _create(cf, [(6, 0, -2.0)])  # Constant not PSD (src/stationary.jl:17-21)
    with pytest.raises(cf.UnsupportedKernel):
        _create(cf, [(42, 0, 0.0)])
    with pytest.raises(cf.UnsupportedKernel):
        _create(cf, [(5, 0, 0.0), (10, 0, 1.0)])  # Lengthscale of a non-isotropic kernel
    with pytest.raises(cf.CovFnError):
        _create(cf, [(1, 0, 0.0), (1, 0, 0.0)])  # stack does not reduce
    with pytest.raises(cf.CovFnError):
        _create(cf, [(7, 2, 0.0)])  # Sum without operands
    with pytest.raises(cf.DimensionMismatch):
        _create(cf, [(1, 0, 0.0)], d=0)
    with pytest.raises(cf.UnsupportedKernel):
        _create(cf, [(1, 0, 0.0)], d=(1 << 20) + 1, n=0)  # point dimension beyond the library's limit
    with pytest.raises(cf.CovFnError):
        _create(cf, [(1, 0, 0.0)], dtype=7)


def test_no_cpu_fallback(cf):
    if cf.device_count() > 0:
        pytest.skip("a GPU is present")
    G = cf.gramian(cf.EQ(), np.random.default_rng(0).standard_normal((3, 10)))
    with pytest.raises(cf.CudaError):
        G @ np.ones(10)
    with pytest.raises(cf.CudaError):
        cf.peak_probe("dfma", 10)
    with pytest.raises(cf.CudaError):
        cf.init([0])


def test_python_mirror_constructors_follow_reference(cf):
    with pytest.raises(cf.DomainError):
        cf.RQ(0)
    with pytest.raises(cf.DomainError):
        cf.MaternP(-1)
    with pytest.raises(cf.DomainError):
        cf.Constant(-1.0)
    with pytest.raises(cf.DomainError):
        cf.Lengthscale(cf.EQ(), 0.0)
    with pytest.raises(TypeError):
        cf.Lengthscale(cf.Dot(), 1.0)
    for p in range(5):
        assert isinstance(cf.MaternP(cf.Matern(p + 0.5)), cf.MaternP)  # test/stationary.jl:90-92
    assert cf.MaternP(cf.Matern(2.5)).p == 2


def test_input_traits(cf):
    # test/properties.jl:10-32 / test/gradient_algebra.jl:13-31 and src/properties.jl:39-63
    iso, dot, gen = cf.IsotropicInput(), cf.DotProductInput(), cf.GenericInput()
    assert cf.input_trait(cf.EQ()) == iso
    assert cf.input_trait(cf.Dot()) == dot
    assert cf.input_trait(cf.Dot() ** 3) == dot
    assert cf.input_trait(cf.EQ() + cf.RQ(1.0)) == iso
    assert cf.input_trait(cf.EQ() * cf.MaternP(2)) == iso
    assert cf.input_trait(2 * cf.EQ() + 1) == iso          # constants are ignored
    assert cf.input_trait(cf.Dot() + 1.0) == dot
    assert cf.input_trait(cf.EQ() + cf.Dot()) == gen       # mixed -> generic
    assert cf.input_trait(cf.Constant(1.0) * cf.Constant(2.0)) == iso
    assert cf.GradientKernel(cf.EQ()).input_trait() == iso
    assert cf.GradientKernel(cf.Dot() ** 3).input_trait() == dot


def test_programs(cf):
    assert cf.EQ().program() == [(1, 0, 0.0)]
    assert cf.RQ(2).program() == [(3, 1, 2.0)]        # Int alpha -> integer power path
    assert cf.RQ(2.0).program() == [(3, 0, 2.0)]
    assert cf.Poly(3, 1.0).program() == [(5, 0, 0.0), (6, 0, 1.0), (7, 2, 0.0), (9, 3, 0.0)]
    k = 0.5 * cf.RQ(2) + cf.Dot() ** 2                # README.md:78-80
    assert k.program() == [(6, 0, 0.5), (3, 1, 2.0), (8, 2, 0.0), (5, 0, 0.0), (9, 2, 0.0), (7, 2, 0.0)]
    assert cf.Lengthscale(cf.MaternP(2), 1.5).program() == [(4, 2, 0.0), (10, 0, 1.5)]
    assert (2 * cf.EQ()).program()[0] == (6, 1, 2.0)  # Constant{Int}


def test_gramian_shapes_and_validation(cf):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((3, 10))                   # d x n matrix, columns are points (src/gramian.jl:154)
    G = cf.gramian(cf.EQ(), X)
    assert G.shape == (10, 10) and G.size(1) == 10 and G.d == 3 and G.issymmetric()
    G2 = cf.gramian(cf.EQ(), [rng.standard_normal(3) for _ in range(4)], [rng.standard_normal(3) for _ in range(6)])
    assert G2.shape == (4, 6) and not G2.issymmetric() and G2.T.shape == (6, 4)
    G1 = cf.gramian(cf.EQ(), rng.standard_normal(8))   # vector of scalars: d = 1 (test/gramian.jl:12-14)
    assert G1.shape == (8, 8) and G1.d == 1
    Gg = cf.gramian(cf.GradientKernel(cf.EQ()), X)
    assert Gg.shape == (30, 30)                        # (d n) x (d n) (test/gradient.jl:35)
    assert cf.gramian(X[:, :2].T.tolist(), X[:, :2].T.tolist()).k.program() == cf.Dot().program()  # gramian(x, y) = Gramian(Dot(), x, y)
    with pytest.raises(cf.DimensionMismatch):
        cf.gramian(cf.EQ(), rng.standard_normal((3, 4)), rng.standard_normal((2, 4)))
    with pytest.raises(cf.DimensionMismatch):
        cf.gramian(cf.EQ(), [np.zeros(2), np.zeros(3)])
    # mul_ validates shapes and dtypes before any device call
    with pytest.raises(cf.DimensionMismatch):
        cf.mul_(np.zeros(10), G, np.zeros(9))
    with pytest.raises(cf.DimensionMismatch):
        cf.mul_(np.zeros(9), G, np.zeros(10))
    with pytest.raises(cf.DimensionMismatch):
        cf.mul_(np.zeros((10, 2), order="F"), G, np.zeros((10, 3), order="F"))
    with pytest.raises(TypeError):
        cf.mul_(np.zeros(10, dtype=np.float32), G, np.zeros(10))
    with pytest.raises(cf.DimensionMismatch):
        G.set_row_range(5, 11)
    assert G.set_row_range(2, 7).row_range == (2, 7)
    D = 1e-6 * cf.I(10)
    G = cf.gramian(cf.EQ(), X)
    assert isinstance(D + G, cf.LazyMatrixSum) and isinstance(G + D, cf.LazyMatrixSum)  # test/gramian.jl:51-53
    with pytest.raises(cf.DimensionMismatch):
        cf.I(9) + G


def test_float32_gramian_eltype(cf):
    X = np.zeros((3, 5), dtype=np.float32)
    assert cf.gramian(cf.EQ(), X).eltype == np.float32  # promote(Union{}, Float32) (src/gramian.jl:30-33)
    assert cf.gramian(cf.EQ(), X.astype(np.float64)).eltype == np.float64


def test_runtime_specialisations_compile_without_a_gpu():
    """the generated evaluators (csrc/cf_jit.h) must build with NVRTC for sm_100a for every kernel family that can be
    specialised, Float64 and Float32 -- checked here on the CPU, through the C ABI (cf_jit_check); the GPU tests then
    check that they compute the right numbers"""
    import covfn_b200 as cf
    from covfn_b200._lib import UnsupportedKernel

    programs = {
        "config3": 0.5 * cf.RQ(2) + cf.Dot() ** 2,
        "matern_times_rq_plus_const": cf.MaternP(3) * cf.RQ(1.5) + cf.Lengthscale(cf.EQ(), 0.7) + 0.25,
        "poly_of_sum": (cf.EQ() + cf.Exp()) ** 2 + (cf.Dot() + 1.0) ** 3,
    }
    try:
        cf.jit_check(cf.EQ(), 3, "mvm")
    except UnsupportedKernel as e:  # no libnvrtc on this machine: nothing to check
        pytest.skip(str(e))
    for name, k in programs.items():
        cf.jit_check(k, 3, "mvm")                 # K1, d = 3 (R = 4)
        for which in ("mm_dmma", "mvm_dmma", "mm_tf32", "mvm_tf32", "mm_tf32_legacy", "mvm_tc5"):  # incl. the tcgen05 kernels (K4u, K1u)
            cf.jit_check(k, 16, which)
    with pytest.raises(UnsupportedKernel):
        cf.jit_check(cf.EQ(), 3, "mm_dmma")       # no tensor-core kernel below d = 8
    # derivative programs: generated jets (product rule) for the tensor-core gradient kernel
    cf.jit_check(cf.EQ() + 0.5 * cf.RQ(2) * cf.MaternP(3), 16, "grad_dmma")
    cf.jit_check(cf.Lengthscale(cf.MaternP(2), 0.7) ** 2 + cf.RQ(1.5), 8, "grad_dmma")


def test_program_limits_are_reported_not_truncated():
    """the C++ lowering (csrc/cf_lower.h) runs on the host: programs beyond the by-value limits (8 terms, 6 atoms, 4 factors)
    must come back as CF_ERR_UNSUPPORTED so that the Julia shim falls through to the reference method"""
    import covfn_b200 as cf
    from covfn_b200._lib import UnsupportedKernel

    try:
        cf.jit_check(cf.EQ(), 3, "mvm")
    except UnsupportedKernel as e:
        pytest.skip(str(e))
    too_many_terms = (cf.EQ() + cf.RQ(1) + cf.MaternP(2) + cf.Lengthscale(cf.EQ(), 0.5)) ** 3  # 20 terms after expansion
    with pytest.raises(UnsupportedKernel):
        cf.jit_check(too_many_terms, 3, "mvm")
    too_many_atoms = sum((cf.Lengthscale(cf.EQ(), 0.1 * (i + 1)) for i in range(1, 8)), cf.EQ())  # 8 distinct atoms
    with pytest.raises(UnsupportedKernel):
        cf.jit_check(too_many_atoms, 3, "mvm")
    ok = cf.jit_check((cf.EQ() + cf.RQ(2)) ** 2 + 0.5, 3, "mvm")  # 4 terms, 2 atoms: fine
    assert isinstance(ok, str)


def test_lowering_merges_like_terms_and_lengthscaled_constants(cf):
    """(a + b)^3 expands to 4 distinct terms, (a + b)^4 to 5 (within the 8-term device limit); Lengthscale(Constant(c), l) is the
    constant (Constant <: IsotropicKernel, reference src/stationary.jl:15, src/transformation.jl:6-19).  Checked through cf_jit_check,
    which lowers the program without a GPU."""
    k4 = (cf.EQ() + cf.MaternP(2)) ** 4
    try:
        cf.jit_check(k4, 3, "mvm")  # lowering succeeds (16 raw terms merge into 5); NVRTC availability is a separate matter
    except cf.UnsupportedKernel as e:
        assert "NVRTC" in str(e) or "nvrtc" in str(e), e
    k5 = (cf.EQ() + cf.MaternP(2) + cf.RQ(2)) ** 4  # 15 distinct terms: beyond the device program
    with pytest.raises(cf.UnsupportedKernel):
        cf.jit_check(k5, 3, "mvm")
    kc = cf.Lengthscale(cf.Constant(2.0), 0.5) * cf.EQ()
    try:
        cf.jit_check(kc + cf.RQ(2), 3, "mvm")
    except cf.UnsupportedKernel as e:
        assert "NVRTC" in str(e) or "nvrtc" in str(e), e


def test_host_api_promotion_rules(cf):
    """gramian eltype = promote_type over x and y (src/gramian.jl:30-33); gramian(x) is the Dot Gramian (src/gramian.jl:151)"""
    x32 = np.ones((3, 5), dtype=np.float32)
    y64 = np.ones((3, 4), dtype=np.float64)
    assert cf.gramian(cf.EQ(), x32, y64).eltype == np.float64
    assert cf.gramian(cf.EQ(), x32).eltype == np.float32
    G = cf.gramian(x32)
    assert isinstance(G.k, cf.Dot) and G.shape == (5, 5)
