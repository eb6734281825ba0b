"""Small exercise of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import covfn_b200 as cf  # noqa: E402

rng = np.random.default_rng(0)
for d in (3, 16):
    n, m = 700, 515
    X, Y = rng.standard_normal((n, d)) / np.sqrt(d), rng.standard_normal((m, d)) / np.sqrt(d)
    a = rng.standard_normal(m)
    for k in (cf.EQ(), cf.MaternP(2), cf.RQ(2), 0.5 * cf.RQ(2) + cf.Dot() ** 2):
        G = cf.gramian(k, X.T, Y.T)
        G @ a
        G @ rng.standard_normal((m, 5))
        G.Matrix()
    Gs = cf.gramian(cf.EQ(), X.T)
    Gs @ rng.standard_normal(n)
    for k in (cf.EQ(), cf.MaternP(2), cf.Dot() ** 3):
        cf.gramian(cf.GradientKernel(k), X.T, Y.T) @ rng.standard_normal(m * d)
        cf.gramian(cf.ValueGradientKernel(k), X.T, Y.T) @ rng.standard_normal(m * (d + 1))
    (1e-2 * cf.I(n) + Gs).solve(rng.standard_normal(n), maxiter=5)
Xf = rng.standard_normal((300, 3)).astype(np.float32)
cf.gramian(cf.EQ(), Xf.T) @ rng.standard_normal(300).astype(np.float32)
for k in (cf.MaternP(2), cf.RQ(2)):  # packed-FP32 kernel (gram_mvm_f32p.cuh), ragged against its 128-column tile
    cf.gramian(k, Xf.T, Xf[:211].T) @ rng.standard_normal(211).astype(np.float32)
cf.gramian(cf.GradientKernel(cf.EQ()), Xf.T) @ rng.standard_normal(900).astype(np.float32)  # Float64 shadow of a Float32 handle
Xf16 = (rng.standard_normal((515, 16)) / 4).astype(np.float32)  # Float32, d = 16: the tcgen05 kernels (gram_mm_tc5.cuh, gram_mvm_tc5.cuh)
for k in (cf.EQ(), cf.MaternP(2), 0.5 * cf.RQ(2) + cf.Dot() ** 2):
    Gf = cf.gramian(k, Xf16.T, Xf16[:300].T)
    Gf @ rng.standard_normal(300).astype(np.float32)
    Gf @ rng.standard_normal((300, 5)).astype(np.float32)
# many column tiles per CTA: the TMA stage ring and the TMEM buffers of the tcgen05 value kernel are re-used several times
# (300 row tiles x 3 column chunks of 16 tiles each; the ring has 4 stages)
Xl = (rng.standard_normal((128 * 300, 16)) / 4).astype(np.float32)
cf.gramian(cf.EQ(), Xl.T, Xl[:3000].T) @ rng.standard_normal(3000).astype(np.float32)
Xb = rng.standard_normal((150, 40)) / 6
cf.gramian(cf.EQ(), Xb.T) @ rng.standard_normal(150)
cf.gramian(cf.GradientKernel(cf.EQ()), Xb.T) @ rng.standard_normal(150 * 40)
if os.environ.get("CF_SAN_SYM"):  # symmetric variant at its minimum size (slow under the sanitizer)
    n = 65536
    Xs = rng.standard_normal((n, 3))
    cf.gramian(cf.EQ(), Xs.T).set_symmetric(True) @ rng.standard_normal(n)
print("sanitize_small: done")
