"""The reference's own published timings (README.md `@time` transcripts) re-run on the B200 path.
Each case reports the end-to-end host-API time of `mul!` (vectors in host memory; the handle already holds the points,
as `K = gramian(k, x)` precedes `@time mul!(b, K, a)` in the README) and the device time of the kernels."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import covfn_b200 as cf  # noqa: E402


def timed_mul(G, a, reps=3):
    b = np.zeros(G.shape[0])
    cf.mul_(b, G, a)  # warm-up (creates the handle)
    ts, ks = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        cf.mul_(b, G, a)
        ts.append(time.perf_counter() - t0)
        ks.append(G.last_timing()[0] * 1e-3)
    return min(ts), min(ks), b


out = []
rng = np.random.default_rng(0)
cases = [
    ("mul!(b,K,a) MaternP(2), d=3, n=16384", "README.md:26-38", 0.584813, cf.MaternP(2), 3, 16384, False),
    ("mul!(b,G,a) EQ, d=32, n=16384", "README.md:369-395", 0.949835, cf.EQ(), 32, 16384, False),
    ("mul!(b,G,a) EQ, d=2, n=65536", "README.md:409-430", 4.938038, cf.EQ(), 2, 65536, False),
    ("GradientKernel(MaternP(2)) mul!, d=1024, n=1024", "README.md:231-245", 0.394388, cf.MaternP(2), 1024, 1024, True),
    ("GradientKernel(EQ) mul!, d=1, n=1024 (figure)", "images/gradient_kernel_mvm_comparison.png", 0.019, cf.EQ(), 1, 1024, True),
    ("GradientKernel(EQ) mul!, d=16, n=1024 (figure)", "images/gradient_kernel_mvm_comparison.png", 0.025, cf.EQ(), 16, 1024, True),
    ("GradientKernel(EQ) mul!, d=64, n=1024 (figure)", "images/gradient_kernel_mvm_comparison.png", 0.043, cf.EQ(), 64, 1024, True),
    ("GradientKernel(EQ) mul!, d=1024, n=1024 (figure)", "images/gradient_kernel_mvm_comparison.png", 0.47, cf.EQ(), 1024, 1024, True),
    ("GradientKernel(EQ) mul!, d=16384, n=1024 (figure)", "images/gradient_kernel_mvm_comparison.png", 7.3, cf.EQ(), 16384, 1024, True),
]
for name, src, ref_s, k, d, n, grad in cases:
    X = rng.standard_normal((n, d))  # x = [randn(d) for _ in 1:n], as in the README
    a = rng.standard_normal(n * (d if grad else 1))
    G = cf.gramian(cf.GradientKernel(k) if grad else k, X.T)
    host_s, dev_s, b = timed_mul(G, a)
    out.append({"case": name, "reference_source": src, "reference_seconds": ref_s, "host_api_seconds": host_s,
                "device_seconds": dev_s, "speedup_vs_published": ref_s / host_s})
    G.close()
# G \ a (README.md:255-258): iterative solve with the d = n = 1024 gradient operator
n = d = 1024
X = rng.standard_normal((n, d))
a = rng.standard_normal(n * d)
G = cf.gramian(cf.GradientKernel(cf.MaternP(2)), X.T)
A = 0.0 * cf.I(n * d) + G
A.solve(a, maxiter=2)
t0 = time.perf_counter()
x, iters, res = A.solve(a)
dt = time.perf_counter() - t0
ok = float(np.linalg.norm((G @ x) - a) / np.linalg.norm(a))
out.append({"case": "G \\ a, GradientKernel(MaternP(2)), d=n=1024", "reference_source": "README.md:255-258", "reference_seconds": 0.817458,
            "host_api_seconds": dt, "iterations": iters, "relative_residual": ok, "speedup_vs_published": 0.817458 / dt})
print(json.dumps(out, indent=1))
