import numpy as np, os, sys
sys.path.insert(0, os.getcwd())
import covfn_b200 as cf
from oracle import oracle as O
rng = np.random.default_rng(5)
for d in (8, 12, 16, 24, 32):
    n, m, p = 301, 517, 5
    X = (rng.standard_normal((n, d)) / np.sqrt(d)).astype(np.float32)
    Y = (rng.standard_normal((m, d)) / np.sqrt(d)).astype(np.float32)
    A = rng.standard_normal((m, p)).astype(np.float32)
    for name, k in {"c3": 0.5 * cf.RQ(2) + cf.Dot() ** 2, "eq": cf.EQ(), "m2": cf.MaternP(2)}.items():
        G = cf.gramian(k, X.T.copy(), Y.T.copy())
        B = G @ A
        ref64 = O.mul_mat(k.program(), X.astype(np.float64), A.astype(np.float64), Y=Y.astype(np.float64))
        os.environ["COVFN_MM_SCALAR"] = "1"
        Bs = G @ A
        del os.environ["COVFN_MM_SCALAR"]
        e = lambda a, b: np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b)
        print(d, name, "tf32 vs f64 truth %.2e" % e(B, ref64), "scalar f32 vs truth %.2e" % e(Bs, ref64), B.dtype)
