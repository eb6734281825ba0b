"""K1e (gram_mvm_eq.cuh) against the oracle and against K1: parity on a row block, device time at n = 2^20.
    python bench_aux/k1e_check.py [--n N]"""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 20)
ap.add_argument("--d", type=int, default=3)
ap.add_argument("--child", default="")
args = ap.parse_args()

if not args.child:
    out = {}
    for mode in ("eq", "scalar"):
        env = dict(os.environ)
        if mode == "scalar":
            env["COVFN_MVM_SCALAR"] = "1"
        r = subprocess.run([sys.executable, __file__, "--n", str(args.n), "--d", str(args.d), "--child", mode], env=env,
                           capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-2000:])
    sys.exit(0)

import torch  # noqa: E402

import covfn_b200 as cf  # noqa: E402
from oracle import oracle as O  # noqa: E402

rng = np.random.default_rng(7)
n, d = args.n, args.d
X = rng.standard_normal((n, d))
a = rng.standard_normal(n)
res = {"mode": args.child, "n": n, "d": d}
for name, k in (("EQ", cf.EQ()), ("EQ_l0.7", cf.Lengthscale(cf.EQ(), 0.7)), ("2.5*EQ_l3", 2.5 * cf.Lengthscale(cf.EQ(), 3.0))):
    G = cf.gramian(k, X.T.copy())
    dev = torch.device("cuda", 0)
    a_dev = torch.from_numpy(a).to(dev)
    b_dev = torch.empty(n, dtype=torch.float64, device=dev)
    ts = []
    for _ in range(3):
        G.mul_device(b_dev.data_ptr(), a_dev.data_ptr(), nrhs=1, ldy=n, ldx=n)
        ts.append(G.last_timing()[0])
    b = b_dev.cpu().numpy()
    rows = (n // 2 - 128, n // 2 + 128)
    ref = O.mul_vec(k.program(), X, a, rows=rows)
    tru = O.truth_mul_vec(k.program(), X, a, rows=rows)
    got = b[rows[0]:rows[1]]
    res[name] = {"ms": min(ts), "pairs_per_s": float(n) * n / min(ts) * 1e3,
                 "err_vs_oracle": float(np.linalg.norm(got - ref) / np.linalg.norm(ref)),
                 "err_vs_truth": float(np.linalg.norm(got - tru) / np.linalg.norm(tru)),
                 "oracle_vs_truth": float(np.linalg.norm(ref - tru) / np.linalg.norm(tru))}
print(json.dumps(res))
