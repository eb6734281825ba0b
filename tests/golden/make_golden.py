"""Generates tests/golden/golden_v1.npz with the CPU oracle (the reference itself cannot run here: no Julia, and its
tests store no vectors).  Seeds are fixed; rerun with `python tests/golden/make_golden.py` after an oracle change and
commit the result.  Each case stores inputs and outputs so that the GPU tests need neither the oracle nor /root/reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import covfn_b200 as cf  # noqa: E402  (kernel objects -> programs only; no device call)
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (kernel, d, n, m, nrhs, gradient)
    "c1_maternp2_d3": (cf.MaternP(2), 3, 96, 96, 1, False),           # BASELINE config 1 shape
    "c2_eq_d3": (cf.EQ(), 3, 80, 144, 1, False),                       # config 2 kernel, rectangular
    "c3_rq_dot_d32_rhs8": (0.5 * cf.RQ(2) + cf.Dot() ** 2, 32, 48, 48, 8, False),  # config 3 kernel
    "c4_grad_eq_d16": (cf.EQ(), 16, 24, 24, 1, True),                  # config 4
    "c5_maternp2_d8": (cf.MaternP(2), 8, 64, 64, 1, False),            # config 5 operator
    "exp_d2": (cf.Exp(), 2, 50, 70, 1, False),
    "rq25_d4": (cf.RQ(2.5), 4, 40, 40, 1, False),
    "poly3_d5": (cf.Poly(3, 1.0), 5, 30, 45, 1, False),
    "grad_maternp3_d5": (cf.MaternP(3), 5, 20, 31, 1, True),
    "ls_eq_d3": (cf.Lengthscale(cf.EQ(), 0.5), 3, 33, 33, 1, False),
}


def main():
    out = {}
    for idx, (name, (k, d, n, m, nrhs, grad)) in enumerate(CASES.items()):
        rng = np.random.Generator(np.random.Philox(0xC0F00000 + idx))
        scale = 1.0 if d <= 3 else 1.0 / np.sqrt(d)
        X = rng.standard_normal((n, d)) * scale
        sym = n == m and name not in ("c2_eq_d3",)
        Y = X if sym else rng.standard_normal((m, d)) * scale
        blk = d if grad else 1
        a = rng.standard_normal(m * blk) if nrhs == 1 else rng.standard_normal((m * blk, nrhs))
        y0 = rng.standard_normal(n * blk) if nrhs == 1 else rng.standard_normal((n * blk, nrhs))
        alpha, beta = 0.75, -1.25
        prog = k.program()
        if grad:
            b = O.gradient_mul(prog, X, a, Y=None if sym else Y)
            b2 = O.gradient_mul(prog, X, a, Y=None if sym else Y, alpha=alpha, beta=beta, y0=y0)
        elif nrhs == 1:
            b = O.mul_vec(prog, X, a, Y=None if sym else Y)
            b2 = O.mul_vec(prog, X, a, Y=None if sym else Y, alpha=alpha, beta=beta, y0=y0)
        else:
            b = O.mul_mat(prog, X, a, Y=None if sym else Y)
            b2 = O.mul_mat(prog, X, a, Y=None if sym else Y, alpha=alpha, beta=beta, B0=y0)
        out[f"{name}/prog"] = np.array(prog, dtype=np.float64)
        out[f"{name}/X"], out[f"{name}/Y"], out[f"{name}/a"], out[f"{name}/y0"] = X, Y, a, y0
        out[f"{name}/sym"] = np.array(sym)
        out[f"{name}/grad"] = np.array(grad)
        out[f"{name}/b"], out[f"{name}/b_alpha_beta"] = b, b2
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz"), **out)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
