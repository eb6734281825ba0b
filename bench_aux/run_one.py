"""Run one workload a few times (device-resident) and print kernel time / pairs per second.  Used under ncu and for tuning.
    python bench_aux/run_one.py --config c2 [--n N] [--reps R] [--dtype f64|f32]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
import covfn_b200 as cf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--dtype", default="f64")
args = ap.parse_args()
w = bench.workload(args.config)
if args.n:
    w["n"] = args.n
X, a = bench.make_inputs(w)
dt = np.float64 if args.dtype == "f64" else np.float32
X, a = X.astype(dt), a.astype(dt)
n, d, nrhs = w["n"], w["d"], w["nrhs"]
blk = d if w["gradient"] else 1
k = cf.GradientKernel(w["kernel"]) if w["gradient"] else w["kernel"]
G = cf.gramian(k, X.T)
G.handle()
dev = torch.device("cuda", 0)
a_dev = torch.from_numpy(np.ascontiguousarray(a.T if nrhs > 1 else a)).to(dev)
b_dev = torch.empty((nrhs, n * blk) if nrhs > 1 else (n * blk,), dtype=a_dev.dtype, device=dev)
times = []
for _ in range(args.reps):
    G.mul_device(b_dev.data_ptr(), a_dev.data_ptr(), nrhs=nrhs, ldy=n * blk, ldx=n * blk)
    ms, launches = G.last_timing()
    times.append(ms)
ms = min(times)
pairs = float(n) * n
slots = bench.SLOTS[args.config]
print(json.dumps({"config": args.config, "n": n, "dtype": args.dtype, "ms": ms, "all_ms": times, "pairs_per_s": pairs / ms * 1e3,
                  "ref_slots_per_s": slots * pairs / ms * 1e3, "launches": launches}))
