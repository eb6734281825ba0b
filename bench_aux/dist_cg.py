"""BASELINE config 5 under torchrun: CG on (K + sigma2 I) x = y with K = gramian(MaternP(2), X), d = 8, n = 2^19,
rows of K sharded over the ranks, ONE NCCL all-gather of the product per iteration (covfn_b200.distributed).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_aux/dist_cg.py [--n N] [--iters K]
Prints one JSON line on rank 0: per-iteration time, pairs/s, residual check against a separately computed true residual."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import covfn_b200 as cf  # noqa: E402
from covfn_b200 import distributed as D  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1 << 19)
ap.add_argument("--d", type=int, default=8)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--sigma2", type=float, default=1e-2)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n, d = args.n, args.d
rng = np.random.Generator(np.random.Philox(0xC0F00005))
X = rng.standard_normal((n, d)) / np.sqrt(d)
y = rng.standard_normal(n)
k = cf.MaternP(2)
r0, r1 = D.row_block(n, rank, world)
G = cf.gramian(k, X.T).set_row_range(r0, r1)
G.handle()
op = D.ShardedOperator(n, D.gpu_local_mul(G), sigma2=args.sigma2)
yt = torch.from_numpy(y).to(dev)
# warm-up product
op.apply(yt)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
x, it, res = D.cg(op, yt, maxiter=args.iters)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
true_res = float(torch.linalg.vector_norm(yt - op.apply(x)))
# every rank must hold bit-identical iterates (no all-reduce is used: scalars are recomputed from gathered vectors)
if world > 1:
    chk = torch.stack([x.sum(), x.abs().max()])
    allc = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    identical = all(torch.equal(allc[0], c) for c in allc)
else:
    identical = True
if rank == 0:
    mvms = it + 1  # initial residual + one per iteration
    print(json.dumps({"config": "c5", "n": n, "d": d, "n_gpus": world, "iterations": it, "seconds": dt,
                      "ms_per_iteration": 1e3 * dt / mvms, "pairs_per_s": mvms * float(n) * n / dt,
                      "recurrence_residual": res, "true_residual": true_res, "rhs_norm": float(np.linalg.norm(y)),
                      "ranks_bit_identical": bool(identical)}))
if world > 1:
    dist.destroy_process_group()
