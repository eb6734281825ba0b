"""Multi-process mode check (launched under torchrun by tests/test_gpu_multidevice.py::test_multi_process_comm_mode): the library's
own NCCL communicator (cf_comm_init), row-block product + cf_comm_allgather_rows, the collective symmetric product
(ncclAllReduce) and the multi-process CG, each against the oracle / the single-process result.  Prints COMM_CHECK_OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import covfn_b200 as cf  # noqa: E402
from covfn_b200 import distributed as D  # noqa: E402
from covfn_b200.gramian import mul_collective_device  # noqa: E402
from oracle import oracle as O  # noqa: E402  (the checker)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
assert D.comm_init_from_torch() == (rank, world)
O.set_num_threads(max(1, (os.cpu_count() or 2) // world))


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


rng = np.random.default_rng(2025)
n, d = 40001, 3  # ragged row blocks
X = rng.standard_normal((n, d))
a = rng.standard_normal(n)
r0, r1 = D.row_block(n, rank, world)
a_dev = torch.from_numpy(a).to(dev)
for k in (cf.EQ(), cf.MaternP(2)):
    G = cf.gramian(k, X.T.copy()).set_row_range(r0, r1)
    rows = (r1 - 40, min(n, r1 + 40)) if rank < world - 1 else (r0 - 40, r0 + 40)
    ref = O.mul_vec(k.program(), X, a, rows=rows)
    # (1) row block in place + the library's all-gather
    G.set_symmetric(False)
    b = torch.full((n,), float("nan"), dtype=torch.float64, device=dev)
    G.mul_device(b[r0:].data_ptr(), a_dev.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    cf.check(cf.lib().cf_comm_allgather_rows(b.data_ptr(), n, 1, 1, None))
    torch.cuda.synchronize()
    b_rows = b.cpu().numpy()
    assert rel(b_rows[rows[0]:rows[1]], ref) < 1e-12, "row blocks + all-gather"
    # (2) collective product: symmetric partial vectors + ncclAllReduce; complete and identical on every rank
    G.set_symmetric(True)
    y0 = rng.standard_normal(n)
    c = torch.from_numpy(y0).to(dev)
    mul_collective_device(G, c.data_ptr(), a_dev.data_ptr(), alpha=-0.5, beta=2.0)
    c = c.cpu().numpy()
    assert rel(c, -0.5 * b_rows + 2.0 * y0) < 1e-13, "collective symmetric product"
    allc = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(allc, torch.from_numpy(c).to(dev))
    assert all(torch.equal(allc[0], t) for t in allc), "ranks must hold identical bits"
# (3) multi-process CG inside the library, against the single-process solve of the same handle type
n2, d2 = 36000, 8
X2 = rng.standard_normal((n2, d2)) / np.sqrt(d2)
y = rng.standard_normal(n2)
q0, q1 = D.row_block(n2, rank, world)
for symm in (True, False):
    Gs = cf.gramian(cf.MaternP(2), X2.T.copy()).set_row_range(q0, q1)
    Gs.set_symmetric(symm)
    x, it, res = (1e-2 * cf.I(n2) + Gs).solve(y, reltol=1e-300, maxiter=6)
    Gf = cf.gramian(cf.MaternP(2), X2.T.copy())
    Gf.set_symmetric(symm)
    true_res = float(np.linalg.norm(y - (Gf @ x) - 1e-2 * x))
    assert it == 6 and abs(true_res - res) < 1e-8 * res, (it, res, true_res)
    chk = torch.from_numpy(x).to(dev)
    allx = [torch.empty_like(chk) for _ in range(world)]
    dist.all_gather(allx, chk)
    assert all(torch.equal(allx[0], t) for t in allx), "CG iterates must be bit-identical on every rank"
dist.barrier()
if rank == 0:
    print("COMM_CHECK_OK")
D.comm_destroy()
dist.destroy_process_group()
