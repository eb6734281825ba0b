"""Row-sharded Gramian products across processes (one process per GPU, torch.distributed for the plumbing).

Rows of K are contiguous blocks, one per rank; x and the weight vector are replicated (the reference
already gives each thread whole rows, src/gramian.jl:81,244).  A single MVM needs no collective on the data path; chained
MVMs (conjugate gradients, reference src/lazy_linear_algebra.jl:126-144) rebuild the full product with ONE all-gather
per iteration (NCCL over NVLink on GPUs; gloo in the CPU tests).  CG scalars are computed redundantly on every rank from the
gathered vectors, so all ranks take bit-identical decisions and no all-reduce is needed.

`local_mul(u) -> K[r0:r1, :] @ u` is injected: on the GPU it is `Gramian.mul_device` on this rank's handle; the CPU tests
inject the oracle so that the partition / gather / CG logic is exercised without a GPU.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


def row_block(n: int, rank: int, world: int, block: int = 1):
    """rows (points) [r0, r1) owned by `rank`; identical to the split in capi.cu (split_rows)"""
    r0 = n * rank // world
    r1 = n * (rank + 1) // world
    return r0, r1


def counts(n: int, world: int, block: int = 1):
    return [(row_block(n, r, world)[1] - row_block(n, r, world)[0]) * block for r in range(world)]


class ShardedOperator:
    """y = (sigma2 I + K) u with K row-sharded over the default process group."""

    def __init__(self, n: int, local_mul: Callable[[torch.Tensor], torch.Tensor], sigma2: float = 0.0, block: int = 1,
                 group=None):
        self.n, self.block, self.sigma2, self.local_mul, self.group = n, block, float(sigma2), local_mul, group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.r0, self.r1 = row_block(n, self.rank, self.world)
        self._counts = counts(n, self.world, block)

    def gather(self, local: torch.Tensor) -> torch.Tensor:
        """all-gather of the row blocks into the full vector (the one collective of a chained MVM)"""
        if self.world == 1:
            return local
        if len(set(self._counts)) == 1:
            out = torch.empty(self.n * self.block, dtype=local.dtype, device=local.device)
            dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
            return out
        # ragged split: pad every block to the largest one (collectives need equal sizes), trim after the gather
        cmax = max(self._counts)
        padded = torch.zeros(cmax, dtype=local.dtype, device=local.device)
        padded[: local.numel()] = local
        out = torch.empty(self.world * cmax, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, padded, group=self.group)
        return torch.cat([out[r * cmax: r * cmax + c] for r, c in enumerate(self._counts)])

    def apply(self, u: torch.Tensor) -> torch.Tensor:
        """full (sigma2 I + K) u on every rank"""
        loc = self.local_mul(u)
        if self.sigma2 != 0.0:
            b = self.block
            loc = loc + self.sigma2 * u[self.r0 * b:self.r1 * b]
        return self.gather(loc)


def cg(op: ShardedOperator, b: torch.Tensor, x0: torch.Tensor | None = None, reltol: float = 0.0, maxiter: int = 0):
    """IterativeSolvers.cg! 0.9.2 [upstream] restated for a sharded operator; every rank holds the full iterates.
    Returns (x, iterations, residual norm)."""
    N = b.numel()
    if reltol <= 0:
        reltol = float(torch.finfo(b.dtype).eps) ** 0.5
    if maxiter <= 0:
        maxiter = N
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    u = torch.zeros_like(b)
    r = b - op.apply(x)
    residual = float(torch.linalg.vector_norm(r))
    prev = 1.0
    tol = reltol * residual
    it = 0
    while residual > tol and it < maxiter:
        beta = residual**2 / prev**2
        u = r + beta * u
        c = op.apply(u)
        alpha = residual**2 / float(torch.dot(u, c))
        x = x + alpha * u
        r = r - alpha * c
        prev = residual
        residual = float(torch.linalg.vector_norm(r))
        it += 1
    return x, it, residual


def gpu_local_mul(G):
    """local_mul for a covfn_b200 Gramian restricted to this rank's row block (device tensors, current stream)."""
    r0, r1 = G.row_range
    blk = G.block

    def f(u: torch.Tensor) -> torch.Tensor:
        out = torch.empty((r1 - r0) * blk, dtype=u.dtype, device=u.device)
        u = u.contiguous()
        st = torch.cuda.current_stream()
        if st.cuda_stream == 0:
            # torch's default stream has handle 0, which the C ABI reads as "use the library's own (non-blocking) stream and block
            # until done": that stream does not order itself after work queued on the legacy default stream, so whatever produced
            # `u` must have finished before the call (the call itself returns only when the product is complete)
            st.synchronize()
        G.mul_device(out.data_ptr(), u.data_ptr(), stream=st.cuda_stream)
        return out

    return f


def comm_init_from_torch(group=None):
    """Bootstrap the library's own NCCL communicator (csrc/cf_comm.h) from an initialised torch.distributed process group:
    rank 0 draws the NCCL unique id (cf_comm_unique_id), torch broadcasts its 128 bytes, every rank calls cf_comm_init with its
    CUDA device current.  Afterwards `(sigma2 * I(n) + G).solve(b)` on a row-restricted Gramian runs the multi-process CG
    inside the library (one NCCL all-gather per product).  Returns (rank, world)."""
    import ctypes as C

    from ._lib import check, lib

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        check(lib().cf_comm_unique_id(buf, 128))
    dev = torch.device("cuda", torch.cuda.current_device())
    t = torch.tensor(list(buf), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=0, group=group)
    raw = bytes(t.cpu().tolist())
    ident = (C.c_ubyte * 128).from_buffer_copy(raw)
    check(lib().cf_comm_init(ident, rank, world))
    return rank, world


def comm_destroy():
    from ._lib import check, lib

    check(lib().cf_comm_destroy())
