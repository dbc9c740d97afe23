"""``imq_KSD`` of sgmcmcjax/ksd.py:43-66 on the fused CUDA path.

``imq_KSD(samples, grads)`` = sqrt(sum_{i,j} k_0(x_i, x_j, g_i, g_j)) / N over all N^2 ordered
pairs with the IMQ kernel c = 1, beta = -1/2 (ksd.py:54).  The N^2 sum is accumulated in fp64
(the reference's fp32 running sum, ksd.py:60-65, is reproduced only up to its own rounding).

Multi-GPU: with ``torch.distributed`` initialised and ``distributed=True`` each rank sums the
Gram row block ``[rank*N/W, (rank+1)*N/W)`` and ONE double is all-reduced.
"""
from __future__ import annotations

import ctypes as C

from . import _native
from ._engine import to_device
from .kernels import _like


def row_block(n: int, rank: int, world: int):
    """Contiguous row-block shard of the Gram matrix owned by `rank`."""
    per = (n + world - 1) // world
    return min(rank * per, n), min((rank + 1) * per, n)


def imq_KSD_partial(samples, grads, row0: int, row1: int):
    """Sum of k_0 over rows [row0, row1) x all columns -> CUDA float64 tensor [1]."""
    torch = _native.require_cuda()
    lib = _native.load()
    x = to_device(samples, torch.float32)
    g = to_device(grads, torch.float32)
    if x.dim() != 2 or x.shape != g.shape:
        raise ValueError("samples and grads must both be [N, D]")
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _native.check(lib.sgmcmc_ksd_imq_partial(C.c_void_p(x.data_ptr()), C.c_void_p(g.data_ptr()), int(x.shape[0]),
                                              int(x.shape[1]), int(row0), int(row1), C.c_void_p(out.data_ptr()),
                                              stream))
    return out


def imq_KSD(samples, grads, distributed: bool = False):
    """Kernel Stein Discrepancy with the IMQ kernel (ksd.py:43-66)."""
    torch = _native.require_cuda()
    n = int(samples.shape[0])
    if distributed:
        import torch.distributed as dist

        r0, r1 = row_block(n, dist.get_rank(), dist.get_world_size())
        part = imq_KSD_partial(samples, grads, r0, r1)
        dist.all_reduce(part)
    else:
        part = imq_KSD_partial(samples, grads, 0, n)
    val = (torch.sqrt(part[0]) / n).to(torch.float32)
    return _like(val, samples)
