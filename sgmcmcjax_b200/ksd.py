"""``imq_KSD`` of sgmcmcjax/ksd.py:43-66 on the fused CUDA path.

``imq_KSD(samples, grads)`` = sqrt(sum_{i,j} k_0(x_i, x_j, g_i, g_j)) / N over all N^2 ordered
pairs with the IMQ kernel c = 1, beta = -1/2 (ksd.py:54).  The N^2 sum is accumulated in fp64
(the reference's fp32 running sum, ksd.py:60-65, is reproduced only up to its own rounding).

Two device paths compute the same sum:

* ``method="tc"`` (default): Gram contractions on the tensor cores (tcgen05, 3xTF32 split, TMA-fed,
  accumulators in TMEM, IMQ epilogue fused) over the upper-triangular 128 x 128 tile list
  (csrc/ksd_tc.cuh).  Multi-GPU: tiles are dealt round-robin to ranks.
* ``method="simt"``: fp32 CUDA-core kernel that forms the differences explicitly like the reference
  does (csrc/ksd.cuh) - the accuracy anchor; shards by Gram row blocks ``[rank*N/W, (rank+1)*N/W)``.

With ``torch.distributed`` initialised and ``distributed=True`` each rank sums its share and ONE
double is all-reduced.
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


_workspaces = {}


def _workspace(torch, nbytes: int):
    dev = torch.cuda.current_device()
    ws = _workspaces.get(dev)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(nbytes + 256, dtype=torch.uint8, device="cuda")
        _workspaces[dev] = ws
    return ws


def imq_KSD_tiles(samples, grads, part: int = 0, nparts: int = 1):
    """Sum of k_0 over partition `part` of `nparts` of the symmetric Gram tile list (tensor cores)
    -> CUDA float64 tensor [1].  Summing over all partitions gives the sum over all N^2 ordered pairs."""
    torch = _native.require_cuda()
    lib = _native.load()
    x = to_device(samples, torch.float32)
    g = to_device(grads, torch.float32)
    if x.dim() != 2 or x.shape != g.shape:
        raise ValueError("samples and grads must both be [N, D]")
    n, d = int(x.shape[0]), int(x.shape[1])
    ws = _workspace(torch, int(lib.sgmcmc_ksd_workspace_bytes(n, d)))
    base = (ws.data_ptr() + 255) & ~255
    out = torch.zeros(1, dtype=torch.float64, device="cuda")
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _native.check(lib.sgmcmc_ksd_imq_tc(C.c_void_p(x.data_ptr()), C.c_void_p(g.data_ptr()), n, d, int(part), int(nparts),
                                         C.c_void_p(base), C.c_void_p(out.data_ptr()), stream))
    return out


def imq_KSD(samples, grads, distributed: bool = False, method: str = "tc"):
    """Kernel Stein Discrepancy with the IMQ kernel (ksd.py:43-66)."""
    torch = _native.require_cuda()
    n = int(samples.shape[0])
    rank, world = 0, 1
    if distributed:
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
    if method == "tc":
        part = imq_KSD_tiles(samples, grads, rank, world)
    elif method == "simt":
        part = imq_KSD_partial(samples, grads, *row_block(n, rank, world))
    else:
        raise ValueError("method must be 'tc' or 'simt'")
    if distributed:
        from . import parallel

        parallel.allreduce_scalar(part)
    val = (torch.sqrt(part[0]) / n).to(torch.float32)
    return _like(val, samples)


def posterior_KSD(engine, samples, distributed: bool = False):
    """imq_KSD of posterior samples [N, P] with their gradients computed here: ``grads = grad_log_post(samples)`` over ALL
    rows of the engine's data (K2; tensor cores for the logistic-regression family), then the Gram contractions.
    ``engine`` is ``sampler.engine`` / ``init_fn.engine`` of any builder of this package."""
    torch = _native.require_cuda()
    x = to_device(samples, torch.float32)
    if x.dim() != 2:
        raise ValueError("samples must be [N, P]")
    grads = engine.fullbatch_grad(x.contiguous())
    val = imq_KSD(x, grads, distributed=distributed)
    return _like(val, samples)
