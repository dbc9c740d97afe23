"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed.

* Sampling: chains are independent units.  Chain c of the job always uses ``split(key, C_total)[c]``, rank r
  owns the contiguous block ``chain_shard(C_total, r, world)``; data is replicated per GPU and there is NO
  collective inside the sampling loop.  ``gather_chains`` is the optional all-gather of the samples at the end.
* Optional data-sharded mode (tables that do not fit one GPU): the ROWS are split over ranks, every rank runs every
  chain with the same keys and processes the minibatch rows it holds; one all-reduce of the [C, D] minibatch sums
  per step (``build_data_sharded_sampler``).
* KSD: the symmetric Gram tile list is dealt round-robin to ranks (``ksd_tile_partition`` mirrors
  csrc/ksd_tc.cuh), each rank produces one fp64 partial, ONE all-reduce of a double (``allreduce_scalar``).

Nothing here computes on the CPU: the callers pass in the per-rank partial results (from the CUDA path in the
product, from the oracle in the gloo CPU tests of this host logic).
"""
from __future__ import annotations

from typing import List, Tuple

KSD_TILE = 128     # csrc/ksd_tc.cuh TILE


def chain_shard(n_chains: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of the job's chains owned by `rank` (sizes differ by at most one)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_chains), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def ksd_tile_count(n: int) -> int:
    t = (int(n) + KSD_TILE - 1) // KSD_TILE
    return t * (t + 1) // 2


def ksd_tile_of(t: int, n: int) -> Tuple[int, int]:
    """Index t of the row-major upper-triangular tile list -> (I, J >= I)  (csrc/ksd_tc.cuh tile_of)."""
    T = (int(n) + KSD_TILE - 1) // KSD_TILE
    i, start = 0, 0
    while t >= start + (T - i):
        start += T - i
        i += 1
    return i, i + (t - start)


def ksd_tile_partition(n: int, part: int, nparts: int) -> List[Tuple[int, int]]:
    """Tiles summed by partition `part` of `nparts`: list indices part, part + nparts, ..."""
    return [ksd_tile_of(t, n) for t in range(part, ksd_tile_count(n), nparts)]


def allreduce_scalar(value, group=None):
    """SUM all-reduce of a one-element float64 tensor (in place) - the only KSD collective."""
    import torch.distributed as dist

    dist.all_reduce(value, op=dist.ReduceOp.SUM, group=group)
    return value


def gather_chains(local, n_chains_total: int, group=None):
    """All-gather per-rank sample blocks [c_local, ...] (chain_shard order) into [n_chains_total, ...]."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    sizes = [chain_shard(n_chains_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)


def prior_grad(spec, leaf_size, thetas):
    """grad logprior at flat parameter rows `thetas` [M, P] for a registered family (all priors are Gaussian:
    gaussian_mean -2 c_l theta_l per leaf, the others -precision * theta)."""
    from . import _native

    out = thetas.clone()
    if spec.family == _native.FAMILY_GAUSSIAN_MEAN:
        off = 0
        for c, n in zip(spec.prior_coef, leaf_size):
            out[:, off:off + n] *= -2.0 * float(c)
            off += n
        return out
    return out.mul_(-float(spec.prior_coef[0]))


def sharded_fullbatch_grad(local_grad, prior, group=None):
    """Full-batch grad_log_post (util.py:43-49 with all rows) when the ROWS of the data are sharded over ranks
    (SURVEY.md section 8e, K2): every rank passes `local_grad` = grad_log_post over ITS rows (which includes the
    prior once) and `prior` = grad logprior; one all-reduce of the [M, P] block, then the (world - 1) surplus
    copies of the prior are removed:  sum_r (L_r + p) - (W - 1) p = sum_r L_r + p."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    total = local_grad.clone()
    dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return total - (world - 1) * prior


def build_data_sharded_sampler(build_kernel_fn, *args, shards, n_total, group=None, distributed=None, **kwargs):
    """Sampler for a table whose ROWS are split into shards (north_star "optional data-sharded mode").

    ``build_kernel_fn(*args, **kwargs)`` is one of the ``kernels.build_*_kernel`` factories (vector families; standard or
    CV estimator; not SGHMC) whose ``data`` argument is replaced, shard by shard, with the entries of
    ``shards = [(row_offset, data_tuple), ...]`` - normally ONE entry: this rank's rows.  With ``torch.distributed``
    initialised (``distributed`` defaults to that) the [C, D] minibatch sums are all-reduced across ranks once per step
    (NCCL over NVLink); several local shards are summed on the device first.  The chains are identical on every rank
    and equal, up to fp32 summation order, to the unsharded sampler with the same keys."""
    import torch

    from . import _native, _tree, kernels
    from ._engine import keys_to_device

    if distributed is None:
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
    args = list(args)
    data_pos = next(k for k, a in enumerate(args) if type(a) is tuple)      # the `data` tuple of the reference signature
    engines = []
    for row0, data in shards:
        a = list(args)
        a[data_pos] = data
        kernels._row_shard = ((int(row0), int(n_total)), group if distributed else False)
        try:
            init_fn, _, _ = build_kernel_fn(*a, **kwargs)
        finally:
            kernels._row_shard = None
        engines.append(init_fn.engine)
    if len(engines) > 1 and engines[0].estimator == "cv" and not distributed:
        # several local shards and no process group: combine the centre's full-batch gradient here
        total = sum(e.fullbatch_grad(e.flatten_params(e._centering_value, None)[0]) for e in engines)
        center = engines[0].flatten_params(engines[0]._centering_value, None)[0]
        total = total - (len(engines) - 1) * prior_grad(engines[0].spec, engines[0].leaf_size, center)
        import ctypes as C
        for e in engines:
            _native.check(e.lib.sgmcmc_set_centering_grad(e._handle, C.c_void_p(center.data_ptr()), C.c_void_p(total.data_ptr()),
                                                          e._stream()))
        torch.cuda.current_stream().synchronize()
    palindrome = engines[0].diffusion in ("baoab", "badodab")

    def reduce_parts(parts):
        total = parts[0] if len(parts) == 1 else torch.stack(parts).sum(0)
        if distributed:
            import torch.distributed as dist
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        return total

    def sampler(key, Nsamples: int, params):
        e0 = engines[0]
        keys, kb = keys_to_device(key)
        flat, treedef, pb = e0.flatten_params(params, broadcast_chains=keys.shape[0])
        batched, nsamp, n_chains = kb or pb, int(Nsamples), flat.shape[0]
        states = [e.new_state(flat) for e in engines]
        total = reduce_parts([e.shard_init(st, keys, True) for e, st in zip(engines, states)])
        samples = torch.empty((nsamp, n_chains, e0.P), dtype=torch.float32, device="cuda")
        for i in range(nsamp):
            flags = _native.SHARD_STEP | (_native.SHARD_UPDATE2 if (palindrome and i > 0) else 0)
            parts = [e.shard_step(st, i, flags, total, samples[i] if k == 0 else None)
                     for k, (e, st) in enumerate(zip(engines, states))]
            total = reduce_parts(parts)
        for e, st in zip(engines, states):                                    # finish the last gradient (and update2)
            e.shard_step(st, nsamp, _native.SHARD_UPDATE2 if (palindrome and nsamp > 0) else 0, total)
        out = samples.permute(1, 0, 2).contiguous()
        lead = (n_chains, nsamp) if batched else (nsamp,)
        tree = e0.unflatten_params(out if batched else out[0], treedef, lead)
        leaves, td = _tree.flatten(tree)
        like = _tree.flatten(params)[0][0]
        sampler.states = states
        return _tree.unflatten(td, [kernels._like(l, like) for l in leaves])

    sampler.engines = engines
    return sampler
