"""The sampler factories of sgmcmcjax/samplers.py:98-136: ``build_*_sampler(...)`` takes exactly
the arguments of the matching ``build_*_kernel`` plus ``compiled`` / ``pbar`` and returns
``sampler(key, Nsamples, params) -> samples``.

The scan of samplers.py:45-58 runs as ONE kernel launch (K steps per launch, one CTA per chain);
``pbar=True`` splits it into at most 20 launches and advances a host-side tqdm bar between them
(the reference updates every Nsamples/20 iterations, util.py:130).  ``compiled=False`` keeps the
reference's return type for that mode: a Python list with one parameter pytree per sample
(samplers.py:81-93).  Multi-chain: ``key`` uint32[C,2] and parameter leaves [C,D] give samples with
leaves [C, Nsamples, D] (``vmap(sampler, in_axes=(0, None, 0))`` semantics).
"""
from __future__ import annotations

import functools
from typing import Callable

from . import _native, _tree
from ._engine import keys_to_device
from .kernels import (
    _like,
    build_badodab_kernel,
    build_badodabCV_kernel,
    build_baoab_kernel,
    build_psgld_kernel,
    build_sghmc_kernel,
    build_sghmc_SVRG_kernel,
    build_sghmcCV_kernel,
    build_sgld_kernel,
    build_sgld_SVRG_kernel,
    build_sgldAdam_kernel,
    build_sgldCV_kernel,
    build_sgnht_kernel,
    build_sgnhtCV_kernel,
)


_PIPELINE_BYTES = 8 << 20
_PIPELINE_CHUNKS = 5
_pinned = {}       # (device, shape) -> two page-locked staging tensors, reused across calls (never handed to the caller)


def _host_result(params) -> bool:
    import numpy as np

    like = _tree.flatten(params)[0][0]
    return isinstance(like, np.ndarray) or np.isscalar(like) or isinstance(like, (list, tuple))


def _run_pipelined(torch, eng, st, samples, nsamp: int, keep):
    """K steps as a few launches; block j's samples go device -> pinned staging (copy engine, side stream) -> the numpy
    result (host memcpy) while block j + 1 computes.  Returns float32 [C, nsamp, P] (numpy)."""
    import numpy as np

    n_chains, P = samples.shape[0], eng.P
    kc = -(-nsamp // _PIPELINE_CHUNKS)
    key = (torch.cuda.current_device(), n_chains, kc, P)
    if key not in _pinned:
        _pinned[key] = [torch.empty((n_chains, kc, P), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    stage = _pinned[key]
    out = np.empty((n_chains, nsamp, P), dtype=np.float32)
    main, side = torch.cuda.current_stream(), _side_stream(torch)
    pending = None                                                 # (event, slot, i0, k) of the block whose copy is in flight
    for j, i0 in enumerate(range(0, nsamp, kc)):
        k = min(kc, nsamp - i0)
        keep.append(eng.run(st, i0, k, samples[:, i0:], sample_chain_stride=nsamp * P))
        done = torch.cuda.Event()
        done.record(main)
        side.wait_event(done)
        with torch.cuda.stream(side):
            stage[j & 1][:, :k].copy_(samples[:, i0:i0 + k], non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(side)
        if pending is not None:                                     # unpack the previous block while this one runs
            ev, slot, p0, pk = pending
            ev.synchronize()
            out[:, p0:p0 + pk] = stage[slot][:, :pk].numpy()
        pending = (copied, j & 1, i0, k)
    ev, slot, p0, pk = pending
    ev.synchronize()
    out[:, p0:p0 + pk] = stage[slot][:, :pk].numpy()
    main.synchronize()
    return out


_side = {}


def _side_stream(torch):
    dev = torch.cuda.current_device()
    if dev not in _side:
        _side[dev] = torch.cuda.Stream()
    return _side[dev]


def _summaries_sampler(eng, thin: int, summaries: bool, store_samples: bool, pbar: bool) -> Callable:
    """Sampler with thinning and on-device posterior summaries (not in the reference, whose scan
    stacks every state, samplers.py:57).  ``sampler(key, Nsamples, params)`` runs Nsamples iterations and keeps the states
    after iterations thin, 2 thin, ...: returns the Nsamples // thin kept samples (None with ``store_samples=False``), or
    ``(samples, {"mean": tree, "var": tree, "count": n})`` with ``summaries=True`` - mean and marginal variance of the kept
    states per chain, accumulated inside the sampling kernel (no [C, K, P] buffer is needed for them)."""
    if thin < 1:
        raise ValueError("thin must be >= 1")

    def sampler(key, Nsamples: int, params):
        torch = _native.require_cuda()
        keys, kb = keys_to_device(key)
        flat, treedef, pb = eng.flatten_params(params, broadcast_chains=keys.shape[0])
        if keys.shape[0] != flat.shape[0]:
            raise ValueError("need one key per chain")
        batched = kb or pb
        n_chains, nsamp = flat.shape[0], int(Nsamples)
        nkeep = nsamp // thin
        st = eng.new_state(flat)
        eng.init(st, keys, sampler_mode=True)
        samples = torch.empty((n_chains, nkeep, eng.P), dtype=torch.float32, device="cuda") if store_samples else None
        moments = None
        if summaries:                                                  # planes: sum (theta - r), sum (theta - r)^2, r = theta_0
            moments = torch.zeros((n_chains, 3, eng.P), dtype=torch.float32, device="cuda")
            moments[:, 2] = st["theta"]
        keep = []
        # launches are multiples of `thin` so that the kept set does not depend on the chunking
        chunk = nsamp if not pbar else max(nsamp // 20, 1)
        chunk = max(-(-chunk // thin) * thin, thin)
        bar = None
        if pbar and nsamp > 0:
            from tqdm.auto import tqdm

            bar = tqdm(total=nsamp)
        for i0 in range(0, nkeep * thin, chunk):
            k = min(chunk, nkeep * thin - i0)
            dst = samples[:, i0 // thin:] if store_samples else None
            keep.append(eng.run(st, i0, k, dst, sample_chain_stride=nkeep * eng.P if store_samples else 0, thin=thin,
                                moments=moments))
            if bar is not None:
                torch.cuda.current_stream().synchronize()
                bar.update(k)
        if nsamp > nkeep * thin:                                       # trailing iterations that keep nothing
            keep.append(eng.run(st, nkeep * thin, nsamp - nkeep * thin, None))
        if bar is not None:
            bar.update(nsamp - nkeep * thin)
            bar.close()
        torch.cuda.current_stream().synchronize()
        like = _tree.flatten(params)[0][0]

        def as_tree(flat_block, lead):
            tree = eng.unflatten_params(flat_block, treedef, lead)
            leaves, td = _tree.flatten(tree)
            return _tree.unflatten(td, [_like(l, like) for l in leaves])

        out = None
        if store_samples:
            out = as_tree(samples if batched else samples[0], (n_chains, nkeep) if batched else (nkeep,))
        if not summaries:
            return out
        n = max(nkeep, 1)
        shift = moments[:, 0].double() / n
        mean = moments[:, 2].double() + shift
        var = (moments[:, 1].double() / n - shift * shift).clamp_min(0.0)
        lead = (n_chains,) if batched else ()
        summ = {"mean": as_tree((mean if batched else mean[0]).float(), lead),
                "var": as_tree((var if batched else var[0]).float(), lead), "count": nkeep}
        return out, summ

    sampler.engine = eng
    return sampler


def _build_native_sampler(init_fn: Callable, as_list: bool, pbar: bool) -> Callable:
    eng = init_fn.engine

    def sampler(key, Nsamples: int, params):
        torch = _native.require_cuda()
        keys, kb = keys_to_device(key)
        flat, treedef, pb = eng.flatten_params(params, broadcast_chains=keys.shape[0])
        if keys.shape[0] != flat.shape[0]:
            raise ValueError("need one key per chain")
        batched = kb or pb
        n_chains, nsamp = flat.shape[0], int(Nsamples)
        st = eng.new_state(flat)
        eng.init(st, keys, sampler_mode=True)                     # samplers.py:53-54
        samples = torch.empty((n_chains, nsamp, eng.P), dtype=torch.float32, device="cuda")
        keep = []
        if pbar and nsamp > 0:
            from tqdm.auto import tqdm

            bar = tqdm(total=nsamp)
            bar.set_description(f"Running for {nsamp:,} iterations", refresh=False)
            chunk = max(nsamp // 20, 1)
            for i0 in range(0, nsamp, chunk):
                k = min(chunk, nsamp - i0)
                keep.append(eng.run(st, i0, k, samples[:, i0:], sample_chain_stride=nsamp * eng.P))
                torch.cuda.current_stream().synchronize()
                bar.update(k)
            bar.close()
        elif nsamp > 0 and _host_result(params) and not as_list and samples.numel() * 4 >= _PIPELINE_BYTES:
            # host (numpy) result: a few launches instead of one, the device->host copy of each block of samples and its
            # unpacking into the result array overlapped with the next launch
            host = _run_pipelined(torch, eng, st, samples, nsamp, keep)
            lead = (n_chains, nsamp) if batched else (nsamp,)
            return eng.unflatten_params(host if batched else host[0], treedef, lead)
        elif nsamp > 0:
            keep.append(eng.run(st, 0, nsamp, samples))           # samplers.py:57
        like = _tree.flatten(params)[0][0]
        if as_list:                                               # samplers.py:81-93
            out = []
            host = samples if batched else samples[0]
            for i in range(nsamp):
                tree = eng.unflatten_params(host[..., i, :], treedef, (n_chains,) if batched else ())
                leaves, td = _tree.flatten(tree)
                out.append(_tree.unflatten(td, [_like(l, like) for l in leaves]))
            return out
        lead = (n_chains, nsamp) if batched else (nsamp,)
        tree = eng.unflatten_params(samples if batched else samples[0], treedef, lead)
        leaves, td = _tree.flatten(tree)
        return _tree.unflatten(td, [_like(l, like) for l in leaves])

    sampler.engine = eng
    return sampler


def sgmcmc_sampler(build_kernel_fn: Callable) -> Callable:
    """Decorator turning a kernel factory into a sampler factory (samplers.py:98-122)."""

    @functools.wraps(build_kernel_fn)
    def wrapper(*args, **kwargs):
        compiled = kwargs.pop("compiled", True)
        pbar = kwargs.pop("pbar", True)
        thin = int(kwargs.pop("thin", 1))                    # extensions (keyword-only; defaults = the reference's behaviour)
        summaries = bool(kwargs.pop("summaries", False))
        store_samples = bool(kwargs.pop("store_samples", True))
        init_fn, _kernel, _get_params = build_kernel_fn(*args, **kwargs)
        if thin != 1 or summaries or not store_samples:
            if not compiled:
                raise ValueError("thin / summaries / store_samples are options of the compiled sampler (compiled=True)")
            return _summaries_sampler(init_fn.engine, thin, summaries, store_samples, pbar)
        return _build_native_sampler(init_fn, as_list=not compiled, pbar=pbar)

    return wrapper


build_sgld_sampler = sgmcmc_sampler(build_sgld_kernel)
build_sgldCV_sampler = sgmcmc_sampler(build_sgldCV_kernel)
build_sgld_SVRG_sampler = sgmcmc_sampler(build_sgld_SVRG_kernel)
build_psgld_sampler = sgmcmc_sampler(build_psgld_kernel)
build_sghmc_sampler = sgmcmc_sampler(build_sghmc_kernel)
build_sghmcCV_sampler = sgmcmc_sampler(build_sghmcCV_kernel)
build_sghmc_SVRG_sampler = sgmcmc_sampler(build_sghmc_SVRG_kernel)
build_baoab_sampler = sgmcmc_sampler(build_baoab_kernel)
build_sgnht_sampler = sgmcmc_sampler(build_sgnht_kernel)
build_sgnhtCV_sampler = sgmcmc_sampler(build_sgnhtCV_kernel)
build_badodab_sampler = sgmcmc_sampler(build_badodab_kernel)
build_badodabCV_sampler = sgmcmc_sampler(build_badodabCV_kernel)
# the reference has the kernel but no sampler for sgldAdam; provided here
build_sgldAdam_sampler = sgmcmc_sampler(build_sgldAdam_kernel)
