"""The 13 ``build_*_kernel`` factories of sgmcmcjax/kernels.py:130-601, same positional
signatures, returning ``(init_fn, kernel, get_params)`` triples backed by the fused CUDA path.

Differences from the reference that a caller can observe:
* ``loglikelihood`` / ``logprior`` must be a registered pair from ``sgmcmcjax_b200.models``.
* Multi-chain: pass ``key`` as uint32[C, 2] and parameter leaves as [C, D]; chain c then equals a
  reference single-chain call with ``key[c]`` (vmap semantics).
* States hold CUDA buffers; ``kernel`` is functional (it returns a new state).
* ``build_badodabCV_kernel`` implements the intended update/update2 palindrome by default;
  ``strict_reference_bugs=True`` reproduces kernels.py:457-460 literally.
"""
from __future__ import annotations

from typing import Callable, Tuple

from . import _native
from ._engine import Engine, keys_to_device
from .types import DiffusionState, SamplerState, SVRGState


def _like(t, like):
    """Return `t` (CUDA torch tensor) in the array type of `like`."""
    import numpy as np

    if isinstance(like, np.ndarray) or np.isscalar(like) or isinstance(like, (list, tuple)):
        if t.numel() * t.element_size() >= (1 << 20):
            # large results (sample blocks): D2H into page-locked memory from torch's caching host allocator -
            # a pageable copy runs at a fraction of the PCIe rate and dominates the end-to-end time
            import torch

            host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            host.copy_(t, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return host.numpy()
        return t.cpu().numpy()
    mod = type(like).__module__
    if mod.startswith("jax"):  # pragma: no cover - jax is not installed in this image
        import jax.dlpack

        return jax.dlpack.from_dlpack(t)
    return t


_row_shard = None      # set by parallel.build_data_sharded_sampler while it calls a build_*_kernel factory


def _build(diffusion, dt, loglikelihood, logprior, data, batch_size, **kw) -> Tuple[Callable, Callable, Callable]:
    if _row_shard is not None:
        kw.update(row_shard=_row_shard[0], group=_row_shard[1])
    eng = Engine(diffusion, dt, loglikelihood, logprior, data, batch_size, **kw)

    def _wrap(st, treedef, batched, like) -> SamplerState:
        lead = (st["theta"].shape[0],) if batched else ()
        un = lambda flat: eng.unflatten_params(flat, treedef, lead)
        svrg = SVRGState()
        if eng.estimator == "svrg":
            svrg = SVRGState(un(st["svrg_center"]), un(st["svrg_grad"]))
        return SamplerState(DiffusionState(st, treedef, (batched, like)), un(st["grad"]), svrg, None)

    def init_fn(key, params) -> SamplerState:                    # kernels.py:48-51
        keys, kb = keys_to_device(key)
        flat, treedef, pb = eng.flatten_params(params, broadcast_chains=keys.shape[0])
        if keys.shape[0] != flat.shape[0]:
            raise ValueError("need one key per chain")
        st = eng.new_state(flat)
        eng.init(st, keys, sampler_mode=False)
        from . import _tree
        like = _tree.flatten(params)[0][0]
        return _wrap(st, treedef, kb or pb, like)

    def kernel(i: int, key, state: SamplerState) -> SamplerState:   # kernels.py:53-64 / :110-122
        st0, treedef, (batched, like) = state.diffusion_state
        keys, _ = keys_to_device(key)
        if keys.shape[0] != st0["theta"].shape[0]:
            raise ValueError("need one key per chain")
        st = eng.clone_state(st0)
        eng.step(st, int(i), keys)
        return _wrap(st, treedef, batched, like)

    def get_params(state: SamplerState):                          # kernels.py:66-67
        st, treedef, (batched, like) = state.diffusion_state
        lead = (st["theta"].shape[0],) if batched else ()
        leaves_tree = eng.unflatten_params(st["theta"], treedef, lead)
        from . import _tree
        leaves, td = _tree.flatten(leaves_tree)
        return _tree.unflatten(td, [_like(l, like) for l in leaves])

    for f in (init_fn, kernel, get_params):
        f.engine = eng
    return init_fn, kernel, get_params


def build_sgld_kernel(dt, loglikelihood, logprior, data, batch_size):
    """kernels.py:130-153."""
    return _build("sgld", dt, loglikelihood, logprior, data, batch_size)


def build_sgldCV_kernel(dt, loglikelihood, logprior, data, batch_size, centering_value):
    """kernels.py:156-185."""
    return _build("sgld", dt, loglikelihood, logprior, data, batch_size, estimator="cv",
                  centering_value=centering_value)


def build_sgld_SVRG_kernel(dt, loglikelihood, logprior, data, batch_size, update_rate):
    """kernels.py:188-217."""
    return _build("sgld", dt, loglikelihood, logprior, data, batch_size, estimator="svrg", update_rate=update_rate)


def build_psgld_kernel(dt, loglikelihood, logprior, data, batch_size, alpha: float = 0.99, eps: float = 1e-5):
    """kernels.py:220-251."""
    return _build("psgld", dt, loglikelihood, logprior, data, batch_size, hyper=dict(alpha=alpha, eps=eps))


def build_sgldAdam_kernel(dt, loglikelihood, logprior, data, batch_size, beta1: float = 0.9,
                          beta2: float = 0.999, eps: float = 1e-8):
    """kernels.py:254-287."""
    return _build("sgldAdam", dt, loglikelihood, logprior, data, batch_size,
                  hyper=dict(beta1=beta1, beta2=beta2, eps=eps))


def build_sgnht_kernel(dt, loglikelihood, logprior, data, batch_size, a: float = 0.01):
    """kernels.py:290-319."""
    return _build("sgnht", dt, loglikelihood, logprior, data, batch_size, hyper=dict(a=a))


def build_sgnhtCV_kernel(dt, loglikelihood, logprior, data, batch_size, centering_value, a: float = 0.01):
    """kernels.py:322-353."""
    return _build("sgnht", dt, loglikelihood, logprior, data, batch_size, estimator="cv",
                  centering_value=centering_value, hyper=dict(a=a))


def build_baoab_kernel(dt, loglikelihood, logprior, data, batch_size, gamma, tau: float = 1.0):
    """kernels.py:356-391."""
    return _build("baoab", dt, loglikelihood, logprior, data, batch_size, hyper=dict(gamma=gamma, tau=tau))


def build_badodab_kernel(dt, loglikelihood, logprior, data, batch_size, a: float = 0.01):
    """kernels.py:394-427."""
    return _build("badodab", dt, loglikelihood, logprior, data, batch_size, hyper=dict(a=a))


def build_badodabCV_kernel(dt, loglikelihood, logprior, data, batch_size, centering_value, a: float = 0.01,
                           strict_reference_bugs: bool = False):
    """kernels.py:430-465 (see the module docstring about kernels.py:457-460)."""
    flags = _native.FLAG_BADODABCV_REFERENCE_BUG if strict_reference_bugs else 0
    return _build("badodab", dt, loglikelihood, logprior, data, batch_size, estimator="cv",
                  centering_value=centering_value, hyper=dict(a=a), flags=flags)


def build_sghmc_kernel(dt, L, loglikelihood, logprior, data, batch_size, alpha: float = 0.01,
                       compiled_leapfrog: bool = True):
    """kernels.py:471-511 (compiled_leapfrog accepted and ignored: the L leapfrogs are always fused)."""
    return _build("sghmc", dt, loglikelihood, logprior, data, batch_size, leapfrog=L, hyper=dict(alpha=alpha))


def build_sghmcCV_kernel(dt, L, loglikelihood, logprior, data, batch_size, centering_value, alpha: float = 0.01,
                         compiled_leapfrog: bool = True):
    """kernels.py:514-556."""
    return _build("sghmc", dt, loglikelihood, logprior, data, batch_size, estimator="cv",
                  centering_value=centering_value, leapfrog=L, hyper=dict(alpha=alpha))


def build_sghmc_SVRG_kernel(dt, L, loglikelihood, logprior, data, batch_size, update_rate, alpha: float = 0.01,
                            compiled_leapfrog: bool = True):
    """kernels.py:559-601."""
    return _build("sghmc", dt, loglikelihood, logprior, data, batch_size, estimator="svrg",
                  update_rate=update_rate, leapfrog=L, hyper=dict(alpha=alpha))
