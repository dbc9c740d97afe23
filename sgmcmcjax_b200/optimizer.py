"""MAP optimisation on the fused CUDA path: ``build_optax_optimizer`` of sgmcmcjax/optimizer.py:12-62.

The reference takes any ``optax.GradientTransformation``; optax is not part of this package, and a CUDA kernel cannot run
an arbitrary Python transformation, so the optimiser is the one the reference's docs and tests use, Adam
(``optax.adam(learning_rate)``; tests/test_optimizers.py:27-31), described by the small ``adam(...)`` object below.
The loop (``key, subkey = split(key)``; value and gradient of the log-posterior on the minibatch drawn with ``subkey``;
Adam ascent step) runs as ONE kernel launch, one CTA per chain: the vector families keep the Adam moments in shared memory
(csrc/vec_sampler.cuh), the Bayesian MLP and PMF stream them with the parameters (csrc/big_sampler.cuh ``adam_step``).
Its main use here is the centring value of the control-variate samplers: the ``*CV`` samplers of
the big families need a MAP estimate as much as logistic regression does.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

from . import _native, _tree
from ._engine import Engine, keys_to_device
from .kernels import _like


@dataclass(frozen=True)
class adam:
    """Stand-in for ``optax.adam(learning_rate, b1, b2, eps)`` (a float or a schedule ``i -> float``)."""
    learning_rate: object
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8


def build_optax_optimizer(optimizer: adam, loglikelihood: Callable, logprior: Callable, data, batch_size: int,
                          pbar: bool = False) -> Callable:
    """optimizer.py:12-62: returns ``run_optimizer(key, Niters, params) -> (params, logpost_array)``.
    Multi-chain: ``key`` uint32[C, 2] and parameter leaves [C, D] run C independent optimisations."""
    if not isinstance(optimizer, adam):
        if hasattr(optimizer, "update") and hasattr(optimizer, "init"):
            raise NotImplementedError("arbitrary optax transformations are out of scope: pass sgmcmcjax_b200.optimizer.adam(lr)")
        raise TypeError("optimizer must be sgmcmcjax_b200.optimizer.adam(learning_rate)")
    eng = Engine("adam_opt", optimizer.learning_rate, loglikelihood, logprior, data, batch_size,
                 hyper=dict(beta1=optimizer.b1, beta2=optimizer.b2, eps=optimizer.eps))

    def run_optimizer(key, Niters: int, params):
        torch = _native.require_cuda()
        keys, kb = keys_to_device(key)
        flat, treedef, pb = eng.flatten_params(params, broadcast_chains=keys.shape[0])
        if keys.shape[0] != flat.shape[0]:
            raise ValueError("need one key per chain")
        batched, n_iters, n_chains = kb or pb, int(Niters), flat.shape[0]
        st = eng.new_state(flat)
        st["key"].copy_(keys)                                       # optimizer.py:46 splits the caller's key itself
        lp = torch.empty((n_chains, n_iters), dtype=torch.float32, device="cuda")
        keep = eng.optimize(st, 0, n_iters, lp) if n_iters > 0 else None
        torch.cuda.current_stream().synchronize()
        del keep
        like = _tree.flatten(params)[0][0]
        lead = (n_chains,) if batched else ()
        tree = eng.unflatten_params(st["theta"] if batched else st["theta"][0], treedef, lead)
        leaves, td = _tree.flatten(tree)
        return _tree.unflatten(td, [_like(l, like) for l in leaves]), _like(lp if batched else lp[0], like)

    run_optimizer.engine = eng
    return run_optimizer
