"""Gradient estimators, transition kernels and the sampling loop (TEST INFRASTRUCTURE).

Restates, in plain numpy and with explicit key trees:

* ``sgmcmcjax/gradient_estimation.py:12-46``  standard estimator
* ``sgmcmcjax/gradient_estimation.py:50-94``  control variates
* ``sgmcmcjax/gradient_estimation.py:97-153`` SVRG
* ``sgmcmcjax/diffusion_util.py:48-96``       pytree lift: leaf j uses split(key, n_leaves)[j]
* ``sgmcmcjax/kernels.py:17-69``              langevin kernel (carried gradient, palindromes)
* ``sgmcmcjax/kernels.py:72-124``             SGHMC kernel (resample + L leapfrogs)
* ``sgmcmcjax/samplers.py:45-58``             compiled sampler loop

Parameters are lists of numpy leaves.  PARITY UNPINNED for floating point (see
oracle/__init__.py): the reference cannot be executed in this image.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import numpy as np

from . import jax_prng as prng
from .diffusions import DIFFUSIONS


class SamplerState(NamedTuple):
    """types.py:30-34 (diffusion_state is a list of per-leaf dicts here)."""
    diffusion_state: list
    param_grad: list
    svrg_state: Optional[tuple] = None   # (centering_value leaves, fb_grad_center leaves)


class Estimator:
    """kind in {"standard", "cv", "svrg"}; data is a tuple of 1 or 2 arrays."""

    def __init__(self, family, data, batch_size, kind="standard", centering_value=None,
                 update_rate=None, dtype=np.float32):
        assert type(data) == tuple and len(data) in (1, 2)
        self.family, self.data, self.kind, self.f = family, data, kind, dtype
        self.n = data[0].shape[0]
        self.batch = batch_size
        self.update_rate = update_rate
        if kind == "cv":
            self.center = [np.asarray(c, dtype=dtype) for c in centering_value]
            self.fb_center = self.full_grad(self.center)        # gradient_estimation.py:70

    def full_grad(self, leaves):
        return self.family.grad_log_post(leaves, self.data, self.n, self.f)

    def minibatch(self, key):
        idx = prng.choice_indices(key, self.n, self.batch)     # :32,:75,:140
        return tuple(a[idx] for a in self.data)                 # :33,:76,:141

    def _cv(self, key, leaves, center, fb_center):
        mb = self.minibatch(key)
        g = self.family.grad_log_post(leaves, mb, self.n, self.f)
        gc = self.family.grad_log_post(center, mb, self.n, self.f)
        return [((c + a).astype(self.f) - b).astype(self.f) for c, a, b in zip(fb_center, g, gc)]

    def init_gradient(self, key, leaves):
        if self.kind == "standard":                             # :31-35 (always subsamples)
            return self.family.grad_log_post(leaves, self.minibatch(key), self.n, self.f), None
        if self.kind == "cv":                                   # :74-86
            return self._cv(key, leaves, self.center, self.fb_center), None
        fb = self.full_grad(leaves)                             # :124-128
        return fb, ([l.copy() for l in leaves], fb)

    def estimate_gradient(self, i, key, leaves, svrg_state):
        if self.kind == "standard":
            if self.batch is None or self.batch == self.n:      # static full-batch bypass :41-42
                return self.full_grad(leaves), svrg_state
            return self.init_gradient(key, leaves)[0], svrg_state
        if self.kind == "cv":
            return self._cv(key, leaves, self.center, self.fb_center), svrg_state
        if i % self.update_rate == 0:                           # lax.cond :134-139
            svrg_state = ([l.copy() for l in leaves], self.full_grad(leaves))
        return self._cv(key, leaves, svrg_state[0], svrg_state[1]), svrg_state


class Kernel:
    """(init_fn, kernel, get_params) for one chain; ``leapfrog`` = L for the SGHMC family."""

    def __init__(self, diffusion, estimator, leapfrog=None, update2_is_update1=False):
        self.d, self.e, self.L = diffusion, estimator, leapfrog
        # kernels.py:457-460: build_badodabCV_kernel binds BOTH palindrome halves to
        # badodab's update2 (``(update_diff1, update_diff1) = badodab(...)[1]``).
        self.ref_bug_both_update2 = update2_is_update1

    # diffusion_util.py:55-76 / :100-125 - one split(key, n_leaves) per tree update
    def _tree(self, fn, i, key, grads, states):
        keys = prng.split(key, len(states))
        return [fn(i, keys[j], grads[j], states[j]) for j in range(len(states))]

    def get_params(self, state):
        return [s["x"] for s in state.diffusion_state]

    def init_fn(self, key, leaves):                             # kernels.py:48-51
        leaves = [np.asarray(l, dtype=self.d.f) for l in leaves]
        ds = [self.d.init(l) for l in leaves]
        g, svrg = self.e.init_gradient(key, leaves)
        return SamplerState(ds, g, svrg)

    def _langevin(self, i, key, state):                         # kernels.py:53-64
        k1, k2 = prng.split(key)
        first = self.d.update2 if self.ref_bug_both_update2 else self.d.update
        ds = self._tree(first, i, k1, state.param_grad, state.diffusion_state)
        g, svrg = self.e.estimate_gradient(i, k2, [s["x"] for s in ds], state.svrg_state)
        if self.d.palindrome:
            ds = self._tree(self.d.update2, i, k1, g, ds)
        return SamplerState(ds, g, svrg)

    def kernel(self, i, key, state):
        if self.L is None:
            return self._langevin(i, key, state)
        k1, k2 = prng.split(key)                                # kernels.py:114
        keys = prng.split(k1, len(state.diffusion_state))       # diffusion_util.py:88-96
        ds = [self.d.resample_momentum(i, keys[j], s) for j, s in enumerate(state.diffusion_state)]
        state = SamplerState(ds, state.param_grad, state.svrg_state)
        for kl in prng.split(k2, self.L):                       # kernels.py:120-121, same i for all L
            state = self._langevin(i, kl, state)
        return state


def run_sampler(kernel: Kernel, key, n_samples: int, leaves):
    """samplers.py:45-58 - returns a list (per leaf) of arrays [n_samples, *leaf.shape]."""
    key = np.asarray(key, dtype=np.uint32)
    key, sub = prng.split(key)
    state = kernel.init_fn(sub, leaves)
    out = [np.empty((n_samples,) + np.shape(l), dtype=kernel.d.f) for l in leaves]
    for i in range(n_samples):
        key, sub = prng.split(key)
        state = kernel.kernel(i, sub, state)
        for dst, p in zip(out, kernel.get_params(state)):
            dst[i] = p
    return out, state, key


def build_kernel(name: str, family, data, batch_size, dt, dtype=np.float32, **kw) -> Kernel:
    """Factory mirroring the 13 ``build_*_kernel`` wirings (kernels.py:130-601).

    name in: sgld sgldCV sgld_SVRG psgld sgldAdam sgnht sgnhtCV baoab badodab
             badodabCV sghmc sghmcCV sghmc_SVRG.  ``strict_reference_bugs`` selects the
    literal badodabCV wiring of kernels.py:457-460."""
    kind = "cv" if name.endswith("CV") else "svrg" if name.endswith("_SVRG") else "standard"
    base = name[:-2] if kind == "cv" else name[:-5] if kind == "svrg" else name
    strict = kw.pop("strict_reference_bugs", False)
    L = kw.pop("L", None)
    est = Estimator(family, data, batch_size, kind, kw.pop("centering_value", None),
                    kw.pop("update_rate", None), dtype)
    diff = DIFFUSIONS[base](dt, dtype=dtype, **kw)
    return Kernel(diff, est, leapfrog=L if base == "sghmc" else None,
                  update2_is_update1=(strict and name == "badodabCV"))
