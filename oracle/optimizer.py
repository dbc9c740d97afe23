"""MAP optimisation loop (TEST INFRASTRUCTURE): restates ``sgmcmcjax/optimizer.py:41-60`` with ``optax.adam``.

Key tree: ``key, subkey = split(key)`` per iteration (optimizer.py:46); the minibatch is drawn with ``subkey`` itself
(estimate_gradient -> ``random.choice(key, ...)``, gradient_estimation.py:32 - no further split).  The gradient and the
log-posterior VALUE are evaluated at the current parameters (optimizer.py:47, ``with_val=True``), then the Adam ascent
step is applied.  optax.adam (third-party, not under /root/reference; published algorithm, optax 0.1.x
``scale_by_adam`` + ``scale(-lr)`` on the negated gradient):
    mu = b1 mu + (1 - b1) g;  nu = b2 nu + (1 - b2) g^2;  t += 1
    params += lr * (mu / (1 - b1^t)) / (sqrt(nu / (1 - b2^t)) + eps)          (eps_root = 0)
PARITY UNPINNED (neither jax nor optax can be installed here); the reference's own test is statistical
(tests/test_optimizers.py:36: allclose(params, mu_true, atol=0.1)) and is reproduced in tests/.
"""
from __future__ import annotations

import numpy as np

from . import jax_prng as prng


def run_adam(family, data, batch_size, lr, key, n_iters, leaves, b1=0.9, b2=0.999, eps=1e-8, dtype=np.float32):
    """-> (final leaves, log-posterior values [n_iters])."""
    f = dtype
    n = data[0].shape[0]
    leaves = [np.asarray(l, dtype=f).copy() for l in leaves]
    mu = [np.zeros_like(l) for l in leaves]
    nu = [np.zeros_like(l) for l in leaves]
    key = np.asarray(key, dtype=np.uint32)
    lps = np.empty(n_iters, dtype=f)
    for i in range(n_iters):
        key, sub = prng.split(key)
        if batch_size == n:                                   # static full-batch bypass, gradient_estimation.py:41-42
            mb = tuple(data)
        else:
            idx = prng.choice_indices(sub, n, batch_size)
            mb = tuple(a[idx] for a in data)
        lps[i] = family.log_post(leaves, mb, n, np.float64)
        g = family.grad_log_post(leaves, mb, n, f)
        c1 = f(1) - np.power(f(b1), f(i + 1))
        c2 = f(1) - np.power(f(b2), f(i + 1))
        for j in range(len(leaves)):
            mu[j] = (f(b1) * mu[j] + (f(1) - f(b1)) * g[j]).astype(f)
            nu[j] = (f(b2) * nu[j] + (f(1) - f(b2)) * (g[j] * g[j])).astype(f)
            leaves[j] = (leaves[j] + f(lr) * (mu[j] / c1) / (np.sqrt(nu[j] / c2) + f(eps))).astype(f)
    return leaves, lps
