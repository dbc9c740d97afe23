"""Logistic regression family (sgmcmcjax/models/logistic_regression.py:65-72):

    loglikelihood(theta, x, y) = -logsumexp([0, (1 - 2y) theta.x])
    logprior(theta)            = -(0.5 / 10) theta.theta
"""
from __future__ import annotations

from .. import _native
from ._family import FamilySpec, TaggedLogDensity


def make(prior_variance: float = 10.0):
    spec = FamilySpec(_native.FAMILY_LOGREG, "logistic_regression", 2, (), (1.0 / float(prior_variance),))

    def loglikelihood(theta, x_val, y_val):
        import torch

        s = (1.0 - 2.0 * y_val) * (theta * x_val).sum()
        return -torch.nn.functional.softplus(torch.as_tensor(s))

    def logprior(theta):
        return -(0.5 / prior_variance) * (theta * theta).sum()

    return TaggedLogDensity(spec, "loglikelihood", loglikelihood), TaggedLogDensity(spec, "logprior", logprior)


loglikelihood, logprior = make()


def _integer_pow_f32(x, n: int):
    """x ** n (n >= 1) by repeated squaring in fp32, as XLA expands lax.integer_pow."""
    import numpy as np

    acc, base = None, np.float32(x)
    while n > 0:
        if n & 1:
            acc = base if acc is None else np.float32(acc * base)
        n >>= 1
        if n:
            base = np.float32(base * base)
    return np.float32(1.0) if acc is None else acc


def genCovMat(key, d: int, rho: float):
    """logistic_regression.py:12-19.  The reference calls ``random.uniform(key)`` with the SAME key for every entry, so
    one scalar u decides the matrix: Sigma[i, j] = (u * 2 rho - rho) ** |i - j|, unit diagonal.  Host numpy [d, d] f32."""
    import numpy as np

    from .. import random

    u = np.float32(random.uniform(key, ()).item())
    base = np.float32(np.float32(np.float32(u * np.float32(2.0)) * np.float32(rho)) - np.float32(rho))
    powers = [np.float32(1.0)] + [_integer_pow_f32(base, k) for k in range(1, d)]
    idx = np.abs(np.arange(d)[:, None] - np.arange(d)[None, :])
    return np.asarray(powers, np.float32)[idx]


def gen_data(key, dim: int, N: int):
    """``gen_data`` of logistic_regression.py:30-62 with the same key tree, on the device:
    ``key, k1, k2, k3 = split(key, 4)``; theta_true = normal(k1, (dim,)) * sqrt(10); covX = genCovMat(k2, dim, 0.4);
    X = normal(k3, (N, dim)) @ cholesky(covX); y_i = bernoulli(split(key, N)[i], sigmoid(theta_true . X_i)) as int32.
    The PRNG draws are bit-identical to jax 0.4.13's; the Cholesky factor and the [N, dim] x [dim, dim] product are
    fp32 library calls here (cuSOLVER / cuBLAS) as they are in the reference (XLA), equal to rounding.  Returns CUDA
    tensors (theta_true [dim], X [N, dim], y int32 [N])."""
    import numpy as np
    import torch

    from .. import random

    keys = random.split(key, 4)
    theta_true = random.normal(keys[1], (dim,)) * float(np.sqrt(np.float32(10.0)))
    cov = torch.from_numpy(genCovMat(keys[2], dim, 0.4)).cuda()
    chol = torch.linalg.cholesky(cov)
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        X = (random.normal(keys[3], (N, dim)) @ chol).contiguous()
        p = 1.0 / (1.0 + torch.exp(-(X @ theta_true)))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    y = random.bernoulli_rows(keys[0], p)
    return theta_true, X, y
