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


def gen_data(key, dim: int, N: int):
    """Synthetic data in the spirit of logistic_regression.py:30-62:
    theta_true = sqrt(10) normal(k1), X = normal(k3, (N, dim)) (identity covariance),
    y ~ Bernoulli(sigmoid(X theta_true)) as int32.  Returns CUDA tensors."""
    import numpy as np
    import torch

    from .. import random

    keys = random.split(key, 4)
    theta_true = random.normal(keys[1], (dim,)) * float(np.float32(np.sqrt(10.0)))
    X = random.normal(keys[3], (N, dim))
    p = torch.sigmoid(X @ theta_true)
    u = (torch.randint(0, 1 << 23, (N,), device=X.device, generator=torch.Generator(X.device).manual_seed(
        int(keys[0][1]))).float() / float(1 << 23))
    y = (u < p).to(torch.int32)
    return theta_true, X, y
