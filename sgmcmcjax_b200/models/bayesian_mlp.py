"""Bayesian MLP family (sgmcmcjax/models/bayesian_NN/NN_model.py:19-58):

    params = [(W_1[n_1, n_0], b_1[n_1]), ..., (W_L, b_L)]     data = (X float32[N, n_0], y float32[N, n_L] one-hot)
    predict: hidden activation softmax (NN_model.py:36-37), output log_softmax (:39-41)
    loglikelihood(params, x, y) = sum(y * predict(params, x))  (:48-50)
    logprior(params)            = sum of N(0, 1) log-densities  (:53-58)

The reference imports MNIST through tensorflow_datasets at import time (NN_data.py:19-21); that loader is
out of scope - pass any (X, y) of this shape.
"""
from __future__ import annotations

import numpy as np

from .. import _native
from ._family import FamilySpec, TaggedLogDensity


def make(prior_precision: float = 1.0):
    spec = FamilySpec(_native.FAMILY_MLP, "bayesian_mlp", 2, (), (float(prior_precision),))

    def predict(params, x):
        import torch

        a = x
        for w, b in params[:-1]:
            a = torch.softmax(w @ a + b, dim=-1)
        w, b = params[-1]
        return torch.log_softmax(w @ a + b, dim=-1)

    def loglikelihood(params, x, y):
        return (y * predict(params, x)).sum()

    def logprior(params):
        lp = 0.0
        for w, b in params:
            lp = lp - 0.5 * prior_precision * ((w * w).sum() + (b * b).sum()) - 0.5 * np.log(2 * np.pi) * (w.numel() + b.numel())
        return lp

    return TaggedLogDensity(spec, "loglikelihood", loglikelihood), TaggedLogDensity(spec, "logprior", logprior)


loglikelihood, logprior = make()


def init_network(key, sizes, scale: float = 1e-2):
    """NN_model.py:19-26 with the device PRNG: keys = split(key, len(sizes)); layer k: (key, subkey) =
    split(k); W = scale * normal(key, (n, m)), b = scale * normal(subkey, (n,)).  Returns CUDA tensors."""
    from .. import random

    keys = random.split(key, len(sizes))
    params = []
    for k, m, n in zip(keys, sizes[:-1], sizes[1:]):
        kk = random.split(k, 2)
        params.append((scale * random.normal(kk[0], (n, m)), scale * random.normal(kk[1], (n,))))
    return params
