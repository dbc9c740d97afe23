"""Multi-chain convergence report from the on-device posterior summaries.

``build_*_sampler(..., summaries=True)`` accumulates, inside the sampling kernel, the per-chain mean and marginal
variance of the kept states (``sgmcmc_run_thinned``); with C chains started from different keys these [C, P] blocks
are all the Gelman-Rubin statistic needs, so a convergence check of thousands of chains never stores a sample.
The reference has no multi-chain API and no diagnostics; this is host-side arithmetic on tiny arrays.
"""
from __future__ import annotations

import numpy as np


def _leaves(tree):
    if isinstance(tree, (list, tuple)):
        out = []
        for t in tree:
            out += _leaves(t)
        return out
    return [tree.detach().cpu().numpy() if hasattr(tree, "detach") else np.asarray(tree)]


def rhat_from_moments(mean, var, count: int):
    """Gelman-Rubin potential scale reduction factor per parameter.

    mean, var: [C, ...] per-chain mean and POPULATION variance (divide by n) of n = `count` draws each, C >= 2.
    W = average within-chain (unbiased) variance, B / n = variance of the chain means,
    R-hat = sqrt(((n - 1) / n * W + B / n) / W)."""
    mean, var = np.asarray(mean, np.float64), np.asarray(var, np.float64)
    n, c = int(count), mean.shape[0]
    if c < 2 or n < 2:
        raise ValueError("R-hat needs at least 2 chains and 2 draws per chain")
    w = (var * n / (n - 1)).mean(axis=0)
    b_over_n = mean.var(axis=0, ddof=1)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.sqrt(((n - 1) / n * w + b_over_n) / w)
    return np.where(w > 0, r, np.nan)


def rhat(summaries):
    """R-hat for every leaf of the ``summaries`` dict a multi-chain sampler returned: same pytree as the parameters,
    without the chain axis."""
    means, variances = summaries["mean"], summaries["var"]
    out = [rhat_from_moments(m, v, summaries["count"]) for m, v in zip(_leaves(means), _leaves(variances))]
    if isinstance(means, (list, tuple)):
        return _rebuild(means, iter(out))
    return out[0]


def _rebuild(like, it):
    if isinstance(like, (list, tuple)):
        return type(like)(_rebuild(t, it) for t in like)
    return next(it)


def rhat_from_samples(samples):
    """Reference implementation on stored draws [C, K, ...] (used by the tests to check the moment route)."""
    s = np.asarray(samples, np.float64)
    return rhat_from_moments(s.mean(axis=1), s.var(axis=1), s.shape[1])


def ess_from_samples(samples):
    """Effective sample size per parameter from stored draws [C, K, ...] (all chains pooled): K C / tau with the
    integrated autocorrelation time tau estimated by Geyer's initial positive sequence on the chain-averaged
    autocovariances (FFT).  Needs the draws themselves - use ``thin=`` to keep the buffer small."""
    x = np.asarray(samples, np.float64)
    c, k = x.shape[0], x.shape[1]
    if k < 4:
        raise ValueError("ESS needs at least 4 draws per chain")
    flat = x.reshape(c, k, -1)
    xc = flat - flat.mean(axis=1, keepdims=True)
    nfft = 1 << (2 * k - 1).bit_length()
    f = np.fft.rfft(xc, n=nfft, axis=1)
    acov = np.fft.irfft(f * np.conj(f), n=nfft, axis=1)[:, :k] / k          # biased autocovariance per chain
    acov = acov.mean(axis=0)                                                 # [K, P]
    w = acov[0]
    b_over_n = flat.mean(axis=1).var(axis=0, ddof=1) if c > 1 else 0.0
    var_plus = w * (k - 1) / k + b_over_n                                    # as in R-hat
    with np.errstate(divide="ignore", invalid="ignore"):
        rho = 1.0 - (w - acov) / var_plus                                    # [K, P], rho[0] = 1 - (w - w)/var+ ~ 1
    out = np.empty(flat.shape[2])
    for p in range(flat.shape[2]):
        r = rho[:, p]
        tau, t = -1.0, 0
        while t + 1 < k:                                                     # pairs rho[2m] + rho[2m+1] while positive
            pair = r[t] + r[t + 1]
            if not np.isfinite(pair) or pair < 0.0:
                break
            tau += 2.0 * pair
            t += 2
        out[p] = c * k / max(tau, 1.0 / np.log10(max(c * k, 10)))
    return out.reshape(x.shape[2:])
