"""Gaussian-mean family: README.md:22-26 (== tests/models.py:6-11) and its two-leaf variant
tests/models.py:15-24, generalised to per-leaf coefficients

    loglikelihood(theta, x) = -sum_l a_l |x - theta_l|^2
    logprior(theta)         = -sum_l c_l |theta_l|^2

``loglikelihood, logprior`` below are the README model (a = 0.5, c = 0.005)."""
from __future__ import annotations

from .. import _native
from ._family import FamilySpec, TaggedLogDensity


def _leaves(theta):
    return list(theta) if isinstance(theta, (list, tuple)) else [theta]


def make(lik_coef=(0.5,), prior_coef=(0.005,)):
    """Return a registered (loglikelihood, logprior) pair."""
    lik_coef = tuple(float(a) for a in lik_coef)
    prior_coef = tuple(float(c) for c in prior_coef)
    if len(lik_coef) != len(prior_coef) or not 1 <= len(lik_coef) <= _native.MAX_LEAVES:
        raise ValueError("lik_coef and prior_coef need one entry per parameter leaf (1..8 leaves)")
    spec = FamilySpec(_native.FAMILY_GAUSSIAN_MEAN, "gaussian_mean", 1, lik_coef, prior_coef)

    def loglikelihood(theta, x):
        return -sum(a * ((x - t) * (x - t)).sum() for a, t in zip(lik_coef, _leaves(theta)))

    def logprior(theta):
        return -sum(c * (t * t).sum() for c, t in zip(prior_coef, _leaves(theta)))

    return TaggedLogDensity(spec, "loglikelihood", loglikelihood), TaggedLogDensity(spec, "logprior", logprior)


loglikelihood, logprior = make()
# tests/models.py:15-24
loglikelihood_list_array, logprior_list_array = make((0.5, 0.1), (0.001, 0.001))
# tests/models.py:6-11 is the README model
loglikelihood_array, logprior_array = loglikelihood, logprior
