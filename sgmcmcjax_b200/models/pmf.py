"""Probabilistic matrix factorisation family.  NOT in the reference (only a TODO mention in
docs/nbs/PyMC_integration.ipynb); defined in SURVEY.md section 8 row a15 in the reference's calling
convention:

    params = (U[n_users, d], V[n_items, d])          data = (ij int32[N, 2], r float32[N])
    loglikelihood(params, ij, r) = -0.5 * alpha * (r - U[i] . V[j])^2        alpha = 2.0
    logprior(params)             = -0.5 * lam * (|U|^2 + |V|^2)              lam = 1.0
"""
from __future__ import annotations

from .. import _native
from ._family import FamilySpec, TaggedLogDensity


def make(alpha: float = 2.0, lam: float = 1.0):
    spec = FamilySpec(_native.FAMILY_PMF, "pmf", 2, (float(alpha),), (float(lam),))

    def loglikelihood(params, ij, r):
        u, v = params
        e = r - (u[ij[0]] * v[ij[1]]).sum()
        return -0.5 * alpha * e * e

    def logprior(params):
        u, v = params
        return -0.5 * lam * ((u * u).sum() + (v * v).sum())

    return TaggedLogDensity(spec, "loglikelihood", loglikelihood), TaggedLogDensity(spec, "logprior", logprior)


loglikelihood, logprior = make()
