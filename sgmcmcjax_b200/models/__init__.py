"""Registered model families.  The native path consumes (loglikelihood, logprior) pairs
created here; arbitrary Python log-densities are out of scope (BASELINE.json north_star)."""
from . import bayesian_mlp, gaussian_mean, logistic_regression, pmf  # noqa: F401
from ._family import FamilySpec, resolve_family  # noqa: F401
