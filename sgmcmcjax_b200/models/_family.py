"""Family registry: how a (loglikelihood, logprior) pair maps onto a native family id."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Tuple

from .. import _native


@dataclass(frozen=True)
class FamilySpec:
    family: int                      # _native.FAMILY_*
    name: str
    n_data_arrays: int               # length of the `data` tuple the family expects
    lik_coef: Tuple[float, ...] = ()
    prior_coef: Tuple[float, ...] = ()
    extra: tuple = field(default_factory=tuple)


class TaggedLogDensity:
    """A callable carrying the family tag; calling it evaluates the log-density with torch
    (convenience only - the samplers never call it)."""

    def __init__(self, spec: FamilySpec, role: str, fn):
        self._sgmcmc_family = spec
        self._sgmcmc_role = role
        self._fn = fn
        self.__name__ = role

    def __call__(self, *args):
        return self._fn(*args)


_REFERENCE_FUNCS = {
    # the reference's own model functions, recognised by module / name if that package is importable
    ("sgmcmcjax.models.logistic_regression", "loglikelihood"): ("logreg", "loglikelihood"),
    ("sgmcmcjax.models.logistic_regression", "logprior"): ("logreg", "logprior"),
}


def resolve_family(loglikelihood, logprior) -> FamilySpec:
    """Map the two callables of the reference signature onto a registered family."""
    spec = getattr(loglikelihood, "_sgmcmc_family", None)
    spec_p = getattr(logprior, "_sgmcmc_family", None)
    if spec is None or spec_p is None:
        key_l = (getattr(loglikelihood, "__module__", None), getattr(loglikelihood, "__name__", None))
        key_p = (getattr(logprior, "__module__", None), getattr(logprior, "__name__", None))
        if _REFERENCE_FUNCS.get(key_l) == ("logreg", "loglikelihood") and \
                _REFERENCE_FUNCS.get(key_p) == ("logreg", "logprior"):
            from . import logistic_regression
            return logistic_regression.loglikelihood._sgmcmc_family
        raise NotImplementedError(
            "arbitrary log-densities are out of scope: pass a (loglikelihood, logprior) pair from "
            "sgmcmcjax_b200.models (gaussian_mean, logistic_regression); a CUDA kernel cannot consume a "
            "user-supplied Python function and there is no CPU fallback")
    if spec != spec_p:
        raise ValueError("loglikelihood and logprior belong to different registered model instances")
    if getattr(loglikelihood, "_sgmcmc_role", None) != "loglikelihood" or \
            getattr(logprior, "_sgmcmc_role", None) != "logprior":
        raise ValueError("expected (loglikelihood, logprior) in that order")
    return spec
