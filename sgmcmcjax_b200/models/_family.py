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

    # ---- geometry: how parameter leaves and data map onto the native config ----------------
    @property
    def vector(self) -> bool:
        """Vector families keep every leaf at data_dim elements (shared-memory sampler)."""
        return self.family in (_native.FAMILY_GAUSSIAN_MEAN, _native.FAMILY_LOGREG)

    def data_dtypes(self):
        """torch dtype names of the data tuple."""
        if self.family == _native.FAMILY_PMF:
            return ("int32", "float32")
        if self.family == _native.FAMILY_LOGREG:
            return ("float32", "int32")
        return ("float32", "float32")[: self.n_data_arrays]

    def leaf_ndims(self, n_leaves: int):
        """Per-chain rank of every parameter leaf (a leaf with one more axis is multi-chain)."""
        if self.family == _native.FAMILY_PMF:
            return [2, 2]
        if self.family == _native.FAMILY_MLP:
            return [2, 1] * (n_leaves // 2)
        return [1] * n_leaves

    def geometry(self, leaf_shapes, data_shapes):
        """-> dims[8] for the native config; raises TypeError on a parameter tree that does not fit the
        model (the reference raises TypeError from its pytree checks, diffusion_util.py:58-64)."""
        dims = [0] * 8
        if self.family == _native.FAMILY_PMF:
            if len(leaf_shapes) != 2 or leaf_shapes[0][1] != leaf_shapes[1][1]:
                raise TypeError("pmf parameters are (U[n_users, rank], V[n_items, rank])")
            dims[:3] = [leaf_shapes[0][0], leaf_shapes[1][0], leaf_shapes[0][1]]
        elif self.family == _native.FAMILY_MLP:
            if len(leaf_shapes) % 2 or not leaf_shapes:
                raise TypeError("mlp parameters are [(W_1, b_1), ..., (W_L, b_L)]")
            n_layers = len(leaf_shapes) // 2
            sizes = [leaf_shapes[0][1]]
            for l in range(n_layers):
                w, b = leaf_shapes[2 * l], leaf_shapes[2 * l + 1]
                if len(w) != 2 or w[1] != sizes[-1] or tuple(b) != (w[0],):
                    raise TypeError("mlp layer shapes are inconsistent")
                sizes.append(w[0])
            if sizes[0] != data_shapes[0][1] or sizes[-1] != data_shapes[1][1]:
                raise TypeError("mlp layer sizes do not match the data arrays")
            dims[0] = n_layers
            dims[1:2 + n_layers] = sizes
        return dims


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
            "sgmcmcjax_b200.models (gaussian_mean, logistic_regression, bayesian_mlp, pmf); a CUDA kernel cannot consume a "
            "user-supplied Python function and there is no CPU fallback")
    if spec != spec_p:
        raise ValueError("loglikelihood and logprior belong to different registered model instances")
    if getattr(loglikelihood, "_sgmcmc_role", None) != "loglikelihood" or \
            getattr(logprior, "_sgmcmc_role", None) != "logprior":
        raise ValueError("expected (loglikelihood, logprior) in that order")
    return spec
