"""Step-size schedules (sgmcmcjax/diffusions.py:351-383) and the host-side evaluation of the
per-step scalar coefficients the fused kernels consume (struct sgmcmc_step_coef).

The diffusion updates themselves (diffusions.py:47-347) are device code:
csrc/vec_sampler.cuh ``ChainCtx::diffuse``.
"""
from __future__ import annotations

import math
from typing import Callable, Union

import numpy as np

from . import _native

f32 = np.float32


def constant(step_size: float) -> Callable:
    """diffusions.py:351-355."""
    def schedule(i):
        return step_size

    schedule._sgmcmc_constant = True
    return schedule


def welling_teh_schedule(a: float, b: float, gamma: float = 0.55) -> Callable:
    """Polynomial schedule, diffusions.py:358-364."""
    def schedule(i):
        return a * (b + i) ** (-gamma)

    return schedule


def cyclical_schedule(alpha_0: float, M: int, K: int) -> Callable:
    """diffusions.py:367-374."""
    period = math.ceil(K / M)

    def schedule(i):
        return alpha_0 * 0.5 * (np.cos(math.pi * ((i - 1) % period) / period) + 1)

    return schedule


def make_schedule(scalar_or_schedule: Union[float, Callable]) -> Callable:
    """diffusions.py:377-383."""
    if callable(scalar_or_schedule):
        return scalar_or_schedule
    if np.ndim(scalar_or_schedule) == 0:
        return constant(float(scalar_or_schedule))
    raise TypeError(type(scalar_or_schedule))


DIFFUSION_IDS = {
    "sgld": _native.DIFF_SGLD, "psgld": _native.DIFF_PSGLD, "sgldAdam": _native.DIFF_SGLDADAM,
    "sghmc": _native.DIFF_SGHMC, "baoab": _native.DIFF_BAOAB, "sgnht": _native.DIFF_SGNHT,
    "badodab": _native.DIFF_BADODAB, "adam_opt": _native.DIFF_ADAM_OPT,
}


def coef_is_constant(diffusion: str, schedule: Callable) -> bool:
    return bool(getattr(schedule, "_sgmcmc_constant", False)) and diffusion not in ("sgldAdam", "adam_opt")


def _schedule_values(schedule: Callable, i0: int, k: int) -> np.ndarray:
    """float64 [k]: schedule(i) for i = i0 .. i0+k-1, in one call on the index array when the schedule accepts one."""
    if getattr(schedule, "_sgmcmc_constant", False):
        return np.full(k, float(schedule(0)), dtype=np.float64)
    idx = np.arange(i0, i0 + k)
    try:
        vals = np.asarray(schedule(idx), dtype=np.float64)
        if vals.shape == (k,):
            return vals
    except Exception:
        pass
    return np.array([float(schedule(int(i))) for i in idx], dtype=np.float64)       # scalar-only user schedule


def step_coefs(diffusion: str, hp: dict, schedule: Callable, i0: int, k: int) -> np.ndarray:
    """float32 [k, 4] rows (h, hh, ca, cb) for i = i0 .. i0+k-1, rounded the way jax's weak
    typing rounds the reference's expressions (Python floats combine in double, then fp32).
    Vectorised over the k steps (a launch of 1e6 iterations costs milliseconds of host time, not seconds)."""
    out = np.zeros((k, 4), dtype=f32)
    if k == 0:
        return out
    dt = _schedule_values(schedule, i0, k)
    dt0 = float(schedule(0))
    ca = cb = np.zeros(k, dtype=f32)
    if diffusion == "sgld":                                       # diffusions.py:65-68
        ca = np.sqrt((2 * dt).astype(f32))
    elif diffusion in ("sgldAdam", "adam_opt"):                    # diffusions.py:139-140 / optax bias correction
        t = np.arange(i0 + 1, i0 + k + 1).astype(f32)
        ca = f32(1) - np.power(f32(hp["beta1"]), t)
        cb = f32(1) - np.power(f32(hp["beta2"]), t)
    elif diffusion == "sghmc":                                    # diffusions.py:181-199
        ca = np.sqrt((2 * (hp["alpha"] - hp.get("beta", 0.0)) * dt).astype(f32))
        cb = np.sqrt(dt.astype(f32))
    elif diffusion == "baoab":                                    # diffusions.py:229-232
        ca = np.exp((-hp["gamma"] * dt).astype(f32).astype(np.float64)).astype(f32)   # correctly rounded fp32 exp
        cb = f32(hp["tau"]) * np.sqrt(f32(1) - ca * ca)
    elif diffusion == "sgnht":                                    # diffusions.py:273,283
        ca = np.sqrt((2 * hp["a"] * dt).astype(f32))
        cb = np.full(k, np.sqrt(f32(dt0)), dtype=f32)
    elif diffusion == "badodab":                                  # diffusions.py:329
        ca = np.sqrt(dt.astype(f32))
    out[:, 0] = dt.astype(f32)
    out[:, 1] = (dt * 0.5).astype(f32)
    out[:, 2] = ca
    out[:, 3] = cb
    return out
