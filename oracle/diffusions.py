"""The seven per-leaf diffusion updates and the step-size schedules (TEST INFRASTRUCTURE).

Restates ``sgmcmcjax/diffusions.py`` (one array leaf at a time; the pytree lift
and its per-leaf key rule live in oracle/sgmcmc.py).  Everything is evaluated
in ``dtype`` (fp32 by default).  jax weak typing: Python-float hyper-parameters
are combined in double first and rounded to fp32 when they meet an array
(SURVEY.md A.5), which is what ``f(...)`` below does.

A leaf state is a dict with keys among ``x, v, m, alpha``.
"""
from __future__ import annotations

import math

import numpy as np

from . import jax_prng as prng


def exp_f32(x):
    """fp32 ``exp`` of an fp32 argument, correctly rounded (evaluated in double, rounded once).
    XLA's, numpy's and CUDA's single-precision ``exp`` each differ from this by at most 1-2 ulp and
    from one another; BAOAB / BADODAB form ``1 - exp(-a*h)**2`` with a*h ~ 1e-5..1e-2, which amplifies
    that last-bit disagreement by ~1/(a*h), so the oracle fixes the correctly rounded value."""
    return np.float32(np.exp(np.float64(np.float32(x))))


def make_schedule(dt):
    """diffusions.py:377-383 - a float becomes a constant schedule."""
    return dt if callable(dt) else (lambda i: dt)


def welling_teh_schedule(a, b, gamma=0.55):
    """diffusions.py:358-364."""
    return lambda i: a * (b + i) ** (-gamma)


def cyclical_schedule(alpha_0, M, K):
    """diffusions.py:367-374."""
    period = math.ceil(K / M)
    return lambda i: alpha_0 * 0.5 * (math.cos(math.pi * ((i - 1) % period) / period) + 1)


class Diffusion:
    palindrome = False
    has_resample = False

    def __init__(self, dt, dtype=np.float32):
        self.dt = make_schedule(dt)
        self.f = dtype

    def normal(self, key, shape):
        return prng.normal(key, shape).astype(self.f)


class SGLD(Diffusion):
    """diffusions.py:47-73."""
    fields = ("x",)

    def init(self, x):
        return {"x": x.astype(self.f)}

    def update(self, i, key, g, s):
        f, h = self.f, self.dt(i)
        x = s["x"] + f(h) * g + np.sqrt(f(2 * h)) * self.normal(key, g.shape)
        return {"x": x.astype(f)}


class PSGLD(Diffusion):
    """diffusions.py:76-112 (RMSprop-preconditioned SGLD)."""
    fields = ("x", "v")

    def __init__(self, dt, alpha=0.99, eps=1e-5, dtype=np.float32):
        super().__init__(dt, dtype)
        self.alpha, self.eps = alpha, eps

    def init(self, x):
        return {"x": x.astype(self.f), "v": np.zeros_like(x, dtype=self.f)}

    def update(self, i, key, g, s):
        f, h = self.f, self.dt(i)
        v = f(self.alpha) * s["v"] + f(1 - self.alpha) * np.square(g)
        G = f(1.0) / (np.sqrt(v) + f(self.eps))
        x = s["x"] + f(h * 0.5) * G * g + np.sqrt(f(h) * G) * self.normal(key, g.shape)
        return {"x": x.astype(f), "v": v.astype(f)}


class SGLDAdam(Diffusion):
    """diffusions.py:115-157."""
    fields = ("x", "m", "v")

    def __init__(self, dt, beta1=0.9, beta2=0.999, eps=1e-8, dtype=np.float32):
        super().__init__(dt, dtype)
        self.beta1, self.beta2, self.eps = beta1, beta2, eps

    def init(self, x):
        z = np.zeros_like(x, dtype=self.f)
        return {"x": x.astype(self.f), "m": z, "v": z.copy()}

    def update(self, i, key, g, s):
        f, h = self.f, self.dt(i)
        m = f(self.beta1) * s["m"] + f(1 - self.beta1) * g
        v = f(self.beta2) * s["v"] + f(1 - self.beta2) * np.square(g)
        m_hat = m / (f(1) - f(self.beta1) ** f(i + 1))
        v_hat = v / (f(1) - f(self.beta2) ** f(i + 1))
        a = f(h) / (np.sqrt(v_hat) + f(self.eps))
        x = s["x"] + a * f(0.5) * m_hat + np.sqrt(a) * self.normal(key, g.shape)
        return {"x": x.astype(f), "m": m.astype(f), "v": v.astype(f)}


class SGHMC(Diffusion):
    """diffusions.py:160-201; x moves with the OLD momentum, then v is refreshed."""
    fields = ("x", "v")
    has_resample = True

    def __init__(self, dt, alpha=0.01, beta=0.0, dtype=np.float32):
        super().__init__(dt, dtype)
        self.alpha, self.beta = alpha, beta

    def init(self, x):
        return {"x": x.astype(self.f), "v": np.zeros_like(x, dtype=self.f)}

    def update(self, i, key, g, s):
        f, h = self.f, self.dt(i)
        x = s["x"] + s["v"]
        v = (s["v"] + f(h) * g - f(self.alpha) * s["v"]
             + np.sqrt(f(2 * (self.alpha - self.beta) * h)) * self.normal(key, g.shape))
        return {"x": x.astype(f), "v": v.astype(f)}

    def resample_momentum(self, i, key, s):
        f, h = self.f, self.dt(i)
        return {"x": s["x"], "v": (np.sqrt(f(h)) * self.normal(key, s["x"].shape)).astype(f)}


class BAOAB(Diffusion):
    """diffusions.py:204-247 (update1 before the new gradient, update2 after)."""
    fields = ("x", "v")
    palindrome = True

    def __init__(self, dt, gamma, tau=1.0, dtype=np.float32):
        super().__init__(dt, dtype)
        self.gamma, self.tau = gamma, tau

    def init(self, x):
        return {"x": x.astype(self.f), "v": np.zeros_like(x, dtype=self.f)}

    def update(self, i, key, g, s):
        f, h = self.f, self.dt(i)
        v = s["v"] + f(h * 0.5) * g
        x = s["x"] + v * f(h) * f(0.5)
        c1 = exp_f32(f(-self.gamma * h))
        c2 = np.sqrt(f(1) - c1 * c1)
        v = c1 * v + f(self.tau) * c2 * self.normal(key, g.shape)
        x = x + v * f(h) * f(0.5)
        return {"x": x.astype(f), "v": v.astype(f)}

    def update2(self, i, key, g, s):
        f, h = self.f, self.dt(i)
        return {"x": s["x"], "v": (s["v"] + f(h * 0.5) * g).astype(f)}


class SGNHT(Diffusion):
    """diffusions.py:250-293; thermostat alpha is ONE scalar per leaf."""
    fields = ("x", "v", "alpha")

    def __init__(self, dt, a=0.01, dtype=np.float32):
        super().__init__(dt, dtype)
        self.a = a

    def init(self, x):
        return {"x": x.astype(self.f), "v": np.zeros_like(x, dtype=self.f), "alpha": self.f(self.a)}

    def update(self, i, key, g, s):
        f, h = self.f, self.dt(i)
        v, alpha = s["v"], s["alpha"]
        if i == 0:  # lax.cond(i == 0, initial_momentum, ...) diffusions.py:270-278
            key, sub = prng.split(key)
            v = np.sqrt(f(self.dt(0))) * self.normal(sub, v.shape)
        v = v - alpha * v + f(h) * g + np.sqrt(f(2 * self.a * h)) * self.normal(key, g.shape)
        v = v.astype(f)
        x = s["x"] + v
        nrm = np.sqrt(np.sum(np.square(v), dtype=f))
        alpha = alpha + (nrm * nrm) / f(v.size) - f(h)
        return {"x": x.astype(f), "v": v, "alpha": f(alpha)}


class BADODAB(Diffusion):
    """diffusions.py:296-347; note |v| (not |v|^2) minus v.size, as the code has it."""
    fields = ("x", "v", "alpha")
    palindrome = True

    def __init__(self, dt, a=0.01, dtype=np.float32):
        super().__init__(dt, dtype)
        self.a = a

    def init(self, x):
        return {"x": x.astype(self.f), "v": np.zeros_like(x, dtype=self.f), "alpha": self.f(self.a)}

    def update(self, i, key, g, s):
        f, h = self.f, self.dt(i)
        dt2 = f(h / 2)
        v = (s["v"] + dt2 * g).astype(f)
        x = s["x"] + dt2 * v
        nrm = lambda a: np.sqrt(np.sum(np.square(a), dtype=f))
        alpha = f(s["alpha"] + dt2 * (nrm(v) - f(v.size)))
        c1 = exp_f32(-alpha * f(h))
        with np.errstate(divide="ignore", invalid="ignore"):
            c2 = np.sqrt(f(h)) if alpha == 0 else np.sqrt(np.abs((f(1) - c1 * c1) / (f(2) * alpha)))
        v = (c1 * v + c2 * self.normal(key, g.shape)).astype(f)
        alpha = f(alpha + dt2 * (nrm(v) - f(v.size)))
        x = x + dt2 * v
        return {"x": x.astype(f), "v": v, "alpha": alpha}

    def update2(self, i, key, g, s):
        f, h = self.f, self.dt(i)
        return {"x": s["x"], "v": (s["v"] + f(h * 0.5) * g).astype(f), "alpha": s["alpha"]}


DIFFUSIONS = {
    "sgld": SGLD, "psgld": PSGLD, "sgldAdam": SGLDAdam, "sghmc": SGHMC,
    "baoab": BAOAB, "sgnht": SGNHT, "badodab": BADODAB,
}
