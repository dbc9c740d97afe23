"""Synthetic logistic-regression data (TEST INFRASTRUCTURE): restates ``gen_data`` / ``genCovMat`` of
``sgmcmcjax/models/logistic_regression.py:12-62`` in numpy with the jax 0.4.13 PRNG of oracle/jax_prng.py.

Key tree (logistic_regression.py:52): ``key, subkey1, subkey2, subkey3 = random.split(key, 4)``;
  theta_true = normal(subkey1, (dim,)) * sqrt(10)                                   (:55)
  covX       = genCovMat(subkey2, dim, rho = 0.4)                                   (:56)
  X          = normal(subkey3, (N, dim)) @ cholesky(covX)                           (:57)
  y_i        = bernoulli(split(key, N)[i], sigmoid(theta_true . X_i)) as int32      (:59-61)
``genCovMat`` (:12-19) calls ``random.uniform(key)`` with the SAME key for every entry, so one scalar u decides the
whole matrix: Sigma[i, j] = (u * 2 rho - rho) ** |i - j| (fp32 scalar, integer power by repeated squaring as XLA's
``integer_pow`` does), unit diagonal.  ``random.bernoulli(key, p)`` with a scalar p is ``uniform(key, ()) < p``; a
shape-() draw is word 0 of the threefry block of counters (0, pad 0).
PARITY UNPINNED (no jax here): Cholesky and the [N, dim] x [dim, dim] product are fp32 library calls on both sides and
agree to rounding only; a y_i can flip where |u_i - p_i| is within that rounding.
"""
from __future__ import annotations

import numpy as np

from . import jax_prng as prng


def integer_pow_f32(x: np.float32, n: int) -> np.float32:
    """x ** n for n >= 1 by repeated squaring in fp32 (XLA's expansion of lax.integer_pow)."""
    acc, base = None, np.float32(x)
    while n > 0:
        if n & 1:
            acc = base if acc is None else np.float32(acc * base)
        n >>= 1
        if n:
            base = np.float32(base * base)
    return np.float32(1.0) if acc is None else acc


def gen_cov_mat(key, d: int, rho: float) -> np.ndarray:
    u = prng.uniform(key, ())
    base = np.float32(np.float32(np.float32(u * np.float32(2.0)) * np.float32(rho)) - np.float32(rho))
    sigma = np.eye(d, dtype=np.float64)
    for i in range(1, d):
        for j in range(i):
            sigma[i, j] = sigma[j, i] = integer_pow_f32(base, i - j)
    return sigma.astype(np.float32)


def bernoulli_rows(key, p: np.ndarray) -> np.ndarray:
    """vmap(random.bernoulli)(split(key, N), p): row i draws uniform(split(key, N)[i], ()) < p[i]."""
    n = p.shape[0]
    keys = prng.split(key, n)
    zeros = np.zeros(n, dtype=np.uint32)
    w0, _ = prng.threefry2x32_block(keys[:, 0], keys[:, 1], zeros, zeros)
    u = ((w0 >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
    return (u < p.astype(np.float32)).astype(np.int32)


def gen_data(key, dim: int, N: int):
    """-> (theta_true f32[dim], X f32[N, dim], y int32[N])."""
    ks = prng.split(np.asarray(key, np.uint32), 4)
    key2, sub1, sub2, sub3 = ks[0], ks[1], ks[2], ks[3]
    theta_true = (prng.normal(sub1, (dim,)) * np.float32(np.sqrt(np.float32(10.0)))).astype(np.float32)
    cov = gen_cov_mat(sub2, dim, 0.4)
    chol = np.linalg.cholesky(cov.astype(np.float64)).astype(np.float32)
    X = (prng.normal(sub3, (N, dim)) @ chol).astype(np.float32)
    z = (X @ theta_true).astype(np.float32)
    p = (np.float32(1.0) / (np.float32(1.0) + np.exp(-z, dtype=np.float32))).astype(np.float32)
    return theta_true, X, bernoulli_rows(key2, p)
