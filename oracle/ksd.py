"""IMQ kernel Stein discrepancy (TEST INFRASTRUCTURE).

Restates ``sgmcmcjax/ksd.py``: ``k_0_fun`` (ksd.py:7-37) for one ordered pair,
and ``imq_KSD`` (ksd.py:43-66) = sqrt(sum over ALL N^2 ordered pairs, i == j
included) / N with c = 1, beta = -1/2 (ksd.py:54).

``imq_ksd_literal`` follows the reference's term-by-term formula row by row
(float64 by default: the "truth" both the reference's fp32 scan and the CUDA
kernel are compared against, SURVEY.md section 7 "KSD fp32 cancellation").
PARITY UNPINNED: the reference has no KSD test and cannot be run here.
"""
from __future__ import annotations

import numpy as np


def k0_pair(x1, x2, g1, g2, c=1.0, beta=-0.5):
    """ksd.py:29-37, one pair, terms kept in the reference's order."""
    diff = x1 - x2
    dim = x1.shape[0]
    base = c ** 2 + diff @ diff
    t1 = (g1 @ g2) * base ** beta
    t2 = -2 * beta * (g1 @ diff) * base ** (beta - 1)
    t3 = 2 * beta * (g2 @ diff) * base ** (beta - 1)
    t4 = -2 * dim * beta * base ** (beta - 1)
    t5 = -4 * beta * (beta - 1) * base ** (beta - 2) * np.sum(np.square(diff))
    return t1 + t2 + t3 + t4 + t5


def k0_rows(xi, X, gi, G, c=1.0, beta=-0.5):
    """ksd.py:40 - k0 of row i against every row j (vectorised over j)."""
    diff = xi[None, :] - X
    dim = X.shape[1]
    sq = np.einsum("nd,nd->n", diff, diff)
    base = c ** 2 + sq
    t1 = (G @ gi) * base ** beta
    t2 = -2 * beta * (diff @ gi) * base ** (beta - 1)
    t3 = 2 * beta * np.einsum("nd,nd->n", G, diff) * base ** (beta - 1)
    t4 = -2 * dim * beta * base ** (beta - 1)
    t5 = -4 * beta * (beta - 1) * base ** (beta - 2) * sq
    return t1 + t2 + t3 + t4 + t5


def imq_ksd_sum(samples, grads, dtype=np.float64, row0=0, row1=None):
    """Sum of k0 over rows [row0, row1) x all columns (the row-block shard unit)."""
    X = np.asarray(samples, dtype=dtype)
    G = np.asarray(grads, dtype=dtype)
    row1 = X.shape[0] if row1 is None else row1
    total = dtype(0.0)
    for i in range(row0, row1):                  # lax.scan over rows, ksd.py:58-65
        total = dtype(total + np.sum(k0_rows(X[i], X, G[i], G), dtype=dtype))
    return total


def imq_ksd_literal(samples, grads, dtype=np.float64):
    """ksd.py:43-66."""
    n = np.shape(samples)[0]
    return dtype(np.sqrt(imq_ksd_sum(samples, grads, dtype)) / dtype(n))
