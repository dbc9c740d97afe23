"""numpy restatement of the jax==0.4.13 PRNG used by the reference (TEST INFRASTRUCTURE).

The reference draws all randomness through ``jax.random`` with raw ``uint32[2]``
threefry keys and ``jax_threefry_partitionable=False`` (the 0.4.13 default).
Call sites in the reference:

* ``random.split``   samplers.py:49,53,83,89  kernels.py:55,114,120
                     diffusion_util.py:66,91,112  diffusions.py:272
* ``random.choice``  gradient_estimation.py:32,75,140
* ``random.normal``  diffusions.py:67,104,148,189,198,232,273,283,332

jax itself is NOT under /root/reference (third-party, pinned in
requirements.txt:3-4), so this file restates its published algorithm:
Threefry-2x32 (Salmon et al., SC'11, 20 rounds), jax's "half-split" counter
layout, ``_split``, ``_random_bits``, ``_uniform``, ``_normal_real`` (XLA's fp32
``ErfInv`` = Giles 2010 single-precision polynomial) and ``_randint``.

Pinned by tests/golden/jax_prng_kat.json (external known answers) except
``randint``/``choice``: PARITY UNPINNED (no external vector; see oracle/__init__.py).
"""
from __future__ import annotations

import numpy as np

U32 = np.uint32
_ROT_A = (13, 15, 26, 6)
_ROT_B = (17, 29, 16, 24)
_PARITY = U32(0x1BD11BDA)


def PRNGKey(seed: int) -> np.ndarray:
    """jax 0.4.13 ``PRNGKey`` for a seed that fits in 32 bits: ``[0, seed]``."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32_block(k0, k1, c0, c1):
    """One Threefry-2x32-20 evaluation per element of (c0, c1); returns (w0, w1)."""
    with np.errstate(over="ignore"):
        k0 = U32(k0)
        k1 = U32(k1)
        ks = (k0, k1, U32(k0 ^ k1 ^ _PARITY))
        x0 = (np.asarray(c0, dtype=U32) + ks[0]).astype(U32)
        x1 = (np.asarray(c1, dtype=U32) + ks[1]).astype(U32)
        for g in range(5):
            for r in _ROT_A if g % 2 == 0 else _ROT_B:
                x0 = (x0 + x1).astype(U32)
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = (x0 + ks[(g + 1) % 3]).astype(U32)
            x1 = (x1 + ks[(g + 2) % 3] + U32(g + 1)).astype(U32)
    return x0, x1


def threefry_2x32(key: np.ndarray, counts: np.ndarray) -> np.ndarray:
    """jax's ``threefry_2x32(keypair, count)``: odd lengths are zero-padded, the
    counter array is cut in two halves that feed the two block inputs, the two
    output halves are concatenated (so element j < h is word 0 of block
    (j, j+h) and element j >= h is word 1 of block (j-h, j))."""
    counts = np.asarray(counts, dtype=U32).ravel()
    n = counts.size
    if n % 2:
        counts = np.concatenate([counts, np.zeros(1, dtype=U32)])
    h = counts.size // 2
    w0, w1 = threefry2x32_block(key[0], key[1], counts[:h], counts[h:])
    return np.concatenate([w0, w1])[:n]


def split(key: np.ndarray, num: int = 2) -> np.ndarray:
    """``random.split`` -> uint32[num, 2]."""
    return threefry_2x32(key, np.arange(2 * num, dtype=U32)).reshape(num, 2)


def random_bits(key: np.ndarray, shape) -> np.ndarray:
    """32-bit ``_random_bits``: counters are a C-order iota over ``shape``."""
    size = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
    return threefry_2x32(key, np.arange(size, dtype=U32)).reshape(shape)


def uniform(key, shape=(), minval=0.0, maxval=1.0) -> np.ndarray:
    """fp32 ``random.uniform``: 23 mantissa bits, [minval, maxval)."""
    bits = random_bits(key, shape)
    f = ((bits >> U32(9)) | U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.float32(minval)
    hi = np.float32(maxval)
    return np.maximum(lo, (f * np.float32(hi - lo) + lo).astype(np.float32))


_ERFINV_CENTRAL = np.array(
    [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06, 0.00021858087,
     -0.00125372503, -0.00417768164, 0.246640727, 1.50140941], dtype=np.float32)
_ERFINV_TAIL = np.array(
    [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844, 0.00573950773,
     -0.0076224613, 0.00943887047, 1.00167406, 2.83297682], dtype=np.float32)


def erf_inv_f32(x: np.ndarray) -> np.ndarray:
    """XLA's single-precision ``ErfInv`` (Giles' polynomial), evaluated in fp32."""
    x = np.asarray(x, dtype=np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (-np.log1p((-(x * x)).astype(np.float32))).astype(np.float32)
        central = w < np.float32(5.0)
        wc = np.where(central, w - np.float32(2.5), np.sqrt(w) - np.float32(3.0)).astype(np.float32)
        p = np.where(central, _ERFINV_CENTRAL[0], _ERFINV_TAIL[0]).astype(np.float32)
        for i in range(1, 9):
            c = np.where(central, _ERFINV_CENTRAL[i], _ERFINV_TAIL[i]).astype(np.float32)
            p = (c + p * wc).astype(np.float32)
        out = (p * x).astype(np.float32)
    return np.where(np.abs(x) == 1.0, np.float32(np.inf) * x, out).astype(np.float32)


def normal(key, shape=()) -> np.ndarray:
    """fp32 ``random.normal``: sqrt(2) * erf_inv(uniform(nextafter(-1, 0), 1))."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0), dtype=np.float32)
    u = uniform(key, shape, lo, np.float32(1.0))
    return (np.float32(np.sqrt(2.0)) * erf_inv_f32(u)).astype(np.float32)


def randint_multiplier(span: int) -> int:
    """The uint32 ``multiplier`` of jax's randint: ((2**16 % span)**2 mod 2**32) % span."""
    m = (1 << 16) % span
    return ((m * m) & 0xFFFFFFFF) % span


def randint(key, shape, minval: int, maxval: int) -> np.ndarray:
    """int32 ``random.randint`` (PARITY UNPINNED): two 32-bit draws per element,
    combined modulo span in wrapping uint32 arithmetic."""
    k_hi, k_lo = split(key)
    hi = random_bits(k_hi, shape)
    lo = random_bits(k_lo, shape)
    span = U32(max(int(maxval) - int(minval), 1))
    mult = U32(randint_multiplier(int(span)))
    with np.errstate(over="ignore"):
        off = (((hi % span) * mult).astype(U32) + (lo % span)).astype(U32) % span
    return (np.int32(minval) + off.astype(np.int32)).astype(np.int32)


def choice_indices(key, n: int, batch: int) -> np.ndarray:
    """``random.choice(key, arange(n), (batch,))`` with replacement and p=None
    (gradient_estimation.py:32,75,140) == ``randint(key, (batch,), 0, n)``."""
    return randint(key, (batch,), 0, n)
