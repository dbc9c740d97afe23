"""jax.random-compatible key helpers backed by the device PRNG (csrc/prng.cuh).

Keys are raw ``uint32[2]`` threefry keys with jax 0.4.13 semantics
(``PRNGKey(seed) == [0, seed]``); ``split`` / ``normal`` / ``choice`` run the same
device code the sampler kernels use, through the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)


def _key_arg(key):
    k = np.ascontiguousarray(np.asarray(key.cpu() if hasattr(key, "cpu") else key).astype(np.uint32).reshape(2))
    return k, k.ctypes.data_as(C.POINTER(C.c_uint32))


def _stream(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def split(key, num: int = 2) -> np.ndarray:
    """random.split(key, num) -> host uint32[num, 2]."""
    torch = _native.require_cuda()
    lib = _native.load()
    out = torch.empty((num, 2), dtype=torch.int32, device="cuda")
    k, kp = _key_arg(key)
    _native.check(lib.sgmcmc_split(kp, num, C.c_void_p(out.data_ptr()), _stream(torch)))
    return out.cpu().numpy().view(np.uint32)


def normal(key, shape=()):
    """random.normal(key, shape) in fp32 -> CUDA torch tensor."""
    torch = _native.require_cuda()
    lib = _native.load()
    n = int(np.prod(shape)) if len(shape) else 1
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    k, kp = _key_arg(key)
    _native.check(lib.sgmcmc_normal(kp, n, C.c_void_p(out.data_ptr()), _stream(torch)))
    return out.reshape(tuple(shape))


def uniform(key, shape=()):
    """random.uniform(key, shape) on [0, 1) in fp32 -> CUDA torch tensor."""
    torch = _native.require_cuda()
    lib = _native.load()
    n = int(np.prod(shape)) if len(shape) else 1
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    k, kp = _key_arg(key)
    _native.check(lib.sgmcmc_uniform(kp, n, C.c_void_p(out.data_ptr()), _stream(torch)))
    return out.reshape(tuple(shape))


def bernoulli_rows(key, p):
    """vmap(random.bernoulli)(random.split(key, N), p) -> CUDA int32 [N] (models/logistic_regression.py:59-61)."""
    torch = _native.require_cuda()
    lib = _native.load()
    p = p.to(device="cuda", dtype=torch.float32).contiguous()
    out = torch.empty(p.shape[0], dtype=torch.int32, device="cuda")
    k, kp = _key_arg(key)
    _native.check(lib.sgmcmc_bernoulli_rows(kp, int(p.shape[0]), C.c_void_p(p.data_ptr()), C.c_void_p(out.data_ptr()),
                                            _stream(torch)))
    return out


def choice_indices(key, n: int, batch: int):
    """random.choice(key, arange(n), (batch,)) -> CUDA int32 tensor (gradient_estimation.py:32)."""
    torch = _native.require_cuda()
    lib = _native.load()
    out = torch.empty(batch, dtype=torch.int32, device="cuda")
    k, kp = _key_arg(key)
    _native.check(lib.sgmcmc_minibatch_indices(kp, int(n), int(batch), C.c_void_p(out.data_ptr()), _stream(torch)))
    return out
