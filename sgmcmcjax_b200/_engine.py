"""Host-side engine: owns a native model handle and drives the fused kernels.

This is the single place where the Python API (samplers.py / kernels.py / gradient_estimation.py)
meets the C ABI (include/sgmcmc_b200.h).  PyTorch is used only for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Sequence

import numpy as np

from . import _native, _tree
from .diffusions import DIFFUSION_IDS, coef_is_constant, make_schedule, step_coefs
from .models import resolve_family

_EST = {"standard": _native.EST_STANDARD, "cv": _native.EST_CV, "svrg": _native.EST_SVRG}


def _vp(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def to_device(a, dtype):
    """numpy / torch / any __dlpack__ producer -> contiguous CUDA torch tensor of `dtype`."""
    torch = _native.require_cuda()
    if isinstance(a, torch.Tensor):
        t = a
    elif isinstance(a, np.ndarray) or np.isscalar(a) or isinstance(a, (list, tuple)):
        arr = np.asarray(a)
        if arr.dtype == np.uint32:
            arr = arr.view(np.int32)
        t = torch.from_numpy(np.ascontiguousarray(arr))
    elif hasattr(a, "__dlpack__"):
        t = torch.from_dlpack(a)
    else:
        t = torch.as_tensor(np.asarray(a))
    return t.to(device="cuda", dtype=dtype, non_blocking=True).contiguous()


def keys_to_device(key):
    """uint32[2] or uint32[C,2] (numpy / torch int32 view) -> CUDA int32 [C,2], and whether it was batched."""
    torch = _native.require_cuda()
    if isinstance(key, torch.Tensor):
        t = key.to(device="cuda").contiguous()
        if t.dtype != torch.int32:
            t = t.to(torch.int64).to(torch.int32) if t.dtype not in (torch.uint32,) else t.view(torch.int32)
    else:
        arr = np.asarray(key)
        arr = np.ascontiguousarray(arr.astype(np.uint32)).view(np.int32)
        t = torch.from_numpy(arr).to("cuda")
    batched = t.dim() == 2
    return t.reshape(-1, 2).contiguous(), batched


class Engine:
    """One model (family + data + estimator + diffusion) bound to a native handle.

    Vector families (gaussian_mean, logreg) know their parameter geometry from the data, so the
    handle exists after construction.  For the big families (bayesian_mlp, pmf) the leaf shapes are
    only known once a parameter pytree is seen (``centering_value`` at build time, else the first
    ``init_fn`` / ``sampler`` call): the handle is created then and is fixed afterwards."""

    def __init__(self, diffusion: str, dt, loglikelihood: Callable, logprior: Callable, data, batch_size: int,
                 estimator: str = "standard", centering_value=None, update_rate: Optional[int] = None,
                 leapfrog: int = 0, hyper: Optional[dict] = None, flags: int = 0, row_shard=None, group=None):
        """row_shard = (row0, n_total): `data` holds only rows [row0, row0 + len) of a table of n_total rows that is
        split over the ranks of `group` (optional data-sharded mode, include/sgmcmc_b200.h)."""
        torch = _native.require_cuda()
        self.lib = _native.load()
        if type(data) != tuple:                                   # gradient_estimation.py:25
            raise AssertionError("data must be a tuple")
        if len(data) not in (1, 2):                               # util.py:34-39
            raise ValueError("'data' must be a tuple of size 1 or 2")
        self.spec = resolve_family(loglikelihood, logprior)
        if len(data) != self.spec.n_data_arrays:
            raise ValueError(f"family {self.spec.name} expects a data tuple of {self.spec.n_data_arrays} arrays")
        self.diffusion, self.estimator = diffusion, estimator
        self.schedule = make_schedule(dt)
        self.hp = dict(hyper or {})
        self.leapfrog = int(leapfrog)
        self.update_rate, self.flags = int(update_rate or 0), int(flags)
        dts = [getattr(torch, n) for n in self.spec.data_dtypes()]
        self._data = [to_device(a, dt_) for a, dt_ in zip(data, dts)]
        x = self._data[0]
        if x.dim() != 2:
            raise ValueError("data[0] must be [N, D]")
        self.n_local, self.data_dim = int(x.shape[0]), int(x.shape[1])
        if len(self._data) == 2 and self._data[1].shape[0] != self.n_local:
            raise ValueError("data arrays must share their leading dimension")
        self.row_shard, self.group = row_shard, group
        self.n_data = int(row_shard[1]) if row_shard is not None else self.n_local
        if row_shard is not None and not self.spec.vector:
            raise NotImplementedError("the data-sharded mode covers the vector families (gaussian_mean, logistic_regression)")
        self.batch = self.n_data if batch_size is None else int(batch_size)
        self._handle = C.c_void_p(0)
        self._coef_cache = {}
        self.leaf_shapes = None
        self._centering_value = centering_value
        if self.spec.vector:
            n_leaves = max(len(self.spec.lik_coef), 1)
            self._bind([(self.data_dim,)] * n_leaves)
        elif estimator == "cv":
            self.flatten_params(centering_value, broadcast_chains=None)   # binds from the centre's shapes
        torch.cuda.current_stream().synchronize()                 # inputs may be freed by the caller now

    def _bind(self, leaf_shapes):
        """Create the native handle for a parameter geometry and ingest data (+ CV centre)."""
        shapes = [tuple(int(d) for d in s) for s in leaf_shapes]
        if not 1 <= len(shapes) <= _native.MAX_LEAVES:
            raise TypeError(f"parameter pytrees need 1..{_native.MAX_LEAVES} leaves")
        dims = self.spec.geometry(shapes, [tuple(d.shape) for d in self._data])   # raises before any state changes
        self.leaf_shapes, self.n_leaves = shapes, len(shapes)
        self.leaf_size = [int(np.prod(s)) for s in self.leaf_shapes]
        self.P = sum(self.leaf_size)

        cfg = _native.Config()
        cfg.abi_version = _native.ABI_VERSION
        cfg.family, cfg.estimator = self.spec.family, _EST[self.estimator]
        cfg.diffusion = DIFFUSION_IDS[self.diffusion]
        cfg.n_data, cfg.batch, cfg.n_leaves, cfg.data_dim = self.n_data, self.batch, self.n_leaves, self.data_dim
        for l, s in enumerate(self.leaf_size):
            cfg.leaf_size[l] = s
        cfg.leapfrog, cfg.update_rate, cfg.flags = self.leapfrog, self.update_rate, self.flags
        for l, a in enumerate(self.spec.lik_coef):
            cfg.lik_coef[l] = a
        for l, c in enumerate(self.spec.prior_coef):
            cfg.prior_coef[l] = c
        for k, v in enumerate(dims):
            cfg.dims[k] = int(v)
        order = {"psgld": ("alpha", "eps"), "sgldAdam": ("beta1", "beta2", "eps"), "adam_opt": ("beta1", "beta2", "eps"),
                 "sghmc": ("alpha",),
                 "sgnht": ("a",), "badodab": ("a",)}.get(self.diffusion, ())
        for k, name in enumerate(order):
            cfg.hyper[k] = float(self.hp[name])
        _native.check(self.lib.sgmcmc_create(C.byref(cfg), C.byref(self._handle)))
        if self.row_shard is not None:
            _native.check(self.lib.sgmcmc_set_row_shard(self._handle, int(self.row_shard[0]), self.n_local))
        y = self._data[1] if len(self._data) == 2 else None
        _native.check(self.lib.sgmcmc_set_data(self._handle, _vp(self._data[0]), _vp(y), self._stream()))
        if self.spec.vector:
            self._data = None                                     # the handle owns its re-laid-out copy
        if self.estimator == "cv":
            center = self.flatten_params(self._centering_value, broadcast_chains=None)[0]
            if center.shape[0] != 1:
                raise ValueError("centering_value must be a single parameter pytree")
            if self.row_shard is None:
                _native.check(self.lib.sgmcmc_set_centering(self._handle, _vp(center), self._stream()))
            else:       # gradient_estimation.py:70 over ALL rows: local rows here, one all-reduce across the ranks
                from . import parallel

                g = self.fullbatch_grad(center)
                if self.group is not False:
                    g = parallel.sharded_fullbatch_grad(g, parallel.prior_grad(self.spec, self.leaf_size, center), self.group)
                _native.check(self.lib.sgmcmc_set_centering_grad(self._handle, _vp(center), _vp(g), self._stream()))

    def __del__(self):
        try:
            if getattr(self, "_handle", None) and self._handle.value:
                self.lib.sgmcmc_destroy(self._handle)
                self._handle = C.c_void_p(0)
        except Exception:
            pass

    @staticmethod
    def _stream():
        import torch

        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ---- parameter pytrees <-> flat [C, P] -------------------------------------------------
    def flatten_params(self, params, broadcast_chains):
        """-> (flat CUDA f32 [C, P], treedef, batched).  A leaf with one axis more than its
        per-chain rank is multi-chain ([C, *leaf_shape])."""
        torch = _native.require_cuda()
        leaves, treedef = _tree.flatten(params)
        ts = [to_device(l, torch.float32) for l in leaves]
        ranks = self.spec.leaf_ndims(len(ts))
        if len(ranks) != len(ts) or any(t.dim() not in (r, r + 1) for t, r in zip(ts, ranks)):
            raise TypeError(f"parameter pytree does not fit the {self.spec.name} family")
        batched = ts[0].dim() == ranks[0] + 1
        if any((t.dim() == r + 1) != batched for t, r in zip(ts, ranks)):
            raise TypeError("either every parameter leaf carries a chain axis or none does")
        shapes = [tuple(t.shape[1:] if batched else t.shape) for t in ts]
        if self.leaf_shapes is None:
            self._bind(shapes)
        if len(shapes) != self.n_leaves:
            raise TypeError(f"parameter pytree has {len(shapes)} leaves, the model expects {self.n_leaves}")
        if shapes != self.leaf_shapes:                            # diffusion_util.py:58-64 raises TypeError too
            raise TypeError(f"parameter leaf shapes {shapes} do not match the model's {self.leaf_shapes}")
        ts = [t.reshape(-1, s) for t, s in zip(ts, self.leaf_size)]
        flat = torch.cat(ts, dim=1).contiguous()
        if broadcast_chains is not None and flat.shape[0] == 1 and broadcast_chains > 1:
            flat = flat.expand(broadcast_chains, -1).contiguous()
        return flat, treedef, batched

    def unflatten_params(self, flat, treedef, lead_shape):
        """flat [..., P] -> pytree with leaves [*lead_shape, *leaf_shape]."""
        leaves, off = [], 0
        for s, shp in zip(self.leaf_size, self.leaf_shapes):
            leaves.append(flat[..., off:off + s].reshape(*lead_shape, *shp))
            off += s
        return _tree.unflatten(treedef, leaves)

    # ---- state ------------------------------------------------------------------------------
    def new_state(self, theta):
        torch = _native.require_cuda()
        c = theta.shape[0]
        z = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device="cuda")
        st = {"theta": theta.clone(), "grad": z(c, self.P), "key": torch.zeros(c, 2, dtype=torch.int32, device="cuda")}
        if self.diffusion != "sgld":
            st["aux1"] = z(c, self.P)
        if self.diffusion in ("sgldAdam", "adam_opt"):
            st["aux2"] = z(c, self.P)
        if self.diffusion in ("sgnht", "badodab"):
            st["alpha"] = z(c, self.n_leaves)
        if self.estimator == "svrg":
            st["svrg_center"], st["svrg_grad"] = z(c, self.P), z(c, self.P)
        return st

    @staticmethod
    def clone_state(st):
        return {k: v.clone() for k, v in st.items()}

    def _cstate(self, st):
        s = _native.State()
        s.n_chains = int(st["theta"].shape[0])
        for name in ("theta", "aux1", "aux2", "alpha", "grad", "svrg_center", "svrg_grad", "key"):
            t = st.get(name)
            setattr(s, name, t.data_ptr() if t is not None else None)
        return s

    def _coefs(self, i0, k):
        torch = _native.require_cuda()
        if coef_is_constant(self.diffusion, self.schedule):
            key = ("const",)
            if key not in self._coef_cache:
                self._coef_cache[key] = torch.from_numpy(
                    step_coefs(self.diffusion, self.hp, self.schedule, 0, 1)).to("cuda")
            return self._coef_cache[key], 0
        return torch.from_numpy(step_coefs(self.diffusion, self.hp, self.schedule, i0, k)).to("cuda"), 1

    # ---- native calls -----------------------------------------------------------------------
    def init(self, st, keys, sampler_mode: bool):
        cs = self._cstate(st)
        _native.check(self.lib.sgmcmc_init(self._handle, C.byref(cs), _vp(keys), 1 if sampler_mode else 0,
                                           self._stream()))

    def run(self, st, i0: int, k: int, samples=None, sample_chain_stride: int = 0, thin: int = 1, moments=None):
        """k iterations from i0.  thin > 1 keeps every thin-th state (slot (s + 1) // thin - 1 of `samples`); `moments`
        (CUDA f32 [C, 3, P]: planes 0, 1 zeroed, plane 2 = a reference point) accumulates the sum and sum of squares of
        (kept state - reference)."""
        coef, stride = self._coefs(i0, k)
        cs = self._cstate(st)
        if thin == 1 and moments is None:
            _native.check(self.lib.sgmcmc_run(self._handle, C.byref(cs), int(i0), int(k), _vp(coef), stride,
                                              _vp(samples), int(sample_chain_stride), self._stream()))
        else:
            _native.check(self.lib.sgmcmc_run_thinned(self._handle, C.byref(cs), int(i0), int(k), _vp(coef), stride, int(thin),
                                                      _vp(samples), int(sample_chain_stride), _vp(moments), self._stream()))
        return coef   # keep alive until the caller synchronises

    def optimize(self, st, i0: int, k: int, lp_out=None):
        torch = _native.require_cuda()
        coef = torch.from_numpy(step_coefs(self.diffusion, self.hp, self.schedule, i0, k)).to("cuda")
        cs = self._cstate(st)
        _native.check(self.lib.sgmcmc_optimize(self._handle, C.byref(cs), int(i0), int(k), _vp(coef), _vp(lp_out), self._stream()))
        return coef

    def step(self, st, i: int, keys):
        coef, _ = self._coefs(i, 1)
        cs = self._cstate(st)
        _native.check(self.lib.sgmcmc_kernel_step(self._handle, C.byref(cs), int(i), _vp(keys), _vp(coef),
                                                  self._stream()))
        return coef

    # ---- data-sharded mode --------------------------------------------------------------------
    @property
    def sum_width(self) -> int:
        return int(self.lib.sgmcmc_sum_width(self._handle))

    def shard_init(self, st, keys, sampler_mode: bool):
        """init_fn up to this shard's partial minibatch sums -> CUDA f32 [C, DS]."""
        torch = _native.require_cuda()
        part = torch.empty((st["theta"].shape[0], self.sum_width), dtype=torch.float32, device="cuda")
        cs = self._cstate(st)
        _native.check(self.lib.sgmcmc_shard_init(self._handle, C.byref(cs), _vp(keys), 1 if sampler_mode else 0, _vp(part),
                                                 self._stream()))
        return part

    def shard_step(self, st, i: int, flags: int, sum_in, sample_out=None):
        """Finish the previous gradient from the complete sums `sum_in`; per flags run update2 / one transition.
        Returns this shard's partial sums of the new estimate (None without SHARD_STEP)."""
        torch = _native.require_cuda()
        coef, _ = self._coefs(i, 1)
        coef_prev, _ = self._coefs(max(i - 1, 0), 1)
        part = None
        if flags & _native.SHARD_STEP:
            part = torch.empty((st["theta"].shape[0], self.sum_width), dtype=torch.float32, device="cuda")
        cs = self._cstate(st)
        _native.check(self.lib.sgmcmc_shard_step(self._handle, C.byref(cs), int(i), _vp(coef), _vp(coef_prev), int(flags),
                                                 _vp(sum_in), _vp(part), _vp(sample_out), self._stream()))
        self._keep = (coef, coef_prev)
        return part

    def fullbatch_grad(self, thetas):
        torch = _native.require_cuda()
        if self.leaf_shapes is None:
            raise _native.NativeError("the parameter geometry is not bound yet: call init_fn / sampler / "
                                      "flatten_params with a parameter pytree first")
        out = torch.empty_like(thetas)
        _native.check(self.lib.sgmcmc_fullbatch_grad(self._handle, _vp(thetas), int(thetas.shape[0]), _vp(out),
                                                     self._stream()))
        return out
