"""GPU parity: the fused CUDA path (through the C ABI) against the numpy oracle, same seeded inputs.

Bars (BASELINE.json north_star):
* minibatch indices, key splits: BIT-EXACT
* single-step updates: <= 1e-5 relative (fp32), measured against the vector's max magnitude
* K-step trajectories: stated tolerance (grows with K; fp32 reduction order differs)
* KSD: <= 1e-4 relative against the fp64 literal formula
"""
import json
import os

import numpy as np
import pytest

from oracle import jax_prng as P
from oracle import ksd as OK
from oracle import models as M
from oracle import sgmcmc as S

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT = json.load(open(os.path.join(GOLD, "jax_prng_kat.json")))

RTOL_STEP = 1e-5      # single-step tolerance of north_star


def close(a, b, rtol=RTOL_STEP, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, f"{what}: max err / max|ref| = {err:.3e} > {rtol:.1e}"


# ------------------------------------------------------------------------------ PRNG
def test_split_bit_exact(cuda):
    from sgmcmcjax_b200 import random as R

    for v in KAT["external"]["split"] + KAT["self_derived"]["split"]:
        assert R.split(np.array(v["key"], np.uint32), v["num"]).tolist() == v["out"]
    for seed, num in [(1, 1), (7, 5), (123456, 1024), (2**31 + 5, 33)]:
        k = P.PRNGKey(seed)
        assert np.array_equal(R.split(k, num), P.split(k, num))


def test_normal_matches_external_kat_and_oracle(cuda):
    from sgmcmcjax_b200 import random as R

    for v in KAT["external"]["normal"]:
        out = R.normal(np.array(v["key"], np.uint32), tuple(v["shape"])).cpu().numpy().ravel()
        np.testing.assert_allclose(out, v["out"], rtol=3e-7)
    for seed, n in [(0, 100), (3, 101), (11, 1), (5, 79510)]:
        k = P.PRNGKey(seed)
        got = R.normal(k, (n,)).cpu().numpy()
        ref = P.normal(k, (n,))
        # identical bits -> identical uniforms; the polynomial is evaluated with FMAs on the device
        np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("n,batch", [(10_000, 1000), (60_000, 100), (40_000, 999), (65_536, 64), (65_537, 65),
                                     (1_000_000, 10_000), (100_000, 1000), (1000, 10), (7, 1), (3, 2), (1, 5),
                                     (2_147_483_647, 257)])
def test_minibatch_indices_bit_exact(cuda, n, batch):
    from sgmcmcjax_b200 import random as R

    for seed in (0, 42, 987654321):
        k = P.split(P.PRNGKey(seed))[1]
        got = R.choice_indices(k, n, batch).cpu().numpy()
        ref = P.choice_indices(k, n, batch)
        assert got.dtype == np.int32 and np.array_equal(got, ref), (n, batch, seed)


def test_minibatch_indices_frozen_vectors(cuda):
    from sgmcmcjax_b200 import random as R

    for v in KAT["self_derived"]["choice"]:
        idx = R.choice_indices(np.array(v["key"], np.uint32), v["n"], v["batch"]).cpu().numpy()
        assert idx[: len(v["head"])].tolist() == v["head"] and int(idx.astype(np.int64).sum()) == v["sum"]


# ------------------------------------------------------------------------------ fixtures
def gauss_case(two_leaves=False, n=1000, d=5):
    from sgmcmcjax_b200.models import gaussian_mean as G

    X = P.normal(P.PRNGKey(0), (n, d))                              # tests/models.py:28-30
    if two_leaves:
        fam = M.GaussianMean((0.5, 0.1), (0.001, 0.001))
        return fam, (G.loglikelihood_list_array, G.logprior_list_array), (X,), [np.zeros(d, np.float32),
                                                                              np.full(d, 0.1, np.float32)]
    return M.GaussianMean(), (G.loglikelihood, G.logprior), (X,), [np.full(d, 0.05, np.float32)]


def logreg_case(n=2000, d=10, seed=3):
    from sgmcmcjax_b200.models import logistic_regression as L

    ks = P.split(P.PRNGKey(seed), 4)
    theta_true = (np.float32(np.sqrt(10.0)) * P.normal(ks[0], (d,))).astype(np.float32)
    X = P.normal(ks[1], (n, d))
    p = 1.0 / (1.0 + np.exp(-(X @ theta_true).astype(np.float64)))
    y = (P.uniform(ks[2], (n,)) < p).astype(np.int32)
    th0 = (theta_true + np.float32(0.05) * P.normal(ks[3], (d,))).astype(np.float32)
    return M.LogisticRegression(), (L.loglikelihood, L.logprior), (X, y), [th0]


# (oracle name, dt, oracle kwargs, native builder name, native extra positional/kw)
KERNELS = [
    ("sgld", 1e-5, {}),
    ("sgldCV", 1e-4, {"cv": True}),
    ("sgld_SVRG", 1e-4, {"update_rate": 3}),
    ("psgld", 1e-3, {}),
    ("sgldAdam", 1e-3, {}),
    ("sgnht", 1e-5, {}),
    ("sgnhtCV", 1e-5, {"cv": True}),
    ("baoab", 1e-3, {"gamma": 5.0}),
    ("badodab", 1e-3, {}),
    ("badodabCV", 1e-3, {"cv": True}),
    ("sghmc", 1e-6, {"L": 3}),
    ("sghmcCV", 1e-6, {"L": 3, "cv": True}),
    ("sghmc_SVRG", 1e-6, {"L": 3, "update_rate": 3}),
]


def build_pair(name, dt, kw, fam, fns, data, batch, center, strict=False):
    """-> (oracle Kernel, native (init_fn, kernel, get_params))."""
    from sgmcmcjax_b200 import kernels as NK

    kw = dict(kw)
    okw, pos, nkw = {}, [], {}
    L = kw.pop("L", None)
    if kw.pop("cv", False):
        okw["centering_value"] = center
        pos.append(center[0] if len(center) == 1 else list(center))
    if "update_rate" in kw:
        okw["update_rate"] = kw["update_rate"]
        pos.append(kw.pop("update_rate"))
    if "gamma" in kw:
        okw["gamma"] = kw["gamma"]
        pos.append(kw.pop("gamma"))
    if L is not None:
        okw["L"] = L
    if strict:
        okw["strict_reference_bugs"] = True
        nkw["strict_reference_bugs"] = True
    ok = S.build_kernel(name, fam, data, batch, dt, **okw)
    builder = getattr(NK, f"build_{name}_kernel")
    args = ([dt, L] if L is not None else [dt]) + [fns[0], fns[1], data, batch] + pos
    return ok, builder(*args, **nkw)


def as_params(leaves):
    return leaves[0] if len(leaves) == 1 else list(leaves)


def leaves_of(p):
    return [np.asarray(l) for l in (p if isinstance(p, (list, tuple)) else [p])]


def to_np(t):
    return t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)


def compare_state(ostate, nstate, get_params, what, rtol=RTOL_STEP):
    """params / momenta / thermostat of chain 0 against the oracle; returns (grad_ref, grad_native)."""
    for j, (ref, got) in enumerate(zip(ostate.diffusion_state, leaves_of(get_params(nstate)))):
        close(got, ref["x"], rtol=rtol, what=f"{what} x leaf {j}")
    packed = nstate.diffusion_state.packed_state
    d = ostate.diffusion_state[0]["x"].shape[0]
    first = ostate.diffusion_state[0]
    bufs = {"m": "aux1", "v": "aux2"} if "m" in first else {"v": "aux1"}
    for fld, buf in bufs.items():
        if fld in first:
            got = to_np(packed[buf])[0]
            for j, ref in enumerate(ostate.diffusion_state):
                close(got[j * d:(j + 1) * d], ref[fld], rtol=2 * rtol, what=f"{what} {fld} leaf {j}")
    if "alpha" in first:
        got = to_np(packed["alpha"])[0]
        ref = np.array([s["alpha"] for s in ostate.diffusion_state], np.float64)
        np.testing.assert_allclose(got, ref, rtol=2 * rtol, err_msg=f"{what} alpha")
    g_ref = np.concatenate([np.ravel(g) for g in ostate.param_grad])
    pg = nstate.param_grad
    g_got = np.concatenate([to_np(l).ravel() for l in (pg if isinstance(pg, (list, tuple)) else [pg])])
    return g_ref, g_got


def inject(ostate, nstate):
    """Overwrite the native per-chain buffers with the oracle's state, so that the next native step
    starts from exactly the oracle's state (every step is then a strict single-step comparison)."""
    import torch

    packed = nstate.diffusion_state.packed_state
    first = ostate.diffusion_state[0]

    def put(name, leaves):
        flat = np.concatenate([np.ravel(np.asarray(l, np.float32)) for l in leaves])[None]
        packed[name].copy_(torch.from_numpy(flat))

    put("theta", [s["x"] for s in ostate.diffusion_state])
    if "m" in first:
        put("aux1", [s["m"] for s in ostate.diffusion_state])
        put("aux2", [s["v"] for s in ostate.diffusion_state])
    elif "v" in first:
        put("aux1", [s["v"] for s in ostate.diffusion_state])
    if "alpha" in first:
        put("alpha", [np.float32(s["alpha"]) for s in ostate.diffusion_state])
    put("grad", ostate.param_grad)
    if ostate.svrg_state is not None:
        put("svrg_center", ostate.svrg_state[0])
        put("svrg_grad", ostate.svrg_state[1])
    return nstate


def grad_tol(name):
    # CV / SVRG: (c + g) - gc cancels; the error is relative to the TERMS' magnitude,
    # which compare_state's max|ref| normalisation only approximates -> 3e-5 instead of 1e-5.
    return 3e-5 if ("CV" in name or "SVRG" in name) else 1e-5


@pytest.mark.parametrize("name,dt,kw", KERNELS, ids=[k[0] for k in KERNELS])
@pytest.mark.parametrize("case", ["gauss1", "gauss2", "logreg"])
def test_init_and_single_steps_match_oracle(cuda, name, dt, kw, case):
    """init_fn, then 4 kernel steps (i = 0..3).  Before every step the native state is overwritten with
    the oracle's, so each step is a strict single-step comparison (<= 1e-5) with the same per-step keys
    (kernels.py:48-64,110-122): params, momenta, thermostat, carried gradient, SVRG state."""
    if case == "logreg":
        fam, fns, data, th0 = logreg_case()
        batch = 100
    else:
        fam, fns, data, th0 = gauss_case(case == "gauss2")
        batch = 10                                                   # tests/test_samplers.py:31
    center = [np.asarray(t + np.float32(0.01)) for t in th0]
    ok, (init_fn, kernel, get_params) = build_pair(name, dt, kw, fam, fns, data, batch, center)
    key = P.PRNGKey(0)
    key, sub = P.split(key)
    ost = ok.init_fn(sub, th0)
    nst = init_fn(sub, as_params(th0))
    g_ref, g_got = compare_state(ost, nst, get_params, f"{name}/{case} init")
    assert np.abs(g_got - g_ref).max() / max(np.abs(g_ref).max(), 1.0) <= grad_tol(name), f"{name}/{case} init grad"
    for i in range(4):
        key, sub = P.split(key)
        nst = kernel(i, sub, inject(ost, nst))
        ost = ok.kernel(i, sub, ost)
        g_ref, g_got = compare_state(ost, nst, get_params, f"{name}/{case} step {i}")
        assert np.abs(g_got - g_ref).max() / max(np.abs(g_ref).max(), 1.0) <= grad_tol(name), f"{name}/{case} grad {i}"
        if ost.svrg_state is not None:
            packed = nst.diffusion_state.packed_state
            close(to_np(packed["svrg_center"])[0], np.concatenate(ost.svrg_state[0]), what="svrg centre")
            close(to_np(packed["svrg_grad"])[0], np.concatenate(ost.svrg_state[1]), rtol=2e-5, what="svrg fb grad")


@pytest.mark.parametrize("name,dt,kw", KERNELS, ids=[k[0] for k in KERNELS])
def test_free_running_trajectory_stays_close(cuda, name, dt, kw):
    """No state injection: 6 free-running steps of the kernel triple; the two fp32 trajectories may only
    drift by accumulated rounding (stated tolerance 1e-4 of max|theta|)."""
    fam, fns, data, th0 = gauss_case()
    ok, (init_fn, kernel, get_params) = build_pair(name, dt, kw, fam, fns, data, 10, th0)
    key = P.PRNGKey(1)
    key, sub = P.split(key)
    ost, nst = ok.init_fn(sub, th0), init_fn(sub, as_params(th0))
    for i in range(6):
        key, sub = P.split(key)
        ost, nst = ok.kernel(i, sub, ost), kernel(i, sub, nst)
    close(leaves_of(get_params(nst))[0], ost.diffusion_state[0]["x"], rtol=1e-4, what=f"{name} free-running")


@pytest.mark.parametrize("name,dt,kw", KERNELS, ids=[k[0] for k in KERNELS])
def test_single_step_from_oracle_state_1e5(cuda, name, dt, kw):
    """The strict single-step bar: native step 0 from the (shared) init state vs the oracle's step 0
    for params, momenta, thermostat and the carried gradient, logistic regression D=100."""
    fam, fns, data, th0 = logreg_case(n=5000, d=100, seed=9)
    center = [np.asarray(th0[0] * np.float32(0.99))]
    ok, (init_fn, kernel, get_params) = build_pair(name, dt, kw, fam, fns, data, 500, center)
    sub_i, sub_0 = P.split(P.PRNGKey(77))
    ost = ok.kernel(0, sub_0, ok.init_fn(sub_i, th0))
    nst = kernel(0, sub_0, init_fn(sub_i, as_params(th0)))
    g_ref, g_got = compare_state(ost, nst, get_params, f"{name} step0")
    assert np.abs(g_got - g_ref).max() / np.abs(g_ref).max() <= 3e-5


def test_badodabcv_strict_reference_bug(cuda):
    """kernels.py:457-460 literal wiring: x never moves, v accumulates dt/2 * g twice per step."""
    fam, fns, data, th0 = gauss_case()
    ok, (init_fn, kernel, get_params) = build_pair("badodabCV", 1e-3, {"cv": True}, fam, fns, data, 10, th0, strict=True)
    key = P.PRNGKey(5)
    key, sub = P.split(key)
    ost, nst = ok.init_fn(sub, th0), init_fn(sub, as_params(th0))
    for i in range(3):
        key, sub = P.split(key)
        ost, nst = ok.kernel(i, sub, ost), kernel(i, sub, nst)
    assert np.array_equal(leaves_of(get_params(nst))[0], th0[0])
    close(nst.diffusion_state.packed_state["aux1"].cpu().numpy()[0], ost.diffusion_state[0]["v"], rtol=3e-5, what="v")


# ------------------------------------------------------------------------------ samplers
SAMPLERS = [k for k in KERNELS if k[0] != "sgldAdam"] + [("sgldAdam", 1e-3, {})]


def build_sampler_pair(name, dt, kw, fam, fns, data, batch, center, **skw):
    from sgmcmcjax_b200 import samplers as NS

    kw = dict(kw)
    okw, pos = {}, []
    L = kw.pop("L", None)
    if kw.pop("cv", False):
        okw["centering_value"] = center
        pos.append(as_params(center))
    if "update_rate" in kw:
        okw["update_rate"] = kw["update_rate"]
        pos.append(kw.pop("update_rate"))
    if "gamma" in kw:
        okw["gamma"] = kw["gamma"]
        pos.append(kw.pop("gamma"))
    if L is not None:
        okw["L"] = L
    ok = S.build_kernel(name, fam, data, batch, dt, **okw)
    args = ([dt, L] if L is not None else [dt]) + [fns[0], fns[1], data, batch] + pos
    return ok, getattr(NS, f"build_{name}_sampler")(*args, pbar=False, **skw)


@pytest.mark.parametrize("name,dt,kw", SAMPLERS, ids=[k[0] for k in SAMPLERS])
@pytest.mark.parametrize("two_leaves", [False, True])
def test_samplers_container_shape_no_nan_and_trajectory(cuda, name, dt, kw, two_leaves):
    """tests/test_samplers.py:95-117 (container type, shapes, no NaN) + 10-step trajectory parity."""
    fam, fns, data, th0 = gauss_case(two_leaves)
    ok, sampler = build_sampler_pair(name, dt, kw, fam, fns, data, 10, th0)
    ref, _, _ = S.run_sampler(ok, P.PRNGKey(0), 10, th0)
    out = sampler(P.PRNGKey(0), 10, as_params(th0))
    if two_leaves:
        assert isinstance(out, list) and len(out) == 2
    else:
        assert isinstance(out, np.ndarray)
    for got, r in zip(leaves_of(out), ref):
        assert got.shape == (10, 5) and np.isfinite(got).all()
        close(got, r, rtol=1e-4, what=f"{name} 10-step trajectory")


def test_noncompiled_sampler_returns_list_of_pytrees(cuda):
    """samplers.py:81-93."""
    fam, fns, data, th0 = gauss_case(True)
    ok, sampler = build_sampler_pair("sgld", 1e-5, {}, fam, fns, data, 10, th0, compiled=False)
    out = sampler(P.PRNGKey(1), 6, list(th0))
    ref, _, _ = S.run_sampler(ok, P.PRNGKey(1), 6, th0)
    assert isinstance(out, list) and len(out) == 6 and isinstance(out[0], list) and len(out[0]) == 2
    for i in range(6):
        for j in range(2):
            close(out[i][j], ref[j][i], rtol=1e-4, what="noncompiled")


def test_pbar_chunked_launches_equal_single_launch(cuda):
    from sgmcmcjax_b200 import samplers as NS

    fam, fns, data, th0 = logreg_case()
    a = NS.build_sgld_sampler(1e-4, fns[0], fns[1], data, 100, pbar=False)(P.PRNGKey(2), 45, th0[0])
    b = NS.build_sgld_sampler(1e-4, fns[0], fns[1], data, 100, pbar=True)(P.PRNGKey(2), 45, th0[0])
    assert np.array_equal(a, b)                                    # resume is bit-reproducible


@pytest.mark.parametrize("name,dt,kw", [KERNELS[0], KERNELS[1], KERNELS[10]], ids=["sgld", "sgldCV", "sghmc"])
def test_thinning_and_on_device_summaries(cuda, name, dt, kw):
    """SURVEY section 8 f3 (sgmcmc_run_thinned): the kept states are exactly every thin-th state of the unthinned run, for
    any launch chunking; the on-device mean / variance equal numpy's over the kept states; the oracle agrees."""
    fam, fns, data, th0 = logreg_case()
    ok, full_sampler = build_sampler_pair(name, dt, kw, fam, fns, data, 100, th0)
    _, thin_sampler = build_sampler_pair(name, dt, kw, fam, fns, data, 100, th0, thin=4, summaries=True)
    full = full_sampler(P.PRNGKey(5), 42, th0[0])
    kept, summ = thin_sampler(P.PRNGKey(5), 42, th0[0])
    assert kept.shape == (10, th0[0].shape[0]) and summ["count"] == 10
    assert np.array_equal(kept, full[3::4][:10])
    ref, _, _ = S.run_sampler(ok, P.PRNGKey(5), 42, th0)
    close(kept, ref[0][3::4][:10], rtol=2e-4, what=f"{name} thinned vs oracle")
    close(summ["mean"], kept.astype(np.float64).mean(0), rtol=1e-6, what="on-device mean")
    np.testing.assert_allclose(summ["var"], kept.astype(np.float64).var(0), rtol=2e-3, atol=1e-9)
    # progress-bar chunking (launches rounded to multiples of thin) keeps the same states
    import sgmcmcjax_b200.samplers as NS
    args = [dt] + ([kw["L"]] if "L" in kw else []) + [fns[0], fns[1], data, 100] + ([th0[0]] if kw.get("cv") else [])
    chunked = getattr(NS, f"build_{name}_sampler")(*args, pbar=True, thin=4)(P.PRNGKey(5), 42, th0[0])
    assert np.array_equal(chunked, kept)
    # many chains, summaries only: no sample buffer at all
    keys = P.split(P.PRNGKey(6), 5)
    only = getattr(NS, f"build_{name}_sampler")(*args, pbar=False, summaries=True, store_samples=False)
    none, s5 = only(keys, 30, np.tile(th0[0], (5, 1)))
    assert none is None and s5["mean"].shape == (5, th0[0].shape[0]) and s5["count"] == 30
    one = full_sampler(keys[2], 30, th0[0])
    close(s5["mean"][2], one.astype(np.float64).mean(0), rtol=1e-6, what="summaries-only mean")


def test_rhat_from_on_device_summaries(cuda):
    """diagnostics.rhat on the kernel's per-chain moments == R-hat of the stored draws (README Gaussian, 8 chains)."""
    from sgmcmcjax_b200 import diagnostics as D
    from sgmcmcjax_b200 import samplers as NS

    fam, fns, data, th0 = gauss_case(n=2000, d=5)
    keys = P.split(P.PRNGKey(11), 8)
    start = np.tile(th0[0], (8, 1)) + np.float32(0.3) * P.normal(P.PRNGKey(12), (8, 5))
    kept, summ = NS.build_sgld_sampler(1e-4, fns[0], fns[1], data, 100, pbar=False, summaries=True)(keys, 400, start)
    want = D.rhat_from_samples(kept)
    np.testing.assert_allclose(D.rhat(summ), want, rtol=1e-4)
    assert (want < 1.2).all()                                     # dt = 1e-4, 400 steps: the chains have mixed


def test_multichain_equals_independent_single_chains(cuda):
    """vmap(sampler, in_axes=(0, None, 0)) semantics: chain c == single-chain run with keys[c]."""
    from sgmcmcjax_b200 import samplers as NS

    fam, fns, data, th0 = logreg_case()
    sampler = NS.build_sgldCV_sampler(1e-4, fns[0], fns[1], data, 100, th0[0], pbar=False)
    keys = P.split(P.PRNGKey(0), 5)
    th = np.stack([th0[0] + np.float32(0.01 * c) for c in range(5)])
    multi = sampler(keys, 20, th)
    assert multi.shape == (5, 20, 10)
    for c in range(5):
        single = sampler(keys[c], 20, th[c])
        assert np.array_equal(multi[c], single)
    ok = S.build_kernel("sgldCV", fam, data, 100, 1e-4, centering_value=th0)
    ref, _, _ = S.run_sampler(ok, keys[3], 20, [th[3]])
    close(multi[3], ref[0], rtol=1e-4, what="chain 3 vs oracle")


@pytest.mark.parametrize("d", [1, 3, 100, 130, 257, 512])
def test_row_widths_and_padding(cuda, d):
    """D not a multiple of 4, D > 128 (2 and 4 float4 chunks per lane)."""
    fam, fns, data, th0 = logreg_case(n=700, d=d, seed=d)
    ok, sampler = build_sampler_pair("sgld", 1e-5, {}, fam, fns, data, 33, th0)
    ref, _, _ = S.run_sampler(ok, P.PRNGKey(4), 3, th0)
    close(sampler(P.PRNGKey(4), 3, th0[0]), ref[0], rtol=5e-5, what=f"D={d}")


@pytest.mark.parametrize("batch", [1, 2, 63, 64, 65, 129, 1000])
def test_batch_sizes_incl_odd_and_full_batch(cuda, batch):
    """Odd batches exercise the zero pad counter; batch == N is the static full-batch bypass
    (gradient_estimation.py:41-42) whose init_gradient still subsamples (:31-35)."""
    fam, fns, data, th0 = gauss_case()
    ok, sampler = build_sampler_pair("sgld", 1e-5, {}, fam, fns, data, batch, th0)
    ref, _, _ = S.run_sampler(ok, P.PRNGKey(8), 4, th0)
    close(sampler(P.PRNGKey(8), 4, th0[0]), ref[0], rtol=5e-5, what=f"batch={batch}")


def test_schedule_step_sizes(cuda):
    from sgmcmcjax_b200 import diffusions as ND
    from sgmcmcjax_b200 import samplers as NS

    from oracle import diffusions as OD

    fam, fns, data, th0 = gauss_case()
    sampler = NS.build_sgld_sampler(ND.welling_teh_schedule(1e-4, 10.0), fns[0], fns[1], data, 10, pbar=False)
    ok = S.build_kernel("sgld", fam, data, 10, OD.welling_teh_schedule(1e-4, 10.0))
    ref, _, _ = S.run_sampler(ok, P.PRNGKey(0), 12, th0)
    close(sampler(P.PRNGKey(0), 12, th0[0]), ref[0], rtol=1e-4, what="welling-teh schedule")


def test_fullbatch_grad_matches_oracle_and_reference_test(cuda):
    """tests/test_gradient_estimation.py:20-30 (full-batch estimator == grad_log_post) + K2 vs oracle."""
    import torch

    from sgmcmcjax_b200 import kernels as NK

    for fam, fns, data, th0 in (gauss_case(True), logreg_case(n=5000, d=100)):
        init_fn, kernel, get_params = NK.build_sgld_kernel(1e-5, fns[0], fns[1], data, 10)
        eng = init_fn.engine
        thetas = np.stack([np.concatenate(th0) * np.float32(s) for s in (1.0, 0.5, -0.3)])
        got = eng.fullbatch_grad(torch.from_numpy(thetas).cuda()).cpu().numpy()
        for m in range(3):
            d = th0[0].shape[0]
            leaves = [thetas[m][j * d:(j + 1) * d] for j in range(len(th0))]
            ref = np.concatenate(fam.grad_log_post(leaves, data, data[0].shape[0]))
            close(got[m], ref, rtol=2e-5, what="fullbatch grad")


@pytest.mark.parametrize("d", [3, 100, 124, 128, 250])
def test_cv_row_layout_with_center_column(cuda, d):
    """logreg + CV estimator: the row table carries center . x~ in an extra float4 (csrc/vec_sampler.cuh CV == 2).  The
    full-batch gradient paths (SIMT and tensor-core) must ignore that column, a step must match the oracle at widths
    where the extra float4 starts a new 32-lane chunk, and re-centring must refresh the column."""
    import torch

    from sgmcmcjax_b200 import kernels as NK

    fam, fns, data, th0 = logreg_case(n=1500, d=d, seed=d)
    ok, (init_fn, kernel, get_params) = build_pair("sgldCV", 1e-4, {"cv": True}, fam, fns, data, 64, th0)
    plain, _, _ = NK.build_sgld_kernel(1e-5, fns[0], fns[1], data, 64)
    rng = np.random.default_rng(d)
    for m in (2, 70):                                              # SIMT path, tensor-core path
        thetas = torch.from_numpy((th0[0][None, :] + 0.05 * rng.normal(size=(m, d))).astype(np.float32)).cuda()
        a, b = init_fn.engine.fullbatch_grad(thetas).cpu().numpy(), plain.engine.fullbatch_grad(thetas).cpu().numpy()
        close(a, b, rtol=2e-5, what=f"K2 with the centre column, m={m}")     # summation order differs (chunking, atomics)
    key = P.PRNGKey(d)
    start = [(th0[0] * np.float32(1.01)).astype(np.float32)]
    ost, nst = ok.init_fn(key, start), init_fn(key, start[0])
    compare_state(ost, nst, get_params, f"cv init d={d}")
    for i in range(3):
        k = P.split(P.PRNGKey(100 + i))[1]
        nst = kernel(i, k, inject(ost, nst))
        ost = ok.kernel(i, k, ost)
        g_ref, g_got = compare_state(ost, nst, get_params, f"cv step {i} d={d}")
        assert np.abs(g_got - g_ref).max() / max(np.abs(g_ref).max(), 1.0) <= grad_tol("sgldCV")
    # at theta == centre the two weights cancel bit for bit: the gradient estimate IS the full-batch gradient
    _, g0 = compare_state(ok.init_fn(key, th0), init_fn(key, th0[0]), get_params, "cv at centre")
    fb = init_fn.engine.fullbatch_grad(torch.from_numpy(th0[0][None, :]).cuda()).cpu().numpy()[0]
    assert np.array_equal(g0, fb)


@pytest.mark.parametrize("n,d,m", [(3000, 100, 130), (777, 10, 64), (5000, 37, 300), (128, 100, 128), (20000, 100, 1000)])
def test_fullbatch_grad_tensor_core_path(cuda, n, d, m):
    """K2 at many parameter vectors (csrc/k2_gemm.cuh: two tcgen05 3xTF32 contractions, the imq_KSD `grads` producer)
    against the oracle's grad_log_post and against the row-streaming SIMT path, ragged shapes included."""
    import torch

    from sgmcmcjax_b200 import kernels as NK

    fam, fns, data, th0 = logreg_case(n=n, d=d, seed=n + d)
    init_fn, _, _ = NK.build_sgld_kernel(1e-5, fns[0], fns[1], data, 10)
    eng = init_fn.engine
    rng = np.random.default_rng(m)
    thetas = (th0[0][None, :] + 0.05 * rng.normal(size=(m, d))).astype(np.float32)
    got = eng.fullbatch_grad(torch.from_numpy(thetas).cuda()).cpu().numpy()        # m >= 64 -> tensor cores
    os.environ["SGMCMC_K2_TC"] = "0"
    try:
        simt = eng.fullbatch_grad(torch.from_numpy(thetas[:5].copy()).cuda()).cpu().numpy()
    finally:
        del os.environ["SGMCMC_K2_TC"]
    for j in (0, 1, 4, m // 2, m - 1):
        ref = fam.grad_log_post([thetas[j]], data, n, np.float64)[0]
        close(got[j], ref, rtol=2e-5, what=f"K2 tensor-core row {j}")
    close(got[:5], simt, rtol=2e-5, what="K2 tensor-core vs SIMT")
    # the same with a 1 MB budget for the intermediate: blocks of 128 parameter vectors x 1024 rows (the blocking that keeps
    # W = -sigmoid(Theta X~^T) in L2 between the two contractions, exercised here with ragged last blocks in both directions)
    os.environ["SGMCMC_K2_W_MB"] = "1"
    try:
        small = eng.fullbatch_grad(torch.from_numpy(thetas).cuda()).cpu().numpy()
    finally:
        del os.environ["SGMCMC_K2_W_MB"]
    close(small, got, rtol=2e-5, what="K2 blocked (1 MB) vs default blocking")


@pytest.mark.parametrize("name,dt,kw", [KERNELS[0], KERNELS[1], KERNELS[3], KERNELS[5], KERNELS[7], KERNELS[9]],
                         ids=["sgld", "sgldCV", "psgld", "sgnht", "baoab", "badodabCV"])
@pytest.mark.parametrize("case", ["gauss2", "logreg"])
def test_data_sharded_mode_equals_unsharded(cuda, name, dt, kw, case):
    """Optional data-sharded mode: the rows split into 3 uneven shards (here all on one GPU, their partial sums added on
    the device - across ranks the same sum is one all-reduce): same chains as the unsharded sampler and the oracle."""
    from sgmcmcjax_b200 import kernels as NK
    from sgmcmcjax_b200 import parallel

    if case == "logreg":
        fam, fns, data, th0 = logreg_case(n=1500, d=10)
        batch = 200
    else:
        fam, fns, data, th0 = gauss_case(True, n=1000)
        batch = 64
    center = [np.asarray(t + np.float32(0.01)) for t in th0]
    n = data[0].shape[0]
    cuts = [0, n // 5, n // 5 + n // 2, n]
    shards = [(lo, tuple(a[lo:hi] for a in data)) for lo, hi in zip(cuts[:-1], cuts[1:])]
    kw = dict(kw)
    pos = []
    okw = {}
    if kw.pop("cv", False):
        okw["centering_value"] = center
        pos.append(as_params(center))
    if "gamma" in kw:
        okw["gamma"] = kw["gamma"]
        pos.append(kw.pop("gamma"))
    builder = getattr(NK, f"build_{name}_kernel")
    sampler = parallel.build_data_sharded_sampler(builder, dt, fns[0], fns[1], data, batch, *pos, shards=shards, n_total=n,
                                                  distributed=False)
    keys = P.split(P.PRNGKey(3), 2)
    th = [np.stack([t, t + np.float32(0.02)]) for t in th0]
    got = leaves_of(sampler(keys, 7, as_params(th)))
    ok = S.build_kernel(name, fam, data, batch, dt, **okw)
    for c in range(2):
        ref, _, _ = S.run_sampler(ok, keys[c], 7, [t[c] for t in th])
        for g, r in zip(got, ref):
            assert g.shape[:2] == (2, 7)
            close(g[c], r, rtol=1e-4, what=f"sharded {name}/{case} chain {c}")
    # the finished gradient in the states of all shards is the complete one
    g0 = to_np(sampler.states[0]["grad"])
    for st in sampler.states[1:]:
        assert np.array_equal(g0, to_np(st["grad"]))


def test_error_behaviour_matches_reference(cuda):
    from sgmcmcjax_b200 import samplers as NS
    from sgmcmcjax_b200.models import gaussian_mean as G

    X = np.zeros((20, 3), np.float32)
    with pytest.raises(AssertionError):                            # gradient_estimation.py:25
        NS.build_sgld_sampler(1e-5, G.loglikelihood, G.logprior, [X], 2)
    with pytest.raises(ValueError):                                # util.py:34-39
        NS.build_sgld_sampler(1e-5, G.loglikelihood, G.logprior, (X, X, X), 2)
    with pytest.raises(NotImplementedError):
        NS.build_sgld_sampler(1e-5, lambda t, x: 0.0, lambda t: 0.0, (X,), 2)
    s = NS.build_sgld_sampler(1e-5, G.loglikelihood, G.logprior, (X,), 2, pbar=False)
    with pytest.raises(TypeError):                                 # diffusion_util.py:58-64 (tree mismatch)
        s(P.PRNGKey(0), 3, [np.zeros(3, np.float32), np.zeros(3, np.float32)])
    assert s(P.PRNGKey(0), 0, np.zeros(3, np.float32)).shape == (0, 3)


# ------------------------------------------------------------------------------ statistics at full size
def test_readme_gaussian_posterior_moments_full_config(cuda):
    """BASELINE config 1 (README.md:29-40): N=10k, D=100, batch=1k, dt=1e-5, 10k samples.
    examples/gauss_sample_sgld.py:36 asserts mean within 0.1; we also check the analytic stationary
    mean / variance and the oracle's first samples."""
    from sgmcmcjax_b200 import samplers as NS
    from sgmcmcjax_b200.models import gaussian_mean as G

    reg = json.load(open(os.path.join(GOLD, "oracle_regression.json")))
    X = P.normal(P.PRNGKey(0), (10000, 100))
    sampler = NS.build_sgld_sampler(1e-5, G.loglikelihood, G.logprior, (X,), 1000, pbar=False)
    out = sampler(P.PRNGKey(0), 10000, np.zeros(100, np.float32))
    assert out.shape == (10000, 100)
    np.testing.assert_allclose(out[:3, :3], reg["readme_sgld_first_samples_head"], rtol=1e-4, atol=1e-6)
    post_mean = X.sum(0, dtype=np.float64) / (10000 + 0.01)
    assert np.allclose(out.mean(0), post_mean, atol=0.1)           # the reference's own bar
    assert np.abs(out[1000:].mean(0) - post_mean).max() < 5e-3
    var = out[1000:].var(0).mean()
    assert abs(var - 1.58e-4) / 1.58e-4 < 0.08


def test_logreg_c2_shape_size_independent_properties(cuda, monkeypatch):
    """BASELINE config 2 geometry (N=1e6, D=100, B=1e4) at a reduced chain count: chain results are
    independent of how chains are grouped into launches (bit for bit at a fixed launch shape: the shape the library
    picks depends on the chain count, and with it the summation order), and SGLD-CV started at the centre with dt -> 0
    keeps theta at the centre with grad == full-batch gradient (CV exactness at theta == theta_hat)."""
    monkeypatch.setenv("SGMCMC_CLUSTER", "4")
    monkeypatch.setenv("SGMCMC_WARPS_PER_CHAIN", "16")
    import torch

    from sgmcmcjax_b200 import random as R
    from sgmcmcjax_b200 import samplers as NS
    from sgmcmcjax_b200.models import logistic_regression as L

    theta_true, X, y = L.gen_data(P.PRNGKey(42), 100, 1_000_000)
    th = theta_true.cpu().numpy()
    keys = P.split(P.PRNGKey(0), 8)
    s = NS.build_sgld_sampler(1e-5, L.loglikelihood, L.logprior, (X, y), 10_000, pbar=False)
    a = s(keys, 5, np.tile(th, (8, 1)))
    b = np.concatenate([s(keys[:3], 5, np.tile(th, (3, 1))), s(keys[3:], 5, np.tile(th, (5, 1)))])
    assert np.array_equal(a, b)
    # indices used on device == oracle indices (bit-exact) at this geometry
    k2 = np.array(KAT["self_derived"]["readme_keys"]["k2"], np.uint32)
    assert np.array_equal(R.choice_indices(k2, 1_000_000, 10_000).cpu().numpy(), P.choice_indices(k2, 1_000_000, 10_000))
    # oracle parity of one chain for 3 steps at full N / B
    Xh, yh = X.cpu().numpy(), y.cpu().numpy()
    ok = S.build_kernel("sgld", M.LogisticRegression(), (Xh, yh), 10_000, 1e-5)
    ref, _, _ = S.run_sampler(ok, keys[0], 3, [th])
    close(a[0][:3], ref[0], rtol=2e-5, what="C2 chain 0 vs oracle")
    # CV exactness at the centre
    from sgmcmcjax_b200 import kernels as NK
    init_fn, kernel, get_params = NK.build_sgldCV_kernel(1e-30, L.loglikelihood, L.logprior, (X, y), 10_000, th)
    st = init_fn(keys[1], th)
    fb = init_fn.engine.fullbatch_grad(torch.from_numpy(th[None]).cuda()).cpu().numpy()[0]
    close(to_np(st.param_grad), fb, rtol=1e-6, what="CV grad at centre == full-batch grad")


# ------------------------------------------------------------------------------ KSD
def ksd_case(n, d, kind="posterior"):
    rng = np.random.default_rng(n * 1000 + d)
    if kind == "posterior":          # tightly clustered samples far from the origin, gradients ~ -precision * (x - mode)
        X = (0.05 * rng.normal(size=(n, d)) + 1.0).astype(np.float32)
        G = (-(X - 1.0) / 0.0025 * 1e-2 + 0.1 * rng.normal(size=(n, d))).astype(np.float32)
    elif kind == "logreg":           # posterior sd ~1e-3 around |theta| ~ 3, gradients O(1e2) with a common offset
        X = (3.0 * rng.normal(size=(1, d)) + 1e-3 * rng.normal(size=(n, d))).astype(np.float32)
        G = (-1e5 * (X - X.mean(0)) + 30.0 + 50.0 * rng.normal(size=(n, d))).astype(np.float32)
    else:                            # spread-out samples, dd >> 1
        X = rng.normal(size=(n, d)).astype(np.float32) * 3.0
        G = (-X + rng.normal(size=(n, d))).astype(np.float32)
    return X, G


@pytest.mark.parametrize("method", ["tc", "simt"])
@pytest.mark.parametrize("n,d", [(1, 3), (17, 1), (64, 100), (129, 100), (200, 100), (333, 7), (1000, 100), (700, 260)])
@pytest.mark.parametrize("kind", ["posterior", "logreg", "wide"])
def test_ksd_matches_fp64_literal(cuda, n, d, kind, method):
    from sgmcmcjax_b200 import ksd as NKSD

    X, G = ksd_case(n, d, kind)
    ref = OK.imq_ksd_literal(X, G)
    got = float(NKSD.imq_KSD(X, G, method=method))
    assert abs(got - ref) <= 1e-4 * abs(ref), (got, ref)


def test_posterior_ksd_pipeline(cuda):
    """samples -> K2 gradients over all rows (tensor cores at M >= 64) -> imq_KSD, against the oracle end to end."""
    from sgmcmcjax_b200 import kernels as NK
    from sgmcmcjax_b200 import ksd as NKSD

    fam, fns, data, th0 = logreg_case(n=4000, d=12, seed=21)
    init_fn, _, _ = NK.build_sgld_kernel(1e-5, fns[0], fns[1], data, 10)
    rng = np.random.default_rng(5)
    pts = (th0[0][None, :] + 0.05 * rng.normal(size=(150, 12))).astype(np.float32)
    got = float(NKSD.posterior_KSD(init_fn.engine, pts))
    grads = np.stack([fam.grad_log_post([p], data, 4000, np.float64)[0] for p in pts])
    ref = OK.imq_ksd_literal(pts, grads)
    assert abs(got - ref) <= 1e-4 * abs(ref), (got, ref)


def test_ksd_tc_tile_partitions_sum_to_whole(cuda):
    from sgmcmcjax_b200 import ksd as NKSD

    X, G = ksd_case(1100, 100)
    whole = float(NKSD.imq_KSD_tiles(X, G, 0, 1).cpu()[0])
    for nparts in (2, 3, 8, 64):
        parts = sum(float(NKSD.imq_KSD_tiles(X, G, p, nparts).cpu()[0]) for p in range(nparts))
        assert abs(whole - parts) <= 1e-9 * abs(whole)
    ref = float(OK.imq_ksd_sum(X, G))
    assert abs(whole - ref) <= 2e-4 * abs(ref)
    simt = float(NKSD.imq_KSD_partial(X, G, 0, 1100).cpu()[0])
    assert abs(whole - simt) <= 2e-4 * abs(ref)


@pytest.mark.parametrize("n,d", [(1, 3), (129, 100), (333, 7), (1000, 100), (1100, 100), (700, 260)])
def test_ksd_tc_cluster_multicast_variant(cuda, monkeypatch, n, d):
    """The opt-in two-CTA-cluster kernel (row-side operand by TMA multicast, csrc/ksd_tc.cuh): same sums as one CTA per tile,
    odd and even tile-row lengths (the odd tile out is a weight-0 repeat), and its item partitions still sum to the whole."""
    from sgmcmcjax_b200 import ksd as NKSD

    X, G = ksd_case(n, d, "logreg")
    monkeypatch.setenv("SGMCMC_KSD_CLUSTER", "1")
    base = float(NKSD.imq_KSD_tiles(X, G, 0, 1).cpu()[0])
    monkeypatch.setenv("SGMCMC_KSD_CLUSTER", "2")
    whole = float(NKSD.imq_KSD_tiles(X, G, 0, 1).cpu()[0])
    assert abs(whole - base) <= 1e-9 * abs(base), (whole, base)
    for nparts in (2, 5):
        parts = sum(float(NKSD.imq_KSD_tiles(X, G, p, nparts).cpu()[0]) for p in range(nparts))
        assert abs(whole - parts) <= 1e-9 * abs(whole)
    ref = OK.imq_ksd_literal(X, G)
    assert abs(float(NKSD.imq_KSD(X, G, method="tc")) - ref) <= 1e-4 * abs(ref)


def test_ksd_row_blocks_sum_to_whole_and_golden(cuda):
    from sgmcmcjax_b200 import ksd as NKSD

    rng = np.random.default_rng(0)
    X = rng.normal(size=(301, 20)).astype(np.float32)
    G = (-X).astype(np.float32)
    whole = float(NKSD.imq_KSD_partial(X, G, 0, 301).cpu()[0])
    parts = sum(float(NKSD.imq_KSD_partial(X, G, *NKSD.row_block(301, r, 4)).cpu()[0]) for r in range(4))
    assert abs(whole - parts) <= 1e-9 * abs(whole)
    assert abs(whole - float(OK.imq_ksd_sum(X, G))) <= 1e-5 * abs(whole)
    assert float(NKSD.imq_KSD_partial(X, G, 5, 5).cpu()[0]) == 0.0


# ------------------------------------------------------------------------------ data generation
def test_gen_data_matches_reference_key_tree(cuda):
    """models.logistic_regression.gen_data == the restatement of logistic_regression.py:30-62 (genCovMat, Cholesky mixing,
    per-row Bernoulli keys): identical PRNG draws, fp32 library rounding in the Cholesky / matmul only."""
    from oracle import gen_data as OG
    from sgmcmcjax_b200 import random as R
    from sgmcmcjax_b200.models import logistic_regression as L

    for seed, dim, n in [(42, 10, 5000), (7, 100, 2000), (0, 3, 1)]:
        key = P.PRNGKey(seed)
        th_r, X_r, y_r = OG.gen_data(key, dim, n)
        th, X, y = L.gen_data(key, dim, n)
        np.testing.assert_allclose(th.cpu().numpy(), th_r, rtol=2e-6, atol=1e-7)
        np.testing.assert_allclose(L.genCovMat(P.split(key, 4)[2], dim, 0.4), OG.gen_cov_mat(P.split(key, 4)[2], dim, 0.4), rtol=0, atol=0)
        close(X.cpu().numpy(), X_r, rtol=1e-5, what="X = Z chol(Sigma)")
        yy = y.cpu().numpy()
        assert yy.dtype == np.int32 and (yy != y_r).mean() <= 2e-3          # flips only where |u - p| is at rounding level
    # the uniform / per-row Bernoulli primitives are bit-exact
    k = P.PRNGKey(5)
    assert np.array_equal(R.uniform(k, (1001,)).cpu().numpy(), P.uniform(k, (1001,)))
    p = P.uniform(P.PRNGKey(6), (3000,))
    import torch
    assert np.array_equal(R.bernoulli_rows(k, torch.from_numpy(p)).cpu().numpy(), OG.bernoulli_rows(k, p))


def test_gaussian_mean_far_from_zero_does_not_cancel(cuda):
    """Data whose mean is 1000 standard deviations from 0: the gradient lik2 (sum x - B theta) is formed from the row SUM,
    so without the column-mean shift of the row table (VecModel::mu) it cancels at ~B |x| 2^-24 in fp32.  Against the fp64
    evaluation of util.py:43-44 on the same minibatch, and one sampler step against the oracle."""
    from sgmcmcjax_b200 import kernels as NK
    from sgmcmcjax_b200.models import gaussian_mean as G

    n, d, batch = 10_000, 16, 1000
    X = (np.float32(1000.0) + P.normal(P.PRNGKey(21), (n, d))).astype(np.float32)
    th0 = np.full(d, 1000.1, np.float32)
    fam = M.GaussianMean()
    init_fn, kernel, get_params = NK.build_sgld_kernel(1e-7, G.loglikelihood, G.logprior, (X,), batch)
    sub = P.split(P.PRNGKey(3))[1]
    st = init_fn(sub, th0)
    idx = P.choice_indices(sub, n, batch)
    ref = fam.grad_log_post([th0], (X[idx],), n, np.float64)[0]
    got = to_np(st.param_grad)
    assert np.abs(got - ref).max() / np.abs(ref).max() <= 1e-5, np.abs(got - ref).max() / np.abs(ref).max()
    fb = init_fn.engine.fullbatch_grad(__import__("torch").from_numpy(th0[None]).cuda()).cpu().numpy()[0]
    ref_fb = fam.grad_log_post([th0], (X,), n, np.float64)[0]
    assert np.abs(fb - ref_fb).max() / np.abs(ref_fb).max() <= 1e-5
