"""GPU parity of the cluster-per-chain launch shape (csrc/vec_sampler.cuh header): the minibatch of ONE chain split
over the CTAs of a thread-block cluster, partial sums exchanged over DSMEM.

This is the shape config 2 runs in when its 1024 chains are sharded over 8 GPUs (128 chains per GPU < 148 SMs) and the
README's single chain (samplers.py:45-58 has no multi-chain API at all).  Same bars as tests/test_gpu_parity.py:
single step <= 1e-5 of max|ref| against the oracle, for all 13 kernels.
"""
import numpy as np
import pytest

from oracle import jax_prng as P
from oracle import sgmcmc as S

from test_gpu_parity import (KERNELS, as_params, build_pair, build_sampler_pair, close, compare_state, gauss_case,
                             grad_tol, inject, leaves_of, logreg_case, to_np)

pytestmark = pytest.mark.gpu

SHAPES = [(2, 16), (3, 4), (8, 8), (16, 2)]          # (CTAs per chain, warps per CTA)


def force_shape(monkeypatch, cluster, warps):
    monkeypatch.setenv("SGMCMC_CLUSTER", str(cluster))
    monkeypatch.setenv("SGMCMC_WARPS_PER_CHAIN", str(warps))


@pytest.mark.parametrize("name,dt,kw", KERNELS, ids=[k[0] for k in KERNELS])
@pytest.mark.parametrize("cluster,warps", SHAPES)
def test_cluster_single_step_from_oracle_state_1e5(cuda, monkeypatch, name, dt, kw, cluster, warps):
    """All 13 kernels, logistic regression D=100 (kernels.py:53-64,110-122; gradient_estimation.py:31-35,74-86,134-151):
    init_fn and step 0 with the chain split over a cluster, vs the oracle."""
    force_shape(monkeypatch, cluster, warps)
    fam, fns, data, th0 = logreg_case(n=5000, d=100, seed=9)
    center = [np.asarray(th0[0] * np.float32(0.99))]
    ok, (init_fn, kernel, get_params) = build_pair(name, dt, kw, fam, fns, data, 500, center)
    sub_i, sub_0 = P.split(P.PRNGKey(77))
    ost = ok.init_fn(sub_i, th0)
    nst = init_fn(sub_i, as_params(th0))
    g_ref, g_got = compare_state(ost, nst, get_params, f"{name} init")
    assert np.abs(g_got - g_ref).max() / np.abs(g_ref).max() <= 3e-5
    ost = ok.kernel(0, sub_0, ost)
    nst = kernel(0, sub_0, nst)
    g_ref, g_got = compare_state(ost, nst, get_params, f"{name} step0")
    assert np.abs(g_got - g_ref).max() / np.abs(g_ref).max() <= 3e-5


@pytest.mark.parametrize("name,dt,kw", KERNELS, ids=[k[0] for k in KERNELS])
@pytest.mark.parametrize("case", ["gauss2", "logreg"])
def test_cluster_injected_steps_match_oracle(cuda, monkeypatch, name, dt, kw, case):
    """4 strict single steps (state injected from the oracle before each), cluster of 4 CTAs x 4 warps; the 2-leaf
    Gaussian exercises the per-leaf keys and thermostats in every replica."""
    force_shape(monkeypatch, 4, 4)
    if case == "logreg":
        fam, fns, data, th0 = logreg_case()
        batch = 100
    else:
        fam, fns, data, th0 = gauss_case(True)
        batch = 10
    center = [np.asarray(t + np.float32(0.01)) for t in th0]
    ok, (init_fn, kernel, get_params) = build_pair(name, dt, kw, fam, fns, data, batch, center)
    key = P.PRNGKey(0)
    key, sub = P.split(key)
    ost = ok.init_fn(sub, th0)
    nst = init_fn(sub, as_params(th0))
    g_ref, g_got = compare_state(ost, nst, get_params, f"{name}/{case} init")
    assert np.abs(g_got - g_ref).max() / max(np.abs(g_ref).max(), 1.0) <= grad_tol(name)
    for i in range(4):
        key, sub = P.split(key)
        nst = kernel(i, sub, inject(ost, nst))
        ost = ok.kernel(i, sub, ost)
        g_ref, g_got = compare_state(ost, nst, get_params, f"{name}/{case} step {i}")
        assert np.abs(g_got - g_ref).max() / max(np.abs(g_ref).max(), 1.0) <= grad_tol(name)
        if ost.svrg_state is not None:
            packed = nst.diffusion_state.packed_state
            close(to_np(packed["svrg_center"])[0], np.concatenate(ost.svrg_state[0]), what="svrg centre")
            close(to_np(packed["svrg_grad"])[0], np.concatenate(ost.svrg_state[1]), rtol=2e-5, what="svrg fb grad")


@pytest.mark.parametrize("cluster,warps", SHAPES + [(5, 1), (7, 32)])
def test_cluster_sampler_many_chains_and_steps(cuda, monkeypatch, cluster, warps):
    """samplers.py:45-58 scan semantics, several chains per launch, 20 steps in ONE launch (20 cluster folds through the
    double-buffered exchange): every chain against its own oracle run, and against the one-CTA-per-chain shape."""
    from sgmcmcjax_b200 import samplers as NS

    fam, fns, data, th0 = logreg_case(n=3000, d=100, seed=4)
    keys = P.split(P.PRNGKey(1), 6)
    th = np.stack([th0[0] + np.float32(0.01 * c) for c in range(6)])
    force_shape(monkeypatch, 1, 8)
    base = NS.build_sgldCV_sampler(1e-4, fns[0], fns[1], data, 300, th0[0], pbar=False)(keys, 20, th)
    force_shape(monkeypatch, cluster, warps)
    sampler = NS.build_sgldCV_sampler(1e-4, fns[0], fns[1], data, 300, th0[0], pbar=False)
    got = sampler(keys, 20, th)
    assert got.shape == (6, 20, 100) and np.isfinite(got).all()
    close(got, base, rtol=2e-5, what="cluster vs single CTA")
    ok = S.build_kernel("sgldCV", fam, data, 300, 1e-4, centering_value=th0)
    for c in (0, 5):
        ref, _, _ = S.run_sampler(ok, keys[c], 20, [th[c]])
        close(got[c], ref[0], rtol=1e-4, what=f"chain {c} vs oracle")
    # the replicas of a cluster are bit-identical, so a launch is reproducible bit for bit
    assert np.array_equal(got, sampler(keys, 20, th))


@pytest.mark.parametrize("batch", [1, 2, 63, 65, 129, 1000])
def test_cluster_ragged_batches_and_full_batch(cuda, monkeypatch, batch):
    """Batches smaller than the cluster's warp count (most CTAs idle), odd batches, and batch == N (the static
    full-batch bypass, gradient_estimation.py:41-42, as a sequential pass split over the cluster)."""
    force_shape(monkeypatch, 8, 4)
    fam, fns, data, th0 = gauss_case()
    ok, sampler = build_sampler_pair("sgld", 1e-5, {}, fam, fns, data, batch, th0)
    ref, _, _ = S.run_sampler(ok, P.PRNGKey(8), 4, th0)
    close(sampler(P.PRNGKey(8), 4, th0[0]), ref[0], rtol=5e-5, what=f"batch={batch}")


def test_cluster_thinning_summaries_and_optimizer(cuda, monkeypatch):
    """The extended instantiation (thinning + on-device moments; the Adam MAP optimiser with its log-posterior values,
    optimizer.py:41-60) in cluster shape: rank 0 is the only writer, the value term crosses the cluster with the sums."""
    from sgmcmcjax_b200 import optimizer as NO
    from sgmcmcjax_b200 import samplers as NS

    from oracle import models as M
    from oracle import optimizer as OO

    fam, fns, data, th0 = logreg_case(n=3000, d=20, seed=6)
    force_shape(monkeypatch, 1, 8)
    full = NS.build_sgld_sampler(1e-4, fns[0], fns[1], data, 300, pbar=False)(P.PRNGKey(5), 40, th0[0])
    p1, lp1 = NO.build_optax_optimizer(NO.adam(1e-2), fns[0], fns[1], data, 300)(P.PRNGKey(3), 30, th0[0])
    force_shape(monkeypatch, 4, 4)
    kept, summ = NS.build_sgld_sampler(1e-4, fns[0], fns[1], data, 300, pbar=False, thin=4, summaries=True)(
        P.PRNGKey(5), 40, th0[0])
    close(kept, full[3::4], rtol=2e-5, what="thinned cluster run vs unthinned single CTA")
    close(summ["mean"], kept.astype(np.float64).mean(0), rtol=1e-6, what="on-device mean")
    p4, lp4 = NO.build_optax_optimizer(NO.adam(1e-2), fns[0], fns[1], data, 300)(P.PRNGKey(3), 30, th0[0])
    close(p4, p1, rtol=2e-5, what="optimiser params, cluster vs single CTA")
    close(lp4, lp1, rtol=2e-5, what="optimiser log-posterior values, cluster vs single CTA")
    po, lpo = OO.run_adam(M.LogisticRegression(), data, 300, 1e-2, P.PRNGKey(3), 30, th0)
    close(p4, po[0], rtol=5e-5, what="optimiser params vs oracle")
    close(lp4, lpo, rtol=5e-5, what="optimiser values vs oracle")


def test_default_shape_uses_a_cluster_for_few_chains(cuda, monkeypatch):
    """Without overrides: 1 chain at the README geometry (README.md:29-40) and 8 chains at a C2-like geometry run as
    clusters (launch-shape query), and still match the oracle."""
    import ctypes as C

    from sgmcmcjax_b200 import _native
    from sgmcmcjax_b200 import samplers as NS
    from sgmcmcjax_b200.models import gaussian_mean as G

    monkeypatch.delenv("SGMCMC_CLUSTER", raising=False)
    monkeypatch.delenv("SGMCMC_WARPS_PER_CHAIN", raising=False)
    X = P.normal(P.PRNGKey(0), (10000, 100))
    sampler = NS.build_sgld_sampler(1e-5, G.loglikelihood, G.logprior, (X,), 1000, pbar=False)
    out = sampler(P.PRNGKey(0), 50, np.zeros(100, np.float32))
    lib = _native.load()
    w, cl = C.c_int32(0), C.c_int32(0)
    _native.check(lib.sgmcmc_launch_shape(sampler.engine._handle, 1, C.byref(w), C.byref(cl)))
    assert cl.value > 1, (w.value, cl.value)
    from oracle import models as M
    ok = S.build_kernel("sgld", M.GaussianMean(), (X,), 1000, 1e-5)
    ref, _, _ = S.run_sampler(ok, P.PRNGKey(0), 50, [np.zeros(100, np.float32)])
    close(out, ref[0], rtol=1e-4, what="README chain in cluster shape vs oracle")


@pytest.mark.parametrize("name,dt,kw", KERNELS, ids=[k[0] for k in KERNELS])
@pytest.mark.parametrize("case", ["gauss2", "logreg"])
def test_cluster_multi_step_launch_all_kernels(cuda, monkeypatch, name, dt, kw, case):
    """K steps in ONE cluster launch (K cluster folds through the double-buffered exchange, the carry key split one
    iteration ahead): every kernel's 12-step trajectory against the oracle's scan (samplers.py:45-58)."""
    if case == "logreg":
        fam, fns, data, th0 = logreg_case(n=3000, d=100, seed=4)
        batch = 300
    else:
        fam, fns, data, th0 = gauss_case(True)
        batch = 64
    center = [np.asarray(t + np.float32(0.01)) for t in th0]
    force_shape(monkeypatch, 4, 4)
    ok, sampler = build_sampler_pair(name, dt, kw, fam, fns, data, batch, center)
    got = leaves_of(sampler(P.PRNGKey(2), 12, as_params(th0)))
    ref, _, _ = S.run_sampler(ok, P.PRNGKey(2), 12, th0)
    for g, r in zip(got, ref):
        close(g, r, rtol=2e-4, what=f"{name}/{case} cluster trajectory vs oracle")
