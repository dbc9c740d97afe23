"""Oracle self-consistency: the checks the reference's own tests make (tests/test_samplers.py,
tests/test_gradient_estimation.py, examples/gauss_sample_sgld.py:36) plus analytic anchors."""
import json
import os

import numpy as np
import pytest

from oracle import jax_prng as P
from oracle import ksd as K
from oracle import models as M
from oracle import sgmcmc as S

REG = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_regression.json")))

ALL = [("sgld", 1e-5, {}), ("sgldCV", 1e-4, {"cv": True}), ("sgld_SVRG", 1e-4, {"update_rate": 5}),
       ("sghmc", 1e-6, {"L": 2}), ("sghmcCV", 1e-6, {"L": 2, "cv": True}), ("sghmc_SVRG", 1e-6, {"L": 2, "update_rate": 5}),
       ("psgld", 1e-2, {}), ("baoab", 1e-3, {"gamma": 5}), ("sgnht", 1e-5, {}), ("sgnhtCV", 1e-5, {"cv": True}),
       ("badodab", 1e-3, {}), ("badodabCV", 1e-3, {"cv": True}), ("sgldAdam", 1e-3, {})]


def toy():
    return P.normal(P.PRNGKey(0), (1000, 5))          # tests/models.py:28-30


@pytest.mark.parametrize("name,dt,kw", ALL, ids=[a[0] for a in ALL])
@pytest.mark.parametrize("two_leaves", [False, True])
def test_all_samplers_shapes_no_nan(name, dt, kw, two_leaves):
    """tests/test_samplers.py:95-117 restated for the oracle (12 samplers x 2 pytrees + sgldAdam)."""
    X = toy()
    fam = M.GaussianMean((0.5, 0.1), (0.001, 0.001)) if two_leaves else M.GaussianMean()
    ic = [np.zeros(5, np.float32)] * (2 if two_leaves else 1)
    kw = dict(kw)
    if kw.pop("cv", False):
        kw["centering_value"] = ic
    kern = S.build_kernel(name, fam, (X,), 10, dt, **kw)
    out, _, _ = S.run_sampler(kern, P.PRNGKey(0), 10, ic)
    assert len(out) == len(ic)
    for leaf in out:
        assert leaf.shape == (10, 5) and np.isfinite(leaf).all()


def test_fullbatch_standard_estimator_equals_grad_log_post():
    """tests/test_gradient_estimation.py:20-30."""
    X = toy()
    fam = M.GaussianMean()
    est = S.Estimator(fam, (X,), X.shape[0], "standard")
    th = [np.ones(5, np.float32)]
    g, svrg = est.estimate_gradient(0, P.PRNGKey(1), th, None)
    np.testing.assert_array_equal(g[0], fam.grad_log_post(th, (X,), 1000)[0])
    assert svrg is None


def test_svrg_state_and_refresh():
    """tests/test_gradient_estimation.py:60-91: centring value equals params at init / refresh."""
    X = toy()
    est = S.Estimator(M.GaussianMean(), (X,), 10, "svrg", update_rate=5)
    th = [np.full(5, 0.3, np.float32)]
    g, st = est.init_gradient(P.PRNGKey(0), th)
    np.testing.assert_array_equal(st[0][0], th[0])
    np.testing.assert_array_equal(g[0], st[1][0])
    th2 = [np.full(5, 0.5, np.float32)]
    _, st2 = est.estimate_gradient(3, P.PRNGKey(1), th2, st)
    np.testing.assert_array_equal(st2[0][0], th[0])          # i % 5 != 0: unchanged
    _, st3 = est.estimate_gradient(5, P.PRNGKey(1), th2, st)
    np.testing.assert_array_equal(st3[0][0], th2[0])         # refreshed at current params


def test_cv_estimator_is_exact_for_gaussian():
    X = toy()
    fam = M.GaussianMean()
    c = [np.full(5, 0.1, np.float32)]
    est = S.Estimator(fam, (X,), 10, "cv", centering_value=c)
    th = [np.full(5, -0.2, np.float32)]
    g, _ = est.estimate_gradient(0, P.PRNGKey(3), th, None)
    np.testing.assert_allclose(g[0], fam.grad_log_post(th, (X,), 1000)[0], rtol=2e-5, atol=2e-3)


def test_logreg_gradient_matches_finite_differences():
    ks = P.split(P.PRNGKey(2), 3)
    X = P.normal(ks[0], (64, 6)).astype(np.float64)
    y = (P.uniform(ks[1], (64,)) < 0.5).astype(np.int32)
    th = P.normal(ks[2], (6,)).astype(np.float64)
    fam = M.LogisticRegression()

    def logpost(t):
        s = (1.0 - 2.0 * y) * (X @ t)
        return -0.05 * t @ t + 64 * np.mean(-np.logaddexp(0.0, s))

    g = fam.grad_log_post([th], (X, y), 64, np.float64)[0]
    fd = np.array([(logpost(th + 1e-6 * e) - logpost(th - 1e-6 * e)) / 2e-6 for e in np.eye(6)])
    np.testing.assert_allclose(g, fd, rtol=1e-6, atol=1e-6)


def test_mlp_and_pmf_gradients_match_finite_differences():
    rng = np.random.default_rng(0)
    sizes = [7, 5, 3]
    leaves = []
    for m, n in zip(sizes[:-1], sizes[1:]):
        leaves += [rng.normal(size=(n, m)) * 0.5, rng.normal(size=(n,)) * 0.5]
    X = rng.uniform(size=(9, 7))
    Y = np.eye(3)[rng.integers(0, 3, 9)]
    fam = M.BayesianMLP()

    def lsm(z):
        z = z - z.max()
        return z - np.log(np.exp(z).sum())

    def logpost(ls):
        lp = sum(-0.5 * (l * l).sum() for l in ls)
        ll = 0.0
        for x, yy in zip(X, Y):
            a = np.exp(lsm(ls[0] @ x + ls[1]))
            ll += (yy * lsm(ls[2] @ a + ls[3])).sum()
        return lp + 20 * ll / 9

    g = fam.grad_log_post(leaves, (X, Y), 20, np.float64)
    for li, l in enumerate(leaves):
        it = np.nditer(l, flags=["multi_index"])
        for _ in it:
            idx = it.multi_index
            lp_, lm_ = [a.copy() for a in leaves], [a.copy() for a in leaves]
            lp_[li][idx] += 1e-6
            lm_[li][idx] -= 1e-6
            assert abs((logpost(lp_) - logpost(lm_)) / 2e-6 - g[li][idx]) < 1e-5
    # PMF
    U, V = rng.normal(size=(4, 3)), rng.normal(size=(5, 3))
    ij = np.stack([rng.integers(0, 4, 11), rng.integers(0, 5, 11)], 1).astype(np.int32)
    r = rng.integers(1, 6, 11).astype(np.float64)
    pmf = M.PMF()

    def lp_pmf(u, v):
        e = r - (u[ij[:, 0]] * v[ij[:, 1]]).sum(1)
        return -0.5 * ((u * u).sum() + (v * v).sum()) + 30 * np.mean(-0.5 * 2.0 * e * e)

    gu, gv = pmf.grad_log_post([U, V], (ij, r), 30, np.float64)
    for a in range(4):
        for b in range(3):
            up, um = U.copy(), U.copy()
            up[a, b] += 1e-6
            um[a, b] -= 1e-6
            assert abs((lp_pmf(up, V) - lp_pmf(um, V)) / 2e-6 - gu[a, b]) < 1e-5


def test_readme_gaussian_stationary_moments():
    """examples/gauss_sample_sgld.py:36 (mean within 0.1) + SURVEY.md Appendix B variance."""
    X = P.normal(P.PRNGKey(0), (10000, 100))
    kern = S.build_kernel("sgld", M.GaussianMean(), (X,), 1000, 1e-5)
    out, _, _ = S.run_sampler(kern, P.PRNGKey(0), 1500, [np.zeros(100, np.float32)])
    np.testing.assert_allclose(out[0][:3, :3], REG["readme_sgld_first_samples_head"], rtol=1e-4, atol=1e-6)
    post_mean = X.sum(0) / (10000 + 0.01)
    assert np.abs(out[0][300:].mean(0) - post_mean).max() < 0.01
    var = out[0][300:].var(0).mean()
    assert abs(var - 1.58e-4) / 1.58e-4 < 0.15


def test_badodabcv_reference_bug_never_moves_x():
    X = toy()
    ic = [np.full(5, 0.25, np.float32)]
    kern = S.build_kernel("badodabCV", M.GaussianMean(), (X,), 10, 1e-3, centering_value=ic, strict_reference_bugs=True)
    out, _, _ = S.run_sampler(kern, P.PRNGKey(0), 5, ic)
    np.testing.assert_array_equal(out[0], np.broadcast_to(ic[0], (5, 5)))


def test_ksd_rows_match_pair_formula_and_regression():
    rng = np.random.default_rng(1)
    X, G = rng.normal(size=(12, 4)), rng.normal(size=(12, 4))
    tot = sum(K.k0_pair(X[i], X[j], G[i], G[j]) for i in range(12) for j in range(12))
    assert abs(K.imq_ksd_sum(X, G) - tot) < 1e-9 * abs(tot)
    assert abs(K.imq_ksd_sum(X, G, row0=0, row1=5) + K.imq_ksd_sum(X, G, row0=5, row1=12) - tot) < 1e-9 * abs(tot)
    assert REG["logreg_small"]["ksd_fp64"] > 0
