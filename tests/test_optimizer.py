"""MAP optimiser (sgmcmcjax/optimizer.py:12-62 with optax.adam): oracle on the CPU, fused CUDA path on the GPU."""
import numpy as np
import pytest

from oracle import jax_prng as P
from oracle import models as M
from oracle import optimizer as O


def reference_test_problem():
    """tests/test_optimizers.py:8-20."""
    key = P.PRNGKey(0)
    mu_true = P.normal(key, (100,))
    X = (P.normal(key, (10_000, 100)) + mu_true).astype(np.float32)
    return key, mu_true, X


def test_oracle_adam_reaches_the_posterior_mode():
    """The reference's own (statistical) bar, tests/test_optimizers.py:23-36, on a shorter run."""
    key, mu_true, X = reference_test_problem()
    leaves, lps = O.run_adam(M.GaussianMean((0.5,), (0.005,)), (X,), 1000, 1e-2, key, 1500, [np.zeros(100, np.float32)])
    assert np.allclose(leaves[0], mu_true, atol=1e-1)
    assert lps.shape == (1500,) and lps[-1] > lps[0]


def test_optimizer_api_errors():
    from sgmcmcjax_b200 import optimizer as NO
    from sgmcmcjax_b200.models import gaussian_mean as G

    with pytest.raises(TypeError):
        NO.build_optax_optimizer("adam", G.loglikelihood, G.logprior, (np.zeros((10, 3), np.float32),), 2)

    class FakeOptax:
        def init(self, p): ...
        def update(self, g, s): ...

    with pytest.raises(NotImplementedError):
        NO.build_optax_optimizer(FakeOptax(), G.loglikelihood, G.logprior, (np.zeros((10, 3), np.float32),), 2)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["gauss", "gauss2", "logreg"])
def test_adam_iterations_match_oracle(cuda, case):
    from sgmcmcjax_b200 import optimizer as NO
    from test_gpu_parity import as_params, close, gauss_case, leaves_of, logreg_case

    if case == "logreg":
        fam, fns, data, th0 = logreg_case(n=3000, d=20)
        batch, lr = 300, 1e-2
    else:
        fam, fns, data, th0 = gauss_case(case == "gauss2", n=2000, d=7)
        batch, lr = 100, 1e-2
    opt = NO.build_optax_optimizer(NO.adam(lr), fns[0], fns[1], data, batch)
    key = P.PRNGKey(4)
    params, lps = opt(key, 12, as_params(th0))
    ref, ref_lps = O.run_adam(fam, data, batch, lr, key, 12, th0)
    for got, r in zip(leaves_of(params), ref):
        close(got, r, rtol=2e-5, what=f"adam params {case}")
    np.testing.assert_allclose(lps, ref_lps, rtol=2e-5)
    assert lps.shape == (12,)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["gauss2", "logreg"])
def test_adam_full_batch_uses_all_rows(cuda, case):
    """batch_size == N: the estimator's static full-batch bypass (gradient_estimation.py:41-42) under the optimiser
    (optimizer.py:41-60) - every row once, no index draw, scale 1."""
    from sgmcmcjax_b200 import optimizer as NO
    from test_gpu_parity import as_params, close, gauss_case, leaves_of, logreg_case

    fam, fns, data, th0 = logreg_case(n=1500, d=20) if case == "logreg" else gauss_case(True, n=1000, d=7)
    n = data[0].shape[0]
    opt = NO.build_optax_optimizer(NO.adam(1e-2), fns[0], fns[1], data, n)
    key = P.PRNGKey(9)
    params, lps = opt(key, 8, as_params(th0))
    ref, ref_lps = O.run_adam(fam, data, n, 1e-2, key, 8, th0)
    for got, r in zip(leaves_of(params), ref):
        close(got, r, rtol=2e-5, what=f"full-batch adam params {case}")
    np.testing.assert_allclose(lps, ref_lps, rtol=2e-5)


@pytest.mark.gpu
def test_reference_optimizer_test(cuda):
    """tests/test_optimizers.py:23-36 verbatim: Adam, lr 1e-2, batch 10 % of N = 10k, 10k iterations."""
    from sgmcmcjax_b200 import optimizer as NO
    from sgmcmcjax_b200.models import gaussian_mean as G

    key, mu_true, X = reference_test_problem()
    optimizer = NO.build_optax_optimizer(NO.adam(1e-2), G.loglikelihood, G.logprior, (X,), int(0.1 * 10_000))
    params, log_post_list = optimizer(key, 10_000, np.zeros(100, np.float32))
    assert log_post_list.shape == (10_000,) and params.shape == (100,)
    assert np.allclose(params, mu_true, atol=1e-1)
    # multi-chain: independent optimisations, chain c == the single run with keys[c]
    keys = P.split(key, 3)
    p3, lp3 = optimizer(keys, 50, np.zeros((3, 100), np.float32))
    p1, lp1 = optimizer(keys[1], 50, np.zeros(100, np.float32))
    assert np.array_equal(p3[1], p1) and np.array_equal(lp3[1], lp1)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["mlp", "pmf"])
def test_adam_iterations_match_oracle_big_families(cuda, case):
    """optimizer.py:41-60 on the Bayesian MLP and PMF (csrc/big_sampler.cuh adam_step): parameters and the log-posterior
    VALUES (util.py:43-44 with_val) against oracle/optimizer.py; pytree in, same pytree out."""
    from sgmcmcjax_b200 import optimizer as NO
    from test_gpu_big_families import CASES, close, flat, tree_leaves

    make, to_params, batch = CASES[case]
    fam, fns, data, th0 = make()
    opt = NO.build_optax_optimizer(NO.adam(1e-2), fns[0], fns[1], data, batch)
    key = P.PRNGKey(4)
    params, lps = opt(key, 10, to_params(th0))
    ref, ref_lps = O.run_adam(fam, data, batch, 1e-2, key, 10, th0)
    assert type(params) is type(to_params(th0))
    close(flat(tree_leaves(params)), flat(ref), rtol=3e-5, what=f"adam params {case}")
    np.testing.assert_allclose(np.asarray(lps), ref_lps, rtol=3e-5)
    # two chains at once == two single runs
    keys = P.split(P.PRNGKey(5), 2)
    multi = to_params([np.stack([t, t * np.float32(1.05)]) for t in th0])
    p2, lp2 = opt(keys, 6, multi)
    p1, lp1 = opt(keys[1], 6, to_params([t * np.float32(1.05) for t in th0]))
    close(flat([l[1] for l in tree_leaves(p2)]), flat(tree_leaves(p1)), rtol=1e-5, what="chain 1 of 2")
    np.testing.assert_allclose(np.asarray(lp2)[1], np.asarray(lp1), rtol=1e-5)


@pytest.mark.gpu
def test_map_centre_for_pmf_control_variates(cuda):
    """The use the optimiser exists for: a MAP estimate as the centring value of a *CV sampler.
    After 300 Adam iterations the log-posterior has risen and BADODAB-CV centred there runs finite."""
    from sgmcmcjax_b200 import optimizer as NO
    from sgmcmcjax_b200 import samplers as NS
    from test_gpu_big_families import pmf_case, pmf_params, tree_leaves

    fam, fns, data, th0 = pmf_case(60, 80, 8, 4000, seed=3)
    theta_map, lps = NO.build_optax_optimizer(NO.adam(5e-2), fns[0], fns[1], data, 400)(P.PRNGKey(0), 300, pmf_params(th0))
    lps = np.asarray(lps)
    assert lps[-20:].mean() > lps[:20].mean()
    sampler = NS.build_badodabCV_sampler(1e-3, fns[0], fns[1], data, 400, theta_map, pbar=False)
    out = sampler(P.PRNGKey(1), 20, theta_map)
    assert all(np.isfinite(l).all() for l in tree_leaves(out))
