"""Host-side logic that runs without a GPU: pytrees, schedules / step coefficients, registry, sharding."""
import numpy as np
import pytest

from sgmcmcjax_b200 import _tree, diffusions, ksd, models
from sgmcmcjax_b200.models import resolve_family


def test_tree_roundtrip_leaf_order():
    tree = {"b": [1, (2, 3)], "a": 0}
    leaves, td = _tree.flatten(tree)
    assert leaves == [0, 1, 2, 3]                          # dict keys sorted, sequences in order
    assert _tree.unflatten(td, [10, 11, 12, 13]) == {"a": 10, "b": [11, (12, 13)]}
    leaves, td = _tree.flatten(5)
    assert leaves == [5] and _tree.unflatten(td, [7]) == 7


def test_schedules_match_reference_formulas():
    s = diffusions.welling_teh_schedule(1e-3, 10.0)
    assert s(5) == pytest.approx(1e-3 * 15.0 ** -0.55)
    c = diffusions.cyclical_schedule(1e-2, 4, 100)          # period ceil(100/4) = 25
    assert c(1) == pytest.approx(1e-2) and c(26) == pytest.approx(1e-2)
    assert diffusions.make_schedule(0.5)(123) == 0.5
    with pytest.raises(TypeError):
        diffusions.make_schedule(np.zeros(3))


def test_step_coefs_weak_type_rounding():
    f = np.float32
    sch = diffusions.make_schedule(1e-5)
    row = diffusions.step_coefs("sgld", {}, sch, 0, 1)[0]
    assert row[0] == f(1e-5) and row[1] == f(0.5e-5) and row[2] == np.sqrt(f(2e-5))
    rows = diffusions.step_coefs("sgldAdam", dict(beta1=0.9, beta2=0.999, eps=1e-8), sch, 3, 2)
    assert rows.shape == (2, 4) and rows[0, 2] == f(1) - np.power(f(0.9), f(4))
    b = diffusions.step_coefs("baoab", dict(gamma=5.0, tau=1.0), diffusions.make_schedule(1e-3), 0, 1)[0]
    assert b[2] == f(np.exp(np.float64(f(-5e-3)))) and b[3] == np.sqrt(f(1) - b[2] * b[2])
    assert diffusions.coef_is_constant("sgld", sch) and not diffusions.coef_is_constant("sgldAdam", sch)
    assert not diffusions.coef_is_constant("sgld", diffusions.welling_teh_schedule(1.0, 1.0))


def test_family_registry():
    spec = resolve_family(models.logistic_regression.loglikelihood, models.logistic_regression.logprior)
    assert spec.name == "logistic_regression" and spec.prior_coef == (0.1,)
    g = resolve_family(models.gaussian_mean.loglikelihood_list_array, models.gaussian_mean.logprior_list_array)
    assert g.lik_coef == (0.5, 0.1) and g.prior_coef == (0.001, 0.001)
    with pytest.raises(NotImplementedError):
        resolve_family(lambda th, x: 0.0, lambda th: 0.0)
    with pytest.raises(ValueError):
        resolve_family(models.gaussian_mean.loglikelihood, models.logistic_regression.logprior)
    with pytest.raises(ValueError):
        resolve_family(models.gaussian_mean.logprior, models.gaussian_mean.loglikelihood)


def test_ksd_row_blocks_partition_rows():
    for n, w in [(50000, 8), (10, 3), (7, 8), (1, 1)]:
        blocks = [ksd.row_block(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(b[1] == nb[0] for b, nb in zip(blocks, blocks[1:]))


def test_no_cpu_fallback():
    import torch

    from sgmcmcjax_b200 import _native

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_native.NativeError):
        from sgmcmcjax_b200.samplers import build_sgld_sampler

        build_sgld_sampler(1e-5, models.gaussian_mean.loglikelihood, models.gaussian_mean.logprior,
                           (np.zeros((10, 4), np.float32),), 2)


def test_rhat_from_moments_matches_textbook_formula():
    """sgmcmcjax_b200/diagnostics.py: Gelman-Rubin R-hat from per-chain moments == the formula on the stored draws."""
    from sgmcmcjax_b200 import diagnostics as D

    rng = np.random.default_rng(0)
    draws = rng.normal(size=(6, 200, 3)) + np.array([0.0, 0.0, 1.0])[None, None, :] * np.arange(6)[:, None, None]
    n = draws.shape[1]
    w = draws.var(axis=1, ddof=1).mean(axis=0)
    b = n * draws.mean(axis=1).var(axis=0, ddof=1)
    want = np.sqrt(((n - 1) / n * w + b / n) / w)
    got = D.rhat_from_moments(draws.mean(axis=1), draws.var(axis=1), n)
    np.testing.assert_allclose(got, want, rtol=1e-12)
    np.testing.assert_allclose(D.rhat_from_samples(draws), want, rtol=1e-12)
    assert got[0] < 1.05 and got[2] > 2.0                     # mixed / separated chains
    summ = {"mean": [draws.mean(axis=1), draws.mean(axis=1)[:, :1]], "var": [draws.var(axis=1), draws.var(axis=1)[:, :1]], "count": n}
    out = D.rhat(summ)
    assert isinstance(out, list) and out[0].shape == (3,) and out[1].shape == (1,)
    with pytest.raises(ValueError):
        D.rhat_from_moments(draws.mean(axis=1)[:1], draws.var(axis=1)[:1], n)


def test_sampler_extension_keywords_are_validated_on_the_host():
    """thin / summaries / store_samples (samplers.py): argument errors are raised before any device work."""
    from sgmcmcjax_b200 import samplers as NS

    with pytest.raises(ValueError, match="thin must be >= 1"):
        NS._summaries_sampler(object(), 0, False, True, False)
    s = NS._summaries_sampler(object(), 3, True, False, False)       # construction needs no device
    assert callable(s)


def test_ess_of_an_ar1_process():
    """diagnostics.ess_from_samples: AR(1) with coefficient rho has ESS / n = (1 - rho) / (1 + rho)."""
    from sgmcmcjax_b200 import diagnostics as D

    rng = np.random.default_rng(3)
    c, k = 8, 4000
    for rho in (0.0, 0.5, 0.9):
        x = np.zeros((c, k, 2))
        e = rng.normal(size=(c, k, 2))
        x[:, 0] = e[:, 0]
        for t in range(1, k):
            x[:, t] = rho * x[:, t - 1] + np.sqrt(1 - rho * rho) * e[:, t]
        ess = D.ess_from_samples(x)
        want = c * k * (1 - rho) / (1 + rho)
        assert ess.shape == (2,)
        assert np.all(np.abs(ess / want - 1.0) < 0.2), (rho, ess, want)


def test_dataset_loaders_read_the_standard_formats(tmp_path):
    """sgmcmcjax_b200.datasets: MNIST IDX files -> the arrays NN_data.py:26-42 builds; MovieLens
    u.data -> 0-based PMF data."""
    import gzip
    import struct

    from sgmcmcjax_b200 import datasets as D

    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, size=(7, 28, 28), dtype=np.uint8)
    lab = rng.integers(0, 10, size=7, dtype=np.uint8)
    pi, pl = tmp_path / "img-idx3-ubyte.gz", tmp_path / "lab-idx1-ubyte"
    with gzip.open(pi, "wb") as f:
        f.write(struct.pack(">HBBIII", 0, 0x08, 3, 7, 28, 28) + img.tobytes())
    with open(pl, "wb") as f:
        f.write(struct.pack(">HBBI", 0, 0x08, 1, 7) + lab.tobytes())
    X, y = D.load_mnist_idx(str(pi), str(pl))
    assert X.dtype == np.float32 and X.shape == (7, 784) and y.shape == (7, 10) and y.dtype == np.float32
    np.testing.assert_array_equal(X, img.reshape(7, -1).astype(np.float32) / np.float32(255.0))
    assert (y.argmax(1) == lab).all() and (y.sum(1) == 1).all()
    with pytest.raises(ValueError):
        D.one_hot([3, 10], 10)

    pu = tmp_path / "u.data"
    pu.write_text("1\t5\t3\t881250949\n3\t2\t5\t891717742\n2\t5\t1\t878887116\n")
    (ij, r), (nu, ni) = D.load_movielens_100k(str(pu))
    assert ij.dtype == np.int32 and ij.tolist() == [[0, 4], [2, 1], [1, 4]] and r.tolist() == [3.0, 5.0, 1.0]
    assert (nu, ni) == (3, 5)
    pu.write_text("0\t5\t3\t1\n")
    with pytest.raises(ValueError):
        D.load_movielens_100k(str(pu))
