"""Pins the oracle's PRNG against external known answers (tests/golden/jax_prng_kat.json)."""
import json
import os

import numpy as np

from oracle import jax_prng as P

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "jax_prng_kat.json")))
key = lambda k: np.array(k, dtype=np.uint32)


def test_threefry_random123_vectors():
    for v in G["external"]["threefry2x32"]:
        w0, w1 = P.threefry2x32_block(v["key"][0], v["key"][1], np.array([v["ctr"][0]]), np.array([v["ctr"][1]]))
        assert [int(w0[0]), int(w1[0])] == v["out"]


def test_split_external_and_frozen():
    for sec in ("external", "self_derived"):
        for v in G[sec]["split"]:
            assert P.split(key(v["key"]), v["num"]).tolist() == v["out"]


def test_uniform_normal_external():
    for v in G["external"]["uniform"]:
        np.testing.assert_allclose(np.ravel(P.uniform(key(v["key"]), tuple(v["shape"]))), v["out"], rtol=2e-7)
    for v in G["external"]["normal"]:
        np.testing.assert_allclose(np.ravel(P.normal(key(v["key"]), tuple(v["shape"]))), v["out"], rtol=3e-7)


def test_choice_frozen_vectors():
    # PARITY UNPINNED: self-derived regression anchors for randint / choice
    for v in G["self_derived"]["choice"]:
        idx = P.choice_indices(key(v["key"]), v["n"], v["batch"])
        assert idx.dtype == np.int32 and idx.shape == (v["batch"],)
        assert idx[: len(v["head"])].tolist() == v["head"] and int(idx.sum()) == v["sum"]
        assert idx.min() >= 0 and idx.max() < v["n"]


def test_readme_key_tree():
    r = G["self_derived"]["readme_keys"]
    k0 = P.PRNGKey(0)
    carry, sub_init = P.split(k0)
    assert sub_init.tolist() == r["init_subkey"]
    carry, sub0 = P.split(carry)
    assert sub0.tolist() == r["step0_subkey"]
    k1, k2 = P.split(sub0)
    assert k1.tolist() == r["k1"] and k2.tolist() == r["k2"]
    assert P.split(k1, 1)[0].tolist() == r["leaf0_noise_key"]


def test_half_split_layout_odd_length_pads_zero_counter():
    k = P.PRNGKey(5)
    bits = P.random_bits(k, (7,))
    h = 4
    w0, w1 = P.threefry2x32_block(k[0], k[1], np.arange(4), np.array([4, 5, 6, 0]))
    assert bits.tolist() == np.concatenate([w0, w1])[:7].tolist() and h == 4


def test_randint_multiplier_cases():
    assert P.randint_multiplier(10_000) == 7296
    assert P.randint_multiplier(60_000) == 47296
    assert P.randint_multiplier(100_000) == 0 and P.randint_multiplier(1_000_000) == 0
    k = P.PRNGKey(9)
    lo = P.random_bits(P.split(k)[1], (33,))
    assert P.randint(k, (33,), 0, 10**6).tolist() == (lo % np.uint32(10**6)).astype(np.int32).tolist()


def test_erfinv_matches_true_erfinv_to_polynomial_accuracy():
    from scipy.special import erfinv

    x = np.linspace(-0.999, 0.999, 2001).astype(np.float32)
    np.testing.assert_allclose(P.erf_inv_f32(x), erfinv(x.astype(np.float64)), rtol=1e-5, atol=1e-7)
    assert np.isinf(P.erf_inv_f32(np.float32(1.0))) and np.isinf(P.erf_inv_f32(np.float32(-1.0)))


def test_gen_data_restatement_key_tree_and_structure():
    """oracle/gen_data.py restates models/logistic_regression.py:12-62: one uniform decides the whole covariance
    (the reference reuses the key), labels come from per-row keys split(key, N)."""
    from oracle import gen_data as G

    key = P.PRNGKey(42)
    ks = P.split(key, 4)
    theta, X, y = G.gen_data(key, 6, 400)
    np.testing.assert_array_equal(theta, P.normal(ks[1], (6,)) * np.float32(np.sqrt(np.float32(10.0))))
    cov = G.gen_cov_mat(ks[2], 6, 0.4)
    u = P.uniform(ks[2], ())
    base = np.float32(np.float32(np.float32(u * np.float32(2.0)) * np.float32(0.4)) - np.float32(0.4))
    assert abs(base) < 0.4 and np.allclose(np.diag(cov), 1.0)
    for k in range(1, 6):
        np.testing.assert_allclose(np.diag(cov, k), float(base) ** k, rtol=1e-6)
    assert np.array_equal(cov, cov.T)
    np.testing.assert_allclose(X, P.normal(ks[3], (400, 6)) @ np.linalg.cholesky(cov.astype(np.float64)).astype(np.float32), rtol=1e-5, atol=1e-6)
    # label i: uniform(split(key', N)[i], ()) < sigmoid(theta . x_i)
    keys = P.split(ks[0], 400)
    p = 1.0 / (1.0 + np.exp(-(X @ theta).astype(np.float64)))
    for i in (0, 1, 57, 399):
        assert y[i] == int(P.uniform(keys[i], ()) < p[i])
    assert y.dtype == np.int32 and 0.2 < y.mean() < 0.8
