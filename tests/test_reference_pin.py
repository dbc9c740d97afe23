"""Consumes tests/golden/reference_run.json - outputs of the UNMODIFIED reference on jax 0.4.13, written by
tests/golden/make_reference_fixtures.py wherever the reference can be imported.  The build image has no jax, so the
file is absent there and these tests SKIP, loudly: until it exists the oracle stays "parity unpinned" for randint and
the floating-point trajectories (oracle/__init__.py).

With the file present: the numpy oracle must reproduce the reference's PRNG bit for bit (split, bits behind normal /
uniform, and the minibatch indices of gradient_estimation.py:32), its kernel trajectories (kernels.py:53-64,110-122)
within the stated tolerances, and imq_KSD (ksd.py:43-66); on a GPU the CUDA path is held to the same fixture.
"""
import json
import os

import numpy as np
import pytest

from oracle import jax_prng as P
from oracle import ksd as OK
from oracle import models as M
from oracle import sgmcmc as S

FIX = os.path.join(os.path.dirname(__file__), "golden", "reference_run.json")
needs_fixture = pytest.mark.skipif(
    not os.path.exists(FIX),
    reason="PARITY UNPINNED: tests/golden/reference_run.json is absent (the reference needs jax==0.4.13, not installable "
           "in the build image); run tests/golden/make_reference_fixtures.py where it imports")

SPECS = {  # name -> (oracle kwargs builder)
    "sgld": {}, "sgldCV": {"cv": True}, "sgld_SVRG": {"update_rate": 3}, "psgld": {}, "sgldAdam": {}, "sgnht": {},
    "sgnhtCV": {"cv": True}, "baoab": {"gamma": 5.0}, "badodab": {}, "badodabCV": {"cv": True, "strict_reference_bugs": True},
    "sghmc": {"L": 3}, "sghmcCV": {"L": 3, "cv": True}, "sghmc_SVRG": {"L": 3, "update_rate": 3},
}


def fixture():
    return json.load(open(FIX))


def rebuild_cases(fx):
    inp = fx["inputs"]
    X = P.normal(np.array(inp["gauss_X_key"], np.uint32), (1000, 5))
    ks = P.split(np.array(inp["logreg_key"], np.uint32), 4)
    n, d = inp["logreg_n"], inp["logreg_d"]
    X_lr = P.normal(ks[1], (n, d))
    y = np.array(inp["logreg_y"], np.int32)
    th0 = np.array(inp["logreg_th0"], np.float32)
    return {
        "gauss1": (M.GaussianMean(), (X,), 10, [np.full(5, 0.05, np.float32)]),
        "gauss2": (M.GaussianMean((0.5, 0.1), (0.001, 0.001)), (X,), 10, [np.zeros(5, np.float32), np.full(5, 0.1, np.float32)]),
        "logreg": (M.LogisticRegression(), (X_lr, y), 100, [th0]),
    }


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@needs_fixture
def test_reference_run_used_the_legacy_threefry():
    fx = fixture()
    assert fx["threefry_partitionable"] is False, "the reference semantics are jax 0.4.13's non-partitionable threefry"


@needs_fixture
def test_oracle_prng_reproduces_the_reference_bit_for_bit():
    kat = fixture()["prng"]
    for v in kat["split"]:
        assert P.split(np.array(v["key"], np.uint32), v["num"]).tolist() == v["out"]
    for v in kat["normal"]:
        np.testing.assert_allclose(P.normal(np.array(v["key"], np.uint32), tuple(v["shape"])), v["out"], rtol=3e-7, atol=1e-7)
    for v in kat["uniform"]:
        np.testing.assert_array_equal(P.uniform(np.array(v["key"], np.uint32), tuple(v["shape"])), np.array(v["out"], np.float32))
    for v in kat["choice"]:                                   # randint: the one piece without an external KAT so far
        idx = P.choice_indices(np.array(v["key"], np.uint32), v["n"], v["batch"])
        assert idx[:16].tolist() == v["head"] and int(idx.astype(np.int64).sum()) == v["sum"], (v["n"], v["batch"])
        assert int(np.bitwise_xor.reduce(idx.astype(np.int64))) == v["xor"]
        if v["indices"] is not None:
            assert idx.tolist() == v["indices"]


@needs_fixture
@pytest.mark.parametrize("case", ["gauss1", "gauss2", "logreg"])
@pytest.mark.parametrize("name", sorted(SPECS))
def test_oracle_trajectories_match_the_reference(case, name):
    fx = fixture()
    rec = fx["trajectories"][f"{case}/{name}"]
    fam, data, batch, th0 = rebuild_cases(fx)[case]
    kw = dict(SPECS[name])
    if kw.pop("cv", False):
        kw["centering_value"] = [t + np.float32(0.01) for t in th0]
    kern = S.build_kernel(name, fam, data, batch, rec["dt"], **kw)
    key = P.PRNGKey(0)
    key, sub = P.split(key)
    st = kern.init_fn(sub, th0)
    assert rel(np.concatenate([np.ravel(g) for g in st.param_grad]), np.concatenate([np.ravel(g) for g in rec["init_grad"]])) <= 3e-5
    for i, step in enumerate(rec["steps"]):
        key, sub = P.split(key)
        st = kern.kernel(i, sub, st)
        got = np.concatenate([s["x"].ravel() for s in st.diffusion_state])
        ref = np.concatenate([np.ravel(l) for l in step["params"]])
        assert rel(got, ref) <= (1e-5 if i == 0 else 1e-4), f"{case}/{name} step {i}"


@needs_fixture
def test_oracle_ksd_matches_the_reference():
    k = fixture()["ksd"]
    got = OK.imq_ksd_literal(np.array(k["samples"], np.float32), np.array(k["grads"], np.float32))
    assert abs(got - k["value"]) <= 1e-4 * abs(k["value"])      # the reference sums in fp32 (ksd.py:60-65)


@needs_fixture
@pytest.mark.gpu
@pytest.mark.parametrize("case", ["gauss1", "gauss2", "logreg"])
@pytest.mark.parametrize("name", sorted(SPECS))
def test_cuda_trajectories_match_the_reference(cuda, case, name):
    from test_gpu_parity import as_params, build_pair, leaves_of
    from sgmcmcjax_b200.models import gaussian_mean as G
    from sgmcmcjax_b200.models import logistic_regression as L

    fx = fixture()
    rec = fx["trajectories"][f"{case}/{name}"]
    fam, data, batch, th0 = rebuild_cases(fx)[case]
    fns = {"gauss1": (G.loglikelihood, G.logprior), "gauss2": (G.loglikelihood_list_array, G.logprior_list_array),
           "logreg": (L.loglikelihood, L.logprior)}[case]
    kw = {k: v for k, v in SPECS[name].items() if k != "strict_reference_bugs"}
    center = [t + np.float32(0.01) for t in th0]
    _, (init_fn, kernel, get_params) = build_pair(name, rec["dt"], kw, fam, fns, data, batch, center,
                                                  strict=SPECS[name].get("strict_reference_bugs", False))
    key = P.PRNGKey(0)
    key, sub = P.split(key)
    st = init_fn(sub, as_params(th0))
    for i, step in enumerate(rec["steps"]):
        key, sub = P.split(key)
        st = kernel(i, sub, st)
        got = np.concatenate([np.ravel(l) for l in leaves_of(get_params(st))])
        ref = np.concatenate([np.ravel(l) for l in step["params"]])
        assert rel(got, ref) <= (1e-5 if i == 0 else 1e-4), f"{case}/{name} step {i}"


@needs_fixture
@pytest.mark.gpu
def test_cuda_indices_and_ksd_match_the_reference(cuda):
    from sgmcmcjax_b200 import ksd as NKSD
    from sgmcmcjax_b200 import random as R

    fx = fixture()
    for v in fx["prng"]["choice"]:
        idx = R.choice_indices(np.array(v["key"], np.uint32), v["n"], v["batch"]).cpu().numpy()
        assert idx[:16].tolist() == v["head"] and int(idx.astype(np.int64).sum()) == v["sum"]
    k = fx["ksd"]
    got = float(NKSD.imq_KSD(np.array(k["samples"], np.float32), np.array(k["grads"], np.float32)))
    assert abs(got - k["value"]) <= 1e-4 * abs(k["value"])
