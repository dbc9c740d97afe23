"""GPU parity for the big-state families (Bayesian MLP, PMF): csrc/big_sampler.cuh + csrc/mlp_grad.cuh
through the C ABI against the numpy oracle, same seeded inputs and keys.

Same bars as tests/test_gpu_parity.py: single-step updates <= 1e-5 of max|ref| (state injected from the
oracle before every step), carried gradient 1e-5 (3e-5 for CV / SVRG), short free-running trajectories 1e-4.
"""
import numpy as np
import pytest

from oracle import jax_prng as P
from oracle import models as M
from oracle import sgmcmc as S

pytestmark = pytest.mark.gpu

RTOL_STEP = 1e-5


def close(a, b, rtol=RTOL_STEP, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, f"{what}: max err / max|ref| = {err:.3e} > {rtol:.1e}"


def to_np(t):
    return t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)


def flat(leaves):
    return np.concatenate([np.ravel(np.asarray(l, np.float32)) for l in leaves])


def tree_leaves(p):
    out = []

    def rec(n):
        if isinstance(n, (list, tuple)):
            for c in n:
                rec(c)
        else:
            out.append(to_np(n))

    rec(p)
    return out


# ------------------------------------------------------------------------------ cases
def mlp_case(sizes=(20, 12, 5), n=300, seed=5, scale=0.3):
    from sgmcmcjax_b200.models import bayesian_mlp as B

    ks = P.split(P.PRNGKey(seed), 3 + len(sizes))
    X = P.uniform(ks[0], (n, sizes[0])).astype(np.float32)
    labels = P.choice_indices(ks[1], sizes[-1], n)
    y = np.eye(sizes[-1], dtype=np.float32)[labels]
    leaves = []
    for l, (m, k) in enumerate(zip(sizes[:-1], sizes[1:])):
        kk = P.split(ks[3 + l])
        leaves += [(np.float32(scale) * P.normal(kk[0], (k, m))).astype(np.float32),
                   (np.float32(scale) * P.normal(kk[1], (k,))).astype(np.float32)]
    return M.BayesianMLP(), (B.loglikelihood, B.logprior), (X, y), leaves


def mlp_params(leaves):
    return [(leaves[2 * l], leaves[2 * l + 1]) for l in range(len(leaves) // 2)]


def pmf_case(nu=30, ni=40, r=4, n=500, seed=6):
    from sgmcmcjax_b200.models import pmf as F

    ks = P.split(P.PRNGKey(seed), 5)
    ij = np.stack([P.choice_indices(ks[0], nu, n), P.choice_indices(ks[1], ni, n)], axis=1).astype(np.int32)
    r_ = (1 + P.choice_indices(ks[2], 5, n)).astype(np.float32)
    U = (np.float32(0.5) * P.normal(ks[3], (nu, r))).astype(np.float32)
    V = (np.float32(0.5) * P.normal(ks[4], (ni, r))).astype(np.float32)
    return M.PMF(), (F.loglikelihood, F.logprior), (ij, r_), [U, V]


def pmf_params(leaves):
    return (leaves[0], leaves[1])


CASES = {"mlp": (mlp_case, mlp_params, 32), "pmf": (pmf_case, pmf_params, 50)}

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


def build_pair(name, dt, kw, fam, fns, data, batch, center_leaves, to_params, sampler=False):
    from sgmcmcjax_b200 import kernels as NK
    from sgmcmcjax_b200 import samplers as NS

    kw = dict(kw)
    okw, pos = {}, []
    L = kw.pop("L", None)
    if kw.pop("cv", False):
        okw["centering_value"] = center_leaves
        pos.append(to_params(center_leaves))
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
    if sampler:
        return ok, getattr(NS, f"build_{name}_sampler")(*args, pbar=False)
    return ok, getattr(NK, f"build_{name}_kernel")(*args)


def compare_state(ostate, nstate, get_params, what, rtol=RTOL_STEP):
    # badodab: c2 = sqrt(|1 - exp(-alpha h)^2| / 2 alpha) with alpha h ~ 1e-4 turns a 1-ulp difference of
    # alpha (the |v| reduction order differs between numpy and the device) into ~1e-4 relative noise
    # amplitude ("BADODAB conditioning"); x only sees it scaled by h/2.
    v_rtol = 2e-4 if "badodab" in what else 2 * rtol
    for j, (ref, got) in enumerate(zip(ostate.diffusion_state, tree_leaves(get_params(nstate)))):
        close(got, ref["x"], rtol=rtol, what=f"{what} x leaf {j}")
    packed = nstate.diffusion_state.packed_state
    first = ostate.diffusion_state[0]
    bufs = {"m": "aux1", "v": "aux2"} if "m" in first else {"v": "aux1"}
    for fld, buf in bufs.items():
        if fld in first:
            close(to_np(packed[buf])[0], flat([s[fld] for s in ostate.diffusion_state]), rtol=v_rtol,
                  what=f"{what} {fld}")
    if "alpha" in first:
        ref = np.array([s["alpha"] for s in ostate.diffusion_state], np.float64)
        np.testing.assert_allclose(to_np(packed["alpha"])[0], ref, rtol=v_rtol, err_msg=f"{what} alpha")
    return flat(ostate.param_grad), flat(tree_leaves(nstate.param_grad))


def inject(ostate, nstate):
    import torch

    packed = nstate.diffusion_state.packed_state
    first = ostate.diffusion_state[0]

    def put(name, leaves):
        packed[name].copy_(torch.from_numpy(flat(leaves)[None]))

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
    return 3e-5 if ("CV" in name or "SVRG" in name) else 1e-5


# ------------------------------------------------------------------------------ tests
@pytest.mark.parametrize("name,dt,kw", KERNELS, ids=[k[0] for k in KERNELS])
@pytest.mark.parametrize("case", ["mlp", "pmf"])
def test_init_and_single_steps_match_oracle(cuda, name, dt, kw, case):
    make, to_params, batch = CASES[case]
    fam, fns, data, th0 = make()
    center = [np.asarray(t * np.float32(0.98) + np.float32(0.003)) for t in th0]
    ok, (init_fn, kernel, get_params) = build_pair(name, dt, kw, fam, fns, data, batch, center, to_params)
    key = P.PRNGKey(0)
    key, sub = P.split(key)
    ost = ok.init_fn(sub, th0)
    nst = init_fn(sub, to_params(th0))
    g_ref, g_got = compare_state(ost, nst, get_params, f"{name}/{case} init")
    assert np.abs(g_got - g_ref).max() / max(np.abs(g_ref).max(), 1.0) <= grad_tol(name), f"{name}/{case} init grad"
    for i in range(3):
        key, sub = P.split(key)
        nst = kernel(i, sub, inject(ost, nst))
        ost = ok.kernel(i, sub, ost)
        g_ref, g_got = compare_state(ost, nst, get_params, f"{name}/{case} step {i}")
        assert np.abs(g_got - g_ref).max() / max(np.abs(g_ref).max(), 1.0) <= grad_tol(name), f"{name}/{case} grad {i}"
        if ost.svrg_state is not None:
            packed = nst.diffusion_state.packed_state
            close(to_np(packed["svrg_center"])[0], flat(ost.svrg_state[0]), what="svrg centre")
            close(to_np(packed["svrg_grad"])[0], flat(ost.svrg_state[1]), rtol=2e-5, what="svrg fb grad")


@pytest.mark.parametrize("name,dt,kw", [KERNELS[0], KERNELS[3], KERNELS[5], KERNELS[9], KERNELS[10]],
                         ids=["sgld", "psgld", "sgnht", "badodabCV", "sghmc"])
@pytest.mark.parametrize("case", ["mlp", "pmf"])
def test_sampler_container_and_trajectory(cuda, name, dt, kw, case):
    """Pytree container / shapes as tests/test_samplers.py:95-117 and an 8-step free-running trajectory."""
    make, to_params, batch = CASES[case]
    fam, fns, data, th0 = make()
    ok, sampler = build_pair(name, dt, kw, fam, fns, data, batch, th0, to_params, sampler=True)
    ref, _, _ = S.run_sampler(ok, P.PRNGKey(3), 8, th0)
    out = sampler(P.PRNGKey(3), 8, to_params(th0))
    assert type(out) is type(to_params(th0))
    got = tree_leaves(out)
    assert len(got) == len(th0)
    for g, r, t in zip(got, ref, th0):
        assert g.shape == (8,) + t.shape and np.isfinite(g).all()
        close(g, r, rtol=1e-4, what=f"{name}/{case} trajectory")


@pytest.mark.parametrize("name,dt,kw", [KERNELS[3], KERNELS[5], KERNELS[10]], ids=["psgld", "sgnht", "sghmc"])
@pytest.mark.parametrize("case", ["mlp", "pmf"])
def test_thinning_and_on_device_summaries(cuda, name, dt, kw, case):
    """sgmcmc_run_thinned on the big-state kernel: kept states == every 3rd state of the unthinned run; moments == numpy."""
    from sgmcmcjax_b200 import samplers as NS

    make, to_params, batch = CASES[case]
    fam, fns, data, th0 = make()
    args = ([dt, kw["L"]] if "L" in kw else [dt]) + [fns[0], fns[1], data, batch]
    build = getattr(NS, f"build_{name}_sampler")
    full = tree_leaves(build(*args, pbar=False)(P.PRNGKey(4), 14, to_params(th0)))
    kept, summ = build(*args, pbar=False, thin=3, summaries=True)(P.PRNGKey(4), 14, to_params(th0))
    assert summ["count"] == 4
    for f, k, mu, var, t0 in zip(full, tree_leaves(kept), tree_leaves(summ["mean"]), tree_leaves(summ["var"]), th0):
        assert k.shape == (4,) + f.shape[1:]
        if case == "pmf":           # float atomics in the gradient: two runs agree to rounding, not bit for bit
            close(k, f[2::3][:4], rtol=1e-5, what="pmf thinned")
        else:
            assert np.array_equal(k, f[2::3][:4])
        close(mu, k.astype(np.float64).mean(0), rtol=1e-6, what="on-device mean")
        # one-pass variance around theta_0 in fp32: exact to ~1e-7 of the mean-square displacement from theta_0
        msd = ((k.astype(np.float64) - t0) ** 2).mean(0)
        assert (np.abs(var - k.astype(np.float64).var(0)) <= 1e-5 * msd + 1e-14).all()


@pytest.mark.parametrize("case", ["mlp", "pmf"])
def test_multichain_equals_single_chains(cuda, case):
    make, to_params, batch = CASES[case]
    fam, fns, data, th0 = make()
    _, sampler = build_pair("sgld", 1e-5, {}, fam, fns, data, batch, th0, to_params, sampler=True)
    keys = P.split(P.PRNGKey(0), 3)
    multi_params = to_params([np.stack([t * np.float32(1 + 0.1 * c) for c in range(3)]) for t in th0])
    multi = tree_leaves(sampler(keys, 5, multi_params))
    for c in range(3):
        single = tree_leaves(sampler(keys[c], 5, to_params([t * np.float32(1 + 0.1 * c) for t in th0])))
        for m, s_ in zip(multi, single):
            assert m.shape[:2] == (3, 5)
            if case == "pmf":       # float atomics: order-dependent rounding
                close(m[c], s_, rtol=1e-5, what="pmf multi-chain")
            else:
                assert np.array_equal(m[c], s_)


@pytest.mark.parametrize("sizes,n,batch", [((784, 100, 10), 600, 100), ((50, 128, 3), 200, 130), ((7, 5), 64, 9),
                                           ((33, 17, 11, 4), 100, 100),
                                           # aligned input width (the cp.async operand ring of the forward GEMM) with: five
                                           # stager items per thread (hidden width 128); a narrow hidden layer and a ragged
                                           # batch; a full 128-row sub-batch with 112 hidden units; two sub-batches
                                           ((64, 128, 5), 300, 100), ((64, 16, 4), 200, 37), ((128, 112, 10), 400, 128),
                                           ((256, 100, 10), 500, 200)])
def test_mlp_shapes_fullbatch_grad(cuda, sizes, n, batch):
    """Layer geometries: the BASELINE config-3 net 784-100-10 (batch 100), widths up to 128, one layer
    (multinomial regression), three layers; full-batch K2 and a minibatch step."""
    import torch

    fam, fns, data, th0 = mlp_case(sizes, n, seed=len(sizes) + n, scale=0.05 if sizes[0] > 100 else 0.3)
    ok, (init_fn, kernel, get_params) = build_pair("sgld", 1e-6, {}, fam, fns, data, batch, th0, mlp_params)
    sub = P.split(P.PRNGKey(2))[1]
    ost, nst = ok.init_fn(sub, th0), init_fn(sub, mlp_params(th0))
    close(flat(tree_leaves(nst.param_grad)), flat(ost.param_grad), rtol=1e-5, what=f"mlp {sizes} minibatch grad")
    thetas = np.stack([flat(th0), flat(th0) * np.float32(0.5)])
    got = init_fn.engine.fullbatch_grad(torch.from_numpy(thetas).cuda()).cpu().numpy()
    ref = flat(fam.grad_log_post(th0, data, n))
    close(got[0], ref, rtol=1e-5, what=f"mlp {sizes} fullbatch grad")


def test_pmf_full_geometry_step(cuda):
    """SURVEY.md a15 geometry: 943 x 1682, rank 20, N = 1e5, batch 1000 - one SGNHT step vs the oracle."""
    fam, fns, data, th0 = pmf_case(943, 1682, 20, 100_000, seed=11)
    ok, (init_fn, kernel, get_params) = build_pair("sgnht", 1e-5, {}, fam, fns, data, 1000, th0, pmf_params)
    key, sub = P.split(P.PRNGKey(1))
    ost, nst = ok.init_fn(sub, th0), init_fn(sub, pmf_params(th0))
    key, sub = P.split(key)
    nst = kernel(0, sub, inject(ost, nst))
    ost = ok.kernel(0, sub, ost)
    g_ref, g_got = compare_state(ost, nst, get_params, "pmf full step")
    assert np.abs(g_got - g_ref).max() / np.abs(g_ref).max() <= 1e-5


def test_big_family_errors(cuda):
    from sgmcmcjax_b200 import samplers as NS

    fam, fns, data, th0 = mlp_case()
    s = NS.build_sgld_sampler(1e-5, fns[0], fns[1], data, 10, pbar=False)
    with pytest.raises(TypeError):
        s(P.PRNGKey(0), 2, [(th0[0], th0[1])])                     # one layer only: sizes do not match y
    with pytest.raises(TypeError):
        s(P.PRNGKey(0), 2, th0[0])                                 # not a (W, b) list
    out = s(P.PRNGKey(0), 2, mlp_params(th0))
    assert isinstance(out, list) and isinstance(out[0], tuple) and out[0][0].shape == (2,) + th0[0].shape
    with pytest.raises(TypeError):                                 # geometry is fixed once bound
        s(P.PRNGKey(0), 2, mlp_params([t[:5] if t.ndim == 1 else t[:5] for t in th0]))


def test_pmf_rejects_out_of_range_ids(cuda):
    """1-based (MovieLens-style) or negative ids would index past U / V in the gradient scatter: refused at ingestion."""
    from sgmcmcjax_b200 import _native
    from sgmcmcjax_b200 import samplers as NS

    fam, fns, data, th0 = pmf_case()
    ij, r = data
    for bad in (ij + np.array([1, 0], np.int32), ij + np.array([0, 1], np.int32), ij - np.array([1, 0], np.int32)):
        assert bad[:, 0].max() >= 30 or bad[:, 1].max() >= 40 or bad.min() < 0
        s = NS.build_sgld_sampler(1e-5, fns[0], fns[1], (bad, r), 50, pbar=False)
        with pytest.raises(_native.NativeError, match="id outside"):
            s(P.PRNGKey(0), 2, pmf_params(th0))
    NS.build_sgld_sampler(1e-5, fns[0], fns[1], (ij, r), 50, pbar=False)(P.PRNGKey(0), 2, pmf_params(th0))


@pytest.mark.parametrize("name,dt,kw", [KERNELS[0], KERNELS[2], KERNELS[5], KERNELS[9], KERNELS[11]],
                         ids=["sgld", "sgld_SVRG", "sgnht", "badodabCV", "sghmcCV"])
def test_pmf_compact_gradient_mode(cuda, monkeypatch, name, dt, kw):
    """SGMCMC_PMF_COMPACT=1: the minibatch gradient as compact shared-memory rows summed by row owners in stream order
    (csrc/big_sampler.cuh PmfCompact).  Same single-step bars against the oracle as the default scatter-add path, the
    two paths agree, and the compact path is bit-reproducible (no float atomics)."""
    fam, fns, data, th0 = pmf_case(60, 80, 8, 3000, seed=8)                 # rank 8: whole float4 factor rows
    center = [np.asarray(t * np.float32(0.98) + np.float32(0.003)) for t in th0]
    runs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("SGMCMC_PMF_COMPACT", mode)
        ok, (init_fn, kernel, get_params) = build_pair(name, dt, kw, fam, fns, data, 200, center, pmf_params)
        key = P.PRNGKey(0)
        key, sub = P.split(key)
        ost = ok.init_fn(sub, th0)
        nst = init_fn(sub, pmf_params(th0))
        g_ref, g_got = compare_state(ost, nst, get_params, f"{name} compact={mode} init")
        assert np.abs(g_got - g_ref).max() / max(np.abs(g_ref).max(), 1.0) <= grad_tol(name)
        for i in range(3):
            key, sub = P.split(key)
            nst = kernel(i, sub, inject(ost, nst))
            ost = ok.kernel(i, sub, ost)
            g_ref, g_got = compare_state(ost, nst, get_params, f"{name} compact={mode} step {i}")
            assert np.abs(g_got - g_ref).max() / max(np.abs(g_ref).max(), 1.0) <= grad_tol(name)
        _, sampler = build_pair(name, dt, kw, fam, fns, data, 200, center, pmf_params, sampler=True)
        runs[mode] = [tree_leaves(sampler(P.PRNGKey(3), 6, pmf_params(th0))) for _ in range(2)]
    if "SVRG" not in name:                 # the SVRG refresh is a FULL pass: dense buffer, float atomics in either mode
        for a, b in zip(runs["1"][0], runs["1"][1]):
            assert np.array_equal(a, b)                                    # deterministic
    for a, b in zip(runs["1"][0], runs["0"][0]):
        close(a, b, rtol=2e-5, what=f"{name}: compact vs scatter-add trajectory")


@pytest.mark.parametrize("name,dt,kw", [KERNELS[3], KERNELS[5], KERNELS[9], KERNELS[10], KERNELS[12]],
                         ids=["psgld", "sgnht", "badodabCV", "sghmc", "sghmc_SVRG"])
@pytest.mark.parametrize("case", ["mlp", "pmf"])
def test_work_queue_equals_one_cta_per_chain(cuda, monkeypatch, name, dt, kw, case):
    """The in-kernel (chain, block of steps) queue of csrc/big_sampler.cuh (more chains than resident CTAs): 5 chains on a
    grid of 2 persistent CTAs, 3 blocks per chain, so chains migrate between CTAs at every block - same samples as one
    CTA per chain (bit for bit for the MLP; PMF's scatter-add is order dependent), thinning and moments included."""
    make, to_params, batch = CASES[case]
    fam, fns, data, th0 = make()
    keys = P.split(P.PRNGKey(9), 5)
    params = to_params([np.stack([t * np.float32(1 + 0.05 * c) for c in range(5)]) for t in th0])
    center = [np.asarray(t * np.float32(0.98) + np.float32(0.003)) for t in th0]
    outs = {}
    for mode in ("plain", "queue"):
        if mode == "plain":
            monkeypatch.setenv("SGMCMC_BIG_QUEUE", "0")
        else:
            monkeypatch.setenv("SGMCMC_BIG_QUEUE", "1")
            monkeypatch.setenv("SGMCMC_BIG_GRID", "2")
            monkeypatch.setenv("SGMCMC_BIG_BLOCKS", "3")
        _, sampler = build_pair(name, dt, kw, fam, fns, data, batch, center, to_params, sampler=True)
        outs[mode] = tree_leaves(sampler(keys, 7, params))
        _, thin_sampler = build_pair(name, dt, kw, fam, fns, data, batch, center, to_params, sampler=True)
    for a, b in zip(outs["plain"], outs["queue"]):
        assert a.shape[:2] == (5, 7)
        if case == "pmf":
            close(b, a, rtol=1e-5, what=f"{name} queue vs plain")
        else:
            assert np.array_equal(a, b), f"{name}: queued launch differs from one CTA per chain"
