"""Small instances of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck), SURVEY.md section 5:
   compute-sanitizer --tool memcheck  python tools/sanitize_cases.py [case ...]
   compute-sanitizer --tool racecheck python tools/sanitize_cases.py [case ...]
cases: vec (K1, one CTA per chain), cluster (K1, thread-block cluster per chain, DSMEM exchange), mlp / pmf (K1-big: tcgen05
layer GEMMs / compact minibatch gradient), ksd (K3: TMA + tcgen05 + TMEM), k2tc (K2 on tensor cores), opt (Adam optimiser).
Shapes are tiny: the tools slow kernels down 10-100x."""
import os, sys
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgmcmcjax_b200 import ksd, optimizer, random, samplers
from sgmcmcjax_b200.models import bayesian_mlp as B, gaussian_mean as G, logistic_regression as L, pmf as F

cases = sys.argv[1:] or ["vec", "cluster", "mlp", "pmf", "ksd", "k2tc", "opt"]


def logreg(n=600, d=20):
    th, X, y = L.gen_data(random.PRNGKey(1), d, n)
    return th.cpu().numpy(), X, y


for c in cases:
    if c in ("vec", "cluster"):
        if c == "cluster":
            os.environ["SGMCMC_CLUSTER"], os.environ["SGMCMC_WARPS_PER_CHAIN"] = "4", "4"
        th, X, y = logreg()
        keys = random.split(random.PRNGKey(0), 3)
        th3 = np.tile(th, (3, 1)).astype(np.float32)
        for name, extra in (("sgld", ()), ("sgldCV", (th,)), ("sgld_SVRG", (3,)), ("sgnht", ()), ("badodab", ())):
            s = getattr(samplers, f"build_{name}_sampler")(1e-4, L.loglikelihood, L.logprior, (X, y), 64, *extra, pbar=False)
            out = s(keys, 6, th3)
            assert np.isfinite(out).all()
        s = samplers.build_sghmc_sampler(1e-6, 2, L.loglikelihood, L.logprior, (X, y), 64, pbar=False)
        assert np.isfinite(s(keys, 4, th3)).all()
        Xg = random.normal(random.PRNGKey(2), (500, 7))
        s = samplers.build_sgld_sampler(1e-5, G.loglikelihood, G.logprior, (Xg,), 33, pbar=False)
        assert np.isfinite(s(random.PRNGKey(3), 5, np.zeros(7, np.float32))).all()
        os.environ.pop("SGMCMC_CLUSTER", None); os.environ.pop("SGMCMC_WARPS_PER_CHAIN", None)
    elif c == "mlp":
        X = torch.rand(300, 40, device="cuda")
        y = torch.nn.functional.one_hot(torch.randint(0, 5, (300,), device="cuda"), 5).float()
        params = B.init_network(random.PRNGKey(1), [40, 24, 5])
        for name, head in (("psgld", (1e-3,)), ("sghmc", (1e-6, 2)), ("sgldCV", (1e-4,))):
            extra = (params,) if name == "sgldCV" else ()
            s = getattr(samplers, f"build_{name}_sampler")(*head, B.loglikelihood, B.logprior, (X, y), 32, *extra, pbar=False)
            out = s(random.PRNGKey(2), 3, params)
            assert all(torch.isfinite(torch.as_tensor(t)).all() for wb in out for t in wb)
    elif c == "pmf":
        nu, ni, r, n = 40, 50, 8, 800
        ij = torch.stack([torch.randint(0, nu, (n,), device="cuda"), torch.randint(0, ni, (n,), device="cuda")], 1).int()
        rt = torch.randint(1, 6, (n,), device="cuda").float()
        U = 0.1 * random.normal(random.PRNGKey(3), (nu, r)); V = 0.1 * random.normal(random.PRNGKey(4), (ni, r))
        for compact in ("0", "1"):
            os.environ["SGMCMC_PMF_COMPACT"] = compact
            for name, extra in (("sgnht", ()), ("badodabCV", ((U, V),)), ("sgld_SVRG", (2,))):
                s = getattr(samplers, f"build_{name}_sampler")(1e-4, F.loglikelihood, F.logprior, (ij, rt), 64, *extra, pbar=False)
                out = s(random.split(random.PRNGKey(5), 2), 4, (U[None].repeat(2, 1, 1), V[None].repeat(2, 1, 1)))
                assert all(torch.isfinite(torch.as_tensor(t)).all() for t in out)
        os.environ.pop("SGMCMC_PMF_COMPACT")
    elif c == "ksd":
        for n, d in ((300, 100), (129, 7)):
            x = random.normal(random.PRNGKey(6), (n, d)); g = -x + 0.1 * random.normal(random.PRNGKey(7), (n, d))
            a, b = float(ksd.imq_KSD(x, g, method="tc")), float(ksd.imq_KSD(x, g, method="simt"))
            assert abs(a - b) <= 1e-4 * abs(b), (a, b)
    elif c == "k2tc":
        th, X, y = logreg(700, 36)
        s = samplers.build_sgld_sampler(1e-5, L.loglikelihood, L.logprior, (X, y), 32, pbar=False)
        pts = torch.from_numpy(th[None] + 0.05 * np.random.default_rng(0).normal(size=(130, 36)).astype(np.float32)).cuda()
        os.environ["SGMCMC_K2_TC"] = "1"; a = s.engine.fullbatch_grad(pts)
        os.environ["SGMCMC_K2_TC"] = "0"; b = s.engine.fullbatch_grad(pts[:4].contiguous())
        os.environ.pop("SGMCMC_K2_TC")
        assert torch.allclose(a[:4], b, rtol=1e-4, atol=1e-3)
    elif c == "opt":
        th, X, y = logreg()
        p, lp = optimizer.build_optax_optimizer(optimizer.adam(1e-2), L.loglikelihood, L.logprior, (X, y), 64)(random.PRNGKey(1), 5, th)
        assert np.isfinite(p).all() and np.isfinite(lp).all()
    torch.cuda.synchronize()
    print(f"case {c}: ok", flush=True)
