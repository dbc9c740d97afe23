"""Profiling driver for the big-state sampler: a few launches of one configuration (used under ncu).
usage: prof_big.py mlp_psgld|mlp_sghmc|pmf_sgnht|pmf_badodabcv [chains] [steps_per_launch] [launches]"""
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from sgmcmcjax_b200 import _native, random, samplers
from sgmcmcjax_b200._engine import keys_to_device
from sgmcmcjax_b200.models import bayesian_mlp as B, pmf as F

which = sys.argv[1] if len(sys.argv) > 1 else "mlp_psgld"
C = int(sys.argv[2]) if len(sys.argv) > 2 else 148
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
launches = int(sys.argv[4]) if len(sys.argv) > 4 else 3
if which.startswith("mlp"):
    X = torch.rand(60000, 784, device="cuda")
    y = torch.nn.functional.one_hot(torch.randint(0, 10, (60000,), device="cuda"), 10).float()
    params = B.init_network(random.PRNGKey(1), [784, 100, 10])
    if which == "mlp_psgld":
        s = samplers.build_psgld_sampler(1e-3, B.loglikelihood, B.logprior, (X, y), 100, pbar=False)
    else:
        s = samplers.build_sghmc_sampler(1e-6, 4, B.loglikelihood, B.logprior, (X, y), 100, pbar=False)
    flat = torch.cat([t.reshape(-1) for wb in params for t in wb])
    s.engine.flatten_params(params, None)
else:
    nu, ni, r, Np = 943, 1682, 20, 100_000
    ij = torch.stack([torch.randint(0, nu, (Np,), device="cuda"), torch.randint(0, ni, (Np,), device="cuda")], 1).int()
    rt = torch.randint(1, 6, (Np,), device="cuda").float()
    U = 0.1 * random.normal(random.PRNGKey(3), (nu, r)); V = 0.1 * random.normal(random.PRNGKey(4), (ni, r))
    if which == "pmf_badodabcv":
        s = samplers.build_badodabCV_sampler(1e-5, F.loglikelihood, F.logprior, (ij, rt), 1000, (U, V), pbar=False)
    else:
        s = samplers.build_sgnht_sampler(1e-5, F.loglikelihood, F.logprior, (ij, rt), 1000, pbar=False)
    s.engine.flatten_params((U, V), None)
    flat = torch.cat([U.reshape(-1), V.reshape(-1)])
eng = s.engine
th = flat[None].repeat(C, 1).contiguous()
keys, _ = keys_to_device(random.split(random.PRNGKey(2), C))
st = eng.new_state(th)
eng.init(st, keys, sampler_mode=True)
samples = torch.empty((C, K, eng.P), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
i0 = 0
for l in range(launches):
    a.record(); eng.run(st, i0, K, samples); b.record(); i0 += K
    torch.cuda.synchronize()
    print(f"launch {l}: {a.elapsed_time(b):.3f} ms for {C} chains x {K} samples", flush=True)
assert torch.isfinite(st["theta"]).all()
