"""C2 SGLD at batch 10000 and 1000 and SGLD-CV at batch 10000 (1024 chains): ms per launch, for A/B runs of variant libraries."""
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from sgmcmcjax_b200 import random, samplers
from sgmcmcjax_b200._engine import keys_to_device
from sgmcmcjax_b200.models import logistic_regression as L
theta_true, X, y = L.gen_data(random.PRNGKey(42), 100, 1_000_000)
th = theta_true.cpu().numpy()
C = 1024
keys, _ = keys_to_device(random.split(random.PRNGKey(0), C))
def leg(name, s, K):
    eng = s.engine
    st = eng.new_state(torch.from_numpy(np.tile(th, (C, 1)).astype(np.float32)).cuda())
    eng.init(st, keys, sampler_mode=True)
    samples = torch.empty((C, K, 100), device="cuda")
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for l in range(6):
        a.record(); eng.run(st, l * K, K, samples); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print(f"{name}: {min(ts[2:]):.3f} ms per {K} steps  ({C * K / min(ts[2:]) / 1e3:.3f} M chain-steps/s)", flush=True)
leg("SGLD B=10000", samplers.build_sgld_sampler(1e-5, L.loglikelihood, L.logprior, (X, y), 10_000, pbar=False), 10)
leg("SGLD B=1000 ", samplers.build_sgld_sampler(1e-5, L.loglikelihood, L.logprior, (X, y), 1_000, pbar=False), 100)
leg("SGLD-CV B=10000", samplers.build_sgldCV_sampler(1e-4, L.loglikelihood, L.logprior, (X, y), 10_000, th, pbar=False), 10)
