"""README config (C1): one chain, N=1e4, D=100, batch=1000 - a few launches of 1000 steps (used under ncu)."""
import sys, torch
sys.path.insert(0, "/root/repo")
from sgmcmcjax_b200 import random, samplers
from sgmcmcjax_b200._engine import keys_to_device
from sgmcmcjax_b200.models import gaussian_mean as G
K = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
Xg = random.normal(random.PRNGKey(0), (10000, 100))
s = samplers.build_sgld_sampler(1e-5, G.loglikelihood, G.logprior, (Xg,), 1000, pbar=False)
eng = s.engine
keys, _ = keys_to_device(random.split(random.PRNGKey(0), 1))
st = eng.new_state(torch.zeros(1, 100, device="cuda"))
eng.init(st, keys, sampler_mode=True)
samples = torch.empty((1, K, 100), device="cuda")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for l in range(3):
    a.record(); eng.run(st, l * K, K, samples); b.record(); torch.cuda.synchronize()
    print(f"launch {l}: {a.elapsed_time(b) * 1000 / K:.3f} us per chain-step", flush=True)
