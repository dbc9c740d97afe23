"""Why is imq_KSD slower inside bench.py than alone?  Same call, timed (a) on fresh data, (b) after gen_data's 400 MB
table exists, (c) after a sampler has run."""
import sys, torch
sys.path.insert(0, "/root/repo")
from sgmcmcjax_b200 import ksd, random, samplers
from sgmcmcjax_b200.models import logistic_regression as L
n, d = 50000, 100
def leg(tag, theta):
    pts = (theta[None, :] + 0.02 * random.normal(random.PRNGKey(7), (n, d))).contiguous()
    grads = (-50.0 * (pts - theta[None, :]) + random.normal(random.PRNGKey(8), (n, d))).contiguous()
    for _ in range(2): ksd.imq_KSD(pts, grads)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): v = ksd.imq_KSD(pts, grads)
    b.record(); torch.cuda.synchronize()
    print(f"{tag}: {a.elapsed_time(b) / 3:.3f} ms  ksd={float(v):.6f}  ptr%2MB={pts.data_ptr() % (2 << 20)}", flush=True)
leg("fresh process, theta ~ 3 N(0,1)", random.normal(random.PRNGKey(1), (d,)) * 3.0)
theta_true, X, y = L.gen_data(random.PRNGKey(42), d, 1_000_000)
leg("after gen_data (400 MB table)", theta_true)
s = samplers.build_sgld_sampler(1e-5, L.loglikelihood, L.logprior, (X, y), 10000, pbar=False)
import numpy as np
keys = random.split(random.PRNGKey(0), 1024)
out = s(keys, 20, np.tile(theta_true.cpu().numpy(), (1024, 1)))
leg("after a 1024-chain sampler call", theta_true)
del X, y, s, out
torch.cuda.empty_cache()
leg("after freeing the table", theta_true)
