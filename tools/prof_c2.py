"""C2 (logreg N=1e6 D=100 B=1e4, 1024 chains): a few launches of SGLD or SGLD-CV (used under ncu)."""
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from sgmcmcjax_b200 import random, samplers
from sgmcmcjax_b200._engine import keys_to_device
from sgmcmcjax_b200.models import logistic_regression as L
cv = len(sys.argv) > 1 and sys.argv[1] == "cv"
theta_true, X, y = L.gen_data(random.PRNGKey(42), 100, 1_000_000)
th = theta_true.cpu().numpy()
if cv:
    s = samplers.build_sgldCV_sampler(1e-4, L.loglikelihood, L.logprior, (X, y), 10_000, th, pbar=False)
else:
    s = samplers.build_sgld_sampler(1e-5, L.loglikelihood, L.logprior, (X, y), 10_000, pbar=False)
eng = s.engine
C, K = 1024, 10
keys, _ = keys_to_device(random.split(random.PRNGKey(0), C))
st = eng.new_state(torch.from_numpy(np.tile(th, (C, 1)).astype(np.float32)).cuda())
eng.init(st, keys, sampler_mode=True)
samples = torch.empty((C, K, 100), device="cuda")
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for l in range(4):
    a.record(); eng.run(st, l * K, K, samples); b.record(); torch.cuda.synchronize()
    print(f"launch {l}: {a.elapsed_time(b):.3f} ms", flush=True)
