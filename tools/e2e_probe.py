"""Where does the end-to-end time of sampler(key, N, params) go? (numpy in -> numpy out, C2 shape)"""
import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from sgmcmcjax_b200 import random, samplers, _tree
from sgmcmcjax_b200._engine import keys_to_device
from sgmcmcjax_b200.kernels import _like
from sgmcmcjax_b200.models import logistic_regression as L
theta_true, X, y = L.gen_data(random.PRNGKey(42), 100, 1_000_000)
s = samplers.build_sgld_sampler(1e-5, L.loglikelihood, L.logprior, (X, y), 10_000, pbar=False)
eng = s.engine
C, K = 1024, 50
keys_host = np.ascontiguousarray(random.split(random.PRNGKey(0), C))
th0 = np.tile(theta_true.cpu().numpy(), (C, 1)).astype(np.float32)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(4):
    t0 = T(); out = s(keys_host, K, th0); t1 = T()
    print(f"rep {rep}: total {1e3*(t1-t0):.2f} ms -> {C*K/(t1-t0)/1e6:.3f} M chain-steps/s")
# staged
for rep in range(2):
    t = [T()]
    keys, kb = keys_to_device(keys_host); t.append(T())
    flat, treedef, pb = eng.flatten_params(th0, broadcast_chains=C); t.append(T())
    st = eng.new_state(flat); t.append(T())
    eng.init(st, keys, sampler_mode=True); t.append(T())
    samples = torch.empty((C, K, eng.P), dtype=torch.float32, device="cuda"); t.append(T())
    eng.run(st, 0, K, samples); t.append(T())
    host = _like(samples, th0); t.append(T())
    names = ["keys H2D", "params H2D+flatten", "new_state", "init_fn", "alloc samples", "run K steps", "D2H (pinned)"]
    print("  ".join(f"{n} {1e3*(b-a):.2f}" for n, a, b in zip(names, t[:-1], t[1:])))
