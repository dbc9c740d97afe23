"""K2 benchmark: full-batch grad_log_post of the C2 logistic regression (N=1e6, D=100) at M sample points, then imq_KSD."""
import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from sgmcmcjax_b200 import ksd, random, samplers
from sgmcmcjax_b200.models import logistic_regression as L
M = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
theta_true, X, y = L.gen_data(random.PRNGKey(42), 100, N)
s = samplers.build_sgld_sampler(1e-5, L.loglikelihood, L.logprior, (X, y), 10000, pbar=False)
eng = s.engine
pts = (theta_true[None, :] + 0.01 * random.normal(random.PRNGKey(7), (M, 100))).contiguous()
for it in range(3):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g = eng.fullbatch_grad(pts); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"K2 M={M} N={N}: {ms:.2f} ms  {2*2*M*N*100/ms/1e9:.1f} algorithmic TFLOP/s  {M*N/ms/1e6:.1f} G (theta,row) pairs/s", flush=True)
a.record(); v = ksd.imq_KSD(pts, g); b.record(); torch.cuda.synchronize()
print(f"imq_KSD on the {M} points with these grads: {float(v):.6g} in {a.elapsed_time(b):.2f} ms")
# spot check vs the SIMT path on 3 points
import os
os.environ["SGMCMC_K2_TC"] = "0"
ref = eng.fullbatch_grad(pts[:3].contiguous())
print("max rel err vs SIMT path:", float((g[:3] - ref).abs().max() / ref.abs().max()))
