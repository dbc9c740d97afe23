"""KSD micro-benchmark: tensor-core vs SIMT path on N x 100 points (CUDA-event timed)."""
import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from sgmcmcjax_b200 import ksd, random
n = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
d = 100
th = random.normal(random.PRNGKey(1), (d,)) * 3.0
pts = (th[None, :] + 0.02 * random.normal(random.PRNGKey(7), (n, d))).contiguous()
grads = (-50.0 * (pts - th[None, :]) + random.normal(random.PRNGKey(8), (n, d))).contiguous()
for method in ("tc", "simt"):
    for _ in range(2):
        v = ksd.imq_KSD(pts, grads, method=method)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5 if method == "tc" else 2
    a.record()
    for _ in range(reps):
        v = ksd.imq_KSD(pts, grads, method=method)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{method}: n={n} {ms:.3f} ms  {n*n/ms/1e6:.1f} Gpairs/s  ksd={float(v):.7g}", flush=True)
