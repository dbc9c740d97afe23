"""Long-run check of the barrier-free cluster exchange: the README chain for 5000 steps in ONE launch as a 16-CTA cluster,
twice (must be bit-identical: a race in the st.async / mbarrier protocol would show as a difference), and against one CTA."""
import os, sys, numpy as np
sys.path.insert(0, "/root/repo")
from sgmcmcjax_b200 import random, samplers
from sgmcmcjax_b200.models import gaussian_mean as G
X = random.normal(random.PRNGKey(0), (10000, 100))
def run(cluster, warps):
    if cluster is None:
        os.environ.pop("SGMCMC_CLUSTER", None); os.environ.pop("SGMCMC_WARPS_PER_CHAIN", None)
    else:
        os.environ["SGMCMC_CLUSTER"], os.environ["SGMCMC_WARPS_PER_CHAIN"] = str(cluster), str(warps)
    s = samplers.build_sgld_sampler(1e-5, G.loglikelihood, G.logprior, (X,), 1000, pbar=False)
    return s(random.PRNGKey(3), 5000, np.zeros(100, np.float32))
a, b = run(None, None), run(None, None)
print("default cluster shape, two runs bit-identical:", np.array_equal(a, b), a.shape)
for cl, w in ((16, 8), (7, 5), (2, 16), (16, 1)):
    c, d = run(cl, w), run(cl, w)
    print(f"cluster {cl} x {w}: two runs bit-identical: {np.array_equal(c, d)}; max |diff| to the default shape over 5000 steps: {np.abs(c - a).max():.3e}")
e = run(1, 8)
print(f"one CTA: max |diff| to the cluster: {np.abs(e - a).max():.3e}  (|theta| max {np.abs(a).max():.3f})")
