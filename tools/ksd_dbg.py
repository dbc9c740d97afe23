import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
from sgmcmcjax_b200 import ksd
from oracle import ksd as OK
for n, d in [(128, 16), (128, 100), (300, 100), (1000, 100)]:
    rng = np.random.default_rng(0)
    X = rng.normal(size=(n, d)).astype(np.float32)
    G = (-X + 0.3*rng.normal(size=(n, d))).astype(np.float32)
    ref = float(OK.imq_ksd_sum(X, G))
    tc = float(ksd.imq_KSD_tiles(X, G).cpu()[0]); torch.cuda.synchronize()
    si = float(ksd.imq_KSD_partial(X, G, 0, n).cpu()[0])
    print(n, d, "ref", ref, "tc", tc, "simt", si, "rel tc", abs(tc-ref)/abs(ref), "rel simt", abs(si-ref)/abs(ref), flush=True)
