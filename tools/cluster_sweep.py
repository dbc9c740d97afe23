"""Launch-shape sweep of the vector sampler (SGMCMC_CLUSTER x SGMCMC_WARPS_PER_CHAIN) for the few-chain regimes:
   c2  : config 2 with CHAINS chains on this GPU (128 = the 8-GPU shard of the 1024 chains), K steps per launch
   c1  : README Gaussian, one chain, 1000 steps per launch
Usage: python tools/cluster_sweep.py c2 [chains] [K] | c1
Prints one line per shape: ms per launch, chain-steps/s, fraction of the measured HBM peak (c2) / us per step (c1)."""
import json, os, sys
import numpy as np, torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgmcmcjax_b200 import random, samplers
from sgmcmcjax_b200._engine import keys_to_device
from sgmcmcjax_b200.models import gaussian_mean as G
from sgmcmcjax_b200.models import logistic_regression as L

mode = sys.argv[1] if len(sys.argv) > 1 else "c2"
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6552.3


def timed(eng, st, K, samples, reps=6):
    for l in range(2):
        eng.run(st, l * K, K, samples)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for l in range(reps):
        eng.run(st, (2 + l) * K, K, samples)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def shapes(arg, default):
    if arg:
        return [tuple(int(v) for v in s.split("x")) for s in arg.split(",")]
    return default


if mode == "c2":
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    theta_true, X, y = L.gen_data(random.PRNGKey(42), 100, 1_000_000)
    th = theta_true.cpu().numpy()
    cv = os.environ.get("SWEEP_CV") == "1"
    if cv:
        s = samplers.build_sgldCV_sampler(1e-4, L.loglikelihood, L.logprior, (X, y), 10_000, th, pbar=False)
    else:
        s = samplers.build_sgld_sampler(1e-5, L.loglikelihood, L.logprior, (X, y), 10_000, pbar=False)
    eng = s.engine
    keys, _ = keys_to_device(random.split(random.PRNGKey(0), C))
    samples = torch.empty((C, K, 100), device="cuda")
    default = [(0, 0), (1, 32), (2, 16), (2, 32), (3, 12), (4, 8), (4, 9), (5, 7), (8, 4), (9, 4), (4, 16), (8, 8), (16, 2), (16, 4)]
    for cl, w in shapes(sys.argv[4] if len(sys.argv) > 4 else None, default):
        os.environ["SGMCMC_CLUSTER"], os.environ["SGMCMC_WARPS_PER_CHAIN"] = str(cl), str(w)
        st = eng.new_state(torch.from_numpy(np.tile(th, (C, 1)).astype(np.float32)).cuda())
        eng.init(st, keys, sampler_mode=True)
        ms = timed(eng, st, K, samples)
        cs = C * K / (ms * 1e-3)
        print(f"c2 chains={C} K={K} cluster={cl} warps={w}: {ms:.4f} ms/launch  {cs/1e6:.3f} M chain-steps/s  "
              f"hbm_frac={cs * 4041200 / 1e9 / PEAK:.3f}", flush=True)
else:
    K = 1000
    Xg = random.normal(random.PRNGKey(0), (10000, 100))
    s = samplers.build_sgld_sampler(1e-5, G.loglikelihood, G.logprior, (Xg,), 1000, pbar=False)
    eng = s.engine
    keys, _ = keys_to_device(random.split(random.PRNGKey(0), 1))
    samples = torch.empty((1, K, 100), device="cuda")
    default = [(0, 0), (1, 16), (2, 16), (4, 8), (4, 16), (8, 4), (8, 8), (8, 16), (16, 2), (16, 4), (16, 8), (12, 4), (7, 8)]
    for cl, w in shapes(sys.argv[2] if len(sys.argv) > 2 else None, default):
        os.environ["SGMCMC_CLUSTER"], os.environ["SGMCMC_WARPS_PER_CHAIN"] = str(cl), str(w)
        st = eng.new_state(torch.zeros(1, 100, device="cuda"))
        eng.init(st, keys, sampler_mode=True)
        ms = timed(eng, st, K, samples)
        print(f"c1 cluster={cl} warps={w}: {ms * 1000 / K:.3f} us per chain-step", flush=True)
