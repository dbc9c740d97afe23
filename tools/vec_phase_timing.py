"""Per-phase cycles of the vector sampler on the README config (one chain), thread 0 of CTA 0.  Build first:
   python tools/variant.py vtiming --only capi.cu,vec_inst_gauss.cu,vec_inst_logreg.cu,vec_inst_logreg_svrg.cu,vec_inst_logreg_cv.cu \
       -DSG_PHASE_TIMING -DSG_PHASE_TIMING_VEC -DSG_NO_MLP -DSG_NO_PMF
Usage: python tools/vec_phase_timing.py [cluster] [warps]"""
import ctypes as C, os, sys
sys.path.insert(0, "/root/repo")
os.environ.setdefault("SGMCMC_B200_LIB", "/root/repo/build/variants/lib_vtiming.so")
if len(sys.argv) > 2:
    os.environ["SGMCMC_CLUSTER"], os.environ["SGMCMC_WARPS_PER_CHAIN"] = sys.argv[1], sys.argv[2]
import torch
from sgmcmcjax_b200 import _native, random, samplers
from sgmcmcjax_b200._engine import keys_to_device
from sgmcmcjax_b200.models import gaussian_mean as G
lib = _native.load()
K = 1000
Xg = random.normal(random.PRNGKey(0), (10000, 100))
s = samplers.build_sgld_sampler(1e-5, G.loglikelihood, G.logprior, (Xg,), 1000, pbar=False)
eng = s.engine
keys, _ = keys_to_device(random.split(random.PRNGKey(0), 1))
st = eng.new_state(torch.zeros(1, 100, device="cuda"))
eng.init(st, keys, sampler_mode=True)
samples = torch.empty((1, K, 100), device="cuda")
eng.run(st, 0, K, samples)
buf = (C.c_ulonglong * 32)()
lib.sgmcmc_debug_phase_cycles(buf, 1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); eng.run(st, K, K, samples); b.record(); torch.cuda.synchronize()
lib.sgmcmc_debug_phase_cycles(buf, 0)
print(f"shape {os.environ.get('SGMCMC_CLUSTER')}x{os.environ.get('SGMCMC_WARPS_PER_CHAIN')}: {a.elapsed_time(b) * 1000 / K:.3f} us per step (instrumented)")
names = {0: "carry split", 1: "expand_step_keys", 2: "diffuse", 3: "randint_plan", 4: "row_pass total", 10: "  index draw",
         11: "  group loop (shuffles, loads, adds)", 12: "fold: first barrier", 5: "fold: W loop", 6: "fold: cluster.sync",
         7: "fold: DSMEM sum + barrier", 8: "finish_from_sums", 13: "estimate total", 14: "transition total", 9: "sample write"}
for k in (0, 1, 2, 3, 4, 10, 11, 12, 5, 6, 7, 8, 13, 14, 9):
    print(f"{names[k]:40s} {buf[k] / K:8.0f} cycles per step")
