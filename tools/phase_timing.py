"""Per-phase cycles of the MLP sampler.  Build the instrumented library first:
   python tools/variant.py timing --only capi.cu,big_inst_mlp.cu -DSG_PHASE_TIMING -DSG_NO_VEC -DSG_NO_PMF
(it is loaded from build/variants/lib_timing.so through SGMCMC_B200_LIB)."""
import ctypes as C, os, subprocess, sys
sys.path.insert(0, "/root/repo")
os.environ.setdefault("SGMCMC_B200_LIB", "/root/repo/build/variants/lib_timing.so")
from sgmcmcjax_b200 import _native
lib = _native.load()
import runpy
buf = (C.c_ulonglong * 32)()
sys.argv = ["prof_big.py", sys.argv[1] if len(sys.argv) > 1 else "mlp_psgld", "148", "4", "3"]
lib.sgmcmc_debug_phase_cycles(buf, 1)
runpy.run_path("/root/repo/tools/prof_big.py", run_name="__main__")
lib.sgmcmc_debug_phase_cycles(buf, 0)
names = ["diffuse passes", "layer-1 forward GEMM", "small layers (softmax, layer 2, dZ1)", "layer-1 backward GEMM",
         "  softmax(z1)", "  softmax + z2 forward", "  output layer dz", "  dW2, db2", "  dW2, db2 + quad dZ1 + db1",
         "  fwd stager: wait stage free (MMA c-2)", "  fwd stager: mask + split + st.shared (incl. wait for the loads)",
         "  fwd stager: fence.proxy.async", "  fwd stager: barrier arrive", "  fwd issuer: MMA issue + commit",
         "  fwd stager: issue loads c+2", "  fwd issuer: wait for the stage",
         "  dZ1 quad: dA sweep", "  dZ1 quad: softmax dot", "  dZ1 quad: operand stores + db1 shuffles", "  dZ1 quad: barrier", "  dZ1 quad: db1 fold",
         "  bwd thread 0: wait stage free", "  bwd thread 0: mask + split + st.shared (incl. wait for the load)", "  bwd thread 0: fence.proxy.async",
         "  bwd thread 0: hand-over barrier", "  bwd thread 0: issue load u+4", "  bwd thread 0: MMA issue + commit",
         "  bwd: wait for the last MMAs", "  bwd: epilogue (TMEM -> gW1)", "  bwd: barrier after the epilogue"]
steps = 3 * 4
for n, v in zip(names, buf):
    print(f"{n:40s} {v / steps / 1.965e3:8.1f} us per chain-step (block 0)")
