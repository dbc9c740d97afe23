"""Build an experiment variant of the library under build/variants/ (git-ignored, shipped to the GPU box):
   python tools/variant.py NAME -DSG_NO_VEC -DSG_NO_MLP -DFOO=1 ...   ->  build/variants/lib_NAME.so
Load it with SGMCMC_B200_LIB=$PWD/build/variants/lib_NAME.so.  Delete build/variants when done."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgmcmcjax_b200 import _build as B
name, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(B.REPO_DIR, "build", "variants", f"lib_{name}.so")
os.makedirs(os.path.dirname(out), exist_ok=True)
t = time.time()
subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *defs, "-I", os.path.join(B.REPO_DIR, "include"), "-o", out,
                *[os.path.join(B.CSRC, s) for s in B.SOURCES]], check=True)
print(out, f"{time.time() - t:.0f} s")
