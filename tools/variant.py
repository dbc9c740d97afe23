"""Build an experiment variant of the library under build/variants/ (git-ignored, shipped to the GPU box):
   python tools/variant.py NAME [--only capi.cu,big_inst_pmf.cu] -DFOO=1 ...   ->  build/variants/lib_NAME.so
--only limits the translation units (the C ABI in capi.cu is always needed; a family whose unit is left out must be
stubbed with -DSG_NO_VEC / -DSG_NO_MLP / -DSG_NO_PMF).  Load with SGMCMC_B200_LIB=$PWD/build/variants/lib_NAME.so and
delete build/variants when done."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sgmcmcjax_b200 import _build as B
name, args = sys.argv[1], sys.argv[2:]
sources = None
if args and args[0] == "--only":
    sources, args = args[1].split(","), args[2:]
out = os.path.join(B.REPO_DIR, "build", "variants", f"lib_{name}.so")
t = time.time()
B.build(force=True, defs=tuple(args), out=out, sources=sources)
print(out, f"{time.time() - t:.0f} s")
