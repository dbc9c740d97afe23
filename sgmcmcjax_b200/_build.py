"""Build libsgmcmc_b200.so in-tree with nvcc for sm_100a (no JIT cache, no torch extension)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_DIR = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libsgmcmc_b200.so")
# capi.cu: the C ABI, KSD and K2 kernels; the sampler kernels' template instantiations are spread over the other
# translation units so that they compile in parallel (one nvcc -c each), then one link.
SOURCES = ["capi.cu", "vec_inst_gauss.cu", "vec_inst_logreg.cu", "vec_inst_logreg_svrg.cu", "vec_inst_logreg_cv.cu",
           "big_inst_mlp.cu", "big_inst_pmf.cu"]
HEADERS = ["prng.cuh", "vec_sampler.cuh", "vec_launch.cuh", "ksd.cuh", "ksd_tc.cuh", "big_sampler.cuh", "big_launch.cuh",
           "mlp_grad.cuh", "k2_gemm.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(REPO_DIR, "build", "obj")


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libsgmcmc_b200.so cannot be built")


def _deps():
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(REPO_DIR, "include", "sgmcmc_b200.h"))
    return deps


def toolchain_missing(exc: BaseException) -> bool:
    """True if a build failed because there is no nvcc at all (as opposed to nvcc rejecting the sources)."""
    return isinstance(exc, RuntimeError) and "nvcc not found" in str(exc)


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False, defs=(), out: str = LIB_PATH, sources=None) -> str:
    """Compile the C-ABI library if it is missing or older than its sources: every translation unit with its own
    `nvcc -c` (in parallel), then `nvcc -shared`.  `defs` / `out` / `sources`: experiment builds (tools/variant.py)."""
    import fcntl

    if out == LIB_PATH and not defs and not force and not is_stale():
        return LIB_PATH
    _nvcc()                                       # fail before taking the lock if there is no toolchain
    os.makedirs(OBJ_DIR, exist_ok=True)
    # one builder at a time per output (every torchrun rank sees the same stale library at once); the others wait on
    # the lock and then find the library fresh
    with open(os.path.join(OBJ_DIR, "." + os.path.basename(out) + ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if out == LIB_PATH and not defs and not force and not is_stale():
            return LIB_PATH
        return _build_locked(verbose, defs, out, sources)


def _build_locked(verbose, defs, out, sources) -> str:
    from concurrent.futures import ThreadPoolExecutor

    sources = list(sources or SOURCES)
    tag = "default" if out == LIB_PATH else os.path.splitext(os.path.basename(out))[0]
    obj_dir = os.path.join(OBJ_DIR, tag)
    os.makedirs(obj_dir, exist_ok=True)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]
    if verbose:
        flags += ["-Xptxas", "-v"]

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, os.path.splitext(src)[0] + ".o")
        cmd = [_nvcc(), *flags, *defs, "-I", os.path.join(REPO_DIR, "include"), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.run(cmd, check=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(sources), os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources))
    tmp = f"{out}.tmp.{os.getpid()}"              # link beside the target, then rename: no reader sees a half-written .so
    subprocess.run([_nvcc(), "-shared", "-o", tmp, *objs], check=True)
    os.replace(tmp, out)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
