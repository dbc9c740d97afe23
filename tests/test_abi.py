"""The C-ABI library loads on a CPU-only box and exports every symbol include/*.h declares."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(?:int|int64_t|const char\*)\s+(sgmcmc_\w+)\s*\(", src)
    return sorted(set(names))


def test_library_builds_loads_and_exports_header_symbols():
    from sgmcmcjax_b200 import _build, _native

    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    fns = declared_functions()
    assert len(fns) >= 17
    for name in fns:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert sorted(_native.EXPORTS) == fns
    lib.sgmcmc_abi_version.restype = ctypes.c_int32
    assert lib.sgmcmc_abi_version() == _native.ABI_VERSION


def test_ctypes_struct_layout_matches_header():
    from sgmcmcjax_b200 import _native

    # sgmcmc_config: 4*i32, i64, 2*i32, 8*i32, 4*i32, 8*f32, 8*f32, 4*f32, 8*i32
    assert ctypes.sizeof(_native.Config) == 16 + 8 + 8 + 32 + 16 + 32 + 32 + 16 + 32
    assert ctypes.sizeof(_native.State) == 8 + 8 * 8
    assert _native.Config.n_data.offset == 16 and _native.Config.leaf_size.offset == 32


def test_argument_validation_without_gpu():
    """Entry points reject bad arguments before touching the device (no compute on CPU boxes)."""
    from sgmcmcjax_b200 import _native

    lib = _native.load()
    cfg = _native.Config()
    cfg.abi_version = 99
    h = ctypes.c_void_p(0)
    assert lib.sgmcmc_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"ABI" in lib.sgmcmc_last_error()
    cfg.abi_version = _native.ABI_VERSION
    cfg.family = 99
    assert lib.sgmcmc_create(ctypes.byref(cfg), ctypes.byref(h)) == -2
    cfg.family, cfg.n_data, cfg.batch, cfg.n_leaves, cfg.data_dim = _native.FAMILY_PMF, 10, 2, 2, 2
    assert lib.sgmcmc_create(ctypes.byref(cfg), ctypes.byref(h)) == -1      # leaf sizes / dims missing
    assert b"leaf" in lib.sgmcmc_last_error() or b"pmf" in lib.sgmcmc_last_error()
    cfg.family, cfg.data_dim, cfg.n_leaves = _native.FAMILY_MLP, 4, 2
    cfg.leaf_size[0], cfg.leaf_size[1] = 12, 3
    cfg.dims[0], cfg.dims[1], cfg.dims[2] = 1, 4, 200                       # width 200 > 128
    assert lib.sgmcmc_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"mlp" in lib.sgmcmc_last_error()
    assert lib.sgmcmc_ksd_imq_partial(None, None, 0, 0, 0, 0, None, None) == -1


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sgmcmcjax_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
