"""Writes tests/golden/jax_prng_kat.json and tests/golden/oracle_regression.json.

Section "external" holds PUBLIC known answers (Random123 threefry2x32 KAT vectors; the
jax.random values for PRNGKey(0) / PRNGKey(42) quoted in SURVEY.md A.4) typed in verbatim -
they PIN the oracle.  Section "self_derived" is produced by running the oracle itself and only
freezes its behaviour (regression anchors; "parity unpinned", see oracle/__init__.py).  The
reference cannot be imported in this image (no jax), so no reference-run fixture exists.

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import jax_prng as P  # noqa: E402
from oracle import ksd as K  # noqa: E402
from oracle import models as M  # noqa: E402
from oracle import sgmcmc as S  # noqa: E402

external = {
    "threefry2x32": [
        {"key": [0, 0], "ctr": [0, 0], "out": [0x6B200159, 0x99BA4EFE]},
        {"key": [0xFFFFFFFF, 0xFFFFFFFF], "ctr": [0xFFFFFFFF, 0xFFFFFFFF], "out": [0x1CB996FC, 0xBB002BE7]},
        {"key": [0x13198A2E, 0x03707344], "ctr": [0x243F6A88, 0x85A308D3], "out": [0xC4923A9C, 0x483DF7A0]},
    ],
    "split": [
        {"key": [0, 0], "num": 2, "out": [[4146024105, 967050713], [2718843009, 1272950319]]},
        {"key": [0, 42], "num": 2, "out": [[2465931498, 3679230171], [255383827, 267815257]]},
    ],
    "uniform": [{"key": [0, 0], "shape": [], "out": [0.41845703]}],
    "normal": [
        {"key": [0, 0], "shape": [3], "out": [1.8160863, -0.48262316, 0.33988908]},
        {"key": [0, 0], "shape": [1], "out": [-0.20584226]},
        {"key": [2718843009, 1272950319], "shape": [1], "out": [-1.2515389]},
    ],
}

key0 = P.PRNGKey(0)
sub_init = P.split(key0)[1]
sub0 = P.split(P.split(key0)[0])[1]
k1, k2 = P.split(sub0)
self_derived = {
    "split": [
        {"key": [4146024105, 967050713], "num": 2, "out": P.split(np.array([4146024105, 967050713], np.uint32)).tolist()},
        {"key": [0, 42], "num": 3, "out": P.split(P.PRNGKey(42), 3).tolist()},
        {"key": [0, 7], "num": 5, "out": P.split(P.PRNGKey(7), 5).tolist()},
    ],
    "choice": [
        {"key": sub_init.tolist(), "n": 10000, "batch": 1000, "head": P.choice_indices(sub_init, 10000, 1000)[:8].tolist(),
         "sum": int(P.choice_indices(sub_init, 10000, 1000).sum())},
        {"key": k2.tolist(), "n": 10000, "batch": 1000, "head": P.choice_indices(k2, 10000, 1000)[:8].tolist(),
         "sum": int(P.choice_indices(k2, 10000, 1000).sum())},
        {"key": k2.tolist(), "n": 1000000, "batch": 10000, "head": P.choice_indices(k2, 10**6, 10**4)[:8].tolist(),
         "sum": int(P.choice_indices(k2, 10**6, 10**4).sum())},
        {"key": [0, 0], "n": 10000, "batch": 5, "head": P.choice_indices(key0, 10000, 5).tolist(),
         "sum": int(P.choice_indices(key0, 10000, 5).sum())},
        {"key": [0, 3], "n": 60000, "batch": 101, "head": P.choice_indices(P.PRNGKey(3), 60000, 101)[:8].tolist(),
         "sum": int(P.choice_indices(P.PRNGKey(3), 60000, 101).sum())},
        {"key": [0, 4], "n": 40000, "batch": 64, "head": P.choice_indices(P.PRNGKey(4), 40000, 64)[:8].tolist(),
         "sum": int(P.choice_indices(P.PRNGKey(4), 40000, 64).sum())},
    ],
    "readme_keys": {"init_subkey": sub_init.tolist(), "step0_subkey": sub0.tolist(), "k1": k1.tolist(), "k2": k2.tolist(),
                    "leaf0_noise_key": P.split(k1, 1)[0].tolist(),
                    "noise_head": P.normal(P.split(k1, 1)[0], (100,))[:4].tolist()},
    "readme_x_head": P.normal(key0, (10000, 100))[0, :4].tolist(),
}

with open(os.path.join(HERE, "jax_prng_kat.json"), "w") as f:
    json.dump({"external": external, "self_derived": self_derived}, f, indent=1)

# ---- small trajectory / KSD regression anchors (self-derived) -----------------------------------
X = P.normal(key0, (10000, 100))
kern = S.build_kernel("sgld", M.GaussianMean(), (X,), 1000, 1e-5)
out, _, _ = S.run_sampler(kern, key0, 3, [np.zeros(100, np.float32)])
ks = P.split(P.PRNGKey(42), 4)
th = (P.normal(ks[1], (10,)) * np.float32(np.sqrt(10))).astype(np.float32)
Xl = P.normal(ks[3], (500, 10))
y = (P.uniform(ks[0], (500,)) < 1 / (1 + np.exp(-(Xl @ th)))).astype(np.int32)
lr = M.LogisticRegression()
kern = S.build_kernel("sgld", lr, (Xl, y), 50, 1e-4)
o, _, _ = S.run_sampler(kern, P.PRNGKey(1), 40, [th])
G = np.stack([lr.grad_log_post([s], (Xl, y), 500)[0] for s in o[0]])
reg = {
    "readme_sgld_first_samples_head": out[0][:, :3].tolist(),
    "logreg_small": {"theta0": th.tolist(), "last_sample": o[0][-1].tolist(),
                     "ksd_fp64": float(K.imq_ksd_literal(o[0], G))},
}
with open(os.path.join(HERE, "oracle_regression.json"), "w") as f:
    json.dump(reg, f, indent=1)
print("golden files written")
