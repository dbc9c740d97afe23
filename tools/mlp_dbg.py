"""Debug helper: layer-1 gradient of the MLP family (tensor-core path) vs the oracle, element pattern."""
import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tests")
np.set_printoptions(linewidth=250, precision=4, suppress=True)
from test_gpu_big_families import mlp_case, mlp_params, build_pair, flat, tree_leaves
from oracle import jax_prng as P
sizes = tuple(int(v) for v in sys.argv[1].split(",")) if len(sys.argv) > 1 else (20, 12, 5)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 16
fam, fns, data, th0 = mlp_case(sizes, n, seed=3)
ok, (init_fn, kernel, get_params) = build_pair("sgld", 1e-6, {}, fam, fns, data, batch, th0, mlp_params)
init_fn.engine.flatten_params(mlp_params(th0), None)
thetas = flat(th0)[None]
got = init_fn.engine.fullbatch_grad(torch.from_numpy(thetas).cuda()).cpu().numpy()[0]
ref = flat(fam.grad_log_post(th0, data, n))
s0, s1 = sizes[0], sizes[1]
gw, rw = got[: s0 * s1].reshape(s1, s0), ref[: s0 * s1].reshape(s1, s0)
# remove the prior part (-w) to look at the likelihood sum alone
w1 = th0[0]
print("max|ref| W1", np.abs(rw).max(), "max err W1", np.abs(gw - rw).max(), " rest err", np.abs(got[s0*s1:] - ref[s0*s1:]).max())
print("got lik part (rows n, cols k):\n", (gw + w1)[:8, :24])
print("ref lik part:\n", (rw + w1)[:8, :24])
bad = np.argwhere(np.abs(gw - rw) > 1e-3 * np.abs(rw).max())
print("bad count", len(bad), "of", gw.size, "first", bad[:10].tolist())
