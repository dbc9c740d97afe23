"""Data-sharded mode over NCCL: each rank holds N/W rows of the C2-shaped table; compare with the unsharded sampler.
torchrun --nproc-per-node 2 tools/sharded_2gpu.py"""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, "/root/repo")
from sgmcmcjax_b200 import kernels, parallel, random, samplers
from sgmcmcjax_b200.models import logistic_regression as L
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, D, B, C, K = 200_000, 100, 10_000, 256, 30
theta, X, y = L.gen_data(random.PRNGKey(42), D, N)            # same table on every rank (deterministic generator)
lo, hi = parallel.chain_shard(N, rank, world)                 # contiguous row shard
th = theta.cpu().numpy()
keys = random.split(random.PRNGKey(0), C)
th0 = np.tile(th, (C, 1)).astype(np.float32)
for name, builder, extra in (("sgld", kernels.build_sgld_kernel, ()), ("sgldCV", kernels.build_sgldCV_kernel, (th,))):
    s = parallel.build_data_sharded_sampler(builder, 1e-5, L.loglikelihood, L.logprior, (X, y), B, *extra,
                                            shards=[(lo, (X[lo:hi], y[lo:hi]))], n_total=N)
    out = s(keys, K, th0)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    out = s(keys, K, th0)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    ref_s = getattr(samplers, f"build_{name}_sampler")(1e-5, L.loglikelihood, L.logprior, (X, y), B, *extra, pbar=False)
    ref = ref_s(keys, K, th0)
    err = np.abs(out - ref).max() / np.abs(ref).max()
    gath = [None] * world
    dist.all_gather_object(gath, float(np.abs(out).sum()))
    if rank == 0:
        print(f"{name}: sharded over {world} ranks vs unsharded: max rel err {err:.2e}; identical on all ranks: {len(set(gath)) == 1}; "
              f"{dt / K * 1e6:.0f} us per step for {C} chains ({C * K / dt:.0f} chain-steps/s)", flush=True)
dist.destroy_process_group()
