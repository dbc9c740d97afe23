"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in sgmcmcjax_b200/parallel.py:
chain sharding + all-gather of samples, KSD tile partition + the single all-reduce of a double.
The per-rank "compute" is done by the oracle here (no CUDA on this box); on the GPU box the same
partition functions drive the CUDA kernels (bench.py --gpus N, ksd.imq_KSD(distributed=True))."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from oracle import jax_prng as P
    from oracle import ksd as OK
    from oracle import models as M
    from oracle import sgmcmc as S
    from sgmcmcjax_b200 import parallel

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- chains: 5 chains over 2 ranks, chain c uses split(key, 5)[c] whatever the world size
        n_chains, steps, d = 5, 4, 3
        X = P.normal(P.PRNGKey(0), (200, d))
        keys = P.split(P.PRNGKey(9), n_chains)
        lo, hi = parallel.chain_shard(n_chains, rank, world)
        kern = S.build_kernel("sgld", M.GaussianMean(), (X,), 10, 1e-4)
        local = np.stack([S.run_sampler(kern, keys[c], steps, [np.zeros(d, np.float32)])[0][0] for c in range(lo, hi)])
        full = parallel.gather_chains(torch.from_numpy(local), n_chains).numpy()
        # ---- KSD: tiles dealt round-robin, one all-reduce of a double
        rng = np.random.default_rng(1)
        n = 300
        Xs = rng.normal(size=(n, 4)); G = -Xs + 0.1 * rng.normal(size=(n, 4))
        part = 0.0
        for I, J in parallel.ksd_tile_partition(n, rank, world):
            r0, r1 = I * parallel.KSD_TILE, min(n, (I + 1) * parallel.KSD_TILE)
            c0, c1 = J * parallel.KSD_TILE, min(n, (J + 1) * parallel.KSD_TILE)
            blk = sum(OK.k0_rows(Xs[i], Xs[c0:c1], G[i], G[c0:c1]).sum() for i in range(r0, r1))
            part += blk * (1.0 if I == J else 2.0)
        total = parallel.allreduce_scalar(torch.tensor([part], dtype=torch.float64))
        # ---- K2 with the rows sharded: local grad_log_post (prior included once per rank) + one all-reduce
        from sgmcmcjax_b200.models import logistic_regression as L
        rngl = np.random.default_rng(2)
        Xl = rngl.normal(size=(101, 6)).astype(np.float32); yl = (rngl.random(101) < 0.5).astype(np.int32)
        thetas = rngl.normal(size=(3, 6)).astype(np.float32)
        lo_r, hi_r = parallel.chain_shard(101, rank, world)          # contiguous row shard
        fam = M.LogisticRegression()
        # grad_log_post over the local rows with Ndata = local rows: prior + sum of the local per-row gradients
        local = np.stack([fam.grad_log_post([t], (Xl[lo_r:hi_r], yl[lo_r:hi_r]), hi_r - lo_r, np.float64)[0] for t in thetas])
        spec = L.loglikelihood._sgmcmc_family
        prior = parallel.prior_grad(spec, [6], torch.from_numpy(thetas.astype(np.float64)))
        k2 = parallel.sharded_fullbatch_grad(torch.from_numpy(local), prior).numpy()
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), full=full, total=total.numpy(), k2=k2)
    finally:
        dist.destroy_process_group()


def test_chain_shard_partitions_chains():
    from sgmcmcjax_b200 import parallel

    for n, w in [(1024, 8), (5, 2), (3, 8), (1, 1), (1000, 3)]:
        blocks = [parallel.chain_shard(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.chain_shard(4, 2, 2)


def test_ksd_tile_partition_covers_upper_triangle_once():
    from sgmcmcjax_b200 import parallel

    for n, nparts in [(300, 2), (1000, 3), (128, 4), (50_000, 8)]:
        T = (n + 127) // 128
        seen = [t for p in range(nparts) for t in parallel.ksd_tile_partition(n, p, nparts)]
        assert len(seen) == T * (T + 1) // 2 == parallel.ksd_tile_count(n)
        assert sorted(seen) == [(i, j) for i in range(T) for j in range(i, T)]
        sizes = [len(parallel.ksd_tile_partition(n, p, nparts)) for p in range(nparts)]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_sharding_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    from oracle import jax_prng as P
    from oracle import ksd as OK
    from oracle import models as M
    from oracle import sgmcmc as S

    port = _free_port()
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    r0, r1 = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    assert np.array_equal(r0["full"], r1["full"]) and np.array_equal(r0["total"], r1["total"])
    X = P.normal(P.PRNGKey(0), (200, 3))
    keys = P.split(P.PRNGKey(9), 5)
    kern = S.build_kernel("sgld", M.GaussianMean(), (X,), 10, 1e-4)
    want = np.stack([S.run_sampler(kern, keys[c], 4, [np.zeros(3, np.float32)])[0][0] for c in range(5)])
    assert np.array_equal(r0["full"], want)                      # result independent of the world size
    rng = np.random.default_rng(1)
    Xs = rng.normal(size=(300, 4)); G = -Xs + 0.1 * rng.normal(size=(300, 4))
    ref = float(OK.imq_ksd_sum(Xs, G))
    assert abs(float(r0["total"][0]) - ref) <= 1e-10 * abs(ref)
    rngl = np.random.default_rng(2)
    Xl = rngl.normal(size=(101, 6)).astype(np.float32); yl = (rngl.random(101) < 0.5).astype(np.int32)
    thetas = rngl.normal(size=(3, 6)).astype(np.float32)
    want = np.stack([M.LogisticRegression().grad_log_post([t], (Xl, yl), 101, np.float64)[0] for t in thetas])
    np.testing.assert_allclose(r0["k2"], want, rtol=1e-6, atol=1e-6)
    assert np.array_equal(r0["k2"], r1["k2"])
