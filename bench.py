"""bench.py - SGLD chain-steps/sec on logistic regression (BASELINE.json metric), plus imq_KSD pairs/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): logistic regression N=1e6 rows, D=100, minibatch 1e4, SGLD
(dt=1e-5), 1024 independent chains IN TOTAL, sharded over the GPUs (chains are the unit that shards:
rank r owns chain_shard(1024, r, N), 128 chains per GPU at N=8; data replicated; no collective on the
data path; scaling "strong").  One bench "step" = one launch of the fused sampler kernel advancing every
chain by --inner SGLD iterations (threefry index draw + random row gather + per-example gradients +
reduction + diffusion update with in-register Gaussian noise, samples written to HBM).  At N > 1 the
same run with 1024 chains PER GPU is reported beside it as "weak" (zero-communication weak scaling is
linear by construction; it is context, not the headline).

value      chain-steps/s over all GPUs, inputs resident in HBM, CUDA-event timed, max over ranks.
e2e        same metric through the public API `sampler(key, Nsamples, params)` with HOST (numpy) keys and
           parameters in and HOST samples out (H2D + D2H inside the timed region).
roofline   algorithmic bytes (B*(row+label bytes) + 8P + 4P per chain-step) / kernel time
           against the measured HBM copy bandwidth of MEASURED_PEAKS.json.
cpu_baseline  the numpy oracle (restatement of the reference's JAX semantics; JAX itself is not installable
           here) on the host cores, one chain per worker process, bounded sample.
ksd        imq_KSD on N=50k x D=100 (configs[4]) in pairs/s; row blocks shard over ranks, one all-reduce.
checks     (N > 1, outside every timed region) one chain of every rank re-run on rank 0 with the same key and launch
           shape must be bit-identical; the distributed KSD sum must equal the single-rank sum to 1e-9.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_DATA, DIM, BATCH, CHAINS, DT = 1_000_000, 100, 10_000, 1024, 1e-5
KSD_N = 50_000
METRIC = "SGLD chain-steps/sec, logistic reg. D=100"
UNIT = "chain-steps/s"


def algorithmic_bytes_per_chain_step(batch=BATCH, dim=DIM, n_state=1):
    # B*(row_bytes(X)+row_bytes(y)) + 4P*(2 n_state) + 4P (sample emitted)
    return batch * (4 * dim + 4) + 4 * dim * 2 * n_state + 4 * dim


def tf32_peak_tflops(bf16_peak):
    """Dense TF32 tensor peak for the KSD roofline: the measured tcgen05 kind::tf32 rate committed under profiles/
    (tools/tf32_peak.py) if present, else half the measured bf16 rate as a proxy."""
    try:
        v = json.load(open(os.path.join(ROOT, "profiles", "tf32_peak.json")))
        return float(v["tf32_tflops"]), f"measured tcgen05 kind::tf32 peak ({v['how']})"
    except Exception:
        return 0.5 * bf16_peak, "0.5 x measured bf16 (no TF32 peak was measured)"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ CPU arm
def _cpu_worker(args):
    """One oracle chain on the shared synthetic table; returns (chain_steps, seconds)."""
    seed, n_steps, budget_s = args
    from oracle import jax_prng as P
    from oracle import models as M
    from oracle import sgmcmc as S

    X, y, theta = _CPU_DATA
    kern = S.build_kernel("sgld", M.LogisticRegression(), (X, y), BATCH, DT)
    key = P.PRNGKey(seed)
    key, sub = P.split(key)
    state = kern.init_fn(sub, [theta])
    t0 = time.perf_counter()
    done = 0
    for i in range(n_steps):
        key, sub = P.split(key)
        state = kern.kernel(i, sub, state)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return done, time.perf_counter() - t0


_CPU_DATA = None


def cpu_data():
    """Synthetic table of the C2 shape for the host arm (numpy RNG: content does not change the cost)."""
    global _CPU_DATA
    if _CPU_DATA is None:
        rng = np.random.default_rng(42)
        theta = (np.sqrt(10.0) * rng.standard_normal(DIM)).astype(np.float32)
        X = rng.standard_normal((N_DATA, DIM), dtype=np.float32)
        y = (rng.random(N_DATA) < 0.5 * (1.0 + np.tanh(0.5 * (X @ theta)))).astype(np.int32)      # sigmoid without overflow
        _CPU_DATA = (X, y, theta)
    return _CPU_DATA


def run_cpu_arm(budget_s: float, max_steps: int = 1_000_000):
    """Oracle chains on every host core (fork before CUDA is initialised; table shared copy-on-write)."""
    import multiprocessing as mp

    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)                      # one chain per core, no BLAS oversubscription
    except Exception:
        pass
    cores = len(os.sched_getaffinity(0))
    cpu_data()
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(s, max_steps, budget_s) for s in range(cores)])
    wall = time.perf_counter() - t0
    steps = sum(r[0] for r in res)
    slowest = max(r[1] for r in res)
    return {"value": steps / slowest, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{cores} oracle chains (one per core), {steps} chain-steps of the C2 workload in "
                      f"{slowest:.1f} s (numpy restatement of SGMCMCJax/JAX-0.4.13 semantics, not XLA)",
            "wall_s": wall, "chain_steps": steps}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_step_budget = 6.0
    vals, times = [], []
    for s in range(args.warmup + args.steps):
        r = run_cpu_arm(per_step_budget if s >= args.warmup else 2.0)
        if s >= args.warmup:
            vals.append(r["value"])
            times.append(r["wall_s"])
    v = float(statistics.mean(vals))
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(statistics.mean(times)), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "note": "reference arm = numpy oracle on host cores; the "
                   "reference itself needs jax==0.4.13, which is not installable in this image"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_name(args):
    return (f"logreg N={N_DATA} D={DIM} batch={args.batch} SGLD dt={DT}, {args.chains} chains in total sharded over the "
            f"GPUs, {args.inner} chain-steps per launch")


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1]
        if not rows and self.rows:       # timed region shorter than the sampling period: the samples nearest to it
            mid = 0.5 * (t0 + t1)
            rows = [r for _, r in sorted(self.rows, key=lambda tr: abs(tr[0] - mid))[:3]]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}



# ------------------------------------------------------------------------------------------ other configs
def _time_engine(torch, eng, lib, keys_host, theta0, K, steps, warmup, with_samples=True):
    """Device-resident timing of `steps` launches of K chain transitions each; returns (ms per launch, launches)."""
    from sgmcmcjax_b200._engine import keys_to_device

    keys_dev, _ = keys_to_device(keys_host)
    st = eng.new_state(theta0)
    eng.init(st, keys_dev, sampler_mode=True)
    C = theta0.shape[0]
    samples = torch.empty((C, K, eng.P), dtype=torch.float32, device="cuda") if with_samples else None
    i0 = 0
    for _ in range(warmup):
        eng.run(st, i0, K, samples); i0 += K
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.sgmcmc_launch_count()
    a.record()
    for _ in range(steps):
        eng.run(st, i0, K, samples); i0 += K
    b.record()
    torch.cuda.synchronize()
    assert torch.isfinite(st["theta"]).all(), "non-finite chain state"
    return a.elapsed_time(b) / steps, lib.sgmcmc_launch_count() - l0


def extra_workloads(args, torch, lib, hbm_peak, X, y, theta_true):
    """The other BASELINE.json configs on one GPU (parity-tested in tests/, reported here for SURVEY.md 8d):
    C1 README Gaussian (latency), C2 at batch 1e3 and with SGLD-CV, C3 Bayesian MLP, C4 PMF."""
    from sgmcmcjax_b200 import random, samplers
    from sgmcmcjax_b200.models import bayesian_mlp as B
    from sgmcmcjax_b200.models import gaussian_mean as G
    from sgmcmcjax_b200.models import logistic_regression as L
    from sgmcmcjax_b200.models import pmf as F

    out = []

    def report(name, ms, C, K, grads_per_step, bytes_per_chain_step, launches, note=""):
        cs = C * K * grads_per_step / (ms * 1e-3)
        gbs = bytes_per_chain_step * cs / 1e9
        out.append({"config": name, "value": cs, "unit": UNIT, "ms_per_launch": ms, "chains": C, "steps_per_launch": K,
                    "algorithmic_bytes_per_chain_step": bytes_per_chain_step, "achieved_gbs": gbs,
                    "hbm_frac": gbs / hbm_peak, "gpu_launches": int(launches), "note": note})

    th = theta_true.cpu().numpy()
    C = args.chains
    keys = random.split(random.PRNGKey(0), C)
    th0 = torch.from_numpy(np.tile(th, (C, 1)).astype(np.float32)).cuda()
    # ---- C2 variants
    s = samplers.build_sgld_sampler(DT, L.loglikelihood, L.logprior, (X, y), 1000, pbar=False)
    ms, nl = _time_engine(torch, s.engine, lib, keys, th0, 100, 5, 3)
    report("C2 logreg SGLD batch=1000 (BASELINE metric string)", ms, C, 100, 1, algorithmic_bytes_per_chain_step(1000), nl)
    s = samplers.build_sgldCV_sampler(1e-4, L.loglikelihood, L.logprior, (X, y), BATCH, th, pbar=False)
    ms, nl = _time_engine(torch, s.engine, lib, keys, th0, 10, 5, 3)
    report("C2 logreg SGLD-CV batch=10000", ms, C, 10, 1, algorithmic_bytes_per_chain_step(BATCH) + 8 * DIM, nl)
    # ---- C5 front half: the `grads` argument of imq_KSD = full-batch grad_log_post at the sample points (K2 on tensor cores)
    Mk = args.ksd_n
    pts = (theta_true[None, :] + 0.01 * random.normal(random.PRNGKey(7), (Mk, DIM))).contiguous()
    s.engine.fullbatch_grad(pts)                                                     # operand prep (once per table), scratch, warm-up
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.sgmcmc_launch_count()
    a.record(); gk = s.engine.fullbatch_grad(pts); b.record(); torch.cuda.synchronize()
    msk = a.elapsed_time(b)
    assert torch.isfinite(gk).all()
    out.append({"config": f"C5 grads: full-batch grad_log_post at {Mk} sample points, logreg N={N_DATA} D={DIM} (tcgen05 3xTF32)",
                "value": Mk * float(N_DATA) / (msk * 1e-3), "unit": "(theta,row) pairs/s", "ms": msk,
                "algorithmic_tflops": 4.0 * Mk * N_DATA * DIM / (msk * 1e-3) / 1e12,
                "gpu_launches": int(lib.sgmcmc_launch_count() - l0)})
    del s, pts, gk
    # ---- C1 README Gaussian: one chain, latency-bound (4 MB table lives in L2)
    Xg = random.normal(random.PRNGKey(0), (10000, 100))
    s = samplers.build_sgld_sampler(1e-5, G.loglikelihood, G.logprior, (Xg,), 1000, pbar=False)
    ms, nl = _time_engine(torch, s.engine, lib, random.split(random.PRNGKey(0), 1), torch.zeros(1, 100, device="cuda"),
                          1000, 5, 3)
    report("C1 README gaussian N=1e4 D=100 batch=1000, 1 chain", ms, 1, 1000, 1, 1000 * 400 + 1200, nl,
           note=f"latency-bound: {ms:.3f} us per chain-step")
    # ---- C3 Bayesian MLP 784-100-10, N=60k, batch=100, 256 chains
    Cm, sizes = 256, [784, 100, 10]
    Xm = torch.rand(60000, 784, device="cuda")
    ym = torch.nn.functional.one_hot(torch.randint(0, 10, (60000,), device="cuda"), 10).float()
    params = B.init_network(random.PRNGKey(1), sizes)
    Pm = sum(w.numel() + b.numel() for w, b in params)
    flat = torch.cat([t.reshape(-1) for wb in params for t in wb])
    thm = flat[None].repeat(Cm, 1).contiguous()
    keys_m = random.split(random.PRNGKey(2), Cm)
    row_bytes = 100 * (784 * 4 + 40)
    s = samplers.build_sghmc_sampler(1e-6, 4, B.loglikelihood, B.logprior, (Xm, ym), 100, pbar=False)
    s.engine.flatten_params(params, None)
    ms, nl = _time_engine(torch, s.engine, lib, keys_m, thm, 4, 3, 3)
    report("C3 bayesian MLP 784-100-10 SGHMC L=4 batch=100", ms, Cm, 4, 4, row_bytes + 16 * Pm + Pm, nl,
           note="one sample = 4 chain-steps (gradient evaluations)")
    s = samplers.build_psgld_sampler(1e-3, B.loglikelihood, B.logprior, (Xm, ym), 100, pbar=False)
    s.engine.flatten_params(params, None)
    ms, nl = _time_engine(torch, s.engine, lib, keys_m, thm, 16, 3, 3)
    report("C3 bayesian MLP 784-100-10 pSGLD batch=100", ms, Cm, 16, 1, row_bytes + 16 * Pm + 4 * Pm, nl)
    del s, Xm, ym, thm
    # ---- C4 PMF 943 x 1682, rank 20, N=1e5, batch=1000, 512 chains
    Cp, nu, ni, r, Np = 512, 943, 1682, 20, 100_000
    ij = torch.stack([torch.randint(0, nu, (Np,), device="cuda"), torch.randint(0, ni, (Np,), device="cuda")], 1).int()
    rt = torch.randint(1, 6, (Np,), device="cuda").float()
    U = 0.1 * random.normal(random.PRNGKey(3), (nu, r)); V = 0.1 * random.normal(random.PRNGKey(4), (ni, r))
    Pp = (nu + ni) * r
    thp = torch.cat([U.reshape(-1), V.reshape(-1)])[None].repeat(Cp, 1).contiguous()
    keys_p = random.split(random.PRNGKey(5), Cp)
    pmf_bytes = 1000 * 12 + 2 * 1000 * 4 * r + 16 * Pp + 4 * Pp
    s = samplers.build_sgnht_sampler(1e-5, F.loglikelihood, F.logprior, (ij, rt), 1000, pbar=False)
    s.engine.flatten_params((U, V), None)
    ms, nl = _time_engine(torch, s.engine, lib, keys_p, thp, 16, 3, 3)
    report("C4 PMF 943x1682 rank 20 SGNHT batch=1000", ms, Cp, 16, 1, pmf_bytes, nl)
    # control variates centred at a MAP estimate: 300 Adam iterations of the native optimiser on the same family
    from sgmcmcjax_b200 import optimizer as NO
    (Um, Vm), lps = NO.build_optax_optimizer(NO.adam(2e-2), F.loglikelihood, F.logprior, (ij, rt), 1000)(random.PRNGKey(6), 300, (U, V))
    lps = lps.float().cpu().numpy() if hasattr(lps, "cpu") else np.asarray(lps)
    # dt = 1e-4 here (SURVEY 8d quotes 1e-3 from the reference's 5-dimensional test): the reference's thermostat update
    # alpha += dt/2 (|v| - v.size) (diffusions.py:326,334) drifts by -dt/2 * 18860 per step on a leaf of 18,860 entries; at
    # dt = 1e-3 the friction turns into exp(+0.9) growth per step within ~100 steps and the chain overflows - a property
    # of that update rule on large leaves, reproduced faithfully.  The cost per step does not depend on dt.
    s = samplers.build_badodabCV_sampler(1e-4, F.loglikelihood, F.logprior, (ij, rt), 1000, (Um, Vm), pbar=False)
    thm_ = torch.cat([Um.reshape(-1), Vm.reshape(-1)])[None].repeat(Cp, 1).contiguous()
    ms, nl = _time_engine(torch, s.engine, lib, keys_p, thm_, 16, 3, 3)
    report("C4 PMF 943x1682 rank 20 BADODAB-CV batch=1000", ms, Cp, 16, 1, pmf_bytes + 8 * Pp, nl,
           note=f"centred at the MAP of 300 native Adam iterations (log-posterior {float(lps[:10].mean()):.4g} -> {float(lps[-10:].mean()):.4g})")
    return out

# ------------------------------------------------------------------------------------------ GPU arm
def _timed_launches(torch, dist, world, eng, lib, keys_host, theta0_host, K, steps, warmup):
    """Device-resident timing of `steps` launches of K chain-steps for this rank's chains (state in HBM, one launch per
    bench step), bracketed by barrier + synchronize; returns (max-over-ranks total ms, this rank's per-launch ms list,
    launches, final samples, state)."""
    from sgmcmcjax_b200._engine import keys_to_device

    C = keys_host.shape[0]
    keys_dev, _ = keys_to_device(keys_host)
    st = eng.new_state(torch.from_numpy(theta0_host).cuda())
    eng.init(st, keys_dev, sampler_mode=True)
    samples = torch.empty((C, K, eng.P), dtype=torch.float32, device="cuda")
    i0 = 0
    for _ in range(warmup):
        eng.run(st, i0, K, samples); i0 += K
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = lib.sgmcmc_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    t_wall0 = time.perf_counter()
    ev[0].record()
    for s in range(steps):
        eng.run(st, i0, K, samples); i0 += K
        ev[s + 1].record()
    torch.cuda.synchronize()
    t_wall1 = time.perf_counter()
    launches = lib.sgmcmc_launch_count() - l0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    per_launch_ms = [ev[s].elapsed_time(ev[s + 1]) for s in range(steps)]
    t = torch.tensor([ev[0].elapsed_time(ev[-1])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert torch.isfinite(samples).all(), "non-finite samples"
    return float(t.item()), per_launch_ms, int(launches), (t_wall0, t_wall1)


def _multi_gpu_checks(torch, dist, rank, world, lib, eng, sampler, all_keys, theta_row, shards, K, ksd_inputs):
    """Outside every timed region.  (1) Chain sharding: the first chain of every rank, run by that rank inside its shard,
    must equal the same key run on rank 0 - same launch shape, so bit for bit.  (2) KSD: the all-reduced tile sums must
    equal the single-rank sum to 1e-9 relative."""
    import ctypes as C

    from sgmcmcjax_b200 import ksd

    checks = {}
    lo, hi = shards[rank]
    w, cl = C.c_int32(0), C.c_int32(0)
    lib.sgmcmc_launch_shape(eng._handle, hi - lo, C.byref(w), C.byref(cl))
    th = np.tile(theta_row, (hi - lo, 1)).astype(np.float32)
    mine = sampler(np.ascontiguousarray(all_keys[lo:hi]), K, th)[0]                 # [K, P] of this rank's first chain
    mine_t = torch.from_numpy(np.ascontiguousarray(mine)).cuda()
    gathered = [torch.empty_like(mine_t) for _ in range(world)]
    dist.all_gather(gathered, mine_t)
    shape_t = torch.tensor([w.value, cl.value], dtype=torch.int32, device="cuda")
    shapes = [torch.empty_like(shape_t) for _ in range(world)]
    dist.all_gather(shapes, shape_t)
    if rank == 0:
        worst, exact = 0.0, True
        for r in range(world):
            rlo, rhi = shards[r]
            ws, cs = (int(v) for v in shapes[r].cpu())
            os.environ["SGMCMC_WARPS_PER_CHAIN"], os.environ["SGMCMC_CLUSTER"] = str(ws), str(cs)
            try:
                again = sampler(np.ascontiguousarray(all_keys[rlo:rlo + 1]), K, theta_row[None].astype(np.float32))[0]
            finally:
                del os.environ["SGMCMC_WARPS_PER_CHAIN"], os.environ["SGMCMC_CLUSTER"]
            got = gathered[r].cpu().numpy()
            exact = exact and bool(np.array_equal(got, again))
            worst = max(worst, float(np.abs(got - again).max() / max(np.abs(again).max(), 1e-30)))
        checks["chain_of_rank_r_equals_same_key_on_rank_0"] = {"bit_identical": exact, "max_rel_diff": worst,
                                                               "chains_checked": world, "steps": K}
        assert exact, f"sharded chains differ from the single-GPU run of the same keys (max rel diff {worst:.3e})"
    if ksd_inputs is not None:
        pts, grads = ksd_inputs
        n = int(pts.shape[0])
        part = ksd.imq_KSD_tiles(pts, grads, rank, world)
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
        if rank == 0:
            whole = ksd.imq_KSD_tiles(pts, grads, 0, 1)
            rel = float((part - whole).abs().item() / abs(whole.item()))
            checks["distributed_ksd_equals_single_rank"] = {"rel_diff": rel, "n": n, "bar": 1e-9}
            assert rel <= 1e-9, f"distributed KSD sum differs from the single-rank sum by {rel:.3e}"
    dist.barrier()
    return checks


def gpu_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_base = run_cpu_arm(args.cpu_budget)           # before CUDA init (fork-safe)
        for k in ("wall_s", "chain_steps"):
            cpu_base.pop(k, None)
        global _CPU_DATA
        _CPU_DATA = None

    import torch
    import torch.distributed as dist

    from sgmcmcjax_b200 import _native, ksd, parallel, random, samplers
    from sgmcmcjax_b200.models import logistic_regression as L

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _native.load()
    hbm_peak, bf16_peak, peak_src = peaks()

    # synthetic C2 data, generated on device with the jax-compatible PRNG (same table on every rank)
    theta_true, X, y = L.gen_data(random.PRNGKey(42), DIM, N_DATA)
    sampler = samplers.build_sgld_sampler(DT, L.loglikelihood, L.logprior, (X, y), args.batch, pbar=False)
    eng = sampler.engine
    CT, K, P = args.chains, args.inner, DIM
    shards = [parallel.chain_shard(CT, r, world) for r in range(world)]
    lo, hi = shards[rank]
    C = hi - lo                                                         # this rank's chains (128 of 1024 at N = 8)
    all_keys = random.split(random.PRNGKey(0), CT)                      # chain c of the job == key c, for any N
    keys_host = np.ascontiguousarray(all_keys[lo:hi])
    theta_row = theta_true.cpu().numpy().astype(np.float32)
    theta0_host = np.tile(theta_row, (C, 1)).astype(np.float32)

    # ---- KSD (configs[4]): 50k x 100 points, row blocks over ranks, one all-reduce of a double.  Timed FIRST: the call is
    # tensor-bound and loses 5-20 % when it runs right after the sampler legs (same clocks reported, board warm: 4.99 ms
    # alone, 5.2-5.9 ms after them, `tools/ksd_in_bench_probe.py`), and its roofline denominator is a burst measurement too.
    def run_ksd():
        n = args.ksd_n
        pts = theta_true[None, :] + 0.02 * random.normal(random.PRNGKey(7), (n, DIM))
        grads = -50.0 * (pts - theta_true[None, :]) + random.normal(random.PRNGKey(8), (n, DIM))
        pts, grads = pts.contiguous(), grads.contiguous()
        for _ in range(2):
            ksd.imq_KSD(pts, grads, distributed=world > 1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        a.record()
        for _ in range(reps):
            val = ksd.imq_KSD(pts, grads, distributed=world > 1)
        b.record()
        torch.cuda.synchronize()
        tk = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tk, op=dist.ReduceOp.MAX)
        ms = float(tk.item())
        pairs = float(n) * n
        # Flop accounting (corrected at the end of round 2).  Algorithmic: three full Gram matrices X X^T, G G^T, G X^T
        # (x_i.g_j is the transpose of g_i.x_j): 6 D N^2 flops for all N^2 ordered pairs.  What the kernel ISSUES: it works on
        # the upper triangle of 128 x 128 tiles, where a tile needs FOUR products - x.x, g.g and both cross products
        # g_I.x_J and x_I.g_J (the second would be tile (J, I) of G X^T) - each as 3 tf32 MMAs: per tile and 8-wide K slice
        # six 128 x 256 x 8 MMAs (csrc/ksd_tc.cuh PROD table), K padded from D to a multiple of 8 = 12 D' N^2 flops.
        # Earlier rounds put 1.5 x 6 D N^2 = 9 D N^2 here (three products per tile): 0.75 of what is issued.
        flops = 6.0 * DIM * pairs                                     # algorithmic, all N^2 ordered pairs
        T = (n + 127) // 128
        tiles = T * (T + 1) // 2
        k_slices = ((DIM + 7) // 8)
        mma_flops = float(tiles) * 6 * k_slices * (2.0 * 128 * 256 * 8)
        tf32_peak, tf32_src = tf32_peak_tflops(bf16_peak)
        issued = mma_flops / (ms * 1e-3) / 1e12                      # incl. the K padding (D -> multiple of 8) and whole diagonal tiles
        tf32_alg = 12.0 * DIM * pairs / (ms * 1e-3) / 1e12            # 4 products x 3 tf32 splits x 2 D flops on N^2 / 2 pairs
        ksd_line = {"metric": "imq_KSD pairs/sec", "value": pairs / (ms * 1e-3), "unit": "pairs/s", "n": n, "d": DIM,
                    "ms": ms, "ksd": float(val), "equiv_tflops": flops / (ms * 1e-3) / 1e12,
                    "method": "tcgen05 3xTF32 Gram tiles (upper triangle, dealt round-robin to ranks)",
                    "roofline": {"bound": "tensor", "achieved": tf32_alg, "peak": world * tf32_peak, "unit": "TFLOP/s",
                                 "frac": tf32_alg / (world * tf32_peak), "issued_tflops": issued,
                                 "mma_instructions": int(tiles) * 6 * k_slices, "tiles": int(tiles),
                                 "note": "achieved = 12 x D x N^2 tf32 flops / time: the upper triangle needs four products per "
                                         "pair (x.x, g.g, g_i.x_j, x_i.g_j), each as 3 tf32 split products; issued_tflops "
                                         "counts the tcgen05.mma instructions themselves (T(T+1)/2 tiles x ceil(D/8) K slices x "
                                         "6 MMAs of 128x256x8: K padding and whole diagonal tiles included).  The whole call is "
                                         "timed (centring + hi/lo split + tile kernel + all-reduce).  Rounds 1-2 reported "
                                         "9 x D x N^2 here (three products per tile instead of four): 0.75 of this figure.  "
                                         "peak = n_gpus x " + tf32_src}}
        return ksd_line, (pts, grads)

    ksd_line, ksd_inputs = None, None
    if not args.no_ksd and not args.ksd_last:
        ksd_line, ksd_inputs = run_ksd()

    # ---- device-resident timing: state lives in HBM, one launch per bench step
    clocks = ClockSampler(local) if rank == 0 else None      # already streaming when the timed region starts
    total_ms_max, per_launch_ms, launches, (t_wall0, t_wall1) = _timed_launches(
        torch, dist, world, eng, lib, keys_host, theta0_host, K, args.steps, args.warmup)
    clk = clocks.stop(t_wall0, t_wall1) if clocks else None
    value = CT * K * args.steps / (total_ms_max * 1e-3)

    import ctypes as ct
    w_, cl_ = ct.c_int32(0), ct.c_int32(0)
    lib.sgmcmc_launch_shape(eng._handle, C, ct.byref(w_), ct.byref(cl_))
    kern_ms = float(statistics.mean(per_launch_ms))
    alg_bytes = algorithmic_bytes_per_chain_step(args.batch) * C * K
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = args.traffic
    if traffic is None and args.batch == BATCH and C == CHAINS and K == 10:
        try:      # dram bytes per launch of the same launch shape, from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["vec_sampler_logreg_c2"]["dram_bytes_per_launch"]
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "kernel": "vec_sampler_kernel<LOGREG>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kern_ms,
                "note": f"rank 0's GPU: {C} chains per launch as {cl_.value} CTA(s) x {w_.value} warps per chain"}

    # ---- weak-scaling context at N > 1: 1024 chains on EVERY GPU (the round-1 headline)
    weak = None
    if world > 1 and not args.no_weak:
        wk = random.split(random.PRNGKey(1), CT * world)
        wt, _, _, _ = _timed_launches(torch, dist, world, eng, lib, np.ascontiguousarray(wk[rank * CT:(rank + 1) * CT]),
                                      np.tile(theta_row, (CT, 1)).astype(np.float32), K, args.steps, args.warmup)
        weak = {"value": world * CT * K * args.steps / (wt * 1e-3), "unit": UNIT, "chains_per_gpu": CT,
                "ms_per_step": wt / args.steps, "scaling": "weak"}

    # ---- end to end through the public API: host keys + host params in, host samples out
    Ke = args.e2e_steps
    sampler(keys_host, Ke, theta0_host)                               # warm-up (allocator, coefficient cache)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps, e2e_t = 3, []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = sampler(keys_host, Ke, theta0_host)                     # numpy in -> numpy out
        e2e_t.append(time.perf_counter() - t0)
    assert isinstance(out, np.ndarray) and out.shape == (C, Ke, P)
    te = torch.tensor([statistics.median(e2e_t)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    # Ke chain-steps are counted; the init_fn gradient (samplers.py:54) is extra uncounted work
    e2e = {"value": CT * Ke / float(te.item()), "unit": UNIT,
           "h2d_bytes_per_step": int((keys_host.nbytes + theta0_host.nbytes) * K / Ke),
           "d2h_bytes_per_step": int(C * K * P * 4),
           "note": f"sampler(key, {Ke}, params) with numpy in/out incl. init_fn; bytes per rank, scaled to one bench step "
                   f"({K} chain-steps per chain)"}

    if not args.no_ksd and args.ksd_last:
        ksd_line, ksd_inputs = run_ksd()

    checks = None
    if world > 1:
        checks = _multi_gpu_checks(torch, dist, rank, world, lib, eng, sampler, all_keys, theta_row, shards, K, ksd_inputs)

    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        del sampler, eng
        torch.cuda.empty_cache()
        extras = extra_workloads(args, torch, lib, hbm_peak, X, y, theta_true)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "l2": "data table 400 MB > 126 MB L2, rows drawn at random "
                       "every step (inputs larger than L2; no explicit flush)", "chains_total": CT,
                       "chains_per_gpu": [h_ - l_ for l_, h_ in shards],
                       "parallelism": f"{CT} chains sharded over {world} GPU(s), no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu_base, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clk, "checks": checks, "weak": weak, "ksd": ksd_line, "extras": extras,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--chains", type=int, default=CHAINS, help="chains IN TOTAL (sharded over the GPUs)")
    ap.add_argument("--no-weak", action="store_true", help="skip the 1024-chains-per-GPU context run at N > 1")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--inner", type=int, default=10, help="chain-steps per launch (= per bench step)")
    ap.add_argument("--e2e-steps", type=int, default=200,
                    help="Nsamples of the end-to-end sampler(key, Nsamples, params) call (its init_fn, launch and copy "
                         "set-up costs are inside the timed region and amortise over these steps)")
    ap.add_argument("--ksd-n", type=int, default=KSD_N)
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of host work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ksd", action="store_true")
    ap.add_argument("--ksd-last", action="store_true", help="time the KSD call after the sampler legs (the order of rounds 1-2)")
    ap.add_argument("--no-extras", action="store_true", help="skip the C1/C3/C4 and C2-variant lines (N=1 only)")
    ap.add_argument("--traffic", type=float, default=None,
                    help="dram bytes per launch from the committed ncu --set full capture (profiles/)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "native":
        print("note: fewer than 3 warm-up steps", file=sys.stderr)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "native" and world != args.gpus and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    if args.impl == "reference":
        return reference_arm(args)
    return gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
