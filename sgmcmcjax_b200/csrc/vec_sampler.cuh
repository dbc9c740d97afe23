// K1: fused persistent sampler for vector-parameter families (gaussian_mean, logreg).
//
// One CTA owns one chain for all K steps of a launch; theta, the diffusion's auxiliary
// arrays, the carried gradient and the CV / SVRG centre live in shared memory.  Each step
// fuses, with no HBM round trip in between:
//   key tree          samplers.py:47-51, kernels.py:53-64,110-122, diffusion_util.py:66,91,112
//   diffusion update  diffusions.py:47-347 (in-register Gaussian noise, XLA erf_inv polynomial)
//   index draw        gradient_estimation.py:32,75,140 (threefry randint, bit-exact)
//   row gather        gradient_estimation.py:33,76,141 (warp-per-row float4 loads from HBM)
//   per-example grad  util.py:43-49 for the registered family, reduced with warp shuffles
//   CV / SVRG         gradient_estimation.py:72-85,134-151
//
// HBM data layout ("rows"): [n_data][DS] fp32, DS = D rounded up to 4.  gaussian_mean stores
// x; logreg stores x~ = (1 - 2y) x, i.e. the label is folded into the sign of the row, which
// is exactly the argument of the reference's loglik  -logsumexp([0, (1-2y) theta.x])
// (models/logistic_regression.py:65-67): d loglik / d theta = -sigmoid(theta.x~) x~.
//
// Few chains (fewer than the GPU has SMs, e.g. 128 chains per GPU when config 2's 1024 chains are split over 8 GPUs,
// or the README's single chain): a THREAD-BLOCK CLUSTER owns the chain.  The B-row index stream (and a full-batch
// pass) is split over the warps of all CTAs of the cluster; every CTA holds a full replica of the chain state in its
// own shared memory and runs the (tiny, D-element) key tree and diffusion update redundantly - same keys, same
// arithmetic, bit-identical replicas - so the ONLY exchange per gradient estimate is the [DS] partial row sum:
// every CTA folds its warps and PUSHES the result into slot recv[buf][its rank] of every CTA of the cluster (remote
// shared-memory stores), one cluster barrier, every CTA adds its own recv[buf][0 .. cs) from local shared memory in
// rank order (same order everywhere -> replicas stay identical).  recv is double-buffered, so one barrier per
// estimate is enough (a CTA can only reach the estimate after next once every peer has passed the next barrier,
// i.e. has finished reading this one).  Rank 0 alone writes samples and state back.
#pragma once
#include <cooperative_groups.h>
#include "prng.cuh"
#include "../../include/sgmcmc_b200.h"

namespace sg {

constexpr int MAXL = SGMCMC_MAX_LEAVES;
constexpr unsigned FULL = 0xffffffffu;

struct VecModel {
  int family, estimator, diffusion, flags;
  int n_leaves, leaf_dim, P, DS;
  unsigned n_data, batch;
  unsigned randint_mult;   // ((2^16 % n_data)^2 mod 2^32) % n_data: the multiplier of jax's randint (prng.cuh RandintPlan)
  int full_batch;          // standard estimator with batch == n_data (gradient_estimation.py:41-42)
  int leapfrog, update_rate;
  float scale;             // fp32(n_data) / fp32(batch)   (util.py:44, mean then * Ndata)
  float lik2[MAXL];        // gaussian_mean: 2 a_l
  float prior2[MAXL];      // gaussian_mean: 2 c_l ; logreg: [0] = prior precision
  float hyper[4];
  const float* rows;       // [own_n][DS]: rows own_lo .. own_lo + own_n - 1 of the table (data-sharded mode: a slice)
  const float* mu;         // gaussian_mean: [DS] column means the stored rows were centred by (nullptr: not centred).  The
                           //   gradient lik2 (sum x - B theta) is formed as lik2 (sum (x - mu) - B (theta - mu)): without the
                           //   shift, data whose mean is far from 0 cancels in fp32 at ~B |x| 2^-24 instead of ~sqrt(B) |x| 2^-24
  unsigned own_lo, own_n;  // own_lo = 0, own_n = n_data unless the rows are sharded over ranks
  const float* cv_center;  // [P]  (CV only)
  const float* cv_grad;    // [P]  full-batch gradient at cv_center
  int zc_f4;               // logreg SGLD-CV & co: float4 index of the row column holding center . x~ (CV == 2), else -1
};
// CV template parameter of the kernels: 0 = standard estimator; 1 = the centre's term is evaluated next to theta's from
// the centre in shared memory (SVRG: the centre moves, per chain; gaussian_mean: free); 2 = logreg with a FIXED centre
// (estimator CV): center . x~ is a per-row constant, written once into a spare column of the row table by
// sgmcmc_set_centering (zc_column_kernel) and simply read with the row - no second dot product per row.

struct VecRun {
  sgmcmc_state st;
  int i0, K;
  const sgmcmc_step_coef* coef;
  int coef_stride;
  float* samples;            // element (c, s, e) at samples[c * sample_chain_stride + s * P + e], or nullptr
  size_t sample_chain_stride;
  const uint32_t* step_keys; // [C][2]: kernel(i, key, state) mode (K == 1), else nullptr
  int mode;                  // 0 = run, 1 = init_fn with keys as given, 2 = init with sampler split
  const uint32_t* init_keys; // [C][2] for mode 1/2
  // data-sharded mode (rows split over ranks, one all-reduce of the [C, DS] minibatch sums per step, done by the host):
  const float* sum_in;       // [C][DS] all-reduced sums of the previous estimate: the gradient is finished from them first
  float* sum_out;            // [C][DS] this rank's partial sums of the new estimate (the gradient is NOT finished)
  float* lp_out;             // [C][K] log-posterior values of the optimizer iterations (SGMCMC_DIFF_ADAM_OPT), else nullptr
  int thin;                  // keep every thin-th state of the launch (0 / 1: all), sgmcmc_run_thinned
  float* moments;            // [C][3][P]: sum and sum of squares of (kept state - plane 2), or nullptr (EXT kernel)
  int shard_flags;           // SHARD_UPDATE2: run the palindrome's update2 (coef_prev) after finishing the gradient;
  const sgmcmc_step_coef* coef_prev;   //   SHARD_STEP: then do one transition and leave its partial sums in sum_out
};
constexpr int SHARD_UPDATE2 = 1, SHARD_STEP = 2;

// ---------------------------------------------------------------------------------------
// shared-memory carve-up (dynamic): all arrays padded to DSP = max(P, DS) rounded to 4
struct Smem {
  float *theta, *aux1, *aux2, *grad, *center, *fb, *S, *part;
  float* mu;          // [DS] column means of the centred row table (gaussian_mean), zeros otherwise
  float* recv;        // [2][cs][DS + 4]: folded partial sums (+ log-likelihood value term) PUSHED here by every CTA of the cluster
  unsigned long long* xbar;   // [2] mbarriers, one per recv half: complete when all cs partial-sum vectors have arrived
  float* alpha;       // [MAXL]
  float* red;         // [32] block-reduction scratch
  uint32_t* keys;     // [2 * (4 + 2*MAXL)] step keys
};

__device__ __forceinline__ Smem carve(float* base, int PP, int DS, int W, int cs) {
  Smem s;
  s.theta = base;            base += PP;
  s.aux1 = base;             base += PP;
  s.aux2 = base;             base += PP;
  s.grad = base;             base += PP;
  s.center = base;           base += PP;
  s.fb = base;               base += PP;
  s.S = base;                base += DS + 4;     // [DS] sums, [DS] = log-likelihood value term (optimiser)
  s.part = base;             base += W * DS;
  s.mu = base;               base += DS;
  s.recv = base;             base += cs > 1 ? 2 * cs * (DS + 4) : 0;
  s.xbar = reinterpret_cast<unsigned long long*>(base);   base += cs > 1 ? 4 : 0;     // 16-byte aligned: every count above is a multiple of 4
  s.alpha = base;            base += MAXL;
  s.red = base;              base += 32;
  s.keys = reinterpret_cast<uint32_t*>(base);
  return s;
}
// key words: [0,4) k1 k2 of the step, [4, 4 + 2 MAXL) leaf keys of k1, [20,24) randint plan keys (split(k2)),
// [24,32) two slots of (next carry, next sub) computed one step ahead (expand_step_keys)
constexpr int SMEM_KEY_WORDS = 2 * (4 + 2 * MAXL);
constexpr int KEY_PLAN = 4 + 2 * MAXL, KEY_NEXT = KEY_PLAN + 4;
static_assert(KEY_NEXT + 8 <= SMEM_KEY_WORDS, "key slots");
__host__ __device__ inline size_t vec_smem_bytes(int PP, int DS, int W, int cs) {
  return sizeof(float) * (size_t)(6 * PP + (DS + 4) + W * DS + DS + (cs > 1 ? 2 * cs * (DS + 4) + 4 : 0) + MAXL + 32) +
         sizeof(uint32_t) * SMEM_KEY_WORDS;
}

// ---------------------------------------------------------------------------------------
// cluster geometry of this CTA (1 / 0 for a launch without a cluster attribute)
__device__ __forceinline__ unsigned cluster_size() { unsigned r; asm("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_rank() { unsigned r; asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

// Cluster exchange without a cluster barrier (SG_VEC_ASYNC_EXCHANGE, default): every CTA pushes its folded partial sums into
// the peers' recv half with st.async, each store signalling complete_tx on the RECEIVER's mbarrier of that half; a CTA only
// waits for its own mbarrier (cs vectors expected).  `barrier.cluster.arrive.release` costs a cluster-scope memory barrier
// per step (ncu on the README chain: 13 % of all stall samples on its ERRBAR, 4 % more in the wait); these stores need none.
#ifndef SG_VEC_ASYNC_EXCHANGE
#define SG_VEC_ASYNC_EXCHANGE 1
#endif
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, unsigned rank) {   // the same location in CTA `rank` of the cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_async_f32(uint32_t remote_addr, float v, uint32_t remote_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];"
               ::"r"(remote_addr), "f"(v), "r"(remote_mbar) : "memory");
}
__device__ __forceinline__ void xbar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr_u32(bar)));
}
__device__ __forceinline__ void xbar_expect(unsigned long long* bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_addr_u32(bar)), "r"(bytes) : "memory");
}
// The st.async data lands in THIS CTA's shared memory and its complete_tx is what completes the phase, so the default
// (.acquire.cta) wait orders it - the pattern of every TMA-multicast consumer.  The .acquire.cluster form compiles to the
// same wait followed by CCTL.IVALL, an invalidation of the whole L1 on every step (9 % of the stall samples of the README chain).
__device__ __forceinline__ void xbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_addr_u32(bar)), "r"(parity) : "memory");
    if (!ok && clock64() - t0 > 8000000000LL) __trap();     // a protocol bug traps instead of hanging the GPU
  }
}

__device__ __forceinline__ float sigmoid_f(float z) {
  // stable logistic; expf (not __expf): per-row relative error must stay ~1e-7
  const float e = expf(-fabsf(z));
  const float s = 1.0f / (1.0f + e);
  return z >= 0.0f ? s : e * s;
}

// one float4 of a data row.  SG_ROW_LOAD selects the cache policy (experiments): 0 = ld.global.nc (__ldg),
// 1 = ld.global.cg (L2 only), 2 = ld.global.nc.L1::no_allocate
#ifndef SG_ROW_LOAD
#define SG_ROW_LOAD 0
#endif
__device__ __forceinline__ float4 load_row4(const float4* p) {
#if SG_ROW_LOAD == 1
  return __ldcg(p);
#elif SG_ROW_LOAD == 2
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}
__device__ __forceinline__ void axpy4(float4& acc, float w, const float4& x) {
  acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y);
  acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
}

// Sum 8 per-lane values across the warp with 9 shuffles.  On return every lane of quad
// {4u .. 4u+3} holds the warp-wide total of p[u].
__device__ __forceinline__ float reduce8(const float (&p)[8], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float q[4], r[2], s;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float send = b4 ? p[k] : p[k + 4];
    const float keep = b4 ? p[k + 4] : p[k];
    q[k] = keep + __shfl_xor_sync(FULL, send, 16);
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const float send = b3 ? q[k] : q[k + 2];
    const float keep = b3 ? q[k + 2] : q[k];
    r[k] = keep + __shfl_xor_sync(FULL, send, 8);
  }
  {
    const float send = b2 ? r[0] : r[1];
    const float keep = b2 ? r[1] : r[0];
    s = keep + __shfl_xor_sync(FULL, send, 4);
  }
  s += __shfl_xor_sync(FULL, s, 2);
  s += __shfl_xor_sync(FULL, s, 1);
  return s;
}

// ---------------------------------------------------------------------------------------
// Row pass: every warp walks its share of the index stream, 64 rows per threefry batch,
// 8 rows in flight per lane.  Result: part[warp][0..DS) = sum over the warp's rows of
//   gaussian_mean: x            logreg: w x~  with  w = -sigmoid(theta.x~) [ + sigmoid(center.x~) if CV ]
// EXT = false is the lean instantiation of the plain samplers (the bench kernel); the row-sharded mode and the
// optimiser's log-posterior values live in the EXT = true instantiation only (they cost registers in the hot loop:
// measured +6 % on the C2 step when they shared one kernel).
template <int FAMILY, int CV, int NCH, bool EXT = false>
__device__ __forceinline__ void row_pass(const VecModel& m, bool sequential, const RandintPlan& plan,
                                         unsigned seq_base, unsigned count, const float* s_theta, const float* s_center,
                                         float* s_part, float* lp_part_in = nullptr) {
  float* lp_part = EXT ? lp_part_in : nullptr;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the stream is dealt to the warps of the whole cluster: W = warps of all its CTAs, gwarp = this warp among them
  const unsigned W = (blockDim.x >> 5) * cluster_size(), gwarp = cluster_rank() * (blockDim.x >> 5) + warp;
  const unsigned h = (count + 1u) >> 1;
  const int DS = m.DS;
  const bool sharded = EXT && m.own_n != m.n_data;
  const float4* __restrict__ rows = reinterpret_cast<const float4*>(m.rows);
  const int row_f4 = DS >> 2;

  float4 acc[NCH], th[NCH], ct[NCH];
  bool ok[NCH];
  float lpv = 0.f;     // log-likelihood VALUE terms (optimizer only): gaussian sum |x|^2, logreg sum -softplus(theta.x~)
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int f4 = lane + 32 * c;
    ok[c] = f4 < row_f4;
    acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    th[c] = ct[c] = acc[c];
    if (FAMILY == SGMCMC_FAMILY_LOGREG && ok[c]) {
      th[c] = reinterpret_cast<const float4*>(s_theta)[f4];
      if (CV == 1) ct[c] = reinterpret_cast<const float4*>(s_center)[f4];
    }
  }

  // short streams (h < 32 W, e.g. one chain with batch 1000): spread the blocks over ALL warps instead of
  // filling 32 lanes of the first few, so the row groups of a step run in parallel rather than back to back
  const unsigned lpw = h < W * 32u ? (h + W - 1u) / W : 32u;
  SG_TA(t_rp);
  for (unsigned b0 = gwarp * lpw; b0 < h; b0 += W * lpw) {
    const unsigned b = b0 + lane;
    int jA = -1, jB = -1;
    if ((unsigned)lane < lpw && b < h) {
      unsigned iA, iB;
      if (sequential) { iA = seq_base + b; iB = seq_base + b + h; } else { randint_block(plan, b, iA, iB); }
      jA = (int)iA;
      if (b + h < count) jB = (int)iB;
      if (sharded && !sequential) {                      // keep only the rows this rank holds, as local row numbers
        jA = (iA - m.own_lo < m.own_n) ? (int)(iA - m.own_lo) : -1;
        if (jB >= 0) jB = (iB - m.own_lo < m.own_n) ? (int)(iB - m.own_lo) : -1;
      }
    }
    // sharded: owned rows are scattered over the lanes - compact them so that every group of 8 is full
    const unsigned maskA = sharded ? __ballot_sync(FULL, jA >= 0) : 0u, maskB = sharded ? __ballot_sync(FULL, jB >= 0) : 0u;
    const int cntA = __popc(maskA), cntB = __popc(maskB);
    SG_TB(10, t_rp);
    // the rows of this batch as one list of 2 lpw entries: [0, lpw) = jA of lanes 0 .. lpw-1, [lpw, 2 lpw) = their jB;
    // short streams (lpw < 32) so fill their groups of 8 from both halves instead of running two half-empty ones
    const int ngrp = sharded ? 8 : (int)((2u * lpw + 7u) >> 3);
#pragma unroll 1
    for (int grp = 0; grp < ngrp; ++grp) {
      const int src = (grp < 4) ? jA : jB;
      const int base = (grp & 3) * 8;
      int r[8];
      if (sharded) {
        const unsigned mask = (grp < 4) ? maskA : maskB;
        const int cnt = (grp < 4) ? cntA : cntB;
        if (base >= cnt) continue;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int from = __fns(mask, 0, base + u + 1);       // lane of the (base + u)-th owned row, or -1
          const int v = __shfl_sync(FULL, src, from & 31);
          r[u] = base + u < cnt ? v : -1;
        }
      } else if (lpw == 32u) {
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u] = __shfl_sync(FULL, src, base + u);
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const unsigned e = 8u * grp + u;
          const int v = __shfl_sync(FULL, e >= lpw ? jB : jA, (e >= lpw ? e - lpw : e) & 31u);
          r[u] = e < 2u * lpw ? v : -1;
        }
      }
      // rows past the end of the stream are -1 (loaded as zeros); a group without any row is skipped
      if ((r[0] & r[1] & r[2] & r[3] & r[4] & r[5] & r[6] & r[7]) < 0) continue;
      float4 x[8][NCH];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          x[u][c] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok[c] && r[u] >= 0) x[u][c] = load_row4(rows + (size_t)r[u] * row_f4 + lane + 32 * c);
        }
      }
      // CV == 2: the lanes of quad u (where reduce8 leaves theta . x~ of row u) read that row's center . x~ themselves;
      // the word shares a sector with the row's last float4, so this is one more request, not more traffic
      float zc = 0.f;
      if (CV == 2) {
        int ru;
        if (sharded || lpw != 32u) {
          ru = r[0];
#pragma unroll
          for (int u = 1; u < 8; ++u) ru = (lane >> 2) == u ? r[u] : ru;
        } else {
          ru = __shfl_sync(FULL, src, base + (lane >> 2));
        }
        if (ru >= 0) zc = __ldg(m.rows + (size_t)ru * DS + 4 * m.zc_f4);
      }
      if (FAMILY == SGMCMC_FAMILY_GAUSSIAN_MEAN) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
          for (int c = 0; c < NCH; ++c) {
            acc[c].x += x[u][c].x; acc[c].y += x[u][c].y; acc[c].z += x[u][c].z; acc[c].w += x[u][c].w;
          }
        if (lp_part) {
#pragma unroll
          for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < NCH; ++c) lpv += dot4(x[u][c], x[u][c]);
        }
      } else {
        float p[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          p[u] = dot4(x[u][0], th[0]);
#pragma unroll
          for (int c = 1; c < NCH; ++c) p[u] += dot4(x[u][c], th[c]);
        }
        const float z = reduce8(p, lane);
        float wgt = -sigmoid_f(z);
        if (lp_part) {                                   // lane 4u holds z of row u: loglik = -softplus(z)
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (lane == 4 * u && r[u] >= 0) lpv -= fmaxf(z, 0.f) + log1pf(expf(-fabsf(z)));
        }
        if (CV == 2) {
          wgt += sigmoid_f(zc);
        } else if (CV) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            p[u] = dot4(x[u][0], ct[0]);
#pragma unroll
            for (int c = 1; c < NCH; ++c) p[u] += dot4(x[u][c], ct[c]);
          }
          wgt += sigmoid_f(reduce8(p, lane));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          // rows past the tail were loaded as zeros, so their weight is irrelevant
          const float wu = __shfl_sync(FULL, wgt, 4 * u);
#pragma unroll
          for (int c = 0; c < NCH; ++c) axpy4(acc[c], wu, x[u][c]);
        }
      }
    }
  }
  SG_TB(11, t_rp);
#pragma unroll
  for (int c = 0; c < NCH; ++c)
    if (ok[c]) reinterpret_cast<float4*>(s_part + warp * DS)[lane + 32 * c] = acc[c];
  if (lp_part) {
#pragma unroll
    for (int o = 16; o; o >>= 1) lpv += __shfl_xor_sync(FULL, lpv, o);
    if (lane == 0) lp_part[warp] = lpv;
  }
}

// rows[i][4 zc_f4] = center . x~_i for every row of the table, with exactly the arithmetic row_pass uses for theta . x~
// (dot4 per lane and chunk, xor butterfly 16..1 = reduce8's tree), so that at theta == center the CV weights cancel
// bit for bit like they do when both dots are taken in the kernel.
template <int NCH>
__global__ void zc_column_kernel(float* __restrict__ rows, unsigned n_rows, int DS, int P, int zc_f4, const float* __restrict__ center) {
  const int lane = threadIdx.x & 31;
  const int row_f4 = DS >> 2;
  float4 ct[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int e = 4 * (lane + 32 * c);
    ct[c] = make_float4(e < P ? center[e] : 0.f, e + 1 < P ? center[e + 1] : 0.f, e + 2 < P ? center[e + 2] : 0.f,
                        e + 3 < P ? center[e + 3] : 0.f);
  }
  const unsigned warps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_rows; r += warps) {
    const float4* row = reinterpret_cast<const float4*>(rows) + (size_t)r * row_f4;
    float4 x[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) x[c] = (lane + 32 * c < row_f4) ? row[lane + 32 * c] : make_float4(0.f, 0.f, 0.f, 0.f);
    float p = dot4(x[0], ct[0]);
#pragma unroll
    for (int c = 1; c < NCH; ++c) p += dot4(x[c], ct[c]);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) p += __shfl_xor_sync(FULL, p, o);
    if (lane == 0) rows[(size_t)r * DS + 4 * zc_f4] = p;
  }
}

// part[W][DS] (+ red[W], the log-likelihood value terms, if with_lp) -> S[DS] (+ S[DS]); fixed summation order.
// In a cluster: the local fold is pushed to recv[buf][rank] of every CTA, ONE cluster barrier, then every CTA adds
// its recv[buf][0 .. cs) in rank order (see the header comment for why the double buffer makes one barrier enough).
__device__ __forceinline__ void fold_parts(const Smem& s, int DS, int& buf, bool with_lp = false) {
  const int W = blockDim.x >> 5;
  const unsigned cs = cluster_size();
  const int n = DS + (with_lp ? 1 : 0), stride = DS + 4;
  SG_TA(t_f);
  __syncthreads();
  SG_TB(12, t_f);
  if (cs == 1) {
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
      float t = 0.f;
      if (d < DS) { for (int w = 0; w < W; ++w) t += s.part[w * DS + d]; }
      else        { for (int w = 0; w < W; ++w) t += s.red[w]; }
      s.S[d] = t;
    }
  } else {
    float* slot = s.recv + ((size_t)(buf & 1) * cs + cluster_rank()) * stride;
#if SG_VEC_ASYNC_EXCHANGE
    // bit 0 of buf: recv half; bits 1, 2: phase parity of xbar[0], xbar[1]
    unsigned long long* bar = s.xbar + (buf & 1);
    if (threadIdx.x == 0) xbar_expect(bar, cs * (unsigned)n * 4u);
    const uint32_t bar_a = smem_addr_u32(bar);
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
      float t = 0.f;
      if (d < DS) { for (int w = 0; w < W; ++w) t += s.part[w * DS + d]; }
      else        { for (int w = 0; w < W; ++w) t += s.red[w]; }
      const uint32_t dst = smem_addr_u32(slot + d);
      for (unsigned r = 0; r < cs; ++r) st_async_f32(map_to_rank(dst, r), t, map_to_rank(bar_a, r));
    }
    SG_TB(5, t_f);
    xbar_wait(bar, (buf >> (1 + (buf & 1))) & 1);
    buf ^= 2 << (buf & 1);                                  // this half's next phase
    SG_TB(6, t_f);
#else
    cooperative_groups::cluster_group cl = cooperative_groups::this_cluster();
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
      float t = 0.f;
      if (d < DS) { for (int w = 0; w < W; ++w) t += s.part[w * DS + d]; }
      else        { for (int w = 0; w < W; ++w) t += s.red[w]; }
      for (unsigned r = 0; r < cs; ++r) cl.map_shared_rank(slot, r)[d] = t;
    }
    SG_TB(5, t_f);
    cl.sync();
    SG_TB(6, t_f);
#endif
    const float* mine = s.recv + (size_t)(buf & 1) * cs * stride;
    for (int d = threadIdx.x; d < n; d += blockDim.x) {
      float t = 0.f;
      for (unsigned r = 0; r < cs; ++r) t += mine[r * stride + d];
      s.S[d] = t;
    }
    buf ^= 1;
  }
  __syncthreads();
  SG_TB(7, t_f);
}

// grad_log_post at theta from the folded row sum S (util.py:43-44).
//   rows_in_S = number of rows summed (as float), scale = Ndata / rows_in_S
//   mu_d = column mean the rows were centred by (gaussian_mean; 0 if not centred): S_d sums x - mu
template <int FAMILY>
__device__ __forceinline__ float grad_from_sum(const VecModel& m, int leaf, int d, float S_d, float theta_e,
                                               float scale, float rows_in_S, float mu_d = 0.f) {
  if (FAMILY == SGMCMC_FAMILY_GAUSSIAN_MEAN)
    return scale * (m.lik2[leaf] * (S_d - rows_in_S * (theta_e - mu_d))) - m.prior2[leaf] * theta_e;
  return scale * S_d - m.prior2[0] * theta_e;
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < W; ++w) t += red[w];   // same order in every thread: deterministic
  return t;
}

// ---------------------------------------------------------------------------------------
template <int FAMILY, int CV, int NCH, bool EXT = false>
struct ChainCtx {
  const VecModel& m;
  Smem s;
  int buf = 0;              // bit 0: recv half of the next cluster fold; bits 1, 2: phase parities of its two mbarriers
  bool plan_cached = false; // keys[KEY_PLAN ..] hold split(k2) of the estimate about to run
  int look_slot = -1;       // >= 0: the next expand_step_keys also splits look_carry into keys[KEY_NEXT + 4 look_slot ..]
  Key look_carry{0u, 0u};
  __device__ ChainCtx(const VecModel& mm, const Smem& ss) : m(mm), s(ss) {}

  // full-batch gradient at s.theta into dst[] (sequential pass over every row)
  __device__ void full_grad(float* dst) {
    RandintPlan dummy{};
    row_pass<FAMILY, false, NCH, EXT>(m, true, dummy, 0u, m.own_n, s.theta, s.center, s.part);
    fold_parts(s, m.DS, buf);
    const float rows = (float)m.own_n;
    for (int e = threadIdx.x; e < m.P; e += blockDim.x) {
      const int leaf = e / m.leaf_dim, d = e - leaf * m.leaf_dim;
      dst[e] = grad_from_sum<FAMILY>(m, leaf, d, s.S[d], s.theta[e], 1.0f, rows, s.mu[d]);
    }
    __syncthreads();
  }

  // estimate_gradient(i, key, theta, svrg_state) -> s.grad
  __device__ void estimate(int i, Key k2, bool is_init) {
    if (m.estimator == SGMCMC_EST_SVRG) {
      if (is_init) {                                   // gradient_estimation.py:124-128
        full_grad(s.fb);
        for (int e = threadIdx.x; e < m.P; e += blockDim.x) { s.center[e] = s.theta[e]; s.grad[e] = s.fb[e]; }
        __syncthreads();
        return;
      }
      if (i % m.update_rate == 0) {                    // gradient_estimation.py:134-139
        full_grad(s.fb);
        for (int e = threadIdx.x; e < m.P; e += blockDim.x) s.center[e] = s.theta[e];
        __syncthreads();
      }
    } else if (m.estimator == SGMCMC_EST_STANDARD && m.full_batch && !is_init) {
      full_grad(s.grad);                               // gradient_estimation.py:41-42
      return;
    }
    partial_sums(k2);
    SG_TA(t_fs);
    finish_from_sums();
    SG_TB(8, t_fs);
  }

  // minibatch sums over the rows of this rank: s.S[d] = sum_i x_i (gaussian) / sum_i (w_i [- w_c,i]) x~_i (logreg)
  __device__ void partial_sums(Key k2, bool want_lp = false, bool all_rows = false) {
    SG_TA(t_ps);
    RandintPlan plan{};
    if (all_rows) {                                      // static full-batch bypass (gradient_estimation.py:41-42): no RNG
      row_pass<FAMILY, false, NCH, EXT>(m, true, plan, 0u, m.own_n, s.theta, s.center, s.part, want_lp ? s.red : nullptr);
      fold_parts(s, m.DS, buf, EXT && want_lp);
      return;
    }
    if (plan_cached) {                                   // split(k2) was taken next to the leaf keys
      plan.k_hi.a = s.keys[KEY_PLAN]; plan.k_hi.b = s.keys[KEY_PLAN + 1];
      plan.k_lo.a = s.keys[KEY_PLAN + 2]; plan.k_lo.b = s.keys[KEY_PLAN + 3];
      plan.span = m.n_data; plan.batch = m.batch; plan.mult = m.randint_mult;
      plan_cached = false;
    } else {
      plan = randint_plan(k2, m.n_data, m.batch);
    }
    SG_TB(3, t_ps);
    row_pass<FAMILY, CV, NCH, EXT>(m, false, plan, 0u, m.batch, s.theta, s.center, s.part, want_lp ? s.red : nullptr);
    SG_TB(4, t_ps);
    fold_parts(s, m.DS, buf, EXT && want_lp);
  }

  // log_post VALUE on the minibatch whose sums are in s.S (value term in s.S[DS]) (util.py:43-44: logprior + Ndata * mean_i loglik_i)
  __device__ float log_post_value() {
    const float lps = s.S[m.DS];
    const float B = (float)m.batch;
    float lik = 0.f, pri = 0.f;
    if (FAMILY == SGMCMC_FAMILY_GAUSSIAN_MEAN) {
      // with x' = x - mu (the stored rows), u = th - mu:  sum_i |x_i - th|^2 = sum |x'|^2 - 2 u.S' + B |u|^2
      for (int leaf = 0; leaf < m.n_leaves; ++leaf) {
        float a = 0.f, b = 0.f, c = 0.f;
        for (int d = threadIdx.x; d < m.leaf_dim; d += blockDim.x) {
          const float t = s.theta[leaf * m.leaf_dim + d], u = t - s.mu[d];
          a = fmaf(u, s.S[d], a); b = fmaf(u, u, b); c = fmaf(t, t, c);
        }
        const float uS = block_sum(a, s.red);
        const float uu = block_sum(b, s.red);
        const float tt = block_sum(c, s.red);
        lik -= 0.5f * m.lik2[leaf] * (lps - 2.0f * uS + B * uu);
        pri -= 0.5f * m.prior2[leaf] * tt;
      }
    } else {
      float b = 0.f;
      for (int d = threadIdx.x; d < m.leaf_dim; d += blockDim.x) b = fmaf(s.theta[d], s.theta[d], b);
      pri = -0.5f * m.prior2[0] * block_sum(b, s.red);
      lik = lps;
    }
    return pri + m.scale * lik;
  }

  // optax.adam ascent step on the gradient in s.grad (oracle/optimizer.py; cf.ca = 1 - b1^t, cf.cb = 1 - b2^t, cf.h = lr)
  __device__ void adam_step(const sgmcmc_step_coef& cf) {
    const float b1 = m.hyper[0], b2 = m.hyper[1], eps = m.hyper[2];
    for (int e = threadIdx.x; e < m.P; e += blockDim.x) {
      const float g = s.grad[e];
      const float mm = b1 * s.aux1[e] + (1.0f - b1) * g;
      const float vv = b2 * s.aux2[e] + (1.0f - b2) * (g * g);
      s.aux1[e] = mm; s.aux2[e] = vv;
      s.theta[e] = s.theta[e] + cf.h * (mm / cf.ca) / (sqrtf(vv / cf.cb) + eps);
    }
    __syncthreads();
  }

  // s.grad from the (complete) minibatch sums in s.S, evaluated at s.theta
  __device__ void finish_from_sums() {
    const float rows = (float)m.batch;
    for (int e = threadIdx.x; e < m.P; e += blockDim.x) {
      const int leaf = e / m.leaf_dim, d = e - leaf * m.leaf_dim;
      float g;
      if (!CV) {
        g = grad_from_sum<FAMILY>(m, leaf, d, s.S[d], s.theta[e], m.scale, rows, s.mu[d]);
      } else if (FAMILY == SGMCMC_FAMILY_GAUSSIAN_MEAN) {
        // (c + g) - gc, gradient_estimation.py:72
        const float gt = grad_from_sum<FAMILY>(m, leaf, d, s.S[d], s.theta[e], m.scale, rows, s.mu[d]);
        const float gc = grad_from_sum<FAMILY>(m, leaf, d, s.S[d], s.center[e], m.scale, rows, s.mu[d]);
        g = (s.fb[e] + gt) - gc;
      } else {
        // S already holds sum (w - w_c) x~ : c + (g - gc) without the fp32 cancellation
        g = s.fb[e] + (m.scale * s.S[d] - m.prior2[0] * (s.theta[e] - s.center[e]));
      }
      s.grad[e] = g;
    }
    __syncthreads();
  }

  // per-leaf noise keys: split(k, n_leaves)[leaf]   (diffusion_util.py:66,91,112)
  __device__ __forceinline__ Key leaf_key(int slot, int leaf) const {
    Key k; k.a = s.keys[slot + 2 * leaf]; k.b = s.keys[slot + 2 * leaf + 1]; return k;
  }

  // The key tree of one langevin step, spread over three warps that run side by side (state-independent work,
  // one barrier):
  //   warp 0: k1, k2 = split(key) -> keys[0..3];  leaf keys split(k1, n_leaves) -> keys[4 + 2j ..]
  //   warp 1: k2 again;  the randint plan's keys split(k2) (jax draws hi / lo bits with them) -> keys[KEY_PLAN ..]
  //   warp 2: if the sampler loop asked for it (look_slot >= 0): split(carry) of the NEXT iteration
  //           (samplers.py:49) -> keys[KEY_NEXT + 4 look_slot ..] = (next carry, next sub)
  // (CTAs with fewer warps: the last warp takes the remaining roles in turn.)
  __device__ void expand_step_keys(Key key) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int w_plan = W > 1 ? 1 : 0, w_next = W > 2 ? 2 : W - 1;
    if (warp == 0 || warp == w_plan) {
      const uint32_t w = lane < 4 ? stream_word(key, lane, 4u) : 0u;        // lanes 0..3: k1.a k1.b k2.a k2.b
      if (warp == 0) {
        Key k1; k1.a = __shfl_sync(FULL, w, 0); k1.b = __shfl_sync(FULL, w, 1);
        if (lane < 4) s.keys[lane] = w;
        const uint32_t nl = m.n_leaves;
        if (lane < 2 * (int)nl) s.keys[4 + lane] = stream_word(k1, lane, 2u * nl);
      }
      if (warp == w_plan) {
        Key k2; k2.a = __shfl_sync(FULL, w, 2); k2.b = __shfl_sync(FULL, w, 3);
        if (lane < 4) s.keys[KEY_PLAN + lane] = stream_word(k2, lane, 4u);   // k_hi.a k_hi.b k_lo.a k_lo.b
      }
    }
    if (look_slot >= 0 && warp == w_next && lane < 4) s.keys[KEY_NEXT + 4 * look_slot + lane] = stream_word(look_carry, lane, 4u);
    look_slot = -1;
    plan_cached = true;
    __syncthreads();
  }

  // diffusion update, phase 1 = `update` (before the new gradient), phase 2 = `update2`
  __device__ void diffuse(int phase, int i, const sgmcmc_step_coef& cf, int key_slot) {
    const int D = m.leaf_dim;
    const bool bug = (m.flags & SGMCMC_FLAG_BADODABCV_REFERENCE_BUG) && m.diffusion == SGMCMC_DIFF_BADODAB;
    if (phase == 2 || bug) {                      // baoab / badodab update2: v += dt/2 * g
      for (int e = threadIdx.x; e < m.P; e += blockDim.x) s.aux1[e] = fmaf(cf.hh, s.grad[e], s.aux1[e]);
      __syncthreads();
      return;
    }
    for (int leaf = 0; leaf < m.n_leaves; ++leaf) {
      const Key lk = leaf_key(key_slot, leaf);
      const int off = leaf * D;
      switch (m.diffusion) {
        case SGMCMC_DIFF_SGLD:
          for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const int e = off + d;
            s.theta[e] = fmaf(cf.ca, normal_at(lk, d, D), fmaf(cf.h, s.grad[e], s.theta[e]));
          }
          break;
        case SGMCMC_DIFF_PSGLD:
          for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const int e = off + d;
            const float g = s.grad[e], alpha = m.hyper[0], eps = m.hyper[1];
            const float v = alpha * s.aux1[e] + (1.0f - alpha) * (g * g);
            const float G = 1.0f / (sqrtf(v) + eps);
            s.aux1[e] = v;
            s.theta[e] = s.theta[e] + cf.hh * G * g + sqrtf(cf.h * G) * normal_at(lk, d, D);
          }
          break;
        case SGMCMC_DIFF_SGLDADAM:
          for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const int e = off + d;
            const float g = s.grad[e], b1 = m.hyper[0], b2 = m.hyper[1], eps = m.hyper[2];
            const float mm = b1 * s.aux1[e] + (1.0f - b1) * g;
            const float vv = b2 * s.aux2[e] + (1.0f - b2) * (g * g);
            const float mh = mm / cf.ca, vh = vv / cf.cb;
            const float a = cf.h / (sqrtf(vh) + eps);
            s.aux1[e] = mm; s.aux2[e] = vv;
            s.theta[e] = s.theta[e] + a * 0.5f * mh + sqrtf(a) * normal_at(lk, d, D);
          }
          break;
        case SGMCMC_DIFF_SGHMC:
          for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const int e = off + d;
            const float v = s.aux1[e], alpha = m.hyper[0];
            s.theta[e] += v;
            s.aux1[e] = v + cf.h * s.grad[e] - alpha * v + cf.ca * normal_at(lk, d, D);
          }
          break;
        case SGMCMC_DIFF_BAOAB:
          for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const int e = off + d;
            float v = fmaf(cf.hh, s.grad[e], s.aux1[e]);
            float x = s.theta[e] + v * cf.h * 0.5f;
            v = cf.ca * v + cf.cb * normal_at(lk, d, D);
            x = x + v * cf.h * 0.5f;
            s.aux1[e] = v; s.theta[e] = x;
          }
          break;
        case SGMCMC_DIFF_SGNHT: {
          Key kn = lk;
          const float alpha = s.alpha[leaf];
          float ss = 0.f;
          if (i == 0) {                              // diffusions.py:268-278
            Key sub; split2(lk, kn, sub);
            for (int d = threadIdx.x; d < D; d += blockDim.x) s.aux1[off + d] = cf.cb * normal_at(sub, d, D);
          }
          for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const int e = off + d;
            float v = s.aux1[e];
            v = v - alpha * v + cf.h * s.grad[e] + cf.ca * normal_at(kn, d, D);
            s.aux1[e] = v; s.theta[e] += v;
            ss = fmaf(v, v, ss);
          }
          const float nrm = sqrtf(block_sum(ss, s.red));
          if (threadIdx.x == 0) s.alpha[leaf] = alpha + (nrm * nrm) / (float)D - cf.h;
          break;
        }
        case SGMCMC_DIFF_BADODAB: {
          float alpha = s.alpha[leaf];
          float ss = 0.f;
          for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const int e = off + d;
            const float v = fmaf(cf.hh, s.grad[e], s.aux1[e]);
            s.aux1[e] = v; s.theta[e] = fmaf(cf.hh, v, s.theta[e]);
            ss = fmaf(v, v, ss);
          }
          alpha = alpha + cf.hh * (sqrtf(block_sum(ss, s.red)) - (float)D);
          // correctly rounded fp32 exp: 1 - c1^2 amplifies a 1-ulp error of expf by ~1/(alpha h)
          const float c1 = (float)exp((double)(-alpha * cf.h));
          const float c2 = (alpha == 0.0f) ? cf.ca : sqrtf(fabsf((1.0f - c1 * c1) / (2.0f * alpha)));
          ss = 0.f;
          for (int d = threadIdx.x; d < D; d += blockDim.x) {
            const int e = off + d;
            const float v = c1 * s.aux1[e] + c2 * normal_at(lk, d, D);
            s.aux1[e] = v; ss = fmaf(v, v, ss);
          }
          alpha = alpha + cf.hh * (sqrtf(block_sum(ss, s.red)) - (float)D);
          for (int d = threadIdx.x; d < D; d += blockDim.x) s.theta[off + d] = fmaf(cf.hh, s.aux1[off + d], s.theta[off + d]);
          __syncthreads();
          if (threadIdx.x == 0) s.alpha[leaf] = alpha;
          break;
        }
      }
    }
    __syncthreads();
  }

  // langevin kernel, kernels.py:53-64
  __device__ void langevin(int i, Key key, const sgmcmc_step_coef& cf) {
    SG_TA(t_l);
    expand_step_keys(key);
    SG_TB(1, t_l);
    Key k2; k2.a = s.keys[2]; k2.b = s.keys[3];
    diffuse(1, i, cf, 4);
    SG_TB(2, t_l);
    estimate(i, k2, false);
    plan_cached = false;
    SG_TB(13, t_l);
    if (m.diffusion == SGMCMC_DIFF_BAOAB || m.diffusion == SGMCMC_DIFF_BADODAB) diffuse(2, i, cf, 4);
  }

  // one sample: langevin kernel, or the SGHMC wrapper of kernels.py:110-122
  __device__ void transition(int i, Key key, const sgmcmc_step_coef& cf) {
    if (m.diffusion != SGMCMC_DIFF_SGHMC) { langevin(i, key, cf); return; }
    Key k1, k2;
    split2(key, k1, k2);
    const int D = m.leaf_dim;
    for (int leaf = 0; leaf < m.n_leaves; ++leaf) {      // resample_momentum, diffusions.py:197-199
      const Key lk = split_at(k1, m.n_leaves, leaf);
      for (int d = threadIdx.x; d < D; d += blockDim.x) s.aux1[leaf * D + d] = cf.cb * normal_at(lk, d, D);
    }
    __syncthreads();
    for (int l = 0; l < m.leapfrog; ++l) langevin(i, split_at(k2, m.leapfrog, l), cf);
  }
};

template <int FAMILY, int CV, int NCH, bool EXT = false>
__global__ void __launch_bounds__(1024 / NCH) vec_sampler_kernel(const VecModel m, const VecRun run) {
  extern __shared__ __align__(16) float smem_raw[];
  const int W = blockDim.x >> 5;
  const int PP = ((m.P > m.DS ? m.P : m.DS) + 3) & ~3;
  const unsigned cs = cluster_size();
  const Smem s = carve(smem_raw, PP, m.DS, W, (int)cs);
  const int c = blockIdx.x / cs;               // one chain per cluster (cs = 1: per CTA)
  const bool writer = cluster_rank() == 0;     // the replicas are identical; rank 0 writes samples and state back
  const int P = m.P;
  const sgmcmc_state& st = run.st;

  for (int e = threadIdx.x; e < PP; e += blockDim.x) {
    const bool in = e < P;
    s.theta[e] = in ? st.theta[(size_t)c * P + e] : 0.f;
    s.aux1[e] = (in && st.aux1 && run.mode == 0) ? st.aux1[(size_t)c * P + e] : 0.f;
    s.aux2[e] = (in && st.aux2 && run.mode == 0) ? st.aux2[(size_t)c * P + e] : 0.f;
    s.grad[e] = (in && run.mode == 0) ? st.grad[(size_t)c * P + e] : 0.f;
    float ce = 0.f, fe = 0.f;
    if (in && m.estimator == SGMCMC_EST_CV) { ce = m.cv_center[e]; fe = m.cv_grad[e]; }
    if (in && m.estimator == SGMCMC_EST_SVRG && run.mode == 0) {
      ce = st.svrg_center[(size_t)c * P + e]; fe = st.svrg_grad[(size_t)c * P + e];
    }
    s.center[e] = ce; s.fb[e] = fe;
  }
  if (threadIdx.x < MAXL) {
    float a = 0.f;
    if (st.alpha && threadIdx.x < m.n_leaves)
      a = run.mode == 0 ? st.alpha[(size_t)c * m.n_leaves + threadIdx.x] : m.hyper[0];
    s.alpha[threadIdx.x] = a;
  }
  for (int d = threadIdx.x; d < m.DS; d += blockDim.x)
    s.mu[d] = (FAMILY == SGMCMC_FAMILY_GAUSSIAN_MEAN && m.mu) ? m.mu[d] : 0.f;
  if (SG_VEC_ASYNC_EXCHANGE && cs > 1) {
    if (threadIdx.x == 0) {
      xbar_init(s.xbar); xbar_init(s.xbar + 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cooperative_groups::this_cluster().sync();          // once per launch: every CTA's barriers exist before the first push
  }
  __syncthreads();

  ChainCtx<FAMILY, CV, NCH, EXT> ctx(m, s);
  Key carry{0u, 0u};

  if (EXT && run.sum_in) {                            // finish the previous estimate from the all-reduced sums
    for (int d = threadIdx.x; d < m.DS; d += blockDim.x) s.S[d] = run.sum_in[(size_t)c * m.DS + d];
    __syncthreads();
    ctx.finish_from_sums();
    if (run.shard_flags & SHARD_UPDATE2) ctx.diffuse(2, run.i0, *run.coef_prev, 4);
  }
  if (run.mode != 0) {
    // init_fn (kernels.py:48-51), optionally preceded by samplers.py:53
    Key k; k.a = run.init_keys[2 * c]; k.b = run.init_keys[2 * c + 1];
    if (run.mode == 2) { Key sub; split2(k, carry, sub); k = sub; }
    if (EXT && run.sum_out) ctx.partial_sums(k); else ctx.estimate(0, k, true);
  } else if (EXT && (run.sum_out || run.sum_in)) {
    // data-sharded step: one transition whose estimate stops at this rank's partial sums
    carry.a = st.key[2 * c]; carry.b = st.key[2 * c + 1];
    if (run.shard_flags & SHARD_STEP) {
      const sgmcmc_step_coef cf = run.coef[0];
      Key nk, sub; split2(carry, nk, sub); carry = nk;           // samplers.py:49
      ctx.expand_step_keys(sub);
      Key k2; k2.a = s.keys[2]; k2.b = s.keys[3];
      ctx.diffuse(1, run.i0, cf, 4);
      if (run.samples && writer) {
        float* dst = run.samples + (size_t)c * run.sample_chain_stride;
        for (int e = threadIdx.x; e < P; e += blockDim.x) dst[e] = s.theta[e];
      }
      ctx.partial_sums(k2);
    }
  } else {
    if (run.step_keys == nullptr) { carry.a = st.key[2 * c]; carry.b = st.key[2 * c + 1]; }
    bool have_next = false;
    sgmcmc_step_coef cf_next = run.coef[0];
    for (int sidx = 0; sidx < run.K; ++sidx) {
      SG_TA(t_s);
      const int i = run.i0 + sidx;
      const sgmcmc_step_coef cf = cf_next;                       // loaded one iteration ahead (global-memory latency)
      if (run.coef_stride && sidx + 1 < run.K) cf_next = run.coef[(size_t)(sidx + 1) * run.coef_stride];
      Key sub;
      if (run.step_keys) { sub.a = run.step_keys[2 * c]; sub.b = run.step_keys[2 * c + 1]; }
      else if (have_next) {                                      // split one iteration ahead by expand_step_keys
        const uint32_t* nx = s.keys + KEY_NEXT + 4 * (sidx & 1);
        carry.a = nx[0]; carry.b = nx[1]; sub.a = nx[2]; sub.b = nx[3];
      }
      else { Key nk; split2(carry, nk, sub); carry = nk; }      // samplers.py:49 / optimizer.py:46
      have_next = false;
      if (run.step_keys == nullptr && sidx + 1 < run.K && !(EXT && m.diffusion == SGMCMC_DIFF_ADAM_OPT)) {
        ctx.look_slot = (sidx + 1) & 1; ctx.look_carry = carry; have_next = true;
      }
      SG_TB(0, t_s);
      if (EXT && m.diffusion == SGMCMC_DIFF_ADAM_OPT) {
        // optimizer.py:45-51: value and gradient on the minibatch drawn with `sub` itself, then the Adam step
        ctx.partial_sums(sub, run.lp_out != nullptr, m.full_batch != 0);   // batch == N: all rows, scale 1 (optimizer.py:47)
        ctx.finish_from_sums();
        if (run.lp_out) {
          const float lp = ctx.log_post_value();
          if (threadIdx.x == 0 && writer) run.lp_out[(size_t)c * run.K + sidx] = lp;
        }
        ctx.adam_step(cf);
        continue;
      }
      ctx.transition(i, sub, cf);
      SG_TB(14, t_s);
      const bool thinned = run.thin > 1;
      if (thinned && (sidx + 1) % run.thin != 0) continue;
      if (run.samples && writer) {
        const int slot = thinned ? (sidx + 1) / run.thin - 1 : sidx;
        float* dst = run.samples + (size_t)c * run.sample_chain_stride + (size_t)slot * P;
        for (int e = threadIdx.x; e < P; e += blockDim.x) dst[e] = s.theta[e];
      }
      if (EXT && run.moments && writer) {
        float* mo = run.moments + (size_t)c * 3 * P;
        for (int e = threadIdx.x; e < P; e += blockDim.x) { const float x = s.theta[e] - mo[2 * P + e]; mo[e] += x; mo[P + e] = fmaf(x, x, mo[P + e]); }
      }
      SG_TB(9, t_s);
    }
  }

  // write the chain state back
  if (cs > 1) cooperative_groups::this_cluster().sync();   // no CTA may exit while a peer still reads its sloc
  if (!writer) return;
  for (int e = threadIdx.x; e < P; e += blockDim.x) {
    const size_t o = (size_t)c * P + e;
    if (run.mode == 0) st.theta[o] = s.theta[e];
    if (st.aux1) st.aux1[o] = s.aux1[e];
    if (st.aux2) st.aux2[o] = s.aux2[e];
    st.grad[o] = s.grad[e];
    if (m.estimator == SGMCMC_EST_SVRG) { st.svrg_center[o] = s.center[e]; st.svrg_grad[o] = s.fb[e]; }
  }
  if (st.alpha && threadIdx.x < m.n_leaves) st.alpha[(size_t)c * m.n_leaves + threadIdx.x] = s.alpha[threadIdx.x];
  if (threadIdx.x == 0 && st.key && (run.mode == 2 || (run.mode == 0 && run.step_keys == nullptr))) {
    st.key[2 * c] = carry.a; st.key[2 * c + 1] = carry.b;
  }
  if (EXT && run.sum_out)
    for (int d = threadIdx.x; d < m.DS; d += blockDim.x) run.sum_out[(size_t)c * m.DS + d] = s.S[d];
}

}  // namespace sg
