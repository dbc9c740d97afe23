// K1-big: fused persistent sampler for families whose parameter vector does not fit in shared
// memory (PMF: P = 52,500; Bayesian MLP 784-100-10: P = 79,510).
//
// One CTA (1024 threads) owns one chain for all K steps of a launch.  theta, the diffusion's
// auxiliary arrays and the carried gradient stay in the caller's [C, P] global buffers; while a
// launch runs, the state of the ~148 chains that are resident at once (<= 148 * 4 * 318 KB) is
// L2-resident (126 MB), so only data rows and emitted samples travel to / from HBM inside the loop.
// Per step (same key tree and operation order as csrc/vec_sampler.cuh):
//   key tree          samplers.py:47-51, kernels.py:53-64,110-122, diffusion_util.py:66,91,112
//   diffusion update  diffusions.py:47-347; one threefry block feeds elements m and m + ceil(n/2)
//                     of a leaf (jax's half-split counter layout), so no PRNG word is wasted
//   index draw        gradient_estimation.py:32,75,140 (bit-exact randint)
//   gradient          util.py:43-49 for the family: family_lik_pass() ADDS sum_i grad loglik_i(theta)
//                     (minus the same at the CV / SVRG centre) into the zeroed gradient buffer,
//                     finalize() applies N/B, the prior and the full-batch centre gradient.
//
// Families
//   PMF: theta = (U[nu, r], V[ni, r]); rating n = (i, j, y) packed as one int4;
//       loglik = -alpha/2 (y - U_i.V_j)^2, prior N(0, 1/lam): sparse row gather + scatter-add (float4 L2 atomics) into
//       the dense gradient buffer.  Opt-in alternative (SGMCMC_PMF_COMPACT=1): a minibatch of B ratings touches at most
//       2 B of the nu + ni factor rows, so its raw gradient can be kept COMPACT in shared memory (PmfCompact below:
//       one slot of r floats per touched row, filled by the row's owner in a fixed order - no float atomics, no dense
//       P-vector that is zeroed, scattered into and read back every step; the diffusion pass looks the slot up).
//       Deterministic, and it removes 2 of the 7 state accesses per element - but measured slower (capi.cu), because the
//       kernel is bound by its dependent RNG chains at 16 warps per SM, not by traffic (DESIGN.md, K1-big, round-2 correction).  Full passes over all N ratings (SVRG refresh,
//       K2, the full-batch bypass) always use the dense buffer.
//   MLP (models/bayesian_NN/NN_model.py:31-58): see mlp_grad.cuh.
#pragma once
#include <cstdlib>
#include "prng.cuh"
#include "vec_sampler.cuh"   // Key helpers, FULL, MAXL, VecRun layout reuse

namespace sg {

// PMF: CTA shape of one chain.  Measured on C4 (512 chains, SGNHT, ms per 4 steps): 1024 threads x 64 registers 1.37 (the
// elementwise passes spill the values they load, so every load is waited for at once), 512 x 128 regs 1.15,
// 2 CTAs x 256 threads x 128 regs with 4 threefry blocks staged per thread 1.08 (default), 4 CTAs x 128 threads 1.11.
// Default: 2 CTAs x 256 threads x 128 registers per SM; with the opt-in compact minibatch gradient (<= 190 KB of shared
// memory for C4) one CTA of 512 threads (big_threads() picks at launch time; the kernel is compiled for <= 512).
#ifndef SG_BIG_THREADS_PMF
#define SG_BIG_THREADS_PMF 512
#endif
#ifndef SG_BIG_MINBLOCKS_PMF
#define SG_BIG_MINBLOCKS_PMF 1
#endif
template <int FAMILY> struct BigThreads { static constexpr int value = SG_BIG_THREADS_PMF; static constexpr int min_blocks = SG_BIG_MINBLOCKS_PMF; };
// MLP: warps 0..15 run the gradient (4 per TMEM lane quarter).  SG_MLP_PRODUCER_WARPS > 0 adds warps that generate
// the NEXT update's Gaussian noise in the shadow of the GEMM phase (all warps run the elementwise passes); measured
// in round 1 with 8 producer warps: the 85-register cap of a 768-thread CTA costs the GEMM phase more than the
// look-ahead saves (1.82 ms vs 1.34 ms per 4 pSGLD steps), so the default is 0.
#ifndef SG_MLP_PRODUCER_WARPS
#define SG_MLP_PRODUCER_WARPS 0
#endif
template <> struct BigThreads<SGMCMC_FAMILY_MLP> { static constexpr int value = 512 + 32 * SG_MLP_PRODUCER_WARPS; static constexpr int min_blocks = 1; };

struct BigModel {
  int family, estimator, diffusion, flags;
  int n_leaves;
  int leaf_off[MAXL + 1];   // element offset of each leaf in the flat parameter vector; leaf_off[n_leaves] = P
  int P;
  unsigned n_data, batch;
  int full_batch, leapfrog, update_rate;
  float scale;              // fp32(n_data) / fp32(batch)
  float hyper[4];
  int dims[8];              // PMF: {n_users, n_items, rank};  MLP: {n_layers, size_0, ..., size_L}
  float coef[4];            // PMF: {alpha, lam, 0};  MLP: {-, prior precision (1), constant of logprior = -P/2 log(2 pi)}
  const void* data0;        // PMF: int4 [N] {i, j, bits(y), 0};  MLP: float [N][size_0]
  const void* data1;        // MLP: float [N][size_L] (one-hot / soft labels)
  const float* cv_center;   // [P]  (CV only)
  const float* cv_grad;     // [P]
  int pmf_compact;          // PMF: the minibatch gradient lives compact in shared memory (PmfCompact)
  unsigned pmf_magic;       // ceil(2^32 / rank): row of element e = umulhi(e, magic)   (e < 2^24)
};

// threads per CTA of big_sampler_kernel<FAMILY> for this model
template <int FAMILY> inline int big_threads(const BigModel& m) {
  if (FAMILY == SGMCMC_FAMILY_PMF) {
    if (m.pmf_compact) return 512;
    const char* v = std::getenv("SGMCMC_PMF_THREADS");          // experiments: 128 / 256 / 384 / 512 (128 registers each)
    const int t = v ? std::atoi(v) : 0;
    return (t >= 64 && t <= 512 && t % 32 == 0) ? t : 256;
  }
  return BigThreads<FAMILY>::value;
}

struct BigRun {
  sgmcmc_state st;
  int i0, K;
  const sgmcmc_step_coef* coef;
  int coef_stride;
  float* samples;
  size_t sample_chain_stride;
  int thin;                  // keep every thin-th state of the launch (0 / 1: all), sgmcmc_run_thinned
  float* moments;            // [C][3][P]: sum and sum of squares of (kept state - plane 2), or nullptr
  const uint32_t* step_keys;
  int mode;                  // 0 run, 1 init_fn, 2 init + sampler split, 3 full-batch gradient at fg_thetas
  const uint32_t* init_keys;
  const float* fg_thetas;    // [M, P] (mode 3)
  float* fg_out;             // [M, P]
  float* noise;              // [C, P] scratch for noise generated one update ahead (families with producer warps), or nullptr
  float* lp_out;             // [C][K] log-posterior values of the optimiser iterations (SGMCMC_DIFF_ADAM_OPT), or nullptr
  // in-kernel work queue (more chains than resident CTAs): sched[0] = next item, sched[1 + c] = blocks of chain c done,
  // sched[1 + C + c] = its gradient buffer holds raw sums, sched[1 + 2 C + c] = their scale (float bits)
  int* sched;                // nullptr: one CTA per chain, all K steps (grid = n_chains)
  int blk;                   // steps per work item
};

// Work for the noise-producer warps during a likelihood pass: normal(split(split(key)[0], n_leaves)[leaf], (n_leaf,))
// for every leaf into dst[P] - the noise of the langevin step that will run with `key`.
struct NoiseJob { Key key; float* dst; bool valid; };

struct BigSmem {
  uint32_t keys[2 * (4 + 2 * MAXL)];
  float alpha[MAXL];
  float red[32];
  float lp;                  // sum of the minibatch log-likelihood VALUES of the current estimate (optimiser only)
  int item;                  // work item claimed by thread 0 (queued launches)
};

template <int FAMILY> struct FamilyGrad;   // lik_pass(...) specialisations below / in mlp_grad.cuh

// The chain-state pointers travel through structs and lambdas, where nvcc loses their address space and emits generic
// LD / ST; telling it they are global memory gives LDG / STG back (non-null pointers only).
#define SG_GLOBAL_PTR(p) __builtin_assume(__isGlobal(p))
#define SG_SHARED_PTR(p) __builtin_assume(__isShared(p))
// ... for the members of BigCtx, at the top of every non-inlined method.  PMF only: measured -3.5 % on its sampler
// (SGNHT 1.077 -> 1.039 ms per 512 chains x 4 steps), +2 % on the MLP's, whose passes schedule worse with them.
#define SG_CTX_SPACES()                                                                       \
  do {                                                                                        \
    if (FAMILY != SGMCMC_FAMILY_PMF) break;                                                   \
    SG_GLOBAL_PTR(theta); SG_GLOBAL_PTR(grad); SG_SHARED_PTR(&s);                             \
    if (aux1) SG_GLOBAL_PTR(aux1);                                                            \
    if (aux2) SG_GLOBAL_PTR(aux2);                                                            \
    if (center) SG_GLOBAL_PTR(center);                                                        \
    if (fb) SG_GLOBAL_PTR(fb);                                                                \
    if (svrg_center) SG_GLOBAL_PTR(svrg_center);                                              \
    if (svrg_fb) SG_GLOBAL_PTR(svrg_fb);                                                      \
  } while (0)


// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float big_block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < W; ++w) t += red[w];
  return t;
}

// One element of a leaf in flight: x, up to two auxiliary values and the (finalised) gradient.
struct Elt { float x, a, b, g, fb, cen; };   // fb / cen: full-batch gradient and centre of the CV / SVRG estimators

// For every element e of an n-element leaf: v = ld(e) ... st(e, v, xi) with xi = normal(key, (n,))[e].
// One threefry block feeds elements m and m + ceil(n/2) (jax's half-split counter layout).  U blocks per
// thread are staged per iteration so that 2U independent sets of global loads are in flight: these passes
// run one CTA per chain and are latency-bound otherwise.
template <int U, class LD, class ST>
__device__ __forceinline__ void leaf_pass(unsigned n, Key key, LD ld, ST st) {
  const unsigned h = (n + 1u) >> 1;
  const unsigned hf = n - h;                     // blocks m < hf have both elements m and m + h in range
  const unsigned bd = blockDim.x;
  unsigned m0 = threadIdx.x;
  // main loop: no guards, so the 2U loads of an iteration issue back to back
#pragma unroll 1
  for (; m0 + (U - 1) * bd < hf; m0 += U * bd) {
    Elt e0[U], e1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { e0[u] = ld(m0 + u * bd); e1[u] = ld(m0 + u * bd + h); }
    // all U threefry blocks first, as independent chains side by side (ILP for the 20 dependent rounds of each), and
    // before the first use of a loaded value: the RNG work hides the memory latency of the loads above
    uint32_t w0[U], w1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) stream_block(key, m0 + u * bd, n, w0[u], w1[u]);
    float z0[U], z1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { z0[u] = bits_to_normal(w0[u]); z1[u] = bits_to_normal(w1[u]); }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned m = m0 + u * bd;
      st(m, e0[u], z0[u]);
      st(m + h, e1[u], z1[u]);
    }
  }
#pragma unroll 1
  for (; m0 < h; m0 += bd) {                     // tail (and the unpaired middle element of an odd leaf)
    const bool two = m0 + h < n;
    const Elt a = ld(m0);
    const Elt b = two ? ld(m0 + h) : Elt{};
    uint32_t w0, w1;
    stream_block(key, m0, n, w0, w1);
    st(m0, a, bits_to_normal(w0));
    if (two) st(m0 + h, b, bits_to_normal(w1));
  }
}

// Same staging without noise: v = ld(e) ... st(e, v).
template <int U, class LD, class ST>
__device__ __forceinline__ void plain_pass(unsigned n, LD ld, ST st) {
  const unsigned bd = blockDim.x;
  unsigned e0 = threadIdx.x;
#pragma unroll 1
  for (; e0 + (2 * U - 1) * bd < n; e0 += 2 * U * bd) {
    Elt v[2 * U];
#pragma unroll
    for (int u = 0; u < 2 * U; ++u) v[u] = ld(e0 + u * bd);
#pragma unroll
    for (int u = 0; u < 2 * U; ++u) st(e0 + u * bd, v[u]);
  }
#pragma unroll 1
  for (; e0 < n; e0 += bd) { const Elt v = ld(e0); st(e0, v); }
}

// Elementwise pass with the noise read from a buffer filled one update ahead: v = ld(e) ... st(e, v, xi[e]).
template <int U, class LD, class ST>
__device__ __forceinline__ void noise_pass(unsigned n, const float* xi, LD ld, ST st) {
  const unsigned bd = blockDim.x;
  unsigned e0 = threadIdx.x;
#pragma unroll 1
  for (; e0 + (2 * U - 1) * bd < n; e0 += 2 * U * bd) {
    Elt v[2 * U];
    float z[2 * U];
#pragma unroll
    for (int u = 0; u < 2 * U; ++u) { v[u] = ld(e0 + u * bd); z[u] = xi[e0 + u * bd]; }
#pragma unroll
    for (int u = 0; u < 2 * U; ++u) st(e0 + u * bd, v[u], z[u]);
  }
#pragma unroll 1
  for (; e0 < n; e0 += bd) { const Elt v = ld(e0); const float z = xi[e0]; st(e0, v, z); }
}

// Pair versions: a thread takes the ADJACENT blocks 2p and 2p + 1, i.e. elements (2p, 2p + 1) and (2p + h, 2p + 1 + h), so
// that every array is read and written 8 bytes at a time (half the load / store instructions and address arithmetic of
// the scalar passes: these passes are bound by their instruction streams, not by bandwidth).  Needs h even and 8-byte aligned
// arrays (the caller checks); the blocks past the last full pair go through the scalar tail.
//   ld2(e, t0, t1): elements e, e + 1;   f(e, t, xi) -> Out;   st2(e, o0, o1)
struct Out { float x, a, b; };                   // new x, new first / second auxiliary value
template <int U, class LD2, class F, class ST2, class LD, class ST>
__device__ __forceinline__ void leaf_pass2(unsigned n, Key key, LD2 ld2, F f, ST2 st2, LD ld, ST st) {
  const unsigned h = (n + 1u) >> 1;
  const unsigned hf = n - h;
  const unsigned np = hf >> 1;                   // full pairs of blocks
  const unsigned bd = blockDim.x;
  unsigned p0 = threadIdx.x;
#pragma unroll 1
  for (; p0 + (U - 1) * bd < np; p0 += U * bd) {
    Elt a0[U], a1[U], b0[U], b1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { const unsigned e = 2u * (p0 + u * bd); ld2(e, a0[u], a1[u]); ld2(e + h, b0[u], b1[u]); }
    uint32_t w[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned e = 2u * (p0 + u * bd);
      stream_block(key, e, n, w[u][0], w[u][2]);          // block e: elements e, e + h
      stream_block(key, e + 1u, n, w[u][1], w[u][3]);     // block e + 1: elements e + 1, e + 1 + h
    }
    float z[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int k = 0; k < 4; ++k) z[u][k] = bits_to_normal(w[u][k]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned e = 2u * (p0 + u * bd);
      st2(e, f(e, a0[u], z[u][0]), f(e + 1u, a1[u], z[u][1]));
      st2(e + h, f(e + h, b0[u], z[u][2]), f(e + h + 1u, b1[u], z[u][3]));
    }
  }
#pragma unroll 1
  for (; p0 < np; p0 += bd) {
    const unsigned e = 2u * p0;
    Elt a0, a1, b0, b1;
    ld2(e, a0, a1); ld2(e + h, b0, b1);
    uint32_t w0, w1, w2, w3;
    stream_block(key, e, n, w0, w2);
    stream_block(key, e + 1u, n, w1, w3);
    st2(e, f(e, a0, bits_to_normal(w0)), f(e + 1u, a1, bits_to_normal(w1)));
    st2(e + h, f(e + h, b0, bits_to_normal(w2)), f(e + h + 1u, b1, bits_to_normal(w3)));
  }
#pragma unroll 1
  for (unsigned m0 = 2u * np + threadIdx.x; m0 < h; m0 += bd) {   // blocks past the last full pair (scalar)
    const bool two = m0 + h < n;
    const Elt a = ld(m0);
    const Elt b = two ? ld(m0 + h) : Elt{};
    uint32_t w0, w1;
    stream_block(key, m0, n, w0, w1);
    st(m0, a, bits_to_normal(w0));
    if (two) st(m0 + h, b, bits_to_normal(w1));
  }
}
// ... and without noise: elements (2q, 2q + 1)
template <int U, class LD2, class F, class ST2, class LD, class ST>
__device__ __forceinline__ void plain_pass2(unsigned n, LD2 ld2, F f, ST2 st2, LD ld, ST st) {
  const unsigned nq = n >> 1, bd = blockDim.x;
  unsigned q0 = threadIdx.x;
#pragma unroll 1
  for (; q0 + (2 * U - 1) * bd < nq; q0 += 2 * U * bd) {
    Elt t0[2 * U], t1[2 * U];
#pragma unroll
    for (int u = 0; u < 2 * U; ++u) ld2(2u * (q0 + u * bd), t0[u], t1[u]);
#pragma unroll
    for (int u = 0; u < 2 * U; ++u) { const unsigned e = 2u * (q0 + u * bd); st2(e, f(e, t0[u]), f(e + 1u, t1[u])); }
  }
#pragma unroll 1
  for (; q0 < nq; q0 += bd) { const unsigned e = 2u * q0; Elt t0, t1; ld2(e, t0, t1); st2(e, f(e, t0), f(e + 1u, t1)); }
  if ((n & 1u) && threadIdx.x == 0) { const Elt t = ld(n - 1u); st(n - 1u, t); }
}

// Producer side: the threads [t0, t0 + nt) of the CTA fill dst[0, n) with normal(key, (n,)).
__device__ __forceinline__ void produce_leaf_noise(unsigned n, Key key, float* dst, unsigned t, unsigned nt) {
  const unsigned h = (n + 1u) >> 1;
  for (unsigned m = t; m < h; m += nt) {
    uint32_t w0, w1;
    stream_block(key, m, n, w0, w1);
    dst[m] = bits_to_normal(w0);
    if (m + h < n) dst[m + h] = bits_to_normal(w1);
  }
}

#ifndef SG_BIG_UNROLL_PMF
#define SG_BIG_UNROLL_PMF 4
#endif
template <int FAMILY> struct BigUnroll { static constexpr int value = SG_BIG_UNROLL_PMF; };  // 256 threads, 2 CTAs per SM: 128 registers
#ifndef SG_BIG_UNROLL_MLP
#define SG_BIG_UNROLL_MLP 4
#endif
template <> struct BigUnroll<SGMCMC_FAMILY_MLP> { static constexpr int value = SG_BIG_UNROLL_MLP; };       // 512 threads: 128 registers

// ---------------------------------------------------------------------------------------
// PMF: compact minibatch gradient in shared memory.
//   ri / rj / e / ec [B]   the minibatch in stream order n: user, item, alpha (y - U_i.V_j) at theta and at the CV centre
//   head[nu + ni]          per factor row (U rows first): the last rating that joined its list, -1 = untouched
//   nxt[2][B]              list links (side 0: by user, side 1: by item)
//   slot[nu + ni]          compact slot of a touched row, -1 = untouched (the diffusion pass reads 0)
//   gc[2 B][r]             raw gradient rows  sum_n e_n other_row(n) [- ec_n other_centre_row(n)], summed in stream order
// Every touched row has exactly one owner - the rating at the head of its list - which walks the list in increasing n,
// so the sum order is fixed: the result is deterministic (unlike float atomics) and costs no read-modify-write traffic.
struct PmfCompact {
  float *gc, *e, *ec;
  int *head, *nslots;
  short *slot, *nxt;
  unsigned short *ri, *rj;
};
__host__ __device__ inline size_t pmf_compact_bytes(int nu, int ni, int r, unsigned batch) {
  size_t b = (size_t)2 * batch * r * 4 + (size_t)2 * batch * 4;          // gc, e, ec
  b += (size_t)(nu + ni) * 4 + 16;                                         // head, nslots
  b += (size_t)(nu + ni) * 2 + (size_t)2 * batch * 2 + (size_t)2 * batch * 2 + 16;   // slot, nxt, ri, rj
  return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline bool pmf_compact_ok(int nu, int ni, int r, unsigned batch, int full_batch, size_t smem_limit) {
  return !full_batch && batch <= 16383u && nu <= 65535 && ni <= 65535 && (r == 8 || r == 16 || r == 20 || r == 32) &&
         pmf_compact_bytes(nu, ni, r, batch) + 1024 <= smem_limit;
}
__device__ __forceinline__ PmfCompact pmf_compact_view(void* fam_smem, int nu, int ni, int r, unsigned batch) {
  PmfCompact c;
  float* f = reinterpret_cast<float*>(fam_smem);
  c.gc = f;                       f += (size_t)2 * batch * r;
  c.e = f;                        f += batch;
  c.ec = f;                       f += batch;
  int* ip = reinterpret_cast<int*>(f);
  c.head = ip;                    ip += nu + ni;
  c.nslots = ip;                  ip += 4;
  short* sp = reinterpret_cast<short*>(ip);
  c.slot = sp;                    sp += (nu + ni + 1) & ~1;
  c.nxt = sp;                     sp += 2 * batch;
  c.ri = reinterpret_cast<unsigned short*>(sp);   sp += batch;
  c.rj = reinterpret_cast<unsigned short*>(sp);
  return c;
}

// ---------------------------------------------------------------------------------------
// PMF likelihood pass
// One rating (i, j, y) at rank 4 R4: g[U_i] += e V_j, g[V_j] += e U_i with e = alpha (y - U_i.V_j); the factor rows sit in
// registers, every load of the rating issues back to back, accumulation by float4 L2 atomics.  The control-variate
// term (minus the same at the centre) is a second round so that the live registers stay at 2 R4 float4.
template <int R4>
__device__ __forceinline__ float pmf_rating_half(const float* th, float* g, int offV, int i, int j, float y, float alpha, float sign) {
  const int r = 4 * R4;
  const float4* u = reinterpret_cast<const float4*>(th + (size_t)i * r);
  const float4* v = reinterpret_cast<const float4*>(th + offV + (size_t)j * r);
  float4 uu[R4], vv[R4];
#pragma unroll
  for (int k = 0; k < R4; ++k) { uu[k] = u[k]; vv[k] = v[k]; }
  float dot = 0.f;
#pragma unroll
  for (int k = 0; k < R4; ++k) {
    dot = fmaf(uu[k].x, vv[k].x, dot); dot = fmaf(uu[k].y, vv[k].y, dot);
    dot = fmaf(uu[k].z, vv[k].z, dot); dot = fmaf(uu[k].w, vv[k].w, dot);
  }
  const float e = sign * alpha * (y - dot);
  float4* gu = reinterpret_cast<float4*>(g + (size_t)i * r);
  float4* gv = reinterpret_cast<float4*>(g + offV + (size_t)j * r);
#pragma unroll
  for (int k = 0; k < R4; ++k) {
    atomicAdd(gu + k, make_float4(e * vv[k].x, e * vv[k].y, e * vv[k].z, e * vv[k].w));
    atomicAdd(gv + k, make_float4(e * uu[k].x, e * uu[k].y, e * uu[k].z, e * uu[k].w));
  }
  return y - dot;
}
// returns the residual y - U_i . V_j at theta (the log-likelihood value is -alpha/2 residual^2)
template <int R4>
__device__ __forceinline__ float pmf_rating(const float* theta, const float* center, float* g, int offV, int i, int j, float y,
                                            float alpha) {
  const float res = pmf_rating_half<R4>(theta, g, offV, i, j, y, alpha, 1.0f);
  if (center) pmf_rating_half<R4>(center, g, offV, i, j, y, alpha, -1.0f);
  return res;
}

template <>
struct FamilyGrad<SGMCMC_FAMILY_PMF> {
  static size_t smem_bytes(const BigModel& m) {
    return m.pmf_compact ? pmf_compact_bytes(m.dims[0], m.dims[1], m.dims[2], m.batch) : 0;
  }
  static constexpr bool kNoiseProducers = false;
  __device__ static void cta_init(const BigModel& m, void* fam_smem) {
    if (!m.pmf_compact) return;
    const PmfCompact c = pmf_compact_view(fam_smem, m.dims[0], m.dims[1], m.dims[2], m.batch);
    for (int t = threadIdx.x; t < m.dims[0] + m.dims[1]; t += blockDim.x) { c.head[t] = -1; c.slot[t] = -1; }
    if (threadIdx.x == 0) *c.nslots = 0;
    __syncthreads();
  }
  __device__ static void cta_fini(const BigModel&, void*) {}

  // alpha (y - U_i . V_j) for one rating; factor rows as R4 float4 each
  template <int R4>
  __device__ __forceinline__ static float resid(const float* th, int offV, int i, int j, float y, float alpha) {
    const int r = 4 * R4;
    const float4* u = reinterpret_cast<const float4*>(th + (size_t)i * r);
    const float4* v = reinterpret_cast<const float4*>(th + offV + (size_t)j * r);
    float4 uu[R4], vv[R4];
#pragma unroll
    for (int k = 0; k < R4; ++k) { uu[k] = u[k]; vv[k] = v[k]; }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < R4; ++k) {
      dot = fmaf(uu[k].x, vv[k].x, dot); dot = fmaf(uu[k].y, vv[k].y, dot);
      dot = fmaf(uu[k].z, vv[k].z, dot); dot = fmaf(uu[k].w, vv[k].w, dot);
    }
    return alpha * (y - dot);
  }

  // owner of a touched row: sum its list in increasing stream position into one compact slot
  template <int R4>
  __device__ __forceinline__ static void owner_row(const PmfCompact& c, const float* theta, const float* center, int offV,
                                                   int side, int n, int row_slot, unsigned batch) {
    const int r = 4 * R4;
    const short* nxt = c.nxt + (size_t)side * batch;
    float4 acc[R4];
#pragma unroll
    for (int k = 0; k < R4; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    int last = -1;
    for (;;) {                                              // lists are short (B / rows ~ 1): selection in stream order
      int best = 0x7fffffff;
      for (int mm = n; mm >= 0; mm = nxt[mm]) best = (mm > last && mm < best) ? mm : best;
      if (best == 0x7fffffff) break;
      last = best;
      // d/dU_i = e V_j, d/dV_j = e U_i: the OTHER side's row of rating `best`
      const size_t off = side ? (size_t)c.ri[best] * r : (size_t)offV + (size_t)c.rj[best] * r;
      const float4* o = reinterpret_cast<const float4*>(theta + off);
      const float ev = c.e[best];
#pragma unroll
      for (int k = 0; k < R4; ++k) {
        const float4 x = o[k];
        acc[k].x = fmaf(ev, x.x, acc[k].x); acc[k].y = fmaf(ev, x.y, acc[k].y);
        acc[k].z = fmaf(ev, x.z, acc[k].z); acc[k].w = fmaf(ev, x.w, acc[k].w);
      }
      if (center) {                                         // control variate: minus the same at the centre
        const float4* oc = reinterpret_cast<const float4*>(center + off);
        const float ecv = -c.ec[best];
#pragma unroll
        for (int k = 0; k < R4; ++k) {
          const float4 x = oc[k];
          acc[k].x = fmaf(ecv, x.x, acc[k].x); acc[k].y = fmaf(ecv, x.y, acc[k].y);
          acc[k].z = fmaf(ecv, x.z, acc[k].z); acc[k].w = fmaf(ecv, x.w, acc[k].w);
        }
      }
    }
    float4* dst = reinterpret_cast<float4*>(c.gc + (size_t)row_slot * r);
#pragma unroll
    for (int k = 0; k < R4; ++k) dst[k] = acc[k];
  }

  // Minibatch estimate into the compact buffer.  Phase 0 forgets the previous minibatch's rows; phase A draws the
  // ratings, takes the residuals and links every rating into the lists of its two rows; phase C: the head of each list
  // owns the row, takes a slot and sums the list.
  template <int R4>
  __device__ static void compact_pass(const BigModel& m, const PmfCompact& c, const float* theta, const float* center,
                                      const RandintPlan& plan, float* lp) {
    const int r = 4 * R4, nu = m.dims[0], offV = m.leaf_off[1];
    const unsigned B = m.batch, h = (B + 1u) >> 1;
    const int4* __restrict__ rat = reinterpret_cast<const int4*>(m.data0);
    const float alpha = m.coef[0];
    SG_GLOBAL_PTR(theta); SG_GLOBAL_PTR(rat);
    if (center) SG_GLOBAL_PTR(center);
    (void)r;
    if (*c.nslots > 0) {                                    // rows of the previous minibatch (still in ri / rj)
      for (unsigned n = threadIdx.x; n < B; n += blockDim.x) {
        const int ru = c.ri[n], rv = nu + c.rj[n];
        c.head[ru] = -1; c.slot[ru] = -1; c.head[rv] = -1; c.slot[rv] = -1;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) *c.nslots = 0;
    float lpv = 0.f;
    for (unsigned b = threadIdx.x; b < h; b += blockDim.x) {
      unsigned i0, i1;
      randint_block(plan, b, i0, i1);
      const bool two = b + h < B;
      const int4 q0 = __ldg(rat + i0);
      const int4 q1 = __ldg(rat + (two ? i1 : i0));
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (t == 1 && !two) break;
        const int4 q = t ? q1 : q0;
        const unsigned n = t ? b + h : b;
        const float y = __int_as_float(q.z);
        const float ev = resid<R4>(theta, offV, q.x, q.y, y, alpha);
        c.e[n] = ev;
        if (center) c.ec[n] = resid<R4>(center, offV, q.x, q.y, y, alpha);
        c.ri[n] = (unsigned short)q.x; c.rj[n] = (unsigned short)q.y;
        c.nxt[n] = (short)atomicExch(c.head + q.x, (int)n);
        c.nxt[B + n] = (short)atomicExch(c.head + nu + q.y, (int)n);
        lpv = fmaf(-0.5f * ev, ev / alpha, lpv);            // -alpha/2 (y - U_i.V_j)^2
      }
    }
    if (lp) {
#pragma unroll
      for (int o = 16; o; o >>= 1) lpv += __shfl_xor_sync(FULL, lpv, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(lp, lpv);
    }
    __syncthreads();
    for (unsigned cnd = threadIdx.x; cnd < 2u * B; cnd += blockDim.x) {
      const int side = cnd >= B, n = (int)(side ? cnd - B : cnd);
      const int row = side ? nu + c.rj[n] : c.ri[n];
      if (c.head[row] != n) continue;
      const int sl = atomicAdd(c.nslots, 1);
      c.slot[row] = (short)sl;
      owner_row<R4>(c, theta, center, offV, side, n, sl, B);
    }
  }

  __device__ static void lik_pass(const BigModel& m, void* fam_smem, const float* theta, const float* center,
                                  bool sequential, const RandintPlan& plan, unsigned seq_base, unsigned count, float* g,
                                  bool /*fresh*/, const NoiseJob& /*job*/, float* lp = nullptr, bool compact = false) {
    const int r = m.dims[2];
    if (compact) {                                   // minibatch of a model with the compact gradient (BigCtx::estimate)
      const PmfCompact c = pmf_compact_view(fam_smem, m.dims[0], m.dims[1], r, m.batch);
      switch (r >> 2) {
        case 2: compact_pass<2>(m, c, theta, center, plan, lp); break;
        case 4: compact_pass<4>(m, c, theta, center, plan, lp); break;
        case 5: compact_pass<5>(m, c, theta, center, plan, lp); break;
        default: compact_pass<8>(m, c, theta, center, plan, lp); break;
      }
      return;
    }
    float lpv = 0.f;                                 // sum of -alpha/2 (y - U_i.V_j)^2 over this thread's ratings (if lp)
    const int offV = m.leaf_off[1];
    const int4* __restrict__ rat = reinterpret_cast<const int4*>(m.data0);
    SG_GLOBAL_PTR(theta); SG_GLOBAL_PTR(g); SG_GLOBAL_PTR(rat);
    if (center) SG_GLOBAL_PTR(center);
    const float alpha = m.coef[0];
    const unsigned h = (count + 1u) >> 1;
    // rank <= 32, rows 16-byte aligned (rank % 4 == 0 and aligned chain rows): whole factor rows in registers, all
    // loads of a rating issued back to back, float4 atomics
    const bool vec = (r == 8 || r == 16 || r == 20 || r == 32) && (offV & 3) == 0 && ((reinterpret_cast<uintptr_t>(theta) | reinterpret_cast<uintptr_t>(g) |
                      reinterpret_cast<uintptr_t>(center ? center : theta)) & 15) == 0;
    const int r4 = r >> 2;
    for (unsigned b = threadIdx.x; b < h; b += blockDim.x) {
      unsigned idx[2];
      if (sequential) { idx[0] = seq_base + b; idx[1] = seq_base + b + h; } else { randint_block(plan, b, idx[0], idx[1]); }
      const int cnt = (b + h < count) ? 2 : 1;
      if (vec) {
        const int4 q0 = __ldg(rat + idx[0]);
        const int4 q1 = __ldg(rat + idx[cnt - 1]);
        for (int t = 0; t < cnt; ++t) {
          const int4 q = t ? q1 : q0;
          const float y = __int_as_float(q.z);
          float res;
          switch (r4) {
            case 2: res = pmf_rating<2>(theta, center, g, offV, q.x, q.y, y, alpha); break;
            case 4: res = pmf_rating<4>(theta, center, g, offV, q.x, q.y, y, alpha); break;
            case 5: res = pmf_rating<5>(theta, center, g, offV, q.x, q.y, y, alpha); break;
            default: res = pmf_rating<8>(theta, center, g, offV, q.x, q.y, y, alpha); break;
          }
          lpv = fmaf(-0.5f * alpha * res, res, lpv);
        }
        continue;
      }
      for (int t = 0; t < cnt; ++t) {
        const int4 q = __ldg(rat + idx[t]);
        const float y = __int_as_float(q.z);
        const float* u = theta + (size_t)q.x * r;
        const float* v = theta + offV + (size_t)q.y * r;
        float dot = 0.f;
        for (int k = 0; k < r; ++k) dot = fmaf(u[k], v[k], dot);
        const float e = alpha * (y - dot);
        lpv = fmaf(-0.5f * e, y - dot, lpv);
        float ec = 0.f;
        const float *uc = nullptr, *vc = nullptr;
        if (center) {
          uc = center + (size_t)q.x * r; vc = center + offV + (size_t)q.y * r;
          float dc = 0.f;
          for (int k = 0; k < r; ++k) dc = fmaf(uc[k], vc[k], dc);
          ec = alpha * (y - dc);
        }
        float* gu = g + (size_t)q.x * r;
        float* gv = g + offV + (size_t)q.y * r;
        for (int k = 0; k < r; ++k) {
          float du = e * v[k], dv = e * u[k];
          if (center) { du -= ec * vc[k]; dv -= ec * uc[k]; }
          atomicAdd(gu + k, du);
          atomicAdd(gv + k, dv);
        }
      }
    }
    if (lp) {
#pragma unroll
      for (int o = 16; o; o >>= 1) lpv += __shfl_xor_sync(FULL, lpv, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(lp, lpv);
    }
  }
};

// ---------------------------------------------------------------------------------------
template <int FAMILY>
struct BigCtx {
  static constexpr int U = BigUnroll<FAMILY>::value;
  const BigModel& m;
  BigSmem& s;
  void* fam_smem;
  float *theta, *aux1, *aux2, *grad;      // this chain's rows of the state arrays
  const float *center, *fb;                // CV: shared model buffers; SVRG: this chain's rows
  float *svrg_center, *svrg_fb;
  // Between an estimate and the next diffusion update the gradient buffer holds the RAW likelihood sum
  // S = sum_i grad loglik_i(theta) [- the same at the centre]; the consumer applies  gscale * S - prior * theta
  // [+ fb + prior * centre] on the fly and zeroes the buffer for the next accumulation (no separate finalise
  // and zero passes over the P-vector).  At launch boundaries the buffer holds the finalised param_grad.
  bool grad_raw = false;
  bool raw_compact = false;   // PMF: the raw sums are the compact shared-memory rows (PmfCompact), not the gradient buffer
  PmfCompact pc{};
  float gscale = 1.0f;
  float* lp_acc = nullptr;    // shared-memory accumulator of the minibatch log-likelihood values (optimiser), or nullptr
  float* noise = nullptr;     // this chain's row of the look-ahead noise buffer (nullptr: generate in place)
  bool noise_ready = false;   // noise[] holds the draws of the NEXT langevin step

  __device__ BigCtx(const BigModel& mm, BigSmem& ss, void* fs) : m(mm), s(ss), fam_smem(fs) {
    if (FAMILY == SGMCMC_FAMILY_PMF && mm.pmf_compact) pc = pmf_compact_view(fs, mm.dims[0], mm.dims[1], mm.dims[2], mm.batch);
  }

  // raw likelihood sum of global element ge from the compact rows: 0 for a row the minibatch did not touch
  __device__ __forceinline__ float compact_raw(unsigned ge) const {
    const unsigned row = __umulhi(ge, m.pmf_magic);
    const int sl = pc.slot[row];
    return sl >= 0 ? pc.gc[(unsigned)sl * (unsigned)m.dims[2] + (ge - row * (unsigned)m.dims[2])] : 0.f;
  }

  __device__ __forceinline__ bool is_cv() const { return m.estimator != SGMCMC_EST_STANDARD; }

  // finalised gradient of global element ge from the buffer value `raw` and the parameter value x it was taken at
  __device__ __forceinline__ float gfin(float raw, float x, unsigned ge) const {
    if (!grad_raw) return raw;
    const float prior = m.coef[1];
    if (is_cv()) return fb[ge] + (gscale * raw - prior * (x - center[ge]));   // gradient_estimation.py:72, no cancellation
    return gscale * raw - prior * x;
  }

  // Gradient consumption in the diffusion passes, specialised at compile time so that the loaders are straight-line
  // batches of independent loads (a branch between two loads serialises them on their memory latency):
  //   MODE 0: the buffer holds the finalised gradient; 1: raw sum, standard estimator; 2: raw sum, CV / SVRG;
  //   3 / 4: as 1 / 2 with the raw sum in the compact shared-memory rows (PMF).
  template <int MODE>
  __device__ __forceinline__ void ld_grad(Elt& t, const float* gbuf, unsigned e, unsigned ge) const {
    if (MODE >= 3) t.g = compact_raw(ge); else t.g = __ldcg(gbuf + e);
    if (MODE == 2 || MODE == 4) { t.fb = fb[ge]; t.cen = center[ge]; }
  }
  template <int MODE>
  __device__ static __forceinline__ float gval(const Elt& t, float gs, float prior) {
    if (MODE == 0) return t.g;
    if (MODE == 2 || MODE == 4) return t.fb + (gs * t.g - prior * (t.x - t.cen));   // gradient_estimation.py:72, no cancellation
    return gs * t.g - prior * t.x;
  }

  __device__ void zero(float* g) {
    if (FAMILY == SGMCMC_FAMILY_PMF) SG_GLOBAL_PTR(g);
    plain_pass<U>((unsigned)m.P, [&](unsigned) { return Elt{}; }, [&](unsigned e, const Elt&) { g[e] = 0.f; });
    __syncthreads();
  }

  // out = grad_log_post(th) over all rows (util.py:43-44 with Ndata * mean over N rows)
  __device__ void full_grad(const float* th, float* out) {
    if (FAMILY == SGMCMC_FAMILY_PMF) { SG_GLOBAL_PTR(th); SG_GLOBAL_PTR(out); }
    zero(out);
    RandintPlan dummy{};
    FamilyGrad<FAMILY>::lik_pass(m, fam_smem, th, nullptr, true, dummy, 0u, m.n_data, out, true, NoiseJob{});
    __syncthreads();
    const float prior = m.coef[1];
    // __ldcg: the sums may have been accumulated with L2 atomics, so read them at L2
    plain_pass<U>((unsigned)m.P, [&](unsigned e) { Elt v{}; v.g = __ldcg(out + e); v.x = th[e]; return v; },
                  [&](unsigned e, const Elt& v) { out[e] = v.g - prior * v.x; });
    __syncthreads();
  }

  // turn a raw buffer into the finalised gradient (launch end, init_fn)
  __device__ void finalize() {
    if (!grad_raw) return;
    SG_CTX_SPACES();
    if (FAMILY == SGMCMC_FAMILY_PMF && raw_compact)
      plain_pass<U>((unsigned)m.P, [&](unsigned e) { Elt v{}; v.g = compact_raw(e); v.x = theta[e]; return v; },
                    [&](unsigned e, const Elt& v) { grad[e] = gfin(v.g, v.x, e); });
    else
      plain_pass<U>((unsigned)m.P, [&](unsigned e) { Elt v{}; v.g = __ldcg(grad + e); v.x = theta[e]; return v; },
                    [&](unsigned e, const Elt& v) { grad[e] = gfin(v.g, v.x, e); });
    __syncthreads();
    grad_raw = false; raw_compact = false;
  }

  // estimate_gradient: on entry the gradient buffer is all zero unless is_init
  __device__ __noinline__ void estimate(int i, Key k2, bool is_init, Key next_key = Key{0u, 0u}, bool has_next = false) {
    SG_CTX_SPACES();
    const bool cv = is_cv();
    // the producer warps generate the next step's noise during the first likelihood pass of this estimate
    NoiseJob job{next_key, noise, has_next && noise != nullptr && FamilyGrad<FAMILY>::kNoiseProducers};
    if (m.estimator == SGMCMC_EST_SVRG) {
      if (is_init) {                                       // gradient_estimation.py:124-128
        full_grad(theta, svrg_fb);
        plain_pass<U>((unsigned)m.P, [&](unsigned e) { Elt v{}; v.x = theta[e]; v.g = svrg_fb[e]; return v; },
                      [&](unsigned e, const Elt& v) { svrg_center[e] = v.x; grad[e] = v.g; });
        __syncthreads();
        grad_raw = false;
        return;
      }
      if (i % m.update_rate == 0) {                        // gradient_estimation.py:134-139
        full_grad(theta, svrg_fb);
        plain_pass<U>((unsigned)m.P, [&](unsigned e) { Elt v{}; v.x = theta[e]; return v; },
                      [&](unsigned e, const Elt& v) { svrg_center[e] = v.x; });
        __syncthreads();
      }
    }
    // PMF minibatch estimates of a compact model never touch the dense buffer (it need not be zero either)
    const bool cmp = FAMILY == SGMCMC_FAMILY_PMF && m.pmf_compact;
    if (is_init && !cmp) zero(grad);
    RandintPlan plan{};
    if (m.estimator == SGMCMC_EST_STANDARD && m.full_batch && !is_init) {
      // static full-batch bypass, gradient_estimation.py:41-42 (init_gradient still subsamples, :31-35)
      FamilyGrad<FAMILY>::lik_pass(m, fam_smem, theta, nullptr, true, plan, 0u, m.n_data, grad, true, job, lp_acc);
      gscale = 1.0f;
    } else {
      plan = randint_plan(k2, m.n_data, m.batch);
      FamilyGrad<FAMILY>::lik_pass(m, fam_smem, theta, cv ? center : nullptr, false, plan, 0u, m.batch, grad, true, job, lp_acc, cmp);
      gscale = m.scale;
      raw_compact = cmp;
    }
    __syncthreads();
    grad_raw = true;
    noise_ready = job.valid;
    if (is_init) finalize();
  }

  __device__ __forceinline__ Key leaf_key(int slot, int leaf) const {
    Key k; k.a = s.keys[slot + 2 * leaf]; k.b = s.keys[slot + 2 * leaf + 1]; return k;
  }

  __device__ void expand_step_keys(Key key) {
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      uint32_t w = lane < 4 ? stream_word(key, lane, 4u) : 0u;
      Key k1; k1.a = __shfl_sync(FULL, w, 0); k1.b = __shfl_sync(FULL, w, 1);
      if (lane < 4) s.keys[lane] = w;
      const uint32_t nl = m.n_leaves;
      if (lane < 2 * (int)nl) s.keys[4 + lane] = stream_word(k1, lane, 2u * nl);
    }
    __syncthreads();
  }

  // phase 1 = `update` (consumes the carried gradient, leaves the buffer zeroed for the next estimate and
  // emits the sample if `sample` is set); phase 2 = `update2` of the palindromes (gradient stays in the buffer).
  __device__ void diffuse(int phase, int i, const sgmcmc_step_coef& cf, int key_slot, float* sample) {
    SG_T0();
    if (!grad_raw) diffuse_mode<0>(phase, i, cf, key_slot, sample);
    else if (FAMILY == SGMCMC_FAMILY_PMF && raw_compact) {
      if (!is_cv()) diffuse_mode<(FAMILY == SGMCMC_FAMILY_PMF ? 3 : 1)>(phase, i, cf, key_slot, sample);
      else diffuse_mode<(FAMILY == SGMCMC_FAMILY_PMF ? 4 : 2)>(phase, i, cf, key_slot, sample);
    }
    else if (!is_cv()) diffuse_mode<1>(phase, i, cf, key_slot, sample);
    else diffuse_mode<2>(phase, i, cf, key_slot, sample);
    SG_T1(0);
  }

  template <int MODE>
  __device__ __noinline__ void diffuse_mode(int phase, int i, const sgmcmc_step_coef& cf, int key_slot, float* sample) {
    const bool bug = (m.flags & SGMCMC_FLAG_BADODABCV_REFERENCE_BUG) && m.diffusion == SGMCMC_DIFF_BADODAB;
    SG_CTX_SPACES();
    // scalars the per-element code needs, read once (the model lives in param space behind a reference)
    const float gs = gscale, prior = m.coef[1];
    const sgmcmc_step_coef cfv = cf;
    const bool use_noise = noise_ready && phase == 1 && !bug;   // draws of this step were generated during the last estimate
    // the consumed buffer is zeroed for the next accumulation - unless minibatch estimates never accumulate into it
    const bool zg_any = !(FAMILY == SGMCMC_FAMILY_PMF && m.pmf_compact);
    // ... nor where the next estimate OVERWRITES it: the MLP's layer-1 weight gradient (leaf 0, 98.6 % of the elements for
    // 784-100-10) is stored, not added, by the first sub-batch's tensor-core epilogue (mlp_grad.cuh layer1_backward)
    const unsigned zfrom = FAMILY == SGMCMC_FAMILY_MLP ? (unsigned)m.leaf_off[1] : 0u;
    if (phase == 2 || bug) {                      // baoab / badodab update2: v += dt/2 * g
      const bool consume = phase == 1;
      plain_pass<U>((unsigned)m.P,
                    [&](unsigned e) { Elt t{}; t.x = theta[e]; t.a = aux1[e]; ld_grad<MODE>(t, grad, e, e); return t; },
                    [&](unsigned e, const Elt& t) {
                      const float gg = gval<MODE>(t, gs, prior);
                      aux1[e] = fmaf(cfv.hh, gg, t.a);
                      if (consume) { if (zg_any && e >= zfrom) grad[e] = 0.f; if (sample) __stcs(sample + e, t.x); }
                      else grad[e] = gg;           // keep the finalised value: the next update reads it as is
                    });
      __syncthreads();
      grad_raw = false; raw_compact = false;
      return;
    }
    for (int leaf = 0; leaf < m.n_leaves; ++leaf) {
      const Key lk = leaf_key(key_slot, leaf);
      const unsigned off = (unsigned)m.leaf_off[leaf];
      const unsigned n = (unsigned)(m.leaf_off[leaf + 1] - m.leaf_off[leaf]);
      float* x = theta + off;
      float* v = aux1 ? aux1 + off : nullptr;
      float* v2 = aux2 ? aux2 + off : nullptr;
      float* g = grad + off;
      float* smp = sample ? sample + off : nullptr;
      const bool zg = zg_any && off >= zfrom;
      if (FAMILY == SGMCMC_FAMILY_PMF) {
        SG_GLOBAL_PTR(x); SG_GLOBAL_PTR(g);
        if (v) SG_GLOBAL_PTR(v);
        if (v2) SG_GLOBAL_PTR(v2);
        if (smp) SG_GLOBAL_PTR(smp);
        if (MODE == 2) { SG_GLOBAL_PTR(fb); SG_GLOBAL_PTR(center); }
      }
      // 8-byte path (leaf_pass2 / plain_pass2): every array 8-byte aligned and the half-split offset h even
      const unsigned hh2 = (n + 1u) >> 1;
      uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(g);
      if (v) al |= reinterpret_cast<uintptr_t>(v);
      if (v2) al |= reinterpret_cast<uintptr_t>(v2);
      if (smp) al |= reinterpret_cast<uintptr_t>(smp);
      if (MODE == 2 || MODE == 4) al |= reinterpret_cast<uintptr_t>(fb + off) | reinterpret_cast<uintptr_t>(center + off);
      // PMF only: measured +5 % there (SGNHT 4.10 -> 3.88 ms per 512 chains x 16 steps), -4.6 % on the MLP sampler, whose
      // elementwise passes share their registers with the GEMM phases and schedule worse with the wider staging
      const bool vec2 = FAMILY == SGMCMC_FAMILY_PMF && (al & 7) == 0 && (hh2 & 1u) == 0 && !use_noise;
      // loaders: pure loads of x, the gradient buffer (+ CV terms) and optionally the auxiliaries; HV / HV2: with v / v2
      auto ld1 = [&](unsigned e, bool hv, bool hv2, bool hg) {
        Elt t{}; t.x = x[e];
        if (hv) t.a = v[e];
        if (hv2) t.b = v2[e];
        if (hg) ld_grad<MODE>(t, g, e, off + e);
        return t;
      };
      auto ld2v = [&](unsigned e, Elt& t0, Elt& t1, bool hv, bool hv2, bool hg) {
        t0 = Elt{}; t1 = Elt{};
        const float2 xx = *reinterpret_cast<const float2*>(x + e); t0.x = xx.x; t1.x = xx.y;
        if (hv) { const float2 a = *reinterpret_cast<const float2*>(v + e); t0.a = a.x; t1.a = a.y; }
        if (hv2) { const float2 b = *reinterpret_cast<const float2*>(v2 + e); t0.b = b.x; t1.b = b.y; }
        if (hg) {
          if (MODE >= 3) { t0.g = compact_raw(off + e); t1.g = compact_raw(off + e + 1u); }
          else { const float2 gg = __ldcg(reinterpret_cast<const float2*>(g + e)); t0.g = gg.x; t1.g = gg.y; }
          if (MODE == 2 || MODE == 4) {
            const float2 f2 = *reinterpret_cast<const float2*>(fb + off + e), c2 = *reinterpret_cast<const float2*>(center + off + e);
            t0.fb = f2.x; t1.fb = f2.y; t0.cen = c2.x; t1.cen = c2.y;
          }
        }
      };
      // stores of the new x (+ sample), auxiliaries, and the zero that re-arms the gradient buffer
      auto st1 = [&](unsigned e, const Out& o, bool sv, bool sv2, bool z, bool sm) {
        x[e] = o.x;
        if (sv) v[e] = o.a;
        if (sv2) v2[e] = o.b;
        if (z && zg) g[e] = 0.f;
        if (sm && smp) __stcs(smp + e, o.x);
      };
      auto st2v = [&](unsigned e, const Out& o0, const Out& o1, bool sv, bool sv2, bool z, bool sm) {
        *reinterpret_cast<float2*>(x + e) = make_float2(o0.x, o1.x);
        if (sv) *reinterpret_cast<float2*>(v + e) = make_float2(o0.a, o1.a);
        if (sv2) *reinterpret_cast<float2*>(v2 + e) = make_float2(o0.b, o1.b);
        if (z && zg) *reinterpret_cast<float2*>(g + e) = make_float2(0.f, 0.f);
        if (sm && smp) __stcs(reinterpret_cast<float2*>(smp + e), make_float2(o0.x, o1.x));
      };
      const float* xi_buf = noise + off;
      // one noisy pass: f(e, t, xi) -> Out for every element with its N(0,1) draw (look-ahead buffer, or generated in
      // place); HV / HV2: the diffusion carries one / two auxiliary arrays, HG: the pass consumes the gradient
      auto noisy = [&](Key key, auto f, bool hv, bool hv2, bool hg, bool sm) {
        auto ld = [&](unsigned e) { return ld1(e, hv, hv2, hg); };
        auto st = [&](unsigned e, const Elt& t, float xi) { st1(e, f(e, t, xi), hv, hv2, hg, sm); };
        if (use_noise) { noise_pass<U>(n, xi_buf, ld, st); return; }
        if (vec2) {
          leaf_pass2<(U > 1 ? U / 2 : 1)>(n, key, [&](unsigned e, Elt& t0, Elt& t1) { ld2v(e, t0, t1, hv, hv2, hg); }, f,
                                          [&](unsigned e, const Out& o0, const Out& o1) { st2v(e, o0, o1, hv, hv2, hg, sm); }, ld, st);
          return;
        }
        leaf_pass<U>(n, key, ld, st);
      };
      switch (m.diffusion) {
        case SGMCMC_DIFF_SGLD:
          noisy(lk, [&](unsigned, const Elt& t, float xi) {
            return Out{fmaf(cfv.ca, xi, fmaf(cfv.h, gval<MODE>(t, gs, prior), t.x)), 0.f, 0.f};
          }, false, false, true, true);
          break;
        case SGMCMC_DIFF_PSGLD: {
          const float alpha = m.hyper[0], eps = m.hyper[1];
          noisy(lk, [&](unsigned, const Elt& t, float xi) {
            const float gg = gval<MODE>(t, gs, prior);
            const float vv = alpha * t.a + (1.0f - alpha) * (gg * gg);
            const float G = fast_rcp(fast_sqrt(vv) + eps);
            return Out{t.x + cfv.hh * G * gg + fast_sqrt(cfv.h * G) * xi, vv, 0.f};
          }, true, false, true, true);
          break;
        }
        case SGMCMC_DIFF_SGLDADAM: {
          const float b1 = m.hyper[0], b2 = m.hyper[1], eps = m.hyper[2];
          noisy(lk, [&](unsigned, const Elt& t, float xi) {
            const float gg = gval<MODE>(t, gs, prior);
            const float mm = b1 * t.a + (1.0f - b1) * gg;
            const float vv = b2 * t.b + (1.0f - b2) * (gg * gg);
            const float mh = mm * fast_rcp(cfv.ca), vh = vv * fast_rcp(cfv.cb);
            const float a = cfv.h * fast_rcp(fast_sqrt(vh) + eps);
            return Out{t.x + a * 0.5f * mh + fast_sqrt(a) * xi, mm, vv};
          }, true, true, true, true);
          break;
        }
        case SGMCMC_DIFF_SGHMC: {
          const float alpha = m.hyper[0];
          noisy(lk, [&](unsigned, const Elt& t, float xi) {
            return Out{t.x + t.a, t.a + cfv.h * gval<MODE>(t, gs, prior) - alpha * t.a + cfv.ca * xi, 0.f};
          }, true, false, true, true);
          break;
        }
        case SGMCMC_DIFF_BAOAB:
          noisy(lk, [&](unsigned, const Elt& t, float xi) {
            float ve = fmaf(cfv.hh, gval<MODE>(t, gs, prior), t.a);
            float xe = t.x + ve * cfv.h * 0.5f;
            ve = cfv.ca * ve + cfv.cb * xi;
            xe = xe + ve * cfv.h * 0.5f;
            return Out{xe, ve, 0.f};
          }, true, false, true, true);
          break;
        case SGMCMC_DIFF_SGNHT: {
          Key kn = lk, sub = lk;
          const float alpha = s.alpha[leaf];
          const bool first = (i == 0);
          if (first) split2(lk, kn, sub);               // diffusions.py:268-278
          float ss = 0.f;
          noisy(kn, [&](unsigned e, const Elt& t, float xi) {
            float ve = first ? cfv.cb * normal_at(sub, e, n) : t.a;
            ve = ve - alpha * ve + cfv.h * gval<MODE>(t, gs, prior) + cfv.ca * xi;
            ss = fmaf(ve, ve, ss);
            return Out{t.x + ve, ve, 0.f};
          }, true, false, true, true);
          const float nrm = sqrtf(big_block_sum(ss, s.red));
          if (threadIdx.x == 0) s.alpha[leaf] = alpha + (nrm * nrm) / (float)n - cfv.h;
          break;
        }
        case SGMCMC_DIFF_BADODAB: {
          float alpha = s.alpha[leaf];
          float ss = 0.f;
          {
            auto f = [&](unsigned, const Elt& t) {
              const float ve = fmaf(cfv.hh, gval<MODE>(t, gs, prior), t.a);
              ss = fmaf(ve, ve, ss);
              return Out{fmaf(cfv.hh, ve, t.x), ve, 0.f};
            };
            auto ld = [&](unsigned e) { return ld1(e, true, false, true); };
            auto st = [&](unsigned e, const Elt& t) { st1(e, f(e, t), true, false, true, false); };
            if (vec2)
              plain_pass2<U>(n, [&](unsigned e, Elt& t0, Elt& t1) { ld2v(e, t0, t1, true, false, true); }, f,
                             [&](unsigned e, const Out& o0, const Out& o1) { st2v(e, o0, o1, true, false, true, false); }, ld, st);
            else
              plain_pass<U>(n, ld, st);
          }
          alpha = alpha + cfv.hh * (sqrtf(big_block_sum(ss, s.red)) - (float)n);
          const float c1 = (float)exp((double)(-alpha * cfv.h));     // correctly rounded, see vec_sampler.cuh
          const float c2 = (alpha == 0.0f) ? cfv.ca : sqrtf(fabsf((1.0f - c1 * c1) / (2.0f * alpha)));
          __syncthreads();
          float ss2 = 0.f;
          noisy(lk, [&](unsigned, const Elt& t, float xi) {
            const float ve = c1 * t.a + c2 * xi;
            ss2 = fmaf(ve, ve, ss2);
            return Out{fmaf(cfv.hh, ve, t.x), ve, 0.f};           // second half step of x with the new v
          }, true, false, false, true);
          alpha = alpha + cfv.hh * (sqrtf(big_block_sum(ss2, s.red)) - (float)n);
          __syncthreads();
          if (threadIdx.x == 0) s.alpha[leaf] = alpha;
          break;
        }
      }
    }
    __syncthreads();
    grad_raw = false; raw_compact = false;        // consumed: the buffer is all zero now (or never used: compact)
    noise_ready = false;
  }

  // One iteration of optimizer.py:45-51 after estimate(): optax.adam ASCENT step on the raw minibatch sums in the
  // gradient buffer (cf.ca = 1 - b1^t, cf.cb = 1 - b2^t, cf.h = learning rate; oracle/optimizer.py), the buffer zeroed for
  // the next estimate.  Returns the log-posterior VALUE at the parameters the gradient was taken at (util.py:43-44):
  // logprior(theta) + Ndata * mean_i loglik_i, the likelihood values summed by the family's pass into s.lp.
  __device__ float adam_step(const sgmcmc_step_coef& cf) {
    SG_CTX_SPACES();
    const float gs = gscale, prior = m.coef[1];
    const float b1 = m.hyper[0], b2 = m.hyper[1], eps = m.hyper[2];
    const sgmcmc_step_coef cfv = cf;
    float sx2 = 0.f;
    const bool cmp = FAMILY == SGMCMC_FAMILY_PMF && raw_compact;
    plain_pass<U>((unsigned)m.P,
                  [&](unsigned e) {
                    Elt t{}; t.x = theta[e]; t.a = aux1[e]; t.b = aux2[e];
                    t.g = cmp ? compact_raw(e) : __ldcg(grad + e);
                    return t;
                  },
                  [&](unsigned e, const Elt& t) {
                    const float gg = gs * t.g - prior * t.x;
                    const float mm = b1 * t.a + (1.0f - b1) * gg;
                    const float vv = b2 * t.b + (1.0f - b2) * (gg * gg);
                    aux1[e] = mm; aux2[e] = vv; grad[e] = 0.f;
                    theta[e] = t.x + cfv.h * (mm / cfv.ca) / (sqrtf(vv / cfv.cb) + eps);
                    sx2 = fmaf(t.x, t.x, sx2);
                  });
    const float tot = big_block_sum(sx2, s.red);
    const float lp = m.coef[2] - 0.5f * prior * tot + gs * s.lp;
    __syncthreads();
    grad_raw = false; raw_compact = false;
    return lp;
  }

  // `next_key` / `has_next`: the key of the NEXT langevin step of this chain, if it is already known (look-ahead noise)
  __device__ void langevin(int i, Key key, const sgmcmc_step_coef& cf, float* sample, Key next_key, bool has_next) {
    expand_step_keys(key);
    Key k2; k2.a = s.keys[2]; k2.b = s.keys[3];
    diffuse(1, i, cf, 4, sample);
    estimate(i, k2, false, next_key, has_next);
    if (m.diffusion == SGMCMC_DIFF_BAOAB || m.diffusion == SGMCMC_DIFF_BADODAB) diffuse(2, i, cf, 4, nullptr);
  }

  // key of the first langevin step of the transition that runs with per-sample key `sub`
  __device__ Key first_langevin_key(Key sub) const {
    if (m.diffusion != SGMCMC_DIFF_SGHMC) return sub;
    Key k1, k2;
    split2(sub, k1, k2);
    return split_at(k2, m.leapfrog, 0);
  }

  __device__ void transition(int i, Key key, const sgmcmc_step_coef& cf, float* sample, Key next_sub, bool has_next) {
    const Key next_first = has_next ? first_langevin_key(next_sub) : Key{0u, 0u};
    if (m.diffusion != SGMCMC_DIFF_SGHMC) { langevin(i, key, cf, sample, next_first, has_next); return; }
    Key k1, k2;
    split2(key, k1, k2);
    for (int leaf = 0; leaf < m.n_leaves; ++leaf) {          // resample_momentum, diffusions.py:197-199
      const Key lk = split_at(k1, m.n_leaves, leaf);
      const int off = m.leaf_off[leaf];
      float* v = aux1 + off;
      leaf_pass<U>((unsigned)(m.leaf_off[leaf + 1] - off), lk, [&](unsigned) { return Elt{}; },
                   [&](unsigned e, const Elt&, float xi) { v[e] = cf.cb * xi; });
    }
    __syncthreads();
    for (int l = 0; l < m.leapfrog; ++l) {
      const bool last = l == m.leapfrog - 1;
      langevin(i, split_at(k2, m.leapfrog, l), cf, last ? sample : nullptr,
               last ? next_first : split_at(k2, m.leapfrog, l + 1), last ? has_next : true);
    }
  }
};

// Work distribution.  Default: CTA c owns chain c for the K steps of the launch.  With more chains than the device holds
// CTAs (256 MLP chains on 148 SMs, 512 PMF chains on 296 slots) that is two waves of which the second is 73 % full; such
// launches run PERSISTENT CTAs over a queue of (chain, block of steps) items instead: items are claimed in order
// (block-major: every chain's block b before any block b + 1), a chain's next block waits for its previous one through a
// per-chain counter (release / acquire at gpu scope; the acquire's fence also drops the SM's stale L1 lines of a chain
// that last ran here two blocks ago), and the chain state travels through the caller's global buffers exactly as it does
// between launches (finalised gradient, alpha, key).  The tail shrinks from a wave to a block: 2.0 -> ~1.75 wave times.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int FAMILY>
__global__ void __launch_bounds__(BigThreads<FAMILY>::value, BigThreads<FAMILY>::min_blocks) big_sampler_kernel(const __grid_constant__ BigModel m, const __grid_constant__ BigRun run) {
  extern __shared__ __align__(16) uint8_t big_smem_raw[];
  BigSmem& s = *reinterpret_cast<BigSmem*>(big_smem_raw);
  void* fam_smem = big_smem_raw + ((sizeof(BigSmem) + 127) & ~(size_t)127);
  const size_t P = (size_t)m.P;
  const sgmcmc_state& st = run.st;
  BigCtx<FAMILY> ctx(m, s, fam_smem);
  FamilyGrad<FAMILY>::cta_init(m, fam_smem);                 // MLP: mbarriers + TMEM allocation

  if (run.mode == 3) {                                       // K2: full-batch gradient at fg_thetas[c]
    const int c = blockIdx.x;
    ctx.full_grad(run.fg_thetas + c * P, run.fg_out + c * P);
    FamilyGrad<FAMILY>::cta_fini(m, fam_smem);
    return;
  }

  // point the context at chain c and load its per-leaf thermostat
  auto bind = [&](int c) {
    ctx.theta = st.theta + c * P;
    ctx.aux1 = st.aux1 ? st.aux1 + c * P : nullptr;
    ctx.aux2 = st.aux2 ? st.aux2 + c * P : nullptr;
    ctx.grad = st.grad + c * P;
    ctx.svrg_center = st.svrg_center ? st.svrg_center + c * P : nullptr;
    ctx.svrg_fb = st.svrg_grad ? st.svrg_grad + c * P : nullptr;
    if (m.estimator == SGMCMC_EST_SVRG) { ctx.center = ctx.svrg_center; ctx.fb = ctx.svrg_fb; }
    else { ctx.center = m.cv_center; ctx.fb = m.cv_grad; }
    ctx.noise = run.noise ? run.noise + c * P : nullptr;
    if (threadIdx.x < MAXL) {
      float a = 0.f;
      if (st.alpha && threadIdx.x < m.n_leaves)
        a = run.mode == 0 ? st.alpha[(size_t)c * m.n_leaves + threadIdx.x] : m.hyper[0];
      s.alpha[threadIdx.x] = a;
    }
    __syncthreads();
  };
  // steps [k0, k1) of the launch for chain c; leaves the finalised gradient, alpha and the scan key in the state buffers
  auto run_steps = [&](int c, int k0, int k1) {
    Key carry{0u, 0u};
    if (run.step_keys == nullptr) { carry.a = st.key[2 * c]; carry.b = st.key[2 * c + 1]; }
    for (int sidx = k0; sidx < k1; ++sidx) {
      const int i = run.i0 + sidx;
      const sgmcmc_step_coef cf = run.coef[(size_t)sidx * run.coef_stride];
      Key sub, next_sub{0u, 0u};
      bool has_next = false;
      if (m.diffusion == SGMCMC_DIFF_ADAM_OPT) {
        // optimizer.py:45-51: key, subkey = split(key); value and gradient on the minibatch drawn with subkey itself
        Key nk; split2(carry, nk, sub); carry = nk;
        if (threadIdx.x == 0) s.lp = 0.f;
        __syncthreads();
        ctx.lp_acc = &s.lp;
        ctx.estimate(i, sub, false);
        const float lp = ctx.adam_step(cf);
        if (run.lp_out && threadIdx.x == 0) run.lp_out[(size_t)c * run.K + sidx] = lp;
        continue;
      }
      if (run.step_keys) { sub.a = run.step_keys[2 * c]; sub.b = run.step_keys[2 * c + 1]; }
      else {
        Key nk; split2(carry, nk, sub); carry = nk;             // samplers.py:49
        if (FamilyGrad<FAMILY>::kNoiseProducers && sidx + 1 < k1) { Key nk2; split2(carry, nk2, next_sub); has_next = true; }   // peek at the next iteration's key
      }
      // get_params after the transition: the sample is written by the diffusion pass that produces it
      const bool thinned = run.thin > 1;
      const bool kept = !thinned || (sidx + 1) % run.thin == 0;
      const int slot = thinned ? (sidx + 1) / run.thin - 1 : sidx;
      ctx.transition(i, sub, cf, (run.samples && kept) ? run.samples + (size_t)c * run.sample_chain_stride + (size_t)slot * P : nullptr,
                     next_sub, has_next);
      if (run.moments && kept) {                                // posterior summaries: one more pass over the kept state
        __syncthreads();
        float* mo = run.moments + (size_t)c * 3 * P;
        for (size_t e = threadIdx.x; e < P; e += blockDim.x) { const float x = ctx.theta[e] - mo[2 * P + e]; mo[e] += x; mo[P + e] = fmaf(x, x, mo[P + e]); }
        __syncthreads();
      }
    }
    // the state's param_grad is the finalised gradient - at the end of the launch.  Between two blocks of a queued launch
    // the raw minibatch sums stay in the buffer (saves a pass over the P-vector per block); the next block's CTA is told
    // through two words beside the chain's done-counter.  (Compact PMF sums live in shared memory: always finalised.)
    const bool last_block = k1 >= run.K;
    const bool keep_raw = run.sched != nullptr && !last_block && ctx.grad_raw && !ctx.raw_compact;
    if (!keep_raw) ctx.finalize();
    __syncthreads();
    if (run.sched != nullptr && threadIdx.x == 0) {
      run.sched[1 + st.n_chains + c] = keep_raw ? 1 : 0;
      run.sched[1 + 2 * st.n_chains + c] = __float_as_int(ctx.gscale);
    }
    if (st.alpha && threadIdx.x < m.n_leaves) st.alpha[(size_t)c * m.n_leaves + threadIdx.x] = s.alpha[threadIdx.x];
    if (threadIdx.x == 0 && st.key && run.step_keys == nullptr) { st.key[2 * c] = carry.a; st.key[2 * c + 1] = carry.b; }
  };

  if (run.mode != 0) {
    // init_fn (kernels.py:48-51): diffusion state = (theta, zeros...), gradient from init_gradient
    const int c = blockIdx.x;
    bind(c);
    Key carry{0u, 0u};
    for (size_t e = threadIdx.x; e < P; e += blockDim.x) {
      if (ctx.aux1) ctx.aux1[e] = 0.f;
      if (ctx.aux2) ctx.aux2[e] = 0.f;
    }
    __syncthreads();
    Key k; k.a = run.init_keys[2 * c]; k.b = run.init_keys[2 * c + 1];
    if (run.mode == 2) { Key sub; split2(k, carry, sub); k = sub; }
    ctx.estimate(0, k, true);
    __syncthreads();
    if (st.alpha && threadIdx.x < m.n_leaves) st.alpha[(size_t)c * m.n_leaves + threadIdx.x] = s.alpha[threadIdx.x];
    if (threadIdx.x == 0 && st.key && run.mode == 2) { st.key[2 * c] = carry.a; st.key[2 * c + 1] = carry.b; }
  } else if (run.sched == nullptr) {
    bind(blockIdx.x);
    run_steps(blockIdx.x, 0, run.K);
  } else {
    const int C = st.n_chains, nblk = (run.K + run.blk - 1) / run.blk;
    const int n_items = C * nblk;
    for (;;) {
      __syncthreads();                                          // everyone is done with the previous item (and s.item)
      if (threadIdx.x == 0) s.item = atomicAdd(run.sched, 1);
      __syncthreads();
      const int item = s.item;
      if (item >= n_items) break;
      const int b = item / C, c = item - b * C;
      if (b > 0) {
        if (threadIdx.x == 0) {
          while (ld_acquire_gpu(run.sched + 1 + c) < b) __nanosleep(200);
          __threadfence();                                      // gpu-scope fence: the SM's L1 forgets this chain's old lines
        }
        __syncthreads();
      }
      bind(c);
      ctx.grad_raw = b > 0 && run.sched[1 + C + c] != 0;        // raw sums left by the previous block of this chain
      ctx.raw_compact = false;
      ctx.gscale = b > 0 ? __int_as_float(run.sched[1 + 2 * C + c]) : 1.0f;
      run_steps(c, b * run.blk, min(run.K, (b + 1) * run.blk));
      __syncthreads();
      if (threadIdx.x == 0) { __threadfence(); st_release_gpu(run.sched + 1 + c, b + 1); }
    }
  }
  __syncthreads();
  FamilyGrad<FAMILY>::cta_fini(m, fam_smem);
}

}  // namespace sg
