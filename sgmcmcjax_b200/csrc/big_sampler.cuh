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
//       loglik = -alpha/2 (y - U_i.V_j)^2, prior N(0, 1/lam): sparse row gather + scatter-add.
//   MLP (models/bayesian_NN/NN_model.py:31-58): see mlp_grad.cuh.
#pragma once
#include "prng.cuh"
#include "vec_sampler.cuh"   // Key helpers, FULL, MAXL, VecRun layout reuse

namespace sg {

template <int FAMILY> struct BigThreads { static constexpr int value = 1024; };
template <> struct BigThreads<SGMCMC_FAMILY_MLP> { static constexpr int value = 512; };   // 8x8 register tiles need > 64 registers

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
  float coef[4];            // PMF: {alpha, lam};  MLP: {-, prior precision (1)}
  const void* data0;        // PMF: int4 [N] {i, j, bits(y), 0};  MLP: float [N][size_0]
  const void* data1;        // MLP: float [N][size_L] (one-hot / soft labels)
  const float* cv_center;   // [P]  (CV only)
  const float* cv_grad;     // [P]
};

struct BigRun {
  sgmcmc_state st;
  int i0, K;
  const sgmcmc_step_coef* coef;
  int coef_stride;
  float* samples;
  size_t sample_chain_stride;
  const uint32_t* step_keys;
  int mode;                  // 0 run, 1 init_fn, 2 init + sampler split, 3 full-batch gradient at fg_thetas
  const uint32_t* init_keys;
  const float* fg_thetas;    // [M, P] (mode 3)
  float* fg_out;             // [M, P]
};

struct BigSmem {
  uint32_t keys[2 * (4 + 2 * MAXL)];
  float alpha[MAXL];
  float red[32];
};

template <int FAMILY> struct FamilyGrad;   // lik_pass(...) specialisations below / in mlp_grad.cuh

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

// f(e, xi) for every element e of an n-element leaf with its N(0,1) draw xi = normal(key, (n,))[e]
template <class F>
__device__ __forceinline__ void for_leaf_normals(unsigned n, Key key, F f) {
  const unsigned h = (n + 1u) >> 1;
  for (unsigned m = threadIdx.x; m < h; m += blockDim.x) {
    uint32_t w0, w1;
    stream_block(key, m, n, w0, w1);
    f(m, bits_to_normal(w0));
    if (m + h < n) f(m + h, bits_to_normal(w1));
  }
}

// ---------------------------------------------------------------------------------------
// PMF likelihood pass
template <>
struct FamilyGrad<SGMCMC_FAMILY_PMF> {
  static constexpr size_t smem_bytes(const BigModel&) { return 0; }

  __device__ static void lik_pass(const BigModel& m, void* /*fam_smem*/, const float* theta, const float* center,
                                  bool sequential, const RandintPlan& plan, unsigned seq_base, unsigned count, float* g) {
    const int r = m.dims[2];
    const int offV = m.leaf_off[1];
    const int4* __restrict__ rat = reinterpret_cast<const int4*>(m.data0);
    const float alpha = m.coef[0];
    const unsigned h = (count + 1u) >> 1;
    for (unsigned b = threadIdx.x; b < h; b += blockDim.x) {
      unsigned idx[2];
      if (sequential) { idx[0] = seq_base + b; idx[1] = seq_base + b + h; } else { randint_block(plan, b, idx[0], idx[1]); }
      const int cnt = (b + h < count) ? 2 : 1;
      for (int t = 0; t < cnt; ++t) {
        const int4 q = __ldg(rat + idx[t]);
        const float y = __int_as_float(q.z);
        const float* u = theta + (size_t)q.x * r;
        const float* v = theta + offV + (size_t)q.y * r;
        float dot = 0.f;
        for (int k = 0; k < r; ++k) dot = fmaf(u[k], v[k], dot);
        const float e = alpha * (y - dot);
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
  }
};

// ---------------------------------------------------------------------------------------
template <int FAMILY>
struct BigCtx {
  const BigModel& m;
  BigSmem& s;
  void* fam_smem;
  float *theta, *aux1, *aux2, *grad;      // this chain's rows of the state arrays
  const float *center, *fb;                // CV: shared model buffers; SVRG: this chain's rows
  float *svrg_center, *svrg_fb;

  __device__ BigCtx(const BigModel& mm, BigSmem& ss, void* fs) : m(mm), s(ss), fam_smem(fs) {}

  __device__ void zero(float* g) {
    for (int e = threadIdx.x; e < m.P; e += blockDim.x) g[e] = 0.f;
    __syncthreads();
  }

  // out = grad_log_post(th) over all rows (util.py:43-44 with Ndata * mean over N rows)
  __device__ void full_grad(const float* th, float* out) {
    zero(out);
    RandintPlan dummy{};
    FamilyGrad<FAMILY>::lik_pass(m, fam_smem, th, nullptr, true, dummy, 0u, m.n_data, out);
    __syncthreads();
    const float prior = m.coef[1];
    // __ldcg: the sums were accumulated with L2 atomics, so read them at L2
    for (int e = threadIdx.x; e < m.P; e += blockDim.x) out[e] = __ldcg(out + e) - prior * th[e];
    __syncthreads();
  }

  __device__ void estimate(int i, Key k2, bool is_init) {
    const bool cv = m.estimator != SGMCMC_EST_STANDARD;
    if (m.estimator == SGMCMC_EST_SVRG) {
      if (is_init) {                                       // gradient_estimation.py:124-128
        full_grad(theta, svrg_fb);
        for (int e = threadIdx.x; e < m.P; e += blockDim.x) { svrg_center[e] = theta[e]; grad[e] = svrg_fb[e]; }
        __syncthreads();
        return;
      }
      if (i % m.update_rate == 0) {                        // gradient_estimation.py:134-139
        full_grad(theta, svrg_fb);
        for (int e = threadIdx.x; e < m.P; e += blockDim.x) svrg_center[e] = theta[e];
        __syncthreads();
      }
    } else if (m.estimator == SGMCMC_EST_STANDARD && m.full_batch && !is_init) {
      full_grad(theta, grad);                              // gradient_estimation.py:41-42
      return;
    }
    zero(grad);
    const RandintPlan plan = randint_plan(k2, m.n_data, m.batch);
    FamilyGrad<FAMILY>::lik_pass(m, fam_smem, theta, cv ? center : nullptr, false, plan, 0u, m.batch, grad);
    __syncthreads();
    const float prior = m.coef[1], scale = m.scale;
    if (cv) {   // fb + (g(theta) - g(centre)): gradient_estimation.py:72 without the fp32 cancellation
      for (int e = threadIdx.x; e < m.P; e += blockDim.x)
        grad[e] = fb[e] + (scale * __ldcg(grad + e) - prior * (theta[e] - center[e]));
    } else {
      for (int e = threadIdx.x; e < m.P; e += blockDim.x) grad[e] = scale * __ldcg(grad + e) - prior * theta[e];
    }
    __syncthreads();
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

  __device__ void diffuse(int phase, int i, const sgmcmc_step_coef& cf, int key_slot) {
    const bool bug = (m.flags & SGMCMC_FLAG_BADODABCV_REFERENCE_BUG) && m.diffusion == SGMCMC_DIFF_BADODAB;
    if (phase == 2 || bug) {                      // baoab / badodab update2: v += dt/2 * g
      for (int e = threadIdx.x; e < m.P; e += blockDim.x) aux1[e] = fmaf(cf.hh, grad[e], aux1[e]);
      __syncthreads();
      return;
    }
    for (int leaf = 0; leaf < m.n_leaves; ++leaf) {
      const Key lk = leaf_key(key_slot, leaf);
      const int off = m.leaf_off[leaf];
      const unsigned n = (unsigned)(m.leaf_off[leaf + 1] - off);
      float* x = theta + off;
      float* v = aux1 ? aux1 + off : nullptr;
      float* v2 = aux2 ? aux2 + off : nullptr;
      const float* g = grad + off;
      switch (m.diffusion) {
        case SGMCMC_DIFF_SGLD:
          for_leaf_normals(n, lk, [&](unsigned e, float xi) { x[e] = fmaf(cf.ca, xi, fmaf(cf.h, g[e], x[e])); });
          break;
        case SGMCMC_DIFF_PSGLD: {
          const float alpha = m.hyper[0], eps = m.hyper[1];
          for_leaf_normals(n, lk, [&](unsigned e, float xi) {
            const float ge = g[e];
            const float vv = alpha * v[e] + (1.0f - alpha) * (ge * ge);
            const float G = 1.0f / (sqrtf(vv) + eps);
            v[e] = vv;
            x[e] = x[e] + cf.hh * G * ge + sqrtf(cf.h * G) * xi;
          });
          break;
        }
        case SGMCMC_DIFF_SGLDADAM: {
          const float b1 = m.hyper[0], b2 = m.hyper[1], eps = m.hyper[2];
          for_leaf_normals(n, lk, [&](unsigned e, float xi) {
            const float ge = g[e];
            const float mm = b1 * v[e] + (1.0f - b1) * ge;
            const float vv = b2 * v2[e] + (1.0f - b2) * (ge * ge);
            const float mh = mm / cf.ca, vh = vv / cf.cb;
            const float a = cf.h / (sqrtf(vh) + eps);
            v[e] = mm; v2[e] = vv;
            x[e] = x[e] + a * 0.5f * mh + sqrtf(a) * xi;
          });
          break;
        }
        case SGMCMC_DIFF_SGHMC: {
          const float alpha = m.hyper[0];
          for_leaf_normals(n, lk, [&](unsigned e, float xi) {
            const float ve = v[e];
            x[e] += ve;
            v[e] = ve + cf.h * g[e] - alpha * ve + cf.ca * xi;
          });
          break;
        }
        case SGMCMC_DIFF_BAOAB:
          for_leaf_normals(n, lk, [&](unsigned e, float xi) {
            float ve = fmaf(cf.hh, g[e], v[e]);
            float xe = x[e] + ve * cf.h * 0.5f;
            ve = cf.ca * ve + cf.cb * xi;
            xe = xe + ve * cf.h * 0.5f;
            v[e] = ve; x[e] = xe;
          });
          break;
        case SGMCMC_DIFF_SGNHT: {
          Key kn = lk, sub = lk;
          const float alpha = s.alpha[leaf];
          const bool first = (i == 0);
          if (first) split2(lk, kn, sub);               // diffusions.py:268-278
          const unsigned h = (n + 1u) >> 1;
          float ss = 0.f;
          for (unsigned mm = threadIdx.x; mm < h; mm += blockDim.x) {
            uint32_t w0, w1, s0 = 0u, s1 = 0u;
            stream_block(kn, mm, n, w0, w1);
            if (first) stream_block(sub, mm, n, s0, s1);
            const int cnt = (mm + h < n) ? 2 : 1;
            for (int t = 0; t < cnt; ++t) {
              const unsigned e = t ? mm + h : mm;
              float ve = first ? cf.cb * bits_to_normal(t ? s1 : s0) : v[e];
              ve = ve - alpha * ve + cf.h * g[e] + cf.ca * bits_to_normal(t ? w1 : w0);
              v[e] = ve; x[e] += ve;
              ss = fmaf(ve, ve, ss);
            }
          }
          const float nrm = sqrtf(big_block_sum(ss, s.red));
          if (threadIdx.x == 0) s.alpha[leaf] = alpha + (nrm * nrm) / (float)n - cf.h;
          break;
        }
        case SGMCMC_DIFF_BADODAB: {
          float alpha = s.alpha[leaf];
          float ss = 0.f;
          for (unsigned e = threadIdx.x; e < n; e += blockDim.x) {
            const float ve = fmaf(cf.hh, g[e], v[e]);
            v[e] = ve; x[e] = fmaf(cf.hh, ve, x[e]);
            ss = fmaf(ve, ve, ss);
          }
          alpha = alpha + cf.hh * (sqrtf(big_block_sum(ss, s.red)) - (float)n);
          const float c1 = (float)exp((double)(-alpha * cf.h));     // correctly rounded, see vec_sampler.cuh
          const float c2 = (alpha == 0.0f) ? cf.ca : sqrtf(fabsf((1.0f - c1 * c1) / (2.0f * alpha)));
          __syncthreads();
          float ss2 = 0.f;
          {
            const unsigned h = (n + 1u) >> 1;
            for (unsigned mm = threadIdx.x; mm < h; mm += blockDim.x) {
              uint32_t w0, w1;
              stream_block(lk, mm, n, w0, w1);
              const int cnt = (mm + h < n) ? 2 : 1;
              for (int t = 0; t < cnt; ++t) {
                const unsigned e = t ? mm + h : mm;
                const float ve = c1 * v[e] + c2 * bits_to_normal(t ? w1 : w0);
                v[e] = ve; ss2 = fmaf(ve, ve, ss2);
              }
            }
          }
          alpha = alpha + cf.hh * (sqrtf(big_block_sum(ss2, s.red)) - (float)n);
          for (unsigned e = threadIdx.x; e < n; e += blockDim.x) x[e] = fmaf(cf.hh, v[e], x[e]);
          __syncthreads();
          if (threadIdx.x == 0) s.alpha[leaf] = alpha;
          break;
        }
      }
    }
    __syncthreads();
  }

  __device__ void langevin(int i, Key key, const sgmcmc_step_coef& cf) {
    expand_step_keys(key);
    Key k2; k2.a = s.keys[2]; k2.b = s.keys[3];
    diffuse(1, i, cf, 4);
    estimate(i, k2, false);
    if (m.diffusion == SGMCMC_DIFF_BAOAB || m.diffusion == SGMCMC_DIFF_BADODAB) diffuse(2, i, cf, 4);
  }

  __device__ void transition(int i, Key key, const sgmcmc_step_coef& cf) {
    if (m.diffusion != SGMCMC_DIFF_SGHMC) { langevin(i, key, cf); return; }
    Key k1, k2;
    split2(key, k1, k2);
    for (int leaf = 0; leaf < m.n_leaves; ++leaf) {          // resample_momentum, diffusions.py:197-199
      const Key lk = split_at(k1, m.n_leaves, leaf);
      const int off = m.leaf_off[leaf];
      float* v = aux1 + off;
      for_leaf_normals((unsigned)(m.leaf_off[leaf + 1] - off), lk, [&](unsigned e, float xi) { v[e] = cf.cb * xi; });
    }
    __syncthreads();
    for (int l = 0; l < m.leapfrog; ++l) langevin(i, split_at(k2, m.leapfrog, l), cf);
  }
};

template <int FAMILY>
__global__ void __launch_bounds__(BigThreads<FAMILY>::value, 1) big_sampler_kernel(const BigModel m, const BigRun run) {
  extern __shared__ __align__(16) uint8_t big_smem_raw[];
  BigSmem& s = *reinterpret_cast<BigSmem*>(big_smem_raw);
  void* fam_smem = big_smem_raw + ((sizeof(BigSmem) + 127) & ~(size_t)127);
  const int c = blockIdx.x;
  const size_t P = (size_t)m.P;
  const sgmcmc_state& st = run.st;
  BigCtx<FAMILY> ctx(m, s, fam_smem);

  if (run.mode == 3) {                                       // K2: full-batch gradient at fg_thetas[c]
    ctx.full_grad(run.fg_thetas + c * P, run.fg_out + c * P);
    return;
  }

  ctx.theta = st.theta + c * P;
  ctx.aux1 = st.aux1 ? st.aux1 + c * P : nullptr;
  ctx.aux2 = st.aux2 ? st.aux2 + c * P : nullptr;
  ctx.grad = st.grad + c * P;
  ctx.svrg_center = st.svrg_center ? st.svrg_center + c * P : nullptr;
  ctx.svrg_fb = st.svrg_grad ? st.svrg_grad + c * P : nullptr;
  if (m.estimator == SGMCMC_EST_SVRG) { ctx.center = ctx.svrg_center; ctx.fb = ctx.svrg_fb; }
  else { ctx.center = m.cv_center; ctx.fb = m.cv_grad; }

  if (threadIdx.x < MAXL) {
    float a = 0.f;
    if (st.alpha && threadIdx.x < m.n_leaves)
      a = run.mode == 0 ? st.alpha[(size_t)c * m.n_leaves + threadIdx.x] : m.hyper[0];
    s.alpha[threadIdx.x] = a;
  }
  __syncthreads();

  Key carry{0u, 0u};
  if (run.mode != 0) {
    // init_fn (kernels.py:48-51): diffusion state = (theta, zeros...), gradient from init_gradient
    for (size_t e = threadIdx.x; e < P; e += blockDim.x) {
      if (ctx.aux1) ctx.aux1[e] = 0.f;
      if (ctx.aux2) ctx.aux2[e] = 0.f;
    }
    __syncthreads();
    Key k; k.a = run.init_keys[2 * c]; k.b = run.init_keys[2 * c + 1];
    if (run.mode == 2) { Key sub; split2(k, carry, sub); k = sub; }
    ctx.estimate(0, k, true);
  } else {
    if (run.step_keys == nullptr) { carry.a = st.key[2 * c]; carry.b = st.key[2 * c + 1]; }
    for (int sidx = 0; sidx < run.K; ++sidx) {
      const int i = run.i0 + sidx;
      const sgmcmc_step_coef cf = run.coef[(size_t)sidx * run.coef_stride];
      Key sub;
      if (run.step_keys) { sub.a = run.step_keys[2 * c]; sub.b = run.step_keys[2 * c + 1]; }
      else { Key nk; split2(carry, nk, sub); carry = nk; }      // samplers.py:49
      ctx.transition(i, sub, cf);
      if (run.samples) {
        float* dst = run.samples + (size_t)c * run.sample_chain_stride + (size_t)sidx * P;
        for (size_t e = threadIdx.x; e < P; e += blockDim.x) __stcs(dst + e, ctx.theta[e]);
      }
    }
  }
  __syncthreads();
  if (st.alpha && threadIdx.x < m.n_leaves) st.alpha[(size_t)c * m.n_leaves + threadIdx.x] = s.alpha[threadIdx.x];
  if (threadIdx.x == 0 && st.key && (run.mode == 2 || (run.mode == 0 && run.step_keys == nullptr))) {
    st.key[2 * c] = carry.a; st.key[2 * c + 1] = carry.b;
  }
}

}  // namespace sg
