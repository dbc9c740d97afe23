// Bayesian MLP family (sgmcmcjax/models/bayesian_NN/NN_model.py:31-58) for the big-state sampler.
//
//   predict(params, x): a_0 = x;  a_l = softmax(W_l a_{l-1} + b_l) for l < L (softmax, NOT relu, :36-37);
//                       log_softmax(W_L a_{L-1} + b_L)                                      (:39-41)
//   loglikelihood = sum_k y_k predict_k (:48-50);  logprior = sum N(0,1).logpdf (:53-58) -> grad = -w
//
// lik_pass ADDS  sign * sum_i grad_theta loglik(theta, x_i, y_i)  into the gradient buffer, in sub-batches
// of <= 128 rows.  Everything is fp32 on the CUDA cores in this round (the two layer-1 GEMMs,
// Z1 = X W1^T (B x s1 x s0) and dW1 = dZ1^T X (s1 x s0 x B), are register-tiled 8x8 SGEMMs with
// shared-memory staging; 2 groups of 256 threads split K (forward) or the output columns (backward)).
// Shapes supported: n_layers <= 3, hidden / output widths <= 128, any input width.
#pragma once
#include "big_sampler.cuh"

namespace sg {

constexpr int MLP_ROWS = 128;     // sub-batch rows
constexpr int MLP_KT = 16;        // staging depth
constexpr int MLP_GROUPS = 2;     // 2 x 256 threads (512-thread CTA: 128 registers per thread for the 8x8 tiles)

inline const char* mlp_validate(const sgmcmc_config* cfg) {
  const int L = cfg->dims[0];
  if (L < 1 || L > 3) return "mlp: 1..3 layers are supported";
  if (cfg->n_leaves != 2 * L) return "mlp: leaves must be (W_1, b_1, ..., W_L, b_L)";
  for (int l = 0; l <= L; ++l)
    if (cfg->dims[1 + l] < 1) return "mlp: layer sizes must be positive";
  for (int l = 1; l <= L; ++l) {
    if (cfg->dims[1 + l] > 128) return "mlp: hidden / output widths above 128 are not supported";
    if (cfg->leaf_size[2 * (l - 1)] != cfg->dims[1 + l] * cfg->dims[l] || cfg->leaf_size[2 * (l - 1) + 1] != cfg->dims[1 + l])
      return "mlp: leaf sizes do not match the layer sizes";
  }
  if (cfg->data_dim != cfg->dims[1]) return "mlp: data[0] must have size_0 columns";
  return nullptr;
}

__host__ __device__ inline int mlp_stride(int w) { return w | 1; }   // odd row stride: conflict-free row-per-lane access

struct MlpSmem {
  int idx[MLP_ROWS];
  float ysum[MLP_ROWS];
};

// shared-memory floats: activations of layers 1..L, two ping-pong dZ buffers, staging for the 4 groups
__host__ __device__ inline size_t mlp_smem_floats(const int* dims) {
  const int L = dims[0];
  size_t f = 0;
  int wmax_even = 0, wmax_odd = 0;
  for (int l = 1; l <= L; ++l) {
    f += (size_t)MLP_ROWS * mlp_stride(dims[1 + l]);
    int& w = ((L - l) & 1) ? wmax_odd : wmax_even;
    if (dims[1 + l] > w) w = dims[1 + l];
  }
  f += (size_t)MLP_ROWS * mlp_stride(wmax_even) + (size_t)MLP_ROWS * mlp_stride(wmax_odd > 0 ? wmax_odd : 1);
  f += (size_t)MLP_GROUPS * 2 * MLP_KT * 128;
  return f;
}

// acc[i][j] += sum_{k < kt} A[k * lda + i0 + i] * B[k * 128 + j0 + j]
template <bool A_VEC>
__device__ __forceinline__ void mlp_microkernel(float (&acc)[8][8], const float* __restrict__ A, int lda, int i0,
                                                const float* __restrict__ B, int j0, int kt) {
#pragma unroll 4
  for (int k = 0; k < kt; ++k) {
    float a[8], b[8];
    if (A_VEC) {
      const float4 a0 = *reinterpret_cast<const float4*>(A + k * lda + i0);
      const float4 a1 = *reinterpret_cast<const float4*>(A + k * lda + i0 + 4);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = A[k * lda + i0 + i];
    }
    const float4 b0 = *reinterpret_cast<const float4*>(B + k * 128 + j0);
    const float4 b1 = *reinterpret_cast<const float4*>(B + k * 128 + j0 + 4);
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

template <>
struct FamilyGrad<SGMCMC_FAMILY_MLP> {
  static size_t smem_bytes(const BigModel& m) { return sizeof(MlpSmem) + 16 + mlp_smem_floats(m.dims) * sizeof(float); }

  // one sub-batch of nb <= 128 rows whose table indices are in ms.idx[0..nb)
  __device__ __noinline__ static void sub_batch(const BigModel& m, MlpSmem& ms, float* fbase, const float* th, float sign, int nb, float* g) {
    const int L = m.dims[0];
    const int* sz = m.dims + 1;                       // sz[0..L]
    const float* X = reinterpret_cast<const float*>(m.data0);
    const float* Y = reinterpret_cast<const float*>(m.data1);
    const int tid = threadIdx.x;
    const int grp = tid >> 8, gt = tid & 255;         // groups of 256 threads
    const int tx = gt & 15, ty = gt >> 4;             // 16 x 16 threads, 8 x 8 outputs each

    // ---- carve shared memory
    float* act[4];
    int astr[4];
    float* p = fbase;
    for (int l = 1; l <= L; ++l) { act[l] = p; astr[l] = mlp_stride(sz[l]); p += (size_t)MLP_ROWS * astr[l]; }
    int wmax_even = 0, wmax_odd = 0;
    for (int l = 1; l <= L; ++l) { int& w = ((L - l) & 1) ? wmax_odd : wmax_even; if (sz[l] > w) w = sz[l]; }
    float* dzbuf[2];
    dzbuf[0] = p; p += (size_t)MLP_ROWS * mlp_stride(wmax_even);
    dzbuf[1] = p; p += (size_t)MLP_ROWS * mlp_stride(wmax_odd > 0 ? wmax_odd : 1);
    p = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
    float* stA = p + (size_t)grp * 2 * MLP_KT * 128;  // [KT][128]
    float* stB = stA + MLP_KT * 128;                  // [KT][128]

    // =================================================================== layer 1 forward: Z1 = X W1^T + b1
    {
      const int s0 = sz[0], s1 = sz[1];
      const float* W1 = th + m.leaf_off[0];
      const float* b1 = th + m.leaf_off[1];
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      const int n_kt = (s0 + MLP_KT - 1) / MLP_KT;
      const int per = (n_kt + MLP_GROUPS - 1) / MLP_GROUPS;
      const int kt_lo = grp * per, kt_hi = min(n_kt, kt_lo + per);
      const int lrow = gt & 127, lhalf = gt >> 7;     // loader mapping: row, 8-float half of the KT slab
      const bool vec_ok = (s0 & 3) == 0;
      // the [C, P] state rows are only 4-byte aligned in general (P = 79,510 for 784-100-10)
      const bool vec_w = vec_ok && (reinterpret_cast<uintptr_t>(W1) & 15) == 0;
      for (int t = kt_lo; t < kt_hi; ++t) {
        const int k0 = t * MLP_KT + lhalf * 8;
        float xa[8], wb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) xa[j] = wb[j] = 0.f;
        if (lrow < nb) {
          const float* src = X + (size_t)ms.idx[lrow] * s0 + k0;
          if (vec_ok && k0 + 8 <= s0) {
            const float4 q0 = __ldg(reinterpret_cast<const float4*>(src)), q1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
            xa[0] = q0.x; xa[1] = q0.y; xa[2] = q0.z; xa[3] = q0.w; xa[4] = q1.x; xa[5] = q1.y; xa[6] = q1.z; xa[7] = q1.w;
          } else {
            for (int j = 0; j < 8; ++j) if (k0 + j < s0) xa[j] = __ldg(src + j);
          }
        }
        if (lrow < s1) {
          const float* src = W1 + (size_t)lrow * s0 + k0;
          if (vec_w && k0 + 8 <= s0) {
            const float4 q0 = *reinterpret_cast<const float4*>(src), q1 = *(reinterpret_cast<const float4*>(src) + 1);
            wb[0] = q0.x; wb[1] = q0.y; wb[2] = q0.z; wb[3] = q0.w; wb[4] = q1.x; wb[5] = q1.y; wb[6] = q1.z; wb[7] = q1.w;
          } else {
            for (int j = 0; j < 8; ++j) if (k0 + j < s0) wb[j] = src[j];
          }
        }
        asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");      // previous slab fully consumed
#pragma unroll
        for (int j = 0; j < 8; ++j) { stA[(lhalf * 8 + j) * 128 + lrow] = xa[j]; stB[(lhalf * 8 + j) * 128 + lrow] = wb[j]; }
        asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
        if (ty * 8 < nb && tx * 8 < s1) mlp_microkernel<true>(acc, stA, 128, ty * 8, stB, tx * 8, MLP_KT);
      }
      // fold the 4 split-K partials into act[1] (bias added by group 0), fixed order -> deterministic
      float* Z = act[1];
      const int zs = astr[1];
      for (int gsel = 0; gsel < MLP_GROUPS; ++gsel) {
        if (grp == gsel) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = ty * 8 + i;
            if (r < MLP_ROWS) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int c = tx * 8 + j;
                if (c < s1) Z[r * zs + c] = (gsel == 0 ? b1[c] : Z[r * zs + c]) + acc[i][j];
              }
            }
          }
        }
        __syncthreads();
      }
    }

    // =================================================================== softmax + small layers forward
    for (int l = 1; l < L; ++l) {
      float* A = act[l];
      const int as = astr[l], w = sz[l];
      if (tid < MLP_ROWS) {                             // a_l = softmax(z_l), one thread per row
        float mx = -INFINITY;
        for (int c = 0; c < w; ++c) mx = fmaxf(mx, A[tid * as + c]);
        float sum = 0.f;
        for (int c = 0; c < w; ++c) { const float e = expf(A[tid * as + c] - mx); A[tid * as + c] = e; sum += e; }
        const float inv = 1.0f / sum;
        for (int c = 0; c < w; ++c) A[tid * as + c] *= inv;
      }
      __syncthreads();
      float* Zn = act[l + 1];
      const int zs = astr[l + 1], wn = sz[l + 1];
      const float* Wn = th + m.leaf_off[2 * l];
      const float* bn = th + m.leaf_off[2 * l + 1];
      for (int o = tid; o < MLP_ROWS * wn; o += blockDim.x) {     // z_{l+1}[b][n] = a_l[b] . W[n] + b[n]
        const int b = o % MLP_ROWS, n = o / MLP_ROWS;
        float accv = 0.f;
        const float* wr = Wn + (size_t)n * w;
        for (int k = 0; k < w; ++k) accv = fmaf(A[b * as + k], wr[k], accv);
        Zn[b * zs + n] = accv + bn[n];
      }
      __syncthreads();
    }

    // =================================================================== output layer: dZ_L = y - sum(y) softmax(z_L)
    float* dz = dzbuf[0];
    int dzs = mlp_stride(wmax_even);
    {
      const int w = sz[L], zs = astr[L];
      const float* Z = act[L];
      if (tid < MLP_ROWS) {
        if (tid < nb) {
          const float* yr = Y + (size_t)ms.idx[tid] * w;
          float mx = -INFINITY, ys = 0.f;
          for (int c = 0; c < w; ++c) { mx = fmaxf(mx, Z[tid * zs + c]); ys += __ldg(yr + c); }
          float sum = 0.f;
          for (int c = 0; c < w; ++c) sum += expf(Z[tid * zs + c] - mx);
          const float inv = ys / sum;
          for (int c = 0; c < w; ++c) dz[tid * dzs + c] = __ldg(yr + c) - inv * expf(Z[tid * zs + c] - mx);
        } else {
          for (int c = 0; c < w; ++c) dz[tid * dzs + c] = 0.f;      // padded rows contribute nothing
        }
      }
      __syncthreads();
    }

    // =================================================================== backward through layers L .. 2
    int cur = 0;
    for (int l = L; l >= 2; --l) {
      const int w = sz[l], wp = sz[l - 1];
      const float* A = act[l - 1];
      const int as = astr[l - 1];
      const float* Wl = th + m.leaf_off[2 * (l - 1)];
      float* gW = g + m.leaf_off[2 * (l - 1)];
      float* gb = g + m.leaf_off[2 * (l - 1) + 1];
      for (int o = tid; o < w * wp; o += blockDim.x) {             // dW_l[n][k] = sum_b dz[b][n] a_{l-1}[b][k]
        const int k = o % wp, n = o / wp;
        float accv = 0.f;
        for (int b = 0; b < nb; ++b) accv = fmaf(dz[b * dzs + n], A[b * as + k], accv);
        gW[o] += sign * accv;
      }
      for (int n = tid; n < w; n += blockDim.x) {
        float accv = 0.f;
        for (int b = 0; b < nb; ++b) accv += dz[b * dzs + n];
        gb[n] += sign * accv;
      }
      // dA_{l-1} = dz W_l ; dZ_{l-1} = a (dA - sum_k a dA)   (softmax Jacobian)
      float* dn = dzbuf[cur ^ 1];
      const int dns = mlp_stride(cur ? wmax_even : (wmax_odd > 0 ? wmax_odd : 1));
      for (int o = tid; o < MLP_ROWS * wp; o += blockDim.x) {
        const int b = o % MLP_ROWS, k = o / MLP_ROWS;
        float accv = 0.f;
        for (int n = 0; n < w; ++n) accv = fmaf(dz[b * dzs + n], Wl[(size_t)n * wp + k], accv);
        dn[b * dns + k] = accv;
      }
      __syncthreads();
      if (tid < MLP_ROWS) {
        float dotv = 0.f;
        for (int k = 0; k < wp; ++k) dotv = fmaf(A[tid * as + k], dn[tid * dns + k], dotv);
        for (int k = 0; k < wp; ++k) dn[tid * dns + k] = tid < nb ? A[tid * as + k] * (dn[tid * dns + k] - dotv) : 0.f;
      }
      __syncthreads();
      dz = dn; dzs = dns; cur ^= 1;
    }

    // =================================================================== layer 1 backward: dW1 = dZ1^T X, db1
    {
      const int s0 = sz[0], s1 = sz[1];
      float* gW = g + m.leaf_off[0];
      float* gb = g + m.leaf_off[1];
      for (int n = tid; n < s1; n += blockDim.x) {
        float accv = 0.f;
        for (int b = 0; b < nb; ++b) accv += dz[b * dzs + n];
        gb[n] += sign * accv;
      }
      const int n_ct = (s0 + 127) / 128;                 // 128-column tiles of dW1, dealt to the 4 groups
      const bool vec_ok = (s0 & 3) == 0;
      const int lk = gt >> 4, lc = (gt & 15) * 8;        // loader mapping: batch row in the slab, 8 columns
      for (int ct = grp; ct < n_ct; ct += MLP_GROUPS) {
        const int c0 = ct * 128;
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        for (int b0 = 0; b0 < nb; b0 += MLP_KT) {
          float xv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) xv[j] = 0.f;
          if (b0 + lk < nb) {
            const float* src = X + (size_t)ms.idx[b0 + lk] * s0 + c0 + lc;
            if (vec_ok && c0 + lc + 8 <= s0) {
              const float4 q0 = __ldg(reinterpret_cast<const float4*>(src)), q1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
              xv[0] = q0.x; xv[1] = q0.y; xv[2] = q0.z; xv[3] = q0.w; xv[4] = q1.x; xv[5] = q1.y; xv[6] = q1.z; xv[7] = q1.w;
            } else {
              for (int j = 0; j < 8; ++j) if (c0 + lc + j < s0) xv[j] = __ldg(src + j);
            }
          }
          asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
          *reinterpret_cast<float4*>(stB + lk * 128 + lc) = make_float4(xv[0], xv[1], xv[2], xv[3]);
          *reinterpret_cast<float4*>(stB + lk * 128 + lc + 4) = make_float4(xv[4], xv[5], xv[6], xv[7]);
          asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
          const int kt = min(MLP_KT, nb - b0);
          // rows of the output = hidden units (ty), A operand = dZ1[b][n] read straight from its buffer
          if (ty * 8 < s1 && c0 + tx * 8 < s0) mlp_microkernel<false>(acc, dz + (size_t)b0 * dzs, dzs, ty * 8, stB, tx * 8, kt);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = ty * 8 + i;
          if (n < s1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c = c0 + tx * 8 + j;
              if (c < s0) gW[(size_t)n * s0 + c] += sign * acc[i][j];
            }
          }
        }
      }
    }
    __syncthreads();
  }

  __device__ static void lik_pass(const BigModel& m, void* fam_smem, const float* theta, const float* center,
                                  bool sequential, const RandintPlan& plan, unsigned seq_base, unsigned count, float* g) {
    MlpSmem& ms = *reinterpret_cast<MlpSmem*>(fam_smem);
    float* fbase = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(&ms + 1) + 15) & ~(uintptr_t)15);
    const unsigned h = (count + 1u) >> 1;
    for (unsigned base = 0; base < count; base += MLP_ROWS) {
      const int nb = (int)min((unsigned)MLP_ROWS, count - base);
      __syncthreads();
      if (threadIdx.x < (unsigned)nb) {
        const unsigned pos = base + threadIdx.x;
        unsigned id;
        if (sequential) {
          id = seq_base + pos;
        } else {
          unsigned i0, i1;
          randint_block(plan, pos < h ? pos : pos - h, i0, i1);
          id = pos < h ? i0 : i1;
        }
        ms.idx[threadIdx.x] = (int)id;
      }
      __syncthreads();
      sub_batch(m, ms, fbase, theta, 1.0f, nb, g);
      if (center) sub_batch(m, ms, fbase, center, -1.0f, nb, g);
    }
  }
};

}  // namespace sg
