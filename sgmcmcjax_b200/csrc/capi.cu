// C-ABI entry points of libsgmcmc_b200.so (see include/sgmcmc_b200.h).
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "big_launch.cuh"
#include "k2_gemm.cuh"
#include "ksd.cuh"
#include "ksd_tc.cuh"
#include "mlp_grad.cuh"
#include "vec_launch.cuh"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(SGMCMC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));         \
  } while (0)

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}

}  // namespace

struct sgmcmc_handle {
  sgmcmc_config cfg;
  sg::VecModel vm;
  sg::BigModel bm;             // families with parameters in global memory (pmf, mlp)
  bool big = false;
  void* data0 = nullptr;       // owned (big families)
  void* data1 = nullptr;       // owned (big families)
  int P = 0, DS = 0, nch = 1;
  float* rows = nullptr;       // owned: [n_data][DS]
  float* mu = nullptr;         // owned: [DS] column means of the centred table (gaussian_mean)
  float* cv_center = nullptr;  // owned: [P]
  float* cv_grad = nullptr;    // owned: [P]
  float* partial = nullptr;    // owned scratch for sgmcmc_fullbatch_grad
  size_t partial_elems = 0;
  float* noise = nullptr;      // owned scratch [n_chains][P]: look-ahead Gaussian noise of the MLP sampler
  size_t noise_elems = 0;
  float* k2_xs = nullptr;      // owned: X~ split K-major chunk-major (tensor-core K2, built on first use)
  float* k2_xt = nullptr;      // owned: X~^T split
  float* k2_scratch = nullptr; // owned: [theta split | G | W] of the current call
  size_t k2_scratch_bytes = 0;
  int num_sms = 148;
  bool has_data = false, has_center = false;
  int* sched = nullptr;        // owned: work-queue words of the big-state sampler ([1 + 3 n_chains] ints)
  size_t sched_elems = 0;
  int big_slots = -1;          // resident CTAs of big_sampler_kernel for this model (-1: not asked yet)
  // (threads, cluster asked, extended kernel, chains) -> cluster shape the device can co-schedule (vec_launch.cuh)
  std::map<std::tuple<int, int, int, int>, int> cluster_ok;
};

namespace sg {

// x[n][D] (+ y[n]) -> rows[n][DS]; logreg folds the label into the sign: (1 - 2y) x.
__global__ void pack_rows_kernel(const float* __restrict__ x, const int* __restrict__ y, float* __restrict__ rows,
                                 size_t n, int D, int DS, int family) {
  const size_t total = n * (size_t)DS;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t r = t / DS;
    const int d = (int)(t - r * DS);
    float v = 0.f;
    if (d < D) {
      v = x[r * D + d];
      if (family == SGMCMC_FAMILY_LOGREG && y[r] != 0) v = -v;
    }
    rows[t] = v;
  }
}

// gaussian_mean: column sums of x[n][D] in double (grid-stride over rows, one atomic per block and column)
__global__ void column_sum_kernel(const float* __restrict__ x, size_t n, int D, double* __restrict__ sums) {
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    double a = 0.0;
    for (size_t r = blockIdx.x; r < n; r += gridDim.x) a += (double)x[r * D + d];
    atomicAdd(&sums[d], a);
  }
}
__global__ void column_mean_kernel(const double* __restrict__ sums, size_t n, int D, int DS, float* __restrict__ mu) {
  for (int d = threadIdx.x; d < DS; d += blockDim.x) mu[d] = d < D ? (float)(sums[d] / (double)n) : 0.f;
}
__global__ void center_rows_kernel(float* __restrict__ rows, size_t n, int D, int DS, const float* __restrict__ mu) {
  const size_t total = n * (size_t)DS;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(t % DS);
    if (d < D) rows[t] -= mu[d];
  }
}

// pmf: (ij[n][2], y[n]) -> int4 {i, j, bits(y), 0}
__global__ void pack_ratings_kernel(const int* __restrict__ ij, const float* __restrict__ y, int4* __restrict__ out, size_t n) {
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x)
    out[t] = make_int4(ij[2 * t], ij[2 * t + 1], __float_as_int(y[t]), 0);
}

// pmf: count the ratings whose (user, item) ids fall outside [0, n_users) x [0, n_items) - they would index past the
// factor matrices in FamilyGrad<PMF>::lik_pass (1-based MovieLens ids do exactly that)
__global__ void pmf_check_ids_kernel(const int* __restrict__ ij, size_t n, int nu, int ni, unsigned* __restrict__ bad) {
  unsigned mine = 0;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
    const int i = ij[2 * t], j = ij[2 * t + 1];
    mine += (i < 0 || i >= nu || j < 0 || j >= ni) ? 1u : 0u;
  }
  if (mine) atomicAdd(bad, mine);
}

// K2 stage 1: CTA (split, m) sums its row range at thetas[m] into partial[m][split][DS].
template <int FAMILY, int NCH>
__global__ void __launch_bounds__(1024 / NCH) fullgrad_partial_kernel(const VecModel m, const float* __restrict__ thetas,
                                                                      float* __restrict__ partial, unsigned rows_per_split) {
  extern __shared__ __align__(16) float smem_raw[];
  const int W = blockDim.x >> 5;
  float* s_theta = smem_raw;
  float* s_part = smem_raw + m.DS;
  const int mi = blockIdx.y, split = blockIdx.x;
  for (int e = threadIdx.x; e < m.DS; e += blockDim.x) s_theta[e] = e < m.P ? thetas[(size_t)mi * m.P + e] : 0.f;
  __syncthreads();
  const unsigned lo = split * rows_per_split;
  const unsigned hi = min(m.own_n, lo + rows_per_split);
  RandintPlan dummy{};
  row_pass<FAMILY, false, NCH>(m, true, dummy, lo, hi > lo ? hi - lo : 0u, s_theta, s_theta, s_part);
  __syncthreads();
  float* dst = partial + ((size_t)mi * gridDim.x + split) * m.DS;
  for (int d = threadIdx.x; d < m.DS; d += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < W; ++w) t += s_part[w * m.DS + d];
    dst[d] = t;
  }
}

// K2 stage 2: fixed-order fold over splits + prior / scaling (util.py:43-44 with all rows).
template <int FAMILY>
__global__ void fullgrad_finish_kernel(const VecModel m, const float* __restrict__ thetas, const float* __restrict__ partial,
                                       int splits, float* __restrict__ out) {
  const int mi = blockIdx.x;
  for (int e = threadIdx.x; e < m.P; e += blockDim.x) {
    const int leaf = e / m.leaf_dim, d = e - leaf * m.leaf_dim;
    float S = 0.f;
    for (int sp = 0; sp < splits; ++sp) S += partial[((size_t)mi * splits + sp) * m.DS + d];
    out[(size_t)mi * m.P + e] = grad_from_sum<FAMILY>(m, leaf, d, S, thetas[(size_t)mi * m.P + e], 1.0f, (float)m.own_n,
                                                      (FAMILY == SGMCMC_FAMILY_GAUSSIAN_MEAN && m.mu) ? m.mu[d] : 0.f);
  }
}

__global__ void indices_kernel(Key key, unsigned n, unsigned batch, int* out) {
  const RandintPlan plan = randint_plan(key, n, batch);
  const unsigned h = (batch + 1u) >> 1;
  for (unsigned b = blockIdx.x * blockDim.x + threadIdx.x; b < h; b += gridDim.x * blockDim.x) {
    unsigned i0, i1;
    randint_block(plan, b, i0, i1);
    out[b] = (int)i0;
    if (b + h < batch) out[b + h] = (int)i1;
  }
}

__global__ void normal_kernel(Key key, unsigned n, float* out) {
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) out[e] = normal_at(key, e, n);
}

// random.uniform(key, (n,)) on [0, 1): 23 mantissa bits of each word
__global__ void uniform_kernel(Key key, unsigned n, float* out) {
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x)
    out[e] = __uint_as_float((stream_word(key, e, n) >> 9) | 0x3F800000u) - 1.0f;
}

// vmap(random.bernoulli)(split(key, n), p) of models/logistic_regression.py:59-61: row i draws
// uniform(split(key, n)[i], ()) < p[i]; a shape-() draw is word 0 of the block of counters (0, pad 0)
__global__ void bernoulli_rows_kernel(Key key, unsigned n, const float* __restrict__ p, int* __restrict__ out) {
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const Key ki = split_at(key, n, i);
    uint32_t w0 = 0u, w1 = 0u;
    threefry(ki, w0, w1);
    const float u = __uint_as_float((w0 >> 9) | 0x3F800000u) - 1.0f;
    out[i] = u < p[i] ? 1 : 0;
  }
}

__global__ void split_kernel(Key key, unsigned num, uint32_t* out) {
  for (unsigned e = blockIdx.x * blockDim.x + threadIdx.x; e < 2u * num; e += gridDim.x * blockDim.x)
    out[e] = stream_word(key, e, 2u * num);
}

}  // namespace sg

namespace {

struct VecShape { int warps, cluster; };     // warps per CTA, CTAs per chain (a thread-block cluster if > 1)

template <int FAMILY, int CV, int NCH, bool EXT>
int launch_vec_kernel(sgmcmc_handle* h, const sg::VecRun& run, VecShape shape, cudaStream_t stream) {
  const int PP = ((h->P > h->DS ? h->P : h->DS) + 3) & ~3;
  const size_t smem = sg::vec_smem_bytes(PP, h->DS, shape.warps, shape.cluster);
  const auto key = std::make_tuple(shape.warps * 32, shape.cluster, EXT ? 1 : 0, run.st.n_chains);
  const auto hit = h->cluster_ok.find(key);
  int used = hit != h->cluster_ok.end() ? hit->second : shape.cluster;
  CUDA_TRY((sg::launch_vec_inst<FAMILY, CV, NCH, EXT>(h->vm, run, shape.warps * 32, used, smem, stream,   // csrc/vec_inst_*.cu
                                                      hit == h->cluster_ok.end(), &used)));
  h->cluster_ok[key] = used;
  ++g_launches;
  return SGMCMC_OK;
}

template <int FAMILY, int CV, int NCH>
int launch_vec(sgmcmc_handle* h, const sg::VecRun& run, VecShape shape, cudaStream_t stream) {
  // the extended instantiation carries the row-sharded mode and the optimiser; the plain samplers run the lean one
  const bool ext = h->vm.own_n != h->vm.n_data || h->cfg.diffusion == SGMCMC_DIFF_ADAM_OPT || run.sum_in || run.sum_out || run.moments;
  return ext ? launch_vec_kernel<FAMILY, CV, NCH, true>(h, run, shape, stream)
             : launch_vec_kernel<FAMILY, CV, NCH, false>(h, run, shape, stream);
}

template <int FAMILY, int CV>
int launch_vec_nch(sgmcmc_handle* h, const sg::VecRun& run, VecShape shape, cudaStream_t stream) {
  switch (h->nch) {
    case 1: return launch_vec<FAMILY, CV, 1>(h, run, shape, stream);
    case 2: return launch_vec<FAMILY, CV, 2>(h, run, shape, stream);
    case 4: return launch_vec<FAMILY, CV, 4>(h, run, shape, stream);
  }
  return fail(SGMCMC_E_UNSUPPORTED, "data_dim > 512 is not supported by the vector-family sampler");
}

// Launch shape of the vector sampler.  The device holds num_sms x 32 resident warps of this kernel (64 registers per
// thread); a chain gets its share of them, but no more warps than the minibatch has groups of 8 rows.  Up to 32 / nch
// warps fit one CTA, and one such CTA per chain already saturates HBM from 128 SMs (measured: 128 chains of config 2,
// 1 x 32 warps 0.91-0.96 of the copy peak, clusters of 2 x 16 / 4 x 8 the same or less).  Only when there are fewer
// chains than HALF the SMs does a chain get a thread-block cluster (csrc/vec_sampler.cuh header): num_sms / n_chains
// CTAs, at most 16 (README config, one chain: 8.3 us per step in one CTA, 4.3 us as 16 CTAs x 8 warps).
// SGMCMC_WARPS_PER_CHAIN (per CTA) / SGMCMC_CLUSTER override.
VecShape choose_shape(const sgmcmc_handle* h, int n_chains) {
  const int max_warps = 32 / h->nch;
  if (n_chains < 1) n_chains = 1;
  const long share = (long)h->num_sms * 32 / n_chains;                 // warps this chain can have
  const long useful = std::max<long>(4, (long)h->vm.batch / 8);       // one full 8-row group per warp
  const long want = std::min(share, useful);
  int w = env_int("SGMCMC_WARPS_PER_CHAIN", 0), cl = env_int("SGMCMC_CLUSTER", 0);
  if (cl <= 0) {
    cl = std::min(16, h->num_sms / n_chains);
    if (cl < 2 || want <= max_warps) cl = 1;
  }
  if (w <= 0) {
    const long per_cta = (want + cl - 1) / cl;
    w = 4;
    if (cl > 1) { while (w < per_cta) w *= 2; }                        // clusters: round up (each CTA sits alone on its SM)
    else { while (w * 2 <= per_cta) w *= 2; }                          // one CTA per chain: round down (CTAs share SMs)
  }
  if (w > max_warps) w = max_warps;
  if (w < 1) w = 1;
  if (cl > 16) cl = 16;
  return VecShape{w, cl};
}

int dispatch_vec(sgmcmc_handle* h, const sg::VecRun& run, cudaStream_t stream) {
#ifdef SG_NO_VEC   // experiment builds of the big-state sampler only (tools/variant.py): skip the 36 vector instantiations
  return fail(SGMCMC_E_UNSUPPORTED, "built with SG_NO_VEC");
#else
  const VecShape warps = choose_shape(h, run.st.n_chains);
  const bool cv = h->cfg.estimator != SGMCMC_EST_STANDARD;
  if (h->cfg.family == SGMCMC_FAMILY_GAUSSIAN_MEAN)
    return cv ? launch_vec_nch<SGMCMC_FAMILY_GAUSSIAN_MEAN, 1>(h, run, warps, stream)
              : launch_vec_nch<SGMCMC_FAMILY_GAUSSIAN_MEAN, 0>(h, run, warps, stream);
  if (h->cfg.family == SGMCMC_FAMILY_LOGREG) {
    if (h->vm.zc_f4 >= 0) return launch_vec_nch<SGMCMC_FAMILY_LOGREG, 2>(h, run, warps, stream);   // fixed centre: z_c column
    return cv ? launch_vec_nch<SGMCMC_FAMILY_LOGREG, 1>(h, run, warps, stream)
              : launch_vec_nch<SGMCMC_FAMILY_LOGREG, 0>(h, run, warps, stream);
  }
  return fail(SGMCMC_E_UNSUPPORTED, "family not implemented by this build");
#endif
}

template <int FAMILY>
int launch_big(sgmcmc_handle* h, const sg::BigRun& run_in, int blocks, cudaStream_t stream) {
  const size_t smem = ((sizeof(sg::BigSmem) + 127) & ~(size_t)127) + sg::FamilyGrad<FAMILY>::smem_bytes(h->bm);
  sg::BigRun run = run_in;
  // Sampling launches with more chains than resident CTAs: persistent CTAs over a queue of (chain, block of steps)
  // items (csrc/big_sampler.cuh).  SGMCMC_BIG_QUEUE=0 disables it, SGMCMC_BIG_GRID / SGMCMC_BIG_BLOCKS force a grid /
  // a number of blocks per chain (tests run few chains on a tiny grid so that chains do migrate between CTAs).
  if (run.mode == 0 && run.step_keys == nullptr && run.K > 1 && env_int("SGMCMC_BIG_QUEUE", 1)) {
    if (h->big_slots < 0) h->big_slots = sg::big_resident_ctas<FAMILY>(h->bm, smem, h->num_sms);
    int grid = env_int("SGMCMC_BIG_GRID", 0);
    if (grid <= 0) grid = h->big_slots;
    int nblk = env_int("SGMCMC_BIG_BLOCKS", 0);
    if (nblk <= 0) nblk = 4;
    if (nblk > run.K) nblk = run.K;
    if (grid > 0 && grid <= h->big_slots && blocks > grid) {
      const size_t need = (size_t)3 * blocks + 1;
      if (need > h->sched_elems) {
        if (h->sched) CUDA_TRY(cudaFree(h->sched));
        h->sched = nullptr; h->sched_elems = 0;
        CUDA_TRY(cudaMalloc(&h->sched, need * sizeof(int)));
        h->sched_elems = need;
      }
      CUDA_TRY(cudaMemsetAsync(h->sched, 0, need * sizeof(int), stream));
      run.sched = h->sched;
      run.blk = (run.K + nblk - 1) / nblk;
      if (run.thin > 1) run.blk = (run.blk + run.thin - 1) / run.thin * run.thin;   // whole thinning periods per item
      blocks = grid;
    }
  }
  CUDA_TRY(sg::launch_big_inst<FAMILY>(h->bm, run, blocks, smem, stream));                        // csrc/big_inst_*.cu
  ++g_launches;
  return SGMCMC_OK;
}

int dispatch_big(sgmcmc_handle* h, const sg::BigRun& run, int blocks, cudaStream_t stream) {
#ifndef SG_NO_PMF
  if (h->cfg.family == SGMCMC_FAMILY_PMF) return launch_big<SGMCMC_FAMILY_PMF>(h, run, blocks, stream);
#endif
#ifndef SG_NO_MLP
  if (h->cfg.family == SGMCMC_FAMILY_MLP) return launch_big<SGMCMC_FAMILY_MLP>(h, run, blocks, stream);
#endif
  return fail(SGMCMC_E_UNSUPPORTED, "family not implemented by this build");
}

int check_state(const sgmcmc_handle* h, const sgmcmc_state* st) {
  if (!h || !st) return fail(SGMCMC_E_INVALID, "null handle or state");
  if (!h->has_data) return fail(SGMCMC_E_STATE, "sgmcmc_set_data has not been called");
  if (st->n_chains <= 0) return fail(SGMCMC_E_INVALID, "n_chains must be positive");
  if (!st->theta || !st->grad) return fail(SGMCMC_E_INVALID, "state.theta and state.grad are required");
  const int d = h->cfg.diffusion;
  if (d != SGMCMC_DIFF_SGLD && !st->aux1) return fail(SGMCMC_E_INVALID, "state.aux1 is required by this diffusion");
  if ((d == SGMCMC_DIFF_SGLDADAM || d == SGMCMC_DIFF_ADAM_OPT) && !st->aux2)
    return fail(SGMCMC_E_INVALID, "state.aux2 is required by sgldAdam / the Adam optimiser");
  if ((d == SGMCMC_DIFF_SGNHT || d == SGMCMC_DIFF_BADODAB) && !st->alpha)
    return fail(SGMCMC_E_INVALID, "state.alpha is required by sgnht / badodab");
  if (h->cfg.estimator == SGMCMC_EST_SVRG && (!st->svrg_center || !st->svrg_grad))
    return fail(SGMCMC_E_INVALID, "state.svrg_center / svrg_grad are required by the SVRG estimator");
  if (h->cfg.estimator == SGMCMC_EST_CV && !h->has_center)
    return fail(SGMCMC_E_STATE, "sgmcmc_set_centering has not been called");
  return SGMCMC_OK;
}

template <int FAMILY>
int fullgrad_impl(sgmcmc_handle* h, const float* thetas, int m, float* out, cudaStream_t stream) {
  const int warps = 8 / (h->nch > 2 ? 2 : 1);
  int splits = (h->num_sms * 4 + m - 1) / m;
  const unsigned min_rows = 512;
  const unsigned max_splits = (h->vm.own_n + min_rows - 1) / min_rows;
  if ((unsigned)splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  const unsigned rows_per_split = (h->vm.own_n + splits - 1) / splits;
  const size_t need = (size_t)m * splits * h->DS;
  if (need > h->partial_elems) {
    if (h->partial) CUDA_TRY(cudaFree(h->partial));
    h->partial = nullptr; h->partial_elems = 0;
    CUDA_TRY(cudaMalloc(&h->partial, need * sizeof(float)));
    h->partial_elems = need;
  }
  const size_t smem = sizeof(float) * (size_t)(h->DS + warps * h->DS);
  dim3 grid(splits, m);
  switch (h->nch) {
    case 1: sg::fullgrad_partial_kernel<FAMILY, 1><<<grid, warps * 32, smem, stream>>>(h->vm, thetas, h->partial, rows_per_split); break;
    case 2: sg::fullgrad_partial_kernel<FAMILY, 2><<<grid, warps * 32, smem, stream>>>(h->vm, thetas, h->partial, rows_per_split); break;
    case 4: sg::fullgrad_partial_kernel<FAMILY, 4><<<grid, warps * 32, smem, stream>>>(h->vm, thetas, h->partial, rows_per_split); break;
    default: return fail(SGMCMC_E_UNSUPPORTED, "data_dim > 512 is not supported");
  }
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  sg::fullgrad_finish_kernel<FAMILY><<<m, 128, 0, stream>>>(h->vm, thetas, h->partial, splits, out);
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  return SGMCMC_OK;
}

int encode_rows16(CUtensorMap* map, float* base, size_t rows, int box_rows) {
  namespace kt = sg::ksdtc;
  kt::EncodeTiledFn encode = kt::encode_tiled_fn();
  if (!encode) return fail(SGMCMC_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)kt::KC, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)kt::KC * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kt::KC, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(SGMCMC_E_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)cr));
  return SGMCMC_OK;
}

// Tensor-core K2 for the logistic-regression family (csrc/k2_gemm.cuh).
int fullgrad_tc(sgmcmc_handle* h, const float* thetas, int m, float* out, cudaStream_t s) {
  namespace k2 = sg::k2;
  namespace kt = sg::ksdtc;
  const int d = h->cfg.data_dim, P = h->P, ds = h->DS;
  const size_t n = (size_t)h->cfg.n_data;
  const int dp = kt::padded_dim(d), chunks = k2::feat_chunks(d), dn = k2::feat_rows(d);
  const size_t nic_all = (n + kt::KC - 1) / kt::KC;
  const int blocks = h->num_sms * 8;
  if (!h->k2_xs) {
    CUDA_TRY(cudaMalloc(&h->k2_xs, (size_t)2 * chunks * n * kt::KC * sizeof(float)));
    k2::split_rows_kernel<<<blocks, 256, 0, s>>>(h->rows, n, ds, d, chunks, h->k2_xs);
    CUDA_TRY(cudaGetLastError()); ++g_launches;
  }
  if (!h->k2_xt) {
    CUDA_TRY(cudaMalloc(&h->k2_xt, (size_t)2 * nic_all * dn * kt::KC * sizeof(float)));
    k2::transpose_rows_kernel<<<blocks, 256, 0, s>>>(h->rows, n, ds, d, dn, h->k2_xt);
    CUDA_TRY(cudaGetLastError()); ++g_launches;
  }
  // Blocking: a block is mb parameter vectors x nc rows whose intermediate W = -sigmoid(Theta X~^T) (hi + lo: 8 bytes per
  // (theta, row) pair) fits SGMCMC_K2_W_MB.  Default 1 GB: all parameter vectors, as many rows as fit.  Small blocks whose W
  // stays in L2 between the two kernels (no HBM round trip) were measured and LOSE: 48 MB blocks 485 ms, 96 MB 357 ms,
  // 256 MB 250 ms against 199-206 ms at 1 GB for M = 5e4, N = 1e6 - two launches of 576-thread / 196 KB / 512-TMEM-column
  // kernels per block cost ~30 us of launch, prologue and drain against ~25 us of work.  Only a persistent fused kernel
  // would collect the L2 residency.  nc = 1024 rows (8 row tiles) if that leaves mb >= 128, all rows of the theta block
  // otherwise; both multiples of 128 so that only the last blocks are ragged.
  const size_t w_budget = (size_t)std::max(1, env_int("SGMCMC_K2_W_MB", 1024)) << 20;
  size_t nc_max = 1024, mb_max = (w_budget / (8 * nc_max)) & ~(size_t)127;
  if (mb_max < 128) mb_max = 128;
  if (mb_max >= (size_t)m) {                       // every parameter vector in one block: spend the budget on rows
    mb_max = ((size_t)m + 127) & ~(size_t)127;
    nc_max = (w_budget / (8 * mb_max)) & ~(size_t)127;
    if (nc_max < 128) nc_max = 128;
  }
  if (nc_max > n) nc_max = (n + 127) & ~(size_t)127;
  const size_t nic_max = nc_max / kt::KC;
  const size_t ts_bytes = kt::align_up((size_t)2 * chunks * m * kt::KC * sizeof(float), 256);
  const size_t g_bytes = kt::align_up((size_t)m * P * sizeof(float), 256);
  const size_t w_bytes = kt::align_up((size_t)2 * nic_max * mb_max * kt::KC * sizeof(float), 256);
  if (ts_bytes + g_bytes + w_bytes > h->k2_scratch_bytes) {
    if (h->k2_scratch) CUDA_TRY(cudaFree(h->k2_scratch));
    h->k2_scratch = nullptr; h->k2_scratch_bytes = 0;
    CUDA_TRY(cudaMalloc(&h->k2_scratch, ts_bytes + g_bytes + w_bytes));
    h->k2_scratch_bytes = ts_bytes + g_bytes + w_bytes;
  }
  float* Ts = h->k2_scratch;
  float* G = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(h->k2_scratch) + ts_bytes);
  float* W = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(h->k2_scratch) + ts_bytes + g_bytes);
  k2::split_thetas_kernel<<<std::min(blocks, (int)((size_t)m * chunks * kt::KC / 256 + 1)), 256, 0, s>>>(thetas, m, P, d, chunks, Ts);
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  CUDA_TRY(cudaMemsetAsync(G, 0, (size_t)m * P * sizeof(float), s));

  CUtensorMap tmap, xmap, xtmap, wmap;
  int rc;
  if ((rc = encode_rows16(&tmap, Ts, (size_t)2 * chunks * m, kt::TILE)) != SGMCMC_OK) return rc;
  if ((rc = encode_rows16(&xmap, h->k2_xs, (size_t)2 * chunks * n, kt::TILE)) != SGMCMC_OK) return rc;
  if ((rc = encode_rows16(&xtmap, h->k2_xt, (size_t)2 * nic_all * dn, dn)) != SGMCMC_OK) return rc;
  const size_t smem_b = k2::smem_b(dn);
  CUDA_TRY(cudaFuncSetAttribute(k2::k2_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k2::SMEM_A));
  CUDA_TRY(cudaFuncSetAttribute(k2::k2_accum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
  for (size_t m0 = 0; m0 < (size_t)m; m0 += mb_max) {
    const int mb = (int)std::min(mb_max, (size_t)m - m0);
    const int MT = (mb + kt::TILE - 1) / kt::TILE;
    for (size_t i0 = 0; i0 < n; i0 += nc_max) {
      const int nc = (int)std::min(nc_max, n - i0);
      const int nic = (nc + kt::KC - 1) / kt::KC;
      const long long tiles = (long long)MT * ((nc + kt::TILE - 1) / kt::TILE);
      if ((rc = encode_rows16(&wmap, W, (size_t)2 * nic * mb, kt::TILE)) != SGMCMC_OK) return rc;
      k2::k2_scores_kernel<<<(int)std::min<long long>(tiles, h->num_sms), k2::THREADS, k2::SMEM_A, s>>>(
          tmap, xmap, m, (int)m0, mb, (int)n, chunks, dp, (int)i0, nc, W);
      CUDA_TRY(cudaGetLastError()); ++g_launches;
      int ksplit = std::max(1, std::min(nic, 2 * h->num_sms / MT));
      k2::k2_accum_kernel<<<MT * ksplit, k2::THREADS, smem_b, s>>>(wmap, xtmap, mb, P, d, dn, nic, (int)nic_all, (int)(i0 / kt::KC), ksplit,
                                                                  G + m0 * (size_t)P);
      CUDA_TRY(cudaGetLastError()); ++g_launches;
    }
  }
  k2::finish_kernel<<<std::min(blocks, (int)((size_t)m * P / 256 + 1)), 256, 0, s>>>(G, thetas, m, P, h->vm.prior2[0], out);
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  return SGMCMC_OK;
}

}  // namespace

extern "C" {

const char* sgmcmc_last_error(void) { return g_err.c_str(); }
int sgmcmc_abi_version(void) { return SGMCMC_ABI_VERSION; }
int64_t sgmcmc_launch_count(void) { return (int64_t)g_launches.load(); }

int sgmcmc_create(const sgmcmc_config* cfg, sgmcmc_handle** out) {
  if (!cfg || !out) return fail(SGMCMC_E_INVALID, "null argument");
  if (cfg->abi_version != SGMCMC_ABI_VERSION) return fail(SGMCMC_E_INVALID, "ABI version mismatch");
  if (cfg->family < SGMCMC_FAMILY_GAUSSIAN_MEAN || cfg->family > SGMCMC_FAMILY_PMF)
    return fail(SGMCMC_E_UNSUPPORTED, "unknown model family");
  if (cfg->estimator < 0 || cfg->estimator > SGMCMC_EST_SVRG) return fail(SGMCMC_E_INVALID, "bad estimator");
  if (cfg->diffusion < 0 || cfg->diffusion > SGMCMC_DIFF_ADAM_OPT) return fail(SGMCMC_E_INVALID, "bad diffusion");
  if (cfg->diffusion == SGMCMC_DIFF_ADAM_OPT && (cfg->estimator != SGMCMC_EST_STANDARD || (int64_t)cfg->batch > cfg->n_data))
    return fail(SGMCMC_E_UNSUPPORTED, "the optimiser runs on the standard minibatch estimator");
  if (cfg->n_data <= 0 || cfg->n_data > 0x7fffffffLL) return fail(SGMCMC_E_INVALID, "n_data out of range");
  if (cfg->batch <= 0) return fail(SGMCMC_E_INVALID, "batch must be positive");
  if (cfg->n_leaves < 1 || cfg->n_leaves > SGMCMC_MAX_LEAVES) return fail(SGMCMC_E_INVALID, "n_leaves out of range");
  if (cfg->data_dim < 1) return fail(SGMCMC_E_INVALID, "data_dim must be positive");
  if (cfg->diffusion == SGMCMC_DIFF_SGHMC && cfg->leapfrog < 1) return fail(SGMCMC_E_INVALID, "sghmc needs leapfrog >= 1");
  if (cfg->estimator == SGMCMC_EST_SVRG && cfg->update_rate < 1) return fail(SGMCMC_E_INVALID, "SVRG needs update_rate >= 1");
  if (cfg->family == SGMCMC_FAMILY_PMF || cfg->family == SGMCMC_FAMILY_MLP) {
    long long P = 0;
    for (int l = 0; l < cfg->n_leaves; ++l) {
      if (cfg->leaf_size[l] < 1) return fail(SGMCMC_E_INVALID, "leaf sizes must be positive");
      P += cfg->leaf_size[l];
    }
    if (P > 0x7fffffffLL) return fail(SGMCMC_E_INVALID, "parameter vector too large");
    if (cfg->family == SGMCMC_FAMILY_PMF) {
      const int nu = cfg->dims[0], ni = cfg->dims[1], r = cfg->dims[2];
      if (cfg->n_leaves != 2 || nu < 1 || ni < 1 || r < 1 || cfg->leaf_size[0] != nu * r || cfg->leaf_size[1] != ni * r)
        return fail(SGMCMC_E_INVALID, "pmf needs leaves (U[n_users, rank], V[n_items, rank]) and dims {n_users, n_items, rank}");
      if (cfg->data_dim != 2) return fail(SGMCMC_E_INVALID, "pmf data[0] is int32 [n, 2]");
    } else {
      const char* err = sg::mlp_validate(cfg);
      if (err) return fail(SGMCMC_E_INVALID, err);
    }
    sgmcmc_handle* h = new sgmcmc_handle();
    h->cfg = *cfg;
    h->big = true;
    h->P = (int)P;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) {
      int sms = 0;
      if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) h->num_sms = sms;
    }
    sg::BigModel& bm = h->bm;
    std::memset(&bm, 0, sizeof(bm));
    bm.family = cfg->family; bm.estimator = cfg->estimator; bm.diffusion = cfg->diffusion; bm.flags = cfg->flags;
    bm.n_leaves = cfg->n_leaves; bm.P = h->P;
    for (int l = 0; l < cfg->n_leaves; ++l) bm.leaf_off[l + 1] = bm.leaf_off[l] + cfg->leaf_size[l];
    bm.n_data = (unsigned)cfg->n_data; bm.batch = (unsigned)cfg->batch;
    bm.full_batch = (cfg->estimator == SGMCMC_EST_STANDARD && (int64_t)cfg->batch == cfg->n_data) ? 1 : 0;
    bm.leapfrog = cfg->leapfrog; bm.update_rate = cfg->update_rate;
    bm.scale = (float)cfg->n_data / (float)cfg->batch;
    for (int k = 0; k < 4; ++k) bm.hyper[k] = cfg->hyper[k];
    for (int k = 0; k < 8; ++k) bm.dims[k] = cfg->dims[k];
    bm.coef[0] = cfg->lik_coef[0];
    bm.coef[1] = cfg->prior_coef[0];
    // constant of the log-prior VALUE (optimiser only): N(0, 1) log-densities carry -1/2 log(2 pi) per entry (NN_model.py:53-58)
    bm.coef[2] = cfg->family == SGMCMC_FAMILY_MLP ? (float)(-0.5 * 1.8378770664093453 * (double)P) : 0.0f;
    if (cfg->family == SGMCMC_FAMILY_PMF) {
      // compact minibatch gradient in shared memory (csrc/big_sampler.cuh PmfCompact) when it fits beside the 227 KB limit
      const size_t limit = (size_t)227 * 1024 - ((sizeof(sg::BigSmem) + 127) & ~(size_t)127);
      // Opt-in (SGMCMC_PMF_COMPACT=1): bit-reproducible gradients, but measured SLOWER on C4 (4.63 vs 3.93 ms per
      // 512 chains x 16 SGNHT steps): its 190 KB leave one CTA per SM, so the latency-bound gather phases no longer
      // overlap a second chain's issue-bound diffusion pass, and the slot lookup adds instructions to that pass.
      bm.pmf_compact = env_int("SGMCMC_PMF_COMPACT", 0) &&
                       sg::pmf_compact_ok(cfg->dims[0], cfg->dims[1], cfg->dims[2], bm.batch, bm.full_batch, limit) ? 1 : 0;
      bm.pmf_magic = (unsigned)((0x100000000ULL + (unsigned)cfg->dims[2] - 1) / (unsigned)cfg->dims[2]);
    }
    *out = h;
    return SGMCMC_OK;
  }
  for (int l = 0; l < cfg->n_leaves; ++l)
    if (cfg->leaf_size[l] != cfg->data_dim)
      return fail(SGMCMC_E_INVALID, "vector families need every leaf to have data_dim elements");
  if (cfg->family == SGMCMC_FAMILY_LOGREG && cfg->n_leaves != 1) return fail(SGMCMC_E_INVALID, "logreg has one leaf");

  sgmcmc_handle* h = new sgmcmc_handle();
  h->cfg = *cfg;
  h->DS = (cfg->data_dim + 3) & ~3;
  // logreg with a fixed centre (estimator CV): one more float4 per row, whose .x holds center . x~ (csrc/vec_sampler.cuh,
  // CV == 2).  A 100-float row already spans 13 sectors of 32 bytes at most offsets, so the wider row costs no traffic.
  const int zc_f4 = h->DS / 4;
  const bool zc = cfg->family == SGMCMC_FAMILY_LOGREG && cfg->estimator == SGMCMC_EST_CV && zc_f4 < 128;
  if (zc) h->DS += 4;
  h->P = cfg->n_leaves * cfg->data_dim;
  const int f4 = h->DS / 4;
  h->nch = f4 <= 32 ? 1 : f4 <= 64 ? 2 : f4 <= 128 ? 4 : 8;
  if (h->nch > 4) { delete h; return fail(SGMCMC_E_UNSUPPORTED, "data_dim > 512 is not supported by the vector-family sampler"); }
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) h->num_sms = sms;
  }
  sg::VecModel& vm = h->vm;
  std::memset(&vm, 0, sizeof(vm));
  vm.family = cfg->family; vm.estimator = cfg->estimator; vm.diffusion = cfg->diffusion; vm.flags = cfg->flags;
  vm.n_leaves = cfg->n_leaves; vm.leaf_dim = cfg->data_dim; vm.P = h->P; vm.DS = h->DS;
  vm.n_data = (unsigned)cfg->n_data; vm.batch = (unsigned)cfg->batch;
  vm.own_lo = 0u; vm.own_n = vm.n_data;
  { const uint32_t mm = 65536u % vm.n_data; vm.randint_mult = (mm * mm) % vm.n_data; }   // uint32 wrap-around as in jax
  vm.zc_f4 = zc ? zc_f4 : -1;
  vm.full_batch = (cfg->estimator == SGMCMC_EST_STANDARD && (int64_t)cfg->batch == cfg->n_data) ? 1 : 0;
  vm.leapfrog = cfg->leapfrog; vm.update_rate = cfg->update_rate;
  vm.scale = (float)cfg->n_data / (float)cfg->batch;
  for (int l = 0; l < SGMCMC_MAX_LEAVES; ++l) {
    vm.lik2[l] = 2.0f * cfg->lik_coef[l];
    vm.prior2[l] = cfg->family == SGMCMC_FAMILY_GAUSSIAN_MEAN ? 2.0f * cfg->prior_coef[l] : cfg->prior_coef[l];
  }
  for (int k = 0; k < 4; ++k) vm.hyper[k] = cfg->hyper[k];
  *out = h;
  return SGMCMC_OK;
}

int sgmcmc_destroy(sgmcmc_handle* h) {
  if (!h) return SGMCMC_OK;
  if (h->rows) cudaFree(h->rows);
  if (h->mu) cudaFree(h->mu);
  if (h->data0) cudaFree(h->data0);
  if (h->data1) cudaFree(h->data1);
  if (h->cv_center) cudaFree(h->cv_center);
  if (h->cv_grad) cudaFree(h->cv_grad);
  if (h->partial) cudaFree(h->partial);
  if (h->noise) cudaFree(h->noise);
  if (h->sched) cudaFree(h->sched);
  if (h->k2_xs) cudaFree(h->k2_xs);
  if (h->k2_xt) cudaFree(h->k2_xt);
  if (h->k2_scratch) cudaFree(h->k2_scratch);
  delete h;
  return SGMCMC_OK;
}

static int write_zc_column(sgmcmc_handle* h, cudaStream_t s);

int sgmcmc_set_data(sgmcmc_handle* h, const void* xv, const void* y, void* stream) {
  if (!h || !xv) return fail(SGMCMC_E_INVALID, "null argument");
  if (h->cfg.family != SGMCMC_FAMILY_GAUSSIAN_MEAN && !y) return fail(SGMCMC_E_INVALID, "this family needs a second data array");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)h->cfg.n_data;
  if (h->cfg.family == SGMCMC_FAMILY_PMF) {
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)h->num_sms * 16);
    {   // one-time range check of the ids (synchronises the stream; ingestion is not a hot path)
      unsigned* bad_dev = nullptr;
      unsigned bad = 0;
      CUDA_TRY(cudaMalloc(&bad_dev, sizeof(unsigned)));
      cudaError_t e = cudaMemsetAsync(bad_dev, 0, sizeof(unsigned), s);
      if (e == cudaSuccess) {
        sg::pmf_check_ids_kernel<<<blocks, 256, 0, s>>>((const int*)xv, n, h->cfg.dims[0], h->cfg.dims[1], bad_dev);
        e = cudaGetLastError(); ++g_launches;
      }
      if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, bad_dev, sizeof(unsigned), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      cudaFree(bad_dev);
      CUDA_TRY(e);
      if (bad)
        return fail(SGMCMC_E_INVALID, std::to_string(bad) + " pmf rating(s) have a (user, item) id outside [0, " +
                                          std::to_string(h->cfg.dims[0]) + ") x [0, " + std::to_string(h->cfg.dims[1]) +
                                          "): ids are 0-based row numbers of U and V");
    }
    if (!h->data0) CUDA_TRY(cudaMalloc(&h->data0, n * sizeof(int4)));
    sg::pack_ratings_kernel<<<blocks, 256, 0, s>>>((const int*)xv, (const float*)y, (int4*)h->data0, n);
    CUDA_TRY(cudaGetLastError()); ++g_launches;
    h->bm.data0 = h->data0;
    h->has_data = true;
    return SGMCMC_OK;
  }
  if (h->cfg.family == SGMCMC_FAMILY_MLP) {
    const size_t b0 = n * (size_t)h->cfg.dims[1] * sizeof(float);
    const size_t b1 = n * (size_t)h->cfg.dims[1 + h->cfg.dims[0]] * sizeof(float);
    if (!h->data0) CUDA_TRY(cudaMalloc(&h->data0, b0));
    if (!h->data1) CUDA_TRY(cudaMalloc(&h->data1, b1));
    CUDA_TRY(cudaMemcpyAsync(h->data0, xv, b0, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->data1, y, b1, cudaMemcpyDeviceToDevice, s));
    h->bm.data0 = h->data0; h->bm.data1 = h->data1;
    h->has_data = true;
    return SGMCMC_OK;
  }
  const float* x = (const float*)xv;
  const size_t nrows = (size_t)h->vm.own_n;                      // == n unless sgmcmc_set_row_shard was called
  if (!h->rows) CUDA_TRY(cudaMalloc(&h->rows, nrows * h->DS * sizeof(float)));
  const size_t total = nrows * h->DS;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, (size_t)h->num_sms * 16);
  sg::pack_rows_kernel<<<blocks, 256, 0, s>>>(x, (const int*)y, h->rows, nrows, h->cfg.data_dim, h->DS, h->cfg.family);
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  h->vm.rows = h->rows;
  h->vm.mu = nullptr;
  if (h->cfg.family == SGMCMC_FAMILY_GAUSSIAN_MEAN && h->vm.own_n == h->vm.n_data) {
    // centre the rows by their column means (vec_sampler.cuh VecModel::mu): a data mean far from 0 would otherwise cancel
    // in fp32 when the gradient is formed from the row SUM.  Row-sharded tables keep mu = 0 (every rank must use one shift).
    if (!h->mu) CUDA_TRY(cudaMalloc(&h->mu, (size_t)h->DS * sizeof(float)));
    double* sums = nullptr;
    CUDA_TRY(cudaMalloc(&sums, (size_t)h->cfg.data_dim * sizeof(double)));
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)h->cfg.data_dim * sizeof(double), s);
    if (e == cudaSuccess) {
      sg::column_sum_kernel<<<(int)std::min<size_t>(nrows, (size_t)h->num_sms * 8), 128, 0, s>>>(x, nrows, h->cfg.data_dim, sums);
      sg::column_mean_kernel<<<1, 128, 0, s>>>(sums, nrows, h->cfg.data_dim, h->DS, h->mu);
      sg::center_rows_kernel<<<blocks, 256, 0, s>>>(h->rows, nrows, h->cfg.data_dim, h->DS, h->mu);
      e = cudaGetLastError(); g_launches += 3;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(sums);
    CUDA_TRY(e);
    h->vm.mu = h->mu;
  }
  h->has_data = true;
  if (h->k2_xs) { cudaFree(h->k2_xs); h->k2_xs = nullptr; }      // operand copies of the previous table
  if (h->k2_xt) { cudaFree(h->k2_xt); h->k2_xt = nullptr; }
  return write_zc_column(h, s);                                  // a new table under an existing centre
}

int sgmcmc_fullbatch_grad(sgmcmc_handle* h, const float* thetas, int32_t m, float* out, void* stream) {
  if (!h || !thetas || !out || m < 1) return fail(SGMCMC_E_INVALID, "bad argument");
  if (!h->has_data) return fail(SGMCMC_E_STATE, "sgmcmc_set_data has not been called");
  cudaStream_t s = (cudaStream_t)stream;
  if (h->big) {
    sg::BigRun run{};
    run.mode = 3; run.fg_thetas = thetas; run.fg_out = out;
    return dispatch_big(h, run, m, s);
  }
  if (h->cfg.family == SGMCMC_FAMILY_GAUSSIAN_MEAN) return fullgrad_impl<SGMCMC_FAMILY_GAUSSIAN_MEAN>(h, thetas, m, out, s);
  // many parameter vectors: two tensor-core contractions instead of m passes over the table (SGMCMC_K2_TC=0/1 overrides)
  const int tc = env_int("SGMCMC_K2_TC", -1);
  if ((tc == 1 || (tc < 0 && m >= 64)) && h->cfg.data_dim <= sg::k2::DN_MAX && h->vm.own_n == h->vm.n_data)
    return fullgrad_tc(h, thetas, m, out, s);
  return fullgrad_impl<SGMCMC_FAMILY_LOGREG>(h, thetas, m, out, s);
}

static int write_zc_column(sgmcmc_handle* h, cudaStream_t s) {
  if (h->vm.zc_f4 < 0 || !h->has_center) return SGMCMC_OK;
  const int blocks = h->num_sms * 8;
  switch (h->nch) {
    case 1: sg::zc_column_kernel<1><<<blocks, 256, 0, s>>>(h->rows, h->vm.own_n, h->DS, h->P, h->vm.zc_f4, h->cv_center); break;
    case 2: sg::zc_column_kernel<2><<<blocks, 256, 0, s>>>(h->rows, h->vm.own_n, h->DS, h->P, h->vm.zc_f4, h->cv_center); break;
    default: sg::zc_column_kernel<4><<<blocks, 256, 0, s>>>(h->rows, h->vm.own_n, h->DS, h->P, h->vm.zc_f4, h->cv_center); break;
  }
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  return SGMCMC_OK;
}

int sgmcmc_set_centering(sgmcmc_handle* h, const float* theta_hat, void* stream) {
  if (!h || !theta_hat) return fail(SGMCMC_E_INVALID, "null argument");
  if (!h->has_data) return fail(SGMCMC_E_STATE, "sgmcmc_set_data has not been called");
  cudaStream_t s = (cudaStream_t)stream;
  if (!h->cv_center) CUDA_TRY(cudaMalloc(&h->cv_center, h->P * sizeof(float)));
  if (!h->cv_grad) CUDA_TRY(cudaMalloc(&h->cv_grad, h->P * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(h->cv_center, theta_hat, h->P * sizeof(float), cudaMemcpyDeviceToDevice, s));
  const int rc = sgmcmc_fullbatch_grad(h, h->cv_center, 1, h->cv_grad, stream);   // gradient_estimation.py:70
  if (rc != SGMCMC_OK) return rc;
  h->vm.cv_center = h->cv_center;
  h->vm.cv_grad = h->cv_grad;
  h->bm.cv_center = h->cv_center;
  h->bm.cv_grad = h->cv_grad;
  h->has_center = true;
  return write_zc_column(h, s);
}

int sgmcmc_init(sgmcmc_handle* h, sgmcmc_state* st, const uint32_t* keys, int32_t sampler_mode, void* stream) {
  int rc = check_state(h, st);
  if (rc != SGMCMC_OK) return rc;
  if (!keys) return fail(SGMCMC_E_INVALID, "keys must not be null");
  if (sampler_mode && !st->key) return fail(SGMCMC_E_INVALID, "state.key is required in sampler mode");
  if (h->big) {
    sg::BigRun run{};
    run.st = *st; run.mode = sampler_mode ? 2 : 1; run.init_keys = keys;
    return dispatch_big(h, run, st->n_chains, (cudaStream_t)stream);
  }
  sg::VecRun run{};
  run.st = *st; run.mode = sampler_mode ? 2 : 1; run.init_keys = keys;
  return dispatch_vec(h, run, (cudaStream_t)stream);
}

int sgmcmc_run(sgmcmc_handle* h, sgmcmc_state* st, int32_t i0, int32_t k, const sgmcmc_step_coef* coef,
               int32_t coef_stride, float* samples, int64_t sample_chain_stride, void* stream) {
  return sgmcmc_run_thinned(h, st, i0, k, coef, coef_stride, 1, samples, sample_chain_stride, nullptr, stream);
}

int sgmcmc_run_thinned(sgmcmc_handle* h, sgmcmc_state* st, int32_t i0, int32_t k, const sgmcmc_step_coef* coef,
                       int32_t coef_stride, int32_t thin, float* samples, int64_t sample_chain_stride, float* moments,
                       void* stream) {
  int rc = check_state(h, st);
  if (rc != SGMCMC_OK) return rc;
  if (k < 0 || i0 < 0) return fail(SGMCMC_E_INVALID, "i0 and k must be non-negative");
  if (thin < 1) return fail(SGMCMC_E_INVALID, "thin must be >= 1");
  if (!coef || (coef_stride != 0 && coef_stride != 1)) return fail(SGMCMC_E_INVALID, "bad step coefficients");
  if (!st->key) return fail(SGMCMC_E_INVALID, "state.key is required");
  if (h->cfg.diffusion == SGMCMC_DIFF_ADAM_OPT) return fail(SGMCMC_E_STATE, "this handle is an optimiser: use sgmcmc_optimize");
  if (k == 0) return SGMCMC_OK;
  if (h->big) {
    sg::BigRun run{};
    run.st = *st; run.i0 = i0; run.K = k; run.coef = coef; run.coef_stride = coef_stride; run.samples = samples;
    run.sample_chain_stride = sample_chain_stride > 0 ? (size_t)sample_chain_stride : (size_t)(k / thin) * h->P;
    run.thin = thin; run.moments = moments;
    if (SG_MLP_PRODUCER_WARPS > 0 && h->cfg.family == SGMCMC_FAMILY_MLP && k > 1 && !env_int("SGMCMC_NO_LOOKAHEAD", 0)) {
      // look-ahead noise scratch of the optional producer warps (stream-ordered reuse: one stream per handle)
      const size_t need = (size_t)st->n_chains * h->P;
      if (need > h->noise_elems) {
        if (h->noise) CUDA_TRY(cudaFree(h->noise));
        h->noise = nullptr; h->noise_elems = 0;
        CUDA_TRY(cudaMalloc(&h->noise, need * sizeof(float)));
        h->noise_elems = need;
      }
      run.noise = h->noise;
    }
    return dispatch_big(h, run, st->n_chains, (cudaStream_t)stream);
  }
  sg::VecRun run{};
  run.st = *st; run.i0 = i0; run.K = k; run.coef = coef; run.coef_stride = coef_stride; run.samples = samples;
  run.sample_chain_stride = sample_chain_stride > 0 ? (size_t)sample_chain_stride : (size_t)(k / thin) * h->P;
  run.thin = thin; run.moments = moments;
  return dispatch_vec(h, run, (cudaStream_t)stream);
}

int sgmcmc_optimize(sgmcmc_handle* h, sgmcmc_state* st, int32_t i0, int32_t k, const sgmcmc_step_coef* coef, float* lp_out,
                    void* stream) {
  int rc = check_state(h, st);
  if (rc != SGMCMC_OK) return rc;
  if (h->cfg.diffusion != SGMCMC_DIFF_ADAM_OPT) return fail(SGMCMC_E_STATE, "the handle was not created with SGMCMC_DIFF_ADAM_OPT");
  if (k < 0 || i0 < 0 || !coef || !st->key) return fail(SGMCMC_E_INVALID, "bad argument");
  if (k == 0) return SGMCMC_OK;
  if (h->big) {
    sg::BigRun run{};
    run.st = *st; run.i0 = i0; run.K = k; run.coef = coef; run.coef_stride = 1; run.lp_out = lp_out; run.thin = 1;
    return dispatch_big(h, run, st->n_chains, (cudaStream_t)stream);
  }
  sg::VecRun run{};
  run.st = *st; run.i0 = i0; run.K = k; run.coef = coef; run.coef_stride = 1; run.lp_out = lp_out;
  return dispatch_vec(h, run, (cudaStream_t)stream);
}

int sgmcmc_kernel_step(sgmcmc_handle* h, sgmcmc_state* st, int32_t i, const uint32_t* keys,
                       const sgmcmc_step_coef* coef, void* stream) {
  int rc = check_state(h, st);
  if (rc != SGMCMC_OK) return rc;
  if (!keys || !coef || i < 0) return fail(SGMCMC_E_INVALID, "bad argument");
  if (h->big) {
    sg::BigRun run{};
    run.st = *st; run.i0 = i; run.K = 1; run.coef = coef; run.coef_stride = 0; run.step_keys = keys;
    return dispatch_big(h, run, st->n_chains, (cudaStream_t)stream);
  }
  sg::VecRun run{};
  run.st = *st; run.i0 = i; run.K = 1; run.coef = coef; run.coef_stride = 0; run.step_keys = keys;
  return dispatch_vec(h, run, (cudaStream_t)stream);
}

int sgmcmc_set_row_shard(sgmcmc_handle* h, int64_t row0, int64_t n_rows) {
  if (!h) return fail(SGMCMC_E_INVALID, "null handle");
  if (h->big) return fail(SGMCMC_E_UNSUPPORTED, "the data-sharded mode covers the vector families (gaussian_mean, logreg)");
  if (h->has_data) return fail(SGMCMC_E_STATE, "sgmcmc_set_row_shard must precede sgmcmc_set_data");
  if (row0 < 0 || n_rows < 1 || row0 + n_rows > h->cfg.n_data) return fail(SGMCMC_E_INVALID, "row shard out of range");
  if (h->cfg.estimator == SGMCMC_EST_SVRG || h->cfg.diffusion == SGMCMC_DIFF_SGHMC || h->vm.full_batch)
    return fail(SGMCMC_E_UNSUPPORTED, "data-sharded mode: SVRG, SGHMC and the full-batch bypass are not supported");
  h->vm.own_lo = (unsigned)row0; h->vm.own_n = (unsigned)n_rows;
  return SGMCMC_OK;
}

int sgmcmc_set_centering_grad(sgmcmc_handle* h, const float* theta_hat, const float* grad, void* stream) {
  if (!h || !theta_hat || !grad) return fail(SGMCMC_E_INVALID, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (!h->cv_center) CUDA_TRY(cudaMalloc(&h->cv_center, h->P * sizeof(float)));
  if (!h->cv_grad) CUDA_TRY(cudaMalloc(&h->cv_grad, h->P * sizeof(float)));
  CUDA_TRY(cudaMemcpyAsync(h->cv_center, theta_hat, h->P * sizeof(float), cudaMemcpyDeviceToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(h->cv_grad, grad, h->P * sizeof(float), cudaMemcpyDeviceToDevice, s));
  h->vm.cv_center = h->cv_center; h->vm.cv_grad = h->cv_grad;
  h->bm.cv_center = h->cv_center; h->bm.cv_grad = h->cv_grad;
  h->has_center = true;
  if (h->has_data) return write_zc_column(h, s);
  return SGMCMC_OK;
}

int sgmcmc_sum_width(const sgmcmc_handle* h) { return h ? h->DS : 0; }

int sgmcmc_launch_shape(const sgmcmc_handle* h, int32_t n_chains, int32_t* warps_per_cta, int32_t* cluster) {
  if (!h || !warps_per_cta || !cluster || n_chains < 1) return fail(SGMCMC_E_INVALID, "bad argument");
  if (h->big) return fail(SGMCMC_E_UNSUPPORTED, "launch shapes are a property of the vector-family sampler");
  const VecShape s = choose_shape(h, n_chains);
  *warps_per_cta = s.warps; *cluster = s.cluster;
  return SGMCMC_OK;
}

int sgmcmc_shard_init(sgmcmc_handle* h, sgmcmc_state* st, const uint32_t* keys, int32_t sampler_mode, float* sum_out,
                      void* stream) {
  int rc = check_state(h, st);
  if (rc != SGMCMC_OK) return rc;
  if (h->big) return fail(SGMCMC_E_UNSUPPORTED, "data-sharded mode covers the vector families");
  if (!keys || !sum_out) return fail(SGMCMC_E_INVALID, "keys and sum_out must not be null");
  if (sampler_mode && !st->key) return fail(SGMCMC_E_INVALID, "state.key is required in sampler mode");
  sg::VecRun run{};
  run.st = *st; run.mode = sampler_mode ? 2 : 1; run.init_keys = keys; run.sum_out = sum_out;
  return dispatch_vec(h, run, (cudaStream_t)stream);
}

int sgmcmc_shard_step(sgmcmc_handle* h, sgmcmc_state* st, int32_t i, const sgmcmc_step_coef* coef,
                      const sgmcmc_step_coef* coef_prev, int32_t flags, const float* sum_in, float* sum_out, float* sample_out,
                      void* stream) {
  int rc = check_state(h, st);
  if (rc != SGMCMC_OK) return rc;
  if (h->big) return fail(SGMCMC_E_UNSUPPORTED, "data-sharded mode covers the vector families");
  if (!sum_in || i < 0 || !st->key) return fail(SGMCMC_E_INVALID, "bad argument");
  if ((flags & sg::SHARD_STEP) && (!coef || !sum_out)) return fail(SGMCMC_E_INVALID, "a step needs coef and sum_out");
  if ((flags & sg::SHARD_UPDATE2) && !coef_prev) return fail(SGMCMC_E_INVALID, "update2 needs coef_prev");
  sg::VecRun run{};
  run.st = *st; run.i0 = i; run.K = 1; run.coef = coef; run.coef_stride = 0; run.coef_prev = coef_prev;
  run.shard_flags = flags; run.sum_in = sum_in; run.sum_out = (flags & sg::SHARD_STEP) ? sum_out : nullptr;
  run.samples = sample_out; run.sample_chain_stride = (size_t)h->P;
  return dispatch_vec(h, run, (cudaStream_t)stream);
}


int sgmcmc_minibatch_indices(const uint32_t* key, int64_t n, int32_t batch, int32_t* out, void* stream) {
  if (!key || !out || n <= 0 || n > 0x7fffffffLL || batch <= 0) return fail(SGMCMC_E_INVALID, "bad argument");
  sg::Key k{key[0], key[1]};
  const unsigned h = ((unsigned)batch + 1u) >> 1;
  sg::indices_kernel<<<(h + 255) / 256, 256, 0, (cudaStream_t)stream>>>(k, (unsigned)n, (unsigned)batch, out);
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  return SGMCMC_OK;
}

int sgmcmc_normal(const uint32_t* key, int32_t n, float* out, void* stream) {
  if (!key || !out || n <= 0) return fail(SGMCMC_E_INVALID, "bad argument");
  sg::Key k{key[0], key[1]};
  const int blocks = std::min((n + 255) / 256, 148 * 8);
  sg::normal_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(k, (unsigned)n, out);
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  return SGMCMC_OK;
}

int sgmcmc_uniform(const uint32_t* key, int32_t n, float* out, void* stream) {
  if (!key || !out || n <= 0) return fail(SGMCMC_E_INVALID, "bad argument");
  sg::Key k{key[0], key[1]};
  const int blocks = std::min((n + 255) / 256, 148 * 8);
  sg::uniform_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(k, (unsigned)n, out);
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  return SGMCMC_OK;
}

int sgmcmc_bernoulli_rows(const uint32_t* key, int32_t n, const float* p, int32_t* out, void* stream) {
  if (!key || !p || !out || n <= 0) return fail(SGMCMC_E_INVALID, "bad argument");
  sg::Key k{key[0], key[1]};
  const int blocks = std::min((n + 255) / 256, 148 * 8);
  sg::bernoulli_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(k, (unsigned)n, p, out);
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  return SGMCMC_OK;
}

int sgmcmc_split(const uint32_t* key, int32_t num, uint32_t* out, void* stream) {
  if (!key || !out || num <= 0) return fail(SGMCMC_E_INVALID, "bad argument");
  sg::Key k{key[0], key[1]};
  const int blocks = std::min((2 * num + 255) / 256, 148 * 8);
  sg::split_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(k, (unsigned)num, out);
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  return SGMCMC_OK;
}

int sgmcmc_ksd_imq_partial(const float* x, const float* g, int32_t n, int32_t d, int32_t row0, int32_t row1,
                           double* partial_sum, void* stream) {
  if (!x || !g || !partial_sum || n <= 0 || d <= 0 || row0 < 0 || row1 > n || row0 > row1)
    return fail(SGMCMC_E_INVALID, "bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  CUDA_TRY(cudaMemsetAsync(partial_sum, 0, sizeof(double), s));
  if (row1 == row0) return SGMCMC_OK;
  dim3 grid((n + sg::KSD_T - 1) / sg::KSD_T, (row1 - row0 + sg::KSD_T - 1) / sg::KSD_T);
  sg::ksd_simt_kernel<<<grid, 256, 0, s>>>(x, g, n, d, row0, row1, partial_sum);
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  return SGMCMC_OK;
}

int64_t sgmcmc_ksd_workspace_bytes(int32_t n, int32_t d) {
  if (n <= 0 || d <= 0) return 0;
  return (int64_t)sg::ksdtc::workspace_bytes(n, d);
}

int sgmcmc_ksd_imq_tc(const float* x, const float* g, int32_t n, int32_t d, int32_t part, int32_t nparts,
                      void* workspace, double* partial_sum, void* stream) {
  namespace kt = sg::ksdtc;
  if (!x || !g || !workspace || !partial_sum || n <= 0 || d <= 0 || nparts <= 0 || part < 0 || part >= nparts)
    return fail(SGMCMC_E_INVALID, "bad argument");
  if (((uintptr_t)workspace & 255) != 0) return fail(SGMCMC_E_INVALID, "workspace must be 256-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int dp = kt::padded_dim(d);
  uint8_t* w = (uint8_t*)workspace;
  float* Z = (float*)w;                       w += kt::align_up(kt::z_floats(n, d) * sizeof(float), 256);
  float4* rt = (float4*)w;                    w += kt::align_up((size_t)n * sizeof(float4), 256);
  double* sums = (double*)w;
  CUDA_TRY(cudaMemsetAsync(partial_sum, 0, sizeof(double), s));
  CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)2 * d * sizeof(double), s));
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (d % 4 == 0 && d / 4 <= 256 && (((uintptr_t)x | (uintptr_t)g) & 15) == 0) {
    const int d4 = d / 4, rpi = std::max(1, 256 / d4), threads = 256;
    const int blocks = std::max(1, std::min((n + rpi - 1) / rpi, sms * 4));
    kt::colsum4_kernel<<<blocks, threads, (size_t)rpi * d4 * 8 * sizeof(double), s>>>((const float4*)x, (const float4*)g, n, d4, sums);
  } else {
    kt::colsum_kernel<<<std::min(n, sms * 8), 128, 0, s>>>(x, g, n, d, sums);
  }
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  kt::split_kernel<<<(n + 7) / 8, 256, 0, s>>>(x, g, n, d, dp, sums, Z, rt);
  CUDA_TRY(cudaGetLastError()); ++g_launches;

  kt::EncodeTiledFn encode = kt::encode_tiled_fn();
  if (!encode) return fail(SGMCMC_E_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  CUtensorMap zmap;
  const cuuint64_t gdim[2] = {(cuuint64_t)kt::KC, (cuuint64_t)(kt::z_floats(n, d) / kt::KC)};
  const cuuint64_t gstride[1] = {(cuuint64_t)kt::KC * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kt::KC, (cuuint32_t)kt::TILE};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult cr = encode(&zmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, Z, gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(SGMCMC_E_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)cr));

  const long long T = (n + kt::TILE - 1) / kt::TILE;
  // SGMCMC_KSD_CLUSTER=2: clusters of two CTAs sharing the row-side operand by TMA multicast (ksd_tc.cuh).  Parity-green and
  // 25 % less L2 -> SM operand traffic, but measured 1.4 % SLOWER (4.91 vs 4.84 ms at N = 50k): the kernel is bound by the
  // MMA -> epilogue hand-over of its single accumulator set, not by the operand stream; one CTA per tile stays the default.
  const char* cl_env = getenv("SGMCMC_KSD_CLUSTER");
  const int cl = cl_env && atoi(cl_env) == 2 ? 2 : 1;
  const long long total = cl == 1 ? T * (T + 1) / 2 : kt::pair_start(T, T);
  const long long n_local = total > part ? (total - part + nparts - 1) / nparts : 0;
  if (n_local == 0) return SGMCMC_OK;
  if (cl == 1) {
    CUDA_TRY(cudaFuncSetAttribute(kt::ksd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kt::SMEM_BYTES));
    const int grid = (int)std::min<long long>(n_local, sms);
    kt::ksd_tc_kernel<1><<<grid, kt::NUM_THREADS, kt::SMEM_BYTES, s>>>(zmap, rt, n, d, dp, part, nparts, partial_sum);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(kt::ksd_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kt::SMEM_BYTES));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * std::min<long long>(n_local, sms / 2)));
    cfg.blockDim = dim3(kt::NUM_THREADS);
    cfg.dynamicSmemBytes = kt::SMEM_BYTES;
    cfg.stream = s;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kt::ksd_tc_kernel<2>, zmap, (const float4*)rt, n, d, dp, (int)part, (int)nparts, partial_sum));
  }
  CUDA_TRY(cudaGetLastError()); ++g_launches;
  return SGMCMC_OK;
}

}  // extern "C"
