// Bayesian MLP family (sgmcmcjax/models/bayesian_NN/NN_model.py:31-58) for the big-state sampler.
//
//   predict(params, x): a_0 = x;  a_l = softmax(W_l a_{l-1} + b_l) for l < L (softmax, NOT relu, :36-37);
//                       log_softmax(W_L a_{L-1} + b_L)                                      (:39-41)
//   loglikelihood = sum_k y_k predict_k (:48-50);  logprior = sum N(0,1).logpdf (:53-58) -> grad = -w
//
// lik_pass ADDS  sign * sum_i grad_theta loglik(theta, x_i, y_i)  into the gradient buffer, in sub-batches
// of <= 128 rows.  The two layer-1 contractions carry ~99 % of the flops and run on the tensor cores
// (tcgen05, fp32 accumulators in TMEM) as 3xTF32:  a = a_hi + a_lo,  a_hi = rna_tf32(a),
// a_lo = rna_tf32(a - a_hi);  a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi  (error ~2^-21 |a||b|).
//
//   forward   Z1[b][n]  = sum_k X[b][k] W1[n][k]          M = b (128), N = n (s1 -> 16), K = k
//             A = gathered rows of X, B = W1, both K-major, 32-float K chunks in 128-byte-swizzled rows;
//             2-stage smem pipeline: all threads load (registers) -> split hi/lo -> st.shared, one thread
//             issues 12 MMAs per chunk, tcgen05.commit frees the stage.
//   backward  dW1^T[k][n] = sum_b X[b][k] dZ1[b][n]       M = k (128-wide tiles), N = n, K = b
//             A = the same gathered rows, B = dZ1: both MN-major (the row-major [b][k] / [b][n] data as it
//             lies, in 4-row x 128-byte SWIZZLE_128B_BASE32B atoms); 4 M-tiles (<= 512 TMEM columns) per pass; dZ1 hi/lo
//             is staged once, X streams through a 2-stage pipeline; the epilogue adds the TMEM tiles into
//             the chain's gradient (coalesced: lane = k).
// The small layers (softmax, layers >= 2, their backward) stay fp32 SIMT.  Shapes: n_layers <= 3,
// hidden / output widths <= 128, any input width; the shared-memory need is checked at create time.
#pragma once
#include "big_sampler.cuh"
#include "ksd_tc.cuh"   // PTX wrappers: mbarrier, tcgen05 alloc / mma / commit / ld, tf32 rounding

namespace sg {

constexpr int MLP_ROWS = 128;      // largest sub-batch (MMA M of the forward contraction)
#ifndef SG_MLP_W_LDCG
#define SG_MLP_W_LDCG 0
#endif
#ifndef SG_MLP_FWD_DEEP_ITEMS
#define SG_MLP_FWD_DEEP_ITEMS 0    // forward GEMM: how many of a stager's 4-5 items are prefetched TWO chunks ahead (the rest: one); 1 and 2 spill
#endif
constexpr int MLP_THREADS = 512;   // gradient worker threads (warps 0..15); the CTA has 8 more noise-producer warps
constexpr int MLP_STAGER_ROWS = (MLP_THREADS - 32) / 8;   // forward GEMM: rows one pass of the 480 operand-staging threads covers
constexpr int MLP_KC = 32;         // forward K chunk (floats): one 128-byte swizzled row
constexpr int MLP_BWD_TILE = 120;  // backward GEMM: input columns per M tile (of 128 lanes): 2 tiles x 8 rows = 480 16-byte slots
// SG_MLP_ISSUER_FENCE = 1: the generic -> async proxy fence that makes a filled stage visible to tcgen05.mma is executed ONCE,
// by the issuing thread, after the hand-over barrier (whose release / acquire makes the stagers' st.shared visible to it;
// the proxy fence is cumulative over what the thread has observed).  The stagers then never execute it: as SASS it is
// MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR waits for every load / cp.async the thread still has in flight -
// i.e. it serialised the operand prefetch with the stage fill.  0: every stager fences its own writes (round 1 .. 2b).
#ifndef SG_MLP_ISSUER_FENCE
#define SG_MLP_ISSUER_FENCE 1
#endif
constexpr int MLP_RING_DEPTH = 2;   // forward GEMM: chunks of operand loads in flight in the cp.async ring
constexpr int MLP_W2S = 128;       // row stride of the staged, zero-padded copy of W_2 (>= any layer-1 width)
constexpr size_t MLP_SMEM_LIMIT = 232448;   // 227 KB opt-in maximum per CTA

__host__ __device__ inline int mlp_stride(int w) { return w | 1; }   // odd row stride: conflict-free row-per-lane access
__host__ __device__ inline int mlp_n16(int w) { return (w + 15) & ~15; }

struct MlpSmem {
  uint64_t bar[4];             // stage-free barriers (tcgen05.commit arrives): forward ring 2 stages, backward ring 2 or 4
  uint32_t parity[4];          // next phase parity of bar[], carried from one GEMM to the next
  uint32_t tmem_base;
  uint32_t pad;
  int idx[MLP_ROWS];
};

// Shared-memory plan: [region | xs | fp32 activations] after a 1024-byte aligned base.
struct MlpPlan {
  int L, s1n;                  // s1n = layer-1 width rounded up to 16 (MMA N)
  int rows;                    // sub-batch capacity: min(128, batch rounded up to 8)
  int tmem_cols;
  size_t region_bytes;         // forward stages (2 x [Xhi | Xlo | Whi | Wlo]) / backward dZ1 hi + lo
  size_t xs_bytes;             // backward X stages: 2 x [hi | lo] of 8 rows x 256 columns (also: bias-gradient scratch)
  int act_off1, act_off2, act_off3;   // float offsets of act[1..L]
  int dz_off0, dz_off1;        // float offsets of the two ping-pong dZ buffers (layers >= 2; layer 1 if L == 1)
  int dz_w0, dz_w1;
  size_t total_bytes;
  __host__ __device__ int act_off(int l) const { return l == 1 ? act_off1 : l == 2 ? act_off2 : act_off3; }
  __host__ __device__ int dz_off(int k) const { return k ? dz_off1 : dz_off0; }
  __host__ __device__ int dz_w(int k) const { return k ? dz_w1 : dz_w0; }
};

__host__ __device__ inline MlpPlan mlp_plan_rows(const int* dims, int rows) {
  MlpPlan p{};
  p.L = dims[0];
  const int* sz = dims + 1;
  p.s1n = mlp_n16(sz[1]);
  p.rows = rows;
  int cols = 32;
  while (cols < 4 * p.s1n && cols < 512) cols <<= 1;
  p.tmem_cols = cols;
  const size_t fwd_stage = 2 * (size_t)MLP_ROWS * 128 + 2 * (size_t)p.s1n * 128;
  const size_t bwd_b = 2 * (size_t)(p.rows / 8) * 4096;
  p.region_bytes = 2 * fwd_stage > bwd_b ? 2 * fwd_stage : bwd_b;
  p.xs_bytes = 2 * 2 * 8 * 256 * sizeof(float);
  int f = 0;
  p.act_off1 = f; f += p.rows * mlp_stride(sz[1]);
  p.act_off2 = f; if (p.L >= 2) f += p.rows * mlp_stride(sz[2]);
  p.act_off3 = f; if (p.L >= 3) f += p.rows * mlp_stride(sz[3]);
  // dZ_l lives in buffer (L - l) & 1; layer 1 only needs one when it is the output layer (L == 1)
  int w0 = 1, w1 = 1;
  for (int l = (p.L == 1 ? 1 : 2); l <= p.L; ++l) {
    if ((p.L - l) & 1) { if (sz[l] > w1) w1 = sz[l]; } else { if (sz[l] > w0) w0 = sz[l]; }
  }
  p.dz_w0 = w0; p.dz_off0 = f; f += p.rows * mlp_stride(w0);
  p.dz_w1 = w1; p.dz_off1 = f; f += p.rows * mlp_stride(w1);
  p.total_bytes = 1024 + p.region_bytes + p.xs_bytes + (size_t)f * sizeof(float);
  return p;
}

// sub-batch capacity: min(128, batch rounded up to 8), lowered in steps of 8 until the plan fits in shared memory
__host__ __device__ inline MlpPlan mlp_plan(const int* dims, unsigned batch) {
  int rows = batch >= (unsigned)MLP_ROWS ? MLP_ROWS : (int)((batch + 7u) & ~7u);
  MlpPlan p = mlp_plan_rows(dims, rows);
  while (rows > 8 && p.total_bytes + 4096 > MLP_SMEM_LIMIT) { rows -= 8; p = mlp_plan_rows(dims, rows); }
  return p;
}

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
  if (mlp_plan(cfg->dims, (unsigned)cfg->batch).total_bytes + 4096 > MLP_SMEM_LIMIT)
    return "mlp: these layer widths need more than 227 KB of shared memory per chain";
  return nullptr;
}

namespace mlptc {
using namespace ksdtc;

// barrier of the 512 gradient worker threads (the noise-producer warps are elsewhere during a likelihood pass)
__device__ __forceinline__ void worker_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
// Stage hand-over to the MMA-issuing thread (thread 0): the 15 other worker warps only ARRIVE on the stage's named barrier
// (2 + stage) and go on to their next loads; warp 0 syncs on it and issues.  With a full barrier here every chunk cost all
// 512 threads the ~1.4 us thread 0 needs to issue its 12 tcgen05.mma (phase timers: 34 of the 95 us of the forward GEMM).
// One barrier per ring stage: a worker cannot arrive for chunk c + 2 before the MMAs of chunk c have retired, i.e. before
// warp 0 has passed this barrier for chunk c, so instances never overlap.
__device__ __forceinline__ void stage_handover(int stage, bool issuer_warp) {
  if (issuer_warp) asm volatile("bar.sync %0, 512;" ::"r"(2 + stage) : "memory");
  else asm volatile("bar.arrive %0, 512;" ::"r"(2 + stage) : "memory");
}

// Per-thread asynchronous copies global -> shared (LDGSTS); src_bytes < the copy size zero-fills the rest.
__device__ __forceinline__ void cp_async16(uint32_t sdst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t sdst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sdst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

// K-major, 128-byte swizzle: rows of 32 floats, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
// MN-major tf32 operands have ONE legal smem layout (cutlass sm100_smem_selector: "for mn-major tf32 operands,
// SW128_32B is the only available smem layout"): SWIZZLE_128B_BASE32B, atoms of 4 K-rows x 32 MN-elements (512 B),
// the 32-byte chunk index of a row XORed with the K-row (Swizzle<2,5,2>); consecutive MN atoms lbo bytes apart,
// consecutive 4-row K groups sbo bytes apart (one K = 8 MMA reads two groups).
__device__ __forceinline__ uint64_t desc_mn_sw128_32b(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                    // SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of the 16-byte chunk c16 (0..7) of K-row r (0..7) inside an 8-row x 128-byte block of that layout
__device__ __forceinline__ uint32_t mn_chunk_off(int r, int c16) {
  return (uint32_t)(r >> 2) * 512u + (uint32_t)(r & 3) * 128u + (uint32_t)(((c16 >> 1) ^ (r & 3)) << 5) + (uint32_t)(c16 & 1) * 16u;
}
// cute::UMMA::InstrDescriptor: fp32 accumulate, tf32 x tf32, M = 128, N = n; bits 15 / 16 select MN-major A / B
__host__ __device__ constexpr uint32_t idesc_tf32(int n, bool mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
  hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
  lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y); lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
}
// 4 consecutive floats of a row that is only guaranteed 4-byte aligned; elements at or beyond `valid` read as 0
__device__ __forceinline__ float4 load4(const float* src, int valid, bool a16, bool a8, bool read_only) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid >= 4 && a16) {
    v = read_only ? __ldg(reinterpret_cast<const float4*>(src)) : *reinterpret_cast<const float4*>(src);
  } else if (valid >= 4 && a8) {
    const float2 p = *reinterpret_cast<const float2*>(src), q = *(reinterpret_cast<const float2*>(src) + 1);
    v = make_float4(p.x, p.y, q.x, q.y);
  } else {
    if (valid > 0) v.x = src[0];
    if (valid > 1) v.y = src[1];
    if (valid > 2) v.z = src[2];
    if (valid > 3) v.w = src[3];
  }
  return v;
}
}  // namespace mlptc

template <>
struct FamilyGrad<SGMCMC_FAMILY_MLP> {
  static constexpr bool kNoiseProducers = SG_MLP_PRODUCER_WARPS > 0;
  static size_t smem_bytes(const BigModel& m) { return sizeof(MlpSmem) + 16 + mlp_plan(m.dims, m.batch).total_bytes; }

  __device__ static void cta_init(const BigModel& m, void* fam_smem) {
    MlpSmem& ms = *reinterpret_cast<MlpSmem*>(fam_smem);
    if (threadIdx.x == 0) {
      for (int i = 0; i < 4; ++i) { mlptc::mbar_init(&ms.bar[i], 1); ms.parity[i] = 0; }
      mlptc::fence_barrier_init();
    }
    if (threadIdx.x < 32) { mlptc::tmem_alloc(&ms.tmem_base, (uint32_t)mlp_plan(m.dims, m.batch).tmem_cols); mlptc::tmem_relinquish(); }
    mlptc::tc_fence_before();
    __syncthreads();
    mlptc::tc_fence_after();
  }
  __device__ static void cta_fini(const BigModel& m, void* fam_smem) {
    MlpSmem& ms = *reinterpret_cast<MlpSmem*>(fam_smem);
    mlptc::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) mlptc::tmem_dealloc(ms.tmem_base, (uint32_t)mlp_plan(m.dims, m.batch).tmem_cols);
  }

  // ---------------------------------------------------------------------------------------------------
  // Z1 = X W1^T + b1 into act1 (fp32, [rows][zs]); rows >= nb come out as b1.
  __device__ __forceinline__ static void layer1_forward(const BigModel& m, MlpSmem& ms, uint8_t* region, const float* th, int nb,
                                        int rows, float* Z, int zs, uint8_t* raw, size_t raw_bytes) {
    const int s0 = m.dims[1];
    const float* W1 = th + m.leaf_off[0];
    // fast path (s0 % 4 == 0, W1 at least 8-byte aligned) as its own instantiation: merged with the general path, ptxas
    // if-converts the two and leaves predicated-off spill stores that still wait for the staged loads
    const bool fast = (s0 & 3) == 0 && (reinterpret_cast<uintptr_t>(W1) & 7) == 0;
    const bool ni4 = mlp_n16(m.dims[2]) <= 4 * MLP_STAGER_ROWS - MLP_ROWS;
    // ... and with room for the raw operand ring (`raw`: shared memory that is idle during this GEMM), the prefetch goes
    // through per-thread cp.async slots there instead of registers (layer1_forward_impl)
    const bool ring = fast && raw_bytes >= (size_t)MLP_RING_DEPTH * (ni4 ? 4 : 5) * (MLP_THREADS - 32) * 16;
    if (ni4) {
      if (ring) layer1_forward_impl<true, 4, true>(m, ms, region, th, nb, rows, Z, zs, raw);
      else if (fast) layer1_forward_impl<true, 4, false>(m, ms, region, th, nb, rows, Z, zs, raw);
      else layer1_forward_impl<false, 4, false>(m, ms, region, th, nb, rows, Z, zs, raw);
    } else {
      if (ring) layer1_forward_impl<true, 5, true>(m, ms, region, th, nb, rows, Z, zs, raw);
      else if (fast) layer1_forward_impl<true, 5, false>(m, ms, region, th, nb, rows, Z, zs, raw);
      else layer1_forward_impl<false, 5, false>(m, ms, region, th, nb, rows, Z, zs, raw);
    }
  }
  // Warp 0 only issues the MMAs; warps 1..15 (480 "stagers") move the operands.  Stager t handles 16-byte chunk t & 7 of
  // rows (t >> 3) + 60 j, j < NI, of the stack [X rows 0..127 | W1 rows 0..s1n): NI = 4 covers s1n <= 112 (the 784-100-10
  // network: 240 rows x 8 chunks = 480 x 4 exactly), NI = 5 the widths up to 128.
  // RING: the operand prefetch does not live in registers (one chunk of them was all that fitted, and its load latency
  // showed in every chunk) but in per-thread slots of a shared-memory ring filled by cp.async: MLP_RING_DEPTH chunks in
  // flight, read back by the thread that requested them (no barrier needed), then split into the tensor-core stages.
  template <bool FAST, int NI, bool RING>
  __device__ __noinline__ static void layer1_forward_impl(const BigModel& m, MlpSmem& ms, uint8_t* region, const float* th, int nb,
                                        int rows, float* Z, int zs, uint8_t* raw) {
    using namespace mlptc;
    const int s0 = m.dims[1], s1 = m.dims[2];
    const int s1n = mlp_n16(s1);
    const int tid = threadIdx.x;
    const float* X = reinterpret_cast<const float*>(m.data0);
    const float* W1 = th + m.leaf_off[0];
    const float* b1 = th + m.leaf_off[1];
    const bool x16 = (s0 & 3) == 0;                                        // cudaMalloc'd table: rows 16-B aligned
    const bool w16 = x16 && (reinterpret_cast<uintptr_t>(W1) & 15) == 0;   // [C, P] state rows: 4-B aligned in general
    const bool w8 = (s0 & 1) == 0 && (reinterpret_cast<uintptr_t>(W1) & 7) == 0;
    const size_t stage_bytes = 2 * (size_t)MLP_ROWS * 128 + 2 * (size_t)s1n * 128;
    const int n_chunks = (s0 + MLP_KC - 1) / MLP_KC;
    const int s0p8 = (s0 + 7) & ~7;
    const uint32_t idesc = idesc_tf32(s1n, false);
    const uint32_t tmem = ms.tmem_base;
    uint32_t par[2] = {ms.parity[0], ms.parity[1]};
    int uses[2] = {0, 0};

    if (tid < 32) {
      // MMA issuer: chunk c is in stage c & 1 once the stagers have arrived on the stage's named barrier; 12 MMAs per
      // chunk, and the commit that frees the stage again
#pragma unroll 1
      for (int c = 0; c < n_chunks; ++c) {
        const int st = c & 1;
        SG_TA(ti);
        if (uses[st] > 0) { mbar_wait(&ms.bar[st], par[st]); par[st] ^= 1; }   // (already retired: the stagers waited for it)
        stage_handover(st, true);
        SG_TB(15, ti);
        if (tid == 0) {
          if (SG_MLP_ISSUER_FENCE) fence_proxy_async();
          tc_fence_after();
          uint8_t* sb = region + (size_t)st * stage_bytes;
          uint8_t* whi = sb + 2 * MLP_ROWS * 128;
          const int slices = min(MLP_KC, s0p8 - c * MLP_KC) / 8;
          const uint64_t dxh = desc_k_sw128(smem_u32(sb)), dxl = desc_k_sw128(smem_u32(sb + MLP_ROWS * 128));
          const uint64_t dwh = desc_k_sw128(smem_u32(whi)), dwl = desc_k_sw128(smem_u32(whi + (size_t)s1n * 128));
          for (int sl = 0; sl < slices; ++sl) {
            const uint64_t adv = (uint64_t)(sl * 2);                           // 32 bytes of K per tf32 MMA
            umma_tf32(tmem, dxl + adv, dwh + adv, idesc, (c | sl) ? 1u : 0u);
            umma_tf32(tmem, dxh + adv, dwl + adv, idesc, 1u);
            umma_tf32(tmem, dxh + adv, dwh + adv, idesc, 1u);
          }
          tc_commit(&ms.bar[st]);
        }
        SG_TB(13, ti);
        __syncwarp();
        ++uses[st];
      }
    } else {
      const int tw = tid - 32;
      const int ch = tw & 7, rb = tw >> 3;
      const float* src[NI];                 // row start (a valid address even when the row is padding)
      uint32_t dhi[NI];                     // byte offset of the hi copy inside a stage (the lo copy: lo_delta(j) further)
      uint32_t flags = 0;                   // bit 2j: the row exists (else it reads as zero); bit 2j + 1: it is stored at all
      auto is_x = [&](int j) { return rb + MLP_STAGER_ROWS * j < MLP_ROWS; };
      auto lo_delta = [&](int j) { return is_x(j) ? (uint32_t)MLP_ROWS * 128u : (uint32_t)s1n * 128u; };
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int R = rb + MLP_STAGER_ROWS * j;
        if (R < MLP_ROWS) {
          const bool real = R < nb;
          src[j] = real ? X + (size_t)ms.idx[R] * s0 : X;
          flags |= (real ? 1u : 0u) << (2 * j) | 2u << (2 * j);
          dhi[j] = (uint32_t)R * 128u + (uint32_t)((ch ^ (R & 7)) << 4);
        } else {
          const int r = R - MLP_ROWS;
          const bool real = r < s1;
          src[j] = real ? W1 + (size_t)r * s0 : W1;
          flags |= (real ? 1u : 0u) << (2 * j) | (r < s1n ? 2u : 0u) << (2 * j);
          dhi[j] = 2u * MLP_ROWS * 128u + (uint32_t)r * 128u + (uint32_t)((ch ^ (r & 7)) << 4);
        }
      }
      if (RING) {
        constexpr int D = MLP_RING_DEPTH;
        constexpr uint32_t SLOT = (MLP_THREADS - 32) * 16;               // one item of every stager
        const uint32_t raw_a = smem_u32(raw) + (uint32_t)tw * 16u;
        // request chunk c into ring slot c % D (always one commit group per call, so the group count stays uniform)
        auto request = [&](int c) {
          if (c < n_chunks) {
            const int k = c * MLP_KC + ch * 4;
            const int kk = min(k, s0 - 4);
            const bool in = k < s0;
#pragma unroll
            for (int j = 0; j < NI; ++j) {
              const uint32_t dst = raw_a + (uint32_t)((c % D) * NI + j) * SLOT;
              const uint32_t ok = (in && (flags >> (2 * j) & 1)) ? 1u : 0u;  // rows / columns out of range arrive as zeros
              if (MLP_STAGER_ROWS * j + MLP_STAGER_ROWS <= MLP_ROWS) {      // always a row of X: 16-byte aligned
                cp_async16(dst, src[j] + kk, 16u * ok);
              } else {                                                      // W1 rows are only 8-byte aligned
                cp_async8(dst, src[j] + kk, 8u * ok);
                cp_async8(dst + 8u, src[j] + kk + 2, 8u * ok);
              }
            }
          }
          cp_async_commit();
        };
#pragma unroll
        for (int c = 0; c < D; ++c) request(c);
#pragma unroll 1
        for (int c = 0; c < n_chunks; ++c) {
          const int st = c & 1;
          uint8_t* sb = region + (size_t)st * stage_bytes;
          if (uses[st] > 0) { mbar_wait(&ms.bar[st], par[st]); par[st] ^= 1; }   // MMAs of chunk c - 2 retired
          cp_async_wait<D - 1>();                                                // chunk c has landed in this thread's slots
#pragma unroll
          for (int j = 0; j < NI; ++j) {
            if (!(flags >> (2 * j) & 2)) continue;
            const float4 x = lds128(raw_a + (uint32_t)((c % D) * NI + j) * SLOT);
            float4 hi, lo;
            split4(x, hi, lo);
            *reinterpret_cast<float4*>(sb + dhi[j]) = hi;
            *reinterpret_cast<float4*>(sb + dhi[j] + lo_delta(j)) = lo;
          }
          if (!SG_MLP_ISSUER_FENCE) fence_proxy_async();
          stage_handover(st, false);
          request(c + D);                                                        // refill the slot just read
          ++uses[st];
        }
        cp_async_wait<0>();
      } else {
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        // Fast path (s0 % 4 == 0, W1 at least 8-byte aligned): every 16-byte chunk is wholly inside or outside a row, so
        // the loads are unconditional (clamped address, result masked in step()) and issue back to back; a branch between
        // two loads, or a use of a loaded value, would serialise them on the memory latency.
        auto fetch = [&](int c, float4 (&v)[NI], int j0, int j1) {                 // items j0 <= j < j1 of chunk c
          const int k = c * MLP_KC + ch * 4;
          if (FAST) {
            const int kk = min(k, s0 - 4);
  #pragma unroll
            for (int j = 0; j < NI; ++j) {
              if (j < j0 || j >= j1) continue;
              if (MLP_STAGER_ROWS * j + MLP_STAGER_ROWS <= MLP_ROWS) {            // always a row of X: 16-byte aligned, read-only
                v[j] = __ldg(reinterpret_cast<const float4*>(src[j] + kk));
              } else {
  #if SG_MLP_W_LDCG   // experiment: explicit ld.global.cg (L2) instead of generic loads for the W1 operand
                const float2 a0 = __ldcg(reinterpret_cast<const float2*>(src[j] + kk)), a1 = __ldcg(reinterpret_cast<const float2*>(src[j] + kk + 2));
  #else
                const float2 a0 = *reinterpret_cast<const float2*>(src[j] + kk), a1 = *reinterpret_cast<const float2*>(src[j] + kk + 2);
  #endif
                v[j] = make_float4(a0.x, a0.y, a1.x, a1.y);
              }
            }
            return;
          }
          const int valid = s0 - k;
  #pragma unroll
          for (int j = 0; j < NI; ++j) {
            if (j < j0 || j >= j1) continue;
            v[j] = (flags >> (2 * j) & 1) ? load4(src[j] + k, valid, is_x(j) ? x16 : w16, is_x(j) ? x16 : w8, false) : zero4;
          }
        };
        // Prefetch depth: ONE chunk of loads in flight per thread.  With two (32 registers of prefetched values) the compiler
        // spills the values it has just loaded - the spill store waits for the load, and the "prefetch" runs at one memory
        // latency per chunk (ncu source view: 7 % of the kernel's stall samples on STL right behind the loads).  Keeping only
        // the first ND items (gathered rows of X: HBM latency) two chunks ahead still spills one float4 per deep item at the
        // 128-register cap of a 512-thread CTA (checked in the SASS for ND = 1, 2), so ND = 0.
        constexpr int ND = SG_MLP_FWD_DEEP_ITEMS;
        // one pipeline step: registers of chunk c -> stage c & 1, hand the stage to warp 0, refill the registers
        auto step = [&](int c, float4 (&vd)[NI], float4 (&vs)[NI]) {
          const int st = c & 1;
          uint8_t* sb = region + (size_t)st * stage_bytes;
          SG_TA(tt);
          if (uses[st] > 0) { mbar_wait(&ms.bar[st], par[st]); par[st] ^= 1; }   // MMAs of chunk c - 2 retired
          SG_TBT(9, tt, 32);
          const bool in = c * MLP_KC + ch * 4 < s0;
  #pragma unroll
          for (int j = 0; j < NI; ++j) {
            if (!(flags >> (2 * j) & 2)) continue;
            float4 x = j < ND ? vd[j] : vs[j];
            if (FAST && !(in && (flags >> (2 * j) & 1))) x = zero4;                // out-of-range rows / columns read as zero
            float4 hi, lo;
            split4(x, hi, lo);
            *reinterpret_cast<float4*>(sb + dhi[j]) = hi;
            *reinterpret_cast<float4*>(sb + dhi[j] + lo_delta(j)) = lo;
          }
          SG_TBT(10, tt, 32);
          if (!SG_MLP_ISSUER_FENCE) fence_proxy_async();   // BEFORE the next loads are issued: the fence would wait for them too
          SG_TBT(11, tt, 32);
          stage_handover(st, false);
          SG_TBT(12, tt, 32);
          if (c + 2 < n_chunks) fetch(c + 2, vd, 0, ND);
          if (c + 1 < n_chunks) fetch(c + 1, vs, ND, NI);
          SG_TBT(14, tt, 32);
          ++uses[st];
        };
        float4 a[NI], b[NI], w[NI];
        fetch(0, a, 0, ND);
        fetch(0, w, ND, NI);
        if (n_chunks > 1) fetch(1, b, 0, ND);
  #pragma unroll 1
        for (int c = 0; c < n_chunks; c += 2) {
          step(c, a, w);
          if (c + 1 < n_chunks) step(c + 1, b, w);
        }
    
      }
    }
    for (int st = 0; st < 2; ++st)
      if (uses[st] > 0) { mbar_wait(&ms.bar[st], par[st]); par[st] ^= 1; }
    tc_fence_after();
    // epilogue: warp w reads TMEM lanes 32 (w & 3) .. + 31 (= batch rows), columns 32 (w >> 2) .. + 31
    {
      const int warp = tid >> 5, lane = tid & 31;
      const int row = (warp & 3) * 32 + lane;
      const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
      const int c_end = min(s1n, (warp >> 2) * 32 + 32);
      for (int c0 = (warp >> 2) * 32; c0 < c_end; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(tl + c0, v);
        tmem_ld_wait();
        if (row < rows) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (c0 + e < s1) Z[row * zs + c0 + e] = __uint_as_float(v[e]) + b1[c0 + e];
        }
      }
    }
    tc_fence_before();
    mlptc::worker_sync();
    if (tid == 0) { ms.parity[0] = par[0]; ms.parity[1] = par[1]; }
    mlptc::worker_sync();
  }

  // ---------------------------------------------------------------------------------------------------
  // dZ1 element (b, j) -> hi / lo copies of the MN-major B operand: K group b / 8, MN atom j / 32
  __device__ static __forceinline__ void put_dz1(uint8_t* bhi, uint8_t* blo, int b, int j, float v) {
    const uint32_t off = (uint32_t)(b >> 3) * 4096u + (uint32_t)(j >> 5) * 1024u + mlptc::mn_chunk_off(b & 7, (j & 31) >> 2) +
                         (uint32_t)(j & 3) * 4u;
    const float hi = mlptc::tf32_rna(v);
    *reinterpret_cast<float*>(bhi + off) = hi;
    *reinterpret_cast<float*>(blo + off) = mlptc::tf32_rna(v - hi);
  }

  // gW1[n][k] += sign * sum_b dZ1[b][n] X[b][k]   (dZ1 hi / lo already staged in `region`)
  __device__ __forceinline__ static void layer1_backward(const BigModel& m, MlpSmem& ms, uint8_t* region, uint8_t* xs, uint8_t* spare,
                                         float sign, int nb, int rows, float* gW, bool overwrite) {
    // the 16-byte-aligned table (s0 % 4 == 0) as its own instantiation, for the reason given in layer1_forward: merged
    // with the element-wise general path the prefetched float4 lives in local memory, and the store that puts it there
    // waits for the load (ncu source view: STL.64 right behind LDG.E.128, 2.5 % of the kernel's stall samples)
    if ((m.dims[1] & 3) == 0) layer1_backward_impl<true>(m, ms, region, xs, spare, sign, nb, rows, gW, overwrite);
    else layer1_backward_impl<false>(m, ms, region, xs, spare, sign, nb, rows, gW, overwrite);
  }
  template <bool X16>
  __device__ __noinline__ static void layer1_backward_impl(const BigModel& m, MlpSmem& ms, uint8_t* region, uint8_t* xs, uint8_t* spare,
                                         float sign, int nb, int rows, float* gW, bool overwrite) {
    using namespace mlptc;
    const int s0 = m.dims[1], s1 = m.dims[2];
    const int s1n = mlp_n16(s1);
    const int tid = threadIdx.x;
    const float* X = reinterpret_cast<const float*>(m.data0);
    constexpr bool x16 = X16;
    const int G = (nb + 7) >> 3;                                  // K groups of 8 batch rows
    const int n_tiles = (s0 + MLP_BWD_TILE - 1) / MLP_BWD_TILE;   // M tiles of 120 input columns (see the stager mapping)
    const int tiles_per_pass = 4;                                 // 4 s1n <= 512 TMEM columns
    const uint32_t idesc = idesc_tf32(s1n, true);
    const uint32_t tmem = ms.tmem_base;
    const uint8_t* bhi = region;
    const uint8_t* blo = region + (size_t)(rows / 8) * 4096;
    constexpr uint32_t stage_bytes = 2 * 8 * 256 * sizeof(float);      // [hi 8 KB | lo 8 KB]
    // Stage ring: the two stages of xs, plus two more in `spare` (the activation buffers, dead by now) when they fit.
    // One unit's chain - wait for the stage, st.shared, fence, hand-over, 6 MMAs, commit - is ~3 us of latencies; the
    // depth of the ring is how many of them overlap.
    const int nst = spare ? 4 : 2;
    auto stage_ptr = [&](int st) { return st < 2 ? xs + (size_t)st * stage_bytes : spare + (size_t)(st - 2) * stage_bytes; };
    uint32_t par = ms.parity[0] | ms.parity[1] << 1 | ms.parity[2] << 2 | ms.parity[3] << 3;   // bit st: next phase parity of bar[st]

    // Warp 0 only issues the MMAs (as in layer1_forward); the 480 threads of warps 1..15 stage X.  So that they divide a
    // unit evenly, an M tile holds MLP_BWD_TILE = 120 input columns (lanes 120..127 of its accumulators are never read):
    // one unit = 8 batch rows x 2 tiles x 30 sixteen-byte chunks = 480 slots, one per stager.
    const int tw = tid >= 32 ? tid - 32 : 0;
    const int lr = tw / 60, c60 = tw - lr * 60;
    const int tl = c60 / 30, cl = c60 - tl * 30;               // tile of the pair, chunk inside the tile
    const uint32_t soff = (uint32_t)(tl * 4 + (cl >> 3)) * 1024u + mn_chunk_off(lr, cl & 7);
    {                                                             // the never-written chunks 30, 31 of every tile row: zero
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int st = 0; st < nst; ++st)
        for (int o = tid; o < (int)(stage_bytes / 16); o += MLP_THREADS) reinterpret_cast<float4*>(stage_ptr(st))[o] = z;
      fence_proxy_async();
      mlptc::worker_sync();
    }

    for (int t0 = 0; t0 < n_tiles; t0 += tiles_per_pass) {
      const int pass_tiles = min(tiles_per_pass, n_tiles - t0);
      const int npp = (pass_tiles + 1) >> 1;                      // tile pairs in this pass
      const int units = G * npp;
      uint32_t used = 0;                                          // bit st: stage st has MMAs in flight
      if (tid < 32) {
#pragma unroll 1
        for (int u = 0; u < units; ++u) {
          const int st = u & (nst - 1);
          const uint8_t* sb = stage_ptr(st);
          if (used >> st & 1) { mbar_wait(&ms.bar[st], par >> st & 1); par ^= 1u << st; }   // (retired: the stagers waited for it)
          stage_handover(st, true);
          if (tid == 0) {
            if (SG_MLP_ISSUER_FENCE) fence_proxy_async();
            tc_fence_after();
            const int g = npp == 2 ? u >> 1 : u, p = npp == 2 ? u & 1 : 0;       // npp <= 2: tiles_per_pass == 4
            const uint64_t dbh = desc_mn_sw128_32b(smem_u32(bhi) + (uint32_t)g * 4096u, 1024u, 512u);
            const uint64_t dbl = desc_mn_sw128_32b(smem_u32(blo) + (uint32_t)g * 4096u, 1024u, 512u);
            for (int t = 0; t < 2; ++t) {
              if (2 * p + t >= pass_tiles) break;
              const uint64_t dah = desc_mn_sw128_32b(smem_u32(sb) + (uint32_t)t * 4096u, 1024u, 512u);
              const uint64_t dal = desc_mn_sw128_32b(smem_u32(sb) + 8192u + (uint32_t)t * 4096u, 1024u, 512u);
              const uint32_t acc = tmem + (uint32_t)((2 * p + t) * s1n);
              umma_tf32(acc, dal, dbh, idesc, g ? 1u : 0u);
              umma_tf32(acc, dah, dbl, idesc, 1u);
              umma_tf32(acc, dah, dbh, idesc, 1u);
            }
            tc_commit(&ms.bar[st]);
          }
          __syncwarp();
          used |= 1u << st;
        }
      } else {
        // unit u = (K group g, tile pair p); npp <= 2 because tiles_per_pass == 4
        auto group_of = [&](int u) { return npp == 2 ? u >> 1 : u; };
        auto pair_of = [&](int u) { return npp == 2 ? u & 1 : 0; };
        // table row of this thread's slot in unit u (read ahead of the load that needs it: ms.idx sits behind a generic
        // pointer, and the gather waited for it - 3 % of the kernel's stall samples on the LDG)
        auto row_of = [&](int u) { return ms.idx[min(group_of(u) * 8 + lr, nb - 1)]; };
        auto fetch = [&](int u, int row, float4& xv) {
          const int b = group_of(u) * 8 + lr;
          const int k = (t0 + 2 * pair_of(u) + tl) * MLP_BWD_TILE + cl * 4;
          if (x16) {                                             // unconditional load, clamped address; masked in step()
            xv = __ldg(reinterpret_cast<const float4*>(X + (size_t)row * s0 + min(k, s0 - 4)));
            return;
          }
          xv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (b < nb && k < s0) xv = load4(X + (size_t)row * s0 + k, s0 - k, x16, x16, true);
        };
        // one pipeline step: registers of unit u -> its ring stage, hand the stage to warp 0, refill them with unit u + 4
        auto step = [&](int u, float4& xv) {
          const int st = u & (nst - 1);
          uint8_t* sb = stage_ptr(st);
          const int next_row = u + 4 < units ? row_of(u + 4) : 0;
          SG_TA(tb);
          if (used >> st & 1) { mbar_wait(&ms.bar[st], par >> st & 1); par ^= 1u << st; }
          SG_TBT(21, tb, 32);
          if (x16) {
            if (!(group_of(u) * 8 + lr < nb && (t0 + 2 * pair_of(u) + tl) * MLP_BWD_TILE + cl * 4 < s0)) xv = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float4 hi, lo;
          split4(xv, hi, lo);
          *reinterpret_cast<float4*>(sb + soff) = hi;
          *reinterpret_cast<float4*>(sb + 8192 + soff) = lo;
          SG_TBT(22, tb, 32);
          if (!SG_MLP_ISSUER_FENCE) fence_proxy_async();         // before the next load is issued (the fence would wait for it)
          SG_TBT(23, tb, 32);
          stage_handover(st, false);
          SG_TBT(24, tb, 32);
          if (u + 4 < units) fetch(u + 4, next_row, xv);         // four units of loads stay in flight
          SG_TBT(25, tb, 32);
          used |= 1u << st;
        };
        float4 x0, x1, x2, x3;
        fetch(0, row_of(0), x0);
        if (units > 1) fetch(1, row_of(1), x1);
        if (units > 2) fetch(2, row_of(2), x2);
        if (units > 3) fetch(3, row_of(3), x3);
#pragma unroll 1
        for (int u = 0; u < units; u += 4) {
          step(u, x0);
          if (u + 1 < units) step(u + 1, x1);
          if (u + 2 < units) step(u + 2, x2);
          if (u + 3 < units) step(u + 3, x3);
        }
      }
      SG_TA(te);
      for (int st = 0; st < 4; ++st)
        if (used >> st & 1) { mbar_wait(&ms.bar[st], par >> st & 1); par ^= 1u << st; }
      tc_fence_after();
      SG_TB(27, te);
      // epilogue: warp group (warp >> 2) owns tile t0 + (warp >> 2); lane = input column k, TMEM column = unit n
      {
        const int warp = tid >> 5, lane = tid & 31;
        const int tq = warp >> 2;
        const int kl = (warp & 3) * 32 + lane;
        const int k = kl < MLP_BWD_TILE ? (t0 + tq) * MLP_BWD_TILE + kl : s0;
        if (tq < pass_tiles) {
          const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tq * s1n);
          for (int n0 = 0; n0 < s1n; n0 += 16) {
            uint32_t v[16];
            tmem_ld16(tl + n0, v);
            tmem_ld_wait();
            if (k < s0) {
              if (overwrite) {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                  if (n0 + e < s1) gW[(size_t)(n0 + e) * s0 + k] = sign * __uint_as_float(v[e]);
              } else {
                float old[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) old[e] = n0 + e < s1 ? __ldcg(gW + (size_t)(n0 + e) * s0 + k) : 0.f;
#pragma unroll
                for (int e = 0; e < 16; ++e)
                  if (n0 + e < s1) gW[(size_t)(n0 + e) * s0 + k] = fmaf(sign, __uint_as_float(v[e]), old[e]);
              }
            }
          }
        }
      }
      SG_TB(28, te);
      tc_fence_before();
      mlptc::worker_sync();                                            // TMEM tiles are free for the next pass
      SG_TB(29, te);
    }
    if (tid == 0) { ms.parity[0] = par & 1; ms.parity[1] = par >> 1 & 1; ms.parity[2] = par >> 2 & 1; ms.parity[3] = par >> 3 & 1; }
    mlptc::worker_sync();
  }

  // ---------------------------------------------------------------------------------------------------
  // one sub-batch of nb <= plan.rows rows whose table indices are in ms.idx[0..nb)
  __device__ __noinline__ static void sub_batch(const BigModel& m, MlpSmem& ms, uint8_t* base, const float* th, float sign,
                                                int nb, float* g, bool overwrite, float* lp = nullptr) {
    const int L = m.dims[0];
    const int* sz = m.dims + 1;                       // sz[0..L]
    const float* Y = reinterpret_cast<const float*>(m.data1);
    const int tid = threadIdx.x;
    const MlpPlan plan = mlp_plan(m.dims, m.batch);
    const int R = plan.rows;
    uint8_t* region = base;
    uint8_t* xs = base + plan.region_bytes;
    float* fbase = reinterpret_cast<float*>(xs + plan.xs_bytes);
    auto act = [&](int l) { return fbase + plan.act_off(l); };
    auto astr = [&](int l) { return mlp_stride(sz[l]); };
    auto dzbuf = [&](int k) { return fbase + plan.dz_off(k); };
    auto dzstr = [&](int k) { return mlp_stride(plan.dz_w(k)); };

    // The labels of the sub-batch (Y rows: a dependent gather from HBM) are requested before the forward GEMM and parked in
    // the output layer's dZ buffer after it, so that the output layer finds them in shared memory.
    float* const ybuf = dzbuf(0);
    const int ybs = dzstr(0), wL = sz[L];
    float ypre[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int o = tid + t * MLP_THREADS;
      ypre[t] = o < nb * wL ? __ldg(Y + (size_t)ms.idx[o / wL] * wL + o % wL) : 0.f;
    }
    // shared memory that is idle during the forward GEMM: the backward X stages and every fp32 activation / dZ buffer
    // (act(1) is only written by the GEMM's epilogue, after the ring has drained)
    const size_t raw_bytes = plan.total_bytes - 1024 - plan.region_bytes;
    { SG_T0(); layer1_forward(m, ms, region, th, nb, R, act(1), astr(1), xs, raw_bytes); SG_T1(1); }
    SG_T0();
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int o = tid + t * MLP_THREADS;
      if (o < nb * wL) ybuf[(o / wL) * ybs + o % wL] = ypre[t];
    }
    for (int o = tid + 4 * MLP_THREADS; o < nb * wL; o += MLP_THREADS) ybuf[(o / wL) * ybs + o % wL] = __ldg(Y + (size_t)ms.idx[o / wL] * wL + o % wL);
    // W_2 (sz[2] x sz[1] floats, 4 KB for 784-100-10) is read by every thread of the layer-2 forward and of the dZ_1
    // back-propagation: from global memory those loops are bound by L2 latency.  Stage it once per sub-batch, rows padded
    // with zeros to MLP_W2S = 128 columns (so the loops over its columns need no bounds checks: a bounds check in front
    // of every load becomes a branch, and a load behind a branch is not hoisted - the loop then runs at one memory
    // latency per multiply-add), in the part of xs the bias-gradient scratch (first 8 KB) leaves free; xs is idle until
    // layer1_backward streams X.
    float* w2s = nullptr;
    if (L >= 2 && (size_t)sz[2] * MLP_W2S * sizeof(float) + 8192 <= plan.xs_bytes) {
      w2s = reinterpret_cast<float*>(xs + 8192);
      const float* W2g = th + m.leaf_off[2];
      for (int o = tid; o < sz[2] * MLP_W2S; o += MLP_THREADS) {
        const int n = o / MLP_W2S, j = o % MLP_W2S;
        w2s[o] = j < sz[1] ? W2g[n * sz[1] + j] : 0.f;
      }
    }

    // =================================================================== softmax + small layers forward
    for (int l = 1; l < L; ++l) {
      float* A = act(l);
      const int as = astr(l), w = sz[l];
      SG_T0();
      {                                                 // a_l = softmax(z_l): four threads per row, columns q, q + 4, ...
        const int b = tid >> 2, q = tid & 3;
        if (b < R) {                                    // warp-uniform: R is a multiple of 8
          float* row = A + b * as;
          float mx = -INFINITY;
          for (int c = q; c < w; c += 4) mx = fmaxf(mx, row[c]);
          mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(FULL, mx, 2));
          float sum = 0.f;
          for (int c = q; c < w; c += 4) { const float e = expf(row[c] - mx); row[c] = e; sum += e; }
          sum += __shfl_xor_sync(FULL, sum, 1);
          sum += __shfl_xor_sync(FULL, sum, 2);
          const float inv = 1.0f / sum;
          for (int c = q; c < w; c += 4) row[c] *= inv;
        }
      }
      mlptc::worker_sync();
      SG_T1(4);
      float* Zn = act(l + 1);
      const int zs = astr(l + 1), wn = sz[l + 1];
      const float* bn = th + m.leaf_off[2 * l + 1];
      if (l == 1 && w2s) {
        // z_2[b][n] = a_1[b] . W_2[n] + b_2[n]: row b = tid & 127, four outputs n = (tid >> 7) + 4 t at a time (the W_2 rows
        // are warp-uniform: broadcast 16-byte reads of the staged copy)
        const int b = tid & 127, ng = tid >> 7;
        if (b < R) {
          const float* arow = A + b * as;
          for (int n0 = ng; n0 < wn; n0 += 16) {
            const float* wr[4];
            float accv[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const int n = min(n0 + 4 * t, wn - 1);
              wr[t] = w2s + n * MLP_W2S;
              accv[t] = bn[n];
            }
            int k = 0;
            for (; k + 4 <= w; k += 4) {
              const float a0 = arow[k], a1 = arow[k + 1], a2 = arow[k + 2], a3 = arow[k + 3];
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const float4 wv = *reinterpret_cast<const float4*>(wr[t] + k);
                accv[t] = fmaf(a0, wv.x, accv[t]); accv[t] = fmaf(a1, wv.y, accv[t]);
                accv[t] = fmaf(a2, wv.z, accv[t]); accv[t] = fmaf(a3, wv.w, accv[t]);
              }
            }
            for (; k < w; ++k) {
#pragma unroll
              for (int t = 0; t < 4; ++t) accv[t] = fmaf(arow[k], wr[t][k], accv[t]);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (n0 + 4 * t < wn) Zn[b * zs + n0 + 4 * t] = accv[t];
          }
        }
      } else {
        const float* Wn = th + m.leaf_off[2 * l];
        for (int o = tid; o < R * wn; o += MLP_THREADS) {            // z_{l+1}[b][n] = a_l[b] . W[n] + b[n]
          const int b = o % R, n = o / R;
          float accv = 0.f;
          const float* wr = Wn + (size_t)n * w;
          for (int k = 0; k < w; ++k) accv = fmaf(A[b * as + k], wr[k], accv);
          Zn[b * zs + n] = accv + bn[n];
        }
      }
      mlptc::worker_sync();
      SG_T1(5);
    }

    // =================================================================== output layer: dZ_L = y - sum(y) softmax(z_L)
    float* dz = dzbuf(0);
    int dzs = dzstr(0);
    {
      SG_T0();
      const int w = sz[L], zs = astr(L);
      const float* Z = act(L);
      if (tid < R) {
        if (tid < nb) {
          const float* yr = dz + tid * dzs;            // the labels parked here after the forward GEMM; overwritten by dZ_L
          float mx = -INFINITY, ys = 0.f;
          for (int c = 0; c < w; ++c) { mx = fmaxf(mx, Z[tid * zs + c]); ys += yr[c]; }
          float sum = 0.f;
          for (int c = 0; c < w; ++c) sum += expf(Z[tid * zs + c] - mx);
          const float inv = ys / sum;
          if (lp) {                                    // loglik = sum_c y_c log_softmax(z)_c  (NN_model.py:48-50)
            const float lse = mx + logf(sum);
            float v = 0.f;
            for (int c = 0; c < w; ++c) v = fmaf(yr[c], Z[tid * zs + c] - lse, v);
            atomicAdd(lp, v);
          }
          for (int c = 0; c < w; ++c) dz[tid * dzs + c] = yr[c] - inv * expf(Z[tid * zs + c] - mx);
        } else {
          for (int c = 0; c < w; ++c) dz[tid * dzs + c] = 0.f;      // padded rows contribute nothing
        }
      }
      mlptc::worker_sync();
      SG_T1(6);
    }

    uint8_t* bhi = region;
    uint8_t* blo = region + (size_t)(R / 8) * 4096;
    const int s1 = sz[1], s1n = mlp_n16(s1);
    float* gb1 = g + m.leaf_off[1];
    float* red = reinterpret_cast<float*>(xs);          // [16 warps][128]: free until layer1_backward streams X

    // =================================================================== backward through layers L .. 2
    int cur = 0;
    for (int l = L; l >= 2; --l) {
      const int w = sz[l], wp = sz[l - 1];
      const float* A = act(l - 1);
      const int as = astr(l - 1);
      const float* Wl = (l == 2 && w2s) ? w2s : th + m.leaf_off[2 * (l - 1)];
      float* gW = g + m.leaf_off[2 * (l - 1)];
      float* gb = g + m.leaf_off[2 * (l - 1) + 1];
      SG_T0();
      // dW_l[n][k] = sum_b dz[b][n] a_{l-1}[b][k]: two rows n per item (independent accumulation chains), the old values
      // requested before the loop; the bias sums go to the threads at the far end, which have the fewest items
      const int npairs = (w + 1) >> 1;
      for (int o = tid; o < npairs * wp; o += MLP_THREADS) {
        const int k = o % wp, n0 = 2 * (o / wp), n1 = min(n0 + 1, w - 1);
        const float old0 = gW[n0 * wp + k], old1 = gW[n1 * wp + k];
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll 4
        for (int b = 0; b < nb; ++b) {
          const float a = A[b * as + k];
          acc0 = fmaf(dz[b * dzs + n0], a, acc0);
          acc1 = fmaf(dz[b * dzs + n1], a, acc1);
        }
        gW[n0 * wp + k] = fmaf(sign, acc0, old0);
        if (n1 != n0) gW[n1 * wp + k] = fmaf(sign, acc1, old1);
      }
      for (int n = MLP_THREADS - 1 - tid; n < w; n += MLP_THREADS) {
        const float old0 = gb[n];
        float accv = 0.f;
#pragma unroll 4
        for (int b = 0; b < nb; ++b) accv += dz[b * dzs + n];
        gb[n] = fmaf(sign, accv, old0);
      }
      SG_T1(7);
      if (l > 2) {
        // dA_{l-1} = dz W_l ; dZ_{l-1} = a (dA - sum_k a dA)   (softmax Jacobian)
        float* dn = dzbuf(cur ^ 1);
        const int dns = dzstr(cur ^ 1);
        for (int o = tid; o < R * wp; o += MLP_THREADS) {
          const int b = o % R, k = o / R;
          float accv = 0.f;
          for (int n = 0; n < w; ++n) accv = fmaf(dz[b * dzs + n], Wl[(size_t)n * wp + k], accv);
          dn[b * dns + k] = accv;
        }
        mlptc::worker_sync();
        if (tid < R) {
          float dotv = 0.f;
          for (int k = 0; k < wp; ++k) dotv = fmaf(A[tid * as + k], dn[tid * dns + k], dotv);
          for (int k = 0; k < wp; ++k) dn[tid * dns + k] = tid < nb ? A[tid * as + k] * (dn[tid * dns + k] - dotv) : 0.f;
        }
        mlptc::worker_sync();
        dz = dn; dzs = dns; cur ^= 1;
      } else {
        // l == 2: dZ_1 of row b by a quad of threads, written straight into the tensor-core operand (hi / lo); bias
        // gradient by fixed-order shuffles + a per-warp fold.
        const int b = tid >> 2, q = tid & 3;
        const bool live = b < nb;
        const int warp = tid >> 5, lane = tid & 31;
        SG_TA(tq);
        if (w2s) {
          // thread q of the quad owns the 16-byte column chunks q, q + 4, ...: columns j = 16 t + 4 q + e.  All 128 columns
          // of the padded W_2 copy are swept (zeros beyond s1), so no load sits behind a bounds check.
          // dA[b][j] = sum_n dz[b][n] W_2[n][j]: one sweep over n, 32 accumulators
          float da[8][4];
#pragma unroll
          for (int t = 0; t < 8; ++t) da[t][0] = da[t][1] = da[t][2] = da[t][3] = 0.f;
          if (live) {
#pragma unroll 2
            for (int n = 0; n < w; ++n) {
              const float dzn = dz[b * dzs + n];
              const float4* wrow = reinterpret_cast<const float4*>(w2s + n * MLP_W2S) + q;
#pragma unroll
              for (int t = 0; t < 8; ++t) {
                const float4 wv = wrow[4 * t];
                da[t][0] = fmaf(dzn, wv.x, da[t][0]); da[t][1] = fmaf(dzn, wv.y, da[t][1]);
                da[t][2] = fmaf(dzn, wv.z, da[t][2]); da[t][3] = fmaf(dzn, wv.w, da[t][3]);
              }
            }
          }
          SG_TB(16, tq);
          // (reads of a_1 beyond its row stay inside the activation buffers; they are discarded by the select)
          const float* arow = A + (live ? b : 0) * as;
          float dotv = 0.f;
#pragma unroll
          for (int t = 0; t < 8; ++t)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = 16 * t + 4 * q + e;
              const float av = arow[j];
              dotv = fmaf(j < wp ? av : 0.f, da[t][e], dotv);
            }
          dotv += __shfl_xor_sync(FULL, dotv, 1);
          dotv += __shfl_xor_sync(FULL, dotv, 2);
          SG_TB(17, tq);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int j0 = 16 * t + 4 * q;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float av = arow[j0 + e];
              v[e] = (live && j0 + e < wp) ? av * (da[t][e] - dotv) : 0.f;
            }
            if (b < R) {                                           // columns in [s1n, 128) land in the operand's padding
              const uint32_t off = (uint32_t)(b >> 3) * 4096u + (uint32_t)(j0 >> 5) * 1024u + mlptc::mn_chunk_off(b & 7, (j0 & 31) >> 2);
              float4 hi, lo;
              mlptc::split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
              *reinterpret_cast<float4*>(bhi + off) = hi;
              *reinterpret_cast<float4*>(blo + off) = lo;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              v[e] += __shfl_xor_sync(FULL, v[e], 4);
              v[e] += __shfl_xor_sync(FULL, v[e], 8);
              v[e] += __shfl_xor_sync(FULL, v[e], 16);
            }
            if (lane < 4) *reinterpret_cast<float4*>(red + warp * 128 + j0) = make_float4(v[0], v[1], v[2], v[3]);
          }
          SG_TB(18, tq);
        } else {
          // W_2 too large to stage: columns j = q, q + 4, ... of the global copy
          float da[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) da[i] = 0.f;
          if (live) {
#pragma unroll 1
            for (int n = 0; n < w; ++n) {
              const float dzn = dz[b * dzs + n];
              const float* wrow = Wl + (size_t)n * wp + q;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (4 * i < s1n) da[i] = fmaf(dzn, 4 * i + q < wp ? wrow[4 * i] : 0.f, da[i]);
            }
          }
          float dotv = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (4 * i < s1n) dotv = fmaf((live && 4 * i + q < wp) ? A[b * as + 4 * i + q] : 0.f, da[i], dotv);
          dotv += __shfl_xor_sync(FULL, dotv, 1);
          dotv += __shfl_xor_sync(FULL, dotv, 2);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = q + 4 * i;
            if (4 * i < s1n) {                                       // warp-uniform
              float v = (live && j < wp) ? A[b * as + j] * (da[i] - dotv) : 0.f;
              if (b < R) put_dz1(bhi, blo, b, j, v);
              v += __shfl_xor_sync(FULL, v, 4);
              v += __shfl_xor_sync(FULL, v, 8);
              v += __shfl_xor_sync(FULL, v, 16);
              if (lane < 4) red[warp * 128 + j] = v;
            }
          }
        }
        mlptc::worker_sync();
        SG_TB(19, tq);
        for (int n = tid; n < s1; n += MLP_THREADS) {
          float accv = 0.f;
          for (int w16 = 0; w16 < 16; ++w16) accv += red[w16 * 128 + n];
          gb1[n] += sign * accv;
        }
        SG_TB(20, tq);
        SG_T1(8);
      }
    }
    if (L == 1) {                                                   // the output layer is layer 1: convert its fp32 dZ
      for (int n = tid; n < s1; n += MLP_THREADS) {
        float accv = 0.f;
        for (int b = 0; b < nb; ++b) accv += dz[b * dzs + n];
        gb1[n] += sign * accv;
      }
      for (int e = tid; e < R * s1n; e += MLP_THREADS) {
        const int b = e / s1n, j = e - b * s1n;
        put_dz1(bhi, blo, b, j, (j < s1 && b < nb) ? dz[b * dzs + j] : 0.f);
      }
    }
    mlptc::fence_proxy_async();
    mlptc::worker_sync();

    SG_T1(2);
    // the fp32 activation / dZ buffers are dead from here on: two more X stages for layer1_backward if they hold them
    uint8_t* spare = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fbase) + 1023) & ~(uintptr_t)1023);
    if (spare + 2 * (2 * 8 * 256 * sizeof(float)) > base + plan.total_bytes - 1024) spare = nullptr;
    { SG_T0(); layer1_backward(m, ms, region, xs, spare, sign, nb, R, g + m.leaf_off[0], overwrite); SG_T1(3); }
  }

  __device__ static void lik_pass(const BigModel& m, void* fam_smem, const float* theta, const float* center,
                                  bool sequential, const RandintPlan& plan, unsigned seq_base, unsigned count, float* g,
                                  bool fresh, const NoiseJob& job, float* lp = nullptr, bool /*compact*/ = false) {
    if (threadIdx.x >= MLP_THREADS) {
      // noise-producer warps: the Gaussian draws of the NEXT diffusion update (its key tree is data independent,
      // diffusion_util.py:66), generated while the worker warps run the GEMM phase; the caller's __syncthreads() after
      // the pass publishes them
      if (job.valid) {
        Key k1, k2;
        split2(job.key, k1, k2);
        for (int leaf = 0; leaf < m.n_leaves; ++leaf)
          produce_leaf_noise((unsigned)(m.leaf_off[leaf + 1] - m.leaf_off[leaf]), split_at(k1, m.n_leaves, leaf),
                             job.dst + m.leaf_off[leaf], threadIdx.x - MLP_THREADS, blockDim.x - MLP_THREADS);
      }
      return;
    }
    MlpSmem& ms = *reinterpret_cast<MlpSmem*>(fam_smem);
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(&ms + 1) + 1023) & ~(uintptr_t)1023);
    const unsigned rows = (unsigned)mlp_plan(m.dims, m.batch).rows;
    const unsigned h = (count + 1u) >> 1;
    for (unsigned base_row = 0; base_row < count; base_row += rows) {
      const int nb = (int)min(rows, count - base_row);
      mlptc::worker_sync();
      if (threadIdx.x < (unsigned)nb) {
        const unsigned pos = base_row + threadIdx.x;
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
      mlptc::worker_sync();
      // the first sub-batch into an all-zero buffer stores dW1 instead of read-modify-writing it
      sub_batch(m, ms, base, theta, 1.0f, nb, g, fresh && base_row == 0, lp);
      if (center) sub_batch(m, ms, base, center, -1.0f, nb, g, false);
    }
  }
};

}  // namespace sg
