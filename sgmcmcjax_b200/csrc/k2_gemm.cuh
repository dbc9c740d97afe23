// K2 on tensor cores: full-batch grad_log_post of the logistic-regression family at M parameter vectors
// (util.py:43-49 with all rows; SURVEY.md section 8 f1: the `grads` producer of imq_KSD, CV / SVRG centres).
//
//   out[m] = sum_i w[m][i] x~_i - prec * theta_m,     w[m][i] = -sigmoid(theta_m . x~_i),   x~ = (1 - 2y) x
//
// For M = 5e4 sample points and N = 1e6 rows that is two [M x N x D] contractions (1e13 flop each) - out of reach
// of the row-streaming SIMT kernel (M passes over the 400 MB table).  Here both run on tcgen05 as 3xTF32, with the
// same TMA ring / TMEM / warp-role skeleton as csrc/ksd_tc.cuh, chunked over the rows (Nc rows at a time):
//
//   k2_scores_kernel   S = Theta . X~^T on a 128 x 128 tile (K = D), epilogue w = -sigmoid(S) -> tf32 hi / lo, written
//                      chunk-major ([2][i-chunks][M][16]) so that it is the K-major A operand of the second kernel.
//                      Two sets of TMEM accumulators (2 x 256 columns): the MMAs of tile t + 1 run under the epilogue of
//                      tile t.  An epilogue warp's 32 rows x 16 columns are 2 KB of CONTIGUOUS W: staged in shared
//                      memory (bank-conflict-free rotated 16-byte chunks) and written by one cp.async.bulk - the
//                      per-lane stores of the first version (16 bytes at a 64-byte stride: 32 write transactions per
//                      warp instruction) cost the load/store unit ~8 k of the 16 k cycles of a tile.
//   k2_accum_kernel    G[m-tile][D] += W . X~ (K = the Nc rows of the chunk, 16 per stage), B operand = X~^T split and
//                      stored [2][i-chunks][DN][16] once per data table; CTAs split K, fp32 atomics into G.
//
// Operand prep (once per table): X~ split K-major chunk-major `Xs[2][chunks][N][16]` and transposed `XT[2][N/16][DN][16]`.
#pragma once
#include "ksd_tc.cuh"
#include "vec_sampler.cuh"

namespace sg {
namespace k2 {
using namespace ksdtc;

constexpr int DN_MAX = 128;            // padded feature count (MMA N of the second contraction), D <= 128
constexpr int A_STAGES = 4;            // scores kernel: 32 KB stages
constexpr int A_STAGE_BYTES = 4 * CHUNK_BYTES;
constexpr int B_STAGES = 6;            // accumulate kernel: <= 32 KB stages
constexpr int THREADS = 576;           // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue

struct TailA { uint64_t full[A_STAGES], empty[A_STAGES], tmem_full[2], tmem_empty[2]; uint32_t tmem_base, pad; };
constexpr int A_EPI_WARPS = 16;
constexpr size_t A_STAGING = (size_t)A_EPI_WARPS * 2 * 32 * KC * 4;      // per epilogue warp: hi and lo block of 32 rows x 16 floats
struct TailB { uint64_t full[B_STAGES], empty[B_STAGES], tmem_full; uint32_t tmem_base, pad; };
constexpr size_t SMEM_A = 1024 + (size_t)A_STAGES * A_STAGE_BYTES + A_STAGING + sizeof(TailA);

// logistic function with the MUFU exponential and reciprocal (relative error ~3e-7: inside the 2e-5 parity bar of K2)
__device__ __forceinline__ float sigmoid_mufu(float z) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-1.44269504089f * fabsf(z)));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + t));
  return z >= 0.0f ? r : t * r;
}
// 16-byte chunk c (0..3) of sixteen floats, c a run-time value: selects instead of a dynamically indexed register array
__device__ __forceinline__ float4 chunk_of(const float (&v)[16], int c) {
  float4 r;
  r.x = c == 0 ? v[0] : c == 1 ? v[4] : c == 2 ? v[8] : v[12];
  r.y = c == 0 ? v[1] : c == 1 ? v[5] : c == 2 ? v[9] : v[13];
  r.z = c == 0 ? v[2] : c == 1 ? v[6] : c == 2 ? v[10] : v[14];
  r.w = c == 0 ? v[3] : c == 1 ? v[7] : c == 2 ? v[11] : v[15];
  return r;
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
inline size_t b_stage_bytes(int dn) { return 2 * (size_t)CHUNK_BYTES + 2 * (size_t)dn * KC * 4; }
inline size_t smem_b(int dn) { return 1024 + (size_t)B_STAGES * b_stage_bytes(dn) + sizeof(TailB); }

__host__ __device__ constexpr uint32_t idesc_n(int n) { return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

inline int feat_chunks(int d) { return (padded_dim(d) + KC - 1) / KC; }
inline int feat_rows(int d) { return (d + 15) & ~15; }                       // DN

// ------------------------------------------------------------------------------------------ operand prep
// rows[n][ds] -> Xs[2][chunks][n][16] (hi, lo)
__global__ void split_rows_kernel(const float* __restrict__ rows, size_t n, int ds, int d, int chunks, float* __restrict__ Xs) {
  const size_t total = n * (size_t)chunks * KC;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(t % KC);
    const size_t r = (t / KC) % n;
    const int c = (int)(t / ((size_t)KC * n));
    const int col = c * KC + w;
    const float v = col < d ? rows[r * ds + col] : 0.f;
    const float hi = tf32_rna(v);
    Xs[t] = hi;
    Xs[total + t] = tf32_rna(v - hi);
  }
}
// rows[n][ds] -> XT[2][ceil(n/16)][dn][16]: XT[seg][ic][col][ii] = split(rows[ic * 16 + ii][col])
__global__ void transpose_rows_kernel(const float* __restrict__ rows, size_t n, int ds, int d, int dn, float* __restrict__ XT) {
  const size_t nic = (n + KC - 1) / KC;
  const size_t total = nic * (size_t)dn * KC;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int ii = (int)(t % KC);
    const int col = (int)((t / KC) % dn);
    const size_t ic = t / ((size_t)KC * dn);
    const size_t r = ic * KC + ii;
    const float v = (r < n && col < d) ? rows[r * ds + col] : 0.f;
    const float hi = tf32_rna(v);
    XT[t] = hi;
    XT[total + t] = tf32_rna(v - hi);
  }
}
// thetas[m][P] -> Ts[2][chunks][m][16]
__global__ void split_thetas_kernel(const float* __restrict__ th, int m, int P, int d, int chunks, float* __restrict__ Ts) {
  const size_t total = (size_t)m * chunks * KC;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(t % KC);
    const int r = (int)((t / KC) % m);
    const int c = (int)(t / ((size_t)KC * m));
    const int col = c * KC + w;
    const float v = col < d ? th[(size_t)r * P + col] : 0.f;
    const float hi = tf32_rna(v);
    Ts[t] = hi;
    Ts[total + t] = tf32_rna(v - hi);
  }
}
// out[m][e] = G[m][e] - prec * theta[m][e]
__global__ void finish_kernel(const float* __restrict__ G, const float* __restrict__ th, int m, int P, float prec, float* __restrict__ out) {
  const size_t total = (size_t)m * P;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
    out[t] = G[t] - prec * th[t];
}

// ------------------------------------------------------------------------------------------ scores
// tmap: Theta split as a [.., 16] matrix (rows (seg * chunks + c) * m_all + r); xmap: Xs likewise with n rows per block.
// This launch: parameter vectors [m0, m0 + m) x rows [i0, i0 + nc);  W[seg][ic][m][16] for the ceil(nc / 16) row chunks.
__global__ void __launch_bounds__(THREADS, 1)
k2_scores_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap xmap, int m_all, int m0, int m,
                 int n, int chunks, int dp, int i0, int nc, float* __restrict__ W) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + (size_t)A_STAGES * A_STAGE_BYTES;
  TailA* tail = reinterpret_cast<TailA*>(staging + A_STAGING);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int MT = (m + TILE - 1) / TILE, IT = (nc + TILE - 1) / TILE;
  const long long total = (long long)MT * IT;
  const int nic = (nc + KC - 1) / KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < A_STAGES; ++s) { mbar_init(&tail->full[s], 1); mbar_init(&tail->empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tail->tmem_full[b], 1); mbar_init(&tail->tmem_empty[b], A_EPI_WARPS); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tail->tmem_base, 512); tmem_relinquish(); }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmap); tma_prefetch_desc(&xmap); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x) {
        const int mt = (int)(t / IT), it = (int)(t % IT);
        for (int c = 0; c < chunks; ++c) {
          mbar_wait(&tail->empty[stage], phase ^ 1);
          uint8_t* sb = smem + (size_t)stage * A_STAGE_BYTES;
          mbar_arrive_expect_tx(&tail->full[stage], A_STAGE_BYTES);
#pragma unroll
          for (int seg = 0; seg < 2; ++seg) {
            tma_load_2d(sb + seg * CHUNK_BYTES, &tmap, 0, (seg * chunks + c) * m_all + m0 + mt * TILE, &tail->full[stage]);
            tma_load_2d(sb + (2 + seg) * CHUNK_BYTES, &xmap, 0, (seg * chunks + c) * n + i0 + it * TILE, &tail->full[stage]);
          }
          if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, u = 0;
      for (long long t = blockIdx.x; t < total; t += gridDim.x, ++u) {
        const uint32_t buf = u & 1u, use = (u >> 1) & 1u;      // accumulator set, parity of its use count
        const uint32_t acc = tmem_base + buf * 256u;
        mbar_wait(&tail->tmem_empty[buf], use ^ 1u);
        tc_fence_after();
        for (int c = 0; c < chunks; ++c) {
          mbar_wait(&tail->full[stage], phase);
          tc_fence_after();
          const uint32_t sb = smem_u32(smem + (size_t)stage * A_STAGE_BYTES);
          const int slices = min(KC, dp - c * KC) / UMMA_K;
          const uint64_t dth = make_desc_sw64(sb), dtl = make_desc_sw64(sb + CHUNK_BYTES);
          const uint64_t dx = make_desc_sw64(sb + 2 * CHUNK_BYTES);          // [Xh | Xl]: 256 rows
          for (int k = 0; k < slices; ++k) {
            const uint32_t accum = (c | k) ? 1u : 0u;
            umma_tf32(acc, dth + (uint64_t)(k * 2), dx + (uint64_t)(k * 2), idesc_n(256), accum);   // Th.Xh | Th.Xl
            umma_tf32(acc, dtl + (uint64_t)(k * 2), dx + (uint64_t)(k * 2), idesc_n(128), 1u);      // += Tl.Xh
          }
          tc_commit(&tail->empty[stage]);
          if (++stage == A_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tail->tmem_full[buf]);
      }
    }
  } else {
    const int q = warp & 3, cgrp = (warp - 2) >> 2;
    const int row_in_tile = q * 32 + lane;
    const size_t seg_stride = (size_t)nic * m * KC;
    float* s_hi = reinterpret_cast<float*>(staging + (size_t)(warp - 2) * 2 * 32 * KC * 4);   // [32 rows][16]
    float* s_lo = s_hi + 32 * KC;
    uint32_t u = 0;
    for (long long t = blockIdx.x; t < total; t += gridDim.x, ++u) {
      const uint32_t buf = u & 1u, use = (u >> 1) & 1u;
      const int mt = (int)(t / IT), it = (int)(t % IT);
      const int gm = mt * TILE + row_in_tile;
      const int gm0 = mt * TILE + q * 32;                     // first row of this warp
      mbar_wait(&tail->tmem_full[buf], use);
      tc_fence_after();
      const uint32_t tl = tmem_base + buf * 256u + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int cb = cgrp * 32; cb < cgrp * 32 + 32; cb += KC) {
        uint32_t p0[16], p1[16];
        tmem_ld16(tl + cb, p0);                                // Th.Xh + Tl.Xh
        tmem_ld16(tl + TILE + cb, p1);                         // Th.Xl
        tmem_ld_wait();
        const int ic = (it * TILE + cb) / KC;                  // row chunk of these 16 columns
        if (ic >= nic) continue;                               // warp-uniform
        float hi[16], lo[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float w = -sigmoid_mufu(__uint_as_float(p1[e]) + __uint_as_float(p0[e]));
          hi[e] = tf32_rna(w);
          lo[e] = tf32_rna(w - hi[e]);
        }
        float* dst = W + ((size_t)ic * m + gm0) * KC;          // this warp's 32 rows x 16 floats: 2 KB contiguous
        if (gm0 + 32 <= m) {
          if (lane == 0) bulk_wait_read();                     // the previous block has left the staging buffer
          __syncwarp();
#pragma unroll
          for (int k = 0; k < 4; ++k) {                        // lane l, instruction k: chunk (k + l / 2) % 4 - no bank conflicts
            const int c = (k + (lane >> 1)) & 3;
            *reinterpret_cast<float4*>(s_hi + lane * KC + 4 * c) = chunk_of(hi, c);
            *reinterpret_cast<float4*>(s_lo + lane * KC + 4 * c) = chunk_of(lo, c);
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            bulk_store(dst, s_hi, 32 * KC * 4);
            bulk_store(dst + seg_stride, s_lo, 32 * KC * 4);
            bulk_commit();
          }
        } else if (gm < m) {                                   // ragged last row tile: per-lane stores
          float* d = dst + (size_t)lane * KC;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            *reinterpret_cast<float4*>(d + 4 * c) = make_float4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
            *reinterpret_cast<float4*>(d + seg_stride + 4 * c) = make_float4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->tmem_empty[buf]);
    }
    if (lane == 0) bulk_wait_read();                           // shared memory must outlive the bulk reads
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------ accumulate
// wmap: W as a [.., 16] matrix (rows (seg * nic + ic) * m + r); xtmap: XT as [.., 16] (rows (seg * nic_all + ic) * dn + col).
// CTA (mt, ks) sums the row chunks ic = ks, ks + ksplit, ... of this N-chunk and adds into G[m][P] (fp32 atomics).
__global__ void __launch_bounds__(THREADS, 1)
k2_accum_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap xtmap, int m, int P, int d, int dn,
                int nic, int nic_all, int ic0, int ksplit, float* __restrict__ G) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const uint32_t xt_bytes = (uint32_t)dn * KC * 4;
  const uint32_t stage_bytes = 2 * CHUNK_BYTES + 2 * xt_bytes;
  TailB* tail = reinterpret_cast<TailB*>(smem + (size_t)B_STAGES * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x / ksplit, ks = blockIdx.x % ksplit;
  const int n_iter = nic > ks ? (nic - ks + ksplit - 1) / ksplit : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < B_STAGES; ++s) { mbar_init(&tail->full[s], 1); mbar_init(&tail->empty[s], 1); }
    mbar_init(&tail->tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tail->tmem_base, 512); tmem_relinquish(); }
  if (warp == 0 && lane == 0) { tma_prefetch_desc(&wmap); tma_prefetch_desc(&xtmap); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int u = 0; u < n_iter; ++u) {
        const int ic = ks + u * ksplit;
        mbar_wait(&tail->empty[stage], phase ^ 1);
        uint8_t* sb = smem + (size_t)stage * stage_bytes;
        mbar_arrive_expect_tx(&tail->full[stage], stage_bytes);
#pragma unroll
        for (int seg = 0; seg < 2; ++seg) {
          tma_load_2d(sb + seg * CHUNK_BYTES, &wmap, 0, (seg * nic + ic) * m + mt * TILE, &tail->full[stage]);
          tma_load_2d(sb + 2 * CHUNK_BYTES + seg * xt_bytes, &xtmap, 0, (int)(((long long)seg * nic_all + ic0 + ic) * dn),
                      &tail->full[stage]);
        }
        if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int u = 0; u < n_iter; ++u) {
        mbar_wait(&tail->full[stage], phase);
        tc_fence_after();
        const uint32_t sb = smem_u32(smem + (size_t)stage * stage_bytes);
        const uint64_t dwh = make_desc_sw64(sb), dwl = make_desc_sw64(sb + CHUNK_BYTES);
        const uint64_t dxt = make_desc_sw64(sb + 2 * CHUNK_BYTES);           // [XTh | XTl]: 2 dn rows
#pragma unroll
        for (int k = 0; k < KC / UMMA_K; ++k) {
          const uint32_t accum = (u | k) ? 1u : 0u;
          umma_tf32(tmem_base, dwh + (uint64_t)(k * 2), dxt + (uint64_t)(k * 2), idesc_n(2 * dn), accum);        // Wh.XTh | Wh.XTl
          umma_tf32(tmem_base + 256, dwl + (uint64_t)(k * 2), dxt + (uint64_t)(k * 2), idesc_n(dn), accum);     // Wl.XTh
        }
        tc_commit(&tail->empty[stage]);
        if (++stage == B_STAGES) { stage = 0; phase ^= 1; }
      }
      tc_commit(&tail->tmem_full);
    }
  } else if (n_iter > 0) {
    const int q = warp & 3, cgrp = (warp - 2) >> 2;
    const int gm = mt * TILE + q * 32 + lane;
    mbar_wait(&tail->tmem_full, 0);
    tc_fence_after();
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int cb = cgrp * 32; cb < cgrp * 32 + 32 && cb < dn; cb += 8) {
      uint32_t p0[8], p1[8], p2[8];
      tmem_ld8(tl + cb, p0);
      tmem_ld8(tl + dn + cb, p1);
      tmem_ld8(tl + 256 + cb, p2);
      tmem_ld_wait();
      if (gm < m) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (cb + e < d)
            atomicAdd(G + (size_t)gm * P + cb + e, (__uint_as_float(p1[e]) + __uint_as_float(p2[e])) + __uint_as_float(p0[e]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace k2
}  // namespace sg
