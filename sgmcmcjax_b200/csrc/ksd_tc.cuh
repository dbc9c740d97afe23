// K3 (tensor-core variant): IMQ kernel Stein discrepancy as Gram contractions on tcgen05.
//
// ksd.py:7-37 with c = 1, beta = -1/2:   k0 = gg b^-1/2 + (gd + D) b^-3/2 - 3 dd b^-5/2,  b = 1 + dd
//   dd = |x_i - x_j|^2 = xx_i + xx_j - 2 x_i.x_j
//   gd = (g_i - g_j).(x_i - x_j) = gx_i + gx_j - (g_i.x_j + x_i.g_j)
//   gg = g_i.g_j
// k0 is symmetric in (i, j), so only tiles J >= I of the 128 x 128 tiling are computed and
// off-diagonal tiles count twice (imq_KSD sums all N^2 ordered pairs, ksd.py:58-66).
//
// Numerics.  dd and gd depend on differences only, so samples and grads are centred by their
// column means first (removes the catastrophic cancellation of xx_i + xx_j - 2 x_i.x_j for
// tightly clustered posterior samples); gg is reconstructed as gc_i.gc_j + r_i + r_j with
// r_i = mg.gc_i + |mg|^2 / 2.  Every dot product runs as 3xTF32: a = a_hi + a_lo with
// a_hi = rna_tf32(a), a_lo = rna_tf32(a - a_hi);  a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi
// (error ~2^-21 |a||b|, fp32 accumulation in TMEM).  The N^2 sum is accumulated in fp64.
//
// Operand matrix  Z[4][chunks][n][16]  fp32 (columns zero padded to whole 16-float chunks; chunk-major so that the
// 128-row x 16-float TMA box of one (segment, chunk) is 8 KB of contiguous global memory), segments
// 0 = xc_hi, 1 = xc_lo, 2 = gc_hi, 3 = gc_lo; row terms rt[n] = (xx, gx, r, |g|^2).
// Exact diagonal: for i == j the reference has x_i - x_j = 0 exactly, k0 = |g_i|^2 + D; the 3xTF32 residual
// (~1e-6 |x|^2 in dd) would otherwise be amplified by (gg/2 + 3D/2), so the diagonal uses dd = gd = 0, gg = |g_i|^2.
// Per 128 x 128 tile and K-chunk of 16 columns, TMA brings the 4 row-side and 4 column-side
// segment chunks into one 64 KB smem stage (64-byte swizzle, K-major); 12 segment products
// (3 for x.x, 3 for g.g, 6 for the cross term), issued as 6 N = 256 MMAs on adjacent column-side
// segments, accumulate into three 128-column fp32 accumulators in TMEM (the two cross products share one).  Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc),
// 2..17 = epilogue (tcgen05.ld -> IMQ formula -> fp64 partial sums; four warps per TMEM lane quarter, 32 columns each).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

namespace sg {
namespace ksdtc {

constexpr int TILE = 128;          // tile edge (rows I and rows J)
constexpr int KC = 16;             // K-chunk (floats) per smem stage = 64 bytes = swizzle span
constexpr int STAGES = 3;
constexpr int CHUNK_BYTES = TILE * KC * 4;           // 8 KB per (segment, side) chunk
constexpr int STAGE_BYTES = 8 * CHUNK_BYTES;         // 64 KB
constexpr int NUM_THREADS = 576;   // warp 0 TMA, warp 1 MMA, warps 2..17 epilogue (4 per TMEM lane quarter)
constexpr int TMEM_COLS = 512;     // 3 accumulators x 128 columns (allocations are powers of two)
constexpr int UMMA_K = 8;          // tf32: 32 bytes of K per instruction

struct SmemTail {                  // after the stage buffers
  uint64_t full[STAGES], empty[STAGES], tmem_full, tmem_empty;
  uint32_t tmem_base;
  uint32_t pad;
  float4 colterm[2][TILE];
};
constexpr size_t SMEM_BYTES = 1024 /*alignment slack*/ + (size_t)STAGES * STAGE_BYTES + sizeof(SmemTail);

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{ .reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1; }" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) {   // ~4 s at 2 GHz
      printf("ksd_tc: mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
// the same box delivered to the same shared-memory offsets (and signalled on the same mbarrier offset) of every CTA in
// cta_mask of this cluster
__device__ __forceinline__ void tma_load_2d_multicast(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols));
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one arrival on the mbarrier at this offset in EVERY CTA of cta_mask, when the MMAs issued so far have retired
__device__ __forceinline__ void tc_commit_multicast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate, M = 128, N = 128, K = 8.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p; }"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 64-byte-swizzled operand tile of TILE rows x 16 floats: row pitch 64 B, 8-row groups 512 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64))
__device__ __forceinline__ uint64_t make_desc_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;                    // leading byte offset: unused for swizzled K-major
  d |= (uint64_t)(512 >> 4) << 32;           // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)4 << 61;                    // SWIZZLE_64B
  return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 [4,6)=1, a/b format TF32 [7,10)/[10,13)=2, K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t IDESC_TF32_128x128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t IDESC_TF32_128x256 = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------ prep
// Column sums of X and G in double: sums[0][d], sums[1][d].
// (the kernels and the product table are `static`: mlp_grad.cuh reuses this header's tcgen05 / mbarrier wrappers from other
// translation units)
static __global__ void colsum_kernel(const float* __restrict__ X, const float* __restrict__ G, int n, int d, double* sums) {
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    double sx = 0.0, sgr = 0.0;
    for (int r = blockIdx.x; r < n; r += gridDim.x) {
      sx += (double)X[(size_t)r * d + c];
      sgr += (double)G[(size_t)r * d + c];
    }
    atomicAdd(&sums[c], sx);
    atomicAdd(&sums[d + c], sgr);
  }
}

// Same sums for rows of whole float4s (d % 4 == 0, 16-byte aligned tables): thread = (row lane, float4 column), four rows
// in flight per thread and array - the scalar kernel above keeps one load in flight per thread and runs at ~1 TB/s.
static __global__ void colsum4_kernel(const float4* __restrict__ X, const float4* __restrict__ G, int n, int d4, double* sums) {
  extern __shared__ double cs_red[];                  // [rows_per_iter][d4][8]
  const int rpi = blockDim.x / d4;                    // rows per iteration of this block
  const int q = threadIdx.x % d4, rl = threadIdx.x / d4;
  double ax[4] = {0, 0, 0, 0}, ag[4] = {0, 0, 0, 0};
  if (rl < rpi) {
    const long long stride = (long long)gridDim.x * rpi;
    long long r = (long long)blockIdx.x * rpi + rl;
    for (; r + 3 * stride < n; r += 4 * stride) {
      float4 x[4], g[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { x[u] = X[(r + u * stride) * d4 + q]; g[u] = G[(r + u * stride) * d4 + q]; }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ax[0] += x[u].x; ax[1] += x[u].y; ax[2] += x[u].z; ax[3] += x[u].w;
        ag[0] += g[u].x; ag[1] += g[u].y; ag[2] += g[u].z; ag[3] += g[u].w;
      }
    }
    for (; r < n; r += stride) {
      const float4 x = X[r * d4 + q], g = G[r * d4 + q];
      ax[0] += x.x; ax[1] += x.y; ax[2] += x.z; ax[3] += x.w;
      ag[0] += g.x; ag[1] += g.y; ag[2] += g.z; ag[3] += g.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { cs_red[((size_t)rl * d4 + q) * 8 + k] = ax[k]; cs_red[((size_t)rl * d4 + q) * 8 + 4 + k] = ag[k]; }
  }
  __syncthreads();
  const int d = 4 * d4;
  for (int t = threadIdx.x; t < 2 * d; t += blockDim.x) {      // t < d: X column t, else G column t - d
    const int c = t < d ? t : t - d, half = t < d ? 0 : 4;
    double v = 0.0;
    for (int j = 0; j < rpi; ++j) v += cs_red[((size_t)j * d4 + (c >> 2)) * 8 + half + (c & 3)];
    atomicAdd(&sums[t], v);
  }
}

// One warp per row: centre, split into tf32 hi / lo, row terms.
static __global__ void split_kernel(const float* __restrict__ X, const float* __restrict__ G, int n, int d, int dp,
                             const double* __restrict__ sums, float* __restrict__ Z, float4* __restrict__ rt) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const double inv_n = 1.0 / (double)n;
  float xx = 0.f, gx = 0.f, r = 0.f, mm = 0.f, g2 = 0.f;
  const int n_chunks = (dp + KC - 1) / KC;
  const int dpc = n_chunks * KC;                 // columns incl. the zero pad of the last chunk
  for (int c = lane; c < dpc; c += 32) {
    float xh = 0.f, xl = 0.f, gh = 0.f, gl = 0.f;
    if (c < d) {
      const float mx = (float)(sums[c] * inv_n), mg = (float)(sums[d + c] * inv_n);
      const float xc = X[(size_t)row * d + c] - mx;
      const float graw = G[(size_t)row * d + c];
      const float gc = graw - mg;
      g2 = fmaf(graw, graw, g2);
      xh = tf32_rna(xc); xl = tf32_rna(xc - xh);
      gh = tf32_rna(gc); gl = tf32_rna(gc - gh);
      xx = fmaf(xc, xc, xx); gx = fmaf(gc, xc, gx); r = fmaf(mg, gc, r); mm = fmaf(mg, mg, mm);
    }
    // chunk-major layout: every (segment, chunk) is one contiguous [n][KC] block, so a 128-row TMA box is 8 KB contiguous
    const size_t blk = (size_t)n * KC;
    float* z = Z + ((size_t)(c / KC) * n + row) * KC + (c % KC);
    z[0] = xh; z[(size_t)n_chunks * blk] = xl; z[2 * (size_t)n_chunks * blk] = gh; z[3 * (size_t)n_chunks * blk] = gl;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    xx += __shfl_xor_sync(0xffffffffu, xx, o); gx += __shfl_xor_sync(0xffffffffu, gx, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);   mm += __shfl_xor_sync(0xffffffffu, mm, o);
    g2 += __shfl_xor_sync(0xffffffffu, g2, o);
  }
  if (lane == 0) rt[row] = make_float4(xx, gx, r + 0.5f * mm, g2);
}

// ------------------------------------------------------------------------------------------ tiles
// Upper-triangular tile list, row-major: t -> (I, J >= I) with T tiles per side.
__device__ __forceinline__ void tile_of(long long t, int T, int& I, int& J) {
  // row I starts at S(I) = I*T - I*(I-1)/2
  const double tt = (double)T + 0.5;
  int i = (int)(tt - sqrt(tt * tt - 2.0 * (double)t));
  if (i < 0) i = 0;
  if (i > T - 1) i = T - 1;
  auto start = [T](int ii) { return (long long)ii * T - (long long)ii * (ii - 1) / 2; };
  while (i > 0 && start(i) > t) --i;
  while (i < T - 1 && start(i + 1) <= t) ++i;
  I = i;
  J = i + (int)(t - start(i));
}

// Cluster items: the two CTAs of a cluster take the tiles (I, J) and (I, J + 1) of one tile ROW at a time, so that the
// row-side operand is fetched once and multicast to both.  Row I has ceil((T - I) / 2) items; item g -> (I, first J).
__host__ __device__ __forceinline__ long long pair_start(long long I, long long T) {   // items in rows 0 .. I-1
  const long long m = T + 1;                                     // sum_{i<I} floor((m - i) / 2)
  const long long odd = (m & 1) ? (I + 1) >> 1 : I >> 1;          // how many of the (m - i) are odd
  return (I * m - I * (I - 1) / 2 - odd) / 2;
}
__device__ __forceinline__ void pair_of(long long g, int T, int& I, int& J) {
  const double tt = (double)T + 1.0;
  int i = (int)(tt - sqrt(fmax(tt * tt - 4.0 * (double)g, 0.0)));
  if (i < 0) i = 0;
  if (i > T - 1) i = T - 1;
  while (i > 0 && pair_start(i, T) > g) --i;
  while (i < T - 1 && pair_start(i + 1, T) <= g) ++i;
  I = i;
  J = i + 2 * (int)(g - pair_start(i, T));
}

__device__ __forceinline__ float imq_k0_tc(float dd, float gg, float gd, float dim) {
  const float b = 1.0f + dd;
  float r;                                      // b >= 1: the denormal guard of rsqrtf() (3 more instructions) is dead weight
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  const float r2 = r * r;
  const float r3 = r2 * r;
  return gg * r + (gd + dim) * r3 - 3.0f * dd * r3 * r2;
}

// Paired products.  The column-side chunks of a stage are laid out [xh_J | gh_J | xl_J | gl_J], so that one N = 256
// MMA multiplies a row-side segment with TWO adjacent column-side segments and writes two adjacent TMEM
// accumulators: half the MMA instructions, and the A operand is read from shared memory once per 256 columns
// (12 KB of operand traffic per 128 tensor cycles instead of 8 KB per 64, the 128 B/clk shared-memory port).
//   TMEM columns: [0,128) x_I.x_J   [128,256) x_I.g_J + g_I.x_J   [256,384) g_I.g_J
// Only the SUM of the two cross products enters the formula, so the g_I pair lands one accumulator to the right of the x_I
// pair and both add into the middle one: three accumulators for the epilogue to read instead of four (TMEM reads, at
// 64 B/clk, are the bulk of the epilogue; round 2).  The one MMA per tile that would have to overwrite [256,384) while
// adding into [128,256) is issued as two N = 128 halves (see the issuer).
// table: (TMEM column of the pair, row-side segment 0 xh 1 xl 2 gh 3 gl, column-side pair 0 = [xh|gh], 1 = [xl|gl])
static __constant__ int16_t PROD[6][3] = {
    {0, 0, 0}, {0, 0, 1}, {0, 1, 0},          // x_I . [x_J | g_J]  (3xTF32: hi.hi, hi.lo, lo.hi)
    {128, 2, 0}, {128, 2, 1}, {128, 3, 0}     // g_I . [x_J | g_J]
};
__device__ __forceinline__ int col_slot(int seg) { return ((seg & 1) << 1) | (seg >> 1); }   // xh 0, gh 1, xl 2, gl 3

// CL = 1 (default): one CTA per tile.  CL = 2 (opt-in, launched as clusters of two CTAs): the operand stream from L2 is
// large (458 KB per 128 x 128 tile: ncu 35.4 GB per N = 50k call at 7.3 TB/s, tensor pipe 56 % busy), so the two CTAs of a
// cluster work on neighbouring tiles of one tile row in lockstep and share its row-side chunks: each CTA fetches two of the
// four row-side segments and TMA-multicasts them into both CTAs' stage (344 KB per tile).  A stage is refilled only when
// BOTH CTAs have consumed it: the MMA issuers commit to the `empty` barrier of both CTAs (arrival count 2).  Measured: same
// result, 4.91 vs 4.84 ms - the stream was not the limit.
template <int CL>
static __global__ void __launch_bounds__(NUM_THREADS, 1)
ksd_tc_kernel(const __grid_constant__ CUtensorMap zmap, const float4* __restrict__ rt, int n, int d, int dp,
              int part, int nparts, double* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  SmemTail* tail = reinterpret_cast<SmemTail*>(smem + (size_t)STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int T = (n + TILE - 1) / TILE;
  const long long total = CL == 1 ? (long long)T * (T + 1) / 2 : pair_start(T, T);
  // items (tiles / tile pairs) of this partition: g = part + nparts * u, u = 0 .. n_local-1; this CTA (cluster) takes
  // u = first + k * stride
  const long long n_local = total > part ? (total - part + nparts - 1) / nparts : 0;
  const int n_chunks = (dp + KC - 1) / KC;
  const uint32_t crank = CL == 1 ? 0u : cluster_ctarank();
  const long long first = blockIdx.x / CL, stride = gridDim.x / CL;
  // item g -> this CTA's tile (I, J); `live` false: the odd tile out at the end of a row (a weight-0 repeat of its neighbour)
  auto item = [&](long long g, int& I, int& J, bool& live) {
    if (CL == 1) { tile_of(g, T, I, J); live = true; return; }
    pair_of(g, T, I, J);
    live = J + (int)crank < T;
    if (live) J += (int)crank;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&tail->full[s], 1); mbar_init(&tail->empty[s], CL); }
    mbar_init(&tail->tmem_full, 1);
    mbar_init(&tail->tmem_empty, 16);      // one arrival per epilogue warp
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tail->tmem_base, TMEM_COLS); tmem_relinquish(); }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&zmap);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // the peer's barriers are initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long u = first; u < n_local; u += stride) {
        int I, J;
        bool live;
        item(part + (long long)nparts * u, I, J, live);
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(&tail->empty[stage], phase ^ 1);
          uint8_t* sbase = smem + (size_t)stage * STAGE_BYTES;
          mbar_arrive_expect_tx(&tail->full[stage], STAGE_BYTES);
#pragma unroll
          for (int seg = 0; seg < 4; ++seg) {
            const int blk_row = (seg * n_chunks + c) * n;          // first row of block (segment, chunk) in the [.., KC] view
            if (CL == 1) tma_load_2d(sbase + seg * CHUNK_BYTES, &zmap, 0, blk_row + I * TILE, &tail->full[stage]);
            else if ((uint32_t)(seg >> 1) == crank)
              tma_load_2d_multicast(sbase + seg * CHUNK_BYTES, &zmap, 0, blk_row + I * TILE, &tail->full[stage], (uint16_t)3);
            tma_load_2d(sbase + (4 + col_slot(seg)) * CHUNK_BYTES, &zmap, 0, blk_row + J * TILE, &tail->full[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one lane) =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc_phase = 0;
      for (long long u = first; u < n_local; u += stride) {
        mbar_wait(&tail->tmem_empty, acc_phase ^ 1);
        tc_fence_after();
        uint32_t started = 0;                                  // bit 0 / 1: the x_I / g_I pair has written its accumulators
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(&tail->full[stage], phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const int valid = min(KC, dp - c * KC);              // dp is a multiple of 8
          const int slices = valid / UMMA_K;
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            const int a = p >= 3;                                  // 0: x_I pair, columns [0,256); 1: g_I pair, columns [128,384)
            const uint64_t da = make_desc_sw64(sbase + PROD[p][1] * CHUNK_BYTES);
            const uint64_t db = make_desc_sw64(sbase + (4 + 2 * PROD[p][2]) * CHUNK_BYTES);
            for (int k = 0; k < slices; ++k) {
              // advancing K inside the swizzle atom = +32 bytes on the start address (>>4 => +2)
              if (a == 1 && !(started & 2u)) {
                // first g_I product of the tile: ADD g_I.x_J into the cross accumulator (the x_I pair, p < 3, has written
                // it), OVERWRITE the g_I.g_J accumulator - two N = 128 MMAs on the two halves of the column-side pair
                const uint64_t db2 = make_desc_sw64(sbase + (4 + 2 * PROD[p][2] + 1) * CHUNK_BYTES);
                umma_tf32(tmem_base + 128, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC_TF32_128x128, 1u);
                umma_tf32(tmem_base + 256, da + (uint64_t)(k * 2), db2 + (uint64_t)(k * 2), IDESC_TF32_128x128, 0u);
              } else {
                umma_tf32(tmem_base + PROD[p][0], da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC_TF32_128x256,
                          a == 1 ? 1u : (started & 1u));
              }
              started |= 1u << a;
            }
          }
          if (CL == 1) tc_commit(&tail->empty[stage]);         // frees the smem stage when these MMAs retire
          else tc_commit_multicast(&tail->empty[stage], (uint16_t)3);   // ... in both CTAs: either may refill both
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tail->tmem_full);                           // accumulators complete
        acc_phase ^= 1;
      }
    }
  } else {
    // ===== epilogue: warps 2..17; warp w reads TMEM lanes 32*(w%4) .. +31, columns 32*((w-2)/4) .. +31 =====
    const int q = warp & 3;
    const int cgrp = (warp - 2) >> 2;                          // 0..3
    const int row_in_tile = q * 32 + lane;
    const int et = threadIdx.x - 64;                           // 0..511
    const float dim = (float)d;
    double acc = 0.0;
    uint32_t acc_phase = 0;
    int buf = 0;
    for (long long u = first; u < n_local; u += stride) {
      int I, J;
      bool live;
      item(part + (long long)nparts * u, I, J, live);
      const int gi = I * TILE + row_in_tile;
      const int gjc = J * TILE + et;
      if (et < TILE) tail->colterm[buf][et] = gjc < n ? rt[gjc] : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4 ri = gi < n ? rt[gi] : make_float4(0.f, 0.f, 0.f, 0.f);
      asm volatile("bar.sync 1, 512;" ::: "memory");          // colterm[buf] visible to the 16 epilogue warps
      mbar_wait(&tail->tmem_full, acc_phase);
      tc_fence_after();
      float tsum = 0.f;
      const uint32_t colterm_s = smem_u32(&tail->colterm[buf][0]);   // read with ld.shared (the generic path is slower)
      auto colterm = [&](int j) {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(colterm_s + 16u * (uint32_t)j));
        return v;
      };
      const bool diag_tile = I == J;
      const int jmax = min(TILE, n - J * TILE);
      const bool interior = !diag_tile && (I + 1) * TILE <= n && jmax == TILE;   // no masks, no exact diagonal
      const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int cb = cgrp * 32; cb < cgrp * 32 + 32; cb += 8) {
        uint32_t p1[8], p2[8], p3[8];
        tmem_ld8(tl + cb, p1);                  // x_I.x_J
        tmem_ld8(tl + 2 * TILE + cb, p2);       // g_I.g_J
        tmem_ld8(tl + TILE + cb, p3);           // x_I.g_J + g_I.x_J
        tmem_ld_wait();
        if (interior) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 cj = colterm(cb + e);
            const float dd = fmaf(-2.0f, __uint_as_float(p1[e]), ri.x + cj.x);
            const float gg = __uint_as_float(p2[e]) + (ri.z + cj.z);
            const float gd = (ri.y + cj.y) - __uint_as_float(p3[e]);
            tsum += imq_k0_tc(dd, gg, gd, dim);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 cj = colterm(cb + e);
            float dd = fmaf(-2.0f, __uint_as_float(p1[e]), ri.x + cj.x);
            float gg = __uint_as_float(p2[e]) + (ri.z + cj.z);
            float gd = (ri.y + cj.y) - __uint_as_float(p3[e]);
            if (diag_tile && cb + e == row_in_tile) { dd = 0.f; gd = 0.f; gg = ri.w; }   // i == j: exact
            const float k0 = imq_k0_tc(dd, gg, gd, dim);
            tsum += (gi < n && cb + e < jmax) ? k0 : 0.f;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->tmem_empty);           // accumulators may be overwritten
      acc += (double)tsum * (!live ? 0.0 : I == J ? 1.0 : 2.0);
      acc_phase ^= 1;
      buf ^= 1;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && acc != 0.0) atomicAdd(out, acc);
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();          // no CTA leaves while its peer may still signal its barriers
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------ host
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int padded_dim(int d) { return (d + 7) & ~7; }
inline size_t z_floats(int n, int d) { return (size_t)4 * ((padded_dim(d) + KC - 1) / KC) * KC * (size_t)n; }
// workspace: Z [4 segments][chunks][n][KC] floats | row terms [n] float4 | column sums [2 d] doubles
inline size_t workspace_bytes(int n, int d) {
  return align_up(z_floats(n, d) * sizeof(float), 256) + align_up((size_t)n * sizeof(float4), 256) +
         align_up((size_t)2 * d * sizeof(double), 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

}  // namespace ksdtc
}  // namespace sg
