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
// Row terms inside the contraction.  The operands carry four extra columns (the zero pad of the last K-chunk, so no
// extra MMA work for D = 100) that differ between the row side (I) and the column side (J):
//     x rows: [-xx_i/2, 1, -gx_i, 0]   x cols: [1, -xx_j/2, 0, -gx_j]      g rows: [0, 0, r_i, 1]   g cols: [0, 0, 1, r_j]
// so that the tensor core itself delivers
//     P1 = x_I.x_J - (xx_i + xx_j)/2 = -dd/2,    P23 = x_I.g_J + g_I.x_J - gx_i - gx_j = -gd,    P4 = g_I.g_J + r_i + r_j = gg
// (x.g and g.x accumulate into ONE accumulator: three TMEM accumulators instead of four), and the epilogue is nine
// instructions per pair with no per-column shared-memory term:  b = 1 - 2 P1,  k0 = P4 b^-1/2 + (D - P23 + 6 P1 / b) b^-3/2.
// The 3xTF32 split carries the extra columns at fp32 accuracy (a "1" is hi = 1, lo = 0).
//
// Operand matrix  Z[4][chunks][n][16]  fp32, followed by the COLUMN-side variant of the last chunk  Zc[4][n][16] (columns zero padded to whole 16-float chunks; chunk-major so that the
// 128-row x 16-float TMA box of one (segment, chunk) is 8 KB of contiguous global memory), segments
// 0 = xc_hi, 1 = xc_lo, 2 = gc_hi, 3 = gc_lo; row terms rt[n] = (xx, gx, r, |g|^2).
// Exact diagonal: for i == j the reference has x_i - x_j = 0 exactly, k0 = |g_i|^2 + D; the 3xTF32 residual
// (~1e-6 |x|^2 in dd) would otherwise be amplified by (gg/2 + 3D/2), so the diagonal uses dd = gd = 0, gg = |g_i|^2.
// Per 128 x 128 tile and K-chunk of 16 columns, TMA brings the 4 row-side and 4 column-side
// segment chunks into one 64 KB smem stage (64-byte swizzle, K-major); 12 segment products
// (3 for x.x, 3 for g.g, 6 for the cross term), issued as 6 N = 256 MMAs on adjacent column-side
// segments, accumulate into four 128-column fp32 accumulators in TMEM.  Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc),
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
constexpr int TMEM_COLS = 512;     // 4 accumulators x 128 columns
constexpr int UMMA_K = 8;          // tf32: 32 bytes of K per instruction

struct SmemTail {                  // after the stage buffers
  uint64_t full[STAGES], empty[STAGES], tmem_full, tmem_empty;
  uint32_t tmem_base;
  uint32_t pad;
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

// One warp per row: centre, split into tf32 hi / lo, row terms; the four row-term columns dp-4 .. dp-1 (see the header) in
// the row-side variant (inside Z) and, for the whole last chunk, the column-side variant Zc.
static __global__ void split_kernel(const float* __restrict__ X, const float* __restrict__ G, int n, int d, int dp,
                             const double* __restrict__ sums, float* __restrict__ Z, float4* __restrict__ rt) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const double inv_n = 1.0 / (double)n;
  float xx = 0.f, gx = 0.f, r = 0.f, mm = 0.f, g2 = 0.f;
  const int n_chunks = (dp + KC - 1) / KC;
  const int dpc = n_chunks * KC;                 // columns incl. the zero pad of the last chunk
  const size_t blk = (size_t)n * KC;             // one (segment, chunk) block: [n][KC], a 128-row TMA box is 8 KB contiguous
  const size_t seg_stride = (size_t)n_chunks * blk;
  float* Zc = Z + 4 * seg_stride;                // column-side variant of the last chunk: [4 segments][n][KC]
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
    if (c >= dp - 4 && c < dp) continue;         // row-term columns: written below, once the sums are complete
    float* z = Z + ((size_t)(c / KC) * n + row) * KC + (c % KC);
    z[0] = xh; z[seg_stride] = xl; z[2 * seg_stride] = gh; z[3 * seg_stride] = gl;
    if (c / KC == n_chunks - 1) {
      float* zc = Zc + (size_t)row * KC + (c % KC);
      zc[0] = xh; zc[blk] = xl; zc[2 * blk] = gh; zc[3 * blk] = gl;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    xx += __shfl_xor_sync(0xffffffffu, xx, o); gx += __shfl_xor_sync(0xffffffffu, gx, o);
    r += __shfl_xor_sync(0xffffffffu, r, o);   mm += __shfl_xor_sync(0xffffffffu, mm, o);
    g2 += __shfl_xor_sync(0xffffffffu, g2, o);
  }
  const float rr = r + 0.5f * mm;
  if (lane == 0) rt[row] = make_float4(xx, gx, rr, g2);
  if (lane < 4) {
    // column dp-4+lane of:   x rows [-xx/2, 1, -gx, 0]   x cols [1, -xx/2, 0, -gx]   g rows [0, 0, rr, 1]   g cols [0, 0, 1, rr]
    const float xr = lane == 0 ? -0.5f * xx : lane == 1 ? 1.0f : lane == 2 ? -gx : 0.0f;
    const float xcv = lane == 0 ? 1.0f : lane == 1 ? -0.5f * xx : lane == 2 ? 0.0f : -gx;
    const float gr = lane == 2 ? rr : lane == 3 ? 1.0f : 0.0f;
    const float gcv = lane == 2 ? 1.0f : lane == 3 ? rr : 0.0f;
    const int c = dp - 4 + lane;
    float* z = Z + ((size_t)(c / KC) * n + row) * KC + (c % KC);
    float* zc = Zc + (size_t)row * KC + (c % KC);
    float h;
    h = tf32_rna(xr);  z[0] = h;               z[seg_stride] = tf32_rna(xr - h);
    h = tf32_rna(gr);  z[2 * seg_stride] = h;  z[3 * seg_stride] = tf32_rna(gr - h);
    h = tf32_rna(xcv); zc[0] = h;              zc[blk] = tf32_rna(xcv - h);
    h = tf32_rna(gcv); zc[2 * blk] = h;        zc[3 * blk] = tf32_rna(gcv - h);
  }
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

__device__ __forceinline__ float imq_k0_tc(float dd, float gg, float gd, float dim) {
  const float b = 1.0f + dd;
  float r;                                      // b >= 1: the denormal guard of rsqrtf() (3 more instructions) is dead weight
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  const float r2 = r * r;
  const float r3 = r2 * r;
  return gg * r + (gd + dim) * r3 - 3.0f * dd * r3 * r2;
}

// k0 from the augmented products: p1 = -dd/2, p23 = -gd, gg
__device__ __forceinline__ float imq_k0_aug(float p1, float p23, float gg, float dim) {
  const float b = fmaf(-2.0f, p1, 1.0f);
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  const float r2 = r * r;
  const float r3 = r2 * r;
  const float u = fmaf(6.0f * p1, r2, dim - p23);          // (gd + D) - 3 dd / b
  return fmaf(gg, r, u * r3);
}

// Paired products.  The column-side chunks of a stage are laid out [xh_J | gh_J | xl_J | gl_J], so that one N = 256
// MMA multiplies a row-side segment with TWO adjacent column-side segments and writes two adjacent TMEM
// accumulators: half the MMA instructions, and the A operand is read from shared memory once per 256 columns
// (12 KB of operand traffic per 128 tensor cycles instead of 8 KB per 64, the 128 B/clk shared-memory port).
//   TMEM columns: [0,128) x_I.x_J   [128,256) x_I.g_J + g_I.x_J   [256,384) g_I.g_J     (384 of the 512 allocated)
// The x_I pairs write columns [0,256), the g_I pairs columns [128,384): both accumulate the cross term in the middle.
// table: (TMEM column of the pair, row-side segment 0 xh 1 xl 2 gh 3 gl, column-side pair 0 = [xh|gh], 1 = [xl|gl])
static __constant__ int16_t PROD[6][3] = {
    {0, 0, 0}, {0, 0, 1}, {0, 1, 0},          // x_I . [x_J | g_J]  (3xTF32: hi.hi, hi.lo, lo.hi)
    {128, 2, 0}, {128, 2, 1}, {128, 3, 0}     // g_I . [x_J | g_J]
};
__device__ __forceinline__ int col_slot(int seg) { return ((seg & 1) << 1) | (seg >> 1); }   // xh 0, gh 1, xl 2, gl 3

static __global__ void __launch_bounds__(NUM_THREADS, 1)
ksd_tc_kernel(const __grid_constant__ CUtensorMap zmap, const float4* __restrict__ rt, int n, int d, int dp,
              int part, int nparts, double* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  SmemTail* tail = reinterpret_cast<SmemTail*>(smem + (size_t)STAGES * STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int T = (n + TILE - 1) / TILE;
  const long long total = (long long)T * (T + 1) / 2;
  // tiles of this partition: g = part + nparts * u, u = 0 .. n_local-1; this CTA takes u = blockIdx.x + k * gridDim.x
  const long long n_local = total > part ? (total - part + nparts - 1) / nparts : 0;
  const int n_chunks = (dp + KC - 1) / KC;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&tail->full[s], 1); mbar_init(&tail->empty[s], 1); }
    mbar_init(&tail->tmem_full, 1);
    mbar_init(&tail->tmem_empty, 16);      // one arrival per epilogue warp
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(&tail->tmem_base, TMEM_COLS); tmem_relinquish(); }
  if (warp == 0 && lane == 0) tma_prefetch_desc(&zmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tail->tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long long u = blockIdx.x; u < n_local; u += gridDim.x) {
        int I, J;
        tile_of(part + (long long)nparts * u, T, I, J);
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(&tail->empty[stage], phase ^ 1);
          uint8_t* sbase = smem + (size_t)stage * STAGE_BYTES;
          mbar_arrive_expect_tx(&tail->full[stage], STAGE_BYTES);
#pragma unroll
          for (int seg = 0; seg < 4; ++seg) {
            const int blk_row = (seg * n_chunks + c) * n;          // first row of block (segment, chunk) in the [.., KC] view
            // the last chunk holds the row-term columns, which differ on the column side: variant block Zc[seg]
            const int col_row = c == n_chunks - 1 ? (4 * n_chunks + seg) * n : blk_row;
            tma_load_2d(sbase + seg * CHUNK_BYTES, &zmap, 0, blk_row + I * TILE, &tail->full[stage]);
            tma_load_2d(sbase + (4 + col_slot(seg)) * CHUNK_BYTES, &zmap, 0, col_row + J * TILE, &tail->full[stage]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one lane) =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc_phase = 0;
      for (long long u = blockIdx.x; u < n_local; u += gridDim.x) {
        mbar_wait(&tail->tmem_empty, acc_phase ^ 1);
        tc_fence_after();
        uint32_t started = 0;                                  // bit a set once pair a (0: x_I, 1: g_I) has been issued
        for (int c = 0; c < n_chunks; ++c) {
          mbar_wait(&tail->full[stage], phase);
          tc_fence_after();
          const uint32_t sbase = smem_u32(smem + (size_t)stage * STAGE_BYTES);
          const int valid = min(KC, dp - c * KC);              // dp is a multiple of 8
          const int slices = valid / UMMA_K;
#pragma unroll
          for (int p = 0; p < 6; ++p) {
            const int a = PROD[p][0] >> 7;                         // 0: columns [0,256), 1: columns [128,384)
            const uint64_t da = make_desc_sw64(sbase + PROD[p][1] * CHUNK_BYTES);
            const uint64_t db = make_desc_sw64(sbase + (4 + 2 * PROD[p][2]) * CHUNK_BYTES);
            for (int k = 0; k < slices; ++k) {
              // advancing K inside the swizzle atom = +32 bytes on the start address (>>4 => +2)
              if (a == 1 && !(started & 2u)) {
                // first g_I MMA of the tile: ADD to the cross accumulator the x_I pair has started, START g.g -
                // one instruction has one accumulate flag, so this slice is issued as two N = 128 halves
                umma_tf32(tmem_base + 128, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC_TF32_128x128, 1u);
                umma_tf32(tmem_base + 256, da + (uint64_t)(k * 2), db + (uint64_t)((CHUNK_BYTES >> 4) + k * 2),
                          IDESC_TF32_128x128, 0u);
              } else {
                umma_tf32(tmem_base + PROD[p][0], da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), IDESC_TF32_128x256,
                          (started >> a) & 1u);
              }
              started |= 1u << a;
            }
          }
          tc_commit(&tail->empty[stage]);                      // frees the smem stage when these MMAs retire
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
    const float dim = (float)d;
    double acc = 0.0;
    uint32_t acc_phase = 0;
    for (long long u = blockIdx.x; u < n_local; u += gridDim.x) {
      int I, J;
      tile_of(part + (long long)nparts * u, T, I, J);
      const int gi = I * TILE + row_in_tile;
      const bool diag_tile = I == J;
      const float g2i = (diag_tile && gi < n) ? rt[gi].w : 0.f;   // |g_i|^2: the exact i == j pair
      mbar_wait(&tail->tmem_full, acc_phase);
      tc_fence_after();
      float tsum = 0.f;
      const int jmax = min(TILE, n - J * TILE);
      const bool interior = !diag_tile && (I + 1) * TILE <= n && jmax == TILE;   // no masks, no exact diagonal
      const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int cb = cgrp * 32; cb < cgrp * 32 + 32; cb += 8) {
        uint32_t p1[8], p2[8], p3[8];
        tmem_ld8(tl + cb, p1);                  // x_I.x_J - (xx_i + xx_j) / 2
        tmem_ld8(tl + TILE + cb, p2);           // x_I.g_J + g_I.x_J - gx_i - gx_j
        tmem_ld8(tl + 2 * TILE + cb, p3);       // g_I.g_J + r_i + r_j
        tmem_ld_wait();
        if (interior) {
#pragma unroll
          for (int e = 0; e < 8; ++e)
            tsum += imq_k0_aug(__uint_as_float(p1[e]), __uint_as_float(p2[e]), __uint_as_float(p3[e]), dim);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float a1 = __uint_as_float(p1[e]), a2 = __uint_as_float(p2[e]), a3 = __uint_as_float(p3[e]);
            if (diag_tile && cb + e == row_in_tile) { a1 = 0.f; a2 = 0.f; a3 = g2i; }   // i == j: dd = gd = 0, gg = |g_i|^2
            const float k0 = imq_k0_aug(a1, a2, a3, dim);
            tsum += (gi < n && cb + e < jmax) ? k0 : 0.f;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tail->tmem_empty);           // accumulators may be overwritten
      acc += (double)tsum * (I == J ? 1.0 : 2.0);
      acc_phase ^= 1;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && acc != 0.0) atomicAdd(out, acc);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------ host
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int padded_dim(int d) { return (d + 7) & ~7; }
inline int ksd_padded_dim(int d) { return (d + 4 + 7) & ~7; }      // + the four row-term columns (dp-4 .. dp-1)
// Z [4 segments][chunks][n][KC] and the column-side variant of the last chunk Zc [4 segments][n][KC]
inline size_t z_floats(int n, int d) { return (size_t)4 * ((ksd_padded_dim(d) + KC - 1) / KC + 1) * KC * (size_t)n; }
// workspace: Z | Zc floats | row terms [n] float4 | column sums [2 d] doubles
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
