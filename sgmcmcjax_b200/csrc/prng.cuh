// Device-side jax 0.4.13 PRNG (threefry2x32, non-partitionable counter layout).
//
// Semantics reproduced (bit-exact for the integer parts):
//   random.split   - samplers.py:49,53  kernels.py:55,114,120  diffusion_util.py:66,91,112
//   random.choice  - gradient_estimation.py:32,75,140  (== randint(key,(B,),0,N))
//   random.normal  - diffusions.py:67,104,148,189,198,232,273,283,332
// jax lays a counter array of n words out as two halves: element j < h is word 0 of
// block (j, j+h), element j >= h is word 1 of block (j-h, j); h = ceil(n/2) and an odd
// n is padded with ONE ZERO counter (not with n).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sg {

// Optional phase timers (-DSG_PHASE_TIMING): cycles of block 0 / thread 0 per phase, read back with sgmcmc_debug_phase_cycles.
#ifdef SG_PHASE_TIMING
static __device__ unsigned long long g_phase_cycles[32];   // per translation unit: read where the kernel is instantiated
#define SG_T0() const long long sg_t0__ = clock64()
#define SG_T1(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_phase_cycles[slot], (unsigned long long)(clock64() - sg_t0__)); } while (0)
#define SG_TA(var) long long var = clock64()
#define SG_TBT(slot, var, t) do { if (blockIdx.x == 0 && threadIdx.x == (t)) { const long long n__ = clock64(); atomicAdd(&g_phase_cycles[slot], (unsigned long long)(n__ - var)); var = n__; } } while (0)
#define SG_TB(slot, var) SG_TBT(slot, var, 0)
#else
#define SG_TBT(slot, var, t) do {} while (0)
#define SG_T0() do {} while (0)
#define SG_T1(slot) do {} while (0)
#define SG_TA(var) do {} while (0)
#define SG_TB(slot, var) do {} while (0)
#endif

struct Key { uint32_t a, b; };

// plain C rotate (ptxas emits SHF.L.W): __funnelshift_l is an `asm volatile` in the CUDA headers, and volatile asms
// keep program order among themselves - independent threefry blocks could then never interleave (no ILP)
__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// One Threefry-2x32-20 block.
__device__ __forceinline__ void threefry(Key k, uint32_t& x0, uint32_t& x1) {
  const uint32_t ks0 = k.a, ks1 = k.b, ks2 = k.a ^ k.b ^ 0x1BD11BDAu;
  x0 += ks0; x1 += ks1;
#define SG_RND(r) x0 += x1; x1 = rotl32(x1, r); x1 ^= x0;
  SG_RND(13) SG_RND(15) SG_RND(26) SG_RND(6)   x0 += ks1; x1 += ks2 + 1u;
  SG_RND(17) SG_RND(29) SG_RND(16) SG_RND(24)  x0 += ks2; x1 += ks0 + 2u;
  SG_RND(13) SG_RND(15) SG_RND(26) SG_RND(6)   x0 += ks0; x1 += ks1 + 3u;
  SG_RND(17) SG_RND(29) SG_RND(16) SG_RND(24)  x0 += ks1; x1 += ks2 + 4u;
  SG_RND(13) SG_RND(15) SG_RND(26) SG_RND(6)   x0 += ks2; x1 += ks0 + 5u;
#undef SG_RND
}

// Block b of an n-word stream: returns words for stream positions b and b+h.
__device__ __forceinline__ void stream_block(Key k, uint32_t b, uint32_t n, uint32_t& w0, uint32_t& w1) {
  const uint32_t h = (n + 1u) >> 1;
  w0 = b;
  w1 = (b + h < n) ? b + h : 0u;   // the odd-length pad counter is 0
  threefry(k, w0, w1);
}

// Word m of an n-word stream (random_bits(key, (n,))[m]).
__device__ __forceinline__ uint32_t stream_word(Key k, uint32_t m, uint32_t n) {
  const uint32_t h = (n + 1u) >> 1;
  uint32_t w0, w1;
  stream_block(k, m < h ? m : m - h, n, w0, w1);
  return m < h ? w0 : w1;
}

// split(key, num)[j]
__device__ __forceinline__ Key split_at(Key k, uint32_t num, uint32_t j) {
  Key r;
  r.a = stream_word(k, 2u * j, 2u * num);
  r.b = stream_word(k, 2u * j + 1u, 2u * num);
  return r;
}

// split(key) -> (first, second) with two block evaluations.
__device__ __forceinline__ void split2(Key k, Key& first, Key& second) {
  uint32_t a0 = 0u, a1 = 2u, b0 = 1u, b1 = 3u;
  threefry(k, a0, a1);
  threefry(k, b0, b1);
  first.a = a0; first.b = b0; second.a = a1; second.b = b1;
}

// 1-instruction MUFU approximations (max relative error ~2^-22): far inside the 1e-5 parity bar, and they
// take the IEEE slow paths (10-30 instructions each) out of the per-element noise / preconditioner math.
__device__ __forceinline__ float fast_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float fast_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// XLA's fp32 ErfInv (Giles 2010) - NOT CUDA's erfinvf, which differs by ~6e-6 relative.
//   w = -log1p(-x*x) is evaluated as -ln(1 - rn(x*x)) with the MUFU logarithm (the product is rounded FIRST,
//   as XLA does: for |x| -> 1 that rounding dominates w, and a fused 1 - x*x would be "too accurate" to match;
//   1 - rn(x*x) itself is exact for x*x >= 1/2): the absolute error of w is
//   <= ~2e-7 in the central branch (|d ln p / dw| ~ 0.2 -> ~4e-8 relative in the result) and <= ~3 ulp of w in
//   the tail branch (w >= 5, 0.3 % of draws; sqrt halves it).  The two polynomial branches are a real branch:
//   the tail is rare, so a warp almost always runs only the 9 central FMAs.
__device__ __forceinline__ float erfinv_xla(float x) {
  const float w = -__logf(__fsub_rn(1.0f, __fmul_rn(x, x)));
  float p;
  if (w < 5.0f) {
    const float t = w - 2.5f;
    p = 2.81022636e-08f;
    p = fmaf(p, t, 3.43273939e-07f);
    p = fmaf(p, t, -3.5233877e-06f);
    p = fmaf(p, t, -4.39150654e-06f);
    p = fmaf(p, t, 0.00021858087f);
    p = fmaf(p, t, -0.00125372503f);
    p = fmaf(p, t, -0.00417768164f);
    p = fmaf(p, t, 0.246640727f);
    p = fmaf(p, t, 1.50140941f);
  } else {
    const float t = fast_sqrt(w) - 3.0f;
    p = -0.000200214257f;
    p = fmaf(p, t, 0.000100950558f);
    p = fmaf(p, t, 0.00134934322f);
    p = fmaf(p, t, -0.00367342844f);
    p = fmaf(p, t, 0.00573950773f);
    p = fmaf(p, t, -0.0076224613f);
    p = fmaf(p, t, 0.00943887047f);
    p = fmaf(p, t, 1.00167406f);
    p = fmaf(p, t, 2.83297682f);
  }
  return fabsf(x) == 1.0f ? copysignf(INFINITY, x) : p * x;
}

// bits -> N(0,1): uniform on [nextafter(-1,0), 1) then sqrt(2) * erf_inv.
__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float f = __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f;
  const float lo = -0.99999994f;                 // nextafter(-1, 0)
  const float u = fmaxf(lo, __fadd_rn(__fmul_rn(f, 2.0f), lo));   // (hi - lo) rounds to 2.0f
  return 1.41421356237f * erfinv_xla(u);
}

// random.normal(key, (n,))[e]
__device__ __forceinline__ float normal_at(Key k, uint32_t e, uint32_t n) {
  return bits_to_normal(stream_word(k, e, n));
}

// randint(key, (batch,), 0, span): jax draws hi bits with split(key)[0] and lo bits with
// split(key)[1]; multiplier = ((2^16 % span)^2 mod 2^32) % span; everything wraps in uint32.
struct RandintPlan {
  Key k_hi, k_lo;
  uint32_t span, mult, batch;
};

__device__ __forceinline__ RandintPlan randint_plan(Key key, uint32_t span, uint32_t batch) {
  RandintPlan p;
  split2(key, p.k_hi, p.k_lo);
  p.span = span; p.batch = batch;
  uint32_t m = 65536u % span;
  p.mult = (m * m) % span;
  return p;
}

__device__ __forceinline__ uint32_t randint_combine(const RandintPlan& p, uint32_t hi, uint32_t lo) {
  return ((hi % p.span) * p.mult + (lo % p.span)) % p.span;
}

// Stream block b -> indices for minibatch positions b and b+h (second one valid iff b+h < batch).
#ifndef SG_RANDINT_ONE_BLOCK
#define SG_RANDINT_ONE_BLOCK 1
#endif
__device__ __forceinline__ void randint_block(const RandintPlan& p, uint32_t b, uint32_t& i0, uint32_t& i1) {
  uint32_t l0, l1, h0 = 0u, h1 = 0u;
#if !SG_RANDINT_ONE_BLOCK
  stream_block(p.k_lo, b, p.batch, l0, l1);
  if (p.mult != 0u) stream_block(p.k_hi, b, p.batch, h0, h1);   // mult == 0 <=> span > 65536: hi unused
#else
  if (p.mult != 0u) {                     // span <= 65536: jax combines a hi and a lo draw
    // both threefry blocks in ONE basic block: two independent ~350-cycle dependency chains that the scheduler interleaves
    // (with the hi block alone behind the branch they ran one after the other: the README chain's index draw)
    stream_block(p.k_lo, b, p.batch, l0, l1);
    stream_block(p.k_hi, b, p.batch, h0, h1);
  } else {                                // mult == 0 <=> span > 65536: hi unused
    stream_block(p.k_lo, b, p.batch, l0, l1);
  }
#endif
  i0 = randint_combine(p, h0, l0);
  i1 = randint_combine(p, h1, l1);
}

}  // namespace sg
