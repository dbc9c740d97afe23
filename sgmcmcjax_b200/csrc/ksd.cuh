// K3 (SIMT fp32 variant): partial sum of the IMQ Stein kernel over a row block.
//
// ksd.py:7-37 with c = 1, beta = -1/2 collapses to
//   k0 = gg * b^-1/2 + (gd + D) * b^-3/2 - 3 dd * b^-5/2,   b = 1 + dd,
//   dd = |x_i - x_j|^2,  gg = g_i . g_j,  gd = (g_i - g_j) . (x_i - x_j).
// This variant forms the differences explicitly (like the reference does), so it has no
// Gram-expansion cancellation; it is the accuracy anchor for the tensor-core variant.
#pragma once
#include <cuda_runtime.h>

namespace sg {

constexpr int KSD_T = 64;    // tile edge (pairs)
constexpr int KSD_DK = 32;   // feature chunk staged in shared memory

__device__ __forceinline__ float imq_k0(float dd, float gg, float gd, float dim) {
  const float b = 1.0f + dd;
  const float r = rsqrtf(b);          // b^-1/2  (<= 2 ulp; b >= 1)
  const float r2 = r * r;             // b^-1
  const float r3 = r2 * r;            // b^-3/2
  return gg * r + (gd + dim) * r3 - 3.0f * dd * r3 * r2;
}

__global__ void __launch_bounds__(256) ksd_simt_kernel(const float* __restrict__ X, const float* __restrict__ G,
                                                       int n, int d, int row0, int row1, double* out) {
  __shared__ float sXI[KSD_T][KSD_DK + 1], sGI[KSD_T][KSD_DK + 1], sXJ[KSD_T][KSD_DK + 1], sGJ[KSD_T][KSD_DK + 1];
  __shared__ double sred[8];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i_base = row0 + blockIdx.y * KSD_T, j_base = blockIdx.x * KSD_T;
  float dd[4][4], gg[4][4], gd[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) dd[a][b] = gg[a][b] = gd[a][b] = 0.f;

  for (int k0 = 0; k0 < d; k0 += KSD_DK) {
    __syncthreads();
    for (int t = tid; t < KSD_T * KSD_DK; t += 256) {
      const int r = t / KSD_DK, k = t % KSD_DK;
      const int gi = i_base + r, gj = j_base + r, kk = k0 + k;
      const bool ki = kk < d;
      sXI[r][k] = (ki && gi < row1) ? X[(size_t)gi * d + kk] : 0.f;
      sGI[r][k] = (ki && gi < row1) ? G[(size_t)gi * d + kk] : 0.f;
      sXJ[r][k] = (ki && gj < n) ? X[(size_t)gj * d + kk] : 0.f;
      sGJ[r][k] = (ki && gj < n) ? G[(size_t)gj * d + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < KSD_DK; ++k) {
      float xi[4], gi[4], xj[4], gj[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) { xi[a] = sXI[ty + 16 * a][k]; gi[a] = sGI[ty + 16 * a][k]; }
#pragma unroll
      for (int b = 0; b < 4; ++b) { xj[b] = sXJ[tx + 16 * b][k]; gj[b] = sGJ[tx + 16 * b][k]; }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float dx = xi[a] - xj[b], dg = gi[a] - gj[b];
          dd[a][b] = fmaf(dx, dx, dd[a][b]);
          gg[a][b] = fmaf(gi[a], gj[b], gg[a][b]);
          gd[a][b] = fmaf(dg, dx, gd[a][b]);
        }
    }
  }
  double acc = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gi = i_base + ty + 16 * a, gj = j_base + tx + 16 * b;
      if (gi < row1 && gj < n) acc += (double)imq_k0(dd[a][b], gg[a][b], gd[a][b], (float)d);
    }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) sred[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sred[w];
    atomicAdd(out, t);
  }
}

}  // namespace sg
