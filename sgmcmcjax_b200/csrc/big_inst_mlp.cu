// big_sampler_kernel<MLP> (csrc/mlp_grad.cuh).
#define SG_BIG_LAUNCH_IMPL
#include "big_launch.cuh"
#include "mlp_grad.cuh"
namespace sg {
template cudaError_t launch_big_inst<SGMCMC_FAMILY_MLP>(const BigModel&, const BigRun&, int, size_t, cudaStream_t);
template int big_resident_ctas<SGMCMC_FAMILY_MLP>(const BigModel&, size_t, int);
}  // namespace sg

// phase timers of this translation unit's kernel (tools/phase_timing.py)
#if defined(SG_PHASE_TIMING) && !defined(SG_PHASE_TIMING_VEC)
extern "C" int sgmcmc_debug_phase_cycles(unsigned long long* out32, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out32, sg::g_phase_cycles, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(sg::g_phase_cycles, z, sizeof(z)); }
  return 0;
}
#endif
