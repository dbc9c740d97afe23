// vec_sampler_kernel instantiations: gaussian_mean (standard and CV / SVRG estimators).
#define SG_VEC_LAUNCH_IMPL
#include "vec_launch.cuh"
namespace sg {
SG_VEC_INSTANTIATE(SGMCMC_FAMILY_GAUSSIAN_MEAN, 0)
SG_VEC_INSTANTIATE(SGMCMC_FAMILY_GAUSSIAN_MEAN, 1)
}  // namespace sg

// phase timers of this translation unit's kernels (tools/vec_phase_timing.py)
#ifdef SG_PHASE_TIMING
extern "C" int sgmcmc_debug_phase_cycles(unsigned long long* out32, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out32, sg::g_phase_cycles, sizeof(unsigned long long) * 32);
  if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(sg::g_phase_cycles, z, sizeof(z)); }
  return 0;
}
#endif
