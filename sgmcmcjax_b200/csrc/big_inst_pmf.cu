// big_sampler_kernel<PMF>.
#define SG_BIG_LAUNCH_IMPL
#include "big_launch.cuh"
namespace sg {
template cudaError_t launch_big_inst<SGMCMC_FAMILY_PMF>(const BigModel&, const BigRun&, int, size_t, cudaStream_t);
template int big_resident_ctas<SGMCMC_FAMILY_PMF>(const BigModel&, size_t, int);
}  // namespace sg
