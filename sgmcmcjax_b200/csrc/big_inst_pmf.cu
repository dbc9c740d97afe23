// big_sampler_kernel<PMF>.
#define SG_BIG_LAUNCH_IMPL
#include "big_launch.cuh"
namespace sg {
template cudaError_t launch_big_inst<SGMCMC_FAMILY_PMF>(const BigModel&, const BigRun&, int, size_t, cudaStream_t);
}  // namespace sg
