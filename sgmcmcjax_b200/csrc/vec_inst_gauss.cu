// vec_sampler_kernel instantiations: gaussian_mean (standard and CV / SVRG estimators).
#define SG_VEC_LAUNCH_IMPL
#include "vec_launch.cuh"
namespace sg {
SG_VEC_INSTANTIATE(SGMCMC_FAMILY_GAUSSIAN_MEAN, 0)
SG_VEC_INSTANTIATE(SGMCMC_FAMILY_GAUSSIAN_MEAN, 1)
}  // namespace sg
