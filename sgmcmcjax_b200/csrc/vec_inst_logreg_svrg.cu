// vec_sampler_kernel instantiations: logistic regression, centre dot product in the kernel (SVRG).
#define SG_VEC_LAUNCH_IMPL
#include "vec_launch.cuh"
namespace sg {
SG_VEC_INSTANTIATE(SGMCMC_FAMILY_LOGREG, 1)
}  // namespace sg
