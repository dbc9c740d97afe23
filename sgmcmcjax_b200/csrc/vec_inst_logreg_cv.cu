// vec_sampler_kernel instantiations: logistic regression, fixed centre with the z_c row column (CV estimator).
#define SG_VEC_LAUNCH_IMPL
#include "vec_launch.cuh"
namespace sg {
SG_VEC_INSTANTIATE(SGMCMC_FAMILY_LOGREG, 2)
}  // namespace sg
