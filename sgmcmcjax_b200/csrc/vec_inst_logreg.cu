// vec_sampler_kernel instantiations: logistic regression, standard estimator.
#define SG_VEC_LAUNCH_IMPL
#include "vec_launch.cuh"
namespace sg {
SG_VEC_INSTANTIATE(SGMCMC_FAMILY_LOGREG, 0)
}  // namespace sg
