// Launchers of the vector-family sampler kernel, one per template instantiation.  The instantiations are spread over
// several translation units (vec_inst_*.cu) that _build.py compiles in parallel; capi.cu sees only this declaration.
#pragma once
#include "vec_sampler.cuh"

namespace sg {

// cudaFuncSetAttribute (if the dynamic shared memory needs it) + launch of vec_sampler_kernel<FAMILY, CV, NCH, EXT>
// with one cluster of `cluster` CTAs (1, 2, 4, 8 or 16) of `threads` threads per chain.  A cluster shape the device
// cannot co-schedule for all chains at once is shrunk until it can (validate = true: asks cudaOccupancyMaxActiveClusters first; the caller
// caches the answer, *cluster_used, and passes validate = false from then on).
template <int FAMILY, int CV, int NCH, bool EXT>
cudaError_t launch_vec_inst(const VecModel& m, const VecRun& run, int threads, int cluster, size_t smem, cudaStream_t stream,
                            bool validate, int* cluster_used);

#ifdef SG_VEC_LAUNCH_IMPL
template <int FAMILY, int CV, int NCH, bool EXT>
cudaError_t launch_vec_inst(const VecModel& m, const VecRun& run, int threads, int cluster, size_t smem, cudaStream_t stream,
                            bool validate, int* cluster_used) {
  auto kern = vec_sampler_kernel<FAMILY, CV, NCH, EXT>;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (cluster <= 1) {
    if (cluster_used) *cluster_used = 1;
    kern<<<run.st.n_chains, threads, smem, stream>>>(m, run);
    return cudaGetLastError();
  }
  if (cluster > 8) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) { (void)cudaGetLastError(); cluster = 8; }
  }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = 1;
  for (; cluster > 1; --cluster) {
    attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)run.st.n_chains * cluster);
    if (!validate) break;
    int active = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&active, kern, &cfg);
    // every chain's cluster should be resident at once (a second wave doubles the launch time); the device places at
    // most `active` clusters of this shape (one GPC holds 16 - 20 SMs)
    if (e == cudaSuccess && active >= run.st.n_chains) break;
    if (e == cudaSuccess && active > 0 && cluster == 2) break;
    (void)cudaGetLastError();
  }
  if (cluster_used) *cluster_used = cluster;
  if (cluster <= 1) {
    kern<<<run.st.n_chains, threads, smem, stream>>>(m, run);
    return cudaGetLastError();
  }
  return cudaLaunchKernelEx(&cfg, kern, m, run);
}
#define SG_VEC_INSTANTIATE(FAMILY, CV)                                                                              \
  template cudaError_t launch_vec_inst<FAMILY, CV, 1, false>(const VecModel&, const VecRun&, int, int, size_t, cudaStream_t, bool, int*); \
  template cudaError_t launch_vec_inst<FAMILY, CV, 1, true>(const VecModel&, const VecRun&, int, int, size_t, cudaStream_t, bool, int*);  \
  template cudaError_t launch_vec_inst<FAMILY, CV, 2, false>(const VecModel&, const VecRun&, int, int, size_t, cudaStream_t, bool, int*); \
  template cudaError_t launch_vec_inst<FAMILY, CV, 2, true>(const VecModel&, const VecRun&, int, int, size_t, cudaStream_t, bool, int*);  \
  template cudaError_t launch_vec_inst<FAMILY, CV, 4, false>(const VecModel&, const VecRun&, int, int, size_t, cudaStream_t, bool, int*); \
  template cudaError_t launch_vec_inst<FAMILY, CV, 4, true>(const VecModel&, const VecRun&, int, int, size_t, cudaStream_t, bool, int*);
#endif

}  // namespace sg
