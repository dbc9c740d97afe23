// Launchers of the vector-family sampler kernel, one per template instantiation.  The instantiations are spread over
// several translation units (vec_inst_*.cu) that _build.py compiles in parallel; capi.cu sees only this declaration.
#pragma once
#include "vec_sampler.cuh"

namespace sg {

// cudaFuncSetAttribute (if the dynamic shared memory needs it) + launch of vec_sampler_kernel<FAMILY, CV, NCH, EXT>
// with one CTA of `threads` threads per chain.
template <int FAMILY, int CV, int NCH, bool EXT>
cudaError_t launch_vec_inst(const VecModel& m, const VecRun& run, int threads, size_t smem, cudaStream_t stream);

#ifdef SG_VEC_LAUNCH_IMPL
template <int FAMILY, int CV, int NCH, bool EXT>
cudaError_t launch_vec_inst(const VecModel& m, const VecRun& run, int threads, size_t smem, cudaStream_t stream) {
  auto kern = vec_sampler_kernel<FAMILY, CV, NCH, EXT>;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<run.st.n_chains, threads, smem, stream>>>(m, run);
  return cudaGetLastError();
}
#define SG_VEC_INSTANTIATE(FAMILY, CV)                                                                              \
  template cudaError_t launch_vec_inst<FAMILY, CV, 1, false>(const VecModel&, const VecRun&, int, size_t, cudaStream_t); \
  template cudaError_t launch_vec_inst<FAMILY, CV, 1, true>(const VecModel&, const VecRun&, int, size_t, cudaStream_t);  \
  template cudaError_t launch_vec_inst<FAMILY, CV, 2, false>(const VecModel&, const VecRun&, int, size_t, cudaStream_t); \
  template cudaError_t launch_vec_inst<FAMILY, CV, 2, true>(const VecModel&, const VecRun&, int, size_t, cudaStream_t);  \
  template cudaError_t launch_vec_inst<FAMILY, CV, 4, false>(const VecModel&, const VecRun&, int, size_t, cudaStream_t); \
  template cudaError_t launch_vec_inst<FAMILY, CV, 4, true>(const VecModel&, const VecRun&, int, size_t, cudaStream_t);
#endif

}  // namespace sg
