// Launcher of the big-state sampler kernel per family; instantiated in big_inst_*.cu (compiled in parallel by _build.py).
#pragma once
#include "big_sampler.cuh"

namespace sg {

template <int FAMILY>
cudaError_t launch_big_inst(const BigModel& m, const BigRun& run, int blocks, size_t smem, cudaStream_t stream);
// CTAs of big_sampler_kernel<FAMILY> the device holds at once for this model (resident slots of the in-kernel queue)
template <int FAMILY>
int big_resident_ctas(const BigModel& m, size_t smem, int num_sms);

#ifdef SG_BIG_LAUNCH_IMPL
template <int FAMILY>
int big_resident_ctas(const BigModel& m, size_t smem, int num_sms) {
  auto kern = big_sampler_kernel<FAMILY>;
  if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, big_threads<FAMILY>(m), smem) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return per_sm * num_sms;
}
template <int FAMILY>
cudaError_t launch_big_inst(const BigModel& m, const BigRun& run, int blocks, size_t smem, cudaStream_t stream) {
  auto kern = big_sampler_kernel<FAMILY>;
  if (smem > 48 * 1024) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  kern<<<blocks, big_threads<FAMILY>(m), smem, stream>>>(m, run);
  return cudaGetLastError();
}
#endif

}  // namespace sg
