// Launcher of the big-state sampler kernel per family; instantiated in big_inst_*.cu (compiled in parallel by _build.py).
#pragma once
#include "big_sampler.cuh"

namespace sg {

template <int FAMILY>
cudaError_t launch_big_inst(const BigModel& m, const BigRun& run, int blocks, size_t smem, cudaStream_t stream);

#ifdef SG_BIG_LAUNCH_IMPL
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
