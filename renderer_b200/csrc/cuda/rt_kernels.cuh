// rt_kernels.cuh — launchers of the device hot path (implemented in *.cu, called from context.cu).
#pragma once
#include "device_types.cuh"

namespace b200r {

cudaError_t launch_raytrace(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, unsigned* d_tileCounter,
                            DeviceCounters* d_ctr, bool count, int numSMs, cudaStream_t stream);

cudaError_t launch_division_selftest(unsigned long long samples, uint32_t seed, unsigned long long* d_mismatches,
                                     float* d_firstBad, int numSMs, cudaStream_t stream);
cudaError_t launch_deinterleave(const uint32_t* gathered, uint32_t* frame, uint32_t W, uint32_t H, uint32_t P,
                                int numSMs, cudaStream_t stream);

}  // namespace b200r
