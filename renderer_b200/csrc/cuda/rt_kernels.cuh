// rt_kernels.cuh — launchers of the device hot path (implemented in *.cu, called from context.cu).
#pragma once
#include "device_types.cuh"

namespace b200r {

cudaError_t launch_raytrace(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, unsigned* d_tileCounter,
                            DeviceCounters* d_ctr, bool count, int numSMs, cudaStream_t stream);

}  // namespace b200r
