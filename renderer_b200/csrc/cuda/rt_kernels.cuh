// rt_kernels.cuh — launchers of the device hot path (implemented in *.cu, called from context.cu).
#pragma once
#include "device_types.cuh"

namespace b200r {

// Scratch of the ray tracer: `counters` = {tile counter / queue head, root-survivor count, hit count, spare};
// `queue` = pixel ids whose primary ray enters the root box (<= one per pixel); `hits` = 32-byte hit records.
struct RtBuffers {
    unsigned* counters = nullptr;
    int* queue = nullptr;
    void* hits = nullptr;
    size_t pixels = 0;
    bool forceMonolithic = false;
    bool noPrune = false;
};
cudaError_t launch_raytrace(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, RtBuffers& rt,
                            DeviceCounters* d_ctr, bool count, unsigned long long* d_tileProf, int numSMs, cudaStream_t stream,
                            int& launches);

// Scratch of the rasteriser: span records (80 B each), their count, and the 64-bit depth keys (one per pixel).
struct RasterBuffers {
    uint32_t* spans = nullptr;
    unsigned* spanCount = nullptr;
    unsigned spanCapacity = 0;
    unsigned long long* zkeys = nullptr;
};
cudaError_t launch_raster(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, RasterBuffers& rb,
                          DeviceCounters* d_ctr, bool count, int numSMs, cudaStream_t st, int& launches);
cudaError_t launch_shadowmap(const DeviceScene& sc, const float light_pos[3], const float world2light[9], unsigned* d_keys,
                             float* d_map, cudaStream_t st);
cudaError_t launch_mlaa(uint32_t* d_frame, uint32_t* d_scratch, int resX, int resY, int numSMs, cudaStream_t st, int& launches);
cudaError_t launch_division_selftest(unsigned long long samples, uint32_t seed, unsigned long long* d_mismatches,
                                     float* d_firstBad, int numSMs, cudaStream_t stream);
cudaError_t launch_deinterleave(const uint32_t* gathered, uint32_t* frame, uint32_t W, uint32_t H, uint32_t P,
                                int numSMs, cudaStream_t stream);

}  // namespace b200r
