// rt_kernels.cuh — launchers of the device hot path (implemented in *.cu, called from context.cu).
#pragma once
#include "device_types.cuh"

namespace b200r {

// Scratch of the ray tracer: `counters` = {tile counter / queue head, root-survivor count, hit count, spare};
// `queue` = (pixel, subtree) jobs of the primary rays that enter the root box (<= 8 per pixel); `hits` = 32-byte hit
// records; `keys`/`pend` = per pixel: best (hitZ, list position) so far and number of jobs still running.
struct RtBuffers {
    unsigned* counters = nullptr;
    void* queue = nullptr;
    unsigned long long* keys = nullptr;
    unsigned* pend = nullptr;
    void* hits = nullptr;
    size_t pixels = 0;
    bool forceMonolithic = false;
    bool noPrune = false;
    int fuseMode = 1;                          // simple config (1 light, no reflections/AO): 0 generic shade kernel, 1 fused lanes, 2 shadow jobs
    bool noRootCull = false;                   // B200R_NO_ROOT_RECT: K0 builds every pixel's ray (no screen rectangle)
    unsigned* sdon = nullptr;                  // fused path: merge words of shadow rays split over lanes (all zero between frames)
    void* srays = nullptr; unsigned* sword = nullptr; void* queue2 = nullptr;   // shadow-job pipeline: ray records (48 B), merge words, jobs
    int refillBelow = 0, innerBurst = 0;      // tuning overrides (B200R_REFILL_BELOW / B200R_INNER_BURST), 0 = built-in
    int sched = 0;                             // 0: rt_primary_kernel (rounds, default); 1: rt_wave_kernel (state-voting scheduler, measured slower) - B200R_RT_SCHED=wave
    int lateWeight = 0, prefetchCur = 1, longT = 0;   // rt_wave_kernel tuning (B200R_LATE_WEIGHT / B200R_NO_PREFETCH_CUR / B200R_LONG_T)
    int splitDepth = -1, blocksPerSM = 0;       // B200R_SPLIT_DEPTH (0..3, default 2) / B200R_BLOCKS_PER_SM (cap on resident CTAs of the persistent kernel)
    unsigned long long* warpProf = nullptr;   // developer tool (B200R_WARP_PROFILE): 4 x u64 per warp of rt_primary_kernel
    unsigned lastPrimaryWarps = 0;
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
    float4* attrs = nullptr;                   // modes 6-8: the winning fragment's interpolants per pixel (2 x float4); nullptr = shade inside the span walk
};
cudaError_t launch_raster(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, RasterBuffers& rb,
                          DeviceCounters* d_ctr, bool count, int numSMs, cudaStream_t st, int& launches);
cudaError_t launch_shadowmap(const DeviceScene& sc, const float light_pos[3], const float world2light[9], unsigned* d_keys,
                             float* d_map, cudaStream_t st);
// Scratch of mode 3 (wireframe): per-pixel fragment counts / offsets, scan block sums, fragment records (8 B each).
struct WireBuffers {
    uint32_t* counts = nullptr; uint32_t* offsets = nullptr; uint32_t* blockSums = nullptr; uint32_t* total = nullptr;
    void* frags = nullptr; uint32_t capacity = 0; size_t pixels = 0;
};
cudaError_t launch_wire_count(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, WireBuffers& wb, cudaStream_t st, int& launches);
cudaError_t launch_wire_emit(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, WireBuffers& wb, int numSMs, cudaStream_t st,
                             int& launches);
// d_lines: mlaa_lines_bytes(resX, resY) bytes of scratch for the two-stage path (line records per row); nullptr = row-scanning kernels
cudaError_t launch_mlaa(uint32_t* d_frame, uint32_t* d_scratch, int resX, int resY, int numSMs, cudaStream_t st, int& launches,
                        void* d_lines);
size_t mlaa_lines_bytes(int resX, int resY);
// SAH BVH build on the device (bvh_build.cu): host arrays in (vertex positions with a float stride, 3 indices per triangle), the
// flattened tree out (32-byte nodes in DFS pre-order + the triangle index list). *depth < 0: deeper than maxLevels - 1.
cudaError_t launch_bvh_build(const float* h_vertPos, int strideFloats, uint32_t nVerts, const uint32_t* h_idx, uint32_t nTris,
                             void* h_nodes_out, uint32_t nodesCap, int32_t* h_order_out, uint32_t* nNodes, int32_t* depth,
                             int maxLevels, cudaStream_t st, int& launches);
cudaError_t launch_division_selftest(unsigned long long samples, uint32_t seed, unsigned long long* d_mismatches,
                                     float* d_firstBad, int numSMs, cudaStream_t stream);
cudaError_t launch_deinterleave(const uint32_t* gathered, uint32_t* frame, uint32_t W, uint32_t H, uint32_t P,
                                int numSMs, cudaStream_t stream);

}  // namespace b200r
