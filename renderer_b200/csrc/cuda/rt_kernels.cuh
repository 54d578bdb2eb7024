// rt_kernels.cuh — launchers of the device hot path (implemented in *.cu, called from context.cu).
#pragma once
#include "device_types.cuh"

namespace b200r {

// Developer switches (b200r_set_switch; defaults from the environment variables B200R_<NAME>, read ONCE at b200r_init).
// They select cross-check variants of the kernels for the parity tests and A/B measurements; none changes a result.
struct Switches {
    int monolithic_rt = 0;        // rt_frame_kernel for every ray-traced frame
    int no_prune = 0;             // no distance pruning in the closest-hit traversal
    int no_fuse = 0;              // simple configuration through hit records + the generic route as well
    int no_wavefront = 0;         // generic configurations: rt_shade_kernel (one thread per hit walks every secondary ray) instead of the wavefront
    int rt_legacy = 0;            // round 1's job pipeline (rt_rootcull_kernel + rt_primary_kernel) instead of rt_pool_kernel
    int no_root_rect = 0;         // no screen rectangle around the root box: every pixel's ray is built
    int pool_small = 0;           // 128-entry pools: exercises the overflow guard of rt_pool_kernel
    int split_depth = -1;         // job pipeline: BVH levels expanded into jobs (0..3; default 2)
    int raster_inline_shade = 0;  // modes 6-8 shade inside the span walk instead of the per-pixel pass
    int mlaa_scan = 0;            // row-scanning MLAA kernels instead of the two-stage path
    int mlaa_fullscan = 0;        // two-stage MLAA, lines walked by the scanning thread
    int mlaa_nobatch = 0;         // MLAA walks load one word per step
    int mlaa_no_tma = 0;          // vertical MLAA blends walk the frame in L2 instead of a TMA-staged strip in shared memory
    int no_frame_overlap = 0;     // b200r_render_async keeps ray-traced frames on one stream
    int bvh_serial_split = 0;     // BVH build: one thread per node in every level
    int pool_policy = 0;          // rt_pool_kernel: how the inner pool is popped (rt_pool.cu PoolParams)
    int pool_leaf_min = 0, pool_sort_min = 0, pool_shade_min = 0, pool_refill_min = 0, pool_low_water = 0, pool_dry = 0;   // rt_pool_kernel thresholds (0 = built-in)
    int pool_tiles_per_warp = 0;  // rt_pool_kernel: grid sized for this many 8x4 tiles per warp (0 = built-in: 1 for a frame alone, 2 with frames in flight)
    int pool_cta_warps = 0;       // rt_pool_kernel: warps per CTA (0 = built-in 4; 8, 2, 1: the same number of warps per SM in larger / smaller CTAs)
    int pool_occ3 = 0;            // rt_pool_kernel: C2-type frames with 3 CTAs per SM (76 registers, 512-entry pools) instead of 4
    int pool_scatter = 0;         // rt_pool_kernel: 0 = scattered 4-pixel groups for a frame alone, whole 8x4 tiles for frames in flight; 1 / 2 force
    int pool_stats = 0;           // rt_pool_kernel adds its per-phase iteration / lane counts to the work counters (tools/pool_stats.py)
};

// Scratch of the ray tracer: `counters` = {tile counter / job queue head, job count, hit count, pixel counter of the pooled kernel};
// `hits` = 32-byte hit records (generic configurations: one per pixel at most). `queue` ((pixel, subtree) jobs, <= 8 per pixel) and
// `keys` (per-pixel merge words) belong to the job pipeline and are only allocated for counting / legacy runs.
struct RtBuffers {
    unsigned* counters = nullptr;
    void* hits = nullptr;
    size_t pixels = 0;
    void* queue = nullptr;
    unsigned long long* keys = nullptr;
    size_t legacyPixels = 0;
    // wavefront of the generic configurations (rt_wavefront.cu): hits of levels 1 and 2, one 64-byte path record per primary hit,
    // reflection-ray records (one per hit), and per chunk of `wfChunk` hits: any-hit ray records (wfStride per hit), their result
    // bytes, the AO cosines and the per-hit shading context
    void* wfHits1 = nullptr; void* wfHits2 = nullptr; void* wfPaths = nullptr;
    float4* wfRefl = nullptr; float4* wfRays = nullptr; unsigned char* wfOcc = nullptr; float* wfCos = nullptr; float4* wfCtx = nullptr;
    size_t wfPixels = 0; unsigned wfChunk = 0, wfStride = 0;
    bool inFlight = false;                    // this frame is one of several in flight (b200r_render_device_slot, b200r_render_async at depth > 1)
};
cudaError_t launch_raytrace(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, RtBuffers& rt, const Switches& sw,
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
                        void* d_lines, const Switches& sw);
size_t mlaa_lines_bytes(int resX, int resY);
// SAH BVH build on the device (bvh_build.cu): host arrays in (vertex positions with a float stride, 3 indices per triangle), the
// flattened tree out (32-byte nodes in DFS pre-order + the triangle index list). *depth < 0: deeper than maxLevels - 1.
cudaError_t launch_bvh_build(const float* h_vertPos, int strideFloats, uint32_t nVerts, const uint32_t* h_idx, uint32_t nTris,
                             void* h_nodes_out, uint32_t nodesCap, int32_t* h_order_out, uint32_t* nNodes, int32_t* depth,
                             int maxLevels, cudaStream_t st, int& launches, bool serialSplit);
// rt_pool.cu: the pooled traversal kernel (clears the frame, then writes lit pixels / hit records)
bool pool_supported(const DeviceScene& sc);
cudaError_t rt_pool_configure();            // once per device: opt in to > 48 KB of dynamic shared memory
cudaError_t launch_rt_pool(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, bool fused, bool prune, const Switches& sw,
                           unsigned* pixelCounter, void* hits, unsigned* hitCount, int numSMs, cudaStream_t stream, int& launches,
                           DeviceCounters* stats = nullptr, bool inFlight = false);
cudaError_t launch_rt_pool_queue(const DeviceScene& sc, const FrameParams& fp, bool anyhit, bool prune, const Switches& sw, unsigned* cursor,
                                 const float4* rays, const unsigned* count, unsigned first, unsigned cap, unsigned stride,
                                 unsigned char* occ, void* hits, unsigned* hitCount, int numSMs, cudaStream_t stream, int& launches);
size_t wavefront_path_bytes();
cudaError_t launch_rt_wavefront(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, RtBuffers& rt, const Switches& sw,
                                bool prune, int numSMs, cudaStream_t stream, int& launches);
cudaError_t launch_division_selftest(unsigned long long samples, uint32_t seed, unsigned long long* d_mismatches,
                                     float* d_firstBad, int numSMs, cudaStream_t stream);
cudaError_t launch_deinterleave(const uint32_t* gathered, uint32_t* frame, uint32_t W, uint32_t H, uint32_t P,
                                int numSMs, cudaStream_t stream);

}  // namespace b200r
