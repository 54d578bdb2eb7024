// bvh_build.cu — the reference's SAH BVH build (src/BVH.cc:96-371) and flattening (src/Raytracer.cc:651-718) on the device.
//
// The per-item work lives in csrc/bvh_steps.h (shared with the host-side test hook); this file only maps items to threads:
//   bvh_triangles_kernel      one thread per triangle: box, centre, identity ordering
//   bvh_rootbox_kernel        one CTA: scene box (exact min/max, order-free)
//   per level of the tree (host loop; the number of nodes of the next level comes back with a 4-byte copy):
//     bvh_candidates_kernel   one thread per (node, axis, candidate): counts + boxes over the node's segment, fp32 SAH cost,
//                             64-bit atomicMin of (cost bits, candidate order) per node = the reference's first best split
//     bvh_split_kernel        one thread per node: leaf, or stable partition of its segment + two children
//   bvh_size/index/emit       DFS pre-order numbering from subtree sizes, 32-byte CacheFriendlyBVHNode records
// Every thread of bvh_candidates_kernel walks the SAME segment (broadcast loads): ~3 * 1024/(depth+1) candidates per node
// times the node's triangles = the reference's own O(n * 1024 * 3) per node, spread over the machine. The top levels have
// few nodes and therefore few threads; C4's 65 534 triangles still build in tens of milliseconds (reference: 1.4-4.1 s).
#include <cstdio>
#include <vector>

#include "../bvh_steps.h"
#include "rt_kernels.cuh"

namespace b200r {
namespace {

__global__ void bvh_triangles_kernel(BvhBuild b, const float* __restrict__ vertPos, int strideFloats, const uint32_t* __restrict__ idx)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < b.nTris) bvh_step_triangle(b, i, vertPos, strideFloats, idx[3 * i], idx[3 * i + 1], idx[3 * i + 2]);
}

__global__ void __launch_bounds__(1024) bvh_rootbox_kernel(BvhBuild b)
{
    __shared__ float s_lo[3][1024], s_hi[3][1024];
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = threadIdx.x; i < b.nTris; i += blockDim.x)
        for (int c = 0; c < 3; c++) { lo[c] = bvh_min2(lo[c], b.tlo[3 * (size_t)i + c]); hi[c] = bvh_max2(hi[c], b.thi[3 * (size_t)i + c]); }
    for (int c = 0; c < 3; c++) { s_lo[c][threadIdx.x] = lo[c]; s_hi[c][threadIdx.x] = hi[c]; }
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
            for (int c = 0; c < 3; c++) {
                s_lo[c][threadIdx.x] = bvh_min2(s_lo[c][threadIdx.x], s_lo[c][threadIdx.x + s]);
                s_hi[c][threadIdx.x] = bvh_max2(s_hi[c][threadIdx.x], s_hi[c][threadIdx.x + s]);
            }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        b.nstart[0] = 0; b.ncount[0] = (int32_t)b.nTris; b.ndepth[0] = 0; b.nleft[0] = -1; b.nright[0] = -1;
        for (int c = 0; c < 3; c++) { b.nlo[c] = s_lo[c][0]; b.nhi[c] = s_hi[c][0]; }
        b.best[0] = BVH_NO_SPLIT; b.ndfs[0] = 0; *b.poolCount = 1;
    }
}

constexpr int CAND_BLOCK = 256, CAND_CHUNKS = BVH_MAX_CANDIDATES / CAND_BLOCK;
static_assert(BVH_MAX_CANDIDATES % CAND_BLOCK == 0, "candidate chunks");

__global__ void __launch_bounds__(CAND_BLOCK) bvh_candidates_kernel(BvhBuild b, int levelBegin)
{
    const int chunk = (int)(blockIdx.x % CAND_CHUNKS), axis = (int)((blockIdx.x / CAND_CHUNKS) % 3);
    const int node = levelBegin + (int)(blockIdx.x / (CAND_CHUNKS * 3));
    const unsigned long long key = bvh_step_candidate(b, node, axis, chunk * CAND_BLOCK + (int)threadIdx.x);
    if (key != BVH_NO_SPLIT) atomicMin(&b.best[node], key);
}

__global__ void bvh_split_kernel(BvhBuild b, int levelBegin, int levelEnd)
{
    const int node = levelBegin + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (node < levelEnd) bvh_step_split(b, node, [](int32_t* pc) { return atomicAdd(pc, 2); });
}

__global__ void bvh_size_kernel(BvhBuild b, int levelBegin, int levelEnd)
{
    const int node = levelBegin + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (node < levelEnd) bvh_step_size(b, node);
}
__global__ void bvh_index_kernel(BvhBuild b, int levelBegin, int levelEnd)
{
    const int node = levelBegin + (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (node < levelEnd) bvh_step_index(b, node);
}
__global__ void bvh_emit_kernel(BvhBuild b, int nNodes, BvhNodeOut* __restrict__ out)
{
    const int node = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (node < nNodes) bvh_step_emit(b, node, out);
}

struct DevBuf {
    std::vector<void*> ptrs;
    template <class T> cudaError_t alloc(T** p, size_t n)
    {
        cudaError_t e = cudaMalloc((void**)p, n * sizeof(T) + 16);
        if (e == cudaSuccess) ptrs.push_back(*p);
        return e;
    }
    ~DevBuf() { for (void* p : ptrs) cudaFree(p); }
};

}  // namespace

#define BCU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return e__; } while (0)

// Returns cudaSuccess and *depth = -1 - (levels) when the tree is deeper than maxDepth allows (caller reports it).
cudaError_t launch_bvh_build(const float* h_vertPos, int strideFloats, uint32_t nVerts, const uint32_t* h_idx, uint32_t nTris,
                             void* h_nodes_out, uint32_t nodesCap, int32_t* h_order_out, uint32_t* nNodes, int32_t* depth,
                             int maxLevels, cudaStream_t st, int& launches)
{
    DevBuf db;
    const size_t N = nTris, cap = 2 * N + 2;
    float* d_vert = nullptr; uint32_t* d_idx = nullptr; BvhNodeOut* d_out = nullptr;
    BvhBuild b{};
    b.nTris = nTris;
    BCU(db.alloc(&d_vert, (size_t)nVerts * strideFloats)); BCU(db.alloc(&d_idx, 3 * N));
    BCU(db.alloc(&b.tlo, 3 * N)); BCU(db.alloc(&b.thi, 3 * N)); BCU(db.alloc(&b.tctr, 3 * N));
    BCU(db.alloc(&b.order, N)); BCU(db.alloc(&b.order2, N));
    BCU(db.alloc(&b.nstart, cap)); BCU(db.alloc(&b.ncount, cap)); BCU(db.alloc(&b.ndepth, cap));
    BCU(db.alloc(&b.nleft, cap)); BCU(db.alloc(&b.nright, cap)); BCU(db.alloc(&b.nlo, 3 * cap)); BCU(db.alloc(&b.nhi, 3 * cap));
    BCU(db.alloc(&b.best, cap)); BCU(db.alloc(&b.nsize, cap)); BCU(db.alloc(&b.ndfs, cap)); BCU(db.alloc(&b.poolCount, 1));
    BCU(db.alloc(&d_out, cap));
    BCU(cudaMemcpyAsync(d_vert, h_vertPos, (size_t)nVerts * strideFloats * sizeof(float), cudaMemcpyHostToDevice, st));
    BCU(cudaMemcpyAsync(d_idx, h_idx, 3 * N * sizeof(uint32_t), cudaMemcpyHostToDevice, st));

    bvh_triangles_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(b, d_vert, strideFloats, d_idx);
    bvh_rootbox_kernel<<<1, 1024, 0, st>>>(b);
    launches += 2;

    std::vector<int> levelBegin; levelBegin.push_back(0);
    int begin = 0, end = 1;
    while (begin < end) {
        if ((int)levelBegin.size() > maxLevels) { *depth = -1 - (int)levelBegin.size(); *nNodes = 0; return cudaGetLastError(); }
        const int n = end - begin;
        bvh_candidates_kernel<<<(unsigned)(n * 3 * CAND_CHUNKS), CAND_BLOCK, 0, st>>>(b, begin);
        bvh_split_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(b, begin, end);
        launches += 2;
        int32_t pool = 0;
        BCU(cudaMemcpyAsync(&pool, b.poolCount, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        BCU(cudaStreamSynchronize(st));
        begin = end; end = pool;
        levelBegin.push_back(begin);
    }
    const int levels = (int)levelBegin.size() - 1, pool = end;
    for (int l = levels - 1; l >= 0; l--) {
        const int n = levelBegin[l + 1] - levelBegin[l];
        bvh_size_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b, levelBegin[l], levelBegin[l + 1]);
    }
    for (int l = 0; l < levels; l++) {
        const int n = levelBegin[l + 1] - levelBegin[l];
        bvh_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(b, levelBegin[l], levelBegin[l + 1]);
    }
    bvh_emit_kernel<<<(unsigned)((pool + 255) / 256), 256, 0, st>>>(b, pool, d_out);
    launches += 2 * levels + 1;
    if ((uint32_t)pool > nodesCap) { *nNodes = (uint32_t)pool; *depth = levels - 1; return cudaErrorInvalidValue; }
    BCU(cudaMemcpyAsync(h_nodes_out, d_out, (size_t)pool * sizeof(BvhNodeOut), cudaMemcpyDeviceToHost, st));
    BCU(cudaMemcpyAsync(h_order_out, b.order, N * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    BCU(cudaStreamSynchronize(st));
    *nNodes = (uint32_t)pool; *depth = levels - 1;
    return cudaGetLastError();
}

}  // namespace b200r
