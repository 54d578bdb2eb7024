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
#include <cstdlib>
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

// Exact min/max are order-free EXCEPT for the sign of a zero: the reference folds `b < a ? b : a` / `a < b ? b : a` over the
// list in order, so when the extreme value is 0 and both +0 and -0 occur, the FIRST zero in list order stays. The block
// reductions below therefore carry the list position of each partial result and prefer the earlier one between equals.
struct MinMaxIdx {
    float v; int i;
    __device__ __forceinline__ void fold_min(float x, int xi) { if (x < v) { v = x; i = xi; } }           // serial fold: later equal values do not replace
    __device__ __forceinline__ void fold_max(float x, int xi) { if (v < x) { v = x; i = xi; } }
    __device__ __forceinline__ void merge_min(float x, int xi) { if (x < v || (!(v < x) && xi < i)) { v = x; i = xi; } }
    __device__ __forceinline__ void merge_max(float x, int xi) { if (v < x || (!(x < v) && xi < i)) { v = x; i = xi; } }
};

constexpr int ROOT_BLOCK = 512;
__global__ void __launch_bounds__(ROOT_BLOCK) bvh_rootbox_kernel(BvhBuild b)
{
    __shared__ float s_v[6][ROOT_BLOCK];
    __shared__ int s_i[6][ROOT_BLOCK];
    MinMaxIdx m[6];
    for (int c = 0; c < 3; c++) { m[c].v = FLT_MAX; m[c].i = 0x7fffffff; m[3 + c].v = -FLT_MAX; m[3 + c].i = 0x7fffffff; }
    for (uint32_t i = threadIdx.x; i < b.nTris; i += ROOT_BLOCK)
        for (int c = 0; c < 3; c++) { m[c].fold_min(b.tlo[3 * (size_t)i + c], (int)i); m[3 + c].fold_max(b.thi[3 * (size_t)i + c], (int)i); }
    for (int k = 0; k < 6; k++) { s_v[k][threadIdx.x] = m[k].v; s_i[k][threadIdx.x] = m[k].i; }
    __syncthreads();
    for (int s = ROOT_BLOCK / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s)
            for (int k = 0; k < 6; k++) {
                MinMaxIdx a; a.v = s_v[k][threadIdx.x]; a.i = s_i[k][threadIdx.x];
                if (k < 3) a.merge_min(s_v[k][threadIdx.x + s], s_i[k][threadIdx.x + s]);
                else a.merge_max(s_v[k][threadIdx.x + s], s_i[k][threadIdx.x + s]);
                s_v[k][threadIdx.x] = a.v; s_i[k][threadIdx.x] = a.i;
            }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        b.nstart[0] = 0; b.ncount[0] = (int32_t)b.nTris; b.ndepth[0] = 0; b.nleft[0] = -1; b.nright[0] = -1;
        for (int c = 0; c < 3; c++) { b.nlo[c] = s_v[c][0]; b.nhi[c] = s_v[3 + c][0]; }
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

// The same step for the few, large nodes near the root, one CTA per node: bvh_step_split() is a serial loop over the node's
// segment (tens of thousands of dependent loads for the root); here the count and the boxes are block reductions (exact
// min/max, integer sums) and the stable partition is a tiled ballot/prefix scatter that keeps list order on both sides.
constexpr int SPLIT_BLOCK = 256;
__global__ void __launch_bounds__(SPLIT_BLOCK) bvh_split_block_kernel(BvhBuild b, int levelBegin)
{
    const int node = levelBegin + (int)blockIdx.x;
    const int n = b.ncount[node];
    const unsigned long long key = b.best[node];
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n < 4 || key == BVH_NO_SPLIT) { if (tid == 0) { b.nleft[node] = -1; b.nright[node] = -1; } return; }
    const int ord = (int)(key & 0xffffffffull), axis = ord / 2048, index = ord % 2048;
    float split = 0.f;
    bvh_candidate(b, node, axis, index, split);
    int32_t* seg = b.order + b.nstart[node];
    int32_t* tmp = b.order2 + b.nstart[node];

    __shared__ float s_box[12][SPLIT_BLOCK];          // llo, lhi, rlo, rhi
    __shared__ int s_bi[12][SPLIT_BLOCK];             // list position of the element that set each partial (sign of zero, see MinMaxIdx)
    __shared__ int s_cnt[SPLIT_BLOCK];
    __shared__ int s_wl[SPLIT_BLOCK / 32], s_wr[SPLIT_BLOCK / 32];
    MinMaxIdx bx[12];
    for (int c = 0; c < 3; c++) {
        bx[c].v = FLT_MAX; bx[3 + c].v = -FLT_MAX; bx[6 + c].v = FLT_MAX; bx[9 + c].v = -FLT_MAX;
        bx[c].i = bx[3 + c].i = bx[6 + c].i = bx[9 + c].i = 0x7fffffff;
    }
    int nlMine = 0;
    for (int i = tid; i < n; i += SPLIT_BLOCK) {
        const size_t t = 3 * (size_t)seg[i];
        if (b.tctr[t + axis] < split) {
            nlMine++;
            for (int c = 0; c < 3; c++) { bx[c].fold_min(b.tlo[t + c], i); bx[3 + c].fold_max(b.thi[t + c], i); }
        } else {
            for (int c = 0; c < 3; c++) { bx[6 + c].fold_min(b.tlo[t + c], i); bx[9 + c].fold_max(b.thi[t + c], i); }
        }
    }
    for (int k = 0; k < 12; k++) { s_box[k][tid] = bx[k].v; s_bi[k][tid] = bx[k].i; }
    s_cnt[tid] = nlMine;
    __syncthreads();
    for (int s = SPLIT_BLOCK / 2; s > 0; s >>= 1) {
        if (tid < s) {
            s_cnt[tid] += s_cnt[tid + s];
            for (int k = 0; k < 12; k++) {
                MinMaxIdx a; a.v = s_box[k][tid]; a.i = s_bi[k][tid];
                if ((k % 6) < 3) a.merge_min(s_box[k][tid + s], s_bi[k][tid + s]);
                else a.merge_max(s_box[k][tid + s], s_bi[k][tid + s]);
                s_box[k][tid] = a.v; s_bi[k][tid] = a.i;
            }
        }
        __syncthreads();
    }
    const int nl = s_cnt[0];
    // stable scatter, tile by tile
    int runL = 0, runR = 0;
    for (int base = 0; base < n; base += SPLIT_BLOCK) {
        const int i = base + tid;
        const bool valid = i < n;
        int32_t tri = 0; bool left = false;
        if (valid) { tri = seg[i]; left = b.tctr[3 * (size_t)tri + axis] < split; }
        const unsigned mL = __ballot_sync(0xffffffffu, valid && left), mR = __ballot_sync(0xffffffffu, valid && !left);
        if (lane == 0) { s_wl[warp] = __popc(mL); s_wr[warp] = __popc(mR); }
        __syncthreads();
        int preL = 0, preR = 0, totL = 0, totR = 0;
        for (int w = 0; w < SPLIT_BLOCK / 32; w++) {
            if (w < warp) { preL += s_wl[w]; preR += s_wr[w]; }
            totL += s_wl[w]; totR += s_wr[w];
        }
        const unsigned lt = (1u << lane) - 1u;
        if (valid) {
            if (left) tmp[runL + preL + __popc(mL & lt)] = tri;
            else tmp[nl + runR + preR + __popc(mR & lt)] = tri;
        }
        runL += totL; runR += totR;
        __syncthreads();
    }
    for (int i = tid; i < n; i += SPLIT_BLOCK) seg[i] = tmp[i];
    if (tid == 0) {
        const int l = atomicAdd(b.poolCount, 2), r = l + 1;
        b.nleft[node] = l; b.nright[node] = r;
        b.nstart[l] = b.nstart[node]; b.ncount[l] = nl; b.ndepth[l] = b.ndepth[node] + 1;
        b.nstart[r] = b.nstart[node] + nl; b.ncount[r] = n - nl; b.ndepth[r] = b.ndepth[node] + 1;
        for (int c = 0; c < 3; c++) {
            b.nlo[3 * (size_t)l + c] = s_box[c][0]; b.nhi[3 * (size_t)l + c] = s_box[3 + c][0];
            b.nlo[3 * (size_t)r + c] = s_box[6 + c][0]; b.nhi[3 * (size_t)r + c] = s_box[9 + c][0];
        }
        b.best[l] = BVH_NO_SPLIT; b.best[r] = BVH_NO_SPLIT;
    }
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
                             int maxLevels, cudaStream_t st, int& launches, bool serialSplit)
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
    bvh_rootbox_kernel<<<1, ROOT_BLOCK, 0, st>>>(b);
    launches += 2;

    std::vector<int> levelBegin; levelBegin.push_back(0);
    int begin = 0, end = 1;
    while (begin < end) {
        if ((int)levelBegin.size() > maxLevels) { *depth = -1 - (int)levelBegin.size(); *nNodes = 0; return cudaGetLastError(); }
        const int n = end - begin;
        bvh_candidates_kernel<<<(unsigned)(n * 3 * CAND_CHUNKS), CAND_BLOCK, 0, st>>>(b, begin);
        if (n <= 2048 && !serialSplit) bvh_split_block_kernel<<<(unsigned)n, SPLIT_BLOCK, 0, st>>>(b, begin);       // few nodes, large segments: a CTA per node
        else bvh_split_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(b, begin, end);
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
