// bvh_steps_host.cpp — the data-parallel BVH build steps (csrc/bvh_steps.h) run in plain loops on the host.
//
// This is NOT a rendering or build path of the product (the host plumbing builds its tree in csrc/host/bvh.cpp, the device
// build is csrc/cuda/bvh_build.cu): it exists so that the step functions the CUDA kernels are made of are exercised by the
// CPU test suite - item by item, in the same level order - against the reference's own .bvh caches.
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/b200render.h"
#include "../bvh_steps.h"

namespace b200r { void set_global_error(const std::string& s); }

extern "C" int b200r_selftest_bvh_steps_host(const b200r_vertex* verts, uint32_t n_verts, const b200r_tri* tris, uint32_t n_tris,
                                             b200r_bvhnode* nodes_out, uint32_t nodes_cap, int32_t* tri_idx_out,
                                             uint32_t* n_nodes, int32_t* depth)
{
    using namespace b200r;
    if (!verts || !tris || !nodes_out || !tri_idx_out || !n_nodes || !depth || n_tris == 0) {
        set_global_error("b200r_selftest_bvh_steps_host: bad argument");
        return B200R_EINVAL;
    }
    (void)n_verts;
    const size_t N = n_tris, cap = 2 * N + 2;
    std::vector<float> tlo(3 * N), thi(3 * N), tctr(3 * N), nlo(3 * cap), nhi(3 * cap);
    std::vector<int32_t> order(N), order2(N), nstart(cap), ncount(cap), ndepth(cap), nleft(cap, -1), nright(cap, -1), nsize(cap), ndfs(cap);
    std::vector<unsigned long long> best(cap, BVH_NO_SPLIT);
    int32_t poolCount = 1;
    BvhBuild b;
    b.nTris = n_tris; b.tlo = tlo.data(); b.thi = thi.data(); b.tctr = tctr.data(); b.order = order.data(); b.order2 = order2.data();
    b.nstart = nstart.data(); b.ncount = ncount.data(); b.ndepth = ndepth.data(); b.nleft = nleft.data(); b.nright = nright.data();
    b.nlo = nlo.data(); b.nhi = nhi.data(); b.best = best.data(); b.nsize = nsize.data(); b.ndfs = ndfs.data(); b.poolCount = &poolCount;

    for (uint32_t i = 0; i < n_tris; i++)
        bvh_step_triangle(b, i, verts[0].pos, (int)(sizeof(b200r_vertex) / sizeof(float)), tris[i].a, tris[i].b, tris[i].c);
    // root = node 0: the whole list, box of all triangles (BVH.cc:341-346, 367-368)
    nstart[0] = 0; ncount[0] = (int32_t)N; ndepth[0] = 0;
    for (int c = 0; c < 3; c++) { nlo[c] = FLT_MAX; nhi[c] = -FLT_MAX; }
    for (size_t i = 0; i < N; i++)
        for (int c = 0; c < 3; c++) { nlo[c] = bvh_min2(nlo[c], tlo[3 * i + c]); nhi[c] = bvh_max2(nhi[c], thi[3 * i + c]); }

    std::vector<int> levelBegin; levelBegin.push_back(0);
    int begin = 0, end = 1;
    while (begin < end) {
        for (int node = begin; node < end; node++)
            for (int axis = 0; axis < 3; axis++)
                for (int index = 0; index < BVH_MAX_CANDIDATES; index++) {
                    float ts;
                    if (ncount[node] < 4 || !bvh_candidate(b, node, axis, index, ts)) break;   // (the device simply runs all indices)
                    const unsigned long long k = bvh_step_candidate(b, node, axis, index);
                    if (k < best[node]) best[node] = k;                     // (atomicMin on the device)
                }
        for (int node = begin; node < end; node++)
            bvh_step_split(b, node, [](int32_t* pc) { const int v = *pc; *pc += 2; return v; });
        begin = end; end = poolCount;
        levelBegin.push_back(begin);
    }
    const int levels = (int)levelBegin.size() - 1;          // levelBegin[levels] == poolCount
    for (int l = levels - 1; l >= 0; l--)
        for (int node = levelBegin[l]; node < levelBegin[l + 1]; node++) bvh_step_size(b, node);
    ndfs[0] = 0;
    for (int l = 0; l < levels; l++)
        for (int node = levelBegin[l]; node < levelBegin[l + 1]; node++) bvh_step_index(b, node);
    if ((uint32_t)poolCount > nodes_cap) { set_global_error("b200r_selftest_bvh_steps_host: nodes_cap too small"); return B200R_EINVAL; }
    for (int node = 0; node < poolCount; node++) bvh_step_emit(b, node, reinterpret_cast<BvhNodeOut*>(nodes_out));
    memcpy(tri_idx_out, order.data(), N * sizeof(int32_t));
    *n_nodes = (uint32_t)poolCount;
    *depth = levels - 1;
    return B200R_OK;
}
