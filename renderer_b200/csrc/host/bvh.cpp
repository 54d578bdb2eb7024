// bvh.cpp — host plumbing: SAH BVH build, DFS flattening and the .bvh on-disk cache.
//
// Produces THE SAME TREE as the reference's scalar builder, because traversal order and the set of
// boxes tested decide which triangles a ray ever sees (SURVEY.md §7 "hard parts"):
//   Recurse / CreateBVH            reference src/BVH.cc:96-371 (scalar path, SIMD_SSE undefined)
//   PopulateCacheFriendlyBVH       reference src/Raytracer.cc:651-682 (DFS pre-order, left = self+1)
//   CreateCFBVH depth check        reference src/Raytracer.cc:711-717
//   .bvh cache                     reference src/Raytracer.cc:747-786
// The split search walks the same (axis, testSplit) candidates in the same order with the same fp32
// expressions; only the per-candidate O(n) pass is replaced by exact prefix/suffix boxes.
#include "scene.h"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <algorithm>

namespace b200r {
namespace {

struct BBoxTmp {            // BVH.cc:73-88
    float lo[3], hi[3], ctr[3];
    int32_t tri;
};

inline float fmin2(float a, float b) { return b < a ? b : a; }
inline float fmax2(float a, float b) { return a < b ? b : a; }

struct Box {
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    void grow(const BBoxTmp& v)
    {
        for (int c = 0; c < 3; c++) { lo[c] = fmin2(lo[c], v.lo[c]); hi[c] = fmax2(hi[c], v.hi[c]); }
    }
    float half_area() const   // side1*side2 + side2*side3 + side3*side1  (BVH.cc:110-115)
    {
        float s1 = hi[0] - lo[0], s2 = hi[1] - lo[1], s3 = hi[2] - lo[2];
        return s1 * s2 + s2 * s3 + s3 * s1;
    }
};

struct TmpNode {
    Box box;
    int left = -1, right = -1;       // children (TmpNode indices) or -1 for a leaf
    std::vector<int32_t> tris;       // leaf payload, in work-list order
};

struct Builder {
    std::vector<TmpNode> pool;

    int recurse(std::vector<BBoxTmp>& work, int depth)
    {
        int me = (int)pool.size();
        pool.emplace_back();
        if (work.size() < 4) {                                   // BVH.cc:99-104
            for (auto& w : work) pool[me].tris.push_back(w.tri);
            return me;
        }
        Box all;
        for (auto& w : work) all.grow(w);
        float minCost = work.size() * all.half_area();           // size_t -> float (BVH.cc:115)
        float bestSplit = FLT_MAX; int bestAxis = -1;

        // The reference evaluates every candidate (axis, testSplit) with a full pass over the work
        // list (BVH.cc:160-206). "center < testSplit" selects a PREFIX of the list sorted by center, and
        // box unions are exact min/max, so prefix/suffix boxes over the sorted order give bit-identical
        // (countLeft, countRight, left box, right box) for every candidate in O(log n) each.
        const size_t n = work.size();
        std::vector<uint32_t> order(n);
        std::vector<float> key(n);
        std::vector<Box> prefix(n + 1), suffix(n + 1);
        for (int axis = 0; axis < 3; axis++) {                   // BVH.cc:120
            float start = all.lo[axis], stop = all.hi[axis];
            if (fabsf(stop - start) < 1e-4) continue;            // float promoted, compared as double
            float step = (stop - start) / (1024.f / (depth + 1.f));
            for (size_t i = 0; i < n; i++) order[i] = (uint32_t)i;
            std::sort(order.begin(), order.end(),
                      [&](uint32_t a, uint32_t b) { return work[a].ctr[axis] < work[b].ctr[axis]; });
            for (size_t i = 0; i < n; i++) key[i] = work[order[i]].ctr[axis];
            prefix[0] = Box();
            for (size_t i = 0; i < n; i++) { prefix[i + 1] = prefix[i]; prefix[i + 1].grow(work[order[i]]); }
            suffix[n] = Box();
            for (size_t i = n; i-- > 0;) { suffix[i] = suffix[i + 1]; suffix[i].grow(work[order[i]]); }
            for (float testSplit = start + step; testSplit < stop - step; testSplit += step) {
                size_t k = std::lower_bound(key.begin(), key.end(), testSplit) - key.begin();
                int countLeft = (int)k, countRight = (int)(n - k);
                if (countLeft <= 1 || countRight <= 1) { if (testSplit + step == testSplit) break; continue; }
                float surfaceLeft = prefix[k].half_area(), surfaceRight = suffix[k].half_area();
                float totalCost = surfaceLeft * countLeft + surfaceRight * countRight;
                if (totalCost < minCost) { minCost = totalCost; bestSplit = testSplit; bestAxis = axis; }
                if (testSplit + step == testSplit) break;        // the reference would spin forever here
            }
        }

        if (bestAxis == -1) {                                    // BVH.cc:210-216
            for (auto& w : work) pool[me].tris.push_back(w.tri);
            return me;
        }
        std::vector<BBoxTmp> left, right; Box lb, rb;            // BVH.cc:219-262
        for (auto& v : work) {
            if (v.ctr[bestAxis] < bestSplit) { left.push_back(v); lb.grow(v); }
            else { right.push_back(v); rb.grow(v); }
        }
        std::vector<BBoxTmp>().swap(work);
        int l = recurse(left, depth + 1);
        pool[l].box = lb;
        int r = recurse(right, depth + 1);
        pool[r].box = rb;
        pool[me].left = l; pool[me].right = r;
        return me;
    }
};

}  // namespace

void Scene::build_bvh_from_scratch()
{
    // CreateBVH, BVH.cc:322-371
    std::vector<BBoxTmp> work; work.reserve(tris.size());
    Box all;
    for (size_t j = 0; j < tris.size(); j++) {
        BBoxTmp b;
        for (int c = 0; c < 3; c++) { b.lo[c] = FLT_MAX; b.hi[c] = -FLT_MAX; }
        const uint32_t idx[3] = {tris[j].a, tris[j].b, tris[j].c};
        for (int k = 0; k < 3; k++)
            for (int c = 0; c < 3; c++) {
                b.lo[c] = fmin2(b.lo[c], verts[idx[k]].pos[c]);
                b.hi[c] = fmax2(b.hi[c], verts[idx[k]].pos[c]);
            }
        all.grow(b);
        for (int c = 0; c < 3; c++) { b.ctr[c] = b.hi[c]; b.ctr[c] += b.lo[c]; b.ctr[c] *= 0.5f; }
        b.tri = (int32_t)j;
        work.push_back(b);
    }
    Builder bld;
    int root = bld.recurse(work, 0);
    bld.pool[root].box = all;

    // CreateCFBVH + PopulateCacheFriendlyBVH, Raytracer.cc:651-718 (iterative DFS pre-order)
    nodes.assign(bld.pool.size(), b200r_bvhnode());
    tri_idx.clear();
    int maxDepth = 0;
    // The pool was filled in DFS pre-order already (node allocated before its children, left subtree
    // before right), so pool index == flattened index; assert that while emitting.
    struct Item { int node, depth; };
    std::vector<Item> stack; stack.push_back({root, 0});
    uint32_t next = 0;
    while (!stack.empty()) {
        Item it = stack.back(); stack.pop_back();
        const TmpNode& n = bld.pool[it.node];
        if ((uint32_t)it.node != next) throw std::runtime_error("internal: BVH pool is not in DFS order");
        b200r_bvhnode& o = nodes[next++];
        for (int c = 0; c < 3; c++) { o.lo[c] = n.box.lo[c]; o.hi[c] = n.box.hi[c]; }
        if (it.depth > maxDepth) maxDepth = it.depth;
        if (n.left >= 0) {
            o.a = (uint32_t)n.left; o.b = (uint32_t)n.right;
            stack.push_back({n.right, it.depth + 1});
            stack.push_back({n.left, it.depth + 1});
        } else {
            o.a = 0x80000000u | (uint32_t)n.tris.size();
            o.b = (uint32_t)tri_idx.size();
            for (int32_t t : n.tris) tri_idx.push_back(t);
        }
    }
    bvh_depth = maxDepth;
    if (maxDepth >= B200R_BVH_STACK_SIZE)
        throw std::runtime_error("Max depth of BVH exceeds BVH_STACK_SIZE");
}

static int depth_of(const std::vector<b200r_bvhnode>& nodes)
{
    if (nodes.empty()) return -1;
    int maxDepth = 0;
    std::vector<std::pair<uint32_t, int>> st; st.push_back({0u, 0});
    while (!st.empty()) {
        auto [i, d] = st.back(); st.pop_back();
        if (d > maxDepth) maxDepth = d;
        if (i >= nodes.size() || d > 4096) return 1 << 20;
        if (!(nodes[i].a & 0x80000000u)) { st.push_back({nodes[i].b, d + 1}); st.push_back({nodes[i].a, d + 1}); }
    }
    return maxDepth;
}

bool Scene::read_bvh_cache(const char* path)
{
    FILE* fp = fopen(path, "rb");
    if (!fp) return false;
    uint32_t nn = 0, ni = 0;
    bool ok = fread(&nn, 4, 1, fp) == 1 && fread(&ni, 4, 1, fp) == 1;
    std::vector<b200r_bvhnode> n; std::vector<int32_t> t;
    if (ok && nn > 0 && nn < (1u << 28) && ni < (1u << 28)) {
        n.resize(nn); t.resize(ni);
        ok = fread(n.data(), sizeof(b200r_bvhnode), nn, fp) == nn && (ni == 0 || fread(t.data(), 4, ni, fp) == ni);
    } else ok = false;
    fclose(fp);
    if (!ok) return false;
    // sanity: indices in range (the reference trusts the file; we do not)
    for (auto& x : n) {
        if (x.a & 0x80000000u) { if ((uint64_t)x.b + (x.a & 0x7fffffffu) > ni) return false; }
        else if (x.a >= nn || x.b >= nn) return false;
    }
    for (auto v : t) if (v < 0 || (size_t)v >= tris.size()) return false;
    int d = depth_of(n);
    if (d >= B200R_BVH_STACK_SIZE) return false;
    nodes.swap(n); tri_idx.swap(t); bvh_depth = d;
    return true;
}

bool Scene::write_bvh_cache(const char* path) const
{
    FILE* fp = fopen(path, "wb");
    if (!fp) return false;
    uint32_t nn = (uint32_t)nodes.size(), ni = (uint32_t)tri_idx.size();
    bool ok = fwrite(&nn, 4, 1, fp) == 1 && fwrite(&ni, 4, 1, fp) == 1 &&
              fwrite(nodes.data(), sizeof(b200r_bvhnode), nn, fp) == nn &&
              (ni == 0 || fwrite(tri_idx.data(), 4, ni, fp) == ni);
    fclose(fp);
    return ok;
}

void Scene::build_bvh(const char* cache_path, bool force_rebuild)
{
    if (!force_rebuild && cache_path && read_bvh_cache(cache_path)) return;
    build_bvh_from_scratch();
    if (cache_path) write_bvh_cache(cache_path);   // silently ignored on failure, like the reference
}

}  // namespace b200r
