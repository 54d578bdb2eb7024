// scene.h — host-side scene container (replaces the reference's Scene data members, src/Scene.h:32-47).
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../../include/b200render.h"
#include "../vec.h"

namespace b200r {

struct Scene {
    std::vector<b200r_vertex>  verts;
    std::vector<b200r_tri>     tris;
    std::vector<b200r_bvhnode> nodes;     // flattened BVH, DFS pre-order (CacheFriendlyBVHNode[])
    std::vector<int32_t>       tri_idx;   // _triIndexList
    int bvh_depth = -1;

    void load(const std::string& filename);          // Scene::load
    void fix_normals();                              // Scene::fix_normals
    void finish_load();                              // recentre/rescale + intersection precompute
    // Scene::UpdateBoundingVolumeHierarchy: cache read, else build + flatten + cache write
    void build_bvh(const char* cache_path, bool force_rebuild);
    void build_bvh_from_scratch();                   // CreateBVH + CreateCFBVH
    bool read_bvh_cache(const char* path);
    bool write_bvh_cache(const char* path) const;
};

}  // namespace b200r
