// api_host.cpp — extern "C" wrappers of the host plumbing (scene load, BVH, accessors).
// Errors: the reference THROWs std::string and exits from main (src/Exceptions.h:28-33,
// src/renderer.cc:637-640); here every failure becomes a negative status + b200r_last_error() text.
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>

#include "scene.h"

namespace b200r {
uint32_t count_unbounded_triangles(const b200r_vertex*, uint32_t, const b200r_tri*, uint32_t, double, unsigned char*);
static std::mutex g_err_mu;
static std::string g_last_error;
void set_global_error(const std::string& s) { std::lock_guard<std::mutex> l(g_err_mu); g_last_error = s; }
const char* global_error() { return g_last_error.c_str(); }
}  // namespace b200r

struct b200r_scene { b200r::Scene s; };

extern "C" {

const char* b200r_version(void) { return "renderer_b200 0.1 (sm_100a)"; }

int b200r_scene_load(const char* filename, b200r_scene** out)
{
    if (!filename || !out) { b200r::set_global_error("b200r_scene_load: NULL argument"); return B200R_EINVAL; }
    *out = nullptr;
    b200r_scene* h = nullptr;
    try {
        h = new b200r_scene();
        h->s.load(filename);
        if (h->s.tris.empty()) throw std::runtime_error(std::string("no triangles in ") + filename);
    } catch (const std::exception& e) {
        delete h;
        b200r::set_global_error(e.what());
        return B200R_EIO;
    }
    *out = h;
    return B200R_OK;
}

void b200r_scene_free(b200r_scene* s) { delete s; }

int b200r_scene_build_bvh(b200r_scene* s, const char* cache_path, int force_rebuild)
{
    if (!s) { b200r::set_global_error("b200r_scene_build_bvh: NULL scene"); return B200R_EINVAL; }
    try {
        s->s.build_bvh(cache_path, force_rebuild != 0);
    } catch (const std::exception& e) {
        b200r::set_global_error(e.what());
        return B200R_EDEPTH;
    }
    return B200R_OK;
}

// Scene::UpdateBoundingVolumeHierarchy with the build itself on the device: cache hit -> as above; else b200r_build_bvh
// (CUDA kernels, same tree) fills the scene's node / index arrays and the cache file is written in the reference's format.
int b200r_scene_build_bvh_device(b200r_scene* s, b200r_ctx* ctx, const char* cache_path, int force_rebuild)
{
    if (!s || !ctx) { b200r::set_global_error("b200r_scene_build_bvh_device: NULL argument"); return B200R_EINVAL; }
    if (!force_rebuild && cache_path && s->s.read_bvh_cache(cache_path)) return B200R_OK;
    const uint32_t nt = (uint32_t)s->s.tris.size();
    std::vector<b200r_bvhnode> nodes(2 * (size_t)nt + 1);
    std::vector<int32_t> idx(nt);
    uint32_t nn = 0; int32_t depth = -1;
    const int rc = b200r_build_bvh(ctx, s->s.verts.data(), (uint32_t)s->s.verts.size(), s->s.tris.data(), nt, nodes.data(),
                                   (uint32_t)nodes.size(), idx.data(), &nn, &depth);
    if (rc) return rc;
    nodes.resize(nn);
    s->s.nodes.swap(nodes); s->s.tri_idx.swap(idx); s->s.bvh_depth = depth;
    if (cache_path) s->s.write_bvh_cache(cache_path);   // silently ignored on failure, like the reference
    return B200R_OK;
}

const b200r_vertex* b200r_scene_vertices(const b200r_scene* s, uint32_t* n)
{ if (n) *n = (uint32_t)s->s.verts.size(); return s->s.verts.data(); }
const b200r_tri* b200r_scene_tris(const b200r_scene* s, uint32_t* n)
{ if (n) *n = (uint32_t)s->s.tris.size(); return s->s.tris.data(); }
const b200r_bvhnode* b200r_scene_nodes(const b200r_scene* s, uint32_t* n)
{ if (n) *n = (uint32_t)s->s.nodes.size(); return s->s.nodes.data(); }
const int32_t* b200r_scene_tri_idx(const b200r_scene* s, uint32_t* n)
{ if (n) *n = (uint32_t)s->s.tri_idx.size(); return s->s.tri_idx.data(); }
int b200r_scene_bvh_depth(const b200r_scene* s) { return s->s.bvh_depth; }

uint32_t b200r_scene_unbounded_triangles(const b200r_scene* s, double tol)
{
    return b200r::count_unbounded_triangles(s->s.verts.data(), (uint32_t)s->s.verts.size(), s->s.tris.data(),
                                            (uint32_t)s->s.tris.size(), tol, nullptr);
}

}  // extern "C"

extern "C" int b200r_upload_scene_handle(b200r_ctx* ctx, const b200r_scene* s)
{
    if (!s) { b200r::set_global_error("b200r_upload_scene_handle: NULL scene"); return B200R_EINVAL; }
    const b200r::Scene& sc = s->s;
    return b200r_upload_scene(ctx, sc.verts.data(), (uint32_t)sc.verts.size(), sc.tris.data(), (uint32_t)sc.tris.size(),
                              sc.nodes.empty() ? nullptr : sc.nodes.data(), (uint32_t)sc.nodes.size(),
                              sc.tri_idx.empty() ? nullptr : sc.tri_idx.data(), (uint32_t)sc.tri_idx.size());
}
