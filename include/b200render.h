/* b200render.h — C-ABI of the B200-native renderer hot path (libb200render.so).
 *
 * The reference (ttsiodras/renderer) has no plugin/FFI interface; the seam this library replaces
 * is the set of Scene::render* member functions dispatched from main()'s switch
 * (reference src/renderer.cc:522-583, declared src/Scene.h:76-85), plus the two satellites the
 * frame depends on: Light::RenderSceneIntoShadowBuffer (src/Light.h:61) and MLAA() (src/MLAA.h:4).
 *
 * Everything here is extern "C", plain pointers and sizes. Every function returns 0 on success or a
 * negative B200R_E* code; b200r_last_error() gives the text. No function ever falls back to a CPU
 * renderer: without a usable CUDA device b200r_init() fails.
 *
 * Two groups:
 *   (1) device path  — b200r_init / upload / render / counters / destroy
 *   (2) host plumbing — scene loading, BVH build + .bvh cache, camera orbit, light matrices: the
 *       load-time numerics of the reference (Loader.cc, BVH.cc, Camera.cc, Light.cc, renderer.cc)
 *       that DEFINE the inputs of the hot path. Plain C++ on the CPU, as in the reference.
 */
#ifndef B200RENDER_H
#define B200RENDER_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- error codes */
#define B200R_OK            0
#define B200R_EINVAL       -1   /* bad argument */
#define B200R_ECUDA        -2   /* CUDA runtime error (text in b200r_last_error) */
#define B200R_ENODEVICE    -3   /* no CUDA device / device is not sm_100 */
#define B200R_EIO          -4   /* file not found / malformed (reference THROWs, Exceptions.h:28-33) */
#define B200R_ESTATE       -5   /* call sequence error (e.g. render before upload) */
#define B200R_EDEPTH       -6   /* BVH deeper than B200R_BVH_STACK_SIZE (reference exit(1), Raytracer.cc:711-717) */
#define B200R_ENOMEM       -7

#define B200R_BVH_STACK_SIZE 32     /* reference src/Defines.h:36 */
#define B200R_SHADOWMAP_SIZE 1024   /* reference src/Defines.h:25 */
#define B200R_MAX_LIGHTS     2      /* reference src/renderer.cc:277-296 (-w adds the second) */
#define B200R_MAX_FRAMES_IN_FLIGHT 16 /* independent sets of per-frame scratch buffers (b200r_set_pipeline_depth, b200r_render_device_slot) */

/* ---------------------------------------------------------------- scene records (POD)
 * These mirror the fields of the reference's Vertex / Triangle / CacheFriendlyBVHNode that the
 * hot path reads (src/Base3d.h:27-66, src/BVH.h:52-65). */

typedef struct b200r_vertex {      /* = reference Vertex: Vector3 + _normal + _ambientOcclusionCoeff */
    float    pos[3];
    float    nrm[3];
    uint32_t ao;
} b200r_vertex;

typedef struct b200r_tri {
    uint32_t a, b, c;              /* indices into the vertex array (reference keeps pointers) */
    float    center[3];            /* Triangle::_center  (ctor value, rescaled; Loader.cc:451-454) */
    float    normal[3];            /* Triangle::_normal  (plane normal, Loader.cc:476-481) */
    float    colorf[3];            /* Triangle::_colorf as r,g,b in 0..255 */
    uint32_t color;                /* Triangle::_color = 0x00RRGGBB (Base3d.cc:41) */
    uint32_t two_sided;            /* Triangle::_twoSided */
    float    d, d1, d2, d3;        /* plane / edge-plane offsets (Loader.cc:482-492) */
    float    e1[3], e2[3], e3[3];  /* edge-plane normals */
} b200r_tri;

typedef struct b200r_bvhnode {     /* byte-identical to CacheFriendlyBVHNode (32 B) */
    float    lo[3], hi[3];
    uint32_t a;                    /* inner: idxLeft          | leaf: 0x80000000 | count   */
    uint32_t b;                    /* inner: idxRight         | leaf: start in triIdx list  */
} b200r_bvhnode;

/* ---------------------------------------------------------------- per-frame state */

typedef struct b200r_light {
    float pos[3];                  /* world position (Light is-a Vector3) */
    float in_camera[3];            /* Light::_inCameraSpace        (Light.cc:162-171), modes >= 5 */
    float cam2light[9];            /* Light::_cameraToLightSpace   (Light.cc:194-216), modes >= 7 */
} b200r_light;

/* render modes = reference RenderMode enum (src/renderer.cc:69-80) */
enum {
    B200R_MODE_POINTS = 1, B200R_MODE_POINTS_TRI = 2, B200R_MODE_LINES = 3, B200R_MODE_AMBIENT = 4,
    B200R_MODE_GOURAUD = 5, B200R_MODE_PHONG = 6, B200R_MODE_PHONG_SHADOWMAPS = 7,
    B200R_MODE_PHONG_SOFTSHADOWMAPS = 8, B200R_MODE_RAYTRACE = 9, B200R_MODE_RAYTRACE_AA = 10
};

/* flags: the reference's compile-time switches at the top of src/Raytracer.cc:53-81, made runtime */
#define B200R_F_SHADOWS      0x01u   /* USE_SHADOWS        (default on)  */
#define B200R_F_REFLECTIONS  0x02u   /* REFLECTIONS        (default on)  */
#define B200R_F_PHONG_NORMAL 0x04u   /* USE_PHONG_NORMAL   (default on)  */
#define B200R_F_AO           0x08u   /* AMBIENT_OCCLUSION  (default off) */
#define B200R_F_MLAA         0x10u   /* --enable-mlaa post filter (Screen.h:130-137) */
#define B200R_F_DEFAULT      (B200R_F_SHADOWS | B200R_F_REFLECTIONS | B200R_F_PHONG_NORMAL)

typedef struct b200r_frame {
    uint32_t   mode;               /* 1..10 */
    uint32_t   width, height;      /* reference: compile-time WIDTH/HEIGHT (Defines.h:26-27) */
    float      eye[3];             /* Camera position */
    float      mv[9];              /* Camera::_mv rows {up, right, forward} (Camera.cc:39-41) */
    uint32_t   n_lights;           /* 1 or 2 */
    b200r_light lights[B200R_MAX_LIGHTS];
    uint32_t   flags;              /* B200R_F_* */
    uint32_t   ao_samples;         /* AMBIENT_SAMPLES (Raytracer.cc:79) */
    uint32_t   max_depth;          /* MAX_RAY_DEPTH   (Raytracer.cc:56), 0 -> 3 */
    uint32_t   frame_index;        /* keys the deterministic AO random stream */
    uint32_t   row_first;          /* multi-GPU row-cyclic sharding: this call renders rows   */
    uint32_t   row_step;           /*   y = row_first + k*row_step (0/1 -> whole frame)       */
} b200r_frame;

/* Work counters of the last frame (for the algorithmic-bytes roofline).
 * Only filled when the frame was rendered with counters enabled (b200r_set_counters). */
typedef struct b200r_counters {
    uint64_t rays_primary, rays_shadow, rays_reflection, rays_ao;
    uint64_t node_tests;           /* RayIntersectsBox calls (inner nodes popped) */
    uint64_t leaf_visits;
    uint64_t tri_tests;            /* triangles reaching the plane test */
    uint64_t tris_setup, spans, z_tests, z_passes;   /* rasteriser */
} b200r_counters;

typedef struct b200r_ctx b200r_ctx;

/* ---------------------------------------------------------------- (1) device path */

/* Replaces: Screen ctor + first use (reference src/Screen.h:59-113). device = CUDA ordinal. */
int  b200r_init(int device, b200r_ctx** out);
void b200r_destroy(b200r_ctx* ctx);
const char* b200r_last_error(const b200r_ctx* ctx);   /* ctx may be NULL: last global error */

/* Replaces: the Scene the render functions read (src/Scene.h:35-47). Uploaded ONCE; re-laid-out
 * on the device (SoA float4 records, leaf triangles baked in triIdx order). nodes may be NULL/0 for
 * raster-only use. */
int  b200r_upload_scene(b200r_ctx* ctx,
                        const b200r_vertex* verts, uint32_t n_verts,
                        const b200r_tri* tris, uint32_t n_tris,
                        const b200r_bvhnode* nodes, uint32_t n_nodes,
                        const int32_t* tri_idx, uint32_t n_tri_idx);

/* Replaces: Light::_shadowBuffer produced by Light::RenderSceneIntoShadowBuffer (src/Light.cc:218-244).
 * Either upload a host-built 1024x1024 map ... */
int  b200r_upload_shadowmap(b200r_ctx* ctx, int light, const float* map1024x1024);
/* ... or have the device render it from the uploaded scene (light world position + world->light rows). */
int  b200r_render_shadowmap(b200r_ctx* ctx, int light, const float light_pos[3], const float world2light[9]);
int  b200r_download_shadowmap(b200r_ctx* ctx, int light, float* map1024x1024);

/* Replaces: Scene::renderPoints/renderWireframe/renderAmbient/renderGouraud/renderPhong/
 * renderPhongAndShadowed/renderPhongAndSoftShadowed/renderRaytracer (src/Scene.h:76-85), selected by
 * f->mode exactly like main()'s switch (src/renderer.cc:522-583), including Screen::ShowScreen's MLAA
 * hook when B200R_F_MLAA is set. host_xrgb receives the frame as the reference's surface words
 * 0x00RRGGBB, pitch = width (rows_rendered x width words when row sharding is used). */
int  b200r_render(b200r_ctx* ctx, const b200r_frame* f, uint32_t* host_xrgb);

/* The same call, pipelined - for loops like main()'s benchmark loop (src/renderer.cc:491-606), where frame i+1 does
 * not depend on frame i: the call returns once the frame is ENQUEUED. Up to depth+1 frames are in flight (depth = 2
 * unless b200r_set_pipeline_depth changed it): one being copied to the host while the next `depth` render; ray-traced
 * frames rotate over `depth` streams and scratch sets, so the head of frame i+1 runs on the SMs that the tail of frame i
 * (its last few long rays) leaves idle.
 * host_xrgb must stay valid, and is only complete, after b200r_wait() - or after the (depth+1)-th following
 * b200r_render_async(), which reuses the slot. Page-locked caller memory is written by
 * DMA directly; pageable memory goes through an internal pinned staging buffer (copied out in b200r_wait / slot reuse).
 * b200r_render() and b200r_destroy() drain the pipeline first. */
int  b200r_render_async(b200r_ctx* ctx, const b200r_frame* f, uint32_t* host_xrgb);
int  b200r_wait(b200r_ctx* ctx);
/* Frames of b200r_render_async that render concurrently: 1 (none overlap) .. B200R_MAX_FRAMES_IN_FLIGHT. Drains the
 * pipeline first. */
int  b200r_set_pipeline_depth(b200r_ctx* ctx, uint32_t depth);

/* Same, but the frame stays in device memory (dev_xrgb is a CUDA device pointer with room for
 * rows_rendered*width words) and the work is enqueued on `cuda_stream` (a cudaStream_t, NULL = the
 * library's own stream, which is then synchronised before returning). Used by the multi-GPU path,
 * where the packed rows feed an NCCL all-gather. */
int  b200r_render_device(b200r_ctx* ctx, const b200r_frame* f, void* dev_xrgb, void* cuda_stream);

/* b200r_render_device for callers that keep several frames in flight themselves (the multi-GPU pipeline: rank r renders
 * its rows of frame i+1 while the all-gather of frame i is on the wire): `scratch_slot` (0 .. B200R_MAX_FRAMES_IN_FLIGHT-1)
 * selects the set of per-frame scratch buffers (job queue, merge words, hit records). Two frames may be in flight at the
 * same time iff they use different slots; frames of the same slot must be ordered by the caller (same stream, or events).
 * `cuda_stream` must not be NULL; nothing is synchronised. Every mode but 3 (the wireframe pass sizes its fragment buffer on the
 * host). The rasteriser's span buffer is sized up front and checked behind the frame: should a frame ever overflow it, the NEXT
 * call on this context returns B200R_ENOMEM (the buffer has been enlarged by then; submit the frames again). */
int  b200r_render_device_slot(b200r_ctx* ctx, const b200r_frame* f, void* dev_xrgb, void* cuda_stream, uint32_t scratch_slot);

/* Replaces: MLAA(fbi, NULL, resX, resY) (src/MLAA.h:4) applied in place on a full device frame. */
int  b200r_mlaa_device(b200r_ctx* ctx, void* dev_xrgb, uint32_t width, uint32_t height, void* cuda_stream);

/* Multi-GPU helper: scatter P packed row-cyclic shards (as all-gathered: shard r holds rows r, r+P, ...)
 * into a scan-order frame. */
int  b200r_deinterleave_device(b200r_ctx* ctx, const void* dev_gathered, void* dev_frame,
                               uint32_t width, uint32_t height, uint32_t n_shards, void* cuda_stream);

/* ---------------------------------------------------------------- (1b) frames in flight / several GPUs
 * One b200r_pipeline per rank - one process (or one thread) per GPU, each with its own b200r_ctx and the same scene uploaded.
 * Replaces the body of main()'s benchmark loop (reference src/renderer.cc:491-606) when frames do not depend on each other:
 * submit() returns as soon as the frame is ENQUEUED; up to `depth` frames are in flight per rank (slot = submission index % depth).
 * With world > 1 rank r renders rows r, r+world, ... of every frame and the rows are assembled into a
 * scan-order frame on EVERY rank, either by one ncclAllGather + a de-interleave kernel (B200R_ASSEMBLE_NCCL) or by stores
 * through NVLink peer mappings with one arrival flag per frame (B200R_ASSEMBLE_PUSH); B200R_F_MLAA is applied to the assembled
 * frame. Every rank must submit the same frames in the same order. Every mode but 3 (b200r_render_device_slot).
 * NCCL is loaded at run time and only when world > 1. */
#define B200R_ASSEMBLE_NCCL 0
#define B200R_ASSEMBLE_PUSH 1
typedef struct b200r_pipeline b200r_pipeline;
/* 128 bytes identifying the job: created on rank 0 (ncclGetUniqueId) and handed to every rank by the launcher. */
int  b200r_dist_unique_id(void* out128);
int  b200r_pipeline_create(b200r_ctx* ctx, uint32_t width, uint32_t height, uint32_t depth, uint32_t rank, uint32_t world,
                           const void* nccl_unique_id /* NULL iff world == 1 */, uint32_t assemble, b200r_pipeline** out);
/* host_xrgb: page-locked host memory that receives the ASSEMBLED frame (complete after b200r_pipeline_drain, or once `depth`
 * further frames were submitted and drained past it), or NULL to leave the frame on the device (b200r_pipeline_slot_frame). */
int  b200r_pipeline_submit(b200r_pipeline* pipe, const b200r_frame* f, uint32_t* host_xrgb);
int  b200r_pipeline_drain(b200r_pipeline* pipe);                       /* host waits for everything submitted on THIS rank */
int  b200r_pipeline_slot_frame(b200r_pipeline* pipe, uint32_t slot, void** dev_xrgb);
/* Order the pipeline against a caller's stream: pipeline_waits != 0 - nothing submitted from now on starts before the stream's
 * current tail; == 0 - the stream waits for everything submitted so far (device side; the host does not block). */
int  b200r_pipeline_fence(b200r_pipeline* pipe, void* cuda_stream, int pipeline_waits);
/* Measurement aids (bench.py): write `bytes` (> L2) before every frame on the frame's stream, then prefetch up to four device
 * buffers (the scene) back into L2. 0 bytes switches it off. */
int  b200r_pipeline_set_l2_flush(b200r_pipeline* pipe, uint64_t bytes);
int  b200r_pipeline_set_prefetch(b200r_pipeline* pipe, uint32_t index, const void* dev_ptr, uint64_t bytes);
int  b200r_pipeline_launches(b200r_pipeline* pipe, uint32_t* n_kernel_launches, int reset);
/* Timing on: CUDA events around this rank's render kernels of every submitted frame (on the frame's own stream);
 * b200r_pipeline_kernel_ms waits for them and returns the sum of the durations and the number of frames since the last call. */
int  b200r_pipeline_set_timing(b200r_pipeline* pipe, int enabled);
int  b200r_pipeline_kernel_ms(b200r_pipeline* pipe, double* sum_ms, uint32_t* n_frames);
const char* b200r_pipeline_last_error(const b200r_pipeline* pipe);
/* All ranks must have drained (and agreed on it, e.g. a barrier) before any rank destroys its pipeline. */
void b200r_pipeline_destroy(b200r_pipeline* pipe);
/* Page-locked host memory for frames (b200r_pipeline_submit / b200r_render_async write into it by DMA). */
int  b200r_host_alloc(uint64_t bytes, void** out);
void b200r_host_free(void* p);
/* Device addresses / sizes of the uploaded scene's traversal buffers (for b200r_pipeline_set_prefetch): index 0 nodes, 1 leaf
 * triangle records, 2 shading records. Returns B200R_EINVAL past the last one. */
int  b200r_scene_buffer(b200r_ctx* ctx, uint32_t index, const void** dev_ptr, uint64_t* bytes);

/* Replaces: CreateBVH + CreateCFBVH (src/BVH.cc:96-371 scalar path, src/Raytracer.cc:651-718) ON THE DEVICE: the SAH
 * build as level-synchronous kernels (csrc/bvh_steps.h, csrc/cuda/bvh_build.cu) producing the same tree bit for bit - the
 * nodes/tri_idx written here are byte-identical to the reference's .bvh cache content. nodes_out needs room for
 * 2*n_tris + 1 nodes, tri_idx_out for n_tris indices. Does not change the uploaded scene: pass the result to
 * b200r_upload_scene (and/or write it as a cache). Fails if the tree is deeper than the reference's BVH_STACK_SIZE. */
int  b200r_build_bvh(b200r_ctx* ctx, const b200r_vertex* verts, uint32_t n_verts, const b200r_tri* tris, uint32_t n_tris,
                     b200r_bvhnode* nodes_out, uint32_t nodes_cap, int32_t* tri_idx_out, uint32_t* n_nodes, int32_t* depth);
/* Test hook (no device needed): the very same per-item step functions run in plain loops on the host. Not a build path. */
int  b200r_selftest_bvh_steps_host(const b200r_vertex* verts, uint32_t n_verts, const b200r_tri* tris, uint32_t n_tris,
                                   b200r_bvhnode* nodes_out, uint32_t nodes_cap, int32_t* tri_idx_out,
                                   uint32_t* n_nodes, int32_t* depth);
/* Test hook (no device needed): the MLAA step functions the device kernels are made of (csrc/mlaa_steps.h: flags, line
 * bounds, split heights, in-place blends - `batched` != 0: with the 8-pixel batched loads the device uses) run in plain
 * loops on the host, in place on a width x height frame. Not a rendering path. */
int  b200r_selftest_mlaa_steps_host(uint32_t* frame_xrgb, uint32_t width, uint32_t height, int batched);
/* Test hook (no device needed): the span walkers of csrc/raster_steps.h (what the depth / resolve passes run per span,
 * pixel by pixel and with 8 depth keys requested together) over n_spans random spans against the per-scanline loop of
 * Screen::RasterizeTriangle (src/Screen.h:244-289) as written there: *mismatches = spans whose pixel sequence (x, interpolant
 * bits, key) differs. Not a rendering path. */
int  b200r_selftest_span_walk_host(uint32_t seed, uint32_t n_spans, uint32_t width, uint64_t* mismatches);


/* Numerics self-test: the slab test's shared-reciprocal divide (DESIGN.md "division") against the compiler's IEEE
 * divide on `samples` random operand pairs drawn from the whole domain in which the fast path is used.
 * *mismatches must come back 0. first_bad (optional) receives {a, d, a/d, fast} of the first mismatch. */
int  b200r_selftest_division(b200r_ctx* ctx, uint64_t samples, uint32_t seed, uint64_t* mismatches, float first_bad[4]);

/* Developer tool: per-tile (8x4 pixels) start/end device timestamps (ns, %globaltimer) of the last ray-traced frame;
 * tile t = (row of tiles)*ceil(W/8) + column. Costs time; off by default. */
int  b200r_set_tile_profile(b200r_ctx* ctx, int enabled);
int  b200r_get_tile_profile(b200r_ctx* ctx, uint64_t* start_end_ns, uint32_t max_tiles, uint32_t* n_tiles);

/* Developer switches: select cross-check variants of the kernels (parity tests, A/B measurements); none changes a result.
 * Names: monolithic_rt, no_prune, no_fuse, no_wavefront, rt_legacy, no_root_rect, pool_small, split_depth, raster_inline_shade, mlaa_scan,
 * mlaa_fullscan, mlaa_nobatch, mlaa_no_tma, no_frame_overlap, bvh_serial_split, pool_stats, pool_policy, pool_scatter, pool_occ3, pool_tiles_per_warp, pool_cta_warps,
 * pool_leaf_min, pool_sort_min, pool_shade_min, pool_refill_min, pool_low_water, pool_dry (csrc/cuda/rt_kernels.cuh `Switches` documents each).
 * Defaults come from the environment variables B200R_<NAME IN CAPITALS>, read once by b200r_init. Waits for frames in flight. */
int  b200r_set_switch(b200r_ctx* ctx, const char* name, int value);
int  b200r_set_counters(b200r_ctx* ctx, int enabled);   /* counting costs time; off by default */
int  b200r_get_counters(b200r_ctx* ctx, b200r_counters* out);
/* Device time (ms, CUDA events on the launching stream) of the kernels of the last frame. */
int  b200r_last_kernel_ms(b200r_ctx* ctx, float* total_ms, float* dominant_ms);
int  b200r_last_launches(b200r_ctx* ctx, uint32_t* n_kernel_launches);

/* ---------------------------------------------------------------- (2) host plumbing */

typedef struct b200r_scene b200r_scene;

/* Replaces Scene::load (src/Loader.cc:85-494): .tri and shadevis .ply, fix_normals, recentre/rescale
 * to +-1.2, per-triangle intersection precompute — bit-for-bit the reference's numerics. */
int  b200r_scene_load(const char* filename, b200r_scene** out);
void b200r_scene_free(b200r_scene* s);
/* Replaces Scene::UpdateBoundingVolumeHierarchy (src/Raytracer.cc:720-789): reads cache_path if it is a
 * valid .bvh cache, else builds the SAH BVH (same tree as src/BVH.cc:96-371 scalar path), flattens it
 * (Raytracer.cc:651-718) and writes the cache (same on-disk format). cache_path may be NULL. */
int  b200r_scene_build_bvh(b200r_scene* s, const char* cache_path, int force_rebuild);

/* The same, with the build on the device (b200r_build_bvh) instead of the host: identical tree, identical cache file. */
int  b200r_scene_build_bvh_device(b200r_scene* s, b200r_ctx* ctx, const char* cache_path, int force_rebuild);
const b200r_vertex*  b200r_scene_vertices(const b200r_scene* s, uint32_t* n);
const b200r_tri*     b200r_scene_tris(const b200r_scene* s, uint32_t* n);
const b200r_bvhnode* b200r_scene_nodes(const b200r_scene* s, uint32_t* n);
const int32_t*       b200r_scene_tri_idx(const b200r_scene* s, uint32_t* n);
int                  b200r_scene_bvh_depth(const b200r_scene* s);
/* Number of triangles whose precomputed edge planes do NOT bound the triangle to within `tol` (fp64 check). The ray
 * tracer prunes subtrees by distance only for scenes where this is 0 at tol = 2e-5 (DESIGN.md "distance pruning"). */
uint32_t             b200r_scene_unbounded_triangles(const b200r_scene* s, double tol);
int  b200r_upload_scene_handle(b200r_ctx* ctx, const b200r_scene* s);   /* convenience */

/* Camera::set + UpdateMV (src/Camera.cc:24-42). */
void b200r_camera_look_at(const float eye[3], const float lookat[3], float mv_out[9]);

/* The benchmark orbit of main() (src/renderer.cc:300-316, 485-496): an fp32 recurrence that must be
 * iterated on the host with the same libm calls. b200r_orbit_init sets the initial eye/angles,
 * b200r_orbit_step advances one frame and returns eye + mv of that frame. */
typedef struct b200r_orbit { float eye[3]; float angle1, angle2, d_angle; } b200r_orbit;
void b200r_orbit_init(b200r_orbit* o);
void b200r_orbit_step(b200r_orbit* o, float eye_out[3], float mv_out[9]);

/* Lights as main() places them (src/renderer.cc:277-296): index 0 = the rotating light at angle 45deg,
 * index 1 = the static second light of "-w". */
void b200r_default_light_pos(int index, float pos_out[3]);
/* Light::CalculatePositionInCameraSpace / CalculateXformFromCameraToLightSpace /
 * CalculateXformFromWorldToLightSpace (src/Light.cc:162-216). */
void b200r_light_in_camera_space(const float light_pos[3], const float eye[3], const float mv[9], float out[3]);
void b200r_light_camera_to_light(const float light_pos[3], const float mv[9], float out[9]);
void b200r_light_world_to_light(const float light_pos[3], float out[9]);

/* Fills a b200r_frame exactly as one iteration of main()'s loop would for the given mode
 * (camera from eye/mv, lights from n_lights defaults, per-mode light matrices). */
void b200r_frame_defaults(b200r_frame* f, uint32_t mode, uint32_t width, uint32_t height,
                          const float eye[3], const float mv[9], uint32_t n_lights);

const char* b200r_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200RENDER_H */
