// context.cu — the device half of the C-ABI (include/b200render.h): context, scene upload, frame dispatch.
//
// b200r_render() stands where main()'s switch(mode) calls scene.renderXxx(sony, canvas)
// (reference src/renderer.cc:522-583). There is deliberately NO CPU rendering path in this library:
// if CUDA is unavailable b200r_init() fails and nothing can be rendered.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "device_types.cuh"
#include "rt_kernels.cuh"

namespace b200r {
void set_global_error(const std::string& s);
const char* global_error();
uint32_t count_unbounded_triangles(const b200r_vertex* verts, uint32_t n_verts, const b200r_tri* tris, uint32_t n_tris, double tol,
                                   unsigned char* bad_out);
}

using namespace b200r;

struct b200r_ctx;
namespace b200r {
int ctx_device(const b200r_ctx* ctx);
int ctx_sms(const b200r_ctx* ctx);
}

struct b200r_ctx {
    int device = -1;
    int numSMs = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::string err;

    // scene
    float4 *d_nodes = nullptr, *d_leaftris = nullptr, *d_shade = nullptr, *d_rverts = nullptr, *d_rtris = nullptr;
    float* d_shadowmap[B200R_MAX_LIGHTS] = {nullptr, nullptr};
    DeviceScene sc{};
    bool have_scene = false, have_bvh = false;
    size_t nodesBytes = 0, leafBytes = 0, shadeBytes = 0;

    // frame resources
    uint32_t* d_frame = nullptr; size_t frame_words = 0;
    uint32_t* h_pinned = nullptr; size_t pinned_words = 0;
    // b200r_render_async: two frames in flight - frame i's device->host copy runs on copyStream while frame i+1 renders
    struct AsyncSlot {
        uint32_t* d = nullptr; size_t words = 0;              // device frame
        uint32_t* staging = nullptr; size_t staging_words = 0; // pinned staging, used when the caller's buffer is pageable
        uint32_t* user = nullptr; size_t user_words = 0;       // where the frame finally goes
        cudaEvent_t rendered = nullptr, copied = nullptr;
        bool inflight = false, staged = false;
        b200r_frame frame{}; int set = 0;                       // what was submitted (a rasterised frame that overflowed its span buffer is redone)
    } slot[B200R_MAX_FRAMES_IN_FLIGHT + 1];
    // frames in flight of b200r_render_async: `depth` rendering (overlapped, see rstream) + one being copied out
    unsigned depth = 2;
    unsigned nslot() const { return depth + 1; }
    cudaStream_t copyStream = nullptr;
    // ... and consecutive ray-traced frames run on different streams, each with its own set of scratch buffers (rts[k],
    // tileCounters[k]), so the head of one frame fills the SMs that the tail of the previous frame's persistent kernel
    // leaves idle (its last few long rays). rts[0] / stream are the set every blocking call uses.
    cudaStream_t rstream[B200R_MAX_FRAMES_IN_FLIGHT] = {};      // [0] is never created: set 0 renders on `stream`
    RtBuffers rts[B200R_MAX_FRAMES_IN_FLIGHT] = {};
    unsigned* tileCounters[B200R_MAX_FRAMES_IN_FLIGHT] = {};    // [0] == d_tileCounter
    unsigned asyncIdx = 0;
    unsigned* d_tileCounter = nullptr;
    DeviceCounters* d_ctr = nullptr;
    Switches sw{};                                              // developer switches (b200r_set_switch / environment at b200r_init)
    // rasteriser / MLAA scratch, one set per scratch slot (frames in flight); set 0 serves the blocking calls
    struct RasterSet {
        RasterBuffers rb{};
        size_t zkey_pixels = 0, attr_pixels = 0;
        float4* d_attrs = nullptr;
        unsigned* h_spanCount = nullptr;          // pinned: the frame's span count lands here behind the frame, without a sync
        cudaEvent_t counted = nullptr;            // recorded behind that copy
        bool pending = false;                     // a frame whose span count has not been looked at yet
        uint32_t* d_mlaaScratch = nullptr; size_t mlaaWords = 0;
        void* d_mlaaLines = nullptr; size_t mlaaLinesBytes = 0;
    } rs[B200R_MAX_FRAMES_IN_FLIGHT];
    unsigned spanCapacityWanted = 0;              // grows when a frame overflowed
    bool framesInFlight = false;                  // the frame being enqueued is one of several in flight (set by the slot / async calls)
    bool spanOverflowSticky = false;              // an un-retried frame overflowed its span buffer (device / slot calls)
    WireBuffers wb{};
    unsigned* h_spanCount = nullptr;          // pinned (mode 3: fragment count)
    unsigned long long* d_tileProf = nullptr; size_t tileProfTiles = 0; bool tileProfile = false; uint32_t lastTiles = 0;
    unsigned* d_shadowKeys = nullptr;
    bool counting = false;
    float last_total_ms = 0.f, last_dominant_ms = 0.f;
    uint32_t last_launches = 0;
};

int b200r::ctx_device(const b200r_ctx* ctx) { return ctx->device; }
int b200r::ctx_sms(const b200r_ctx* ctx) { return ctx->numSMs; }

namespace {

int fail(b200r_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg;
    set_global_error(msg);
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, B200R_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));   \
    } while (0)

template <class T>
cudaError_t upload(T** dptr, const std::vector<T>& h, cudaStream_t s)
{
    if (*dptr) { cudaFree(*dptr); *dptr = nullptr; }
    if (h.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)dptr, h.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s);
}

inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

int make_frame_params(b200r_ctx* ctx, const b200r_frame* f, FrameParams& fp)
{
    if (!f) return fail(ctx, B200R_EINVAL, "NULL frame");
    memset(&fp, 0, sizeof fp);
    fp.mode = f->mode == 0 ? (uint32_t)B200R_MODE_RAYTRACE_AA : f->mode;
    if (fp.mode < 1 || fp.mode > 10) return fail(ctx, B200R_EINVAL, "mode must be 1..10");
    if (f->width == 0 || f->height == 0 || f->width > 16384 || f->height > 16384)
        return fail(ctx, B200R_EINVAL, "width/height out of range");
    if (f->n_lights < 1 || f->n_lights > B200R_MAX_LIGHTS) return fail(ctx, B200R_EINVAL, "n_lights must be 1 or 2");
    fp.W = f->width; fp.H = f->height;
    fp.n_lights = f->n_lights; fp.flags = f->flags; fp.ao_samples = f->ao_samples;
    fp.max_depth = f->max_depth ? f->max_depth : 3;
    if (fp.max_depth > 8) return fail(ctx, B200R_EINVAL, "max_depth > 8");
    if ((fp.flags & B200R_F_AO) && (fp.ao_samples == 0 || fp.ao_samples > 4096))
        return fail(ctx, B200R_EINVAL, "ao_samples must be 1..4096 when AO is on");
    fp.frame_index = f->frame_index;
    fp.row_step = f->row_step ? f->row_step : 1;
    fp.row_first = f->row_first;
    if (fp.row_first >= fp.row_step && !(fp.row_step == 1 && fp.row_first == 0))
        return fail(ctx, B200R_EINVAL, "row_first must be < row_step");
    fp.n_rows = (fp.H - fp.row_first + fp.row_step - 1) / fp.row_step;
    memcpy(fp.eye, f->eye, 12); memcpy(fp.mv, f->mv, 36);
    for (uint32_t i = 0; i < f->n_lights; i++) {
        memcpy(fp.light_pos[i], f->lights[i].pos, 12);
        memcpy(fp.light_cam[i], f->lights[i].in_camera, 12);
        memcpy(fp.cam2light[i], f->lights[i].cam2light, 36);
    }
    return B200R_OK;
}

int ensure_frame(b200r_ctx* ctx, size_t words)
{
    if (ctx->frame_words < words) {
        if (ctx->d_frame) cudaFree(ctx->d_frame);
        ctx->d_frame = nullptr; ctx->frame_words = 0;
        CU(cudaMalloc((void**)&ctx->d_frame, words * 4));
        ctx->frame_words = words;
    }
    return B200R_OK;
}

int mlaa_on(b200r_ctx* ctx, uint32_t* d_frame, uint32_t width, uint32_t height, cudaStream_t s, int scratchSet = 0)
{
    b200r_ctx::RasterSet& R = ctx->rs[scratchSet];
    if ((width % 4) || (height % 8))
        return fail(ctx, B200R_EINVAL, "MLAA needs width % 4 == 0 and height % 8 == 0 (the reference's SSE code assumes it, MLAA.cc:396,453)");
    if (reinterpret_cast<uintptr_t>(d_frame) & 15u)
        return fail(ctx, B200R_EINVAL, "MLAA needs a 16-byte aligned frame (so does the reference's SSE code, MLAA.cc:453-457)");
    const size_t words = (size_t)width * height;
    if (R.mlaaWords < words) {
        if (R.d_mlaaScratch) cudaFree(R.d_mlaaScratch);
        R.d_mlaaScratch = nullptr; R.mlaaWords = 0;
        CU(cudaMalloc((void**)&R.d_mlaaScratch, words * 4));
        R.mlaaWords = words;
    }
    int launches = 0;
    void* lines = nullptr;
    if (!ctx->sw.mlaa_scan) {           // default: two-stage blending (all separation lines first, then the ordered blends)
        const size_t need = mlaa_lines_bytes((int)width, (int)height);
        if (R.mlaaLinesBytes < need) {
            if (R.d_mlaaLines) cudaFree(R.d_mlaaLines);
            R.d_mlaaLines = nullptr; R.mlaaLinesBytes = 0;
            CU(cudaMalloc(&R.d_mlaaLines, need));
            R.mlaaLinesBytes = need;
        }
        lines = R.d_mlaaLines;
    }
    CU(launch_mlaa(d_frame, R.d_mlaaScratch, (int)width, (int)height, ctx->numSMs, s, launches, lines, ctx->sw));
    ctx->last_launches += (uint32_t)launches;
    return B200R_OK;
}

// Look at the span count of the last rasterised frame of a scratch set (waits for that frame's rasteriser part). Returns true
// if it overflowed its span buffer - the frame is incomplete; the next allocation is large enough.
bool span_overflowed(b200r_ctx* ctx, int set)
{
    b200r_ctx::RasterSet& R = ctx->rs[set];
    if (!R.pending) return false;
    cudaEventSynchronize(R.counted);
    R.pending = false;
    const unsigned need = *R.h_spanCount;
    if (need <= R.rb.spanCapacity) return false;
    unsigned cap = R.rb.spanCapacity;
    while (cap < need) cap *= 2;
    ctx->spanCapacityWanted = std::max(ctx->spanCapacityWanted, cap);
    return true;
}

int render_common(b200r_ctx* ctx, const b200r_frame* f, uint32_t* d_out, cudaStream_t stream, FrameParams& fp, int scratchSet = 0)
{
    int rc = make_frame_params(ctx, f, fp);
    if (rc) return rc;
    if (!ctx->have_scene) return fail(ctx, B200R_ESTATE, "b200r_render before b200r_upload_scene");
    if (ctx->counting || ctx->sw.pool_stats) CU(cudaMemsetAsync(ctx->d_ctr, 0, sizeof(DeviceCounters), stream));
    ctx->last_launches = 0;
    CU(cudaEventRecord(ctx->ev0, stream));
    switch (fp.mode) {
    case B200R_MODE_RAYTRACE:
    case B200R_MODE_RAYTRACE_AA:
        if (!ctx->have_bvh) return fail(ctx, B200R_ESTATE, "ray tracing needs a BVH (nodes/tri_idx were not uploaded)");
        {
            const uint32_t nTiles = ((fp.W + 7) / 8) * ((fp.n_rows + 3) / 4);
            unsigned long long* prof = nullptr;
            if (ctx->tileProfile) {
                if (ctx->tileProfTiles < nTiles) {
                    if (ctx->d_tileProf) cudaFree(ctx->d_tileProf);
                    ctx->d_tileProf = nullptr; ctx->tileProfTiles = 0;
                    CU(cudaMalloc((void**)&ctx->d_tileProf, (size_t)nTiles * 16));
                    ctx->tileProfTiles = nTiles;
                }
                prof = ctx->d_tileProf; ctx->lastTiles = nTiles;
            }
            const size_t px32 = (size_t)nTiles * 32;
            if (scratchSet < 0 || scratchSet >= B200R_MAX_FRAMES_IN_FLIGHT) return fail(ctx, B200R_EINVAL, "scratch set out of range");
            RtBuffers& rt = ctx->rts[scratchSet];
            if (!ctx->tileCounters[scratchSet]) CU(cudaMalloc((void**)&ctx->tileCounters[scratchSet], 64 * sizeof(unsigned)));
            if (rt.pixels < px32) {
                cudaFree(rt.hits); rt.hits = nullptr; rt.pixels = 0;
                CU(cudaMalloc((void**)&rt.hits, px32 * 32));             // one 32-byte hit record per pixel at most
                rt.pixels = px32;
            }
            if ((ctx->counting || ctx->sw.rt_legacy) && rt.legacyPixels < px32) {      // job pipeline: counting / legacy runs only
                cudaFree(rt.queue); cudaFree(rt.keys); rt.queue = nullptr; rt.keys = nullptr; rt.legacyPixels = 0;
                CU(cudaMalloc((void**)&rt.queue, px32 * 8 * 8));        // <= 8 jobs of 8 bytes per pixel
                CU(cudaMalloc((void**)&rt.keys, px32 * 8));
                rt.legacyPixels = px32;
            }
            rt.counters = ctx->tileCounters[scratchSet];
            rt.inFlight = ctx->framesInFlight;
            {   // wavefront buffers of the generic configurations (AO / reflections / two lights), sized on first use
                const bool simple = fp.n_lights == 1 && !(fp.flags & (B200R_F_REFLECTIONS | B200R_F_AO));
                const unsigned stride = ((fp.flags & B200R_F_AO) ? fp.ao_samples : 0u) + ((fp.flags & B200R_F_SHADOWS) ? fp.n_lights : 0u);
                if ((!simple || ctx->sw.no_fuse) && !ctx->sw.no_wavefront && !ctx->counting && fp.mode == B200R_MODE_RAYTRACE && fp.max_depth <= 3 &&
                    (rt.wfPixels < px32 || rt.wfStride < stride)) {
                    cudaFree(rt.wfHits1); cudaFree(rt.wfHits2); cudaFree(rt.wfPaths); cudaFree(rt.wfRefl);
                    cudaFree(rt.wfRays); cudaFree(rt.wfOcc); cudaFree(rt.wfCos); cudaFree(rt.wfCtx);
                    rt.wfHits1 = rt.wfHits2 = rt.wfPaths = nullptr; rt.wfRefl = rt.wfRays = rt.wfCtx = nullptr; rt.wfOcc = nullptr; rt.wfCos = nullptr;
                    rt.wfPixels = 0;
                    const size_t cap = std::max(px32, rt.pixels);
                    // hits are processed in chunks of at most 1 Mi (at most 16 chunks per level: very large frames get larger chunks)
                    const unsigned chunk = (unsigned)std::min<size_t>(cap, std::max<size_t>((size_t)1 << 20, (cap + 15) / 16));
                    const unsigned st = std::max(stride, std::max(rt.wfStride, 1u));
                    CU(cudaMalloc(&rt.wfHits1, cap * 32)); CU(cudaMalloc(&rt.wfHits2, cap * 32));
                    CU(cudaMalloc(&rt.wfPaths, cap * wavefront_path_bytes()));
                    CU(cudaMalloc((void**)&rt.wfRefl, cap * 48));
                    CU(cudaMalloc((void**)&rt.wfRays, (size_t)chunk * st * 48));
                    CU(cudaMalloc((void**)&rt.wfOcc, (size_t)chunk * st));
                    CU(cudaMalloc((void**)&rt.wfCos, (size_t)chunk * st * 4));
                    CU(cudaMalloc((void**)&rt.wfCtx, (size_t)chunk * 16));
                    rt.wfPixels = cap; rt.wfChunk = chunk; rt.wfStride = st;
                }
            }
            int launches = 0;
            CU(launch_raytrace(ctx->sc, fp, d_out, rt, ctx->sw, ctx->d_ctr, ctx->counting, prof, ctx->numSMs, stream, launches));
            ctx->last_launches += (uint32_t)launches;
        }
        break;
    case B200R_MODE_PHONG_SHADOWMAPS:
    case B200R_MODE_PHONG_SOFTSHADOWMAPS:
        for (uint32_t i = 0; i < fp.n_lights; i++)
            if (!ctx->sc.shadowmap[i])
                return fail(ctx, B200R_ESTATE, "modes 7/8 need a shadow map per light (b200r_render_shadowmap / b200r_upload_shadowmap)");
        /* fallthrough */
    case B200R_MODE_POINTS:
    case B200R_MODE_POINTS_TRI:
    case B200R_MODE_AMBIENT:
    case B200R_MODE_GOURAUD:
    case B200R_MODE_PHONG: {
        const size_t px = (size_t)fp.W * fp.n_rows;
        if (scratchSet < 0 || scratchSet >= B200R_MAX_FRAMES_IN_FLIGHT) return fail(ctx, B200R_EINVAL, "scratch set out of range");
        b200r_ctx::RasterSet& R = ctx->rs[scratchSet];
        if (span_overflowed(ctx, scratchSet)) ctx->spanOverflowSticky = true;       // a frame nobody retried (stream / slot calls)
        if (R.zkey_pixels < px) {
            if (R.rb.zkeys) cudaFree(R.rb.zkeys);
            R.rb.zkeys = nullptr; R.zkey_pixels = 0;
            CU(cudaMalloc((void**)&R.rb.zkeys, px * 8));
            R.zkey_pixels = px;
        }
        if (fp.mode >= B200R_MODE_PHONG && !ctx->sw.raster_inline_shade) {      // per-pixel lighting pass (default)
            if (R.attr_pixels < px) {
                if (R.d_attrs) cudaFree(R.d_attrs);
                R.d_attrs = nullptr; R.attr_pixels = 0;
                CU(cudaMalloc((void**)&R.d_attrs, px * 32));
                R.attr_pixels = px;
            }
            R.rb.attrs = R.d_attrs;
        } else R.rb.attrs = nullptr;
        if (!R.rb.spanCount) {
            CU(cudaMalloc((void**)&R.rb.spanCount, 64));
            CU(cudaMallocHost((void**)&R.h_spanCount, 64));
            CU(cudaEventCreateWithFlags(&R.counted, cudaEventDisableTiming));
        }
        // The span buffer is sized up front (one 80-byte record per triangle and scanline it covers: 48 scanlines per triangle
        // on average is far beyond any of the reference's models at 4K) and NOT read back inside the frame: the count follows the
        // frame to pinned memory and is looked at when the frame is retired (blocking calls: the frame is simply rendered again
        // with a larger buffer; stream / slot calls: reported by the next call).
        const unsigned wantCap = std::max(ctx->spanCapacityWanted, std::max(1u << 20, 48u * ctx->sc.n_tris));
        if (R.rb.spanCapacity < wantCap) {
            if (R.rb.spans) cudaFree(R.rb.spans);
            R.rb.spans = nullptr; R.rb.spanCapacity = 0;
            CU(cudaMalloc((void**)&R.rb.spans, (size_t)wantCap * 80));
            R.rb.spanCapacity = wantCap;
        }
        {
            int launches = 0;
            CU(launch_raster(ctx->sc, fp, d_out, R.rb, ctx->d_ctr, ctx->counting, ctx->numSMs, stream, launches));
            ctx->last_launches += (uint32_t)launches;
            if (fp.mode > B200R_MODE_POINTS_TRI) {
                CU(cudaMemcpyAsync(R.h_spanCount, R.rb.spanCount, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
                CU(cudaEventRecord(R.counted, stream));
                R.pending = true;
            }
        }
        break;
    }
    case B200R_MODE_LINES: {
        const size_t px = (size_t)fp.W * fp.n_rows;
        if (fp.W > 32767 || fp.H > 32767) return fail(ctx, B200R_EINVAL, "mode 3 uses 16-bit screen coordinates (Sint16) like the reference");
        if (ctx->wb.pixels < px) {
            cudaFree(ctx->wb.counts); cudaFree(ctx->wb.offsets); cudaFree(ctx->wb.blockSums);
            ctx->wb.counts = ctx->wb.offsets = ctx->wb.blockSums = nullptr; ctx->wb.pixels = 0;
            CU(cudaMalloc((void**)&ctx->wb.counts, px * 4));
            CU(cudaMalloc((void**)&ctx->wb.offsets, (px + 1) * 4));
            CU(cudaMalloc((void**)&ctx->wb.blockSums, ((px + 1023) / 1024 + 1) * 4));
            ctx->wb.pixels = px;
        }
        if (!ctx->wb.total) CU(cudaMalloc((void**)&ctx->wb.total, 64));
        int launches = 0;
        CU(launch_wire_count(ctx->sc, fp, d_out, ctx->wb, stream, launches));
        // the fragment buffer is sized from the count pass (one small read-back per frame)
        CU(cudaMemcpyAsync(ctx->h_spanCount, ctx->wb.total, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        CU(cudaStreamSynchronize(stream));
        const unsigned need = *ctx->h_spanCount;
        if (need > ctx->wb.capacity) {
            cudaFree(ctx->wb.frags); ctx->wb.frags = nullptr; ctx->wb.capacity = 0;
            unsigned cap = 1u << 20;
            while (cap < need) cap *= 2;
            CU(cudaMalloc(&ctx->wb.frags, (size_t)cap * 8));
            ctx->wb.capacity = cap;
        }
        CU(launch_wire_emit(ctx->sc, fp, d_out, ctx->wb, ctx->numSMs, stream, launches));
        ctx->last_launches += (uint32_t)launches;
        break;
    }
    default:
        return fail(ctx, B200R_EINVAL, "render mode must be 1..10");
    }
    // Screen::ShowScreen's hook (reference src/Screen.h:130-137): MLAA over the finished frame. A row-sharded frame is
    // filtered after the all-gather instead (b200r_mlaa_device), because the filter needs the neighbouring rows.
    if ((fp.flags & B200R_F_MLAA) && fp.row_step == 1) {
        int rc2 = mlaa_on(ctx, d_out, fp.W, fp.H, stream, scratchSet);
        if (rc2) return rc2;
    }
    CU(cudaEventRecord(ctx->ev1, stream));
    return B200R_OK;
}

}  // namespace

struct SwitchName { const char* name; int Switches::*field; };
const SwitchName kSwitches[] = {
    {"monolithic_rt", &Switches::monolithic_rt}, {"no_prune", &Switches::no_prune}, {"no_fuse", &Switches::no_fuse}, {"no_wavefront", &Switches::no_wavefront},
    {"rt_legacy", &Switches::rt_legacy}, {"no_root_rect", &Switches::no_root_rect}, {"pool_small", &Switches::pool_small},
    {"split_depth", &Switches::split_depth}, {"raster_inline_shade", &Switches::raster_inline_shade}, {"mlaa_scan", &Switches::mlaa_scan},
    {"mlaa_fullscan", &Switches::mlaa_fullscan}, {"mlaa_nobatch", &Switches::mlaa_nobatch}, {"mlaa_no_tma", &Switches::mlaa_no_tma},
    {"no_frame_overlap", &Switches::no_frame_overlap}, {"bvh_serial_split", &Switches::bvh_serial_split},
    {"pool_stats", &Switches::pool_stats}, {"pool_policy", &Switches::pool_policy}, {"pool_scatter", &Switches::pool_scatter}, {"pool_occ3", &Switches::pool_occ3}, {"pool_tiles_per_warp", &Switches::pool_tiles_per_warp}, {"pool_cta_warps", &Switches::pool_cta_warps},
    {"pool_leaf_min", &Switches::pool_leaf_min}, {"pool_sort_min", &Switches::pool_sort_min}, {"pool_shade_min", &Switches::pool_shade_min},
    {"pool_refill_min", &Switches::pool_refill_min}, {"pool_low_water", &Switches::pool_low_water}, {"pool_dry", &Switches::pool_dry},
};

extern "C" {

int b200r_set_switch(b200r_ctx* ctx, const char* name, int value)
{
    if (!ctx || !name) return fail(ctx, B200R_EINVAL, "b200r_set_switch: NULL argument");
    for (const SwitchName& n : kSwitches)
        if (!strcmp(n.name, name)) {
            int rc = b200r_wait(ctx);               // frames in flight were enqueued under the old setting
            if (rc) return rc;
            ctx->sw.*(n.field) = value;
            return B200R_OK;
        }
    return fail(ctx, B200R_EINVAL, std::string("b200r_set_switch: unknown switch '") + name + "'");
}

const char* b200r_last_error(const b200r_ctx* ctx) { return ctx ? ctx->err.c_str() : global_error(); }

int b200r_init(int device, b200r_ctx** out)
{
    b200r_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, B200R_EINVAL, "b200r_init: NULL out");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, B200R_ENODEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                                  " (this library has no CPU rendering path)");
    if (device < 0 || device >= n) return fail(nullptr, B200R_EINVAL, "device ordinal out of range");
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) return fail(nullptr, B200R_ECUDA, cudaGetErrorString(e));
    if (p.major != 10)
        return fail(nullptr, B200R_ENODEVICE, std::string("device '") + p.name + "' is sm_" + std::to_string(p.major) +
                                                  std::to_string(p.minor) + "; this build contains sm_100a code only");
    ctx = new b200r_ctx();
    ctx->device = device; ctx->numSMs = p.multiProcessorCount;
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(cudaEventCreate(&ctx->ev0)); CU(cudaEventCreate(&ctx->ev1));
    CU(cudaMalloc((void**)&ctx->d_tileCounter, 64 * sizeof(unsigned)));
    ctx->tileCounters[0] = ctx->d_tileCounter;
    CU(cudaMalloc((void**)&ctx->d_ctr, sizeof(DeviceCounters)));
    CU(cudaMemset(ctx->d_ctr, 0, sizeof(DeviceCounters)));
    CU(cudaMallocHost((void**)&ctx->h_spanCount, 64));
    CU(rt_pool_configure());
    for (const SwitchName& n : kSwitches) {      // defaults from the environment, read ONCE here - never on the per-frame path
        std::string env = "B200R_";
        for (const char* c = n.name; *c; c++) env += (char)toupper((unsigned char)*c);
        if (const char* v = getenv(env.c_str())) ctx->sw.*(n.field) = (*v >= '0' && *v <= '9') || *v == '-' ? atoi(v) : 1;
    }
    *out = ctx;
    return B200R_OK;
}

void b200r_destroy(b200r_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaFree(ctx->d_nodes); cudaFree(ctx->d_leaftris); cudaFree(ctx->d_shade); cudaFree(ctx->d_rverts); cudaFree(ctx->d_rtris);
    for (int i = 0; i < B200R_MAX_LIGHTS; i++) cudaFree(ctx->d_shadowmap[i]);
    cudaFree(ctx->d_frame); cudaFree(ctx->d_tileCounter); cudaFree(ctx->d_ctr);
    cudaFree(ctx->wb.counts); cudaFree(ctx->wb.offsets); cudaFree(ctx->wb.blockSums); cudaFree(ctx->wb.total); cudaFree(ctx->wb.frags);
    for (int k = 0; k < B200R_MAX_FRAMES_IN_FLIGHT; k++) {
        RtBuffers& rt = ctx->rts[k];
        cudaFree(rt.queue); cudaFree(rt.hits); cudaFree(rt.keys);
        cudaFree(rt.wfHits1); cudaFree(rt.wfHits2); cudaFree(rt.wfPaths); cudaFree(rt.wfRefl); cudaFree(rt.wfRays); cudaFree(rt.wfOcc); cudaFree(rt.wfCos); cudaFree(rt.wfCtx);
        if (k) cudaFree(ctx->tileCounters[k]);
        if (ctx->rstream[k]) cudaStreamDestroy(ctx->rstream[k]);
    }
    cudaFree(ctx->d_tileProf); cudaFree(ctx->d_shadowKeys);
    for (auto& R : ctx->rs) {
        cudaFree(R.rb.zkeys); cudaFree(R.rb.spans); cudaFree(R.rb.spanCount); cudaFree(R.d_attrs); cudaFree(R.d_mlaaScratch); cudaFree(R.d_mlaaLines);
        if (R.h_spanCount) cudaFreeHost(R.h_spanCount);
        if (R.counted) cudaEventDestroy(R.counted);
    }
    if (ctx->h_spanCount) cudaFreeHost(ctx->h_spanCount);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    for (auto& S : ctx->slot) {
        if (S.inflight) cudaEventSynchronize(S.copied);
        cudaFree(S.d);
        if (S.staging) cudaFreeHost(S.staging);
        if (S.rendered) cudaEventDestroy(S.rendered);
        if (S.copied) cudaEventDestroy(S.copied);
    }
    if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int b200r_upload_scene(b200r_ctx* ctx, const b200r_vertex* verts, uint32_t n_verts, const b200r_tri* tris,
                       uint32_t n_tris, const b200r_bvhnode* nodes, uint32_t n_nodes, const int32_t* tri_idx,
                       uint32_t n_tri_idx)
{
    if (!ctx) return fail(nullptr, B200R_EINVAL, "NULL ctx");
    if (!verts || !tris || n_verts == 0 || n_tris == 0) return fail(ctx, B200R_EINVAL, "empty scene");
    if ((nodes != nullptr && n_nodes != 0) != (tri_idx != nullptr && n_tri_idx != 0))
        return fail(ctx, B200R_EINVAL, "nodes and tri_idx must be given together");
    CU(cudaSetDevice(ctx->device));
    for (uint32_t i = 0; i < n_tris; i++)
        if (tris[i].a >= n_verts || tris[i].b >= n_verts || tris[i].c >= n_verts)
            return fail(ctx, B200R_EINVAL, "triangle vertex index out of range");

    // ---- validate the BVH like CreateCFBVH does (depth < stack size), plus index ranges
    if (nodes && n_nodes) {
        // every node, reachable or not: the re-layout below walks all of them
        for (uint32_t i = 0; i < n_nodes; i++) {
            if (nodes[i].a & 0x80000000u) {
                if ((uint64_t)nodes[i].b + (nodes[i].a & 0x7fffffffu) > n_tri_idx)
                    return fail(ctx, B200R_EINVAL, "BVH leaf range outside the triangle index list");
            } else if (nodes[i].a >= n_nodes || nodes[i].b >= n_nodes)
                return fail(ctx, B200R_EINVAL, "BVH child index out of range");
        }
        std::vector<std::pair<uint32_t, int>> st; st.push_back({0u, 0});
        size_t visited = 0;
        while (!st.empty()) {
            auto [i, d] = st.back(); st.pop_back();
            if (i >= n_nodes) return fail(ctx, B200R_EINVAL, "BVH child index out of range");
            if (++visited > n_nodes) return fail(ctx, B200R_EINVAL, "BVH is not a tree");
            if (d >= B200R_BVH_STACK_SIZE) return fail(ctx, B200R_EDEPTH, "BVH deeper than BVH_STACK_SIZE (32)");
            if (nodes[i].a & 0x80000000u) {
                if ((uint64_t)nodes[i].b + (nodes[i].a & 0x7fffffffu) > n_tri_idx)
                    return fail(ctx, B200R_EINVAL, "BVH leaf range outside the triangle index list");
            } else { st.push_back({nodes[i].b, d + 1}); st.push_back({nodes[i].a, d + 1}); }
        }
        for (uint32_t i = 0; i < n_tri_idx; i++)
            if (tri_idx[i] < 0 || (uint32_t)tri_idx[i] >= n_tris) return fail(ctx, B200R_EINVAL, "tri_idx entry out of range");
    }

    // ---- may the closest-hit kernel prune by distance?  (rt_kernels.cu "Distance pruning"; check in host/edgecheck.cpp)
    // Triangles that fail the check make every BVH node above them "unprunable" (flag bits in the node record); the
    // rest of the tree is pruned as usual.
    std::vector<unsigned char> bad_tri(n_tris, 0);
    b200r::count_unbounded_triangles(verts, n_verts, tris, n_tris, 2e-5, bad_tri.data());
    // The pruning slack (1e-4 box margin, hits within ~3e-6 of their triangle) is an ABSOLUTE error budget: it covers the fp32
    // rounding of hit = o + d*s only while coordinates stay small. The loader rescales every model to |coord| <= 1.2; geometry
    // handed in directly at another scale is traversed without pruning (identical results, more work).
    uint32_t prune_ok = 1;
    for (uint32_t i = 0; i < n_verts && prune_ok; i++)
        for (int c = 0; c < 3; c++) if (!(fabsf(verts[i].pos[c]) <= 8.0f)) prune_ok = 0;

    std::vector<float4> hn, hl, hs, hv, ht;
    uint32_t root_ref = 0xFFFFFFFFu, fast_ok = 1;
    float root_lo[3] = {0, 0, 0}, root_hi[3] = {0, 0, 0};
    if (nodes && n_nodes) {
        // inner nodes get consecutive record ids in DFS order; a child ref is a record id or a leaf ref
        std::vector<uint32_t> inner_id(n_nodes, 0xFFFFFFFFu);
        uint32_t n_inner = 0;
        for (uint32_t i = 0; i < n_nodes; i++) if (!(nodes[i].a & 0x80000000u)) inner_id[i] = n_inner++;
        if (n_inner >= 0x40000000u || n_tri_idx >= 0x3fffffffu) return fail(ctx, B200R_EINVAL, "scene too large");
        auto ref_of = [&](uint32_t i) -> uint32_t {
            if (!(nodes[i].a & 0x80000000u)) return inner_id[i];
            return (nodes[i].a & 0x7fffffffu) ? (0x80000000u | nodes[i].b) : 0xFFFFFFFFu;
        };
        auto check = [&](float v) {
            const float a = fabsf(v);
            if (!(a == 0.f || (a >= 2.9103830456733704e-11f && a <= 1.125899906842624e15f))) fast_ok = 0;
        };
        hn.resize(4 * (size_t)n_inner);
        for (uint32_t i = 0; i < n_nodes; i++) {
            for (int c = 0; c < 3; c++) { check(nodes[i].lo[c]); check(nodes[i].hi[c]); }
            if (nodes[i].a & 0x80000000u) continue;
            const b200r_bvhnode &L = nodes[nodes[i].a], &R = nodes[nodes[i].b];
            float4* rec = &hn[4 * (size_t)inner_id[i]];
            rec[0] = make_float4(L.lo[0], L.hi[0], R.lo[0], R.hi[0]);
            rec[1] = make_float4(L.lo[1], L.hi[1], R.lo[1], R.hi[1]);
            rec[2] = make_float4(L.lo[2], L.hi[2], R.lo[2], R.hi[2]);
            rec[3] = make_float4(u2f(ref_of(nodes[i].a)), u2f(ref_of(nodes[i].b)), 0.f, 0.f);
        }
        // subtree-contains-an-unbounded-triangle flags: children have larger indices than their parent (DFS pre-order),
        // so one reverse sweep is a post-order pass
        {
            std::vector<unsigned char> sub(n_nodes, 0);
            std::vector<float> tlo(3 * (size_t)n_nodes), thi(3 * (size_t)n_nodes);
            for (uint32_t ii = n_nodes; ii-- > 0;) {
                // tlo/thi: actual bounds of the triangles below the node. Pruning assumes they lie inside the node's box
                // (the reference's builder guarantees it, a caller's own tree may not): if not, the subtree is never pruned.
                if (nodes[ii].a & 0x80000000u) {
                    const uint32_t cnt = nodes[ii].a & 0x7fffffffu;
                    for (int c = 0; c < 3; c++) { tlo[3 * (size_t)ii + c] = FLT_MAX; thi[3 * (size_t)ii + c] = -FLT_MAX; }
                    for (uint32_t k = 0; k < cnt; k++) {
                        const uint32_t ti = (uint32_t)tri_idx[nodes[ii].b + k];
                        if (bad_tri[ti]) sub[ii] = 1;
                        const uint32_t vi[3] = {tris[ti].a, tris[ti].b, tris[ti].c};
                        for (int v = 0; v < 3; v++)
                            for (int c = 0; c < 3; c++) {
                                const float x = verts[vi[v]].pos[c];
                                if (!(x >= tlo[3 * (size_t)ii + c])) tlo[3 * (size_t)ii + c] = x;      // NaN propagates into the bounds
                                if (!(x <= thi[3 * (size_t)ii + c])) thi[3 * (size_t)ii + c] = x;
                            }
                    }
                    if (cnt == 0) for (int c = 0; c < 3; c++) { tlo[3 * (size_t)ii + c] = nodes[ii].lo[c]; thi[3 * (size_t)ii + c] = nodes[ii].hi[c]; }
                } else {
                    if (nodes[ii].a <= ii || nodes[ii].b <= ii) { std::fill(sub.begin(), sub.end(), 1); break; }   // not pre-order: be safe
                    sub[ii] = sub[nodes[ii].a] | sub[nodes[ii].b];
                    for (int c = 0; c < 3; c++) {
                        tlo[3 * (size_t)ii + c] = fminf(tlo[3 * (size_t)nodes[ii].a + c], tlo[3 * (size_t)nodes[ii].b + c]);
                        thi[3 * (size_t)ii + c] = fmaxf(thi[3 * (size_t)nodes[ii].a + c], thi[3 * (size_t)nodes[ii].b + c]);
                    }
                }
                for (int c = 0; c < 3; c++)
                    if (!(tlo[3 * (size_t)ii + c] >= nodes[ii].lo[c] - 1e-5f && thi[3 * (size_t)ii + c] <= nodes[ii].hi[c] + 1e-5f)) sub[ii] = 1;
            }
            for (uint32_t ii = 0; ii < n_nodes; ii++) {
                if (nodes[ii].a & 0x80000000u) continue;
                const uint32_t fl = (sub[nodes[ii].a] ? 1u : 0u) | (sub[nodes[ii].b] ? 2u : 0u);
                hn[4 * (size_t)inner_id[ii] + 3].z = u2f(fl);
            }
        }
        root_ref = ref_of(0);
        for (int c = 0; c < 3; c++) { root_lo[c] = nodes[0].lo[c]; root_hi[c] = nodes[0].hi[c]; }
        std::vector<char> last(n_tri_idx, 0);
        for (uint32_t i = 0; i < n_nodes; i++)
            if ((nodes[i].a & 0x80000000u) && (nodes[i].a & 0x7fffffffu)) last[nodes[i].b + (nodes[i].a & 0x7fffffffu) - 1] = 1;
        hl.resize(5 * (size_t)n_tri_idx);
        for (uint32_t i = 0; i < n_tri_idx; i++) {
            const uint32_t ti = (uint32_t)tri_idx[i];
            const b200r_tri& t = tris[ti];
            hl[5 * i + 0] = make_float4(t.normal[0], t.normal[1], t.normal[2], t.d);
            hl[5 * i + 1] = make_float4(t.e1[0], t.e1[1], t.e1[2], t.d1);
            hl[5 * i + 2] = make_float4(t.e2[0], t.e2[1], t.e2[2], t.d2);
            hl[5 * i + 3] = make_float4(t.e3[0], t.e3[1], t.e3[2], t.d3);
            hl[5 * i + 4] = make_float4(t.center[0], t.center[1], t.center[2],
                                        u2f((t.two_sided ? 0x80000000u : 0u) | (last[i] ? 0x40000000u : 0u) | ti));
        }
        if (n_tris >= 0x3fffffffu) return fail(ctx, B200R_EINVAL, "scene too large");
    }
    hs.resize(6 * (size_t)n_tris); ht.resize(4 * (size_t)n_tris);
    for (uint32_t i = 0; i < n_tris; i++) {
        const b200r_tri& t = tris[i];
        const b200r_vertex &A = verts[t.a], &B = verts[t.b], &C = verts[t.c];
        float s[24] = {A.pos[0], A.pos[1], A.pos[2], B.pos[0], B.pos[1], B.pos[2], C.pos[0], C.pos[1], C.pos[2],
                       A.nrm[0], A.nrm[1], A.nrm[2], B.nrm[0], B.nrm[1], B.nrm[2], C.nrm[0], C.nrm[1], C.nrm[2],
                       u2f(A.ao), u2f(B.ao), u2f(C.ao), t.colorf[0], t.colorf[1], t.colorf[2]};
        memcpy(&hs[6 * (size_t)i], s, sizeof s);
        ht[4 * (size_t)i + 0] = make_float4(u2f(t.a), u2f(t.b), u2f(t.c), u2f(t.two_sided));
        ht[4 * (size_t)i + 1] = make_float4(t.center[0], t.center[1], t.center[2], u2f(t.color));
        ht[4 * (size_t)i + 2] = make_float4(t.normal[0], t.normal[1], t.normal[2], 0.f);
        ht[4 * (size_t)i + 3] = make_float4(t.colorf[0], t.colorf[1], t.colorf[2], 0.f);
    }
    hv.resize(2 * (size_t)n_verts);
    for (uint32_t i = 0; i < n_verts; i++) {
        hv[2 * (size_t)i] = make_float4(verts[i].pos[0], verts[i].pos[1], verts[i].pos[2], u2f(verts[i].ao));
        hv[2 * (size_t)i + 1] = make_float4(verts[i].nrm[0], verts[i].nrm[1], verts[i].nrm[2], 0.f);
    }
    CU(upload(&ctx->d_nodes, hn, ctx->stream));
    CU(upload(&ctx->d_leaftris, hl, ctx->stream));
    CU(upload(&ctx->d_shade, hs, ctx->stream));
    CU(upload(&ctx->d_rverts, hv, ctx->stream));
    CU(upload(&ctx->d_rtris, ht, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));   // host staging vectors go out of scope
    ctx->nodesBytes = hn.size() * sizeof(float4); ctx->leafBytes = hl.size() * sizeof(float4); ctx->shadeBytes = hs.size() * sizeof(float4);
    ctx->sc.wnodes = ctx->d_nodes; ctx->sc.leaftris = ctx->d_leaftris; ctx->sc.shade = ctx->d_shade;
    ctx->sc.rverts = ctx->d_rverts; ctx->sc.rtris = ctx->d_rtris;
    ctx->sc.n_nodes = n_nodes; ctx->sc.n_list = n_tri_idx; ctx->sc.n_tris = n_tris; ctx->sc.n_verts = n_verts;
    ctx->sc.root_ref = root_ref; ctx->sc.fast_div_ok = fast_ok; ctx->sc.prune_ok = prune_ok;
    memcpy(ctx->sc.root_lo, root_lo, 12); memcpy(ctx->sc.root_hi, root_hi, 12);
    ctx->have_scene = true; ctx->have_bvh = (nodes && n_nodes);
    return B200R_OK;
}

int b200r_upload_shadowmap(b200r_ctx* ctx, int light, const float* map)
{
    if (!ctx || !map || light < 0 || light >= B200R_MAX_LIGHTS) return fail(ctx, B200R_EINVAL, "bad shadow map argument");
    CU(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)B200R_SHADOWMAP_SIZE * B200R_SHADOWMAP_SIZE * 4;
    if (!ctx->d_shadowmap[light]) CU(cudaMalloc((void**)&ctx->d_shadowmap[light], bytes));
    CU(cudaMemcpyAsync(ctx->d_shadowmap[light], map, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->sc.shadowmap[light] = ctx->d_shadowmap[light];
    return B200R_OK;
}

int b200r_render_shadowmap(b200r_ctx* ctx, int light, const float light_pos[3], const float world2light[9])
{
    if (!ctx || !light_pos || !world2light || light < 0 || light >= B200R_MAX_LIGHTS)
        return fail(ctx, B200R_EINVAL, "bad shadow map argument");
    if (!ctx->have_scene) return fail(ctx, B200R_ESTATE, "b200r_render_shadowmap before b200r_upload_scene");
    CU(cudaSetDevice(ctx->device));
    const size_t n = (size_t)B200R_SHADOWMAP_SIZE * B200R_SHADOWMAP_SIZE;
    if (!ctx->d_shadowmap[light]) CU(cudaMalloc((void**)&ctx->d_shadowmap[light], n * 4));
    if (!ctx->d_shadowKeys) CU(cudaMalloc((void**)&ctx->d_shadowKeys, n * 4));
    CU(launch_shadowmap(ctx->sc, light_pos, world2light, ctx->d_shadowKeys, ctx->d_shadowmap[light], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->sc.shadowmap[light] = ctx->d_shadowmap[light];
    return B200R_OK;
}

int b200r_download_shadowmap(b200r_ctx* ctx, int light, float* map)
{
    if (!ctx || !map || light < 0 || light >= B200R_MAX_LIGHTS || !ctx->d_shadowmap[light])
        return fail(ctx, B200R_EINVAL, "no such shadow map");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpy(map, ctx->d_shadowmap[light], (size_t)B200R_SHADOWMAP_SIZE * B200R_SHADOWMAP_SIZE * 4, cudaMemcpyDeviceToHost));
    return B200R_OK;
}

int b200r_render_device(b200r_ctx* ctx, const b200r_frame* f, void* dev_xrgb, void* cuda_stream)
{
    if (!ctx) return fail(nullptr, B200R_EINVAL, "NULL ctx");
    if (!dev_xrgb) return fail(ctx, B200R_EINVAL, "NULL device frame pointer");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    FrameParams fp;
    int rc = b200r_wait(ctx);                   // frames of b200r_render_async still in flight use the same scratch buffers
    if (rc) return rc;
    for (int attempt = 0;; attempt++) {
        rc = render_common(ctx, f, (uint32_t*)dev_xrgb, s, fp);
        if (rc) return rc;
        if (cuda_stream) break;
        CU(cudaStreamSynchronize(s));
        if (!span_overflowed(ctx, 0)) break;
        if (attempt >= 3) return fail(ctx, B200R_ENOMEM, "span buffer overflow persisted");
    }
    if (!cuda_stream) {
        CU(cudaEventElapsedTime(&ctx->last_total_ms, ctx->ev0, ctx->ev1));
        ctx->last_dominant_ms = ctx->last_total_ms;
    }
    return B200R_OK;
}

int b200r_render_device_slot(b200r_ctx* ctx, const b200r_frame* f, void* dev_xrgb, void* cuda_stream, uint32_t scratch_slot)
{
    if (!ctx) return fail(nullptr, B200R_EINVAL, "NULL ctx");
    if (!dev_xrgb || !cuda_stream) return fail(ctx, B200R_EINVAL, "b200r_render_device_slot needs a device frame pointer and a stream");
    if (scratch_slot >= B200R_MAX_FRAMES_IN_FLIGHT) return fail(ctx, B200R_EINVAL, "scratch_slot out of range");
    if (!f || f->mode == B200R_MODE_LINES)
        return fail(ctx, B200R_EINVAL, "b200r_render_device_slot: every mode but 3 (the wireframe pass sizes its fragment buffer on the host)");
    if (ctx->spanOverflowSticky) {
        ctx->spanOverflowSticky = false;
        return fail(ctx, B200R_ENOMEM, "an earlier rasterised frame overflowed its span buffer and is incomplete; the buffer has been enlarged - submit again");
    }
    if (ctx->counting || ctx->tileProfile) return fail(ctx, B200R_ESTATE, "b200r_render_device_slot: counters / tile profile use one shared buffer; switch them off");
    CU(cudaSetDevice(ctx->device));
    for (auto& S : ctx->slot)
        if (S.inflight) return fail(ctx, B200R_ESTATE, "b200r_render_device_slot while b200r_render_async frames are in flight (b200r_wait first)");
    FrameParams fp;
    ctx->framesInFlight = true;
    const int rc = render_common(ctx, f, (uint32_t*)dev_xrgb, (cudaStream_t)cuda_stream, fp, (int)scratch_slot);
    ctx->framesInFlight = false;
    return rc;
}

int b200r_set_pipeline_depth(b200r_ctx* ctx, uint32_t depth)
{
    if (!ctx) return fail(nullptr, B200R_EINVAL, "NULL ctx");
    if (depth < 1 || depth > B200R_MAX_FRAMES_IN_FLIGHT) return fail(ctx, B200R_EINVAL, "pipeline depth must be 1..B200R_MAX_FRAMES_IN_FLIGHT");
    int rc = b200r_wait(ctx);
    if (rc) return rc;
    ctx->depth = depth; ctx->asyncIdx = 0;
    return B200R_OK;
}

namespace {
bool host_pointer_is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

int retire_slot(b200r_ctx* ctx, b200r_ctx::AsyncSlot& S)
{
    if (!S.inflight) return B200R_OK;
    CU(cudaEventSynchronize(S.copied));
    if (span_overflowed(ctx, S.set)) {          // the frame's span buffer was too small: render it again, now, with the enlarged one
        FrameParams fp;
        for (int attempt = 0;; attempt++) {
            int rc = render_common(ctx, &S.frame, S.d, ctx->stream, fp, S.set);
            if (rc) return rc;
            CU(cudaMemcpyAsync(S.staged ? S.staging : S.user, S.d, S.user_words * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CU(cudaStreamSynchronize(ctx->stream));
            if (!span_overflowed(ctx, S.set)) break;
            if (attempt >= 3) return fail(ctx, B200R_ENOMEM, "span buffer overflow persisted");
        }
    }
    if (S.staged) memcpy(S.user, S.staging, S.user_words * 4);
    S.inflight = false;
    return B200R_OK;
}
}  // namespace

int b200r_render(b200r_ctx* ctx, const b200r_frame* f, uint32_t* host_xrgb)
{
    if (!ctx) return fail(nullptr, B200R_EINVAL, "NULL ctx");
    if (!host_xrgb) return fail(ctx, B200R_EINVAL, "NULL host frame pointer");
    CU(cudaSetDevice(ctx->device));
    FrameParams fp;
    int rc = make_frame_params(ctx, f, fp);
    if (rc) return rc;
    const size_t words = (size_t)fp.W * fp.n_rows;
    rc = ensure_frame(ctx, words);
    if (rc) return rc;
    rc = b200r_wait(ctx);                               // frames still in flight from b200r_render_async
    if (rc) return rc;
    const bool direct = host_pointer_is_pinned(host_xrgb);      // page-locked caller memory: DMA straight into it
    if (!direct && ctx->pinned_words < words) {
        if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
        ctx->h_pinned = nullptr; ctx->pinned_words = 0;
        CU(cudaMallocHost((void**)&ctx->h_pinned, words * 4));
        ctx->pinned_words = words;
    }
    for (int attempt = 0;; attempt++) {
        rc = render_common(ctx, f, ctx->d_frame, ctx->stream, fp);
        if (rc) return rc;
        CU(cudaMemcpyAsync(direct ? host_xrgb : ctx->h_pinned, ctx->d_frame, words * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (!span_overflowed(ctx, 0)) break;                 // the span buffer was too small: the frame is simply rendered again
        if (attempt >= 3) return fail(ctx, B200R_ENOMEM, "span buffer overflow persisted");
    }
    CU(cudaEventElapsedTime(&ctx->last_total_ms, ctx->ev0, ctx->ev1));
    ctx->last_dominant_ms = ctx->last_total_ms;
    if (!direct) memcpy(host_xrgb, ctx->h_pinned, words * 4);
    return B200R_OK;
}


int b200r_render_async(b200r_ctx* ctx, const b200r_frame* f, uint32_t* host_xrgb)
{
    if (!ctx) return fail(nullptr, B200R_EINVAL, "NULL ctx");
    if (!host_xrgb) return fail(ctx, B200R_EINVAL, "NULL host frame pointer");
    CU(cudaSetDevice(ctx->device));
    FrameParams fp;
    int rc = make_frame_params(ctx, f, fp);
    if (rc) return rc;
    const size_t words = (size_t)fp.W * fp.n_rows;
    if (!ctx->copyStream) {
        CU(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
        for (auto& S : ctx->slot) {
            CU(cudaEventCreateWithFlags(&S.rendered, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&S.copied, cudaEventDisableTiming));
        }
    }
    b200r_ctx::AsyncSlot& S = ctx->slot[ctx->asyncIdx % ctx->nslot()];
    rc = retire_slot(ctx, S);                    // the frame submitted nslot() calls ago: its copy must be out of S.d
    if (rc) return rc;
    if (S.words < words) {
        cudaFree(S.d); S.d = nullptr; S.words = 0;
        CU(cudaMalloc((void**)&S.d, words * 4));
        S.words = words;
    }
    S.staged = !host_pointer_is_pinned(host_xrgb);
    if (S.staged && S.staging_words < words) {
        if (S.staging) cudaFreeHost(S.staging);
        S.staging = nullptr; S.staging_words = 0;
        CU(cudaMallocHost((void**)&S.staging, words * 4));
        S.staging_words = words;
    }
    // Frames rotate over `depth` streams / scratch sets: frame i+1 starts while the tail of frame i is still running and takes
    // over the SMs it frees. Mode 3 (its fragment buffer is sized on the host) and profiling / counting runs stay on the one
    // stream and are therefore serialised.
    const bool overlap = fp.mode != B200R_MODE_LINES && !ctx->counting && !ctx->tileProfile && !ctx->sw.no_frame_overlap;
    const int set = overlap ? (int)(ctx->asyncIdx % ctx->depth) : 0;
    if (set && !ctx->rstream[set]) CU(cudaStreamCreateWithFlags(&ctx->rstream[set], cudaStreamNonBlocking));
    cudaStream_t rs = set ? ctx->rstream[set] : ctx->stream;
    ctx->framesInFlight = overlap && ctx->depth > 1;
    rc = render_common(ctx, f, S.d, rs, fp, set);
    ctx->framesInFlight = false;
    if (rc) return rc;
    CU(cudaEventRecord(S.rendered, rs));
    CU(cudaStreamWaitEvent(ctx->copyStream, S.rendered, 0));
    CU(cudaMemcpyAsync(S.staged ? S.staging : host_xrgb, S.d, words * 4, cudaMemcpyDeviceToHost, ctx->copyStream));
    CU(cudaEventRecord(S.copied, ctx->copyStream));
    S.user = host_xrgb; S.user_words = words; S.inflight = true; S.frame = *f; S.set = set;
    ctx->asyncIdx++;
    return B200R_OK;
}

int b200r_wait(b200r_ctx* ctx)
{
    if (!ctx) return fail(nullptr, B200R_EINVAL, "NULL ctx");
    CU(cudaSetDevice(ctx->device));
    for (unsigned k = 0; k < ctx->nslot(); k++) {               // oldest submission first
        int rc = retire_slot(ctx, ctx->slot[(ctx->asyncIdx + k) % ctx->nslot()]);
        if (rc) return rc;
    }
    return B200R_OK;
}

int b200r_mlaa_device(b200r_ctx* ctx, void* dev_xrgb, uint32_t width, uint32_t height, void* cuda_stream)
{
    if (!ctx || !dev_xrgb || !width || !height) return fail(ctx, B200R_EINVAL, "b200r_mlaa_device: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    int rc = mlaa_on(ctx, (uint32_t*)dev_xrgb, width, height, s);
    if (rc) return rc;
    if (!cuda_stream) CU(cudaStreamSynchronize(s));
    return B200R_OK;
}

int b200r_deinterleave_device(b200r_ctx* ctx, const void* dev_gathered, void* dev_frame, uint32_t width,
                              uint32_t height, uint32_t n_shards, void* cuda_stream)
{
    if (!ctx || !dev_gathered || !dev_frame || !width || !height || !n_shards)
        return fail(ctx, B200R_EINVAL, "b200r_deinterleave_device: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
    CU(launch_deinterleave((const uint32_t*)dev_gathered, (uint32_t*)dev_frame, width, height, n_shards, ctx->numSMs, s));
    if (!cuda_stream) CU(cudaStreamSynchronize(s));
    return B200R_OK;
}

int b200r_build_bvh(b200r_ctx* ctx, const b200r_vertex* verts, uint32_t n_verts, const b200r_tri* tris, uint32_t n_tris,
                    b200r_bvhnode* nodes_out, uint32_t nodes_cap, int32_t* tri_idx_out, uint32_t* n_nodes, int32_t* depth)
{
    if (!ctx) return fail(nullptr, B200R_EINVAL, "NULL ctx");
    if (!verts || !tris || !nodes_out || !tri_idx_out || !n_nodes || !depth || !n_verts || !n_tris)
        return fail(ctx, B200R_EINVAL, "b200r_build_bvh: bad argument");
    static_assert(sizeof(b200r_vertex) % sizeof(float) == 0 && offsetof(b200r_vertex, pos) == 0, "vertex positions lead the record");
    CU(cudaSetDevice(ctx->device));
    std::vector<uint32_t> idx(3 * (size_t)n_tris);
    for (uint32_t i = 0; i < n_tris; i++) {
        if (tris[i].a >= n_verts || tris[i].b >= n_verts || tris[i].c >= n_verts) return fail(ctx, B200R_EINVAL, "b200r_build_bvh: vertex index out of range");
        idx[3 * (size_t)i] = tris[i].a; idx[3 * (size_t)i + 1] = tris[i].b; idx[3 * (size_t)i + 2] = tris[i].c;
    }
    int launches = 0;
    // levels <= BVH_STACK_SIZE: the reference refuses deeper trees (Raytracer.cc:711-717)
    CU(launch_bvh_build(verts[0].pos, (int)(sizeof(b200r_vertex) / sizeof(float)), n_verts, idx.data(), n_tris, nodes_out, nodes_cap,
                        tri_idx_out, n_nodes, depth, B200R_BVH_STACK_SIZE, ctx->stream, launches, ctx->sw.bvh_serial_split != 0));
    ctx->last_launches = (uint32_t)launches;
    if (*depth < 0) return fail(ctx, B200R_EDEPTH, "Max depth of BVH exceeds BVH_STACK_SIZE");
    return B200R_OK;
}

int b200r_selftest_division(b200r_ctx* ctx, uint64_t samples, uint32_t seed, uint64_t* mismatches, float first_bad[4])
{
    if (!ctx || !mismatches) return fail(ctx, B200R_EINVAL, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    unsigned long long* d_m = nullptr; float* d_f = nullptr;
    CU(cudaMalloc((void**)&d_m, 8)); CU(cudaMalloc((void**)&d_f, 16));
    CU(cudaMemsetAsync(d_m, 0, 8, ctx->stream)); CU(cudaMemsetAsync(d_f, 0, 16, ctx->stream));
    CU(launch_division_selftest(samples, seed, d_m, d_f, ctx->numSMs, ctx->stream));
    unsigned long long m = 0; float fb[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(&m, d_m, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(fb, d_f, 16, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_m); cudaFree(d_f);
    *mismatches = m;
    if (first_bad) memcpy(first_bad, fb, 16);
    return B200R_OK;
}

int b200r_set_tile_profile(b200r_ctx* ctx, int enabled)
{
    if (!ctx) return fail(nullptr, B200R_EINVAL, "NULL ctx");
    ctx->tileProfile = enabled != 0;
    return B200R_OK;
}

int b200r_get_tile_profile(b200r_ctx* ctx, uint64_t* start_end_ns, uint32_t max_tiles, uint32_t* n_tiles)
{
    if (!ctx || !n_tiles) return fail(ctx, B200R_EINVAL, "NULL argument");
    *n_tiles = ctx->lastTiles;
    if (!start_end_ns || !ctx->d_tileProf) return B200R_OK;
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = ctx->lastTiles < max_tiles ? ctx->lastTiles : max_tiles;
    CU(cudaMemcpy(start_end_ns, ctx->d_tileProf, (size_t)n * 16, cudaMemcpyDeviceToHost));
    return B200R_OK;
}

int b200r_host_alloc(uint64_t bytes, void** out)
{
    if (!out || !bytes) return fail(nullptr, B200R_EINVAL, "b200r_host_alloc: bad argument");
    cudaError_t e = cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(nullptr, B200R_ENOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    return B200R_OK;
}

void b200r_host_free(void* p) { if (p) cudaFreeHost(p); }

int b200r_scene_buffer(b200r_ctx* ctx, uint32_t index, const void** dev_ptr, uint64_t* bytes)
{
    if (!ctx || !dev_ptr || !bytes) return fail(ctx, B200R_EINVAL, "NULL argument");
    if (!ctx->have_scene) return fail(ctx, B200R_ESTATE, "no scene uploaded");
    switch (index) {
    case 0: *dev_ptr = ctx->d_nodes; *bytes = (uint64_t)ctx->nodesBytes; return B200R_OK;
    case 1: *dev_ptr = ctx->d_leaftris; *bytes = (uint64_t)ctx->leafBytes; return B200R_OK;
    case 2: *dev_ptr = ctx->d_shade; *bytes = (uint64_t)ctx->shadeBytes; return B200R_OK;
    default: return fail(ctx, B200R_EINVAL, "no such scene buffer");
    }
}

int b200r_set_counters(b200r_ctx* ctx, int enabled)
{
    if (!ctx) return fail(nullptr, B200R_EINVAL, "NULL ctx");
    ctx->counting = enabled != 0;
    return B200R_OK;
}

int b200r_get_counters(b200r_ctx* ctx, b200r_counters* out)
{
    if (!ctx || !out) return fail(ctx, B200R_EINVAL, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    DeviceCounters h;
    CU(cudaMemcpy(&h, ctx->d_ctr, sizeof h, cudaMemcpyDeviceToHost));
    uint64_t* o = (uint64_t*)out;
    for (int i = 0; i < 11; i++) o[i] = h.v[i];
    return B200R_OK;
}

int b200r_last_kernel_ms(b200r_ctx* ctx, float* total_ms, float* dominant_ms)
{
    if (!ctx) return fail(nullptr, B200R_EINVAL, "NULL ctx");
    if (total_ms) *total_ms = ctx->last_total_ms;
    if (dominant_ms) *dominant_ms = ctx->last_dominant_ms;
    return B200R_OK;
}

int b200r_last_launches(b200r_ctx* ctx, uint32_t* n)
{
    if (!ctx || !n) return fail(ctx, B200R_EINVAL, "NULL argument");
    *n = ctx->last_launches;
    return B200R_OK;
}

}  // extern "C"
