// api_port.cpp — entry points of the CPU restatement. TEST INFRASTRUCTURE ONLY (see oracle_port.h).
#include <cstring>
#include <omp.h>
#include "oracle_port.h"
#include "port_common.h"

static inline uint32_t mix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

extern "C" {

// Same counter-based stream as oracle/sdl_stub/sdl_stub.cc (oracle_seed / oracle_rand).
uint32_t oracle_ao_key(uint32_t frame, uint32_t x, uint32_t y)
{
    uint32_t k = mix32(frame * 0x9E3779B9u + 0x7F4A7C15u);
    k = mix32(k ^ (x * 0x85EBCA77u));
    k = mix32(k ^ (y * 0xC2B2AE3Du));
    return k;
}
int32_t oracle_ao_draw(uint32_t key, uint32_t n)
{
    uint32_t v = mix32(key + 0x9E3779B9u * n);
    v = mix32(v ^ key);
    return (int32_t)(v >> 1);
}

int oracle_supports(int mode, int mlaa)
{
    (void)mlaa;     // the MLAA post filter is restated (mlaa_port.cpp) and applies to every supported mode
    return mode >= B200R_MODE_POINTS && mode <= B200R_MODE_RAYTRACE_AA;
}

int oracle_render(const oracle_scene* s, const b200r_frame* f, uint32_t* out, b200r_counters* ctr, int threads)
{
    if (!s || !f || !out) return -1;
    if (threads <= 0) threads = omp_get_max_threads();
    if (ctr) memset(ctr, 0, sizeof *ctr);
    switch (f->mode) {
    case B200R_MODE_RAYTRACE:
    case B200R_MODE_RAYTRACE_AA:
        if (!s->nodes) return -1;
        oport::render_raytrace(s, f, out, ctr, threads);
        break;
    case B200R_MODE_PHONG_SHADOWMAPS:
    case B200R_MODE_PHONG_SOFTSHADOWMAPS:
        for (uint32_t i = 0; i < f->n_lights; i++) if (!s->shadowmap[i]) return -4;
        /* fallthrough */
    case B200R_MODE_POINTS:
    case B200R_MODE_POINTS_TRI:
    case B200R_MODE_AMBIENT:
    case B200R_MODE_GOURAUD:
    case B200R_MODE_PHONG:
        oport::render_raster(s, f, out, ctr, threads);
        break;
    case B200R_MODE_LINES:
        oport::render_wireframe(s, f, out, ctr);
        break;
    default:
        return -2;
    }
    // Screen::ShowScreen (reference src/Screen.h:130-137): MLAA over the finished frame when built --enable-mlaa
    if ((f->flags & B200R_F_MLAA) && (f->row_step <= 1)) {
        int rc = oracle_mlaa(out, (int)f->width, (int)f->height);
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"

// (raster / shadow-map / MLAA restatements live in raster_port.cpp / mlaa_port.cpp)
