/* Headless SDL 1.2 stand-in — TEST INFRASTRUCTURE ONLY (part of oracle/). See SDL.h.
 *
 * Presents are counted: present #0 is the blank screen shown before the main loop
 * (reference src/renderer.cc:314), present #k+1 is benchmark-orbit frame k.
 *   ORACLE_DUMP=<prefix>   write <prefix>_<frame>.xrgb (raw W*H uint32 0x00RRGGBB)
 *   ORACLE_FRAMES=a,b,c    only dump these frame numbers (default: all)
 */
#include "SDL.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static SDL_PixelFormat g_fmt;
static SDL_Surface g_surf;
static long g_presents = 0;          /* number of presents so far */

extern "C" {

int SDL_Init(Uint32) { return 0; }
void SDL_Quit(void) {}
char *SDL_GetError(void) { static char e[] = "headless stub"; return e; }

SDL_Surface *SDL_SetVideoMode(int w, int h, int bpp, Uint32 flags)
{
    if (bpp != 32) return NULL;
    memset(&g_fmt, 0, sizeof g_fmt);
    g_fmt.BitsPerPixel = 32; g_fmt.BytesPerPixel = 4;
    g_fmt.Rshift = 16; g_fmt.Gshift = 8; g_fmt.Bshift = 0; g_fmt.Ashift = 0;
    g_fmt.Rmask = 0x00FF0000u; g_fmt.Gmask = 0x0000FF00u; g_fmt.Bmask = 0x000000FFu; g_fmt.Amask = 0;
    g_fmt.Aloss = 8; g_fmt.alpha = 255;
    g_surf.flags = flags; g_surf.format = &g_fmt; g_surf.w = w; g_surf.h = h;
    g_surf.pitch = (Uint16)(w * 4);
    g_surf.pixels = aligned_alloc(64, ((size_t)w * h * 4 + 63) & ~(size_t)63);
    memset(g_surf.pixels, 0, (size_t)w * h * 4);
    g_surf.clip_rect.x = 0; g_surf.clip_rect.y = 0;
    g_surf.clip_rect.w = (Uint16)w; g_surf.clip_rect.h = (Uint16)h;
    g_surf.refcount = 1;
    return &g_surf;
}

Uint32 SDL_MapRGB(const SDL_PixelFormat *, Uint8 r, Uint8 g, Uint8 b)
{
    return ((Uint32)r << 16) | ((Uint32)g << 8) | (Uint32)b;
}

/* Amask == 0: SDL 1.2 drops alpha for a surface without an alpha channel. */
Uint32 SDL_MapRGBA(const SDL_PixelFormat *, Uint8 r, Uint8 g, Uint8 b, Uint8)
{
    return ((Uint32)r << 16) | ((Uint32)g << 8) | (Uint32)b;
}

int SDL_LockSurface(SDL_Surface *) { return 0; }
void SDL_UnlockSurface(SDL_Surface *) {}

int SDL_FillRect(SDL_Surface *dst, SDL_Rect *r, Uint32 color)
{
    int x0 = 0, y0 = 0, x1 = dst->w, y1 = dst->h;
    if (r) { x0 = r->x; y0 = r->y; x1 = r->x + r->w; y1 = r->y + r->h; }
    if (x0 < 0) x0 = 0; if (y0 < 0) y0 = 0;
    if (x1 > dst->w) x1 = dst->w; if (y1 > dst->h) y1 = dst->h;
    for (int y = y0; y < y1; y++) {
        Uint32 *row = (Uint32 *)((Uint8 *)dst->pixels + (size_t)y * dst->pitch);
        for (int x = x0; x < x1; x++) row[x] = color;
    }
    return 0;
}

static void present(SDL_Surface *s)
{
    long frame = g_presents - 1;   /* -1 = the blank pre-loop present */
    g_presents++;
    const char *prefix = getenv("ORACLE_DUMP");
    if (!prefix || frame < 0) return;
    const char *sel = getenv("ORACLE_FRAMES");
    if (sel && *sel) {
        int wanted = 0;
        const char *p = sel;
        while (*p) {
            char *end; long v = strtol(p, &end, 10);
            if (end == p) break;
            if (v == frame) { wanted = 1; break; }
            p = (*end == ',') ? end + 1 : end;
        }
        if (!wanted) return;
    }
    char name[4096];
    snprintf(name, sizeof name, "%s_%ld.xrgb", prefix, frame);
    FILE *fp = fopen(name, "wb");
    if (!fp) { perror(name); exit(2); }
    for (int y = 0; y < s->h; y++)
        fwrite((Uint8 *)s->pixels + (size_t)y * s->pitch, 4, (size_t)s->w, fp);
    fclose(fp);
}

void SDL_UpdateRect(SDL_Surface *s, Sint32, Sint32, Uint32, Uint32) { present(s); }
int SDL_Flip(SDL_Surface *s) { present(s); return 0; }
void SDL_WM_SetCaption(const char *, const char *) {}

Uint32 SDL_GetTicks(void)
{
    static struct timespec t0; static int init = 0;
    struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t);
    if (!init) { t0 = t; init = 1; }
    return (Uint32)((t.tv_sec - t0.tv_sec) * 1000L + (t.tv_nsec - t0.tv_nsec) / 1000000L);
}

void SDL_Delay(Uint32) {}
int SDL_PollEvent(SDL_Event *) { return 0; }
int SDL_EnableUNICODE(int) { return 0; }

/* ---- deterministic AO random stream (oracle-only) ----
 * A counter-based generator keyed by (frame, x, y): the n-th draw inside a pixel is
 * mix(key, n) >> 1, in [0, RAND_MAX]. Thread-local so OpenMP threads do not interact.
 * The product's CUDA shader implements the same function. */
static inline Uint32 mix32(Uint32 h)
{
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}
static __thread Uint32 t_key = 0, t_ctr = 0;

void oracle_seed(int x, int y)
{
    long frame = g_presents - 1;
    Uint32 k = mix32((Uint32)frame * 0x9E3779B9u + 0x7F4A7C15u);
    k = mix32(k ^ ((Uint32)x * 0x85EBCA77u));
    k = mix32(k ^ ((Uint32)y * 0xC2B2AE3Du));
    t_key = k; t_ctr = 0;
}

int oracle_rand(void)
{
    Uint32 v = mix32(t_key + 0x9E3779B9u * (t_ctr++));
    v = mix32(v ^ t_key);
    return (int)(v >> 1);
}

} /* extern "C" */
