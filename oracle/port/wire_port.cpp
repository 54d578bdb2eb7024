// wire_port.cpp — CPU restatement of the reference's anti-aliased WIREFRAME mode. TEST INFRASTRUCTURE ONLY.
//
// Serial (triangle order), for the headless surface the oracle uses: 32 bpp, Rmask 0xFF0000 / Gmask 0xFF00 /
// Bmask 0xFF, no alpha channel, clip rectangle = the whole frame.
//   Scene::renderWireframe                reference src/Rasterizers.cc:117-183
//   my_aalineColor / _aalineColor         reference src/Wu.cc:1282-1512  (Wu lines, 32-bit fixed point)
//   _clipLine / _clipEncode               reference src/Wu.cc:949-1052   (Cohen-Sutherland, float slope, Sint16 casts)
//   lineColor (alpha branch)              reference src/Wu.cc:1070-1255  (Bresenham, used for dx == dy)
//   hlineColor / vlineColor (alpha)       reference src/Wu.cc:652-937 -> filledRectAlpha -> _filledRectAlpha :320-575
//   pixelColorNolock / WeightNolock       reference src/Wu.cc:280-312, 621-637
//   _putPixelAlpha (32 bpp branch)        reference src/Wu.cc:163-193
// The colour handed down is SDL_MapRGB(200,200,200) = 0x00C8C8C8 but every callee decodes 0xRRGGBBAA, so lines are
// R=0, G=200, B=200 with alpha 200: every fragment blends the same colour with some alpha.
#include <cstdlib>
#include <cstring>

#include "oracle_port.h"
#include "port_common.h"

namespace {
using namespace oport;

typedef int16_t Sint16;
typedef uint32_t Uint32;
typedef uint8_t Uint8;

struct Surface {
    uint32_t* px; int W, H; int rowFirst, rowStep;       // px holds packed owned rows
    bool owns(int y) const { return y >= rowFirst && ((y - rowFirst) % rowStep) == 0; }
    uint32_t* at(int x, int y) { return owns(y) ? &px[(size_t)((y - rowFirst) / rowStep) * W + x] : nullptr; }
};

// the per-channel blend of _putPixelAlpha / _filledRectAlpha (Wu.cc:170-190, 469-486), Amask == 0
inline Uint32 blend(Uint32 dc, Uint32 color, Uint8 alpha)
{
    const Uint32 Rmask = 0x00FF0000u, Gmask = 0x0000FF00u, Bmask = 0x000000FFu;
    Uint32 R = ((dc & Rmask) + (((((color & Rmask) - (dc & Rmask)) >> 16) * alpha >> 8) << 16)) & Rmask;
    Uint32 G = ((dc & Gmask) + (((((color & Gmask) - (dc & Gmask)) >> 8) * alpha >> 8) << 8)) & Gmask;
    Uint32 B = ((dc & Bmask) + (((((color & Bmask) - (dc & Bmask)) >> 0) * alpha >> 8) << 0)) & Bmask;
    return R | G | B;
}

inline Uint32 map_rgba(Uint32 color) { return (((color >> 24) & 0xff) << 16) | (((color >> 16) & 0xff) << 8) | ((color >> 8) & 0xff); }

// _putPixelAlpha with the clip test (Wu.cc:47-246)
inline void putPixelAlpha(Surface& s, Sint16 x, Sint16 y, Uint32 mcolor, Uint8 alpha)
{
    if (x >= 0 && x <= s.W - 1 && y >= 0 && y <= s.H - 1) {
        uint32_t* p = s.at(x, y);
        if (!p) return;
        if (alpha == 255) *p = mcolor; else *p = blend(*p, mcolor, alpha);
    }
}
inline void pixelColorNolock(Surface& s, Sint16 x, Sint16 y, Uint32 color)
{ putPixelAlpha(s, x, y, map_rgba(color), (Uint8)(color & 0xff)); }
inline void pixelColorWeightNolock(Surface& s, Sint16 x, Sint16 y, Uint32 color, Uint32 weight)
{
    Uint32 a = (color & 0xffu);
    a = ((a * weight) >> 8);
    pixelColorNolock(s, x, y, (color & 0xffffff00u) | a);
}

// filledRectAlpha -> _filledRectAlpha, 32 bpp, no clipping inside (Wu.cc:320-575)
void filledRectAlpha(Surface& s, Sint16 x1, Sint16 y1, Sint16 x2, Sint16 y2, Uint32 color)
{
    const Uint8 alpha = color & 0xff;
    const Uint32 mcolor = map_rgba(color);
    for (Sint16 y = y1; y <= y2; y++)
        for (Sint16 x = x1; x <= x2; x++) {
            uint32_t* p = s.at(x, y);
            if (p) *p = blend(*p, mcolor, alpha);          // note: no alpha==255 shortcut on this path
            if (x == 32767) break;
        }
}

void hlineColor(Surface& s, Sint16 x1, Sint16 x2, Sint16 y, Uint32 color)
{
    if (x1 > x2) { Sint16 t = x1; x1 = x2; x2 = t; }
    const Sint16 left = 0, right = (Sint16)(s.W - 1), top = 0, bottom = (Sint16)(s.H - 1);
    if (x2 < left) return;
    if (x1 > right) return;
    if ((y < top) || (y > bottom)) return;
    if (x1 < left) x1 = left;
    if (x2 > right) x2 = right;
    const int dx = x2 - x1;
    if ((color & 255) == 255) {
        const Uint32 c = map_rgba(color);
        for (int x = x1; x <= x1 + dx; x++) { uint32_t* p = s.at(x, y); if (p) *p = c; }
    } else {
        filledRectAlpha(s, x1, y, (Sint16)(x1 + dx), y, color);
    }
}

void vlineColor(Surface& s, Sint16 x, Sint16 y1, Sint16 y2, Uint32 color)
{
    if (y1 > y2) { Sint16 t = y1; y1 = y2; y2 = t; }
    const Sint16 left = 0, right = (Sint16)(s.W - 1), top = 0, bottom = (Sint16)(s.H - 1);
    if ((x < left) || (x > right)) return;
    if (y2 < top) return;
    if (y1 > bottom) return;
    if (y1 < top) y1 = top;
    if (y2 > bottom) y2 = bottom;
    const Sint16 h = (Sint16)(y2 - y1);
    if ((color & 255) == 255) {
        const Uint32 c = map_rgba(color);
        for (int y = y1; y <= y1 + h; y++) { uint32_t* p = s.at(x, y); if (p) *p = c; }
    } else {
        filledRectAlpha(s, x, y1, x, (Sint16)(y1 + h), color);
    }
}

inline int clipEncode(Sint16 x, Sint16 y, Sint16 left, Sint16 top, Sint16 right, Sint16 bottom)
{
    int code = 0;
    if (x < left) code |= 1; else if (x > right) code |= 2;
    if (y < top) code |= 8; else if (y > bottom) code |= 4;
    return code;
}

// _clipLine (Wu.cc:990-1052)
int clipLine(const Surface& s, Sint16* x1, Sint16* y1, Sint16* x2, Sint16* y2)
{
    const Sint16 left = 0, right = (Sint16)(s.W - 1), top = 0, bottom = (Sint16)(s.H - 1);
    int draw = 0;
    while (1) {
        int code1 = clipEncode(*x1, *y1, left, top, right, bottom);
        int code2 = clipEncode(*x2, *y2, left, top, right, bottom);
        if (!(code1 | code2)) { draw = 1; break; }
        else if (code1 & code2) break;
        else {
            if (!code1) {
                Sint16 t = *x2; *x2 = *x1; *x1 = t;
                t = *y2; *y2 = *y1; *y1 = t;
                code1 = code2;
            }
            float m;
            if (*x2 != *x1) m = (*y2 - *y1) / (float)(*x2 - *x1); else m = 1.0f;
            if (code1 & 1) { *y1 += (Sint16)((left - *x1) * m); *x1 = left; }
            else if (code1 & 2) { *y1 += (Sint16)((right - *x1) * m); *x1 = right; }
            else if (code1 & 4) { if (*x2 != *x1) *x1 += (Sint16)((bottom - *y1) / m); *y1 = bottom; }
            else if (code1 & 8) { if (*x2 != *x1) *x1 += (Sint16)((top - *y1) / m); *y1 = top; }
        }
    }
    return draw;
}

// lineColor (Wu.cc:1070-1255), alpha != 255 branch reachable from _aalineColor's dx == dy case
void lineColor(Surface& s, Sint16 x1, Sint16 y1, Sint16 x2, Sint16 y2, Uint32 color)
{
    if (!clipLine(s, &x1, &y1, &x2, &y2)) return;
    if (x1 == x2) {
        if (y1 < y2) { vlineColor(s, x1, y1, y2, color); return; }
        else if (y1 > y2) { vlineColor(s, x1, y2, y1, color); return; }
        else { pixelColorNolock(s, x1, y1, color); return; }
    }
    if (y1 == y2) {
        if (x1 < x2) { hlineColor(s, x1, x2, y1, color); return; }
        else if (x1 > x2) { hlineColor(s, x2, x1, y1, color); return; }
    }
    int dx = x2 - x1, dy = y2 - y1;
    const int sx = (dx >= 0) ? 1 : -1, sy = (dy >= 0) ? 1 : -1;
    if ((color & 255) == 255) {
        const Uint32 c = map_rgba(color);
        dx = sx * dx + 1; dy = sy * dy + 1;
        int px = x1, py = y1, stepAx = sx, stepAy = 0, stepBx = 0, stepBy = sy;    // pixel += pixx / pixy
        if (dx < dy) { int t = dx; dx = dy; dy = t; stepAx = 0; stepAy = sy; stepBx = sx; stepBy = 0; }
        int x = 0, y = 0;
        for (; x < dx; x++, px += stepAx, py += stepAy) {
            uint32_t* p = s.at(px, py); if (p) *p = c;
            y += dy;
            if (y >= dx) { y -= dx; px += stepBx; py += stepBy; }
        }
    } else {
        const int ax = abs(dx) << 1, ay = abs(dy) << 1;
        int x = x1, y = y1;
        if (ax > ay) {
            int d = ay - (ax >> 1);
            while (x != x2) {
                pixelColorNolock(s, (Sint16)x, (Sint16)y, color);
                if (d > 0 || (d == 0 && sx == 1)) { y += sy; d -= ax; }
                x += sx; d += ay;
            }
        } else {
            int d = ax - (ay >> 1);
            while (y != y2) {
                pixelColorNolock(s, (Sint16)x, (Sint16)y, color);
                if (d > 0 || ((d == 0) && (sy == 1))) { x += sx; d -= ay; }
                y += sy; d += ax;
            }
        }
        pixelColorNolock(s, (Sint16)x, (Sint16)y, color);
    }
}

// _aalineColor(dst, x1, y1, x2, y2, color, draw_endpoint = 1) (Wu.cc:1282-1495)
void aalineColor(Surface& s, Sint16 x1, Sint16 y1, Sint16 x2, Sint16 y2, Uint32 color)
{
    if (!clipLine(s, &x1, &y1, &x2, &y2)) return;
    int32_t xx0 = x1, yy0 = y1, xx1 = x2, yy1 = y2;
    if (yy0 > yy1) { int t = yy0; yy0 = yy1; yy1 = t; t = xx0; xx0 = xx1; xx1 = t; }
    int dx = xx1 - xx0, dy = yy1 - yy0;
    if (dx == 0) { vlineColor(s, x1, y1, y2, color); return; }
    else if (dy == 0) { hlineColor(s, x1, x2, y1, color); return; }
    else if (dx == dy) { lineColor(s, x1, y1, x2, y2, color); return; }
    int xdir;
    if (dx >= 0) xdir = 1; else { xdir = -1; dx = -dx; }
    Uint32 erracc = 0;
    const Uint32 intshift = 32 - 8;
    pixelColorNolock(s, x1, y1, color);
    if (dy > dx) {
        const Uint32 erradj = (Uint32)(((dx << 16) / dy) << 16);
        int x0pxdir = xx0 + xdir;
        while (--dy) {
            const Uint32 erracctmp = erracc;
            erracc += erradj;
            if (erracc <= erracctmp) { xx0 = x0pxdir; x0pxdir += xdir; }
            yy0++;
            const Uint32 wgt = (erracc >> intshift) & 255;
            pixelColorWeightNolock(s, (Sint16)xx0, (Sint16)yy0, color, 255 - wgt);
            pixelColorWeightNolock(s, (Sint16)x0pxdir, (Sint16)yy0, color, wgt);
        }
    } else {
        const Uint32 erradj = (Uint32)(((dy << 16) / dx) << 16);
        int y0p1 = yy0 + 1;
        while (--dx) {
            const Uint32 erracctmp = erracc;
            erracc += erradj;
            if (erracc <= erracctmp) { yy0 = y0p1; y0p1++; }
            xx0 += xdir;
            const Uint32 wgt = (erracc >> intshift) & 255;
            pixelColorWeightNolock(s, (Sint16)xx0, (Sint16)yy0, color, 255 - wgt);
            pixelColorWeightNolock(s, (Sint16)xx0, (Sint16)y0p1, color, wgt);
        }
    }
    pixelColorNolock(s, x2, y2, color);
}

}  // namespace

namespace oport {

// Scene::renderWireframe, reference src/Rasterizers.cc:117-183
void render_wireframe(const oracle_scene* s, const b200r_frame* f, uint32_t* out, b200r_counters*)
{
    Surface sf;
    sf.px = out; sf.W = (int)f->width; sf.H = (int)f->height;
    sf.rowStep = f->row_step ? (int)f->row_step : 1; sf.rowFirst = (int)f->row_first;
    const int nRows = (sf.H - sf.rowFirst + sf.rowStep - 1) / sf.rowStep;
    memset(out, 0, (size_t)nRows * sf.W * 4);
    const Vec eye(f->eye);
    const Uint32 greyPixel = 0x00C8C8C8u;     // SDL_MapRGB(200,200,200)
    const int W = sf.W, H = sf.H, SCREEN_DIST = H * 2;
    for (uint32_t j = 0; j < s->n_tris; j++) {
        const b200r_tri& t = s->tris[j];
        Vec triToEye = eye; triToEye -= Vec(t.center);
        if (dot(triToEye, Vec(t.normal)) < 0) continue;
        Vec A = xform(Vec(s->verts[t.a].pos), eye, f->mv);
        Vec B = xform(Vec(s->verts[t.b].pos), eye, f->mv);
        Vec C = xform(Vec(s->verts[t.c].pos), eye, f->mv);
#define SCREENSPACE(P_, xx, yy) xx = int(W / 2 + SCREEN_DIST * P_.v[1] / P_.v[2]); yy = int(H / 2 - SCREEN_DIST * P_.v[0] / P_.v[2]);
        const bool agood = A.v[2] > 0.2f, bgood = B.v[2] > 0.2f, cgood = C.v[2] > 0.2f;
        if (agood) {
            int ax, ay; SCREENSPACE(A, ax, ay)
            if (bgood) {
                int bx, by; SCREENSPACE(B, bx, by)
                aalineColor(sf, (Sint16)ax, (Sint16)ay, (Sint16)bx, (Sint16)by, greyPixel);
                if (cgood) {
                    int cx, cy; SCREENSPACE(C, cx, cy)
                    aalineColor(sf, (Sint16)ax, (Sint16)ay, (Sint16)cx, (Sint16)cy, greyPixel);
                    aalineColor(sf, (Sint16)bx, (Sint16)by, (Sint16)cx, (Sint16)cy, greyPixel);
                }
            } else if (cgood) {
                int cx, cy; SCREENSPACE(C, cx, cy)
                aalineColor(sf, (Sint16)ax, (Sint16)ay, (Sint16)cx, (Sint16)cy, greyPixel);
            }
        } else if (bgood && cgood) {
            int bx, by, cx, cy; SCREENSPACE(B, bx, by) SCREENSPACE(C, cx, cy)
            aalineColor(sf, (Sint16)bx, (Sint16)by, (Sint16)cx, (Sint16)cy, greyPixel);
        }
#undef SCREENSPACE
    }
}

}  // namespace oport
