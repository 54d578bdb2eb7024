// raster_port.cpp — CPU restatement of the reference RASTERISER. TEST INFRASTRUCTURE ONLY (see oracle_port.h).
//
// Serial (triangle-index order), i.e. the semantics of the reference run with one thread; the reference's OpenMP
// build races on the Z-buffer test-and-set (src/Screen.h:209-213 inside src/Rasterizers.cc:250) and only differs
// from this on exact 1/z ties.
//   ProjectAndPlot / Scene::renderPoints          reference src/Rasterizers.cc:46-111
//   RasterizeScene<T>::DrawTriangles              reference src/Rasterizers.cc:242-310
//   Filler<T> (Ambient/Gouraud/Phong*)            reference src/Fillers.h:176-300
//   ScanConverter                                 reference src/ScanConverter.h:27-137
//   Screen::RasterizeTriangle / CheckZBuffer...   reference src/Screen.h:194-291
//   Screen::Plot<T> / IlluminatePixel             reference src/Screen.cc:34-112
//   LightingEquation<mode>::ComputePixel          reference src/LightingEq.h:45-170
//   Light::RenderSceneIntoShadowBuffer & friends  reference src/Light.cc:84-296
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "oracle_port.h"
#include "port_common.h"

namespace {
using namespace oport;

const float ClipPlaneDistance = 0.2f;     // src/Rasterizers.cc:39
const int SMAP = B200R_SHADOWMAP_SIZE;

// FatPoint: N interpolated floats, member-wise += -= *= /= (the X-macro operators of src/Fillers.h:35-140).
//   Ambient/Gouraud: [0]=_projx [1]=_z [2..4]=_color b,g,r
//   Phong*:          [0]=_projx [1]=_x [2]=_y [3]=_z [4]=_ambientOcclusionCoeff [5..7]=_normal
template <int N>
struct FP {
    float v[N];
    FP& operator+=(const FP& r) { for (int i = 0; i < N; i++) v[i] += r.v[i]; return *this; }
    FP& operator-=(const FP& r) { for (int i = 0; i < N; i++) v[i] -= r.v[i]; return *this; }
    FP& operator*=(float r) { for (int i = 0; i < N; i++) v[i] *= r; return *this; }
    FP& operator/=(float r) { for (int i = 0; i < N; i++) v[i] /= r; return *this; }
};
template <int N> inline float projx(const FP<N>& p) { return p.v[0]; }
inline float projx(const Vec& p) { return p.v[0]; }

// src/ScanConverter.h:27-137
template <class T>
struct ScanConverter {
    unsigned* lines; T* left; T* right; int height; int minimum, maximum;
    ScanConverter(unsigned* l, T* le, T* ri, int h) : lines(l), left(le), right(ri), height(h), minimum(h), maximum(-1)
    { std::fill_n(lines, h, 0u); }
    void ScanlineAdd(int idx, const T& v)
    {
        if (!lines[idx]) { left[idx] = v; lines[idx]++; }
        else if (lines[idx] == 1) {
            if (projx(left[idx]) <= projx(v)) right[idx] = v;
            else { right[idx] = left[idx]; left[idx] = v; }
            lines[idx]++;
        } else {
            if (projx(v) < projx(left[idx])) left[idx] = v;
            else if (projx(v) > projx(right[idx])) right[idx] = v;
        }
        minimum = std::min<int>(minimum, idx);
        maximum = std::max<int>(maximum, idx);
    }
    void InnerLoop(int y1, int y2, const T& v1, const T& v2)
    {
        if (y1 < 0 && y2 < 0) return;
        if (y1 >= height && y2 >= height) return;
        T vtc = v1;
        T d12 = v2; d12 -= v1; d12 /= (float)(y2 - y1);
        if (y1 < 0) { T d = d12; d *= (float)-y1; vtc += d; y1 = 0; }
        y2 = std::min(y2, height - 1);
        int steps = y2 - y1;
        ScanlineAdd(y1, vtc);
        while (steps--) { y1++; vtc += d12; ScanlineAdd(y1, vtc); }
    }
    void ScanConvert(int y1, const T& v1, int y2, const T& v2)
    {
        if (y1 == y2) { if (y1 >= 0 && y1 < height) { ScanlineAdd(y1, v1); ScanlineAdd(y1, v2); } }
        else if (y1 < y2) InnerLoop(y1, y2, v1, v2);
        else InnerLoop(y2, y1, v2, v1);
    }
};

struct Raster {
    const oracle_scene* s; const b200r_frame* f;
    int W, H; Vec eye; const float* mv;
    std::vector<float> zbuf; uint32_t* out;
    int rowFirst, rowStep;
    uint64_t trisSetup = 0, zTests = 0, zPasses = 0, spans = 0;

    bool owns(int y) const { return y >= rowFirst && ((y - rowFirst) % rowStep) == 0; }
    void put(int y, int x, uint32_t c) { if (owns(y)) out[(size_t)((y - rowFirst) / rowStep) * W + x] = c; }

    // src/LightingEq.h:45-170; mode: 0 NoShadows, 1 ShadowMapping, 2 SoftShadowMapping
    void ComputePixel(int mode, const Vec& inCam, const Vec& nrm, const Pix& material, float aoc, Pix& target) const
    {
        target = material;
        float ambient = (float)((96.f * aoc / 255.0) / 255.0);
        target *= ambient;
        for (uint32_t i = 0; i < f->n_lights; i++) {
            const b200r_light& light = f->lights[i];
            Pix dColor;
            Vec pointToLight(light.in_camera); pointToLight -= inCam;
            int cntInShadow = 0;
            if (mode != 0) {
                Vec lightToPoint = pointToLight; lightToPoint *= -1;
                Vec inLight = matmul(light.cam2light, lightToPoint);
                inLight.v[0] = SMAP / 2 + SMAP * 2 * inLight.v[0] / inLight.v[2];
                inLight.v[1] = SMAP / 2 + SMAP * 2 * inLight.v[1] / inLight.v[2];
                inLight.v[2] = 1.0f / inLight.v[2];
                int sx = (int)inLight.v[0];
                int sy = (int)inLight.v[1];
                const float* sb = s->shadowmap[i];
                if (mode == 1) {
                    if ((sx < 0) || (sx >= SMAP) || (sy < 0) || (sy >= SMAP)) continue;
                    if (!(sb[(size_t)sy * SMAP + sx] < (inLight.v[2] + 0.001))) continue;
                } else {
                    int basex = sx, basey = sy;
                    for (int d = -1; d <= 1; d++) {
                        sy = (int)((unsigned)basey + (unsigned)d);     // wraps like the reference's int add does in practice
                        if ((sy < 0) || (sy >= SMAP)) continue;
                        for (int e = -1; e <= 1; e++) {
                            sx = (int)((unsigned)basex + (unsigned)e);
                            if ((sx < 0) || (sx >= SMAP)) continue;
                            if (sb[(size_t)sy * SMAP + sx] > (inLight.v[2] + 0.001)) cntInShadow++;
                        }
                    }
                }
            }
            pointToLight.normalize();
            float intensity = dot(nrm, pointToLight);
            if (intensity < 0.) {
            } else {
                Pix diffuse = material;
                diffuse *= (float)(128.f * intensity / 255.);
                dColor += diffuse;
                Vec pointToCamera = inCam; pointToCamera *= -1.f; pointToCamera.normalize();
                Vec half = pointToLight; half += pointToCamera; half.normalize();
                float intensity2 = dot(half, nrm);
                if (intensity2 > 0.) {
                    intensity2 *= intensity2; intensity2 *= intensity2; intensity2 *= intensity2;
                    intensity2 *= intensity2; intensity2 *= intensity2;
                    dColor += Pix((unsigned char)(192.f * intensity2), (unsigned char)(192.f * intensity2),
                                  (unsigned char)(192.f * intensity2));
                }
            }
            if (mode == 2) { if (cntInShadow) dColor *= (9.0f - cntInShadow) / 9.0f; }
            target += dColor;
        }
        if (target.b > 255) target.b = 255;
        if (target.g > 255) target.g = 255;
        if (target.r > 255) target.r = 255;
    }

    // Screen::Plot<T>, src/Screen.cc:34-112
    void plot5(int y, int x, const FP<5>& v)
    { put(y, x, map_rgb((unsigned char)v.v[4], (unsigned char)v.v[3], (unsigned char)v.v[2])); }
    void plot8(int y, int x, const FP<8>& v, const Pix& triColor, int lmode)
    {
        Vec point(v.v[1], v.v[2], v.v[3]);
        point.v[0] /= point.v[2]; point.v[1] /= point.v[2]; point.v[2] = 1.0f / point.v[2];
        Vec normal(v.v[5], v.v[6], v.v[7]); normal.normalize();
        Pix color;
        ComputePixel(lmode, point, normal, triColor, v.v[4], color);
        put(y, x, map_rgb((uint8_t)color.r, (uint8_t)color.g, (uint8_t)color.b));
    }

    static int myfloor(float val) { if (val < 0.) return int(val - 0.5f); return int(val + 0.5f); }   // Screen.h:218-221

    template <int N, class PlotFn>
    void check_z(bool xr, int y, int x, const FP<N>& v, PlotFn&& plot)
    {
        if (xr && (x < 0 || x >= W)) return;
        zTests++;
        const float z = (N == 5) ? v.v[1] : v.v[3];
        float& zb = zbuf[(size_t)y * W + x];
        if (zb < z) { zb = z; zPasses++; plot(y, x, v); }
    }

    // Screen::RasterizeTriangle, src/Screen.h:223-291
    template <int N, class PlotFn>
    void RasterizeTriangle(int ay, int by, int cy, const FP<N>& A, const FP<N>& B, const FP<N>& C,
                           unsigned* lines, FP<N>* left, FP<N>* right, PlotFn&& plot)
    {
        ScanConverter<FP<N>> sc(lines, left, right, H);
        sc.ScanConvert(ay, A, by, B);
        sc.ScanConvert(ay, A, cy, C);
        sc.ScanConvert(by, B, cy, C);
        for (int i = sc.minimum; i <= sc.maximum; i++) {
            spans++;
            if (lines[i] == 1) {
                check_z<N>(true, i, myfloor(left[i].v[0]), left[i], plot);
            } else {
                int x1 = myfloor(left[i].v[0]); if (x1 >= W) continue;
                int x2 = myfloor(right[i].v[0]); if (x2 < 0) continue;
                int steps = abs(x2 - x1);
                if (!steps) {
                    check_z<N>(true, i, myfloor(left[i].v[0]), left[i], plot);
                } else {
                    FP<N> start = left[i]; FP<N> dLR = right[i];
                    dLR -= start; dLR /= (float)steps;
                    if (x1 < 0) { FP<N> jump = dLR; jump *= (float)-x1; start += jump; steps -= (-x1); x1 = 0; }
                    if (x2 >= W) steps -= (x2 - W + 1);
                    check_z<N>(false, i, x1, start, plot);
                    while (steps--) { x1++; start += dLR; check_z<N>(false, i, x1, start, plot); }
                }
            }
        }
    }

    // RasterizeScene<T>::DrawTriangles (src/Rasterizers.cc:242-310) + Filler<T> (src/Fillers.h:176-300)
    void draw_triangles(int mode)
    {
        std::vector<unsigned> lines(H);
        std::vector<FP<5>> l5(H), r5(H);
        std::vector<FP<8>> l8(H), r8(H);
        const int SCREEN_DIST = H * 2;
        for (uint32_t j = 0; j < s->n_tris; j++) {
            const b200r_tri& t = s->tris[j];
            if (!t.two_sided) {
                Vec triToEye = eye; triToEye -= Vec(t.center);
                if (dot(triToEye, Vec(t.normal)) < 0) continue;
            }
            const b200r_vertex &VA = s->verts[t.a], &VB = s->verts[t.b], &VC = s->verts[t.c];
            Vec cA = xform(Vec(VA.pos), eye, mv); if (cA.v[2] < ClipPlaneDistance) continue;
            Vec cB = xform(Vec(VB.pos), eye, mv); if (cB.v[2] < ClipPlaneDistance) continue;
            Vec cC = xform(Vec(VC.pos), eye, mv); if (cC.v[2] < ClipPlaneDistance) continue;
            float ax, ay, bx, by, cx, cy;
            ay = H / 2 - SCREEN_DIST * cA.v[0] / cA.v[2];
            by = H / 2 - SCREEN_DIST * cB.v[0] / cB.v[2];
            cy = H / 2 - SCREEN_DIST * cC.v[0] / cC.v[2];
            if (ay < 0 && by < 0 && cy < 0) continue;
            if (ay >= H && by >= H && cy >= H) continue;
            ax = W / 2 + SCREEN_DIST * cA.v[1] / cA.v[2];
            bx = W / 2 + SCREEN_DIST * cB.v[1] / cB.v[2];
            cx = W / 2 + SCREEN_DIST * cC.v[1] / cC.v[2];
            trisSetup++;
            const int iay = (int)ay, iby = (int)by, icy = (int)cy;
            const Pix colorf(t.colorf[0], t.colorf[1], t.colorf[2]);
            if (mode == B200R_MODE_AMBIENT || mode == B200R_MODE_GOURAUD) {
                FP<5> P[3];
                const b200r_vertex* VV[3] = {&VA, &VB, &VC};
                const Vec* CC[3] = {&cA, &cB, &cC};
                const float XX[3] = {ax, bx, cx};
                for (int k = 0; k < 3; k++) {
                    P[k].v[0] = XX[k];
                    P[k].v[1] = 1.0f / CC[k]->v[2];
                    Pix col;
                    if (mode == B200R_MODE_AMBIENT) {
                        col = colorf; col *= VV[k]->ao / 255.f;
                    } else {
                        Vec nrm = matmul(mv, Vec(VV[k]->nrm));
                        ComputePixel(0, *CC[k], nrm, colorf, (float)VV[k]->ao, col);
                    }
                    P[k].v[2] = col.b; P[k].v[3] = col.g; P[k].v[4] = col.r;
                }
                RasterizeTriangle<5>(iay, iby, icy, P[0], P[1], P[2], lines.data(), l5.data(), r5.data(),
                                     [&](int y, int x, const FP<5>& v) { plot5(y, x, v); });
            } else {
                FP<8> P[3];
                const b200r_vertex* VV[3] = {&VA, &VB, &VC};
                const Vec* CC[3] = {&cA, &cB, &cC};
                const float XX[3] = {ax, bx, cx};
                for (int k = 0; k < 3; k++) {
                    P[k].v[0] = XX[k];
                    P[k].v[3] = 1.0f / CC[k]->v[2];
                    P[k].v[1] = CC[k]->v[0] / CC[k]->v[2];
                    P[k].v[2] = CC[k]->v[1] / CC[k]->v[2];
                    P[k].v[4] = (float)VV[k]->ao;
                    Vec nrm = matmul(mv, Vec(VV[k]->nrm));
                    P[k].v[5] = nrm.v[0]; P[k].v[6] = nrm.v[1]; P[k].v[7] = nrm.v[2];
                }
                const int lmode = mode == B200R_MODE_PHONG ? 0 : (mode == B200R_MODE_PHONG_SHADOWMAPS ? 1 : 2);
                RasterizeTriangle<8>(iay, iby, icy, P[0], P[1], P[2], lines.data(), l8.data(), r8.data(),
                                     [&](int y, int x, const FP<8>& v) { plot8(y, x, v, colorf, lmode); });
            }
        }
    }

    // ProjectAndPlot, src/Rasterizers.cc:46-54
    void project_and_plot(const Vec& p, uint32_t color)
    {
        const int SCREEN_DIST = H * 2;
        if (p.v[2] > ClipPlaneDistance) {
            int x = (int)(W / 2 + SCREEN_DIST * p.v[1] / p.v[2]);
            int y = (int)(H / 2 - SCREEN_DIST * p.v[0] / p.v[2]);
            if (y >= 0 && y < H && x >= 0 && x < W) put(y, x, color);
        }
    }
    // Scene::renderPoints, src/Rasterizers.cc:56-111
    void render_points(bool asTriangles)
    {
        if (!asTriangles) {
            for (uint32_t j = 0; j < s->n_verts; j++)
                project_and_plot(xform(Vec(s->verts[j].pos), eye, mv), 0x00FFFFFFu);
        } else {
            for (uint32_t j = 0; j < s->n_tris; j++) {
                const b200r_tri& t = s->tris[j];
                Vec triToEye = eye; triToEye -= Vec(t.center);
                if (dot(triToEye, Vec(t.normal)) < 0) continue;
                project_and_plot(xform(Vec(s->verts[t.a].pos), eye, mv), t.color);
                project_and_plot(xform(Vec(s->verts[t.b].pos), eye, mv), t.color);
                project_and_plot(xform(Vec(s->verts[t.c].pos), eye, mv), t.color);
            }
        }
    }
};

}  // namespace

namespace oport {

void render_raster(const oracle_scene* s, const b200r_frame* f, uint32_t* out, b200r_counters* ctr, int)
{
    Raster r;
    r.s = s; r.f = f; r.W = (int)f->width; r.H = (int)f->height; r.eye = Vec(f->eye); r.mv = f->mv; r.out = out;
    r.rowStep = f->row_step ? (int)f->row_step : 1; r.rowFirst = (int)f->row_first;
    const int nRows = (r.H - r.rowFirst + r.rowStep - 1) / r.rowStep;
    memset(out, 0, (size_t)nRows * r.W * 4);                       // ClearScreen
    if (f->mode == B200R_MODE_POINTS || f->mode == B200R_MODE_POINTS_TRI) {
        r.render_points(f->mode == B200R_MODE_POINTS_TRI);
    } else {
        r.zbuf.assign((size_t)r.W * r.H, 0.f);                      // ClearZbuffer
        r.draw_triangles((int)f->mode);
    }
    if (ctr) { ctr->tris_setup = r.trisSetup; ctr->spans = r.spans; ctr->z_tests = r.zTests; ctr->z_passes = r.zPasses; }
}

}  // namespace oport

// ---------------------------------------------------------------- shadow map (src/Light.cc:84-296)
extern "C" int oracle_render_shadowmap(const oracle_scene* s, const float light_pos[3], const float world2light[9],
                                       float* map)
{
    using namespace oport;
    if (!s || !map) return -1;
    memset(map, 254, (size_t)SMAP * SMAP * 4);                      // ClearShadowBuffer: 0xFEFEFEFE, a huge negative
    std::vector<unsigned> lines(SMAP);
    std::vector<Vec> left(SMAP), right(SMAP);
    const Vec light(light_pos);
    auto plotShadow = [&](int y, const Vec& v) {                    // Light::PlotShadowPixel, :253-259
        int idx = (int)v.v[0];
        if (idx >= 0 && idx < SMAP) { float& d = map[(size_t)y * SMAP + idx]; if (d < v.v[2]) d = v.v[2]; }
    };
    for (uint32_t j = 0; j < s->n_tris; j++) {
        const b200r_tri& t = s->tris[j];
        Vec P[3] = {Vec(s->verts[t.a].pos), Vec(s->verts[t.b].pos), Vec(s->verts[t.c].pos)};
        for (int k = 0; k < 3; k++) {
            P[k] -= light;
            P[k] = matmul(world2light, P[k]);
            Vec& x = P[k];
            x.v[0] = SMAP / 2 + SMAP * 2 * x.v[0] / x.v[2];
            x.v[1] = SMAP / 2 + SMAP * 2 * x.v[1] / x.v[2];
            x.v[2] = 1.0f / x.v[2];
        }
        if (P[0].v[1] < 0 && P[1].v[1] < 0 && P[2].v[1] < 0) continue;
        if (P[0].v[1] >= SMAP && P[1].v[1] >= SMAP && P[2].v[1] >= SMAP) continue;
        // Light::InterpolateTriangleOnShadowBuffer, :261-296 (edge order 12, 23, 13)
        ScanConverter<Vec> sc(lines.data(), left.data(), right.data(), SMAP);
        sc.ScanConvert(int(P[0].v[1]), P[0], int(P[1].v[1]), P[1]);
        sc.ScanConvert(int(P[1].v[1]), P[1], int(P[2].v[1]), P[2]);
        sc.ScanConvert(int(P[0].v[1]), P[0], int(P[2].v[1]), P[2]);
        for (int y = sc.minimum; y <= sc.maximum; y++) {
            if (lines[y] == 1) plotShadow(y, left[y]);
            else {
                int x1 = (int)left[y].v[0], x2 = (int)right[y].v[0];
                int steps = abs(x2 - x1);
                if (!steps) { plotShadow(y, left[y]); plotShadow(y, right[y]); }
                else {
                    Vec start = left[y];
                    Vec dLR = right[y]; dLR -= start; dLR /= (float)steps;
                    plotShadow(y, start);
                    while (steps--) { start += dLR; plotShadow(y, start); }
                }
            }
        }
    }
    return 0;
}
