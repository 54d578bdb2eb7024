// raster_kernels.cu — the scan-conversion rasteriser hot path as hand-written CUDA for sm_100a.
//
// Replaces, for render modes 1, 2 and 4..8 (and the shadow-map pre-pass):
//   Scene::renderPoints / ProjectAndPlot          reference src/Rasterizers.cc:46-111
//   RasterizeScene<T>::DrawTriangles              reference src/Rasterizers.cc:242-310   (cull, transform, near reject, project)
//   Filler<T> / PhongSetup                        reference src/Fillers.h:176-300        (per-vertex interpolants)
//   ScanConverter::ScanConvert/InnerLoop/Add      reference src/ScanConverter.h:27-137   (incremental edge walk)
//   Screen::RasterizeTriangle / CheckZBuffer...   reference src/Screen.h:194-291         (incremental span walk, 1/z test)
//   Screen::Plot<T> / IlluminatePixel             reference src/Screen.cc:34-112
//   LightingEquation<mode>::ComputePixel          reference src/LightingEq.h:45-170
//   Light::RenderSceneIntoShadowBuffer            reference src/Light.cc:84-296
//
// The CPU code is a serial loop over triangles whose order decides Z ties ("first triangle wins", strict `<` at
// Screen.h:209). The GPU formulation is order-free and deterministic:
//   K1 setup+edges : one thread per triangle does the per-triangle setup and walks the three edges with the SAME
//                    incremental float adds (vtc += d12 per scanline), emitting one span record per scanline;
//   K2 depth       : one thread per span walks its pixels with the SAME incremental adds (start += dLR) and does
//                    atomicMax on a 64-bit key  (bits(1/z) << 32 | ~triangleIndex): max 1/z, earliest triangle on ties;
//   K3 resolve     : the same walk again; the fragment whose key equals the stored key is shaded and written.
// Edge and span interpolants are produced by repeated float `+=` in the reference, not in closed form, which is why
// the walks are sequential per edge / per span (the work per triangle is small).
#include <cfloat>

#include "device_types.cuh"
#include "rt_kernels.cuh"
#include "../raster_steps.h"

namespace b200r {

namespace {

constexpr float kClipPlaneDistance = 0.2f;        // reference src/Rasterizers.cc:39
constexpr int SMAP = B200R_SHADOWMAP_SIZE;
constexpr int SPAN_WORDS = 20;                     // tri, y|flags, 8+8 interpolants (+2 pad): five 16-byte chunks

struct Mat9 { float m[9]; };

// ---- one edge of ScanConverter::ScanConvert / InnerLoop (reference src/ScanConverter.h:90-136)
template <int N>
struct Edge {
    int y0, y1;            // rows this edge adds to (inclusive); y0 > y1 = none
    bool horizontal;       // y1 == y2 in the reference: adds v1 then v2 on that single row
    FPd<N> vtc, d12;       // running value / per-row delta  (horizontal: vtc = v1, d12 = v2)
};

template <int N>
__device__ __forceinline__ void edge_init(Edge<N>& e, int height, int ya, const FPd<N>& A, int yb, const FPd<N>& B)
{
    e.horizontal = false;
    if (ya == yb) {
        e.horizontal = true;
        if (ya >= 0 && ya < height) { e.y0 = e.y1 = ya; } else { e.y0 = 1; e.y1 = 0; }
        e.vtc = A; e.d12 = B;
        return;
    }
    int y1 = ya, y2 = yb;
    const FPd<N>* v1 = &A; const FPd<N>* v2 = &B;
    if (!(y1 < y2)) { y1 = yb; y2 = ya; v1 = &B; v2 = &A; }
    if ((y1 < 0 && y2 < 0) || (y1 >= height && y2 >= height)) { e.y0 = 1; e.y1 = 0; return; }
    e.vtc = *v1;
    const float dy = (float)(y2 - y1);
#pragma unroll
    for (int i = 0; i < N; i++) { float t = v2->v[i]; t -= v1->v[i]; t /= dy; e.d12.v[i] = t; }
    if (y1 < 0) {
        const float k = (float)-y1;
#pragma unroll
        for (int i = 0; i < N; i++) { float t = e.d12.v[i]; t *= k; e.vtc.v[i] += t; }
        y1 = 0;
    }
    y2 = min(y2, height - 1);
    e.y0 = y1; e.y1 = y2;
}

// ScanlineAdd (reference src/ScanConverter.h:33-53): cnt in {0,1,2}
template <int N>
__device__ __forceinline__ void scanline_add(int& cnt, FPd<N>& L, FPd<N>& R, const FPd<N>& v)
{
    if (cnt == 0) { L = v; cnt = 1; }
    else if (cnt == 1) {
        if (L.v[0] <= v.v[0]) R = v; else { R = L; L = v; }
        cnt = 2;
    } else {
        if (v.v[0] < L.v[0]) L = v;
        else if (v.v[0] > R.v[0]) R = v;
    }
}

// Walk the three edges (AB, AC, BC order unless `lightOrder`: 12, 23, 13) and call emit(y, cnt, L, R) per scanline.
template <int N, class Emit>
__device__ __forceinline__ void scan_triangle(int height, int ya, const FPd<N>& A, int yb, const FPd<N>& B, int yc,
                                              const FPd<N>& C, bool lightOrder, Emit&& emit)
{
    Edge<N> e0, e1, e2;
    if (!lightOrder) {
        edge_init<N>(e0, height, ya, A, yb, B);
        edge_init<N>(e1, height, ya, A, yc, C);
        edge_init<N>(e2, height, yb, B, yc, C);
    } else {
        edge_init<N>(e0, height, ya, A, yb, B);
        edge_init<N>(e1, height, yb, B, yc, C);
        edge_init<N>(e2, height, ya, A, yc, C);
    }
    int ymin = height, ymax = -1;
    if (e0.y0 <= e0.y1) { ymin = min(ymin, e0.y0); ymax = max(ymax, e0.y1); }
    if (e1.y0 <= e1.y1) { ymin = min(ymin, e1.y0); ymax = max(ymax, e1.y1); }
    if (e2.y0 <= e2.y1) { ymin = min(ymin, e2.y0); ymax = max(ymax, e2.y1); }
    for (int y = ymin; y <= ymax; y++) {
        int cnt = 0; FPd<N> L, R;
#define B2_EDGE(e)                                                                    \
        if (y >= e.y0 && y <= e.y1) {                                                 \
            if (e.horizontal) { scanline_add<N>(cnt, L, R, e.vtc); scanline_add<N>(cnt, L, R, e.d12); } \
            else { if (y > e.y0) fp_add<N>(e.vtc, e.d12); scanline_add<N>(cnt, L, R, e.vtc); }          \
        }
        B2_EDGE(e0) B2_EDGE(e1) B2_EDGE(e2)
#undef B2_EDGE
        emit(y, cnt, L, R);
    }
}

// ---- LightingEquation<mode>::ComputePixel (reference src/LightingEq.h:45-170). LM: 0 none, 1 hard, 2 soft shadows
struct Pix3 { float r, g, b; };

template <int LM>
__device__ __forceinline__ Pix3 compute_pixel(const DeviceScene& sc, const FrameParams& fp, const V3& inCam, const V3& nrm,
                                              const Pix3& material, float aoc)
{
    const float ambient = (float)(((double)(96.f * aoc) / 255.0) / 255.0);
    Pix3 target; target.b = ambient * material.b; target.g = ambient * material.g; target.r = ambient * material.r;
    for (uint32_t i = 0; i < fp.n_lights; i++) {
        Pix3 dColor; dColor.r = dColor.g = dColor.b = 0.f;
        V3 pointToLight = mkv3(fp.light_cam[i][0], fp.light_cam[i][1], fp.light_cam[i][2]) - inCam;
        int cntInShadow = 0;
        if (LM != 0) {
            const V3 lightToPoint = pointToLight * -1.0f;
            V3 inLight = mat3_mul(fp.cam2light[i], lightToPoint);
            inLight.x = (float)(SMAP / 2) + ((float)(SMAP * 2) * inLight.x) / inLight.z;
            inLight.y = (float)(SMAP / 2) + ((float)(SMAP * 2) * inLight.y) / inLight.z;
            inLight.z = 1.0f / inLight.z;
            int sx = cvtt_x86(inLight.x), sy = cvtt_x86(inLight.y);
            const float* sb = sc.shadowmap[i];
            const double zl = (double)inLight.z + 0.001;          // float + double literal, compared in double
            if (LM == 1) {
                if ((sx < 0) || (sx >= SMAP) || (sy < 0) || (sy >= SMAP)) continue;
                if (!((double)__ldg(&sb[(size_t)sy * SMAP + sx]) < zl)) continue;
            } else {
                const int basex = sx, basey = sy;
                for (int d = -1; d <= 1; d++) {
                    sy = (int)((unsigned)basey + (unsigned)d);
                    if ((sy < 0) || (sy >= SMAP)) continue;
                    for (int e = -1; e <= 1; e++) {
                        sx = (int)((unsigned)basex + (unsigned)e);
                        if ((sx < 0) || (sx >= SMAP)) continue;
                        if ((double)__ldg(&sb[(size_t)sy * SMAP + sx]) > zl) cntInShadow++;
                    }
                }
            }
        }
        pointToLight = normalize3(pointToLight);
        const float intensity = dot3(nrm, pointToLight);
        if (intensity < 0.f) {
        } else {
            const float df = (128.f * intensity) / 255.f;      // == (coord)(DIFFUSE*intensity/255.) (innocuous double rounding)
            dColor.b += df * material.b; dColor.g += df * material.g; dColor.r += df * material.r;
            const V3 pointToCamera = normalize3(inCam * -1.0f);
            const V3 half = normalize3(pointToLight + pointToCamera);
            float intensity2 = dot3(half, nrm);
            if (intensity2 > 0.f) {
                intensity2 *= intensity2; intensity2 *= intensity2; intensity2 *= intensity2;
                intensity2 *= intensity2; intensity2 *= intensity2;
                const float sp = (float)u8_x86(192.f * intensity2);
                dColor.r += sp; dColor.g += sp; dColor.b += sp;
            }
        }
        if (LM == 2) {
            if (cntInShadow) { const float k = (9.0f - (float)cntInShadow) / 9.0f; dColor.b = k * dColor.b; dColor.g = k * dColor.g; dColor.r = k * dColor.r; }
        }
        target.b += dColor.b; target.g += dColor.g; target.r += dColor.r;
    }
    if (target.b > 255.f) target.b = 255.f;
    if (target.g > 255.f) target.g = 255.f;
    if (target.r > 255.f) target.r = 255.f;
    return target;
}

// Screen::Plot<T> (reference src/Screen.cc:34-112): colour word of one fragment.
//   N==5: v = {projx, z, b, g, r}     N==8: v = {projx, x/z, y/z, 1/z, ao, nx, ny, nz}
template <int N, int LM>
__device__ __forceinline__ uint32_t shade_fragment(const DeviceScene& sc, const FrameParams& fp, const FPd<N>& v, uint32_t tri)
{
    if constexpr (N == 5) {
        return (u8_x86(v.v[4]) << 16) | (u8_x86(v.v[3]) << 8) | u8_x86(v.v[2]);
    } else {
        V3 point = mkv3(v.v[1], v.v[2], v.v[3]);
        point.x /= point.z; point.y /= point.z; point.z = 1.0f / point.z;
        const V3 normal = normalize3(mkv3(v.v[5], v.v[6], v.v[7]));
        const float4 cf = __ldg(sc.rtris + 4 * (size_t)tri + 3);
        Pix3 mat; mat.r = cf.x; mat.g = cf.y; mat.b = cf.z;
        const Pix3 c = compute_pixel<LM>(sc, fp, point, normal, mat, v.v[4]);
        return (u8_x86(c.r) << 16) | (u8_x86(c.g) << 8) | u8_x86(c.b);
    }
}

// ---------------------------------------------------------------- K1: per-triangle setup + edge walk -> span records
// MODE: 4 ambient, 5 gouraud, 6/7/8 phong.  N = 5 for 4/5, 8 for 6..8.
template <int MODE>
__global__ void __launch_bounds__(128)
ras_setup_kernel(DeviceScene sc, FrameParams fp, uint32_t* __restrict__ spans, unsigned* __restrict__ spanCount,
                 unsigned spanCapacity, DeviceCounters* __restrict__ ctr, int count)
{
    constexpr int N = (MODE <= 5) ? 5 : 8;
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= sc.n_tris) return;
    const int W = (int)fp.W, H = (int)fp.H;
    const V3 eye = mkv3(fp.eye[0], fp.eye[1], fp.eye[2]);
    const float4 t0 = __ldg(sc.rtris + 4 * (size_t)j + 0);
    const float4 t1 = __ldg(sc.rtris + 4 * (size_t)j + 1);
    const float4 t2 = __ldg(sc.rtris + 4 * (size_t)j + 2);
    if (!__float_as_uint(t0.w)) {       // !_twoSided: backface cull (Rasterizers.cc:263-271)
        const V3 triToEye = eye - mkv3(t1.x, t1.y, t1.z);
        if (dot3(triToEye, mkv3(t2.x, t2.y, t2.z)) < 0.f) return;
    }
    const uint32_t ia = __float_as_uint(t0.x), ib = __float_as_uint(t0.y), ic = __float_as_uint(t0.z);
    const float4 pa = __ldg(sc.rverts + 2 * (size_t)ia), pb = __ldg(sc.rverts + 2 * (size_t)ib), pc = __ldg(sc.rverts + 2 * (size_t)ic);
    const V3 cA = transform3(mkv3(pa.x, pa.y, pa.z), eye, fp.mv); if (cA.z < kClipPlaneDistance) return;
    const V3 cB = transform3(mkv3(pb.x, pb.y, pb.z), eye, fp.mv); if (cB.z < kClipPlaneDistance) return;
    const V3 cC = transform3(mkv3(pc.x, pc.y, pc.z), eye, fp.mv); if (cC.z < kClipPlaneDistance) return;
    const float SD = (float)(H * 2), H2 = (float)(H / 2), W2 = (float)(W / 2);
    const float ay = H2 - (SD * cA.x) / cA.z, by = H2 - (SD * cB.x) / cB.z, cy = H2 - (SD * cC.x) / cC.z;
    if (ay < 0.f && by < 0.f && cy < 0.f) return;
    const float Hf = (float)H;
    if (ay >= Hf && by >= Hf && cy >= Hf) return;
    const float ax = W2 + (SD * cA.y) / cA.z, bx = W2 + (SD * cB.y) / cB.z, cx = W2 + (SD * cC.y) / cC.z;
    if (count) atomicAdd(&ctr->v[C_TRIS_SETUP], 1ull);

    FPd<N> P[3];
    const V3 cc[3] = {cA, cB, cC};
    const float xx[3] = {ax, bx, cx};
    const float4 pp[3] = {pa, pb, pc};
    const uint32_t vi[3] = {ia, ib, ic};
    const float4 cf = __ldg(sc.rtris + 4 * (size_t)j + 3);
    Pix3 colorf; colorf.r = cf.x; colorf.g = cf.y; colorf.b = cf.z;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        P[k].v[0] = xx[k];
        const float ao = (float)__float_as_uint(pp[k].w);
        if constexpr (N == 5) {
            P[k].v[1] = 1.0f / cc[k].z;
            Pix3 col;
            if constexpr (MODE == 4) {
                const float f = ao / 255.f;
                col.b = f * colorf.b; col.g = f * colorf.g; col.r = f * colorf.r;
            } else {
                const float4 nn = __ldg(sc.rverts + 2 * (size_t)vi[k] + 1);
                const V3 nrm = mat3_mul(fp.mv, mkv3(nn.x, nn.y, nn.z));
                col = compute_pixel<0>(sc, fp, cc[k], nrm, colorf, ao);
            }
            P[k].v[2] = col.b; P[k].v[3] = col.g; P[k].v[4] = col.r;
        } else {
            P[k].v[3] = 1.0f / cc[k].z;
            P[k].v[1] = cc[k].x / cc[k].z;
            P[k].v[2] = cc[k].y / cc[k].z;
            P[k].v[4] = ao;
            const float4 nn = __ldg(sc.rverts + 2 * (size_t)vi[k] + 1);
            const V3 nrm = mat3_mul(fp.mv, mkv3(nn.x, nn.y, nn.z));
            P[k].v[5] = nrm.x; P[k].v[6] = nrm.y; P[k].v[7] = nrm.z;
        }
    }
    const int iay = cvtt_x86(ay), iby = cvtt_x86(by), icy = cvtt_x86(cy);

    // rows this rank owns inside [ymin, ymax]: reserve their span slots with one atomic
    int lo = H, hi = -1;
    {
        const int ys[3][2] = {{iay, iby}, {iay, icy}, {iby, icy}};
#pragma unroll
        for (int e = 0; e < 3; e++) {
            int y1 = min(ys[e][0], ys[e][1]), y2 = max(ys[e][0], ys[e][1]);
            if ((y1 < 0 && y2 < 0) || (y1 >= H && y2 >= H)) continue;
            y1 = max(y1, 0); y2 = min(y2, H - 1);
            lo = min(lo, y1); hi = max(hi, y2);
        }
    }
    if (hi < lo) return;
    const int rf = (int)fp.row_first, rs = (int)fp.row_step;
    int firstOwned = lo <= rf ? rf : rf + ((lo - rf + rs - 1) / rs) * rs;
    const int nOwned = firstOwned > hi ? 0 : (hi - firstOwned) / rs + 1;
    if (count) atomicAdd(&ctr->v[C_SPANS], (unsigned long long)(hi - lo + 1));
    if (nOwned == 0) return;
    unsigned base = atomicAdd(spanCount, (unsigned)nOwned);
    unsigned slot = base;
    scan_triangle<N>(H, iay, P[0], iby, P[1], icy, P[2], false,
        [&](int y, int cnt, const FPd<N>& L, const FPd<N>& R) {
            if (y < rf || ((y - rf) % rs) != 0) return;
            const unsigned s = slot++;
            if (s >= spanCapacity) return;                 // overflow: the host sees spanCount > capacity and retries
            uint32_t* rec = spans + (size_t)s * SPAN_WORDS;
            uint4 w0;
            float buf[18];
#pragma unroll
            for (int i = 0; i < 8; i++) { buf[i] = (i < N && cnt >= 1) ? L.v[i < N ? i : 0] : 0.f; buf[8 + i] = (i < N && cnt >= 2) ? R.v[i < N ? i : 0] : 0.f; }
            buf[16] = 0.f; buf[17] = 0.f;
            w0.x = j; w0.y = (uint32_t)y | (cnt == 1 ? 0x80000000u : 0u) | (cnt == 0 ? 0x40000000u : 0u);
            w0.z = __float_as_uint(buf[0]); w0.w = __float_as_uint(buf[1]);
            uint4* r4 = reinterpret_cast<uint4*>(rec);
            r4[0] = w0;
#pragma unroll
            for (int q = 0; q < 4; q++)
                r4[1 + q] = make_uint4(__float_as_uint(buf[2 + 4 * q]), __float_as_uint(buf[3 + 4 * q]),
                                       __float_as_uint(buf[4 + 4 * q]), __float_as_uint(buf[5 + 4 * q]));
        });
}

template <int N>
__device__ __forceinline__ void load_span(const uint32_t* __restrict__ spans, unsigned s, uint32_t& tri, int& y, bool& single,
                                          bool& empty, FPd<N>& L, FPd<N>& R)
{
    const uint4* r4 = reinterpret_cast<const uint4*>(spans + (size_t)s * SPAN_WORDS);
    float buf[18];
    const uint4 w0 = __ldg(r4);
    tri = w0.x; y = (int)(w0.y & 0x3fffffffu); single = (w0.y & 0x80000000u) != 0;
    empty = (w0.y & 0x40000000u) != 0;
    buf[0] = __uint_as_float(w0.z); buf[1] = __uint_as_float(w0.w);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint4 w = __ldg(r4 + 1 + q);
        buf[2 + 4 * q] = __uint_as_float(w.x); buf[3 + 4 * q] = __uint_as_float(w.y);
        buf[4 + 4 * q] = __uint_as_float(w.z); buf[5 + 4 * q] = __uint_as_float(w.w);
    }
#pragma unroll
    for (int i = 0; i < N; i++) { L.v[i] = buf[i]; R.v[i] = buf[8 + i]; }
}

// (walk_span / walk_span_keyed: csrc/raster_steps.h)
constexpr int KEY_BATCH = 8;      // depth keys requested together by the resolve passes

__device__ __forceinline__ unsigned long long depth_key(float z, uint32_t tri)
{
    return ((unsigned long long)__float_as_uint(z) << 32) | (unsigned long long)(0xFFFFFFFFu - tri);
}

// ---------------------------------------------------------------- K2: depth pass
template <int N>
__global__ void __launch_bounds__(256)
ras_depth_kernel(FrameParams fp, const uint32_t* __restrict__ spans, const unsigned* __restrict__ spanCount,
                 unsigned spanCapacity, unsigned long long* __restrict__ zkeys, DeviceCounters* __restrict__ ctr, int count)
{
    const unsigned n = min(*spanCount, spanCapacity);
    const int W = (int)fp.W;
    unsigned long long tests = 0;
    for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        uint32_t tri; int y; bool single, empty; FPd<N> L, R;
        load_span<N>(spans, s, tri, y, single, empty, L, R);
        if (empty) continue;
        unsigned long long* row = zkeys + (size_t)((y - (int)fp.row_first) / (int)fp.row_step) * W;
        walk_span<N>(W, single, L, R, [&](int x, const FPd<N>& v) {
            tests++;
            const float z = v.v[N == 5 ? 1 : 3];
            if (z > 0.f) atomicMax(&row[x], depth_key(z, tri));     // Zbuffer starts at 0: `0 < z` then max 1/z wins
        });
    }
    if (count && tests) atomicAdd(&ctr->v[C_Z_TESTS], tests);
}

// ---------------------------------------------------------------- K3: resolve + shade
template <int N, int LM>
__global__ void __launch_bounds__(256)
ras_resolve_kernel(DeviceScene sc, FrameParams fp, const uint32_t* __restrict__ spans, const unsigned* __restrict__ spanCount,
                   unsigned spanCapacity, const unsigned long long* __restrict__ zkeys, uint32_t* __restrict__ out,
                   DeviceCounters* __restrict__ ctr, int count)
{
    const unsigned n = min(*spanCount, spanCapacity);
    const int W = (int)fp.W;
    unsigned long long wins = 0;
    for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        uint32_t tri; int y; bool single, empty; FPd<N> L, R;
        load_span<N>(spans, s, tri, y, single, empty, L, R);
        if (empty) continue;
        const size_t rowOff = (size_t)((y - (int)fp.row_first) / (int)fp.row_step) * W;
        walk_span_keyed<N, KEY_BATCH>(W, single, L, R, [&](int x) { return zkeys[rowOff + x]; },
                                      [&](int x, const FPd<N>& v, unsigned long long stored) {
            const float z = v.v[N == 5 ? 1 : 3];
            if (z > 0.f && stored == depth_key(z, tri)) {
                out[rowOff + x] = shade_fragment<N, LM>(sc, fp, v, tri);
                wins++;
            }
        });
    }
    if (count && wins) atomicAdd(&ctr->v[C_Z_PASSES], wins);
}

// ---------------------------------------------------------------- K3 split in two for per-pixel Phong (modes 6-8)
// ras_resolve_kernel shades inside the span walk: one thread per span, spans of ~5 pixels of which a part wins, ~300
// instructions per shaded pixel -> 9.8 of 32 lanes active per instruction (ncu, profiles/r01h_ncu_c4.txt). Every pixel has at
// most one winning fragment, so the walk only has to leave the winner's interpolants in a per-pixel slot (32 bytes) ...
__global__ void __launch_bounds__(256)
ras_resolve_attr_kernel(FrameParams fp, const uint32_t* __restrict__ spans, const unsigned* __restrict__ spanCount,
                        unsigned spanCapacity, const unsigned long long* __restrict__ zkeys, float4* __restrict__ attrs,
                        DeviceCounters* __restrict__ ctr, int count)
{
    constexpr int N = 8;
    const unsigned n = min(*spanCount, spanCapacity);
    const int W = (int)fp.W;
    unsigned long long wins = 0;
    for (unsigned s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
        uint32_t tri; int y; bool single, empty; FPd<N> L, R;
        load_span<N>(spans, s, tri, y, single, empty, L, R);
        if (empty) continue;
        const size_t rowOff = (size_t)((y - (int)fp.row_first) / (int)fp.row_step) * W;
        walk_span_keyed<N, KEY_BATCH>(W, single, L, R, [&](int x) { return zkeys[rowOff + x]; },
                                      [&](int x, const FPd<N>& v, unsigned long long stored) {
            const float z = v.v[3];
            if (z > 0.f && stored == depth_key(z, tri)) {
                float4* a = attrs + 2 * (rowOff + x);
                a[0] = make_float4(v.v[1], v.v[2], v.v[3], v.v[4]);
                a[1] = make_float4(v.v[5], v.v[6], v.v[7], 0.f);
                wins++;
            }
        });
    }
    if (count && wins) atomicAdd(&ctr->v[C_Z_PASSES], wins);
}
// ... and the lighting runs one thread per PIXEL (neighbouring pixels are covered or not together; this pass also writes the
// black of the uncovered ones, i.e. it is Screen::ClearScreen too). Same shade_fragment(), same inputs: same pixels.
template <int LM>
__global__ void __launch_bounds__(256)
ras_shade_pixels_kernel(DeviceScene sc, FrameParams fp, const unsigned long long* __restrict__ zkeys, const float4* __restrict__ attrs,
                        uint32_t* __restrict__ out, size_t nPixels)
{
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < nPixels; p += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long key = zkeys[p];
        uint32_t c = 0u;
        if (key) {
            const uint32_t tri = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
            const float4 a0 = __ldg(attrs + 2 * p), a1 = __ldg(attrs + 2 * p + 1);
            FPd<8> v;
            v.v[0] = 0.f; v.v[1] = a0.x; v.v[2] = a0.y; v.v[3] = a0.z; v.v[4] = a0.w; v.v[5] = a1.x; v.v[6] = a1.y; v.v[7] = a1.z;
            c = shade_fragment<8, LM>(sc, fp, v, tri);
        }
        out[p] = c;
    }
}

// ---------------------------------------------------------------- points (modes 1, 2)
__device__ __forceinline__ bool project_point(const FrameParams& fp, const V3& p, int& x, int& y)
{
    const int W = (int)fp.W, H = (int)fp.H;
    if (!(p.z > kClipPlaneDistance)) return false;
    const float SD = (float)(H * 2);
    x = cvtt_x86((float)(W / 2) + (SD * p.y) / p.z);
    y = cvtt_x86((float)(H / 2) - (SD * p.x) / p.z);
    return y >= 0 && y < H && x >= 0 && x < W;
}

__device__ __forceinline__ bool owned_row(const FrameParams& fp, int y, size_t& rowOff)
{
    const int rf = (int)fp.row_first, rs = (int)fp.row_step;
    if (y < rf || ((y - rf) % rs) != 0) return false;
    rowOff = (size_t)((y - rf) / rs) * fp.W;
    return true;
}

// mode 1: every vertex -> white (order-free)
__global__ void points_vertices_kernel(DeviceScene sc, FrameParams fp, uint32_t* __restrict__ out)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= sc.n_verts) return;
    const float4 p = __ldg(sc.rverts + 2 * (size_t)j);
    const V3 c = transform3(mkv3(p.x, p.y, p.z), mkv3(fp.eye[0], fp.eye[1], fp.eye[2]), fp.mv);
    int x, y; size_t ro;
    if (project_point(fp, c, x, y) && owned_row(fp, y, ro)) out[ro + x] = 0x00FFFFFFu;
}

// mode 2: the serial reference lets the LAST writer win (triangle order, then A, B, C). PASS 0 records
// max(sequence number) per pixel, PASS 1 writes the colour of the fragment that owns the maximum.
template <int PASS>
__global__ void points_triangles_kernel(DeviceScene sc, FrameParams fp, unsigned long long* __restrict__ zkeys,
                                        uint32_t* __restrict__ out)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= sc.n_tris) return;
    const V3 eye = mkv3(fp.eye[0], fp.eye[1], fp.eye[2]);
    const float4 t0 = __ldg(sc.rtris + 4 * (size_t)j + 0);
    const float4 t1 = __ldg(sc.rtris + 4 * (size_t)j + 1);
    const float4 t2 = __ldg(sc.rtris + 4 * (size_t)j + 2);
    if (dot3(eye - mkv3(t1.x, t1.y, t1.z), mkv3(t2.x, t2.y, t2.z)) < 0.f) return;   // note: ignores _twoSided, like the reference
    const uint32_t vi[3] = {__float_as_uint(t0.x), __float_as_uint(t0.y), __float_as_uint(t0.z)};
    const uint32_t color = __float_as_uint(t1.w);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float4 p = __ldg(sc.rverts + 2 * (size_t)vi[k]);
        const V3 c = transform3(mkv3(p.x, p.y, p.z), eye, fp.mv);
        int x, y; size_t ro;
        if (!project_point(fp, c, x, y) || !owned_row(fp, y, ro)) continue;
        const unsigned long long seq = (unsigned long long)j * 3ull + (unsigned long long)k + 1ull;
        if (PASS == 0) atomicMax(&zkeys[ro + x], seq);
        else if (zkeys[ro + x] == seq) out[ro + x] = color;
    }
}

// ---------------------------------------------------------------- shadow map (reference src/Light.cc:84-296)
__device__ __forceinline__ unsigned f2ord(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void shadow_clear_kernel(unsigned* __restrict__ keys)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)SMAP * SMAP) keys[i] = f2ord(__uint_as_float(0xFEFEFEFEu));   // ClearShadowBuffer: memset 254
}

__device__ __forceinline__ void plot_shadow(unsigned* __restrict__ keys, int y, const FPd<3>& v)
{
    // Light::PlotShadowPixel (:253-259): buffer = max(buffer, 1/z) — order-free, so a plain atomicMax on the
    // order-preserving integer image of the float. NaN never passes `buffer < v` and is skipped.
    const int idx = cvtt_x86(v.v[0]);
    if (idx >= 0 && idx < SMAP && v.v[2] == v.v[2]) atomicMax(&keys[(size_t)y * SMAP + idx], f2ord(v.v[2]));
}

__global__ void __launch_bounds__(128)
shadow_raster_kernel(DeviceScene sc, float lx, float ly, float lz, Mat9 w2lm, unsigned* __restrict__ keys)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= sc.n_tris) return;
    const float* w2l = w2lm.m;
    const float4 t0 = __ldg(sc.rtris + 4 * (size_t)j + 0);
    const uint32_t vi[3] = {__float_as_uint(t0.x), __float_as_uint(t0.y), __float_as_uint(t0.z)};
    const V3 light = mkv3(lx, ly, lz);
    FPd<3> P[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float4 p = __ldg(sc.rverts + 2 * (size_t)vi[k]);
        V3 x = mat3_mul(w2l, mkv3(p.x, p.y, p.z) - light);
        P[k].v[0] = (float)(SMAP / 2) + ((float)(SMAP * 2) * x.x) / x.z;
        P[k].v[1] = (float)(SMAP / 2) + ((float)(SMAP * 2) * x.y) / x.z;
        P[k].v[2] = 1.0f / x.z;
    }
    if (P[0].v[1] < 0.f && P[1].v[1] < 0.f && P[2].v[1] < 0.f) return;
    const float S = (float)SMAP;
    if (P[0].v[1] >= S && P[1].v[1] >= S && P[2].v[1] >= S) return;
    scan_triangle<3>(SMAP, cvtt_x86(P[0].v[1]), P[0], cvtt_x86(P[1].v[1]), P[1], cvtt_x86(P[2].v[1]), P[2], true,
        [&](int y, int cnt, const FPd<3>& L, const FPd<3>& R) {
            if (cnt == 0) return;
            if (cnt == 1) { plot_shadow(keys, y, L); return; }
            const int x1 = cvtt_x86(L.v[0]), x2 = cvtt_x86(R.v[0]);
            int steps = abs(x2 - x1);
            if (!steps) { plot_shadow(keys, y, L); plot_shadow(keys, y, R); return; }
            FPd<3> start = L, dLR;
            const float fs = (float)steps;
#pragma unroll
            for (int i = 0; i < 3; i++) { float t = R.v[i]; t -= start.v[i]; t /= fs; dLR.v[i] = t; }
            plot_shadow(keys, y, start);
            while (steps-- > 0) { fp_add<3>(start, dLR); plot_shadow(keys, y, start); }
        });
}

__global__ void shadow_finalize_kernel(const unsigned* __restrict__ keys, float* __restrict__ map)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (size_t)SMAP * SMAP) map[i] = ord2f(keys[i]);
}

template <int MODE>
cudaError_t run_raster_mode(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, RasterBuffers& rb,
                            DeviceCounters* d_ctr, bool count, int numSMs, cudaStream_t st, int& launches)
{
    constexpr int N = (MODE <= 5) ? 5 : 8;
    constexpr int LM = MODE == 7 ? 1 : (MODE == 8 ? 2 : 0);
    const unsigned tb = 128;
    ras_setup_kernel<MODE><<<(sc.n_tris + tb - 1) / tb, tb, 0, st>>>(sc, fp, rb.spans, rb.spanCount, rb.spanCapacity, d_ctr, count ? 1 : 0);
    const int grid = numSMs * 8;
    ras_depth_kernel<N><<<grid, 256, 0, st>>>(fp, rb.spans, rb.spanCount, rb.spanCapacity, rb.zkeys, d_ctr, count ? 1 : 0);
    if constexpr (N == 8) {
        if (rb.attrs) {
            const size_t px = (size_t)fp.W * fp.n_rows;
            ras_resolve_attr_kernel<<<grid, 256, 0, st>>>(fp, rb.spans, rb.spanCount, rb.spanCapacity, rb.zkeys, rb.attrs, d_ctr, count ? 1 : 0);
            ras_shade_pixels_kernel<LM><<<numSMs * 16, 256, 0, st>>>(sc, fp, rb.zkeys, rb.attrs, d_out, px);
            launches += 4;
            return cudaGetLastError();
        }
    }
    ras_resolve_kernel<N, LM><<<grid, 256, 0, st>>>(sc, fp, rb.spans, rb.spanCount, rb.spanCapacity, rb.zkeys, d_out, d_ctr, count ? 1 : 0);
    launches += 3;
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_raster(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, RasterBuffers& rb,
                          DeviceCounters* d_ctr, bool count, int numSMs, cudaStream_t st, int& launches)
{
    const size_t px = (size_t)fp.W * fp.n_rows;
    cudaError_t e = cudaSuccess;
    const bool perPixelShade = fp.mode >= B200R_MODE_PHONG && fp.mode <= B200R_MODE_PHONG_SOFTSHADOWMAPS && rb.attrs;
    if (!perPixelShade) e = cudaMemsetAsync(d_out, 0, px * 4, st);               // Screen::ClearScreen (else: ras_shade_pixels_kernel writes every pixel)
    if (e != cudaSuccess) return e;
    if (fp.mode == B200R_MODE_POINTS) {
        points_vertices_kernel<<<(sc.n_verts + 255) / 256, 256, 0, st>>>(sc, fp, d_out);
        launches += 1;
        return cudaGetLastError();
    }
    e = cudaMemsetAsync(rb.zkeys, 0, px * 8, st);                                 // Screen::ClearZbuffer
    if (e != cudaSuccess) return e;
    if (fp.mode == B200R_MODE_POINTS_TRI) {
        points_triangles_kernel<0><<<(sc.n_tris + 255) / 256, 256, 0, st>>>(sc, fp, rb.zkeys, d_out);
        points_triangles_kernel<1><<<(sc.n_tris + 255) / 256, 256, 0, st>>>(sc, fp, rb.zkeys, d_out);
        launches += 2;
        return cudaGetLastError();
    }
    e = cudaMemsetAsync(rb.spanCount, 0, sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    switch (fp.mode) {
    case B200R_MODE_AMBIENT: return run_raster_mode<4>(sc, fp, d_out, rb, d_ctr, count, numSMs, st, launches);
    case B200R_MODE_GOURAUD: return run_raster_mode<5>(sc, fp, d_out, rb, d_ctr, count, numSMs, st, launches);
    case B200R_MODE_PHONG: return run_raster_mode<6>(sc, fp, d_out, rb, d_ctr, count, numSMs, st, launches);
    case B200R_MODE_PHONG_SHADOWMAPS: return run_raster_mode<7>(sc, fp, d_out, rb, d_ctr, count, numSMs, st, launches);
    case B200R_MODE_PHONG_SOFTSHADOWMAPS: return run_raster_mode<8>(sc, fp, d_out, rb, d_ctr, count, numSMs, st, launches);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_shadowmap(const DeviceScene& sc, const float light_pos[3], const float world2light[9], unsigned* d_keys,
                             float* d_map, cudaStream_t st)
{
    Mat9 m;
    for (int i = 0; i < 9; i++) m.m[i] = world2light[i];
    const unsigned n = SMAP * SMAP;
    shadow_clear_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_keys);
    shadow_raster_kernel<<<(sc.n_tris + 127) / 128, 128, 0, st>>>(sc, light_pos[0], light_pos[1], light_pos[2], m, d_keys);
    shadow_finalize_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_keys, d_map);
    return cudaGetLastError();
}

}  // namespace b200r
