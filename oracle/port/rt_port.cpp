// rt_port.cpp — CPU restatement of the reference RAY TRACER. TEST INFRASTRUCTURE ONLY (see oracle_port.h).
//
// Follows, expression for expression (same association, same float/double promotions, same casts):
//   RayIntersectsBox                 reference src/Raytracer.cc:99-151
//   BVH_IntersectTriangles<s,c>      reference src/Raytracer.cc:183-308
//   Raytrace<doCulling>              reference src/Raytracer.cc:315-553
//   RaytraceHorizontalSegment        reference src/Raytracer.cc:555-606
// The compile-time switches of src/Raytracer.cc:53-81 (USE_SHADOWS, REFLECTIONS, AMBIENT_OCCLUSION,
// AMBIENT_SAMPLES, USE_PHONG_NORMAL, MAX_RAY_DEPTH) and WIDTH/HEIGHT (src/Defines.h:26-27) are runtime
// fields of b200r_frame here. Built with the oracle's strict flags (-O2/-O3, no fast-math, no FMA).
#include <cfloat>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "oracle_port.h"
#include "port_common.h"

namespace {

using namespace oport;

struct Ctx {
    const oracle_scene* s;
    const b200r_frame* f;
    Vec eye;
    bool shadows, reflections, phong, ao;
    int aoSamples, maxDepth;
    // per-thread counters
    uint64_t nodeTests = 0, leafVisits = 0, triTests = 0;
    uint64_t raysPrimary = 0, raysShadow = 0, raysRefl = 0, raysAO = 0;
    // AO stream
    uint32_t aoKey = 0, aoCtr = 0;
    int rnd() { return oracle_ao_draw(aoKey, aoCtr++); }
};

// reference src/Raytracer.cc:99-151
inline bool RayIntersectsBox(const Vec& o, const Vec& d, const b200r_bvhnode& box)
{
    float Tnear = -FLT_MAX, Tfar = FLT_MAX;
    for (int c = 0; c < 3; c++) {
        if (d.v[c] == 0.) {
            if (o.v[c] < box.lo[c]) return false;
            if (o.v[c] > box.hi[c]) return false;
        } else {
            float T1 = (box.lo[c] - o.v[c]) / d.v[c];
            float T2 = (box.hi[c] - o.v[c]) / d.v[c];
            if (T1 > T2) { float tmp = T1; T1 = T2; T2 = tmp; }
            if (T1 > Tnear) Tnear = T1;
            if (T2 < Tfar) Tfar = T2;
            if (Tnear > Tfar) return false;
            if (Tfar < 0.) return false;
        }
    }
    return true;
}

// reference src/Raytracer.cc:183-308
template <bool stopAtfirstRayHit, bool doCulling>
bool Intersect(Ctx& c, const Vec& origin, const Vec& ray, int avoidSelf, int& bestTri,
               Vec& pointHit /* in: light position for shadow rays; out: hit point */,
               float& kAB, float& kBC, float& kCA)
{
    const oracle_scene& sc = *c.s;
    bestTri = -1;
    float bestTriDist;
    const Vec lightPos = pointHit;
    if (stopAtfirstRayHit) bestTriDist = distancesq(origin, lightPos);
    else bestTriDist = FLT_MAX;

    uint32_t stack[B200R_BVH_STACK_SIZE];
    int stackIdx = 0;
    stack[stackIdx++] = 0;
    while (stackIdx) {
        const b200r_bvhnode& cur = sc.nodes[stack[stackIdx - 1]];
        stackIdx--;
        if (!(cur.a & 0x80000000u)) {
            c.nodeTests++;
            if (RayIntersectsBox(origin, ray, cur)) {
                stack[stackIdx++] = cur.b;   // right
                stack[stackIdx++] = cur.a;   // left (popped first)
            }
        } else {
            c.leafVisits++;
            for (uint32_t i = cur.b; i < cur.b + (cur.a & 0x7fffffffu); i++) {
                const int ti = sc.tri_idx[i];
                const b200r_tri& t = sc.tris[ti];
                c.triTests++;
                if (avoidSelf == ti) continue;
                if (doCulling && !t.two_sided) {
                    Vec fromTriToOrigin = origin; fromTriToOrigin -= Vec(t.center);
                    if (dot(fromTriToOrigin, Vec(t.normal)) < 0) continue;
                }
                float k = dot(Vec(t.normal), ray);
                if (k == 0.0) continue;
                float s = (t.d - dot(Vec(t.normal), origin)) / k;
                if (s <= 0.0) continue;
                if (s <= 1e-5f) continue;            // NUDGE_FACTOR
                Vec hit = ray * s; hit += origin;
                float kt1 = dot(Vec(t.e1), hit) - t.d1; if (kt1 < 0.0) continue;
                float kt2 = dot(Vec(t.e2), hit) - t.d2; if (kt2 < 0.0) continue;
                float kt3 = dot(Vec(t.e3), hit) - t.d3; if (kt3 < 0.0) continue;
                if (stopAtfirstRayHit) {
                    float dist = distancesq(lightPos, hit);
                    if (dist < bestTriDist) return true;
                } else {
                    float hitZ = distancesq(origin, hit);
                    if (hitZ < bestTriDist) {
                        bestTriDist = hitZ; bestTri = ti; pointHit = hit;
                        kAB = kt1; kBC = kt2; kCA = kt3;
                    }
                }
            }
        }
    }
    if (!stopAtfirstRayHit) return bestTri != -1;
    return false;
}

// reference src/Raytracer.cc:315-553
template <bool doCulling>
Pix Raytrace(Ctx& c, Vec origin, Vec ray, int avoidSelf, int depth)
{
    if (depth >= c.maxDepth) return Pix(0.f, 0.f, 0.f);
    const oracle_scene& sc = *c.s;
    if (depth == 0) c.raysPrimary++; else c.raysRefl++;

    int best = -1; Vec hitp; float kAB = 0.f, kBC = 0.f, kCA = 0.f;
    if (!Intersect<false, doCulling>(c, origin, ray, avoidSelf, best, hitp, kAB, kBC, kCA))
        return Pix(0.f, 0.f, 0.f);
    avoidSelf = best;
    const b200r_tri& T = sc.tris[best];
    const b200r_vertex &VA = sc.verts[T.a], &VB = sc.verts[T.b], &VC = sc.verts[T.c];
    Pix color(T.colorf[0], T.colorf[1], T.colorf[2]);

    Vec phongNormal; float ABx = 0, BCx = 0, CAx = 0, area = 1;
    if (c.phong) {
        Vec A(VA.pos), B(VB.pos), C(VC.pos);
        Vec AB = B; AB -= A;
        Vec BC = C; BC -= B;
        Vec crossAB_BC = cross(AB, BC);
        area = crossAB_BC.length();
        ABx = kAB * distance(A, B);
        BCx = kBC * distance(B, C);
        CAx = kCA * distance(C, A);
        Vec nA(VA.nrm); nA *= BCx / area;
        Vec nB(VB.nrm); nB *= CAx / area;
        Vec nC(VC.nrm); nC *= ABx / area;
        phongNormal = nA + nB + nC;
        phongNormal.normalize();
    } else {
        phongNormal = Vec(T.normal);
    }

    if (c.ao) {
        // src/Raytracer.cc:386-417
        int i = 0; float totalLight = 0.f, maxLight = 0.f;
        const int RM2 = RAND_MAX / 2;
        while (i < c.aoSamples) {
            Vec ambientRay = phongNormal;
            ambientRay.v[0] += float(c.rnd() - RM2) / (RM2);
            ambientRay.v[1] += float(c.rnd() - RM2) / (RM2);
            ambientRay.v[2] += float(c.rnd() - RM2) / (RM2);
            float cosangle = dot(ambientRay, phongNormal);
            if (cosangle < 0.f) continue;
            i++;
            maxLight += cosangle;
            ambientRay.normalize();
            Vec temp(hitp);
            temp += ambientRay * 0.15f;      // AMBIENT_RANGE
            int dummy; float k0 = 0;
            c.raysAO++;
            if (!Intersect<true, true>(c, hitp, ambientRay, avoidSelf, dummy, temp, k0, k0, k0))
                totalLight += cosangle;
        }
        color *= (float)((96.f / 255.0) * (totalLight / maxLight));
    } else {
        float coeff;
        if (c.phong)
            coeff = VA.ao * BCx / area + VB.ao * CAx / area + VC.ao * ABx / area;
        else
            coeff = (VA.ao + VB.ao + VC.ao) / 3.f;
        float ambientFactor = (float)((96.f * coeff / 255.0) / 255.0);
        color *= ambientFactor;
    }

    for (uint32_t li = 0; li < c.f->n_lights; li++) {
        Vec light(c.f->lights[li].pos);
        Pix dColor;
        Vec pointToLight = light; pointToLight -= hitp;
        if (c.shadows) {
            float distanceFromLightSq = pointToLight.lengthsq();
            Vec shadowray = pointToLight; shadowray /= sqrtf(distanceFromLightSq);
            int dummy; float k0 = 0; Vec lp = light;
            c.raysShadow++;
            if (Intersect<true, doCulling>(c, hitp, shadowray, avoidSelf, dummy, lp, k0, k0, k0))
                continue;
        }
        pointToLight.normalize();
        float intensity = dot(phongNormal, pointToLight);
        if (intensity < 0.) {
        } else {
            Pix diffuse(T.colorf[0], T.colorf[1], T.colorf[2]);
            diffuse *= (float)(128.f * intensity / 255.);
            dColor += diffuse;
            Vec pointToCamera = c.eye; pointToCamera -= hitp; pointToCamera.normalize();
            Vec half = pointToLight; half += pointToCamera; half.normalize();
            float intensity2 = dot(half, phongNormal);
            if (intensity2 > 0.) {
                intensity2 *= intensity2; intensity2 *= intensity2; intensity2 *= intensity2;
                intensity2 *= intensity2; intensity2 *= intensity2;
                dColor += Pix((unsigned char)(192.f * intensity2),
                              (unsigned char)(192.f * intensity2),
                              (unsigned char)(192.f * intensity2));
            }
        }
        color += dColor;
    }

    if (!c.reflections) return color;
    origin = hitp;
    float c1 = -dot(ray, phongNormal);
    Vec reflected = ray; reflected += phongNormal * (2.0f * c1);
    reflected.normalize();
    return color + Raytrace<true>(c, origin, reflected, avoidSelf, depth + 1) * 0.375f;
}

}  // namespace

namespace oport {

// reference src/Raytracer.cc:555-606 + the scanline loop of Scene::renderRaytracer (:814-838)
void render_raytrace(const oracle_scene* s, const b200r_frame* f, uint32_t* out, b200r_counters* ctr, int threads)
{
    const int W = (int)f->width, H = (int)f->height;
    const bool antialias = (f->mode == B200R_MODE_RAYTRACE_AA);
    const int rowStep = f->row_step ? (int)f->row_step : 1;
    const int nRows = (H - (int)f->row_first + rowStep - 1) / rowStep;
    uint64_t tot[7] = {0, 0, 0, 0, 0, 0, 0};
    const int SCREEN_DIST = H * 2;
#pragma omp parallel num_threads(threads)
    {
        Ctx c;
        c.s = s; c.f = f; c.eye = Vec(f->eye);
        c.shadows = f->flags & B200R_F_SHADOWS; c.reflections = f->flags & B200R_F_REFLECTIONS;
        c.phong = f->flags & B200R_F_PHONG_NORMAL; c.ao = f->flags & B200R_F_AO;
        c.aoSamples = (int)f->ao_samples; c.maxDepth = f->max_depth ? (int)f->max_depth : 3;
        const Vec row1(f->mv), row2(f->mv + 3), row3(f->mv + 6);
#pragma omp for schedule(dynamic, 1)
        for (int r = 0; r < nRows; r++) {
            const int y = (int)f->row_first + r * rowStep;
            for (int x = 0; x < W; x++) {
                c.aoKey = oracle_ao_key(f->frame_index, (uint32_t)x, (uint32_t)y); c.aoCtr = 0;
                Pix finalColor(0, 0, 0);
                int pixelsTraced = antialias ? 4 : 1;
                while (pixelsTraced--) {
                    float xx = (float)x, yy = (float)y;
                    if (antialias) {
                        xx += 0.25f - .5f * (pixelsTraced & 1);
                        yy += 0.25f - .5f * ((pixelsTraced & 2) >> 1);
                    }
                    float lx = float((H / 2) - yy) / SCREEN_DIST;
                    float ly = float(xx - (W / 2)) / SCREEN_DIST;
                    float lz = 1.0;
                    Vec rayCam(lx, ly, lz); rayCam.normalize();
                    Vec rayWorld = row1 * rayCam.v[0];
                    rayWorld += row2 * rayCam.v[1];
                    rayWorld += row3 * rayCam.v[2];
                    rayWorld.normalize();
                    finalColor += Raytrace<true>(c, c.eye, rayWorld, -1, 0);
                }
                if (antialias) finalColor /= 4.f;
                if (finalColor.r > 255.0f) finalColor.r = 255.0f;
                if (finalColor.g > 255.0f) finalColor.g = 255.0f;
                if (finalColor.b > 255.0f) finalColor.b = 255.0f;
                out[(size_t)r * W + x] = map_rgb((uint8_t)finalColor.r, (uint8_t)finalColor.g, (uint8_t)finalColor.b);
            }
        }
#pragma omp critical
        {
            tot[0] += c.raysPrimary; tot[1] += c.raysShadow; tot[2] += c.raysRefl; tot[3] += c.raysAO;
            tot[4] += c.nodeTests; tot[5] += c.leafVisits; tot[6] += c.triTests;
        }
    }
    if (ctr) {
        ctr->rays_primary = tot[0]; ctr->rays_shadow = tot[1]; ctr->rays_reflection = tot[2]; ctr->rays_ao = tot[3];
        ctr->node_tests = tot[4]; ctr->leaf_visits = tot[5]; ctr->tri_tests = tot[6];
    }
}

}  // namespace oport
