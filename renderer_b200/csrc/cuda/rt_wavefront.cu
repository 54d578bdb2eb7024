// rt_wavefront.cu — Raytrace() for the generic configurations (ambient occlusion, reflections, two lights) as a WAVEFRONT:
// every secondary ray goes through the pooled traversal kernel (rt_pool.cu) instead of being walked by the thread that shades
// the hit.
//
// Reference: Raytrace<doCulling> (src/Raytracer.cc:315-553) recurses per pixel: at a hit it casts AMBIENT_SAMPLES occlusion
// rays (:386-417), one shadow ray per light (:440-505) and one reflection ray (:508-541, depth <= 3), each a full BVH walk.
// Round 1 ran that recursion with one thread per primary hit (rt_shade_kernel: ~50 dependent traversals per thread, 97 % of a
// C3 frame). Here the recursion is unrolled over the whole frame, level by level:
//
//   level d:   rt_spawn_kernel    one thread per hit of the level: Phong normal, the AO sample directions (the pixel's random
//                                 stream, consumed in the reference's order), the shadow-ray directions -> 48-byte any-hit ray
//                                 records; the reflection ray of the level -> one closest-hit ray record
//              rt_pool_kernel     POOL_ANYHIT over those records: one byte per ray, "something is in the way"
//              rt_combine_kernel  one thread per hit: the rest of Raytrace() at that hit from the bytes (AO factor = ordered sum
//                                 over the unoccluded samples, lights that are not blocked) -> the level's colour
//              rt_pool_kernel     POOL_CLOSEST over the reflection rays -> the hits of level d + 1
//   finally:   rt_compose_kernel  one thread per primary hit: colour = clamp(c0 + 0.375 clamp(c1 + 0.375 c2)) (Pixel::operator+,
//                                 src/Types.h:137-142), the final clamp and the XRGB store of RaytraceHorizontalSegment (:598-606).
//
// Same rays, same arithmetic, same order of every floating-point sum as shade_hit()/trace() in rt_common.cuh (which stay in use
// for mode 0, counting runs and the `no_wavefront` cross-check): the frames are bit-identical. Hits of a level are processed in
// chunks of at most `chunk` hits so that the ray records of a 4K frame with 16 AO samples stay under 1 GB.
#include "rt_common.cuh"
#include "rt_kernels.cuh"

namespace b200r {
using namespace rt;

namespace {

struct __align__(16) PathRec {          // one per primary hit
    uint32_t pix, nlev, rngCtr, pad;
    float lev[3][3];                     // the colour each level contributes (r, g, b)
    float ray[3];                        // direction of the ray that arrives at the level being processed
};
static_assert(sizeof(PathRec) == 64, "PathRec layout");

struct WfLevel {
    const HitRecord* hits; const unsigned* hitCount;    // hits of this level
    unsigned first, cap;                                 // chunk: hits [first, first + cap)
    int depth;
    unsigned stride, aoN;                                // any-hit rays per hit; how many of them are AO rays
    int emitRefl;
};

__device__ __forceinline__ uint32_t ao_key(const FrameParams& fp, int x, int y)
{
    uint32_t k = mix32(fp.frame_index * 0x9E3779B9u + 0x7F4A7C15u);
    k = mix32(k ^ ((uint32_t)x * 0x85EBCA77u));
    k = mix32(k ^ ((uint32_t)y * 0xC2B2AE3Du));
    return k;
}

// The interpolated normal at a hit (reference src/Raytracer.cc:362-381), exactly as shade_hit() computes it.
__device__ __forceinline__ V3 hit_normal(const DeviceScene& sc, const FrameParams& fp, int tri, float kAB, float kBC, float kCA,
                                         float& ABx, float& BCx, float& CAx, float& area, unsigned ao[3], Pix3& colorf)
{
    const float4* S = sc.shade + 6 * (size_t)tri;
    const float4 s0 = __ldg(S + 0), s1 = __ldg(S + 1), s2 = __ldg(S + 2);
    const float4 s3 = __ldg(S + 3), s4 = __ldg(S + 4), s5 = __ldg(S + 5);
    const V3 A = mkv3(s0.x, s0.y, s0.z), B = mkv3(s0.w, s1.x, s1.y), C = mkv3(s1.z, s1.w, s2.x);
    const V3 nA = mkv3(s2.y, s2.z, s2.w), nB = mkv3(s3.x, s3.y, s3.z), nC = mkv3(s3.w, s4.x, s4.y);
    ao[0] = __float_as_uint(s4.z); ao[1] = __float_as_uint(s4.w); ao[2] = __float_as_uint(s5.x);
    colorf = mkpix(s5.y, s5.z, s5.w);
    ABx = 0.f; BCx = 0.f; CAx = 0.f; area = 1.f;
    if (fp.flags & B200R_F_PHONG_NORMAL) {
        const V3 AB = B - A, BC = C - B;
        area = length3(cross3(AB, BC));
        ABx = kAB * distance3(A, B);
        BCx = kBC * distance3(B, C);
        CAx = kCA * distance3(C, A);
        const V3 pA = nA * (BCx / area), pB = nB * (CAx / area), pC = nC * (ABx / area);
        return normalize3((pA + pB) + pC);
    }
    const float4 nn = __ldg(sc.rtris + 4 * (size_t)tri + 2);
    return mkv3(nn.x, nn.y, nn.z);
}

__global__ void __launch_bounds__(256)
rt_spawn_kernel(DeviceScene sc, FrameParams fp, WfLevel L, PathRec* __restrict__ paths, float4* __restrict__ rays,
                unsigned char* __restrict__ occ, float* __restrict__ cosv, float4* __restrict__ ctx, float4* __restrict__ refl)
{
    const unsigned n = *L.hitCount;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.cap) return;
    const unsigned h = L.first + t;
    if (h >= n) return;
    const float4* src = reinterpret_cast<const float4*>(L.hits + h);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    const int tri = __float_as_int(a.y);
    const V3 hitp = mkv3(a.z, a.w, b.x);
    const unsigned pathId = L.depth == 0 ? h : (unsigned)__float_as_int(a.x);
    PathRec* P = paths + pathId;
    const uint32_t pix = L.depth == 0 ? (uint32_t)__float_as_int(a.x) : P->pix;
    const int x = (int)(pix & 0xffffu), r = (int)(pix >> 16);
    const int y = (int)fp.row_first + r * (int)fp.row_step;
    AoStream rng; rng.key = ao_key(fp, x, y);
    V3 ray;
    if (L.depth == 0) { ray = primary_ray(fp, x, y); rng.ctr = 0u; P->pix = pix; P->nlev = 0u; }
    else { ray = mkv3(P->ray[0], P->ray[1], P->ray[2]); rng.ctr = P->rngCtr; }

    float ABx, BCx, CAx, area; unsigned aoc[3]; Pix3 colorf;
    const V3 phongNormal = hit_normal(sc, fp, tri, b.y, b.z, b.w, ABx, BCx, CAx, area, aoc, colorf);

    float4* out = rays + 3 * (size_t)t * L.stride;
    float maxLight = 0.f;
    if (fp.flags & B200R_F_AO) {
        // reference src/Raytracer.cc:386-417: rejection sampling around the normal, AMBIENT_SAMPLES accepted directions
        int i = 0;
        const int RM2 = 2147483647 / 2;
        while (i < (int)L.aoN) {
            V3 ambientRay = phongNormal;
            ambientRay.x += float(rng.draw() - RM2) / float(RM2);
            ambientRay.y += float(rng.draw() - RM2) / float(RM2);
            ambientRay.z += float(rng.draw() - RM2) / float(RM2);
            const float cosangle = dot3(ambientRay, phongNormal);
            if (cosangle < 0.f) continue;
            maxLight += cosangle;
            ambientRay = normalize3(ambientRay);
            const V3 temp = hitp + ambientRay * 0.15f;   // AMBIENT_RANGE
            const unsigned ri = t * L.stride + (unsigned)i;
            out[3 * i + 0] = make_float4(hitp.x, hitp.y, hitp.z, 0.f);
            out[3 * i + 1] = make_float4(ambientRay.x, ambientRay.y, ambientRay.z, __int_as_float(tri));
            out[3 * i + 2] = make_float4(temp.x, temp.y, temp.z, __uint_as_float(ri));
            cosv[(size_t)t * L.aoN + i] = cosangle;
            occ[ri] = 0;
            i++;
        }
    }
    if (fp.flags & B200R_F_SHADOWS) {
        for (uint32_t li = 0; li < fp.n_lights; li++) {
            const V3 light = mkv3(fp.light_pos[li][0], fp.light_pos[li][1], fp.light_pos[li][2]);
            const V3 pointToLight = light - hitp;
            const float distanceFromLightSq = lengthsq3(pointToLight);
            const V3 shadowray = pointToLight / sqrtf(distanceFromLightSq);
            const unsigned j = L.aoN + li, ri = t * L.stride + j;
            out[3 * j + 0] = make_float4(hitp.x, hitp.y, hitp.z, 0.f);
            out[3 * j + 1] = make_float4(shadowray.x, shadowray.y, shadowray.z, __int_as_float(tri));
            out[3 * j + 2] = make_float4(light.x, light.y, light.z, __uint_as_float(ri));
            occ[ri] = 0;
        }
    }
    ctx[t] = make_float4(phongNormal.x, phongNormal.y, phongNormal.z, maxLight);
    if (L.emitRefl) {
        // reference src/Raytracer.cc:508-519
        const float c1 = -dot3(ray, phongNormal);
        const V3 nray = normalize3(ray + phongNormal * (2.0f * c1));
        float4* ro = refl + 3 * (size_t)h;
        ro[0] = make_float4(hitp.x, hitp.y, hitp.z, 0.f);
        ro[1] = make_float4(nray.x, nray.y, nray.z, __int_as_float(tri));
        ro[2] = make_float4(0.f, 0.f, 0.f, __uint_as_float(pathId));
        P->ray[0] = nray.x; P->ray[1] = nray.y; P->ray[2] = nray.z;
        P->rngCtr = rng.ctr;
    }
}

__global__ void __launch_bounds__(256)
rt_combine_kernel(DeviceScene sc, FrameParams fp, WfLevel L, PathRec* __restrict__ paths, const unsigned char* __restrict__ occ,
                  const float* __restrict__ cosv, const float4* __restrict__ ctx)
{
    const unsigned n = *L.hitCount;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= L.cap) return;
    const unsigned h = L.first + t;
    if (h >= n) return;
    const float4* src = reinterpret_cast<const float4*>(L.hits + h);
    const float4 a = __ldg(src), b = __ldg(src + 1);
    const int tri = __float_as_int(a.y);
    const V3 hitp = mkv3(a.z, a.w, b.x);
    const unsigned pathId = L.depth == 0 ? h : (unsigned)__float_as_int(a.x);
    const V3 eye = mkv3(fp.eye[0], fp.eye[1], fp.eye[2]);

    float ABx, BCx, CAx, area; unsigned aoc[3]; Pix3 colorf;
    (void)hit_normal(sc, fp, tri, b.y, b.z, b.w, ABx, BCx, CAx, area, aoc, colorf);
    const float4 c = ctx[t];
    const V3 phongNormal = mkv3(c.x, c.y, c.z);
    Pix3 color = colorf;
    const unsigned char* myOcc = occ + (size_t)t * L.stride;
    if (fp.flags & B200R_F_AO) {
        float totalLight = 0.f;
        const float maxLight = c.w;
        for (unsigned i = 0; i < L.aoN; i++)
            if (!myOcc[i]) totalLight += cosv[(size_t)t * L.aoN + i];
        // (AMBIENT/255.0)*(totalLight/maxLight): double constant x float quotient, rounded once to float
        const float f = (float)((96.0 / 255.0) * (double)(totalLight / maxLight));
        color.b = f * color.b; color.g = f * color.g; color.r = f * color.r;
    } else {
        float coeff;
        if (fp.flags & B200R_F_PHONG_NORMAL)
            coeff = (float)aoc[0] * BCx / area + (float)aoc[1] * CAx / area + (float)aoc[2] * ABx / area;
        else
            coeff = (float)(aoc[0] + aoc[1] + aoc[2]) / 3.f;
        // (coord)((AMBIENT*coeff/255.0)/255.0): float product, two double divides, one rounding
        const float f = (float)(((double)(96.f * coeff) / 255.0) / 255.0);
        color.b = f * color.b; color.g = f * color.g; color.r = f * color.r;
    }
    for (uint32_t li = 0; li < fp.n_lights; li++) {
        const V3 light = mkv3(fp.light_pos[li][0], fp.light_pos[li][1], fp.light_pos[li][2]);
        Pix3 dColor = mkpix(0.f, 0.f, 0.f);
        V3 pointToLight = light - hitp;
        if ((fp.flags & B200R_F_SHADOWS) && myOcc[L.aoN + li]) continue;           // the light is blocked
        pointToLight = normalize3(pointToLight);
        const float intensity = dot3(phongNormal, pointToLight);
        if (intensity < 0.f) {
        } else {
            // (coord)(DIFFUSE*intensity/255.) == float divide (innocuous double rounding)
            const float df = (128.f * intensity) / 255.f;
            dColor.b += df * colorf.b; dColor.g += df * colorf.g; dColor.r += df * colorf.r;
            const V3 pointToCamera = normalize3(eye - hitp);
            const V3 half = normalize3(pointToLight + pointToCamera);
            float intensity2 = dot3(half, phongNormal);
            if (intensity2 > 0.f) {
                intensity2 *= intensity2; intensity2 *= intensity2; intensity2 *= intensity2;
                intensity2 *= intensity2; intensity2 *= intensity2;
                const float sp = (float)u8_x86(192.f * intensity2);
                dColor.r += sp; dColor.g += sp; dColor.b += sp;
            }
        }
        color.b += dColor.b; color.g += dColor.g; color.r += dColor.r;
    }
    PathRec* P = paths + pathId;
    P->lev[L.depth][0] = color.r; P->lev[L.depth][1] = color.g; P->lev[L.depth][2] = color.b;
    P->nlev = (uint32_t)L.depth + 1u;
}

__global__ void __launch_bounds__(256)
rt_compose_kernel(FrameParams fp, const PathRec* __restrict__ paths, const unsigned* __restrict__ nPaths, uint32_t* __restrict__ out)
{
    const unsigned n = *nPaths;
    for (unsigned p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const PathRec& P = paths[p];
        Pix3 c;
        if (!(fp.flags & B200R_F_REFLECTIONS)) c = P.nlev ? mkpix(P.lev[0][0], P.lev[0][1], P.lev[0][2]) : mkpix(0.f, 0.f, 0.f);
        else {
            // color + Raytrace(depth+1)*0.375 with the clamping Pixel::operator+, innermost level first
            c = mkpix(0.f, 0.f, 0.f);
            for (int k = (int)P.nlev - 1; k >= 0; k--) {
                c.r = clamp255(P.lev[k][0] + 0.375f * c.r);
                c.g = clamp255(P.lev[k][1] + 0.375f * c.g);
                c.b = clamp255(P.lev[k][2] + 0.375f * c.b);
            }
        }
        if (c.r > 255.0f) c.r = 255.0f;
        if (c.g > 255.0f) c.g = 255.0f;
        if (c.b > 255.0f) c.b = 255.0f;
        const uint32_t pix = P.pix;
        out[(size_t)(pix >> 16) * fp.W + (pix & 0xffffu)] = (u8_x86(c.r) << 16) | (u8_x86(c.g) << 8) | u8_x86(c.b);
    }
}

}  // namespace

size_t wavefront_path_bytes() { return sizeof(PathRec); }

// Everything after the primary rays of a generic frame. `counters`: the frame's zeroed counter words ([2] = hits of level 0,
// [4], [5] = hits of levels 1, 2, [8..63] = read cursors of the queue launches).
cudaError_t launch_rt_wavefront(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, RtBuffers& rt, const Switches& sw,
                                bool prune, int numSMs, cudaStream_t stream, int& launches)
{
    const unsigned aoN = (fp.flags & B200R_F_AO) ? fp.ao_samples : 0u;
    const unsigned stride = aoN + ((fp.flags & B200R_F_SHADOWS) ? fp.n_lights : 0u);
    const bool reflections = (fp.flags & B200R_F_REFLECTIONS) != 0;
    const int maxDepth = reflections ? (int)(fp.max_depth < 3 ? fp.max_depth : 3) : 1;
    const unsigned chunk = rt.wfChunk;
    const unsigned px32 = ((fp.W + 7) / 8) * ((fp.n_rows + 3) / 4) * 32u;           // no level has more hits than the frame has pixels
    const unsigned nChunks = (px32 + chunk - 1) / chunk;
    HitRecord* hits[3] = {reinterpret_cast<HitRecord*>(rt.hits), reinterpret_cast<HitRecord*>(rt.wfHits1), reinterpret_cast<HitRecord*>(rt.wfHits2)};
    unsigned* hitCount[3] = {rt.counters + 2, rt.counters + 4, rt.counters + 5};
    PathRec* paths = reinterpret_cast<PathRec*>(rt.wfPaths);
    unsigned cursor = 8;
    cudaError_t e = cudaSuccess;
    for (int d = 0; d < maxDepth; d++) {
        const int emitRefl = reflections && d + 1 < maxDepth ? 1 : 0;
        for (unsigned c = 0; c < nChunks; c++) {
            WfLevel L = {hits[d], hitCount[d], c * chunk, chunk, d, stride, aoN, emitRefl};
            const unsigned blocks = (chunk + 255u) / 256u;
            rt_spawn_kernel<<<blocks, 256, 0, stream>>>(sc, fp, L, paths, rt.wfRays, rt.wfOcc, rt.wfCos, rt.wfCtx, rt.wfRefl);
            launches += 1;
            if (stride) {
                if (cursor >= 64) return cudaErrorInvalidValue;
                e = launch_rt_pool_queue(sc, fp, true, prune, sw, rt.counters + cursor++, rt.wfRays, hitCount[d], c * chunk, chunk, stride,
                                         rt.wfOcc, nullptr, nullptr, numSMs, stream, launches);
                if (e != cudaSuccess) return e;
            }
            rt_combine_kernel<<<blocks, 256, 0, stream>>>(sc, fp, L, paths, rt.wfOcc, rt.wfCos, rt.wfCtx);
            launches += 1;
        }
        if (emitRefl) {
            if (cursor >= 64) return cudaErrorInvalidValue;
            e = launch_rt_pool_queue(sc, fp, false, prune, sw, rt.counters + cursor++, rt.wfRefl, hitCount[d], 0u, 0xFFFFFFFFu, 1u, nullptr,
                                     hits[d + 1], hitCount[d + 1], numSMs, stream, launches);
            if (e != cudaSuccess) return e;
        }
    }
    rt_compose_kernel<<<numSMs * 4, 256, 0, stream>>>(fp, paths, hitCount[0], d_out);
    launches += 1;
    return cudaGetLastError();
}

}  // namespace b200r
