// port_common.h — value types of the CPU restatement. TEST INFRASTRUCTURE ONLY (see oracle_port.h).
// Vec  ~ reference Vector3 (src/Types.h:32-117), Pix ~ reference Pixel (src/Types.h:121-144),
// free functions ~ src/Algebra.h:44-78. Same member-wise operation order as the reference.
#pragma once
#include <math.h>
#include <stdint.h>
#include "oracle_port.h"

namespace oport {

struct Vec {
    float v[3];
    Vec() { v[0] = v[1] = v[2] = 0.f; }
    Vec(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
    explicit Vec(const float* p) { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; }
    float length() const { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    float lengthsq() const { return v[0] * v[0] + v[1] * v[1] + v[2] * v[2]; }
    void normalize() { float n = length(); v[0] /= n; v[1] /= n; v[2] /= n; }
    Vec& operator+=(const Vec& r) { v[0] += r.v[0]; v[1] += r.v[1]; v[2] += r.v[2]; return *this; }
    Vec& operator-=(const Vec& r) { v[0] -= r.v[0]; v[1] -= r.v[1]; v[2] -= r.v[2]; return *this; }
    Vec& operator*=(float r) { v[0] *= r; v[1] *= r; v[2] *= r; return *this; }
    Vec& operator/=(float r) { v[0] /= r; v[1] /= r; v[2] /= r; return *this; }
    Vec operator*(float r) const { return Vec(v[0] * r, v[1] * r, v[2] * r); }
    Vec operator+(const Vec& r) const { return Vec(v[0] + r.v[0], v[1] + r.v[1], v[2] + r.v[2]); }
};

inline float dot(const Vec& l, const Vec& r) { return l.v[0] * r.v[0] + l.v[1] * r.v[1] + l.v[2] * r.v[2]; }
inline float distancesq(const Vec& a, const Vec& b)
{
    float dx = a.v[0] - b.v[0], dy = a.v[1] - b.v[1], dz = a.v[2] - b.v[2];
    return dx * dx + dy * dy + dz * dz;
}
inline float distance(const Vec& a, const Vec& b)
{
    float dx = a.v[0] - b.v[0], dy = a.v[1] - b.v[1], dz = a.v[2] - b.v[2];
    return sqrtf(dx * dx + dy * dy + dz * dz);
}
inline Vec cross(const Vec& l, const Vec& r)
{
    return Vec(l.v[1] * r.v[2] - r.v[1] * l.v[2],
               r.v[0] * l.v[2] - l.v[0] * r.v[2],
               l.v[0] * r.v[1] - l.v[1] * r.v[0]);
}
inline Vec matmul(const float* m, const Vec& r)   // Matrix3::multiplyRightWith, src/Algebra.h:28-34
{
    return Vec(m[0] * r.v[0] + m[1] * r.v[1] + m[2] * r.v[2],
               m[3] * r.v[0] + m[4] * r.v[1] + m[5] * r.v[2],
               m[6] * r.v[0] + m[7] * r.v[1] + m[8] * r.v[2]);
}
inline Vec xform(Vec p, const Vec& origin, const float* mv) { p -= origin; return matmul(mv, p); }   // Transform

struct Pix {
    float b, g, r;
    Pix(float r_ = 0.f, float g_ = 0.f, float b_ = 0.f) : b(b_), g(g_), r(r_) {}
    Pix& operator+=(const Pix& o) { b += o.b; g += o.g; r += o.r; return *this; }
    Pix& operator-=(const Pix& o) { b -= o.b; g -= o.g; r -= o.r; return *this; }
    Pix& operator*=(float s) { b = s * b; g = s * g; r = s * r; return *this; }
    Pix& operator/=(float s) { b = b / s; g = g / s; r = r / s; return *this; }
    Pix operator+(const Pix& o) const   // the clamping operator+, src/Types.h:137-142
    {
        float rr = r + o.r; if (rr < 0.f) rr = 0.f; if (rr > 255.f) rr = 255.f;
        float gg = g + o.g; if (gg < 0.f) gg = 0.f; if (gg > 255.f) gg = 255.f;
        float bb = b + o.b; if (bb < 0.f) bb = 0.f; if (bb > 255.f) bb = 255.f;
        return Pix(rr, gg, bb);
    }
    Pix operator*(float s) const { return Pix(s * r, s * g, s * b); }
};

inline uint32_t map_rgb(uint8_t r, uint8_t g, uint8_t b) { return ((uint32_t)r << 16) | ((uint32_t)g << 8) | b; }

void render_raytrace(const oracle_scene*, const b200r_frame*, uint32_t*, b200r_counters*, int threads);
void render_raster(const oracle_scene*, const b200r_frame*, uint32_t*, b200r_counters*, int threads);
void render_points(const oracle_scene*, const b200r_frame*, uint32_t*, b200r_counters*);
void render_wireframe(const oracle_scene*, const b200r_frame*, uint32_t*, b200r_counters*);

}  // namespace oport
