// vec.h — 3-vector helpers whose operation ORDER matches the reference's expressions, so that
// fp32 results are bit-identical when compiled without FMA contraction (host: -ffp-contract=off,
// device: -fmad=false) and with IEEE div/sqrt (nvcc defaults -prec-div=true -prec-sqrt=true).
// Follows reference src/Types.h:32-117 (Vector3) and src/Algebra.h:26-78.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define B2_HD __host__ __device__ __forceinline__
#else
#define B2_HD inline
#endif

struct V3 { float x, y, z; };

B2_HD V3 mkv3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
B2_HD V3 v3_from(const float* p) { return mkv3(p[0], p[1], p[2]); }
B2_HD V3 operator+(V3 a, V3 b) { return mkv3(a.x + b.x, a.y + b.y, a.z + b.z); }
B2_HD V3 operator-(V3 a, V3 b) { return mkv3(a.x - b.x, a.y - b.y, a.z - b.z); }
B2_HD V3 operator*(V3 a, float s) { return mkv3(a.x * s, a.y * s, a.z * s); }
B2_HD V3 operator/(V3 a, float s) { return mkv3(a.x / s, a.y / s, a.z / s); }

// Algebra.h:74-77: l._x*r._x + l._y*r._y + l._z*r._z  == ((xx + yy) + zz)
B2_HD float dot3(V3 l, V3 r) { return l.x * r.x + l.y * r.y + l.z * r.z; }
// Types.h:66-69
B2_HD float lengthsq3(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }
// Types.h:61-64: sqrt on coord(float) resolves to the float overload (sqrtss)
B2_HD float length3(V3 v) { return sqrtf(v.x * v.x + v.y * v.y + v.z * v.z); }
// Types.h:72-76: three IEEE divides by the length, not a reciprocal multiply
B2_HD V3 normalize3(V3 v) { float n = length3(v); return mkv3(v.x / n, v.y / n, v.z / n); }
// Algebra.h:44-50
B2_HD float distancesq3(V3 a, V3 b)
{
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return dx * dx + dy * dy + dz * dz;
}
// Algebra.h:52-58
B2_HD float distance3(V3 a, V3 b) { return sqrtf(distancesq3(a, b)); }
// Algebra.h:60-72:  x=aay*bbz-bby*aaz; y=bbx*aaz-aax*bbz; z=aax*bby-aay*bbx
B2_HD V3 cross3(V3 l, V3 r)
{
    return mkv3(l.y * r.z - r.y * l.z, r.x * l.z - l.x * r.z, l.x * r.y - l.y * r.x);
}
// Algebra.h:26-35 Matrix3::multiplyRightWith (rows r1,r2,r3 stored as m[0..8])
B2_HD V3 mat3_mul(const float* m, V3 r)
{
    return mkv3(m[0] * r.x + m[1] * r.y + m[2] * r.z,
                m[3] * r.x + m[4] * r.y + m[5] * r.z,
                m[6] * r.x + m[7] * r.y + m[8] * r.z);
}
// Algebra.h:38-42 Transform(worldPoint, origin, mv)
B2_HD V3 transform3(V3 p, V3 origin, const float* mv) { return mat3_mul(mv, p - origin); }
