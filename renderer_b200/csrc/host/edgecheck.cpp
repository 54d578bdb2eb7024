// edgecheck.cpp - scene validation behind the ray tracer's distance pruning (see rt_kernels.cu "Distance pruning").
//
// The reference accepts a hit when s > 1e-5 and the three edge tests dot(e_i, hit) - d_i >= 0 pass
// (src/Raytracer.cc:263-275), with e_i, d_i precomputed at load (src/Loader.cc:465-493). The accepted region of a
// triangle is therefore {p in its plane : e_i.p >= d_i}. Pruning by box distance is only sound if that region really
// is the triangle, i.e. its corners (pairwise edge-line intersections, solved here in fp64) coincide with the
// vertices to within `tol`. A triangle whose PLANE (normal, d) is not finite is never accepted by the reference's
// tests (k, s, hit, hitZ all become NaN and every `<` is false) and is harmless; any other non-finite field makes
// the region unbounded and counts as bad.
#include <cmath>

#include "../../../include/b200render.h"

namespace b200r {

static bool solve3(const double m[3][3], const double r[3], double out[3])
{
    const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                       m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    if (!(std::fabs(det) > 1e-6)) return false;
    for (int c = 0; c < 3; c++) {
        double a[3][3];
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = (j == c) ? r[i] : m[i][j];
        out[c] = (a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                  a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0])) / det;
    }
    return true;
}

uint32_t count_unbounded_triangles(const b200r_vertex* verts, uint32_t n_verts, const b200r_tri* tris, uint32_t n_tris, double tol,
                                   unsigned char* bad_out)
{
    uint32_t bad = 0;
    if (bad_out) for (uint32_t i = 0; i < n_tris; i++) bad_out[i] = 0;
    for (uint32_t i = 0; i < n_tris; i++) {
        const b200r_tri& t = tris[i];
        if (t.a >= n_verts || t.b >= n_verts || t.c >= n_verts) { bad++; if (bad_out) bad_out[i] = 1; continue; }
        if (!(std::isfinite(t.normal[0]) && std::isfinite(t.normal[1]) && std::isfinite(t.normal[2]) && std::isfinite(t.d)))
            continue;                                               // never accepted: harmless
        const float* e[3] = {t.e1, t.e2, t.e3};
        const float dd[3] = {t.d1, t.d2, t.d3};
        bool ok = true;
        for (int k = 0; k < 3; k++)
            ok = ok && std::isfinite(e[k][0]) && std::isfinite(e[k][1]) && std::isfinite(e[k][2]) && std::isfinite(dd[k]);
        // corner of edge planes (1,3) = A, (1,2) = B, (2,3) = C   (e1 <-> AB, e2 <-> BC, e3 <-> CA)
        const int pairs[3][2] = {{0, 2}, {0, 1}, {1, 2}};
        const float* V[3] = {verts[t.a].pos, verts[t.b].pos, verts[t.c].pos};
        for (int c = 0; c < 3 && ok; c++) {
            double m[3][3], r[3], p[3];
            for (int j = 0; j < 3; j++) { m[0][j] = t.normal[j]; m[1][j] = e[pairs[c][0]][j]; m[2][j] = e[pairs[c][1]][j]; }
            r[0] = t.d; r[1] = dd[pairs[c][0]]; r[2] = dd[pairs[c][1]];
            if (!solve3(m, r, p)) { ok = false; break; }
            for (int j = 0; j < 3; j++) if (!(std::fabs(p[j] - (double)V[c][j]) <= tol)) ok = false;
        }
        if (!ok) { bad++; if (bad_out) bad_out[i] = 1; }
    }
    return bad;
}

}  // namespace b200r
