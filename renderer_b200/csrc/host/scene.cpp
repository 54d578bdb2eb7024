// scene.cpp — host plumbing: scene loading with the reference's load-time numerics.
//
// New code (not a copy) that reproduces, operation for operation, what the reference's loader does
// to the numbers the hot path later reads:
//   Scene::load            reference src/Loader.cc:85-494   (.tri :103-222, shadevis .ply :354-409,
//                          recentre/rescale :418-454, intersection precompute :465-493)
//   Scene::fix_normals     reference src/Loader.cc:496-518
//   Triangle::Triangle     reference src/Base3d.cc:27-55    (_center from UNscaled vertices)
// Compiled with -ffp-contract=off so no FMA is ever formed.
#include "scene.h"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>

namespace b200r {

static const uint32_t TRI_MAGIC = 0xDEADBEEFu;        // reference src/Defines.h:23
static const uint32_t TRI_MAGICNORMAL = 0xDEADC0DEu;  // reference src/Defines.h:24
static const float MaxCoordAfterRescale = 1.2f;       // reference src/Loader.cc:77

namespace {

struct Reader {
    const unsigned char* p; size_t n, off;
    bool have(size_t k) const { return off + k <= n; }
    template <class T> T get()
    {
        if (!have(sizeof(T))) throw std::runtime_error("Malformed 3D file");
        T v; memcpy(&v, p + off, sizeof(T)); off += sizeof(T); return v;
    }
};

// Triangle ctor, reference src/Base3d.cc:27-55. The ctor's averaged _normal is always overwritten
// later (fix_normals and/or the intersection precompute), so only _center/_colorf/_color matter.
b200r_tri make_tri(const std::vector<b200r_vertex>& v, uint32_t a, uint32_t b, uint32_t c,
                   unsigned r, unsigned g, unsigned bl)
{
    b200r_tri t;
    memset(&t, 0, sizeof t);
    t.a = a; t.b = b; t.c = c;
    for (int k = 0; k < 3; k++)
        t.center[k] = (v[a].pos[k] + v[b].pos[k] + v[c].pos[k]) / 3.0f;
    t.colorf[0] = (float)r; t.colorf[1] = (float)g; t.colorf[2] = (float)bl;
    // SDL_MapRGB takes Uint8 components: the unsigned values are narrowed mod 256
    t.color = ((uint32_t)(uint8_t)r << 16) | ((uint32_t)(uint8_t)g << 8) | (uint32_t)(uint8_t)bl;
    t.two_sided = 0;
    return t;
}

void load_tri(Scene& s, const std::string& filename)
{
    FILE* fp = fopen(filename.c_str(), "rb");
    if (!fp) throw std::runtime_error("File '" + filename + "' not found!");
    std::vector<unsigned char> buf;
    fseek(fp, 0, SEEK_END); long sz = ftell(fp); fseek(fp, 0, SEEK_SET);
    buf.resize(sz > 0 ? (size_t)sz : 0);
    if (sz > 0 && fread(buf.data(), 1, (size_t)sz, fp) != (size_t)sz) { fclose(fp); throw std::runtime_error("Malformed 3D file"); }
    fclose(fp);

    Reader rd{buf.data(), buf.size(), 0};
    uint32_t magic = rd.get<uint32_t>();
    if (magic != TRI_MAGIC && magic != TRI_MAGICNORMAL) rd.off = 0;   // Loader.cc:128-131
    const bool hasNormals = (magic == TRI_MAGICNORMAL);
    const bool hasColors = (magic == TRI_MAGIC || magic == TRI_MAGICNORMAL);

    uint32_t totalPoints = 0;
    // Loader.cc:164-219: blocks until EOF; vertex indices are GLOBAL (not per block)
    while (rd.have(4)) {
        uint32_t noOfPoints = rd.get<uint32_t>();
        for (uint32_t i = 0; i < noOfPoints; i++) {
            b200r_vertex v; memset(&v, 0, sizeof v);
            v.pos[0] = rd.get<float>(); v.pos[1] = rd.get<float>(); v.pos[2] = rd.get<float>();
            if (hasNormals) { v.nrm[0] = rd.get<float>(); v.nrm[1] = rd.get<float>(); v.nrm[2] = rd.get<float>(); }
            v.ao = 60;   // Vertex ctor default, Base3d.h:32
            s.verts.push_back(v);
        }
        uint32_t noOfTris = rd.get<uint32_t>();
        for (uint32_t i = 0; i < noOfTris; i++) {
            uint32_t i1 = rd.get<uint32_t>(), i2 = rd.get<uint32_t>(), i3 = rd.get<uint32_t>();
            if (i1 >= totalPoints + noOfPoints) throw std::runtime_error("Malformed 3D file (idx1)");
            if (i2 >= totalPoints + noOfPoints) throw std::runtime_error("Malformed 3D file (idx2)");
            if (i3 >= totalPoints + noOfPoints) throw std::runtime_error("Malformed 3D file (idx3)");
            float r, g, b;
            if (hasColors) {
                r = rd.get<float>(); g = rd.get<float>(); b = rd.get<float>();
                // "r*=255." : float * double constant, rounded back to float (Loader.cc:205)
                r = (float)((double)r * 255.); g = (float)((double)g * 255.); b = (float)((double)b * 255.);
            } else {
                r = g = b = 255.0f;
            }
            s.tris.push_back(make_tri(s.verts, i1, i2, i3, (unsigned)r, (unsigned)g, (unsigned)b));
        }
        totalPoints += noOfPoints;
    }
    if (!hasNormals) s.fix_normals();   // Loader.cc:221-222
}

void load_ply(Scene& s, const std::string& filename)
{
    // Loader.cc:354-409 — only shadevis-generated objects, same iostream extraction semantics.
    std::ifstream file(filename.c_str(), std::ios::in);
    if (!file) throw std::runtime_error("Missing " + filename);
    std::string line;
    unsigned totalVertices = 0, totalTriangles = 0;
    bool inside = false;
    while (getline(file, line)) {
        if (!inside) {
            if (line.substr(0, 14) == "element vertex") {
                std::istringstream str(line); std::string w; str >> w; str >> w; str >> totalVertices;
                s.verts.reserve(totalVertices);
            } else if (line.substr(0, 12) == "element face") {
                std::istringstream str(line); std::string w; str >> w; str >> w; str >> totalTriangles;
            } else if (line.substr(0, 10) == "end_header")
                inside = true;
        } else {
            if (totalVertices) {
                totalVertices--;
                float x, y, z; unsigned ao;
                std::istringstream str(line);
                str >> x >> y >> z >> ao;
                b200r_vertex v; memset(&v, 0, sizeof v);
                v.pos[0] = x; v.pos[1] = y; v.pos[2] = z;
                v.ao = (unsigned)(unsigned char)ao;   // Vertex ctor takes "unsigned char amb" (Base3d.h:32)
                s.verts.push_back(v);
            } else if (totalTriangles) {
                totalTriangles--;
                unsigned dummy, i1, i2, i3;
                std::istringstream str(line);
                if (str >> dummy >> i1 >> i2 >> i3) {
                    unsigned r, g, b;
                    if (str >> r >> g >> b) {} else { r = 255; g = 255; b = 255; }
                    if (i1 >= s.verts.size() || i2 >= s.verts.size() || i3 >= s.verts.size())
                        throw std::runtime_error("Malformed 3D file (ply index)");
                    s.tris.push_back(make_tri(s.verts, i1, i2, i3, r, g, b));
                }
            }
        }
    }
    s.fix_normals();
}

inline float fmin2(float a, float b) { return b < a ? b : a; }   // std::min(a,b)
inline float fmax2(float a, float b) { return a < b ? b : a; }   // std::max(a,b)

}  // namespace

// reference src/Loader.cc:496-518
void Scene::fix_normals()
{
    for (size_t j = 0; j < tris.size(); j++) {
        b200r_tri& t = tris[j];
        V3 A = v3_from(verts[t.a].pos), B = v3_from(verts[t.b].pos), C = v3_from(verts[t.c].pos);
        V3 cr = normalize3(cross3(B - A, C - A));
        t.normal[0] = cr.x; t.normal[1] = cr.y; t.normal[2] = cr.z;
        uint32_t idx[3] = {t.a, t.b, t.c};
        for (int k = 0; k < 3; k++) {          // sequential: a vertex used twice accumulates twice
            float* n = verts[idx[k]].nrm;
            n[0] += cr.x; n[1] += cr.y; n[2] += cr.z;
        }
    }
    // every vertex normal is normalised once per incident triangle corner (not idempotent in fp32)
    for (size_t j = 0; j < tris.size(); j++) {
        uint32_t idx[3] = {tris[j].a, tris[j].b, tris[j].c};
        for (int k = 0; k < 3; k++) {
            float* n = verts[idx[k]].nrm;
            V3 v = normalize3(v3_from(n));
            n[0] = v.x; n[1] = v.y; n[2] = v.z;
        }
    }
}

void Scene::load(const std::string& filename)
{
    verts.clear(); tris.clear(); nodes.clear(); tri_idx.clear(); bvh_depth = -1;
    size_t dot = filename.rfind('.');
    if (dot == std::string::npos)
        throw std::runtime_error("No extension in filename (only .tri or .ply accepted)");
    std::string ext = filename.substr(dot + 1);
    if (ext == "tri") load_tri(*this, filename);
    else if (ext == "ply" || ext == "PLY") load_ply(*this, filename);
    else throw std::runtime_error("Unknown extension (only .tri or .ply accepted)");
    finish_load();
}

// reference src/Loader.cc:418-493
void Scene::finish_load()
{
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (size_t i = 0; i < tris.size(); i++) {
        uint32_t idx[3] = {tris[i].a, tris[i].b, tris[i].c};
        for (int k = 0; k < 3; k++)
            for (int c = 0; c < 3; c++) {
                mn[c] = fmin2(mn[c], verts[idx[k]].pos[c]);
                mx[c] = fmax2(mx[c], verts[idx[k]].pos[c]);
            }
    }
    float ctr[3];
    for (int c = 0; c < 3; c++) ctr[c] = (mx[c] + mn[c]) / 2;
    for (int c = 0; c < 3; c++) { mn[c] -= ctr[c]; mx[c] -= ctr[c]; }
    float maxi = 0;
    for (int c = 0; c < 3; c++) maxi = fmax2(maxi, (float)fabs(mn[c]));   // fabs(float)->double, exact
    for (int c = 0; c < 3; c++) maxi = fmax2(maxi, (float)fabs(mx[c]));
    const float scale = MaxCoordAfterRescale / maxi;
    for (size_t i = 0; i < verts.size(); i++)
        for (int c = 0; c < 3; c++) { verts[i].pos[c] -= ctr[c]; verts[i].pos[c] *= scale; }
    for (size_t i = 0; i < tris.size(); i++)
        for (int c = 0; c < 3; c++) { tris[i].center[c] -= ctr[c]; tris[i].center[c] *= scale; }

    // intersection precompute (Loader.cc:465-493)
    for (size_t i = 0; i < tris.size(); i++) {
        b200r_tri& t = tris[i];
        V3 A = v3_from(verts[t.a].pos), B = v3_from(verts[t.b].pos), C = v3_from(verts[t.c].pos);
        V3 vc1 = B - A, vc2 = C - B, vc3 = A - C;
        V3 n = cross3(vc1, vc2);
        V3 alt1 = cross3(vc2, vc3);
        if (length3(alt1) > length3(n)) n = alt1;
        V3 alt2 = cross3(vc3, vc1);
        if (length3(alt2) > length3(n)) n = alt2;
        n = normalize3(n);
        t.normal[0] = n.x; t.normal[1] = n.y; t.normal[2] = n.z;
        t.d = dot3(n, A);
        V3 e1 = normalize3(cross3(n, vc1)); t.d1 = dot3(e1, A);
        V3 e2 = normalize3(cross3(n, vc2)); t.d2 = dot3(e2, B);
        V3 e3 = normalize3(cross3(n, vc3)); t.d3 = dot3(e3, C);
        t.e1[0] = e1.x; t.e1[1] = e1.y; t.e1[2] = e1.z;
        t.e2[0] = e2.x; t.e2[1] = e2.y; t.e2[2] = e2.z;
        t.e3[0] = e3.x; t.e3[1] = e3.y; t.e3[2] = e3.z;
    }
}

}  // namespace b200r
