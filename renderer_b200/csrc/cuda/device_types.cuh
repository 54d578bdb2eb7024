// device_types.cuh — device-side layout of the scene and per-frame parameters.
//
// HBM layout (uploaded once by b200r_upload_scene, see DESIGN.md "data layout"):
//   wnodes     4 x float4 (64 B) per INNER BVH node: the boxes of its two children side by side, so one fetch
//              feeds both child slab tests:  {L.lo.x, L.hi.x, R.lo.x, R.hi.x} {..y..} {..z..} {Lref, Rref, 0, 0}
//              ref = index of the child's own record if the child is an inner node, else 0x80000000 | start of the
//              child leaf in the list below (0xFFFFFFFF for an empty leaf). The reference tests a node's box when
//              it is popped and never tests leaf boxes (src/Raytracer.cc:224-229); testing a child's box at push
//              time instead visits exactly the same leaves. The root's own box lives in DeviceScene (root_lo/hi).
//   leaftris   5 x float4 per entry of the triangle index list, IN LIST ORDER (the triIdx indirection of
//              reference src/Raytracer.cc:240 is baked out; order inside each leaf is preserved):
//                {n.xyz, d} {e1.xyz, d1} {e2.xyz, d2} {e3.xyz, d3}
//                {center.xyz, bits(twoSided<<31 | lastInLeaf<<30 | triIndex)}
//   shade      6 x float4 per triangle (by triangle index): A, B, C positions, nA, nB, nC, ao[3], colorf[3]
//   rverts     vertices as {pos.xyz, bits(ao)} {nrm.xyz, 0}  (rasteriser / points)
//   rtris      per-triangle raster record (rasteriser setup)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../../include/b200render.h"
#include "../vec.h"

namespace b200r {

struct DeviceScene {
    const float4* wnodes;     // 4 per inner node
    const float4* leaftris;   // 5 per list entry
    const float4* shade;      // 6 per triangle
    const float4* rverts;     // 2 per vertex
    const float4* rtris;      // 4 per triangle: {a,b,c,two_sided} {center, color bits} {normal, 0} {colorf, 0}
    const float*  shadowmap[B200R_MAX_LIGHTS];
    uint32_t n_nodes, n_list, n_tris, n_verts;
    uint32_t root_ref;        // ref of the root (inner record 0, or a leaf ref for tiny scenes)
    uint32_t fast_div_ok;     // every node bound is 0 or in [2^-35, 2^50]: precondition of the shared-reciprocal divide
    uint32_t prune_ok;        // every triangle's edge planes bound it to within 2e-5 (checked in fp64 at upload): distance pruning allowed
    float root_lo[3], root_hi[3];
};

struct FrameParams {
    uint32_t mode, W, H;
    uint32_t n_lights, flags, ao_samples, max_depth, frame_index;
    uint32_t row_first, row_step, n_rows;     // virtual row r <-> y = row_first + r*row_step
    float eye[3];
    float mv[9];
    float light_pos[B200R_MAX_LIGHTS][3];
    float light_cam[B200R_MAX_LIGHTS][3];
    float cam2light[B200R_MAX_LIGHTS][9];
};

struct DeviceCounters {
    unsigned long long v[11];   // same order as b200r_counters
};
enum { C_RAYS_PRIMARY = 0, C_RAYS_SHADOW, C_RAYS_REFL, C_RAYS_AO, C_NODE_TESTS, C_LEAF_VISITS, C_TRI_TESTS,
       C_TRIS_SETUP, C_SPANS, C_Z_TESTS, C_Z_PASSES };

// x86 cvttss2si semantics: out-of-range and NaN give INT_MIN ("integer indefinite"); CUDA's own
// float->int conversion saturates and maps NaN to 0 (SURVEY.md §8a "float -> 8-bit casts").
__device__ __forceinline__ int cvtt_x86(float f)
{
    return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000;
}
// (unsigned char)f / (Uint8)f as g++ compiles it on x86-64: cvttss2si to 32 bits, keep the low byte.
__device__ __forceinline__ unsigned u8_x86(float f) { return (unsigned)cvtt_x86(f) & 0xFFu; }

__device__ __forceinline__ uint32_t mix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

}  // namespace b200r
