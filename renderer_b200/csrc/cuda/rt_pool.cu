// rt_pool.cu — the ray tracer's hot kernel: BVH traversal with a warp-wide WORK POOL in shared memory.
//
// What it computes is BVH_IntersectTriangles (reference src/Raytracer.cc:183-308) for the primary ray of every pixel and,
// in the common configuration (one light, no reflections, no AO), the shading of the hit and its shadow ray
// (Raytrace, reference src/Raytracer.cc:315-505) - the same rays, the same box and triangle arithmetic (rt_common.cuh), the
// same winner per ray. How the work is scheduled has nothing in common with the reference's one-ray-one-stack loop:
//
//   * A warp owns up to 32 rays at a time ("slots": origin, direction, refined reciprocals, best hit so far, pending count -
//     all in shared memory) and ONE pool of pending BVH nodes for all of them. A pool entry is 8 bytes:
//     {leaf flag | slot | node or list index, entry distance of the node's box}.
//   * Every iteration the 32 lanes pop the top 32 entries - whichever rays they belong to - and each lane processes one node:
//     one 64-byte record holds both children's boxes, so a lane does the two slab tests (true IEEE quotients, the reference's
//     per-axis rules) and pushes the surviving children back, far child first. Positions come from warp ballots; there is
//     no per-ray stack and no lane is ever tied to a ray. Leaves go to a second pool and are intersected 32 at a time.
//   * A ray's pending subtrees are therefore walked by as many lanes as the pool can feed: a ray that crosses hundreds of
//     boxes (C2's horizon pixels: 280 dependent steps in round 1's lane-per-ray kernel, the whole frame's critical path)
//     finishes in a few dozen iterations, and the lanes never idle behind the longest ray of their warp.
//   * The closest hit of a ray is the minimum of (hitZ, position in the triangle list) over every triangle of every leaf
//     the reference would reach - the reference's strict `<` in its list-order visit (src/Raytracer.cc:287-296) - so the order
//     in which lanes find hits does not matter: one 64-bit atomicMin per improving hit on the slot's key in shared memory.
//     Subtrees that can no longer win are dropped when pushed and again when popped (distance pruning, see
//     rt_common.cuh/primary pruning contract in DESIGN.md section 4); shadow rays stop at the first occluder.
//   * A slot whose pending count reaches zero is resolved: black, or (FUSED) shaded with both outcomes of the light test and
//     re-armed as the SHADOW ray of its hit, or (generic configurations) appended as a hit record for rt_shade_kernel.
//     Resolution is deferred until 8 slots wait (or nothing else is left) so the shading code runs with more than one lane.
//   * Persistent CTAs (3 per SM x 148), 8 independent warps each - no CTA-wide barrier anywhere. Warps pull 8x4-pixel tiles of
//     the screen rectangle that can contain the model (centre-out), build the primary rays themselves and test the root box
//     against kernel arguments; the frame is cleared beforehand, so the 81 % of C2's pixels that miss are never touched.
//   * The pool cannot overflow: when it is nearly full the top 32 entries are walked depth-first by their lanes with a private
//     stack (same tests, same merges) instead of being expanded.
#include "rt_common.cuh"
#include "rt_kernels.cuh"

namespace b200r {
using namespace rt;

namespace {

constexpr int POOL_WARPS = 8;            // warps per CTA (independent of each other)
constexpr int POOL_CTAS_PER_SM = 3;
constexpr int CAP_L = 128;               // leaf pool: < 32 waiting + at most 64 pushed per inner iteration (+ 32 roots)
constexpr int LEAF_MIN = 32;             // run a leaf iteration as soon as this many leaves wait
constexpr int SHADE_MIN = 8;             // resolve finished slots once this many wait
constexpr int REFILL_MIN = 8;            // take new pixels once this many slots are free ...
constexpr int LOW_WATER = 48;            // ... and fewer than this many inner entries are pending
constexpr uint32_t ITEM_LEAF = 0x80000000u;
constexpr uint32_t ITEM_INDEX_MASK = 0x03FFFFFFu;     // 26 bits: inner record id or list position
constexpr int ITEM_SLOT_SHIFT = 26;
constexpr unsigned long long KEY_EMPTY = ((unsigned long long)0x7F7FFFFFu << 32) | 0xFFFFFFFFull;   // (FLT_MAX, no list position)

template <int CAP_I>
struct __align__(16) WarpPool {
    uint2 ipool[CAP_I];                  // pending inner nodes of all slots
    uint2 lpool[CAP_L];                  // pending leaves
    float4 ro[32];                       // slot: ray origin, pruning slack (+inf: never prune)
    float4 rd[32];                       // slot: ray direction, w = 1 if the shared-reciprocal divide applies
    float4 rr[32];                       // slot: refined reciprocals of the direction, w = bits(triangle to skip, or -1)
    unsigned long long key[32];          // slot: (bits(best hitZ) << 32) | list position; shadow slots: (bits(light distance^2) << 32)
    int pend[32];                        // slot: pool entries not yet processed
    uint32_t pix[32];                    // slot: (packed row << 16) | x
    uint32_t lit[32], shd[32];           // shadow slots: the two possible pixel words
    uint32_t state[32];                  // bit 0: shadow ray, bit 1: occluded
};

__device__ __forceinline__ uint32_t make_item(uint32_t ref, uint32_t slot)
{
    return (ref & (ITEM_LEAF | ITEM_INDEX_MASK)) | (slot << ITEM_SLOT_SHIFT);
}

struct SlotRay { RayPrep rp; float slack; int avoid; };

template <class POOL>
__device__ __forceinline__ SlotRay load_slot_ray(const POOL& P, uint32_t slot, const float4& ro)
{
    const float4 rd = P.rd[slot], rr = P.rr[slot];
    SlotRay s;
    s.rp.o = mkv3(ro.x, ro.y, ro.z); s.rp.d = mkv3(rd.x, rd.y, rd.z); s.rp.r = mkv3(rr.x, rr.y, rr.z);
    s.rp.fast = rd.w != 0.f;
    s.slack = ro.w;
    s.avoid = __float_as_int(rr.w);
    return s;
}

__device__ __forceinline__ bool pruned(float tnear, float slack, float best)
{
    const float e = tnear - slack;
    return e > 0.f && (e * e) * 0.99999f > best;
}

// The triangles of one leaf against one ray, in list order (reference src/Raytracer.cc:235-298). Closest-hit rays fold
// improving hits into `bestK`; shadow rays return true at the first triangle that is nearer to the light than the origin is.
__device__ __forceinline__ bool intersect_leaf(const DeviceScene& sc, const RayPrep& rp, uint32_t li, bool isShadow, int avoid,
                                               const V3& lightPos, float lightDistSq, unsigned long long& bestK)
{
    const float4* rec = sc.leaftris + 5 * (size_t)li;
    for (;; rec += 5, li++) {
        const float4 q4 = __ldg(rec + 4), q0 = __ldg(rec + 0), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2), q3 = __ldg(rec + 3);
        const uint32_t tw = __float_as_uint(q4.w);
        const bool last = (tw & 0x40000000u) != 0;
        const V3 n = mkv3(q0.x, q0.y, q0.z);
        bool alive = !(isShadow && (int)(tw & 0x3fffffffu) == avoid);      // avoidSelf
        if (alive && !(tw & 0x80000000u)) {                                 // doCulling && !twoSided
            const V3 fromTriToOrigin = rp.o - mkv3(q4.x, q4.y, q4.z);
            if (dot3(fromTriToOrigin, n) < 0.f) alive = false;
        }
        if (alive) {
            const float k = dot3(n, rp.d);
            if (k != 0.f) {
                const float s = (q0.w - dot3(n, rp.o)) / k;
                if (s > 0.f && s > 1e-5f) {                                 // behind the origin / NUDGE_FACTOR
                    const V3 hit = rp.d * s + rp.o;
                    const float kt1 = dot3(mkv3(q1.x, q1.y, q1.z), hit) - q1.w;
                    if (!(kt1 < 0.f)) {
                        const float kt2 = dot3(mkv3(q2.x, q2.y, q2.z), hit) - q2.w;
                        if (!(kt2 < 0.f)) {
                            const float kt3 = dot3(mkv3(q3.x, q3.y, q3.z), hit) - q3.w;
                            if (!(kt3 < 0.f)) {
                                if (isShadow) {
                                    if (distancesq3(lightPos, hit) < lightDistSq) return true;
                                } else {
                                    const float hitZ = distancesq3(rp.o, hit);
                                    if (hitZ < FLT_MAX) {                   // the reference starts from FLT_MAX with a strict `<`
                                        const unsigned long long k64 = ((unsigned long long)__float_as_uint(hitZ) << 32) | li;
                                        if (k64 < bestK) bestK = k64;
                                    }
                                }
                            }
                        }
                    }
                }
            }
        }
        if (last) return false;
    }
}

// Both children of inner record `id` against the ray: which survive, and the entry distances of their boxes.
// A leaf child has no box test in the reference (src/Raytracer.cc:224-229): it survives unless empty, and carries `tHere`, the
// entry distance of the node being processed (its triangles lie inside that box too).
__device__ __forceinline__ void test_children(const DeviceScene& sc, const RayPrep& rp, uint32_t id, float tHere, float slack, float best,
                                              uint32_t& L, uint32_t& R, bool& hitL, bool& hitR, float& tL, float& tR)
{
    const float4* rec = sc.wnodes + 4 * (size_t)id;
    const float4 bx = __ldg(rec + 0), by = __ldg(rec + 1), bz = __ldg(rec + 2), rf = __ldg(rec + 3);
    L = __float_as_uint(rf.x); R = __float_as_uint(rf.y);
    tL = tHere; tR = tHere;
    if (L & REF_LEAF) hitL = (L != REF_EMPTY);
    else hitL = rp.fast ? ray_box<true>(rp, bx.x, bx.y, by.x, by.y, bz.x, bz.y, &tL) : ray_box<false>(rp, bx.x, bx.y, by.x, by.y, bz.x, bz.y, &tL);
    if (R & REF_LEAF) hitR = (R != REF_EMPTY);
    else hitR = rp.fast ? ray_box<true>(rp, bx.z, bx.w, by.z, by.w, bz.z, bz.w, &tR) : ray_box<false>(rp, bx.z, bx.w, by.z, by.w, bz.z, bz.w, &tR);
    const uint32_t unprunable = __float_as_uint(rf.z);        // bit0/bit1: L/R subtree holds a triangle that failed the upload check
    if (unprunable & 1u) tL = -FLT_MAX;
    if (unprunable & 2u) tR = -FLT_MAX;
    if (hitL && pruned(tL, slack, best)) hitL = false;
    if (hitR && pruned(tR, slack, best)) hitR = false;
}

// STATS (developer switch pool_stats): per-phase iteration / lane counts are added to DeviceCounters (tools/pool_stats.py).
template <bool FUSED, int CAP_I, bool STATS = false>
__global__ void __launch_bounds__(POOL_WARPS * 32, POOL_CTAS_PER_SM)
rt_pool_kernel(DeviceScene sc, FrameParams fp, uint32_t* __restrict__ out, unsigned* __restrict__ pixelCounter, int4 tiles,
               HitRecord* __restrict__ hits, unsigned* __restrict__ hitCount, int prune, DeviceCounters* __restrict__ stats)
{
    unsigned st_it[5] = {0, 0, 0, 0, 0}, st_ln[5] = {0, 0, 0, 0, 0}, st_drop = 0, st_max = 0;   // inner, leaf, resolve, refill, guard
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Pool = WarpPool<CAP_I>;
    Pool& P = reinterpret_cast<Pool*>(smem_raw)[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned FULL = 0xffffffffu;
    const V3 eye = mkv3(fp.eye[0], fp.eye[1], fp.eye[2]);
    const V3 lightPos = mkv3(fp.light_pos[0][0], fp.light_pos[0][1], fp.light_pos[0][2]);
    const int tx0 = tiles.x, ty0 = tiles.y, ntx = tiles.z, nty = tiles.w;
    const unsigned total = (unsigned)(ntx * nty) * 32u;
    const float INF = __int_as_float(0x7f800000);

    // warp-uniform bookkeeping
    int icount = 0, lcount = 0;
    unsigned freeMask = FULL, doneMask = 0u;
    bool exhausted = (total == 0u);

    for (;;) {
        // ------------------------------------------------------------------ leaves, 32 at a time
        if (lcount >= LEAF_MIN || (lcount > 0 && icount == 0)) {
            const int n = min(32, lcount);
            lcount -= n;
            if (STATS) { st_it[1]++; st_ln[1] += n; }
            bool fin = false; uint32_t slot = 0;
            if ((int)lane < n) {
                const uint2 it = P.lpool[lcount + n - 1 - (int)lane];
                slot = (it.x >> ITEM_SLOT_SHIFT) & 31u;
                const float4 ro = P.ro[slot];
                const uint32_t st = P.state[slot];
                const unsigned long long k0 = P.key[slot];
                const float best = __uint_as_float((uint32_t)(k0 >> 32));
                if (!(st & 2u) && !pruned(__uint_as_float(it.y), ro.w, best)) {
                    const SlotRay s = load_slot_ray(P, slot, ro);
                    unsigned long long bestK = k0;
                    const bool occ = intersect_leaf(sc, s.rp, it.x & ITEM_INDEX_MASK, (st & 1u) != 0u, s.avoid, lightPos, best, bestK);
                    if (occ) P.state[slot] = 3u;
                    else if (bestK < k0) atomicMin(&P.key[slot], bestK);
                }
                fin = atomicSub(&P.pend[slot], 1) == 1;
            }
            doneMask |= __reduce_or_sync(FULL, fin ? (1u << slot) : 0u);
            __syncwarp();
            continue;
        }
        // ------------------------------------------------------------------ finished slots: black / shade + shadow ray / hit record
        const int ndone = __popc(doneMask);
        if (ndone >= SHADE_MIN || (ndone > 0 && icount == 0)) {
            if (STATS) { st_it[2]++; st_ln[2] += ndone; }
            bool freed = false, arm = false; uint32_t slot = 0;
            bool record = false; int tri = -1; V3 hitp = eye; float kAB = 0.f, kBC = 0.f, kCA = 0.f;
            if ((int)lane < ndone) {
                slot = __fns(doneMask, 0u, (int)lane + 1);
                const uint32_t st = P.state[slot];
                const uint32_t pix = P.pix[slot];
                const size_t o = (size_t)(pix >> 16) * fp.W + (pix & 0xffffu);
                freed = true;
                if (st & 1u) out[o] = (st & 2u) ? P.shd[slot] : P.lit[slot];
                else {
                    const unsigned long long k = P.key[slot];
                    if (k != KEY_EMPTY) {          // else: pierced nothing - the frame was cleared to black
                        const float4 rd = P.rd[slot];
                        reconstruct_hit(sc, eye, mkv3(rd.x, rd.y, rd.z), (uint32_t)k, tri, hitp, kAB, kBC, kCA);
                        if (FUSED) {
                            uint32_t pixLit, pixShadow; V3 sdir; float ldsq;
                            shade_one_light(sc, fp, eye, tri, hitp, kAB, kBC, kCA, pixLit, pixShadow, sdir, ldsq);
                            if (!(fp.flags & B200R_F_SHADOWS) || pixLit == pixShadow) out[o] = pixLit;   // the shadow ray cannot change this pixel
                            else {
                                const RayPrep rp = prep_ray(sc, hitp, sdir);
                                bool enter = true;
                                if (!(sc.root_ref & REF_LEAF))
                                    enter = rp.fast ? ray_box<true>(rp, sc.root_lo[0], sc.root_hi[0], sc.root_lo[1], sc.root_hi[1], sc.root_lo[2], sc.root_hi[2])
                                                    : ray_box<false>(rp, sc.root_lo[0], sc.root_hi[0], sc.root_lo[1], sc.root_hi[1], sc.root_lo[2], sc.root_hi[2]);
                                else if (sc.root_ref == REF_EMPTY) enter = false;
                                if (!enter) out[o] = pixLit;
                                else {
                                    P.ro[slot] = make_float4(hitp.x, hitp.y, hitp.z, INF);         // any-hit ray: no distance pruning
                                    P.rd[slot] = make_float4(sdir.x, sdir.y, sdir.z, rp.fast ? 1.f : 0.f);
                                    P.rr[slot] = make_float4(rp.r.x, rp.r.y, rp.r.z, __int_as_float(tri));
                                    P.key[slot] = (unsigned long long)__float_as_uint(ldsq) << 32;
                                    P.pend[slot] = 1; P.lit[slot] = pixLit; P.shd[slot] = pixShadow; P.state[slot] = 1u;
                                    freed = false; arm = true;
                                }
                            }
                        } else record = true;
                    }
                }
            }
            if (FUSED) {
                const unsigned am = __ballot_sync(FULL, arm);
                if (am) {
                    const uint2 item = make_uint2(make_item(sc.root_ref, slot), __float_as_uint(-FLT_MAX));
                    if (sc.root_ref & REF_LEAF) { if (arm) P.lpool[lcount + __popc(am & lt)] = item; lcount += __popc(am); }
                    else { if (arm) P.ipool[icount + __popc(am & lt)] = item; icount += __popc(am); }
                }
            } else {
                const unsigned hm = __ballot_sync(FULL, record);
                if (hm) {
                    unsigned base = 0;
                    if (lane == (unsigned)(__ffs(hm) - 1)) base = atomicAdd(hitCount, (unsigned)__popc(hm));
                    base = __shfl_sync(FULL, base, __ffs(hm) - 1);
                    if (record) {
                        float4* dst = reinterpret_cast<float4*>(hits + base + __popc(hm & lt));
                        dst[0] = make_float4(__int_as_float((int)P.pix[slot]), __int_as_float(tri), hitp.x, hitp.y);
                        dst[1] = make_float4(hitp.z, kAB, kBC, kCA);
                    }
                }
            }
            freeMask |= __reduce_or_sync(FULL, freed ? (1u << slot) : 0u);
            doneMask = 0u;
            __syncwarp();
            continue;
        }
        // ------------------------------------------------------------------ new pixels into free slots
        if (!exhausted && icount < LOW_WATER && (__popc(freeMask) >= REFILL_MIN || icount == 0)) {
            const int nfree = __popc(freeMask);
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(pixelCounter, (unsigned)nfree);
            base = __shfl_sync(FULL, base, 0);
            if (base + (unsigned)nfree >= total) exhausted = true;
            const unsigned g = base + lane;
            bool enter = false; RayPrep rp; int x = 0, r = 0;
            if ((int)lane < nfree && g < total) {
                const unsigned tile = g >> 5, l = g & 31u;
                const int qrow = (int)(tile / (unsigned)ntx), off = (qrow + 1) >> 1;
                const int trow = (qrow & 1) ? (nty >> 1) - off : (nty >> 1) + off;        // centre-out: the expensive tiles first
                x = (tx0 + (int)(tile % (unsigned)ntx)) * 8 + (int)(l & 7u);
                r = (ty0 + trow) * 4 + (int)(l >> 3);
                if (x < (int)fp.W && r < (int)fp.n_rows) {
                    const int y = (int)fp.row_first + r * (int)fp.row_step;
                    rp = prep_ray(sc, eye, primary_ray(fp, x, y));
                    if (sc.root_ref & REF_LEAF) enter = (sc.root_ref != REF_EMPTY);
                    else enter = rp.fast ? ray_box<true>(rp, sc.root_lo[0], sc.root_hi[0], sc.root_lo[1], sc.root_hi[1], sc.root_lo[2], sc.root_hi[2])
                                         : ray_box<false>(rp, sc.root_lo[0], sc.root_hi[0], sc.root_lo[1], sc.root_hi[1], sc.root_lo[2], sc.root_hi[2]);
                }
            }
            const unsigned em = __ballot_sync(FULL, enter);
            if (STATS) { st_it[3]++; st_ln[3] += __popc(em); }
            uint32_t slot = 0;
            if (enter) {
                slot = __fns(freeMask, 0u, __popc(em & lt) + 1);
                float slack = INF;
                if (prune) {
                    // 1/|d| per axis (IEEE divide; +inf for a zero component switches pruning off for this ray)
                    const float m = fmaxf(fmaxf(1.0f / fabsf(rp.d.x), 1.0f / fabsf(rp.d.y)), 1.0f / fabsf(rp.d.z));
                    slack = 1e-4f * m + 1e-4f;
                }
                P.ro[slot] = make_float4(rp.o.x, rp.o.y, rp.o.z, slack);
                P.rd[slot] = make_float4(rp.d.x, rp.d.y, rp.d.z, rp.fast ? 1.f : 0.f);
                P.rr[slot] = make_float4(rp.r.x, rp.r.y, rp.r.z, __int_as_float(-1));
                P.key[slot] = KEY_EMPTY;
                P.pend[slot] = 1; P.pix[slot] = ((uint32_t)r << 16) | (uint32_t)x; P.state[slot] = 0u;
                const uint2 item = make_uint2(make_item(sc.root_ref, slot), __float_as_uint(-FLT_MAX));
                if (sc.root_ref & REF_LEAF) P.lpool[lcount + __popc(em & lt)] = item;
                else P.ipool[icount + __popc(em & lt)] = item;
            }
            if (sc.root_ref & REF_LEAF) lcount += __popc(em); else icount += __popc(em);
            freeMask &= ~__reduce_or_sync(FULL, enter ? (1u << slot) : 0u);
            __syncwarp();
            continue;
        }
        if (icount == 0) break;            // nothing pending, nothing waiting, no pixels left

        // ------------------------------------------------------------------ inner nodes, 32 at a time
        const int n = min(32, icount);
        icount -= n;
        bool fin = false; uint32_t slot = 0;
        if (icount + n > CAP_I - 64) {
            // (overflow guard) the pool is nearly full: each lane walks its entry's whole subtree depth-first with a private stack
            if (STATS) { st_it[4]++; st_ln[4] += n; }
            if ((int)lane < n) {
                const uint2 it = P.ipool[icount + n - 1 - (int)lane];
                slot = (it.x >> ITEM_SLOT_SHIFT) & 31u;
                const float4 ro = P.ro[slot];
                const SlotRay s = load_slot_ray(P, slot, ro);
                const bool isShadow = (P.state[slot] & 1u) != 0u;
                uint32_t stk[B200R_BVH_STACK_SIZE]; float tst[B200R_BVH_STACK_SIZE];
                int sp = 0;
                uint32_t cur = it.x & (ITEM_LEAF | ITEM_INDEX_MASK); float tcur = __uint_as_float(it.y);
                for (;;) {
                    const unsigned long long k0 = P.key[slot];
                    const float best = __uint_as_float((uint32_t)(k0 >> 32));
                    bool pop = true;
                    if (!(P.state[slot] & 2u) && !pruned(tcur, s.slack, best)) {
                        if (cur & ITEM_LEAF) {
                            unsigned long long bestK = k0;
                            if (intersect_leaf(sc, s.rp, cur & ITEM_INDEX_MASK, isShadow, s.avoid, lightPos, best, bestK)) P.state[slot] = 3u;
                            else if (bestK < k0) atomicMin(&P.key[slot], bestK);
                        } else {
                            uint32_t L, R; bool hitL, hitR; float tL, tR;
                            test_children(sc, s.rp, cur, tcur, s.slack, best, L, R, hitL, hitR, tL, tR);
                            if (hitL && hitR) {
                                const bool rFirst = tR < tL;
                                stk[sp] = (rFirst ? L : R) & (ITEM_LEAF | ITEM_INDEX_MASK); tst[sp] = rFirst ? tL : tR; sp++;
                                cur = (rFirst ? R : L) & (ITEM_LEAF | ITEM_INDEX_MASK); tcur = rFirst ? tR : tL; pop = false;
                            } else if (hitL) { cur = L & (ITEM_LEAF | ITEM_INDEX_MASK); tcur = tL; pop = false; }
                            else if (hitR) { cur = R & (ITEM_LEAF | ITEM_INDEX_MASK); tcur = tR; pop = false; }
                        }
                    }
                    if (pop) {
                        if (sp == 0) break;
                        --sp; cur = stk[sp]; tcur = tst[sp];
                    }
                }
                fin = atomicSub(&P.pend[slot], 1) == 1;
            }
            doneMask |= __reduce_or_sync(FULL, fin ? (1u << slot) : 0u);
            __syncwarp();
            continue;
        }
        if (STATS) { st_it[0]++; st_ln[0] += n; st_max = max(st_max, (unsigned)(icount + n)); }
        uint32_t c0 = 0, c1 = 0; float t0 = 0.f, t1 = 0.f;          // children to push: c0 (far) first, c1 (near) on top of it
        bool p0 = false, p1 = false;
        if ((int)lane < n) {
            const uint2 it = P.ipool[icount + n - 1 - (int)lane];
            slot = (it.x >> ITEM_SLOT_SHIFT) & 31u;
            const float4 ro = P.ro[slot];
            const uint32_t st = P.state[slot];
            const float best = __uint_as_float((uint32_t)(P.key[slot] >> 32));
            const float tHere = __uint_as_float(it.y);
            int delta = -1;
            if (!(st & 2u) && !pruned(tHere, ro.w, best)) {
                const SlotRay s = load_slot_ray(P, slot, ro);
                uint32_t L, R; bool hitL, hitR; float tL, tR;
                test_children(sc, s.rp, it.x & ITEM_INDEX_MASK, tHere, s.slack, best, L, R, hitL, hitR, tL, tR);
                if (hitL && hitR) {
                    const bool rFirst = tR < tL;                     // nearer child on top
                    c0 = rFirst ? L : R; t0 = rFirst ? tL : tR; p0 = true;
                    c1 = rFirst ? R : L; t1 = rFirst ? tR : tL; p1 = true;
                    delta = 1;
                } else if (hitL) { c1 = L; t1 = tL; p1 = true; delta = 0; }
                else if (hitR) { c1 = R; t1 = tR; p1 = true; delta = 0; }
            } else if (STATS) st_drop++;
            if (delta != 0) fin = (atomicAdd(&P.pend[slot], delta) + delta) == 0;
        }
        {
            const bool i0 = p0 && !(c0 & REF_LEAF), i1 = p1 && !(c1 & REF_LEAF);
            const bool l0 = p0 && (c0 & REF_LEAF), l1 = p1 && (c1 & REF_LEAF);
            const unsigned bI0 = __ballot_sync(FULL, i0), bI1 = __ballot_sync(FULL, i1);
            const unsigned bL0 = __ballot_sync(FULL, l0), bL1 = __ballot_sync(FULL, l1);
            int io = icount + __popc(bI0 & lt) + __popc(bI1 & lt);
            int lo = lcount + __popc(bL0 & lt) + __popc(bL1 & lt);
            if (i0) P.ipool[io++] = make_uint2(make_item(c0, slot), __float_as_uint(t0));
            if (l0) P.lpool[lo++] = make_uint2(make_item(c0, slot), __float_as_uint(t0));
            if (i1) P.ipool[io] = make_uint2(make_item(c1, slot), __float_as_uint(t1));
            if (l1) P.lpool[lo] = make_uint2(make_item(c1, slot), __float_as_uint(t1));
            icount += __popc(bI0) + __popc(bI1);
            lcount += __popc(bL0) + __popc(bL1);
        }
        doneMask |= __reduce_or_sync(FULL, fin ? (1u << slot) : 0u);
        __syncwarp();
    }
    if (STATS) {
        st_drop = __reduce_add_sync(FULL, st_drop);
        if (lane == 0) {
            for (int i = 0; i < 5; i++) { atomicAdd(&stats->v[2 * i], (unsigned long long)st_it[i]); atomicAdd(&stats->v[2 * i + 1], (unsigned long long)st_ln[i]); }
            atomicAdd(&stats->v[10], (unsigned long long)st_drop);
        }
        (void)st_max;
    }
}

}  // namespace

// Tiles (8 x 4 pixels, in packed-row space) that can contain a pixel of the screen rectangle `b` = (x0, y0, x1, y1), inclusive,
// full-frame coordinates: (first tile column, first tile row, columns, rows).
static int4 tile_rect(const FrameParams& fp, int4 b)
{
    const int rs = (int)fp.row_step, rf = (int)fp.row_first;
    if (b.x > b.z || b.y > b.w) return make_int4(0, 0, 0, 0);
    int r0 = b.y <= rf ? 0 : (b.y - rf + rs - 1) / rs;                 // first packed row with y >= y0
    int r1 = b.w < rf ? -1 : (b.w - rf) / rs;                          // last packed row with y <= y1
    if (r1 > (int)fp.n_rows - 1) r1 = (int)fp.n_rows - 1;
    if (r0 > r1) return make_int4(0, 0, 0, 0);
    const int tx0 = b.x >> 3, tx1 = b.z >> 3, ty0 = r0 >> 2, ty1 = r1 >> 2;
    return make_int4(tx0, ty0, tx1 - tx0 + 1, ty1 - ty0 + 1);
}

bool pool_supported(const DeviceScene& sc)
{
    return sc.n_nodes <= ITEM_INDEX_MASK && sc.n_list <= ITEM_INDEX_MASK;
}

cudaError_t rt_pool_configure()
{
    cudaError_t e = cudaSuccess;
    const int big = (int)(sizeof(WarpPool<512>) * POOL_WARPS), small = (int)(sizeof(WarpPool<128>) * POOL_WARPS);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<true, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<false, 512, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<true, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<false, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
    return e;
}

cudaError_t launch_rt_pool(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, bool fused, bool prune, bool smallCap,
                           bool noRootRect, unsigned* pixelCounter, void* hits, unsigned* hitCount, int numSMs, cudaStream_t stream,
                           int& launches, DeviceCounters* stats)
{
    cudaError_t e = cudaMemsetAsync(d_out, 0, (size_t)fp.W * fp.n_rows * 4, stream);        // black; the kernel writes lit pixels only
    if (e != cudaSuccess) return e;
    const int4 bounds = noRootRect ? make_int4(0, 0, (int)fp.W - 1, (int)fp.H - 1) : root_screen_bounds(sc, fp);
    const int4 tiles = tile_rect(fp, bounds);
    if (tiles.z <= 0 || tiles.w <= 0) return cudaSuccess;
    using K = void (*)(DeviceScene, FrameParams, uint32_t*, unsigned*, int4, HitRecord*, unsigned*, int, DeviceCounters*);
    K k; size_t smem;
    if (smallCap) { k = fused ? rt_pool_kernel<true, 128> : rt_pool_kernel<false, 128>; smem = sizeof(WarpPool<128>) * POOL_WARPS; }
    else { k = fused ? rt_pool_kernel<true, 512> : rt_pool_kernel<false, 512>; smem = sizeof(WarpPool<512>) * POOL_WARPS; }
    if (stats && !smallCap) k = fused ? rt_pool_kernel<true, 512, true> : rt_pool_kernel<false, 512, true>;
    int grid = numSMs * POOL_CTAS_PER_SM;
    const int needed = (tiles.z * tiles.w + POOL_WARPS - 1) / POOL_WARPS;       // one tile per warp is the least a warp can take
    if (grid > needed) grid = needed;
    k<<<grid, POOL_WARPS * 32, smem, stream>>>(sc, fp, d_out, pixelCounter, tiles, reinterpret_cast<HitRecord*>(hits), hitCount, prune ? 1 : 0, stats);
    launches += 1;
    return cudaGetLastError();
}

}  // namespace b200r
