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
//   * Persistent CTAs of 4 independent warps (24 or 32 warps per SM x 148 SMs) - no CTA-wide barrier anywhere. Warps pull 8x4-pixel tiles of
//     the screen rectangle that can contain the model (centre-out), build the primary rays themselves and test the root box
//     against kernel arguments; the frame is cleared beforehand, so the 81 % of C2's pixels that miss are never touched.
//   * The pool cannot overflow: when it is nearly full the top 32 entries are walked depth-first by their lanes with a private
//     stack (same tests, same merges) instead of being expanded.
#include <cstring>

#include "rt_common.cuh"
#include "rt_kernels.cuh"

namespace b200r {
using namespace rt;

namespace {

constexpr int POOL_WARPS = 8;            // most warps per CTA (independent of each other); the default is 4 - pool_cta_warps = 8 / 2 / 1: as many warps per SM
constexpr int POOL_CTAS_PER_SM = 3;
constexpr int CAP_L = 128;               // leaf pool: < LEAF_MIN waiting + at most 64 pushed per inner iteration (+ 32 roots)
// scheduling thresholds (defaults; developer switches pool_* override them for tuning)
constexpr int LEAF_MIN = 24;             // run a leaf iteration as soon as this many leaves wait
constexpr int SORT_MIN = 8;              // look at finished slots once this many wait (or the warp is running dry)
constexpr int SHADE_MIN = 16;            // shade resolved hits once this many wait (or the warp is running dry)
constexpr int REFILL_MIN = 8;            // take new pixels once this many slots are free ...
constexpr int LOW_WATER = 32;            // ... and fewer than this many hot entries are pending
constexpr int DRY = 16;                  // "running dry": fewer inner entries than this are pending
constexpr uint32_t ITEM_LEAF = 0x80000000u;
constexpr uint32_t ITEM_INDEX_MASK = 0x03FFFFFFu;     // 26 bits: inner record id or list position
constexpr uint32_t ITEM_REF_MASK = ITEM_LEAF | ITEM_INDEX_MASK;
constexpr int ITEM_SLOT_SHIFT = 26;
constexpr unsigned long long KEY_EMPTY = ((unsigned long long)0x7F7FFFFFu << 32) | 0xFFFFFFFFull;   // (FLT_MAX, no list position)

// Per-warp state in shared memory. The inner-node pool is ONE array with two stacks: HOT entries (the nearer child of the
// node just processed - the depth-first continuation of its ray) grow up from index 0, COLD entries (the farther child, which
// the reference's loop would visit after the whole near subtree) grow down from the end. Lanes pop hot entries first and fill
// up with cold ones, newest first: with many rays in the warp every ray advances depth-first, nearest box first, and what it
// finds prunes its cold entries before anyone fetches them; with few rays left the spare lanes walk the cold entries of the
// same rays, which is what keeps a long ray from becoming the frame's critical path.
template <int CAP_I>
struct __align__(16) WarpPool {
    uint2 ipool[CAP_I];                  // {leaf flag | slot | index, entry distance of the node's box}
    uint2 lpool[CAP_L];                  // pending leaves
    float4 ro[32];                       // slot: ray origin, pruning slack (+inf: never prune; -inf: ray finished, drop its entries)
    float4 rd[32];                       // slot: ray direction
    float4 rr[32];                       // slot: refined reciprocals of the direction, w = bits(triangle to skip): >= 0 marks a SHADOW ray
    unsigned long long key[32];          // slot: (bits(best hitZ) << 32) | list position; shadow slots: (bits(light distance^2) << 32)
    int pend[32];                        // slot: pool entries not yet processed
    uint32_t pix[32];                    // slot: (packed row << 16) | x; queue modes: where the ray's result goes
    uint32_t lit[32], shd[32];           // shadow slots: the two possible pixel words
    float4 rl[32];                       // any-hit slots: the point the ray must reach (light / end of an AO ray)
};

__device__ __forceinline__ uint32_t make_item(uint32_t ref, uint32_t slot) { return (ref & ITEM_REF_MASK) | (slot << ITEM_SLOT_SHIFT); }

__device__ __forceinline__ bool pruned(float tnear, float slack, float best)
{
    const float e = tnear - slack;
    return e > 0.f && (e * e) * 0.99999f > best;
}

// RayIntersectsBox (reference src/Raytracer.cc:99-151) for rays inside the shared-reciprocal domain (rt_common.cuh "Division"):
// every quotient is finite there, so the compare-and-swap ladder of the reference collapses to min/max - the same values up to
// the sign of a zero, which neither `Tnear > Tfar` nor `Tfar < 0` can see - and the per-axis early returns to the final test.
__device__ __forceinline__ bool box_fast(const float4& o, const float4& d, const float4& r, float lox, float hix, float loy, float hiy,
                                         float loz, float hiz, float& tnear)
{
    const float x1 = div_shared_rcp(lox - o.x, d.x, r.x), x2 = div_shared_rcp(hix - o.x, d.x, r.x);
    const float y1 = div_shared_rcp(loy - o.y, d.y, r.y), y2 = div_shared_rcp(hiy - o.y, d.y, r.y);
    const float z1 = div_shared_rcp(loz - o.z, d.z, r.z), z2 = div_shared_rcp(hiz - o.z, d.z, r.z);
    const float tn = fmaxf(fmaxf(fminf(x1, x2), fminf(y1, y2)), fminf(z1, z2));
    const float tf = fminf(fminf(fmaxf(x1, x2), fmaxf(y1, y2)), fmaxf(z1, z2));
    tnear = tn;
    return !(tn > tf) && !(tf < 0.f);
}

// The triangles of one leaf against one ray, in list order (reference src/Raytracer.cc:235-298). Closest-hit rays fold
// improving hits into `bestK`; shadow rays return true at the first triangle that is nearer to the light than the origin is.
__device__ __forceinline__ bool intersect_leaf(const DeviceScene& sc, const V3& o, const V3& d, uint32_t li, int avoid, bool isShadow,
                                               const V3& lightPos, float lightDistSq, unsigned long long& bestK)
{
    const float4* rec = sc.leaftris + 5 * (size_t)li;
    for (;; rec += 5, li++) {
        const float4 q4 = __ldg(rec + 4), q0 = __ldg(rec + 0), q1 = __ldg(rec + 1), q2 = __ldg(rec + 2), q3 = __ldg(rec + 3);
        const uint32_t tw = __float_as_uint(q4.w);
        const bool last = (tw & 0x40000000u) != 0;
        const V3 n = mkv3(q0.x, q0.y, q0.z);
        bool alive = (int)(tw & 0x3fffffffu) != avoid;                      // avoidSelf (avoid = -1 for primary rays)
        if (alive && !(tw & 0x80000000u)) {                                 // doCulling && !twoSided
            const V3 fromTriToOrigin = o - mkv3(q4.x, q4.y, q4.z);
            if (dot3(fromTriToOrigin, n) < 0.f) alive = false;
        }
        if (alive) {
            const float k = dot3(n, d);
            if (k != 0.f) {
                const float s = (q0.w - dot3(n, o)) / k;
                if (s > 0.f && s > 1e-5f) {                                 // behind the origin / NUDGE_FACTOR
                    const V3 hit = d * s + o;
                    const float kt1 = dot3(mkv3(q1.x, q1.y, q1.z), hit) - q1.w;
                    if (!(kt1 < 0.f)) {
                        const float kt2 = dot3(mkv3(q2.x, q2.y, q2.z), hit) - q2.w;
                        if (!(kt2 < 0.f)) {
                            const float kt3 = dot3(mkv3(q3.x, q3.y, q3.z), hit) - q3.w;
                            if (!(kt3 < 0.f)) {
                                if (isShadow) {
                                    if (distancesq3(lightPos, hit) < lightDistSq) return true;
                                } else {
                                    const float hitZ = distancesq3(o, hit);
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

// Cold path: one lane walks a whole subtree depth-first with a private stack - the reference's own loop with near-first order
// and pruning. Used for rays outside the shared-reciprocal domain (a zero direction component: they take the reference's
// `dir == 0` rule, ray_box<false>) and when a pool is about to overflow. Results go where the pooled lanes put theirs.
__device__ __noinline__ void walk_subtree(const DeviceScene& sc, const RayPrep rp, const int avoid, const bool anyhit, const V3 lightPos, uint32_t cur, float tcur,
                                          unsigned long long* key, const volatile float* slackWord)
{
    uint32_t stk[B200R_BVH_STACK_SIZE]; float tst[B200R_BVH_STACK_SIZE];
    int sp = 0;
    for (;;) {
        const unsigned long long k0 = *reinterpret_cast<volatile unsigned long long*>(key);
        const float best = __uint_as_float((uint32_t)(k0 >> 32));
        const float slack = *slackWord;
        bool pop = true;
        if (cur != REF_EMPTY && !pruned(tcur, slack, best)) {
            if (cur & REF_LEAF) {
                unsigned long long bestK = k0;
                if (intersect_leaf(sc, rp.o, rp.d, cur & ITEM_INDEX_MASK, avoid, anyhit, lightPos, best, bestK)) {
                    *const_cast<float*>(slackWord) = -__int_as_float(0x7f800000);     // occluded: every other entry of the ray is dropped
                    return;
                }
                if (bestK < k0) atomicMin(key, bestK);
            } else {
                const float4* rec = sc.wnodes + 4 * (size_t)cur;
                const float4 bx = __ldg(rec + 0), by = __ldg(rec + 1), bz = __ldg(rec + 2), rf = __ldg(rec + 3);
                const uint32_t L = __float_as_uint(rf.x), R = __float_as_uint(rf.y);
                bool hitL, hitR; float tL = tcur, tR = tcur;
                if (L & REF_LEAF) hitL = (L != REF_EMPTY);
                else hitL = rp.fast ? ray_box<true>(rp, bx.x, bx.y, by.x, by.y, bz.x, bz.y, &tL) : ray_box<false>(rp, bx.x, bx.y, by.x, by.y, bz.x, bz.y, &tL);
                if (R & REF_LEAF) hitR = (R != REF_EMPTY);
                else hitR = rp.fast ? ray_box<true>(rp, bx.z, bx.w, by.z, by.w, bz.z, bz.w, &tR) : ray_box<false>(rp, bx.z, bx.w, by.z, by.w, bz.z, bz.w, &tR);
                const uint32_t unprunable = __float_as_uint(rf.z);
                if (unprunable & 1u) tL = -FLT_MAX;
                if (unprunable & 2u) tR = -FLT_MAX;
                if (hitL && pruned(tL, slack, best)) hitL = false;
                if (hitR && pruned(tR, slack, best)) hitR = false;
                if (hitL && hitR) {
                    const bool rFirst = tR < tL;
                    stk[sp] = rFirst ? L : R; tst[sp] = rFirst ? tL : tR; sp++;
                    cur = rFirst ? R : L; tcur = rFirst ? tR : tL; pop = false;
                } else if (hitL) { cur = L; tcur = tL; pop = false; }
                else if (hitR) { cur = R; tcur = tR; pop = false; }
            }
        }
        if (pop) {
            if (sp == 0) return;
            --sp; cur = stk[sp]; tcur = tst[sp];
        }
    }
}

// How pixels are dealt to warps and how the inner pool is popped (PoolParams.policy, developer switch pool_policy):
//   0  hot entries first, spare lanes take cold ones (depth-first per ray while the warp has enough rays)
//   1  one stack: the farther child is pushed under the nearer one and both are popped as they come (breadth grows fast)
//   2  like 0, but up to 8 lanes always go to cold entries
//   3  one stack, each lane's farther child directly under its nearer one: a node's two children are popped together,
//      a ray's pending subtrees spread over the lanes at once (shortest chains of dependent iterations, weakest pruning)
struct PoolParams {
    int4 tiles;                     // first tile column / row, tile columns / rows of the screen rectangle that can contain the model
    int prune, policy;
    int leafMin, sortMin, shadeMin, refillMin, lowWater, dry;
    unsigned scatterMul;            // 0: tiles are dealt in centre-out order, 32 neighbouring pixels per grab; else: 4-pixel groups
    unsigned nGroups, groupsPerRow; //    are dealt in the order (q * scatterMul) mod nGroups, so every warp holds a cross-section
    unsigned long long scatterInv;  //    of the frame instead of one tile (floor(2^64 / nGroups), for the modulo)
};

// Where the rays of a launch come from and where their results go (MODE):
//   POOL_FUSED    the pixels of the frame; a resolved hit is shaded here and its slot re-armed as the hit's shadow ray (C2-type frames)
//   POOL_PRIMARY  the pixels of the frame; resolved hits are appended to `hits` (generic configurations, level 0)
//   POOL_ANYHIT   a queue of 48-byte ray records {origin, -}{direction, bits(triangle to skip)}{point to reach, bits(result index)}:
//                 shadow and ambient-occlusion rays; `occ[result index]` becomes 1 if anything lies between origin and that point
//   POOL_CLOSEST  the same records, closest hit wanted (reflection rays); hits are appended to `hits` tagged with the result index
// Queue modes take the rays [0, min(queueCap, *queueCount - queueFirst) * queueStride) of `rays`.
enum { POOL_FUSED = 0, POOL_PRIMARY = 1, POOL_ANYHIT = 2, POOL_CLOSEST = 3 };
struct PoolQueue {
    const float4* rays; const unsigned* count; unsigned first, cap, stride;
    unsigned char* occ;
};

// STATS (developer switch pool_stats): per-phase iteration / lane counts are added to DeviceCounters (tools/pool_stats.py).
template <int MODE, int CAP_I, bool STATS = false, int CTAS = POOL_CTAS_PER_SM, int WARPS = POOL_WARPS>
__global__ void __launch_bounds__(WARPS * 32, CTAS * (POOL_WARPS / WARPS))
rt_pool_kernel(DeviceScene sc, FrameParams fp, uint32_t* __restrict__ out, unsigned* __restrict__ pixelCounter, PoolParams pp,
               HitRecord* __restrict__ hits, unsigned* __restrict__ hitCount, DeviceCounters* __restrict__ stats, PoolQueue q)
{
    constexpr bool FUSED = MODE == POOL_FUSED;
    constexpr bool QUEUE = MODE == POOL_ANYHIT || MODE == POOL_CLOSEST;
    const int4 tiles = pp.tiles;
    const int prune = pp.prune, policy = pp.policy;
    const int LEAF_MIN = pp.leafMin, SORT_MIN = pp.sortMin, SHADE_MIN = pp.shadeMin, REFILL_MIN = pp.refillMin, LOW_WATER = pp.lowWater, DRY = pp.dry;
    unsigned st_it[5] = {0, 0, 0, 0, 0}, st_ln[5] = {0, 0, 0, 0, 0}, st_drop = 0, st_cold = 0;   // inner, leaf, resolve, refill, guard
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using Pool = WarpPool<CAP_I>;
    Pool& P = reinterpret_cast<Pool*>(smem_raw)[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned FULL = 0xffffffffu;
    const V3 eye = mkv3(fp.eye[0], fp.eye[1], fp.eye[2]);
    const V3 lightPos = mkv3(fp.light_pos[0][0], fp.light_pos[0][1], fp.light_pos[0][2]);
    const int tx0 = tiles.x, ty0 = tiles.y, ntx = tiles.z, nty = tiles.w;
    unsigned total = (unsigned)(ntx * nty) * 32u;
    if (QUEUE) {
        const unsigned n = *q.count;
        total = n > q.first ? min(q.cap, n - q.first) * q.stride : 0u;
    }
    const float INF = __int_as_float(0x7f800000);

    // warp-uniform bookkeeping
    int hcount = 0, ccount = 0, lcount = 0;         // hot / cold inner entries, leaves
    unsigned freeMask = FULL, doneMask = 0u, shadeMask = 0u;       // slots: free / finished, not looked at yet / resolved hits waiting to be shaded
    bool exhausted = (total == 0u);

    for (;;) {
        // ------------------------------------------------------------------ leaves, 32 at a time
        if (lcount >= LEAF_MIN || (lcount > 0 && hcount + ccount == 0)) {
            const int n = min(32, lcount);
            lcount -= n;
            if (STATS) { st_it[1]++; st_ln[1] += n; }
            bool fin = false, occluded = false; uint32_t slot = 0;
            if ((int)lane < n) {
                const uint2 it = P.lpool[lcount + n - 1 - (int)lane];
                slot = (it.x >> ITEM_SLOT_SHIFT) & 31u;
                const float4 ro = P.ro[slot];
                const unsigned long long k0 = P.key[slot];
                const float best = __uint_as_float((uint32_t)(k0 >> 32));
                if (!pruned(__uint_as_float(it.y), ro.w, best)) {
                    const float4 rd = P.rd[slot];
                    unsigned long long bestK = k0;
                    const int avoid = __float_as_int(P.rr[slot].w);
                    const bool anyhit = MODE == POOL_ANYHIT || (FUSED && avoid >= 0);
                    const float4 rl = P.rl[slot];
                    const bool occ = intersect_leaf(sc, mkv3(ro.x, ro.y, ro.z), mkv3(rd.x, rd.y, rd.z), it.x & ITEM_INDEX_MASK,
                                                    avoid, anyhit, mkv3(rl.x, rl.y, rl.z), best, bestK);
                    occluded = occ;
                    if (!occ && bestK < k0) atomicMin(&P.key[slot], bestK);
                }
                fin = atomicSub(&P.pend[slot], 1) == 1;
            }
            __syncwarp();               // every lane has read its slot's state: now the occlusion marks may be written
            if (occluded) P.ro[slot].w = -INF;
            doneMask |= __reduce_or_sync(FULL, fin ? (1u << slot) : 0u);
            __syncwarp();
            continue;
        }
        // ------------------------------------------------------------------ finished slots: shadow rays write their pixel, primary rays
        // that pierced nothing are done (the frame was cleared to black), resolved hits queue up for shading
        // (every branch that pushes checks the room left in the inner pool: an inner iteration needs 64 free entries - each of
        // its <= 32 lanes may push two - or it runs as the overflow guard; shading / refilling push at most 32)
        const int room = CAP_I - (hcount + ccount);
        const bool dry = hcount + ccount < DRY;
        const int ndone = __popc(doneMask);
        if (ndone >= SORT_MIN || (ndone > 0 && dry)) {
            bool freed = false, hit = false; uint32_t slot = 0;
            if ((int)lane < ndone) {
                slot = __fns(doneMask, 0u, (int)lane + 1);
                if (MODE == POOL_ANYHIT) {
                    if (P.ro[slot].w == -INF) q.occ[P.pix[slot]] = 1;
                    freed = true;
                } else if (FUSED && __float_as_int(P.rr[slot].w) >= 0) {
                    const uint32_t pix = P.pix[slot];
                    out[(size_t)(pix >> 16) * fp.W + (pix & 0xffffu)] = (P.ro[slot].w == -INF) ? P.shd[slot] : P.lit[slot];
                    freed = true;
                } else if (P.key[slot] == KEY_EMPTY) freed = true;
                else hit = true;
            }
            freeMask |= __reduce_or_sync(FULL, freed ? (1u << slot) : 0u);
            shadeMask |= __reduce_or_sync(FULL, hit ? (1u << slot) : 0u);
            doneMask = 0u;
            __syncwarp();               // the slots read here are re-armed / refilled by other lanes in the next pass
            continue;
        }
        // ------------------------------------------------------------------ resolved hits: shade + shadow ray / hit record
        const int nshade = __popc(shadeMask);
        if ((nshade >= SHADE_MIN || (nshade > 0 && dry)) && room >= 32) {
            if (STATS) { st_it[2]++; st_ln[2] += nshade; }
            bool freed = false, arm = false; uint32_t slot = 0;
            bool record = false; int tri = -1; V3 hitp = eye; float kAB = 0.f, kBC = 0.f, kCA = 0.f;
            if ((int)lane < nshade) {
                slot = __fns(shadeMask, 0u, (int)lane + 1);
                const uint32_t pix = P.pix[slot];
                const size_t o = (size_t)(pix >> 16) * fp.W + (pix & 0xffffu);
                freed = true;
                const unsigned long long k = P.key[slot];
                const float4 rd = P.rd[slot], ro = P.ro[slot];
                reconstruct_hit(sc, mkv3(ro.x, ro.y, ro.z), mkv3(rd.x, rd.y, rd.z), (uint32_t)k, tri, hitp, kAB, kBC, kCA);
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
                            P.rd[slot] = make_float4(sdir.x, sdir.y, sdir.z, rp.fast ? 0.f : 1.f);
                            P.rr[slot] = make_float4(rp.r.x, rp.r.y, rp.r.z, __int_as_float(tri));
                            P.key[slot] = (unsigned long long)__float_as_uint(ldsq) << 32;
                            P.rl[slot] = make_float4(lightPos.x, lightPos.y, lightPos.z, 0.f);
                            P.pend[slot] = 1; P.lit[slot] = pixLit; P.shd[slot] = pixShadow; freed = false; arm = true;
                        }
                    }
                } else record = true;
            }
            if (FUSED) {
                const unsigned am = __ballot_sync(FULL, arm);
                if (am) {
                    const uint2 item = make_uint2(make_item(sc.root_ref, slot), __float_as_uint(-FLT_MAX));
                    if (sc.root_ref & REF_LEAF) { if (arm) P.lpool[lcount + __popc(am & lt)] = item; lcount += __popc(am); }
                    else { if (arm) P.ipool[hcount + __popc(am & lt)] = item; hcount += __popc(am); }
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
            shadeMask = 0u;
            __syncwarp();
            continue;
        }
        // ------------------------------------------------------------------ new pixels into free slots
        if (!exhausted && hcount < LOW_WATER && room >= 32 && (__popc(freeMask) >= REFILL_MIN || hcount + ccount == 0)) {
            const int nfree = __popc(freeMask);
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(pixelCounter, (unsigned)nfree);
            base = __shfl_sync(FULL, base, 0);
            if (base + (unsigned)nfree >= total) exhausted = true;
            const unsigned g = base + lane;
            bool enter = false; RayPrep rp; int x = 0, r = 0;
            int qAvoid = -1; uint32_t qResult = 0; V3 qTarget = eye;
            if (QUEUE) {
                if ((int)lane < nfree && g < total) {
                    const float4* rec = q.rays + 3 * (size_t)g;
                    const float4 a = __ldg(rec), b = __ldg(rec + 1), c = __ldg(rec + 2);
                    qAvoid = __float_as_int(b.w); qResult = __float_as_uint(c.w); qTarget = mkv3(c.x, c.y, c.z);
                    rp = prep_ray(sc, mkv3(a.x, a.y, a.z), mkv3(b.x, b.y, b.z));
                    if (sc.root_ref & REF_LEAF) enter = (sc.root_ref != REF_EMPTY);
                    else enter = rp.fast ? ray_box<true>(rp, sc.root_lo[0], sc.root_hi[0], sc.root_lo[1], sc.root_hi[1], sc.root_lo[2], sc.root_hi[2])
                                         : ray_box<false>(rp, sc.root_lo[0], sc.root_hi[0], sc.root_lo[1], sc.root_hi[1], sc.root_lo[2], sc.root_hi[2]);
                }
            } else
            if ((int)lane < nfree && g < total) {
                if (pp.scatterMul) {
                    const unsigned long long prod = (unsigned long long)(g >> 2) * pp.scatterMul;
                    unsigned long long rem = prod - __umul64hi(prod, pp.scatterInv) * pp.nGroups;
                    if (rem >= pp.nGroups) rem -= pp.nGroups;
                    const unsigned q = (unsigned)rem;
                    x = tx0 * 8 + (int)(q % pp.groupsPerRow) * 4 + (int)(g & 3u);
                    r = ty0 * 4 + (int)(q / pp.groupsPerRow);
                } else {
                    const unsigned tile = g >> 5, l = g & 31u;
                    const int qrow = (int)(tile / (unsigned)ntx), off = (qrow + 1) >> 1;
                    const int trow = (qrow & 1) ? (nty >> 1) - off : (nty >> 1) + off;        // centre-out: the expensive tiles first
                    x = (tx0 + (int)(tile % (unsigned)ntx)) * 8 + (int)(l & 7u);
                    r = (ty0 + trow) * 4 + (int)(l >> 3);
                }
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
            uint32_t slot = 0; const bool walked = false;
            if (enter) {
                slot = __fns(freeMask, 0u, __popc(em & lt) + 1);
                float slack = INF;
                if (prune && MODE != POOL_ANYHIT) {
                    // 1/|d| per axis (IEEE divide; +inf for a zero component switches pruning off for this ray)
                    const float m = fmaxf(fmaxf(1.0f / fabsf(rp.d.x), 1.0f / fabsf(rp.d.y)), 1.0f / fabsf(rp.d.z));
                    slack = 1e-4f * m + 1e-4f;
                }
                P.ro[slot] = make_float4(rp.o.x, rp.o.y, rp.o.z, slack);
                P.rd[slot] = make_float4(rp.d.x, rp.d.y, rp.d.z, rp.fast ? 0.f : 1.f);
                P.rr[slot] = make_float4(rp.r.x, rp.r.y, rp.r.z, __int_as_float(QUEUE ? qAvoid : -1));
                // any-hit: "best" is the squared distance origin - target, as BVH_IntersectTriangles<true> starts (src/Raytracer.cc:212)
                P.key[slot] = MODE == POOL_ANYHIT ? (unsigned long long)__float_as_uint(distancesq3(rp.o, qTarget)) << 32 : KEY_EMPTY;
                if (MODE == POOL_ANYHIT) P.rl[slot] = make_float4(qTarget.x, qTarget.y, qTarget.z, 0.f);
                P.pix[slot] = QUEUE ? qResult : (((uint32_t)r << 16) | (uint32_t)x);
                P.pend[slot] = 1;
            }
            const unsigned pm = __ballot_sync(FULL, enter && !walked);
            if (enter && !walked) {
                const uint2 item = make_uint2(make_item(sc.root_ref, slot), __float_as_uint(-FLT_MAX));
                if (sc.root_ref & REF_LEAF) P.lpool[lcount + __popc(pm & lt)] = item;
                else P.ipool[hcount + __popc(pm & lt)] = item;
            }
            if (sc.root_ref & REF_LEAF) lcount += __popc(pm); else hcount += __popc(pm);
            freeMask &= ~__reduce_or_sync(FULL, enter ? (1u << slot) : 0u);
            doneMask |= __reduce_or_sync(FULL, walked ? (1u << slot) : 0u);
            __syncwarp();
            continue;
        }
        if (hcount + ccount == 0) break;            // nothing pending, nothing waiting, no pixels left

        // ------------------------------------------------------------------ inner nodes, 32 at a time: hot entries first, then cold
        int nh = min(32, hcount), nc = min(32 - nh, ccount);
        if (policy == 2 && nc < 8 && ccount > nc) { nc = min(8, ccount); nh = min(nh, 32 - nc); }
        const bool guard = room < 64;
        uint2 it = make_uint2(0u, 0u);
        const bool have = (int)lane < nh + nc;
        if ((int)lane < nh) it = P.ipool[hcount - 1 - (int)lane];
        else if (have) it = P.ipool[CAP_I - ccount + ((int)lane - nh)];
        hcount -= nh; ccount -= nc;
        bool fin = false;
        const uint32_t slot = (it.x >> ITEM_SLOT_SHIFT) & 31u;
        if (guard) {
            // (overflow guard) the pool is nearly full: each lane walks its entry's whole subtree itself instead of expanding it
            if (STATS) { st_it[4]++; st_ln[4] += nh + nc; }
            if (have) {
                const float4 ro = P.ro[slot], rd = P.rd[slot], rr = P.rr[slot];
                RayPrep rp; rp.o = mkv3(ro.x, ro.y, ro.z); rp.d = mkv3(rd.x, rd.y, rd.z); rp.r = mkv3(rr.x, rr.y, rr.z); rp.fast = rd.w == 0.f;
                { const int avoid = __float_as_int(rr.w); const float4 rl = P.rl[slot];
                  walk_subtree(sc, rp, avoid, MODE == POOL_ANYHIT || (FUSED && avoid >= 0), mkv3(rl.x, rl.y, rl.z), it.x & ITEM_REF_MASK, __uint_as_float(it.y), &P.key[slot], &P.ro[slot].w); }
                fin = atomicSub(&P.pend[slot], 1) == 1;
            }
            doneMask |= __reduce_or_sync(FULL, fin ? (1u << slot) : 0u);
            __syncwarp();
            continue;
        }
        if (STATS) { st_it[0]++; st_ln[0] += nh + nc; st_cold += nc; }
        uint32_t cN = 0, cF = 0; float tN = 0.f, tF = 0.f;          // surviving children: the nearer one, the farther one
        bool pN = false, pF = false;
        if (have) {
            const float4 ro = P.ro[slot];
            const float best = __uint_as_float(reinterpret_cast<const uint32_t*>(&P.key[slot])[1]);
            const float tHere = __uint_as_float(it.y);
            int delta = -1;
            if (!pruned(tHere, ro.w, best)) {
                const float4 rd = P.rd[slot], rr = P.rr[slot];
                const float4* rec = sc.wnodes + 4 * (size_t)(it.x & ITEM_INDEX_MASK);
                const float4 bx = __ldg(rec + 0), by = __ldg(rec + 1), bz = __ldg(rec + 2), rf = __ldg(rec + 3);
                const uint32_t L = __float_as_uint(rf.x), R = __float_as_uint(rf.y);
                // both boxes, unconditionally (straight-line code). A LEAF child has no box test in the reference
                // (src/Raytracer.cc:224-229): it survives unless empty; its box only supplies a tighter entry distance.
                float tL, tR; bool hitL, hitR;
                if (rd.w == 0.f) {
                    hitL = box_fast(ro, rd, rr, bx.x, bx.y, by.x, by.y, bz.x, bz.y, tL);
                    hitR = box_fast(ro, rd, rr, bx.z, bx.w, by.z, by.w, bz.z, bz.w, tR);
                } else {            // a ray outside the shared-reciprocal domain (e.g. a zero direction component): the reference's own rules
                    RayPrep rp; rp.o = mkv3(ro.x, ro.y, ro.z); rp.d = mkv3(rd.x, rd.y, rd.z); rp.r = rp.d; rp.fast = false;
                    hitL = ray_box<false>(rp, bx.x, bx.y, by.x, by.y, bz.x, bz.y, &tL);
                    hitR = ray_box<false>(rp, bx.z, bx.w, by.z, by.w, bz.z, bz.w, &tR);
                }
                if (L & REF_LEAF) { if (!hitL) tL = tHere; hitL = (L != REF_EMPTY); }
                if (R & REF_LEAF) { if (!hitR) tR = tHere; hitR = (R != REF_EMPTY); }
                const uint32_t unprunable = __float_as_uint(rf.z);        // bit0/bit1: L/R subtree holds a triangle that failed the upload check
                if (unprunable & 1u) tL = -FLT_MAX;
                if (unprunable & 2u) tR = -FLT_MAX;
                if (hitL && pruned(tL, ro.w, best)) hitL = false;
                if (hitR && pruned(tR, ro.w, best)) hitR = false;
                const bool rFirst = hitR && (!hitL || tR < tL);
                cN = rFirst ? R : L; tN = rFirst ? tR : tL; pN = hitL || hitR;
                cF = rFirst ? L : R; tF = rFirst ? tL : tR; pF = hitL && hitR;
                delta += (pN ? 1 : 0) + (pF ? 1 : 0);
            } else if (STATS) st_drop++;
            if (delta != 0) fin = (atomicAdd(&P.pend[slot], delta) + delta) == 0;
        }
        __syncwarp();                   // every lane has read its entry before the pushes below reuse the popped part of the pool
        {
            const bool oneStack = policy == 1 || policy == 3;
            const bool hN = pN && !(cN & REF_LEAF), hF = pF && !(cF & REF_LEAF);      // inner children: near -> hot, far -> cold
            const bool lN = pN && (cN & REF_LEAF), lF = pF && (cF & REF_LEAF);        // leaves: straight to the leaf pool
            const unsigned bH = __ballot_sync(FULL, hN), bC = __ballot_sync(FULL, hF);
            const unsigned bL0 = __ballot_sync(FULL, lF), bL1 = __ballot_sync(FULL, lN);
            if (policy == 3) {          // one stack, every lane's far child directly under its near child: both are popped together
                const int io = hcount + __popc(bC & lt) + __popc(bH & lt);
                if (hF) P.ipool[io] = make_uint2(make_item(cF, slot), __float_as_uint(tF));
                if (hN) P.ipool[io + (hF ? 1 : 0)] = make_uint2(make_item(cN, slot), __float_as_uint(tN));
            } else if (oneStack) {      // far children first, the near ones on top of them
                if (hF) P.ipool[hcount + __popc(bC & lt)] = make_uint2(make_item(cF, slot), __float_as_uint(tF));
                if (hN) P.ipool[hcount + __popc(bC) + __popc(bH & lt)] = make_uint2(make_item(cN, slot), __float_as_uint(tN));
            } else {
                if (hN) P.ipool[hcount + __popc(bH & lt)] = make_uint2(make_item(cN, slot), __float_as_uint(tN));
                if (hF) P.ipool[CAP_I - 1 - ccount - __popc(bC & lt)] = make_uint2(make_item(cF, slot), __float_as_uint(tF));
            }
            if (lF) P.lpool[lcount + __popc(bL0 & lt)] = make_uint2(make_item(cF, slot), __float_as_uint(tF));
            if (lN) P.lpool[lcount + __popc(bL0) + __popc(bL1 & lt)] = make_uint2(make_item(cN, slot), __float_as_uint(tN));
            hcount += __popc(bH) + (oneStack ? __popc(bC) : 0); ccount += oneStack ? 0 : __popc(bC);
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
            atomicAdd(&stats->v[8], (unsigned long long)st_cold << 32);      // packed beside the guard iterations
            atomicMax(&stats->v[9], (unsigned long long)(st_it[0] + st_it[1] + st_it[2] + st_it[3] + st_it[4]));   // most iterations of any warp
        }
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

namespace {
using PoolKernel = void (*)(DeviceScene, FrameParams, uint32_t*, unsigned*, PoolParams, HitRecord*, unsigned*, DeviceCounters*, PoolQueue);

// `warps` = warps per CTA of the kernel picked, `ctas` = CTAs of that size per SM (always 24 or 32 warps per SM). Warps never
// synchronise with each other, so the CTA size only decides how soon a finished warp's share of the SM goes to a waiting CTA
// - of the NEXT frame in flight: a CTA leaves when its slowest warp does.
template <int MODE>
PoolKernel pick_kernel(const Switches& sw, bool stats, size_t& smem, int& ctas, int& warps)
{
    ctas = POOL_CTAS_PER_SM; warps = POOL_WARPS;
    if (sw.pool_small) { smem = sizeof(WarpPool<128>) * POOL_WARPS; return rt_pool_kernel<MODE, 128>; }
    smem = sizeof(WarpPool<512>) * POOL_WARPS;
    if (stats && MODE <= POOL_PRIMARY) return rt_pool_kernel<(MODE <= POOL_PRIMARY ? MODE : 0), 512, true>;
    // Measured on one B200 (C2, frames in flight, L2 flush per frame; profiles/README.md sessions r03b-e): 3 in flight 8 warps 4287 fps,
    // 4 warps 4411, 2 warps 4409; 6 in flight 4386 / 4510 / 4538; alone 0.3205 / 0.3200 / 0.327 ms (1 warp: 0.341); one rank of 8
    // emulated, 8 in flight: 15 264 / 15 438 / 15 118 fps. Default: 4 warps per CTA.
    int w = sw.pool_cta_warps == 8 || sw.pool_cta_warps == 2 || sw.pool_cta_warps == 1 ? sw.pool_cta_warps : 4;
    if (w == 1 && (MODE != POOL_FUSED || sw.pool_occ3)) w = 2;        // one-warp CTAs (32 per SM) exist for the C2-type kernel only
    if (!sw.pool_occ3 && MODE == POOL_FUSED) {          // C2-type frames: 4 CTAs per SM (64 registers per thread, 256-entry pools): 0.339 ms vs 0.355 with 3
        smem = sizeof(WarpPool<256>) * w; ctas = 4 * (POOL_WARPS / w); warps = w;
        if (w == 4) return rt_pool_kernel<POOL_FUSED, 256, false, 4, 4>;
        if (w == 2) return rt_pool_kernel<POOL_FUSED, 256, false, 4, 2>;
        if (w == 1) return rt_pool_kernel<POOL_FUSED, 256, false, 4, 1>;
        return rt_pool_kernel<POOL_FUSED, 256, false, 4>;
    }
    smem = sizeof(WarpPool<512>) * w; ctas = POOL_CTAS_PER_SM * (POOL_WARPS / w); warps = w;
    if (w == 4) return rt_pool_kernel<MODE, 512, false, POOL_CTAS_PER_SM, 4>;
    if (w == 2) return rt_pool_kernel<MODE, 512, false, POOL_CTAS_PER_SM, 2>;
    return rt_pool_kernel<MODE, 512>;
}

void fill_thresholds(PoolParams& pp, const Switches& sw, bool prune)
{
    pp.prune = prune ? 1 : 0;
    pp.policy = sw.pool_policy;
    pp.leafMin = sw.pool_leaf_min > 0 ? sw.pool_leaf_min : LEAF_MIN; pp.sortMin = sw.pool_sort_min > 0 ? sw.pool_sort_min : SORT_MIN;
    pp.shadeMin = sw.pool_shade_min > 0 ? sw.pool_shade_min : SHADE_MIN; pp.refillMin = sw.pool_refill_min > 0 ? sw.pool_refill_min : REFILL_MIN;
    pp.lowWater = sw.pool_low_water > 0 ? sw.pool_low_water : LOW_WATER; pp.dry = sw.pool_dry > 0 ? sw.pool_dry : DRY;
    if (pp.leafMin > 32) pp.leafMin = 32;             // the leaf pool holds < leafMin + 64 + 32 entries
    if (pp.lowWater > 64) pp.lowWater = 64;
}
}  // namespace

cudaError_t rt_pool_configure()
{
    cudaError_t e = cudaSuccess;
    const int big = (int)(sizeof(WarpPool<512>) * POOL_WARPS), small = (int)(sizeof(WarpPool<128>) * POOL_WARPS);
    const int mid = (int)(sizeof(WarpPool<256>) * POOL_WARPS);
    auto set = [&](PoolKernel k, int bytes) { if (e == cudaSuccess) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); };
    set(rt_pool_kernel<POOL_FUSED, 512>, big); set(rt_pool_kernel<POOL_PRIMARY, 512>, big);
    set(rt_pool_kernel<POOL_ANYHIT, 512>, big); set(rt_pool_kernel<POOL_CLOSEST, 512>, big);
    set(rt_pool_kernel<POOL_FUSED, 512, true>, big); set(rt_pool_kernel<POOL_PRIMARY, 512, true>, big);
    set(rt_pool_kernel<POOL_FUSED, 128>, small); set(rt_pool_kernel<POOL_PRIMARY, 128>, small);
    set(rt_pool_kernel<POOL_ANYHIT, 128>, small); set(rt_pool_kernel<POOL_CLOSEST, 128>, small);
    set(rt_pool_kernel<POOL_FUSED, 256, false, 4>, mid);
    set(rt_pool_kernel<POOL_FUSED, 256, false, 4, 4>, mid / 2); set(rt_pool_kernel<POOL_FUSED, 256, false, 4, 2>, mid / 4);
    set(rt_pool_kernel<POOL_FUSED, 256, false, 4, 1>, mid / 8);
    set(rt_pool_kernel<POOL_FUSED, 512, false, POOL_CTAS_PER_SM, 4>, big / 2); set(rt_pool_kernel<POOL_FUSED, 512, false, POOL_CTAS_PER_SM, 2>, big / 4);
    set(rt_pool_kernel<POOL_PRIMARY, 512, false, POOL_CTAS_PER_SM, 4>, big / 2); set(rt_pool_kernel<POOL_PRIMARY, 512, false, POOL_CTAS_PER_SM, 2>, big / 4);
    set(rt_pool_kernel<POOL_ANYHIT, 512, false, POOL_CTAS_PER_SM, 4>, big / 2); set(rt_pool_kernel<POOL_ANYHIT, 512, false, POOL_CTAS_PER_SM, 2>, big / 4);
    set(rt_pool_kernel<POOL_CLOSEST, 512, false, POOL_CTAS_PER_SM, 4>, big / 2); set(rt_pool_kernel<POOL_CLOSEST, 512, false, POOL_CTAS_PER_SM, 2>, big / 4);
    return e;
}

// Primary rays of a frame: clears the frame, then the pooled kernel over the screen rectangle of the root box.
cudaError_t launch_rt_pool(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, bool fused, bool prune, const Switches& sw,
                           unsigned* pixelCounter, void* hits, unsigned* hitCount, int numSMs, cudaStream_t stream, int& launches,
                           DeviceCounters* stats, bool inFlight)
{
    cudaError_t e = cudaMemsetAsync(d_out, 0, (size_t)fp.W * fp.n_rows * 4, stream);        // black; the kernel writes lit pixels only
    if (e != cudaSuccess) return e;
    const int4 bounds = sw.no_root_rect ? make_int4(0, 0, (int)fp.W - 1, (int)fp.H - 1) : root_screen_bounds(sc, fp);
    PoolParams pp;
    pp.tiles = tile_rect(fp, bounds);
    if (pp.tiles.z <= 0 || pp.tiles.w <= 0) return cudaSuccess;
    fill_thresholds(pp, sw, prune);
    pp.nGroups = (unsigned)pp.tiles.z * 2u * (unsigned)pp.tiles.w * 4u;
    pp.groupsPerRow = (unsigned)pp.tiles.z * 2u;
    pp.scatterMul = 0; pp.scatterInv = 0;
    // How pixels are dealt to warps. A frame rendered ALONE is bound by its slowest warp: scattered 4-pixel groups give every warp a
    // cross-section of the frame (C2 kernel alone 0.355 ms vs 0.386 with whole tiles). A frame rendered with others IN FLIGHT is bound
    // by throughput - the other frames fill the gaps - and whole 8x4 tiles keep a warp's rays coherent (2 frames in flight: 3850 fps
    // vs 3170 scattered). pool_scatter = 1 / 2 forces scattered / whole tiles.
    // (a rank's row shard of a many-GPU job stays scattered: few rays per warp, balance matters more - the measured 8-GPU setting)
    const bool scatter = sw.pool_scatter == 1 || (sw.pool_scatter == 0 && (!inFlight || fp.row_step > 1));
    if (scatter && pp.nGroups > 64) {
        // a multiplier near nGroups / golden ratio, coprime to nGroups: consecutive groups land far apart
        unsigned m = (unsigned)((double)pp.nGroups * 0.6180339887498949) | 1u;
        auto gcd = [](unsigned a, unsigned b) { while (b) { const unsigned t = a % b; a = b; b = t; } return a; };
        while (gcd(m, pp.nGroups) != 1u) m += 2;
        pp.scatterMul = m % pp.nGroups;
        pp.scatterInv = ~0ull / pp.nGroups;
    }
    size_t smem; int ctas, warps;
    PoolKernel k = fused ? pick_kernel<POOL_FUSED>(sw, stats != nullptr, smem, ctas, warps) : pick_kernel<POOL_PRIMARY>(sw, stats != nullptr, smem, ctas, warps);
    int grid = numSMs * ctas;
    // Grid: a frame rendered alone (or a rank's row shard of it) is latency-bound and every warp that can take rays shortens it: one 8x4
    // tile per warp (kernel alone on every 8th row of C2: 0.146 ms with 1 tile per warp, 0.235 with 4). With frames IN FLIGHT the other
    // frames hide that latency and fuller warps waste fewer lanes (a warp with its one tile's ~27 rays runs its inner passes at 16.6 of
    // 32 lanes; a full frame's warps at 25.7): one rank of 8 emulated on one GPU, 8 frames in flight, L2 flush per frame:
    // 1 tile per warp 15 260 fps, 2: 17 040, 4: 16 180 (12 in flight: 15 360 / 16 730 / -).
    const int tpw = sw.pool_tiles_per_warp > 0 ? sw.pool_tiles_per_warp : (inFlight ? 2 : 1);
    const int needed = (pp.tiles.z * pp.tiles.w + warps * tpw - 1) / (warps * tpw);
    if (grid > needed) grid = needed;
    PoolQueue q = {nullptr, nullptr, 0u, 0u, 0u, nullptr};
    k<<<grid, warps * 32, smem, stream>>>(sc, fp, d_out, pixelCounter, pp, reinterpret_cast<HitRecord*>(hits), hitCount, stats, q);
    launches += 1;
    return cudaGetLastError();
}

// Secondary rays from a queue of 48-byte records (rt_wavefront.cu): any-hit rays set occ[result] = 1 when blocked, closest-hit rays
// append hit records tagged with their result index. `cursor` must be zero (the queue's read position).
cudaError_t launch_rt_pool_queue(const DeviceScene& sc, const FrameParams& fp, bool anyhit, bool prune, const Switches& sw, unsigned* cursor,
                                 const float4* rays, const unsigned* count, unsigned first, unsigned cap, unsigned stride,
                                 unsigned char* occ, void* hits, unsigned* hitCount, int numSMs, cudaStream_t stream, int& launches)
{
    PoolParams pp;
    memset(&pp, 0, sizeof pp);
    fill_thresholds(pp, sw, prune);
    size_t smem; int ctas, warps;
    PoolKernel k = anyhit ? pick_kernel<POOL_ANYHIT>(sw, false, smem, ctas, warps) : pick_kernel<POOL_CLOSEST>(sw, false, smem, ctas, warps);
    PoolQueue q = {rays, count, first, cap, stride, occ};
    k<<<numSMs * ctas, warps * 32, smem, stream>>>(sc, fp, nullptr, cursor, pp, reinterpret_cast<HitRecord*>(hits), hitCount, nullptr, q);
    launches += 1;
    return cudaGetLastError();
}

}  // namespace b200r
