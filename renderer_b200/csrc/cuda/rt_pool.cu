// rt_pool.cu — the ray tracer's hot kernel: BVH traversal with warp-wide work lists in shared memory.
//
// What it computes is BVH_IntersectTriangles (reference src/Raytracer.cc:183-308) for the primary ray of every pixel and,
// in the common configuration (one light, no reflections, no AO), the shading of the hit and its shadow ray
// (Raytrace, reference src/Raytracer.cc:315-505) - the same rays, the same box and triangle arithmetic (rt_common.cuh), the
// same winner per ray. How the work is scheduled has nothing in common with the reference's one-ray-one-stack loop:
//
//   * A warp owns up to 32 rays at a time ("slots": origin, direction, refined reciprocals, best hit so far, a private stack of
//     deferred subtrees - all in shared memory) and two work lists: the inner nodes and the leaves its rays are standing at.
//     A list entry is 8 bytes: {leaf flag | slot | node or list index, entry distance of the node's box}.
//   * Lanes are not tied to rays. An INNER iteration pops up to 32 entries of the node list and each lane advances one ray by
//     one node: one 64-byte record holds both children's boxes, so the lane does the two slab tests (true IEEE quotients, the
//     reference's per-axis rules), keeps the farther surviving child on the ray's private stack and puts the nearer one back on
//     a list - positions come from warp ballots. A LEAF iteration does the same for up to 32 rays standing at leaves
//     (triangles in list order, then the ray's next subtree is popped from its stack, skipping what its new bound rules out).
//     Slab tests and triangle tests therefore run in separate, densely populated passes instead of sharing a diverged warp.
//   * Per ray this is the reference's depth-first walk, nearest child first, with distance pruning: the
//     closest hit is the minimum of (hitZ, position in the triangle list) over every triangle of every leaf the reference
//     would reach that can still win - the reference's strict `<` in its list-order visit (src/Raytracer.cc:287-296).
//   * A ray that has taken many steps (C2's horizon pixels cross hundreds of boxes: 280 dependent steps in round 1's kernel,
//     the whole frame's critical path) is turned WIDE: its private stack is emptied onto the lists and from then on both
//     children of its nodes go there, so all its pending subtrees are walked at once by as many lanes as are free; results
//     merge with one 64-bit atomicMin on the slot's key, a pending count says when the ray is done. The same happens to every
//     ray with deferred subtrees once the warp runs out of new pixels.
//   * A finished ray is black, or (FUSED) shaded with both outcomes of the light test and re-armed in place as the SHADOW ray
//     of its hit, or (generic configurations) appended as a hit record for rt_shade_kernel. Shading is deferred until several
//     hits wait so that its long arithmetic runs with more than one lane.
//   * Persistent CTAs (3 per SM x 148), 8 independent warps each - no CTA-wide barrier anywhere. Warps take pixels in scattered
//     groups of four from the screen rectangle that can contain the model, build the primary rays themselves and test the root
//     box against kernel arguments; the frame is cleared beforehand, so the 81 % of C2's pixels that miss are never touched.
//   * Nothing can overflow: a private stack that is full, or lists that are nearly full, make the lane walk that subtree
//     depth-first on the spot (walk_subtree - also the path of rays outside the shared-reciprocal divide's domain).
#include "rt_common.cuh"
#include "rt_kernels.cuh"

namespace b200r {
using namespace rt;

namespace {

constexpr int POOL_WARPS = 8;            // warps per CTA (independent of each other)
constexpr int POOL_CTAS_PER_SM = 3;
// scheduling thresholds (defaults; developer switches pool_* override them for tuning)
constexpr int LEAF_MIN = 16;             // run a leaf iteration as soon as this many rays stand at leaves
constexpr int SORT_MIN = 4;              // look at finished slots once this many wait (or the warp is running dry)
constexpr int SHADE_MIN = 8;             // shade resolved hits once this many wait (or the warp is running dry)
constexpr int REFILL_MIN = 8;            // take new pixels once this many slots are free
constexpr int DRY = 16;                  // "running dry": fewer list entries than this
constexpr int WIDE_AFTER = 48;           // a ray turns wide after this many steps
constexpr uint32_t ITEM_LEAF = 0x80000000u;
constexpr uint32_t ITEM_INDEX_MASK = 0x03FFFFFFu;     // 26 bits: inner record id or list position
constexpr uint32_t ITEM_REF_MASK = ITEM_LEAF | ITEM_INDEX_MASK;
constexpr int ITEM_SLOT_SHIFT = 26;
constexpr uint32_t SP_WIDE = 0xFFu;      // meta & 0xFF: depth of the private stack, or this mark
constexpr unsigned long long KEY_EMPTY = ((unsigned long long)0x7F7FFFFFu << 32) | 0xFFFFFFFFull;   // (FLT_MAX, no list position)

// Per-warp state in shared memory. DEPTH = entries of a ray's private stack, CAP = entries of each work list.
template <int DEPTH, int CAP>
struct __align__(16) WarpState {
    uint2 hot[CAP];                      // rays standing at inner nodes (+ every pending inner node of the wide rays)
    uint2 leaf[CAP];                     // rays standing at leaves (+ every pending leaf of the wide rays)
    uint2 stk[DEPTH][32];                // [depth][slot]: deferred subtrees {ref, entry distance}; one lane per ray at a time -> no bank conflicts
    float4 ro[32];                       // slot: ray origin, pruning slack (+inf: never prune; -inf: ray finished, drop its entries)
    float4 rd[32];                       // slot: ray direction
    float4 rr[32];                       // slot: refined reciprocals of the direction, w = bits(triangle to skip): >= 0 marks a SHADOW ray
    unsigned long long key[32];          // slot: (bits(best hitZ) << 32) | list position; shadow slots: (bits(light distance^2) << 32)
    uint32_t meta[32];                   // slot: [7:0] private stack depth or SP_WIDE, [31:8] steps taken
    int pend[32];                        // wide slots: list entries not yet processed
    uint32_t pix[32];                    // slot: (packed row << 16) | x
    uint32_t lit[32], shd[32];           // shadow slots: the two possible pixel words
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
__device__ __forceinline__ bool intersect_leaf(const DeviceScene& sc, const V3& o, const V3& d, uint32_t li, int avoid,
                                               const V3& lightPos, float lightDistSq, unsigned long long& bestK)
{
    const bool isShadow = avoid >= 0;
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

// Cold path: one lane walks a whole subtree depth-first with a stack of its own - the reference's loop with near-first order
// and pruning. Used for rays outside the shared-reciprocal domain (a zero direction component: they take the reference's
// `dir == 0` rule, ray_box<false>) and whenever a private stack or a work list is full. Results go where the lists' lanes put theirs.
__device__ __noinline__ void walk_subtree(const DeviceScene& sc, const RayPrep rp, const int avoid, const V3 lightPos, uint32_t cur, float tcur,
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
                if (intersect_leaf(sc, rp.o, rp.d, cur & ITEM_INDEX_MASK, avoid, lightPos, best, bestK)) {
                    *const_cast<float*>(slackWord) = -__int_as_float(0x7f800000);     // occluded: every other entry of the ray is dropped
                    return;
                }
                if (bestK < k0) atomicMin(key, bestK);
            } else {
                const float4* rec = sc.wnodes + 4 * (size_t)(cur & ITEM_INDEX_MASK);
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

struct PoolParams {
    int4 tiles;                     // first tile column / row, tile columns / rows of the screen rectangle that can contain the model
    int prune;
    int leafMin, sortMin, shadeMin, refillMin, dry, wideAfter;
    unsigned scatterMul;            // 0: tiles are dealt in centre-out order, 32 neighbouring pixels per grab; else: 4-pixel groups
    unsigned nGroups, groupsPerRow; //    are dealt in the order (q * scatterMul) mod nGroups, so every warp holds a cross-section
    unsigned long long scatterInv;  //    of the frame instead of one tile (floor(2^64 / nGroups), for the modulo)
};

// STATS (developer switch pool_stats): per-phase iteration / lane counts are added to DeviceCounters (tools/pool_stats.py).
template <bool FUSED, int DEPTH, int CAP, bool STATS = false>
__global__ void __launch_bounds__(POOL_WARPS * 32, POOL_CTAS_PER_SM)
rt_pool_kernel(DeviceScene sc, FrameParams fp, uint32_t* __restrict__ out, unsigned* __restrict__ pixelCounter, PoolParams pp,
               HitRecord* __restrict__ hits, unsigned* __restrict__ hitCount, DeviceCounters* __restrict__ stats)
{
    static_assert(DEPTH >= 1 && DEPTH <= 32 && CAP >= 160, "list capacity: see the room checks (`tight`, conversion, shading, refill)");
    unsigned st_it[4] = {0, 0, 0, 0}, st_ln[4] = {0, 0, 0, 0}, st_conv = 0, st_walk = 0, st_wide = 0;   // inner, leaf, shade, refill
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using State = WarpState<DEPTH, CAP>;
    State& P = reinterpret_cast<State*>(smem_raw)[threadIdx.x >> 5];
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned FULL = 0xffffffffu;
    const V3 eye = mkv3(fp.eye[0], fp.eye[1], fp.eye[2]);
    const V3 lightPos = mkv3(fp.light_pos[0][0], fp.light_pos[0][1], fp.light_pos[0][2]);
    const int4 tiles = pp.tiles;
    const int tx0 = tiles.x, ty0 = tiles.y, ntx = tiles.z, nty = tiles.w;
    const unsigned total = (unsigned)(ntx * nty) * 32u;
    const float INF = __int_as_float(0x7f800000);
    const int LEAF_MIN = pp.leafMin, SORT_MIN = pp.sortMin, SHADE_MIN = pp.shadeMin, REFILL_MIN = pp.refillMin, DRY = pp.dry;
    const unsigned WIDE_AFTER = (unsigned)pp.wideAfter;

    // warp-uniform bookkeeping
    int hcount = 0, lcount = 0;                     // entries of the two lists
    unsigned freeMask = FULL, doneMask = 0u, shadeMask = 0u, convMask = 0u;   // slots: free / finished / hit waiting for shading / to be turned wide
    bool exhausted = (total == 0u);

    // One ray has been advanced by its lane (inner or leaf iteration): what is left to do for it. Ordinary ray: `sp`/`steps` are its
    // private stack depth and step count; if it has no node to go on with (!go) its stack is popped, skipping subtrees its bound
    // rules out. Returns true when the ray is finished.
    auto next_of_ordinary = [&](uint32_t slot, uint32_t sp, uint32_t steps, float slack, float best, bool& go, uint32_t& ref, float& t,
                                bool allWide, bool& wantWide) -> bool {
        while (!go && sp > 0u) {
            --sp;
            const uint2 e = P.stk[sp][slot];
            if (pruned(__uint_as_float(e.y), slack, best)) continue;
            ref = e.x; t = __uint_as_float(e.y); go = true;
        }
        steps++;
        wantWide = go && sp > 0u && (steps >= WIDE_AFTER || allWide);
        P.meta[slot] = sp | (steps << 8);
        return !go;
    };

    for (;;) {
        // Room in the lists: an inner iteration of wide rays may push 64 entries onto each list, a leaf iteration / shading / a refill 32,
        // a conversion DEPTH. `tight`: wide rays stop spreading (their farther subtrees are walked on the spot) until there is room again.
        const bool tight = CAP - hcount < 72 || CAP - lcount < 72;
        const bool room32 = CAP - hcount >= 40 && CAP - lcount >= 40;
        // ------------------------------------------------------------------ rays that took many steps (or all, once the warp runs out of
        // pixels): the private stack goes onto the lists, from now on every pending subtree of the ray is walked in parallel
        if (convMask) {
            while (convMask) {
                const uint32_t slot = (uint32_t)__ffs(convMask) - 1u;
                convMask &= convMask - 1u;
                const uint32_t sp = P.meta[slot] & 0xFFu;
                if (sp == SP_WIDE || sp == 0u || CAP - hcount < 96 || CAP - lcount < 96) continue;
                bool mine = lane < sp; uint2 e = make_uint2(0u, 0u);
                if (mine) { e = P.stk[lane][slot]; e.x = make_item(e.x, slot); }
                const unsigned bI = __ballot_sync(FULL, mine && !(e.x & ITEM_LEAF)), bL = __ballot_sync(FULL, mine && (e.x & ITEM_LEAF));
                if (mine) { if (e.x & ITEM_LEAF) P.leaf[lcount + __popc(bL & lt)] = e; else P.hot[hcount + __popc(bI & lt)] = e; }
                hcount += __popc(bI); lcount += __popc(bL);
                if (lane == 0) { P.pend[slot] = (int)sp + 1; P.meta[slot] = SP_WIDE; }     // + 1: the node the ray is standing at
                if (STATS) st_conv++;
            }
            __syncwarp();
            continue;
        }
        // ------------------------------------------------------------------ leaves
        if (lcount >= LEAF_MIN || (lcount > 0 && hcount == 0)) {
            const int n = min(32, lcount);
            lcount -= n;
            if (STATS) { st_it[1]++; st_ln[1] += n; }
            bool fin = false, go = false, wantWide = false; uint32_t slot = 0, ref = 0; float t = 0.f;
            const bool allWide = exhausted && hcount + lcount < DRY;
            if ((int)lane < n) {
                const uint2 it = P.leaf[lcount + n - 1 - (int)lane];
                slot = (it.x >> ITEM_SLOT_SHIFT) & 31u;
                const float4 ro = P.ro[slot];
                const uint32_t meta = P.meta[slot];
                const bool wide = (meta & 0xFFu) == SP_WIDE;
                const unsigned long long k0 = P.key[slot];
                float best = __uint_as_float((uint32_t)(k0 >> 32));
                bool occ = false;
                if (!pruned(__uint_as_float(it.y), ro.w, best)) {
                    const float4 rd = P.rd[slot];
                    unsigned long long bestK = k0;
                    occ = intersect_leaf(sc, mkv3(ro.x, ro.y, ro.z), mkv3(rd.x, rd.y, rd.z), it.x & ITEM_INDEX_MASK,
                                         __float_as_int(P.rr[slot].w), lightPos, best, bestK);
                    if (occ) P.ro[slot].w = -INF;
                    else if (bestK < k0) {
                        if (wide) atomicMin(&P.key[slot], bestK); else P.key[slot] = bestK;
                        best = __uint_as_float((uint32_t)(bestK >> 32));
                    }
                }
                if (wide) fin = atomicSub(&P.pend[slot], 1) == 1;
                else fin = next_of_ordinary(slot, occ ? 0u : (meta & 0xFFu), meta >> 8, ro.w, best, go, ref, t, allWide, wantWide);
            }
            const unsigned bI = __ballot_sync(FULL, go && !(ref & REF_LEAF)), bL = __ballot_sync(FULL, go && (ref & REF_LEAF));
            if (go) {
                const uint2 item = make_uint2(make_item(ref, slot), __float_as_uint(t));
                if (ref & REF_LEAF) P.leaf[lcount + __popc(bL & lt)] = item; else P.hot[hcount + __popc(bI & lt)] = item;
            }
            hcount += __popc(bI); lcount += __popc(bL);
            doneMask |= __reduce_or_sync(FULL, fin ? (1u << slot) : 0u);
            convMask |= __reduce_or_sync(FULL, wantWide ? (1u << slot) : 0u);
            __syncwarp();
            continue;
        }
        // ------------------------------------------------------------------ finished slots: shadow rays write their pixel, primary rays
        // that pierced nothing are done (the frame was cleared to black), resolved hits queue up for shading
        const bool dry = hcount + lcount < DRY;
        const int ndone = __popc(doneMask);
        if (ndone >= SORT_MIN || (ndone > 0 && dry)) {
            bool freed = false, hit = false; uint32_t slot = 0;
            if ((int)lane < ndone) {
                slot = __fns(doneMask, 0u, (int)lane + 1);
                if (__float_as_int(P.rr[slot].w) >= 0) {
                    const uint32_t pix = P.pix[slot];
                    out[(size_t)(pix >> 16) * fp.W + (pix & 0xffffu)] = (P.ro[slot].w == -INF) ? P.shd[slot] : P.lit[slot];
                    freed = true;
                } else if (P.key[slot] == KEY_EMPTY) freed = true;
                else hit = true;
            }
            freeMask |= __reduce_or_sync(FULL, freed ? (1u << slot) : 0u);
            shadeMask |= __reduce_or_sync(FULL, hit ? (1u << slot) : 0u);
            doneMask = 0u;
            continue;
        }
        // ------------------------------------------------------------------ resolved hits: shade + shadow ray / hit record
        const int nshade = __popc(shadeMask);
        if ((nshade >= SHADE_MIN || (nshade > 0 && dry)) && room32) {
            if (STATS) { st_it[2]++; st_ln[2] += nshade; }
            bool freed = false, arm = false; uint32_t slot = 0;
            bool record = false; int tri = -1; V3 hitp = eye; float kAB = 0.f, kBC = 0.f, kCA = 0.f;
            if ((int)lane < nshade) {
                slot = __fns(shadeMask, 0u, (int)lane + 1);
                const uint32_t pix = P.pix[slot];
                const size_t o = (size_t)(pix >> 16) * fp.W + (pix & 0xffffu);
                freed = true;
                const unsigned long long k = P.key[slot];
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
                            P.rd[slot] = make_float4(sdir.x, sdir.y, sdir.z, 0.f);
                            P.rr[slot] = make_float4(rp.r.x, rp.r.y, rp.r.z, __int_as_float(tri));
                            P.key[slot] = (unsigned long long)__float_as_uint(ldsq) << 32;
                            if (rp.fast) { P.meta[slot] = 0u; P.lit[slot] = pixLit; P.shd[slot] = pixShadow; freed = false; arm = true; }
                            else {          // outside the shared-reciprocal domain: walked here, by this lane
                                walk_subtree(sc, rp, tri, lightPos, sc.root_ref, -FLT_MAX, &P.key[slot], &P.ro[slot].w);
                                out[o] = (P.ro[slot].w == -INF) ? pixShadow : pixLit;
                            }
                        }
                    }
                } else record = true;
            }
            if (FUSED) {
                const unsigned am = __ballot_sync(FULL, arm);
                if (am) {
                    const uint2 item = make_uint2(make_item(sc.root_ref, slot), __float_as_uint(-FLT_MAX));
                    if (sc.root_ref & REF_LEAF) { if (arm) P.leaf[lcount + __popc(am & lt)] = item; lcount += __popc(am); }
                    else { if (arm) P.hot[hcount + __popc(am & lt)] = item; hcount += __popc(am); }
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
        if (!exhausted && room32 && (__popc(freeMask) >= REFILL_MIN || hcount + lcount == 0)) {
            const int nfree = __popc(freeMask);
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(pixelCounter, (unsigned)nfree);
            base = __shfl_sync(FULL, base, 0);
            if (base + (unsigned)nfree >= total) exhausted = true;
            const unsigned g = base + lane;
            bool enter = false; RayPrep rp; int x = 0, r = 0;
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
            uint32_t slot = 0; bool walked = false;
            if (enter) {
                slot = __fns(freeMask, 0u, __popc(em & lt) + 1);
                float slack = INF;
                if (pp.prune) {
                    // 1/|d| per axis (IEEE divide; +inf for a zero component switches pruning off for this ray)
                    const float m = fmaxf(fmaxf(1.0f / fabsf(rp.d.x), 1.0f / fabsf(rp.d.y)), 1.0f / fabsf(rp.d.z));
                    slack = 1e-4f * m + 1e-4f;
                }
                P.ro[slot] = make_float4(rp.o.x, rp.o.y, rp.o.z, slack);
                P.rd[slot] = make_float4(rp.d.x, rp.d.y, rp.d.z, 0.f);
                P.rr[slot] = make_float4(rp.r.x, rp.r.y, rp.r.z, __int_as_float(-1));
                P.key[slot] = KEY_EMPTY;
                P.pix[slot] = ((uint32_t)r << 16) | (uint32_t)x;
                P.meta[slot] = 0u;
                if (!rp.fast) {     // outside the shared-reciprocal domain: walked here, by this lane, then resolved like any other
                    walk_subtree(sc, rp, -1, lightPos, sc.root_ref, -FLT_MAX, &P.key[slot], &P.ro[slot].w);
                    walked = true;
                }
            }
            const unsigned pm = __ballot_sync(FULL, enter && !walked);
            if (enter && !walked) {
                const uint2 item = make_uint2(make_item(sc.root_ref, slot), __float_as_uint(-FLT_MAX));
                if (sc.root_ref & REF_LEAF) P.leaf[lcount + __popc(pm & lt)] = item;
                else P.hot[hcount + __popc(pm & lt)] = item;
            }
            if (sc.root_ref & REF_LEAF) lcount += __popc(pm); else hcount += __popc(pm);
            freeMask &= ~__reduce_or_sync(FULL, enter ? (1u << slot) : 0u);
            doneMask |= __reduce_or_sync(FULL, walked ? (1u << slot) : 0u);
            __syncwarp();
            continue;
        }
        if (hcount == 0) break;            // nothing pending, nothing waiting, no pixels left

        // ------------------------------------------------------------------ inner nodes
        const int n = min(32, hcount);
        hcount -= n;
        if (STATS) { st_it[0]++; st_ln[0] += n; }
        bool fin = false, go = false, goFar = false, wantWide = false;
        uint32_t slot = 0, ref = 0, refFar = 0; float t = 0.f, tFar = 0.f;
        const bool allWide = exhausted && hcount + lcount < DRY;
        if ((int)lane < n) {
            const uint2 it = P.hot[hcount + n - 1 - (int)lane];
            slot = (it.x >> ITEM_SLOT_SHIFT) & 31u;
            const float4 ro = P.ro[slot];
            const uint32_t meta = P.meta[slot];
            const bool wide = (meta & 0xFFu) == SP_WIDE;
            const float best = __uint_as_float(reinterpret_cast<const uint32_t*>(&P.key[slot])[1]);
            const float tHere = __uint_as_float(it.y);
            if (!pruned(tHere, ro.w, best)) {
                const float4 rd = P.rd[slot], rr = P.rr[slot];
                const float4* rec = sc.wnodes + 4 * (size_t)(it.x & ITEM_INDEX_MASK);
                const float4 bx = __ldg(rec + 0), by = __ldg(rec + 1), bz = __ldg(rec + 2), rf = __ldg(rec + 3);
                const uint32_t L = __float_as_uint(rf.x), R = __float_as_uint(rf.y);
                // both boxes, unconditionally (straight-line code). A LEAF child has no box test in the reference
                // (src/Raytracer.cc:224-229): it survives unless empty; its box only supplies a tighter entry distance.
                float tL, tR;
                bool hitL = box_fast(ro, rd, rr, bx.x, bx.y, by.x, by.y, bz.x, bz.y, tL);
                bool hitR = box_fast(ro, rd, rr, bx.z, bx.w, by.z, by.w, bz.z, bz.w, tR);
                if (L & REF_LEAF) { if (!hitL) tL = tHere; hitL = (L != REF_EMPTY); }
                if (R & REF_LEAF) { if (!hitR) tR = tHere; hitR = (R != REF_EMPTY); }
                const uint32_t unprunable = __float_as_uint(rf.z);        // bit0/bit1: L/R subtree holds a triangle that failed the upload check
                if (unprunable & 1u) tL = -FLT_MAX;
                if (unprunable & 2u) tR = -FLT_MAX;
                if (hitL && pruned(tL, ro.w, best)) hitL = false;
                if (hitR && pruned(tR, ro.w, best)) hitR = false;
                const bool rFirst = hitR && (!hitL || tR < tL);
                ref = rFirst ? R : L; t = rFirst ? tR : tL; go = hitL || hitR;
                refFar = rFirst ? L : R; tFar = rFirst ? tL : tR; goFar = hitL && hitR;
            }
            if (wide) {
                if (STATS) st_wide++;
                if (tight && goFar) {           // no room to spread: the farther subtree is walked here and now
                    RayPrep rp; rp.o = mkv3(ro.x, ro.y, ro.z);
                    const float4 rd = P.rd[slot], rr = P.rr[slot];
                    rp.d = mkv3(rd.x, rd.y, rd.z); rp.r = mkv3(rr.x, rr.y, rr.z); rp.fast = true;
                    walk_subtree(sc, rp, __float_as_int(rr.w), lightPos, refFar, tFar, &P.key[slot], &P.ro[slot].w);
                    goFar = false;
                    if (STATS) st_walk++;
                }
                const int delta = -1 + (go ? 1 : 0) + (goFar ? 1 : 0);
                if (delta != 0) fin = (atomicAdd(&P.pend[slot], delta) + delta) == 0;
            } else {
                uint32_t sp = meta & 0xFFu;
                if (goFar) {
                    if (sp < (uint32_t)DEPTH) { P.stk[sp][slot] = make_uint2(refFar, __float_as_uint(tFar)); sp++; }
                    else {                      // private stack full: the farther subtree is walked here and now
                        RayPrep rp; rp.o = mkv3(ro.x, ro.y, ro.z);
                        const float4 rd = P.rd[slot], rr = P.rr[slot];
                        rp.d = mkv3(rd.x, rd.y, rd.z); rp.r = mkv3(rr.x, rr.y, rr.z); rp.fast = true;
                        walk_subtree(sc, rp, __float_as_int(rr.w), lightPos, refFar, tFar, &P.key[slot], &P.ro[slot].w);
                        if (STATS) st_walk++;
                    }
                    goFar = false;
                }
                const float bestNow = __uint_as_float(reinterpret_cast<const volatile uint32_t*>(&P.key[slot])[1]);
                fin = next_of_ordinary(slot, sp, meta >> 8, P.ro[slot].w, bestNow, go, ref, t, allWide, wantWide);
            }
        }
        {
            const unsigned bI = __ballot_sync(FULL, go && !(ref & REF_LEAF)), bL = __ballot_sync(FULL, go && (ref & REF_LEAF));
            if (go) {
                const uint2 item = make_uint2(make_item(ref, slot), __float_as_uint(t));
                if (ref & REF_LEAF) P.leaf[lcount + __popc(bL & lt)] = item; else P.hot[hcount + __popc(bI & lt)] = item;
            }
            hcount += __popc(bI); lcount += __popc(bL);
            if (__any_sync(FULL, goFar)) {          // wide rays only: the farther child goes onto the lists as well
                const unsigned cI = __ballot_sync(FULL, goFar && !(refFar & REF_LEAF)), cL = __ballot_sync(FULL, goFar && (refFar & REF_LEAF));
                if (goFar) {
                    const uint2 item = make_uint2(make_item(refFar, slot), __float_as_uint(tFar));
                    if (refFar & REF_LEAF) P.leaf[lcount + __popc(cL & lt)] = item; else P.hot[hcount + __popc(cI & lt)] = item;
                }
                hcount += __popc(cI); lcount += __popc(cL);
            }
        }
        doneMask |= __reduce_or_sync(FULL, fin ? (1u << slot) : 0u);
        convMask |= __reduce_or_sync(FULL, wantWide ? (1u << slot) : 0u);
        __syncwarp();
    }
    if (STATS) {
        st_walk = __reduce_add_sync(FULL, st_walk); st_wide = __reduce_add_sync(FULL, st_wide);
        if (lane == 0) {
            for (int i = 0; i < 4; i++) { atomicAdd(&stats->v[2 * i], (unsigned long long)st_it[i]); atomicAdd(&stats->v[2 * i + 1], (unsigned long long)st_ln[i]); }
            atomicAdd(&stats->v[8], ((unsigned long long)st_walk << 32) | st_conv);      // subtrees walked on the spot | rays turned wide
            atomicMax(&stats->v[9], (unsigned long long)(st_it[0] + st_it[1] + st_it[2] + st_it[3]));   // most iterations of any warp
            atomicAdd(&stats->v[10], (unsigned long long)st_wide);                       // inner steps of wide rays
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

cudaError_t rt_pool_configure()
{
    cudaError_t e = cudaSuccess;
    const int big = (int)(sizeof(WarpState<16, 160>) * POOL_WARPS), small = (int)(sizeof(WarpState<2, 160>) * POOL_WARPS);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<true, 16, 160>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<false, 16, 160>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<true, 16, 160, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<false, 16, 160, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<true, 2, 160>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(rt_pool_kernel<false, 2, 160>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
    return e;
}

cudaError_t launch_rt_pool(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, bool fused, bool prune, const Switches& sw,
                           unsigned* pixelCounter, void* hits, unsigned* hitCount, int numSMs, cudaStream_t stream, int& launches,
                           DeviceCounters* stats)
{
    cudaError_t e = cudaMemsetAsync(d_out, 0, (size_t)fp.W * fp.n_rows * 4, stream);        // black; the kernel writes lit pixels only
    if (e != cudaSuccess) return e;
    const int4 bounds = sw.no_root_rect ? make_int4(0, 0, (int)fp.W - 1, (int)fp.H - 1) : root_screen_bounds(sc, fp);
    PoolParams pp;
    pp.tiles = tile_rect(fp, bounds);
    if (pp.tiles.z <= 0 || pp.tiles.w <= 0) return cudaSuccess;
    pp.prune = prune ? 1 : 0;
    pp.leafMin = sw.pool_leaf_min > 0 ? sw.pool_leaf_min : LEAF_MIN; pp.sortMin = sw.pool_sort_min > 0 ? sw.pool_sort_min : SORT_MIN;
    pp.shadeMin = sw.pool_shade_min > 0 ? sw.pool_shade_min : SHADE_MIN; pp.refillMin = sw.pool_refill_min > 0 ? sw.pool_refill_min : REFILL_MIN;
    pp.dry = sw.pool_dry > 0 ? sw.pool_dry : DRY; pp.wideAfter = sw.pool_wide_after > 0 ? sw.pool_wide_after : WIDE_AFTER;
    if (pp.leafMin > 32) pp.leafMin = 32;
    pp.nGroups = (unsigned)pp.tiles.z * 2u * (unsigned)pp.tiles.w * 4u;
    pp.groupsPerRow = (unsigned)pp.tiles.z * 2u;
    pp.scatterMul = 0; pp.scatterInv = 0;
    if (!sw.pool_no_scatter && pp.nGroups > 64) {
        // a multiplier near nGroups / golden ratio, coprime to nGroups: consecutive groups land far apart
        unsigned m = (unsigned)((double)pp.nGroups * 0.6180339887498949) | 1u;
        auto gcd = [](unsigned a, unsigned b) { while (b) { const unsigned t = a % b; a = b; b = t; } return a; };
        while (gcd(m, pp.nGroups) != 1u) m += 2;
        pp.scatterMul = m % pp.nGroups;
        pp.scatterInv = ~0ull / pp.nGroups;
    }
    using K = void (*)(DeviceScene, FrameParams, uint32_t*, unsigned*, PoolParams, HitRecord*, unsigned*, DeviceCounters*);
    K k; size_t smem;
    // pool_small: 2-entry private stacks, so the parity tests exercise walk_subtree and the wide conversion all the time
    if (sw.pool_small) { k = fused ? rt_pool_kernel<true, 2, 160> : rt_pool_kernel<false, 2, 160>; smem = sizeof(WarpState<2, 160>) * POOL_WARPS; }
    else { k = fused ? rt_pool_kernel<true, 16, 160> : rt_pool_kernel<false, 16, 160>; smem = sizeof(WarpState<16, 160>) * POOL_WARPS; }
    if (stats && !sw.pool_small) k = fused ? rt_pool_kernel<true, 16, 160, true> : rt_pool_kernel<false, 16, 160, true>;
    int grid = numSMs * POOL_CTAS_PER_SM;
    const int needed = (pp.tiles.z * pp.tiles.w + POOL_WARPS - 1) / POOL_WARPS;       // one tile per warp is the least a warp can take
    if (grid > needed) grid = needed;
    k<<<grid, POOL_WARPS * 32, smem, stream>>>(sc, fp, d_out, pixelCounter, pp, reinterpret_cast<HitRecord*>(hits), hitCount, stats);
    launches += 1;
    return cudaGetLastError();
}

}  // namespace b200r
