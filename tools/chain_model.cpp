// chain_model.cpp — developer tool (CPU, no GPU): how long are the dependency chains of a ray-traced frame, and what would other
// job-splitting policies do to them?
//
//   g++ -O2 -std=c++17 -ffp-contract=off -fopenmp tools/chain_model.cpp -Iinclude -Lrenderer_b200 -lb200render \
//       -Wl,-rpath,$PWD/renderer_b200 -o /tmp/chain_model && /tmp/chain_model oracle/_ref/models/chessboard.tri 1920 1080 10
//
// profiles/README.md: the C2 frame is bound by its longest chain (a ~100-step primary job of a horizon pixel followed by its
// shadow ray), not by throughput. This tool restates the SCHEDULING-RELEVANT part of rt_primary_kernel on the CPU - the same jobs
// ((pixel, subtree) after `split` BVH levels), near-first traversal with the same distance pruning rule, one "step" = one inner
// node or one triangle test, the shadow ray after the pixel's jobs - and reports, per policy:
//   total steps (work), jobs, and the critical path per pixel = longest primary job + longest shadow job (steps).
// Policies: primary split depth; whether a pixel's jobs share their best hit while they run ("shared": what pruning against the
// pixel's merge word on every pop would give, idealised as lock-step execution); shadow rays split into subtree jobs as well.
// It is a MODEL (flat shading normal for the cast/no-cast decision, no unprunable flags, ideal sharing) - it is validated against
// the job-length histograms the GPU job profiler measured (tools/job_profile.py; printed first), not against pixels.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "b200render.h"

struct V3 { float x, y, z; };
static inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
static inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 norm(V3 v) { const float n = sqrtf(dot(v, v)); return mk(v.x / n, v.y / n, v.z / n); }
static inline float dsq(V3 a, V3 b) { const V3 d = a - b; return dot(d, d); }
static inline V3 v3(const float* p) { return mk(p[0], p[1], p[2]); }

struct Scene {
    const b200r_vertex* v; const b200r_tri* t; const b200r_bvhnode* n; const int32_t* idx;
    uint32_t nv, nt, nn, nidx;
};

static bool ray_box(const V3& o, const V3& d, const float* lo, const float* hi, float& tnear)
{
    float Tn = -FLT_MAX, Tf = FLT_MAX;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    bool ok = true;
    for (int a = 0; a < 3; a++) {
        if (dd[a] == 0.f) { if (oo[a] < lo[a] || oo[a] > hi[a]) ok = false; }
        else {
            float T1 = (lo[a] - oo[a]) / dd[a], T2 = (hi[a] - oo[a]) / dd[a];
            if (T1 > T2) std::swap(T1, T2);
            if (T1 > Tn) Tn = T1;
            if (T2 < Tf) Tf = T2;
        }
    }
    if (Tn > Tf || Tf < 0.f) ok = false;
    tnear = Tn;
    return ok;
}

static inline bool is_leaf(const b200r_bvhnode& n) { return (n.a & 0x80000000u) != 0; }

// One traversal job as a state machine: step() advances by one inner node or one triangle test.
struct Job {
    const Scene* sc; V3 o, d; float slack;
    bool shadow; V3 light; float lightBest; int avoid;
    float* best; uint32_t* bestLi; int* bestTri; V3* bestHit;       // closest hit so far (may be shared by the jobs of a pixel)
    bool* occluded;                                                 // shadow: some job found an occluder (may be shared)
    std::vector<std::pair<uint32_t, float>> stack;
    uint32_t cur; int leafPos; bool done; int steps;

    void init(const Scene* s, V3 o_, V3 d_, uint32_t node)
    {
        sc = s; o = o_; d = d_; cur = node; leafPos = 0; done = false; steps = 0; stack.clear();
        const float m = std::max(std::max(1.0f / fabsf(d.x), 1.0f / fabsf(d.y)), 1.0f / fabsf(d.z));
        slack = shadow ? INFINITY : 1e-4f * m + 1e-4f;
    }
    bool pruned(float t) const { const float e = t - slack; return e > 0.f && (e * e) * 0.99999f > *best; }
    void pop()
    {
        while (!stack.empty()) {
            auto [n, t] = stack.back(); stack.pop_back();
            if (!shadow && pruned(t)) continue;
            cur = n; leafPos = 0; return;
        }
        done = true;
    }
    void step()
    {
        if (done) return;
        if (shadow && *occluded) { done = true; return; }
        const b200r_bvhnode& n = sc->n[cur];
        steps++;
        if (is_leaf(n)) {
            const uint32_t cnt = n.a & 0x7fffffffu;
            if (cnt == 0) { steps--; pop(); return; }
            const uint32_t li = n.b + (uint32_t)leafPos;
            const int ti = sc->idx[li];
            const b200r_tri& T = sc->t[ti];
            bool alive = !(shadow && ti == avoid);
            const V3 nn = v3(T.normal);
            if (alive && !T.two_sided && dot(o - v3(T.center), nn) < 0.f) alive = false;
            if (alive) {
                const float k = dot(nn, d);
                if (k != 0.f) {
                    const float s = (T.d - dot(nn, o)) / k;
                    if (s > 0.f && s > 1e-5f) {
                        const V3 hit = d * s + o;
                        if (!(dot(v3(T.e1), hit) - T.d1 < 0.f) && !(dot(v3(T.e2), hit) - T.d2 < 0.f) && !(dot(v3(T.e3), hit) - T.d3 < 0.f)) {
                            if (shadow) { if (dsq(light, hit) < lightBest) { *occluded = true; done = true; return; } }
                            else {
                                const float z = dsq(o, hit);
                                if (z < *best || (z == *best && li < *bestLi)) { *best = z; *bestLi = li; *bestTri = ti; *bestHit = hit; }
                            }
                        }
                    }
                }
            }
            if (++leafPos >= (int)cnt) pop();
            return;
        }
        float tL = -FLT_MAX, tR = -FLT_MAX;
        const b200r_bvhnode &L = sc->n[n.a], &R = sc->n[n.b];
        bool hL = is_leaf(L) ? (L.a & 0x7fffffffu) != 0 : ray_box(o, d, L.lo, L.hi, tL);
        bool hR = is_leaf(R) ? (R.a & 0x7fffffffu) != 0 : ray_box(o, d, R.lo, R.hi, tR);
        if (is_leaf(L)) tL = -FLT_MAX;
        if (is_leaf(R)) tR = -FLT_MAX;
        if (!shadow) { if (hL && pruned(tL)) hL = false; if (hR && pruned(tR)) hR = false; }
        if (hL && hR) {
            const bool rFirst = tR < tL;
            stack.push_back({rFirst ? n.a : n.b, rFirst ? tL : tR});
            cur = rFirst ? n.b : n.a; leafPos = 0;
        } else if (hL) { cur = n.a; leafPos = 0; }
        else if (hR) { cur = n.b; leafPos = 0; }
        else pop();
    }
};

// the subtrees still alive after `depth` levels below `root` (the child-box tests the traversal would do)
static void expand(const Scene& sc, const V3& o, const V3& d, uint32_t root, int depth, std::vector<uint32_t>& out)
{
    out.clear(); out.push_back(root);
    std::vector<uint32_t> nxt;
    for (int l = 0; l < depth; l++) {
        nxt.clear();
        for (uint32_t r : out) {
            const b200r_bvhnode& n = sc.n[r];
            if (is_leaf(n)) { nxt.push_back(r); continue; }
            for (uint32_t c : {n.a, n.b}) {
                const b200r_bvhnode& C = sc.n[c];
                float t;
                if (is_leaf(C) ? (C.a & 0x7fffffffu) != 0 : ray_box(o, d, C.lo, C.hi, t)) nxt.push_back(c);
            }
        }
        out.swap(nxt);
    }
}

// donateAfter > 0: a job that has run that many steps hands the BOTTOM entry of its stack (its largest pending subtree) to a new
// job every step from then on (what rt_primary_kernel's donation does once the queue is empty, here: always, to unlimited lanes)
struct Policy { int split; bool shared; int shadowSplit; int donateAfter; const char* name; bool needHit = false; bool shadowDonate = true; };

// Lock-step execution of the jobs of one ray: every running job advances one step per round; with donation, a job that has run
// `donateAfter` steps gives its bottom stack entry to a new job (same ray, same shared state) that starts in the next round.
// Returns the number of rounds = the length of the ray's dependency chain under that policy.
struct Priv { float best; uint32_t li; int tri; V3 hit; };
static int run(std::vector<Job>& jobs, int donateAfter, std::deque<Priv>* priv = nullptr, bool needHit = false)
{
    int rounds = 0;
    for (bool any = true; any;) {
        any = false;
        const size_t n = jobs.size();
        for (size_t j = 0; j < n; j++) {
            if (jobs[j].done) continue;
            jobs[j].step(); any = true;
            if (donateAfter > 0 && !jobs[j].done && jobs[j].steps >= donateAfter && !jobs[j].stack.empty() &&
                (jobs[j].shadow || !needHit || *jobs[j].best < FLT_MAX)) {
                const auto entry = jobs[j].stack.front();
                jobs[j].stack.erase(jobs[j].stack.begin());
                Job D = jobs[j];                       // same ray, same pointers to the shared state
                D.stack.clear(); D.cur = entry.first; D.leafPos = 0; D.done = false; D.steps = 0;
                if (!D.shadow && D.pruned(entry.second)) continue;
                if (priv && !D.shadow) {               // not shared: the part starts from its donor's bound and then goes its own way
                    priv->push_back(Priv{*D.best, *D.bestLi, -1, *D.bestHit});
                    Priv& p = priv->back();
                    D.best = &p.best; D.bestLi = &p.li; D.bestTri = &p.tri; D.bestHit = &p.hit;
                }
                jobs.push_back(D);
            }
        }
        if (any) rounds++;
    }
    return rounds;
}

struct Stats {
    unsigned long long steps = 0, jobs = 0, shadowRays = 0;
    std::vector<int> primHit, primMiss, shadowLit, shadowBlocked, cp;
};

static void pct(const char* what, std::vector<int>& v)
{
    if (v.empty()) { printf("    %-16s none\n", what); return; }
    std::sort(v.begin(), v.end());
    double sum = 0; for (int x : v) sum += x;
    auto q = [&](double p) { return v[std::min(v.size() - 1, (size_t)(p * v.size()))]; };
    printf("    %-16s n %8zu  mean %6.1f  p50 %4d  p90 %4d  p99 %4d  p99.9 %4d  max %4d\n", what, v.size(), sum / v.size(), q(.5), q(.9), q(.99), q(.999), v.back());
}

int main(int argc, char** argv)
{
    if (argc < 2) { fprintf(stderr, "usage: chain_model MODEL [W H FRAME]\n"); return 2; }
    const int W = argc > 2 ? atoi(argv[2]) : 1920, H = argc > 3 ? atoi(argv[3]) : 1080, frameNo = argc > 4 ? atoi(argv[4]) : 10;
    b200r_scene* s = nullptr;
    if (b200r_scene_load(argv[1], &s)) { fprintf(stderr, "%s\n", b200r_last_error(nullptr)); return 1; }
    const std::string cache = std::string(argv[1]) + ".bvh";
    if (b200r_scene_build_bvh(s, cache.c_str(), 0)) { fprintf(stderr, "%s\n", b200r_last_error(nullptr)); return 1; }
    Scene sc;
    sc.v = b200r_scene_vertices(s, &sc.nv); sc.t = b200r_scene_tris(s, &sc.nt);
    sc.n = b200r_scene_nodes(s, &sc.nn); sc.idx = b200r_scene_tri_idx(s, &sc.nidx);
    b200r_orbit orbit; b200r_orbit_init(&orbit);
    float eye[3], mv[9];
    for (int k = 0; k <= frameNo; k++) b200r_orbit_step(&orbit, eye, mv);
    b200r_frame f; b200r_frame_defaults(&f, B200R_MODE_RAYTRACE, W, H, eye, mv, 1);
    const V3 E = v3(eye), Lp = v3(f.lights[0].pos);
    printf("%s %dx%d frame %d: %u triangles, %u nodes\n", argv[1], W, H, frameNo, sc.nt, sc.nn);

    const Policy policies[] = {
        {2, false, 0, 0, "split 2 (what rt_primary_kernel does while its queue is not empty)"},
        {4, false, 0, 0, "split 4"},
        {2, true, 0, 0, "split 2, jobs of a pixel share their best hit"},
        {6, true, 4, 0, "split 6 shared, shadow split 4"},
        {8, true, 5, 0, "split 8 shared, shadow split 5"},
        {2, true, 0, 48, "split 2 shared, donate after 48 steps"},
        {2, true, 0, 32, "split 2 shared, donate after 32 steps"},
        {2, true, 0, 16, "split 2 shared, donate after 16 steps"},
        {2, true, 0, 8, "split 2 shared, donate after 8 steps"},
        {2, true, 2, 16, "split 2 shared, shadow split 2, donate after 16 steps"},
        {2, false, 0, 32, "split 2 NOT shared, primary jobs with a hit donate after 32 steps (= B200R_URGENT_T=32)", true, false},
        {2, false, 0, 16, "split 2 NOT shared, primary jobs with a hit donate after 16 steps (= B200R_URGENT_T=16)", true, false},
        {2, false, 0, 32, "split 2 NOT shared, primary jobs donate after 32 steps, hit or not (= B200R_URGENT_T=32 B200R_URGENT_NOHIT=1)", false, false},
        {2, false, 0, 32, "split 2 NOT shared, donate after 32 steps (a donated part starts from its donor's bound only)"},
        {2, false, 0, 16, "split 2 NOT shared, donate after 16 steps"},
    };
    for (const Policy& P : policies) {
        Stats S;
#pragma omp parallel
        {
            Stats T;
            std::vector<uint32_t> subs, ssubs;
            std::vector<Job> jobs, sjobs;
#pragma omp for schedule(dynamic, 8)
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    const float SD = (float)(H * 2);
                    const V3 rc = norm(mk(((float)(H / 2) - (float)y) / SD, ((float)x - (float)(W / 2)) / SD, 1.0f));
                    V3 rw = mk(mv[0], mv[1], mv[2]) * rc.x;
                    rw = rw + mk(mv[3], mv[4], mv[5]) * rc.y;
                    rw = rw + mk(mv[6], mv[7], mv[8]) * rc.z;
                    const V3 d = norm(rw);
                    float t;
                    if (is_leaf(sc.n[0]) ? false : !ray_box(E, d, sc.n[0].lo, sc.n[0].hi, t)) continue;
                    expand(sc, E, d, 0, P.split, subs);
                    if (subs.empty()) continue;
                    jobs.clear(); jobs.reserve(subs.size() + 64);
                    float best = FLT_MAX; uint32_t bestLi = 0xFFFFFFFFu; int bestTri = -1; V3 bestHit = E;
                    std::vector<float> jb(subs.size(), FLT_MAX); std::vector<uint32_t> jl(subs.size(), 0xFFFFFFFFu);
                    std::vector<int> jt(subs.size(), -1); std::vector<V3> jh(subs.size(), E);
                    jobs.resize(subs.size());
                    for (size_t j = 0; j < subs.size(); j++) {
                        Job& J = jobs[j];
                        J.shadow = false; J.avoid = -1; J.occluded = nullptr;
                        if (P.shared) { J.best = &best; J.bestLi = &bestLi; J.bestTri = &bestTri; J.bestHit = &bestHit; }
                        else { J.best = &jb[j]; J.bestLi = &jl[j]; J.bestTri = &jt[j]; J.bestHit = &jh[j]; }
                        J.init(&sc, E, d, subs[j]);
                    }
                    // lock-step: every running job advances one step per round (sharing, if on, is then immediate)
                    std::deque<Priv> priv;
                    int longest = run(jobs, P.donateAfter, P.shared ? nullptr : &priv, P.needHit);
                    for (const Priv& p : priv)
                        if (p.tri >= 0 && (p.best < best || (p.best == best && p.li < bestLi))) { best = p.best; bestLi = p.li; bestTri = p.tri; bestHit = p.hit; }
                    for (size_t j = subs.size(); j < jobs.size(); j++) { T.steps += jobs[j].steps; T.jobs++; }     // donated parts
                    for (size_t j = 0; j < subs.size(); j++) {
                        const Job& J = jobs[j];
                        T.steps += J.steps; T.jobs++;
                        if (!P.shared) {
                            if (jb[j] < best || (jb[j] == best && jl[j] < bestLi)) { best = jb[j]; bestLi = jl[j]; bestTri = jt[j]; bestHit = jh[j]; }
                            (jt[j] >= 0 ? T.primHit : T.primMiss).push_back(J.steps);
                        }
                    }
                    int shadowLongest = 0;
                    if (bestTri >= 0) {
                        // flat-normal stand-in for "the light faces the surface" (the kernel decides with the Phong normal)
                        V3 nn = v3(sc.t[bestTri].normal);
                        if (dot(nn, E - bestHit) < 0.f) nn = nn * -1.f;
                        const V3 toL = Lp - bestHit;
                        if (dot(nn, toL) > 0.f) {
                            const V3 sd = norm(toL);
                            bool occ = false;
                            if (!is_leaf(sc.n[0]) && ray_box(bestHit, sd, sc.n[0].lo, sc.n[0].hi, t)) {
                                expand(sc, bestHit, sd, 0, P.shadowSplit, ssubs);
                                sjobs.resize(ssubs.size());
                                for (size_t j = 0; j < ssubs.size(); j++) {
                                    Job& J = sjobs[j];
                                    J.shadow = true; J.light = Lp; J.lightBest = dsq(bestHit, Lp); J.avoid = bestTri; J.occluded = &occ;
                                    J.best = &best; J.bestLi = &bestLi; J.bestTri = &bestTri; J.bestHit = &bestHit;
                                    J.init(&sc, bestHit, sd, ssubs[j]);
                                }
                                shadowLongest = run(sjobs, P.shadowDonate ? P.donateAfter : 0);
                                for (const Job& J : sjobs) { T.steps += J.steps; T.jobs++; }
                                T.shadowRays++;
                                if (P.shadowSplit == 0) (occ ? T.shadowBlocked : T.shadowLit).push_back(sjobs[0].steps);
                            }
                        }
                    }
                    T.cp.push_back(longest + shadowLongest);
                }
#pragma omp critical
            {
                S.steps += T.steps; S.jobs += T.jobs; S.shadowRays += T.shadowRays;
                S.primHit.insert(S.primHit.end(), T.primHit.begin(), T.primHit.end());
                S.primMiss.insert(S.primMiss.end(), T.primMiss.begin(), T.primMiss.end());
                S.shadowLit.insert(S.shadowLit.end(), T.shadowLit.begin(), T.shadowLit.end());
                S.shadowBlocked.insert(S.shadowBlocked.end(), T.shadowBlocked.begin(), T.shadowBlocked.end());
                S.cp.insert(S.cp.end(), T.cp.begin(), T.cp.end());
            }
        }
        printf("== %s\n    jobs %llu  shadow rays %llu  total steps %llu (%.2f M)\n", P.name, S.jobs, S.shadowRays, S.steps, S.steps / 1e6);
        if (!P.shared) { pct("prim-hit jobs", S.primHit); pct("prim-miss jobs", S.primMiss); }
        if (P.shadowSplit == 0) { pct("shadow lit", S.shadowLit); pct("shadow blocked", S.shadowBlocked); }
        pct("critical path/px", S.cp);
    }
    b200r_scene_free(s);
    return 0;
}
