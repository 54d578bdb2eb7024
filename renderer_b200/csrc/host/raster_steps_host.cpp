// raster_steps_host.cpp — the span walkers of csrc/raster_steps.h run on the host over random spans.
//
// NOT a rendering path: a test hook for the CPU suite. It checks that walk_span and the batched walk_span_keyed (what the
// device's depth / resolve passes call per span) visit exactly the pixels, with exactly the interpolant bits, of the
// per-scanline loop of Screen::RasterizeTriangle as the reference writes it (src/Screen.h:244-289), restated below.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/b200render.h"
#include "../raster_steps.h"

namespace b200r { void set_global_error(const std::string& s); }

namespace {
using namespace b200r;

struct Visit { int x; uint32_t bits[8]; unsigned long long key; };

template <int N> Visit mk(int x, const FPd<N>& v, unsigned long long key)
{
    Visit r; memset(&r, 0, sizeof r);          // padding too: sequences are compared with memcmp
    r.x = x; r.key = key; memcpy(r.bits, v.v, 4 * N); return r;
}

// the loop as the reference has it: first pixel, then `while (steps-- > 0) { x1++; start += dLR; plot }`
template <int N, class Frag>
void reference_walk(int W, bool single, const FPd<N>& L, const FPd<N>& R, Frag&& frag)
{
    if (single) { const int x = myfloor_x86(L.v[0]); if (x < 0 || x >= W) return; frag(x, L); return; }
    int x1 = myfloor_x86(L.v[0]); if (x1 >= W) return;
    const int x2 = myfloor_x86(R.v[0]); if (x2 < 0) return;
    int steps = abs(x2 - x1);
    if (!steps) { const int x = myfloor_x86(L.v[0]); if (x < 0 || x >= W) return; frag(x, L); return; }
    FPd<N> start = L, dLR;
    const float fs = (float)steps;
    for (int i = 0; i < N; i++) { float t = R.v[i]; t -= start.v[i]; t /= fs; dLR.v[i] = t; }
    if (x1 < 0) {
        const float k = (float)-x1;
        for (int i = 0; i < N; i++) { float t = dLR.v[i]; t *= k; start.v[i] += t; }
        steps -= (-x1);
        x1 = 0;
    }
    if (x2 >= W) steps -= (x2 - W + 1);
    frag(x1, start);
    while (steps-- > 0) { x1++; fp_add<N>(start, dLR); frag(x1, start); }
}

uint32_t rnd(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }
float frnd(uint32_t& s, float lo, float hi) { return lo + (hi - lo) * (float)(rnd(s) >> 8) * (1.0f / 16777216.0f); }

template <int N>
uint64_t run(uint32_t seed, uint32_t n_spans, int W)
{
    uint32_t s = seed * 2654435761u + 12345u + N;
    std::vector<unsigned long long> keys((size_t)W);
    for (auto& k : keys) k = ((unsigned long long)rnd(s) << 32) | rnd(s);
    uint64_t bad = 0;
    for (uint32_t i = 0; i < n_spans; i++) {
        FPd<N> L, R;
        for (int c = 0; c < N; c++) { L.v[c] = frnd(s, -3.f, 3.f); R.v[c] = frnd(s, -3.f, 3.f); }
        const uint32_t kind = rnd(s) % 8;
        float a = frnd(s, -60.f, (float)W + 60.f), b = a + frnd(s, 0.f, kind < 2 ? 2.f : (kind < 6 ? 12.f : (float)W * 1.5f));
        if (kind == 7) { a = frnd(s, -2000.f, -1.f); b = frnd(s, (float)W, (float)W + 2000.f); }    // clipped on both sides
        L.v[0] = a; R.v[0] = b;
        const bool single = (rnd(s) % 16) == 0;
        std::vector<Visit> ref, w1, w2;
        reference_walk<N>(W, single, L, R, [&](int x, const FPd<N>& v) { ref.push_back(mk<N>(x, v, keys[(size_t)x])); });
        walk_span<N>(W, single, L, R, [&](int x, const FPd<N>& v) { w1.push_back(mk<N>(x, v, keys[(size_t)x])); });
        walk_span_keyed<N, 8>(W, single, L, R, [&](int x) { return keys[(size_t)x]; },
                              [&](int x, const FPd<N>& v, unsigned long long k) { w2.push_back(mk<N>(x, v, k)); });
        auto same = [](const std::vector<Visit>& p, const std::vector<Visit>& q) {
            return p.size() == q.size() && (p.empty() || memcmp(p.data(), q.data(), p.size() * sizeof(Visit)) == 0);
        };
        if (!same(ref, w1)) bad++;
        if (!same(ref, w2)) bad++;
    }
    return bad;
}
}  // namespace

extern "C" int b200r_selftest_span_walk_host(uint32_t seed, uint32_t n_spans, uint32_t width, uint64_t* mismatches)
{
    if (!mismatches || !n_spans || width < 8 || width > 16384) {
        b200r::set_global_error("b200r_selftest_span_walk_host: bad argument");
        return B200R_EINVAL;
    }
    *mismatches = run<5>(seed, n_spans, (int)width) + run<8>(seed, n_spans, (int)width);
    return B200R_OK;
}
