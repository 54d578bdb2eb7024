// raster_steps.h — the per-span pixel walk of the scan-conversion rasteriser (reference Screen::RasterizeTriangle,
// src/Screen.h:244-289) as host+device functions: csrc/cuda/raster_kernels.cu calls them per span, and
// csrc/host/raster_steps_host.cpp runs them on the host so that the CPU suite can check the batched walker against the
// pixel-by-pixel one (same pattern as bvh_steps.h / mlaa_steps.h).
#pragma once
#include <stdint.h>

#include "vec.h"      // B2_HD

namespace b200r {

template <int N> struct FPd { float v[N]; };      // a FatPoint: v[0] = projx, then the mode's interpolants

template <int N> B2_HD void fp_add(FPd<N>& a, const FPd<N>& b)
{
#pragma unroll
    for (int i = 0; i < N; i++) a.v[i] += b.v[i];
}

// x86 cvttss2si: out of range and NaN give INT_MIN (CUDA's own conversion saturates)
B2_HD int ras_cvtt_x86(float f) { return (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000; }

// Screen::myfloor (reference src/Screen.h:218-221) with x86 float->int conversion semantics
B2_HD int myfloor_x86(float val)
{
    if (val < 0.f) return ras_cvtt_x86(val - 0.5f);
    return ras_cvtt_x86(val + 0.5f);
}

// The span's pixel range and interpolant deltas - everything of the per-scanline body of Screen::RasterizeTriangle that
// comes before its pixel loop. Returns the number of pixels (0: nothing to draw); pixel j is at x1 + j with value
// start + j additions of dLR (additions, not a product: the reference accumulates).
template <int N>
B2_HD int span_setup(int W, bool single, const FPd<N>& L, const FPd<N>& R, int& x1, FPd<N>& start, FPd<N>& dLR)
{
    start = L;
#pragma unroll
    for (int i = 0; i < N; i++) dLR.v[i] = 0.f;
    if (single) {
        x1 = myfloor_x86(L.v[0]);
        return (x1 < 0 || x1 >= W) ? 0 : 1;
    }
    x1 = myfloor_x86(L.v[0]); if (x1 >= W) return 0;
    const int x2 = myfloor_x86(R.v[0]); if (x2 < 0) return 0;
    int steps = x2 - x1; if (steps < 0) steps = -steps;
    if (!steps) return (x1 < 0 || x1 >= W) ? 0 : 1;
    const float fs = (float)steps;
#pragma unroll
    for (int i = 0; i < N; i++) { float t = R.v[i]; t -= start.v[i]; t /= fs; dLR.v[i] = t; }
    if (x1 < 0) {
        const float k = (float)-x1;
#pragma unroll
        for (int i = 0; i < N; i++) { float t = dLR.v[i]; t *= k; start.v[i] += t; }
        steps -= (-x1);
        x1 = 0;
    }
    if (x2 >= W) steps -= (x2 - W + 1);
    return 1 + (steps > 0 ? steps : 0);
}

// The pixel loop: frag(x, v) per pixel, in order.
template <int N, class Frag>
B2_HD void walk_span(int W, bool single, const FPd<N>& L, const FPd<N>& R, Frag&& frag)
{
    int x1; FPd<N> start, dLR;
    const int total = span_setup<N>(W, single, L, R, x1, start, dLR);
    for (int j = 0; j < total; j++) {
        if (j) fp_add<N>(start, dLR);
        frag(x1 + j, start);
    }
}

// The same loop for passes that need a word of per-pixel state (the stored depth key) at every pixel: key(x) for B pixels
// is requested up front, then the pixels are walked in order with their words - one memory round trip per B pixels instead
// of one per pixel (spans average ~5 pixels). frag(x, v, key(x)) sees exactly the values walk_span would produce.
template <int N, int B, class Key, class Frag>
B2_HD void walk_span_keyed(int W, bool single, const FPd<N>& L, const FPd<N>& R, Key&& key, Frag&& frag)
{
    int x1; FPd<N> start, dLR;
    const int total = span_setup<N>(W, single, L, R, x1, start, dLR);
    for (int j = 0; j < total;) {
        const int nb = (total - j) < B ? (total - j) : B;
        decltype(key(0)) kv[B];
#pragma unroll
        for (int k = 0; k < B; k++)
            if (k < nb) kv[k] = key(x1 + j + k);
#pragma unroll
        for (int k = 0; k < B; k++)
            if (k < nb) {
                if (j + k) fp_add<N>(start, dLR);
                frag(x1 + j + k, start, kv[k]);
            }
        j += nb;
    }
}

}  // namespace b200r
