// mlaa_steps.h — the per-item steps of the MLAA post filter (reference src/MLAA.cc:47-372: ssedif, mixColor,
// getSplitHeight, computeUpperBounds / computeLowerBounds, blendInterval, the one-cell blend), as host+device functions.
//
// The CUDA kernels (csrc/cuda/mlaa_kernels.cu) are made of these; csrc/host/mlaa_steps_host.cpp runs the very same functions
// in plain loops in the reference's job order, so that the CPU suite checks them bit for bit against the oracle without a
// GPU (same pattern as bvh_steps.h). Float maths as in the reference: plain products and sums (device -fmad=false, host
// -ffp-contract=off) and x86 byte truncation.
#pragma once
#include <stdint.h>

#include "vec.h"      // B2_HD

namespace b200r {

constexpr unsigned MLAA_HF = 1u << 31, MLAA_VF = 1u << 30;      // discontinuity towards the pixel below / to the right

// ssedif (MLAA.cc:47-55): some byte differs by >= 16
B2_HD bool mlaa_differs(unsigned a, unsigned b)
{
#ifdef __CUDA_ARCH__
    return (__vabsdiffu4(a, b) & 0xF0F0F0F0u) != 0;
#else
    for (int s = 0; s < 32; s += 8) {
        const int x = (int)((a >> s) & 0xffu), y = (int)((b >> s) & 0xffu);
        if ((x > y ? x - y : y - x) >= 16) return true;
    }
    return false;
#endif
}

B2_HD int mlaa_sumColor(unsigned c) { return (int)((c >> 16) & 0xff) + (int)((c >> 8) & 0xff) + (int)(c & 0xff); }

// (unsigned char)f as g++ compiles it on x86-64: cvttss2si to 32 bits (out of range / NaN -> INT_MIN), keep the low byte
B2_HD unsigned mlaa_u8(float f)
{
    const int i = (f >= -2147483648.0f && f < 2147483648.0f) ? (int)f : (int)0x80000000;
    return (unsigned)i & 0xFFu;
}

// mixColor (MLAA.cc:93-119)
B2_HD unsigned mlaa_mixColor(float w1, unsigned c1, float w2, unsigned c2)
{
    const float r1 = (float)((c1 >> 16) & 0xff), g1 = (float)((c1 >> 8) & 0xff), b1 = (float)(c1 & 0xff);
    const float r2 = (float)((c2 >> 16) & 0xff), g2 = (float)((c2 >> 8) & 0xff), b2 = (float)(c2 & 0xff);
    const unsigned r = mlaa_u8(r1 * w1 + r2 * w2), g = mlaa_u8(g1 * w1 + g2 * w2), b = mlaa_u8(b1 * w1 + b2 * w2);
    return (r << 16) | (g << 8) | b;
}

B2_HD float mlaa_getSplitHeight(const uint32_t* fb, int l, int icb, int icm, int ipb, int ipm)
{
    const int cc = mlaa_sumColor(fb[icb]), cu = mlaa_sumColor(fb[icm]), pc = mlaa_sumColor(fb[ipb]), pu = mlaa_sumColor(fb[ipm]);
    return (float)(l * (pc - cu) + (cc - cu) - (pc - pu)) / (float)(l * ((cc - cu) + (pc - pu)) + (cc - cu) - (pc - pu));
}

// computeUpperBounds / computeLowerBounds (MLAA.cc:178-280): two searches along the line [x0, x1] - forwards for the first
// usable split (s0, h0), backwards for the last one (s1, h1) - over the immutable flag copy `fb0` (sz words).
// The loops are the reference's do-while loops with their trip counts computed up front and the flag words of BATCH steps
// loaded together: a search without a hit walks the whole line four times, and in the vertical passes every step is its own
// L2 access. A speculative load is of a word the loop would read if it got that far without a hit (second words are clamped
// into the buffer: the reference reads them only behind a first-word test that fails at the frame border). Decisions are
// taken in step order from the loaded words, so results, `nsteps` and the tie-breaks are those of the step-by-step loops.
// BATCH == 1 is the loop as written in the reference.
B2_HD int mlaa_trips(int d, int stepx) { return d <= stepx ? 1 : (d + stepx - 1) / stepx; }     // do { x += stepx; } while (x < x + d)
B2_HD int mlaa_clampi(int i, int sz) { return i < 0 ? 0 : (i >= sz ? sz - 1 : i); }

template <int BATCH>
B2_HD void mlaa_computeUpperBounds(int& s0, int& s1, float& h0, float& h1, const uint32_t* fb0, unsigned fc,
                                   int x0, int x1, int len, int stepx, int befor, int after, int sz)
{
    s0 = s1 = -1;
    int nsteps = 0, xi = x0, t0 = -1, t1 = -1;
    const unsigned fo = fc ^ (MLAA_HF | MLAA_VF);
    // do { ... xi += stepx; nsteps++; } while (xi < x1);
    for (int n = mlaa_trips(x1 - xi, stepx); n > 0 && s0 == -1;) {
        const int nb = n < BATCH ? n : BATCH;
        uint32_t f[BATCH], g[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (k < nb) { f[k] = fb0[xi + k * stepx]; g[k] = fb0[mlaa_clampi(xi + k * stepx + befor, sz)]; }
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (k < nb && s0 == -1) {
                const int xk = xi + k * stepx;
                bool hit = false;
                if ((f[k] & fo) && (g[k] & fc)) {
                    h0 = mlaa_getSplitHeight(fb0, len - nsteps, xk + stepx, xk + stepx + after, xk + befor, xk);
                    if (0.f < h0 && h0 < 1.f) { s0 = xk + stepx; hit = true; }
                }
                if (!hit) {
                    if ((f[k] & fo) && t0 == -1) t0 = xk;
                    nsteps++;
                }
            }
        xi += nb * stepx; n -= nb;
    }
    if (s0 == -1 && t0 != -1) { h0 = 0.5f; s0 = t0 + stepx; }
    if (x1 + stepx >= sz) { if (fb0[x1] & fo) t1 = x1; x1 -= stepx; }
    xi = x1;
    // do { ... xi -= stepx; nsteps++; } while (xi > x0);
    for (int n = mlaa_trips(xi - x0, stepx); n > 0 && s1 == -1;) {
        const int nb = n < BATCH ? n : BATCH;
        uint32_t f[BATCH], g[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (k < nb) { f[k] = fb0[xi - k * stepx]; g[k] = fb0[mlaa_clampi(xi - k * stepx + stepx + befor, sz)]; }
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (k < nb && s1 == -1) {
                const int xk = xi - k * stepx;
                bool hit = false;
                if ((f[k] & fo) && (g[k] & fc)) {
                    h1 = mlaa_getSplitHeight(fb0, nsteps, xk + stepx, xk + stepx + befor, xk + after, xk);
                    if (0.f < h1 && h1 < 1.f) { s1 = xk; hit = true; }
                }
                if (!hit) {
                    if ((f[k] & fo) && t1 == -1) t1 = xk;
                    nsteps++;
                }
            }
        xi -= nb * stepx; n -= nb;
    }
    if (s1 == -1 && t1 != -1) { h1 = 0.5f; s1 = t1; }
}

template <int BATCH>
B2_HD void mlaa_computeLowerBounds(int& s0, int& s1, float& h0, float& h1, const uint32_t* fb0, unsigned fc,
                                   int x0, int x1, int len, int stepx, int after, int sz)
{
    s0 = s1 = -1;
    int nsteps = 0, xi = x0, t0 = -1, t1 = -1;
    const unsigned fo = fc ^ (MLAA_HF | MLAA_VF);
    for (int n = mlaa_trips(x1 - xi, stepx); n > 0 && s0 == -1;) {
        const int nb = n < BATCH ? n : BATCH;
        uint32_t f[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (k < nb) f[k] = fb0[xi + k * stepx + after];
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (k < nb && s0 == -1) {
                const int xk = xi + k * stepx, xia = xk + after;
                bool hit = false;
                if ((f[k] & fo) && (f[k] & fc)) {
                    if (xia + after < sz) h0 = mlaa_getSplitHeight(fb0, len - nsteps, xia + stepx, xk + stepx, xia + after, xia);
                    else h0 = 0.5f;
                    if (0.f < h0 && h0 < 1.f) { s0 = xk + stepx; hit = true; }
                }
                if (!hit) {
                    if ((f[k] & fo) && t0 == -1) t0 = xk;
                    nsteps++;
                }
            }
        xi += nb * stepx; n -= nb;
    }
    if (s0 == -1 && t0 != -1) { h0 = 0.5f; s0 = t0 + stepx; }
    if (x1 + stepx >= sz) { if (fb0[x1] & fo) t1 = x1; x1 -= stepx; }
    xi = x1;
    for (int n = mlaa_trips(xi - x0, stepx); n > 0 && s1 == -1;) {
        const int nb = n < BATCH ? n : BATCH;
        uint32_t f[BATCH], g[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (k < nb) { f[k] = fb0[xi - k * stepx + after]; g[k] = fb0[mlaa_clampi(xi - k * stepx + after + stepx, sz)]; }
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (k < nb && s1 == -1) {
                const int xk = xi - k * stepx, xia = xk + after;
                bool hit = false;
                if ((f[k] & fo) && (g[k] & fo)) {
                    if (xia + after < sz) h1 = mlaa_getSplitHeight(fb0, nsteps, xia + stepx, xia + after + stepx, xk, xia);
                    else h1 = 0.5f;
                    if (0.f < h1 && h1 < 1.f) { s1 = xk; hit = true; }
                }
                if (!hit) {
                    if ((f[k] & fo) && t1 == -1) t1 = xk;
                    nsteps++;
                }
            }
        xi -= nb * stepx; n -= nb;
    }
    if (s1 == -1 && t1 != -1) { h1 = 0.5f; s1 = t1; }
}

// One run of blendInterval's two pixel loops (MLAA.cc:311-316, 340-345):
//     do { fbi[x + wshift] = mixColor(area, fbi[x], 1 - area, fbi[x + other]); area += dh; x += stepx; } while (...)
// for `n` pixels. The pixels of a run are independent of each other - iteration k reads x+k*stepx and x+k*stepx+other and
// writes one of the two; another iteration's addresses differ from them by a non-zero multiple of stepx, possibly +-other,
// which is never 0 while a run is shorter than a row (horizontal: stepx 1, other +-resX; vertical: stepx resX, other +-1) -
// so BATCH pixels are loaded together, then blended and stored in order: one memory round trip per BATCH pixels instead of
// one per pixel (the vertical passes walk with a stride of a whole row: every pixel is its own L2 access). The areas are the
// same sequence of float additions. BATCH == 1 is the loop as written in the reference.
template <int BATCH>
B2_HD void mlaa_blendRun(uint32_t* fbi, int& x, int n, float& area, float dh, int stepx, int other, int wshift)
{
    while (n > 0) {
        const int nb = n < BATCH ? n : BATCH;
        uint32_t a[BATCH], b[BATCH];
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (k < nb) { a[k] = fbi[x + k * stepx]; b[k] = fbi[x + k * stepx + other]; }
#pragma unroll
        for (int k = 0; k < BATCH; k++)
            if (k < nb) {
                fbi[x + k * stepx + wshift] = mlaa_mixColor(area, a[k], 1.f - area, b[k]);
                area += dh;
            }
        x += nb * stepx;
        n -= nb;
    }
}

// blendInterval (MLAA.cc:282-347). The two do-while loops run max(1, .) times: their trip counts are computed up front.
template <int BATCH>
B2_HD void mlaa_blendInterval(uint32_t* fbi, int x0, int x1, float h0, float h1, int stepx, int other, bool ushape)
{
    float dh0 = ((2.f * (1.f - h0)) * (float)stepx) / (float)(x1 - x0 + stepx);
    float dh1 = ((2.f * (1.f - h1)) * (float)stepx) / (float)(x1 - x0 + stepx);
    int shift = other < 0 ? -other : 0;
    x0 += shift; x1 += shift;
    const int middle = (x0 + x1) / 2;
    float area = h0 + 0.5f * dh0;
    if (h0 == 0.f) {
        x0 += 1 + (x1 - x0) / stepx;
        area = dh1;
    } else {
        // do { ...; x0 += stepx; } while (x0 < middle);
        const int d = middle - x0;
        const int n1 = mlaa_trips(d, stepx);
        mlaa_blendRun<BATCH>(fbi, x0, n1, area, dh0, stepx, other, 0);
        if (x0 == middle) {
            fbi[x0] = mlaa_mixColor((1.f - dh0 / 8.f), fbi[x0], dh0 / 8.f, fbi[x0 + other]);
            if (!ushape) fbi[x0 + other] = mlaa_mixColor(dh1 / 8.f, fbi[x0], (1.f - dh1 / 8.f), fbi[x0 + other]);
            x0 += stepx;
            area = dh1;
        } else {
            area = 0.5f * dh1;
        }
    }
    if (h1 == 0.f) return;
    if (ushape) { area = 1.f - area; dh1 = -dh1; }
    shift = ushape ? 0 : other;
    // do { ...; x0 += stepx; } while (x0 <= x1);
    const int d2 = x1 - x0;
    const int n2 = d2 < 0 ? 1 : d2 / stepx + 1;
    mlaa_blendRun<BATCH>(fbi, x0, n2, area, dh1, stepx, other, shift);
}

B2_HD void mlaa_blend_one_cell(uint32_t* fbi, int x0, int after)
{
    const float weightc = 7.0f / 8;
    fbi[x0] = mlaa_mixColor(weightc, fbi[x0], 1.f - weightc, fbi[x0 + after]);
    fbi[x0 + after] = mlaa_mixColor(1.f - weightc, fbi[x0], weightc, fbi[x0 + after]);
}

// Everything blending needs to know about one separation line; ui1 == -2: a one-pixel line at ui0.
struct __attribute__((aligned(16))) MlaaLineRec { int ui0, ui1, li0, li1; float uh0, uh1, lh0, lh1; };

// The body of the while loop at MLAA.cc:565-699 up to the blends: end points and split heights of the line [x0, x1] (len pixels)
// of row/column yc. Reads the immutable flag copy only.
template <int BATCH>
B2_HD MlaaLineRec mlaa_line_bounds(const uint32_t* fb0, unsigned fc, int yc, int x0, int x1, int len, int stepx, int befor,
                                   int after, int sz)
{
    MlaaLineRec r; r.ui0 = x0; r.ui1 = -2; r.li0 = r.li1 = -1; r.uh0 = r.uh1 = r.lh0 = r.lh1 = 0.f;
    if (len != 1) {
        if (x0 == yc) { x0 += stepx; len--; }
        mlaa_computeUpperBounds<BATCH>(r.ui0, r.ui1, r.uh0, r.uh1, fb0, fc, x0 - stepx, x1, len, stepx, befor, after, sz);
        mlaa_computeLowerBounds<BATCH>(r.li0, r.li1, r.lh0, r.lh1, fb0, fc, x0 - stepx, x1, len, stepx, after, sz);
    }
    return r;
}

// ... and its blends (in place on the frame)
template <int BATCH>
B2_HD void mlaa_line_blend(uint32_t* fbi, const MlaaLineRec& r, int stepx, int befor, int after)
{
    if (r.ui1 == -2) { mlaa_blend_one_cell(fbi, r.ui0, after); return; }
    bool done = false;
    if (r.ui0 != -1 && r.li1 != -1 && r.ui0 < r.li1) { mlaa_blendInterval<BATCH>(fbi, r.ui0, r.li1, r.uh0, r.lh1, stepx, after, false); done = true; }
    if (r.li0 != -1 && r.ui1 != -1 && r.li0 < r.ui1) { mlaa_blendInterval<BATCH>(fbi, r.li0, r.ui1, r.lh0, r.uh1, stepx, befor, false); done = true; }
    if (!done) {
        if (r.ui0 != -1 && r.ui1 != -1 && r.ui0 < r.ui1) mlaa_blendInterval<BATCH>(fbi, r.ui0, r.ui1, r.uh0, r.uh1, stepx, after, true);
        if (r.li0 != -1 && r.li1 != -1 && r.li0 < r.li1) mlaa_blendInterval<BATCH>(fbi, r.li0, r.li1, r.lh0, r.lh1, stepx, befor, true);
    }
}

}  // namespace b200r
