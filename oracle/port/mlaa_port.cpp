// mlaa_port.cpp — CPU restatement (scalar, no SSE) of the reference's MLAA post filter. TEST INFRASTRUCTURE ONLY.
//
// Follows MLAA(fbi, NULL, resX, resY, job=0, threadID=0) as Screen::ShowScreen calls it (reference
// src/Screen.h:133-134), i.e. ONE thread running every job in order: all "find fragments" blocks, then the
// horizontal-line blocks (even blocks first, then odd), then the vertical-line blocks (even, odd).
//   ssedif / mixColor / sumColor              reference src/MLAA.cc:47-119
//   findSeparationLine                         reference src/MLAA.cc:122-176
//   getSplitHeight                             reference src/MLAA.cc:179-186
//   computeUpperBounds / computeLowerBounds    reference src/MLAA.cc:205-321
//   blendInterval                              reference src/MLAA.cc:324-371
//   MLAA (job loop, find fragments, blending)  reference src/MLAA.cc:374-714
// Requires resX % 4 == 0 and resY % 8 == 0 (the reference reads/writes out of bounds or leaves its scratch
// uninitialised otherwise, :453-457 and :396).
#include <cstdlib>
#include <cstring>
#include <vector>

#include "oracle_port.h"

namespace {

const unsigned HF = 1u << 31, VF = 1u << 30;

// ssedif (:47-55): any of the 4 bytes differs by >= 16  (|a-b| & 0xF0 != 0)
inline bool differs(unsigned a, unsigned b)
{
    for (int s = 0; s < 32; s += 8) {
        int x = (a >> s) & 0xff, y = (b >> s) & 0xff;
        int d = x > y ? x - y : y - x;
        if (d & 0xf0) return true;
    }
    return false;
}

inline int sumColor(unsigned c1) { return ((c1 >> 16) & 0xff) + ((c1 >> 8) & 0xff) + (c1 & 0xff); }

inline unsigned mixColor(float w1, unsigned c1, float w2, unsigned c2)
{
    unsigned char r1 = (c1 >> 16) & 0xff, g1 = (c1 >> 8) & 0xff, b1 = c1 & 0xff;
    unsigned char r2 = (c2 >> 16) & 0xff, g2 = (c2 >> 8) & 0xff, b2 = c2 & 0xff;
    r1 = (unsigned char)(r1 * w1 + r2 * w2);
    g1 = (unsigned char)(g1 * w1 + g2 * w2);
    b1 = (unsigned char)(b1 * w1 + b2 * w2);
    return (r1 << 16) | (g1 << 8) | b1;
}

// findSeparationLine (:122-176). The stepx==1 branch scans with SSE: first pixel by pixel up to a multiple of 4,
// then 4 pixels per step, checking `xstart >= xend` only AFTER an empty group. Restated literally (movemask of four
// sign bits == H flags of four pixels), because that control flow has an observable quirk: when the search starts
// 1-3 pixels before the end of a row and finds nothing there, the first group of the NEXT row is examined too and
// a flagged pixel in it is returned as a one-pixel line.
inline int findSeparationLine(int& x0, int& x1, const unsigned* fb0, unsigned fc, int xstart, int xend, int stepx)
{
    if (xstart >= xend) return 0;
    x0 = -1;
    if (stepx > 1) {
        while (true) {
            if (fb0[xstart] & fc) { x0 = xstart; break; }
            xstart += stepx;
            if (xstart > xend) return 0;
        }
    } else {
        bool found = false;
        while (xstart & 3) {
            if (fb0[xstart] & HF) { x0 = xstart; found = true; break; }
            xstart++;
        }
        while (!found) {
            int f = 0;
            for (int k = 0; k < 4; k++) if (fb0[xstart + k] & HF) f |= 1 << k;
            if (f) {
                xstart += (f & 1) ? 0 : (f & 2) ? 1 : (f & 4) ? 2 : 3;
                x0 = xstart;
                break;
            }
            xstart += 4;
            if (xstart >= xend) return 0;
        }
    }
    int len = 1;
    xstart += stepx;
    while (xstart <= xend && (fb0[xstart] & fc)) { len++; xstart += stepx; }
    x1 = xstart - stepx;
    return len;
}

inline float getSplitHeight(const unsigned* fb, int l, int icb, int icm, int ipb, int ipm)
{
    int cc = sumColor(fb[icb]), cu = sumColor(fb[icm]), pc = sumColor(fb[ipb]), pu = sumColor(fb[ipm]);
    return float(l * (pc - cu) + (cc - cu) - (pc - pu)) / (l * ((cc - cu) + (pc - pu)) + (cc - cu) - (pc - pu));
}

void computeUpperBounds(int& s0, int& s1, float& h0, float& h1, const unsigned* fb0, unsigned fc, int x0, int x1, int len,
                        int stepx, int befor, int after, int sz)
{
    s0 = s1 = -1;
    int nsteps = 0, xi = x0, t0 = -1, t1 = -1;
    unsigned fo = fc ^ (HF | VF);
    do {
        if ((fb0[xi] & fo) && (fb0[xi + befor] & fc)) {
            h0 = getSplitHeight(fb0, len - nsteps, xi + stepx, xi + stepx + after, xi + befor, xi);
            if (0 < h0 && h0 < 1) { s0 = xi + stepx; break; }
        }
        if ((fb0[xi] & fo) && t0 == -1) t0 = xi;
        xi += stepx;
        nsteps++;
    } while (xi < x1);
    if (s0 == -1 && t0 != -1) { h0 = 0.5f; s0 = t0 + stepx; }
    if (x1 + stepx >= sz) { if (fb0[x1] & fo) t1 = x1; x1 -= stepx; }
    xi = x1;
    do {
        if ((fb0[xi] & fo) && (fb0[xi + stepx + befor] & fc)) {
            h1 = getSplitHeight(fb0, nsteps, xi + stepx, xi + stepx + befor, xi + after, xi);
            if (0 < h1 && h1 < 1) { s1 = xi; break; }
        }
        if ((fb0[xi] & fo) && t1 == -1) t1 = xi;
        xi -= stepx;
        nsteps++;
    } while (xi > x0);
    if (s1 == -1 && t1 != -1) { h1 = 0.5f; s1 = t1; }
}

void computeLowerBounds(int& s0, int& s1, float& h0, float& h1, const unsigned* fb0, unsigned fc, int x0, int x1, int len,
                        int stepx, int after, int sz)
{
    s0 = s1 = -1;
    int nsteps = 0, xi = x0, t0 = -1, t1 = -1;
    unsigned fo = fc ^ (HF | VF);
    do {
        int xia = xi + after;
        if ((fb0[xia] & fo) && (fb0[xia] & fc)) {
            if (xia + after < sz) h0 = getSplitHeight(fb0, len - nsteps, xia + stepx, xi + stepx, xia + after, xia);
            else h0 = 0.5f;
            if (0 < h0 && h0 < 1) { s0 = xi + stepx; break; }
        }
        if ((fb0[xia] & fo) && t0 == -1) t0 = xi;
        xi += stepx;
        nsteps++;
    } while (xi < x1);
    if (s0 == -1 && t0 != -1) { h0 = 0.5f; s0 = t0 + stepx; }
    if (x1 + stepx >= sz) { if (fb0[x1] & fo) t1 = x1; x1 -= stepx; }
    xi = x1;
    do {
        int xia = xi + after;
        if ((fb0[xia] & fo) && (fb0[xia + stepx] & fo)) {
            if (xia + after < sz) h1 = getSplitHeight(fb0, nsteps, xia + stepx, xia + after + stepx, xi, xia);
            else h1 = 0.5f;
            if (0 < h1 && h1 < 1) { s1 = xi; break; }
        }
        if ((fb0[xia] & fo) && t1 == -1) t1 = xi;
        xi -= stepx;
        nsteps++;
    } while (xi > x0);
    if (s1 == -1 && t1 != -1) { h1 = 0.5f; s1 = t1; }
}

void blendInterval(unsigned* fbi, int x0, int x1, float h0, float h1, int stepx, int other, bool ushape)
{
    float dh0 = 2 * (1 - h0) * stepx / (x1 - x0 + stepx);
    float dh1 = 2 * (1 - h1) * stepx / (x1 - x0 + stepx);
    int shift = other < 0 ? -other : 0;
    x0 += shift; x1 += shift;
    int middle = (x0 + x1) / 2;
    float area = h0 + 0.5f * dh0;
    if (h0 == 0) {
        x0 += 1 + (x1 - x0) / stepx;
        area = dh1;
    } else {
        do {
            fbi[x0] = mixColor(area, fbi[x0], 1 - area, fbi[x0 + other]);
            area += dh0;
            x0 += stepx;
        } while (x0 < middle);
        if (x0 == middle) {
            fbi[x0] = mixColor((1 - dh0 / 8), fbi[x0], dh0 / 8, fbi[x0 + other]);
            if (!ushape) fbi[x0 + other] = mixColor(dh1 / 8, fbi[x0], (1 - dh1 / 8), fbi[x0 + other]);
            x0 += stepx;
            area = dh1;
        } else {
            area = 0.5f * dh1;
        }
    }
    if (h1 == 0) return;
    if (ushape) { area = 1 - area; dh1 = -dh1; }
    shift = ushape ? 0 : other;
    do {
        fbi[x0 + shift] = mixColor(area, fbi[x0], 1 - area, fbi[x0 + other]);
        area += dh1;
        x0 += stepx;
    } while (x0 <= x1);
}

}  // namespace

extern "C" int oracle_mlaa(uint32_t* fbi, int resX, int resY)
{
    if (!fbi || resX <= 0 || resY <= 0 || (resX % 4) || (resY % 8)) return -1;
    std::vector<unsigned> scratch((size_t)resX * resY);
    unsigned* fb0 = scratch.data();
    const int rows_per_job = 8;
    const int n_find_fragment_jobs = resY / rows_per_job;
    const int n_hscan_jobs = (resY / rows_per_job) + ((resY % rows_per_job) ? 1 : 0);
    const int n_vscan_jobs = (resX / rows_per_job) + ((resX % rows_per_job) ? 1 : 0);
    const int njobs = n_find_fragment_jobs + n_hscan_jobs + n_vscan_jobs;

    for (int job = 0; job < njobs; job++) {
        int jobindex = job;
        if (jobindex < n_find_fragment_jobs) {
            // find fragments (:437-503): H flag = differs from the pixel below, V flag = differs from the pixel to the
            // right; the last row / last column compare with themselves (no flag).
            const int yfrst = jobindex * rows_per_job, ylast = yfrst + rows_per_job;
            for (int y = yfrst; y < ylast; y++)
                for (int x = 0; x < resX; x++) {
                    const int ci = y * resX + x;
                    const unsigned c = fbi[ci];
                    const unsigned below = (y == resY - 1) ? c : fbi[ci + resX];
                    const unsigned right = (x == resX - 1) ? c : fbi[ci + 1];
                    fb0[ci] = c | (differs(c, below) ? HF : 0) | (differs(c, right) ? VF : 0);
                }
            continue;
        }
        // main blending loop (:524-704)
        jobindex -= n_find_fragment_jobs;
        unsigned fc; int resx, resy, stepy, stepx, scanjobs;
        if (jobindex < n_hscan_jobs) { fc = HF; resx = resX; resy = resY; stepy = resX; stepx = 1; scanjobs = n_hscan_jobs; }
        else { jobindex -= n_hscan_jobs; fc = VF; resx = resY; resy = resX; stepy = 1; stepx = resX; scanjobs = n_vscan_jobs; }
        int yodd;
        if (jobindex >= scanjobs / 2) { jobindex -= scanjobs / 2; yodd = 1; } else yodd = 0;
        int yfrst = (2 * jobindex + yodd) * rows_per_job * stepy;
        int ylast = yfrst + rows_per_job * stepy;
        if (ylast >= resy * stepy) ylast = resy * stepy - stepy;
        int befor = yfrst ? -stepy : 0;
        const int after = stepy;
        const int sz = resX * resY;
        for (int yc = yfrst; yc < ylast; yc += stepy, befor = -stepy) {
            int x0, x1, len;
            const int xend = yc + (resx - 1) * stepx;
            int xstart = yc;
            while ((len = findSeparationLine(x0, x1, fb0, fc, xstart, xend, stepx))) {
                if (len == 1) {
                    const float weightc = 7.0f / 8;
                    if (x0 + after >= sz) { xstart = x1 + stepx; continue; }   // only reachable through the next-row quirk on the
                                                                               // last processed row, where the reference writes out of bounds
                    fbi[x0] = mixColor(weightc, fbi[x0], 1 - weightc, fbi[x0 + after]);
                    fbi[x0 + after] = mixColor(1 - weightc, fbi[x0], weightc, fbi[x0 + after]);
                } else {
                    if (x0 == yc) { x0 += stepx; len--; }
                    int ui0, ui1, li0, li1; float uh0, uh1, lh0, lh1;
                    computeUpperBounds(ui0, ui1, uh0, uh1, fb0, fc, x0 - stepx, x1, len, stepx, befor, after, sz);
                    computeLowerBounds(li0, li1, lh0, lh1, fb0, fc, x0 - stepx, x1, len, stepx, after, sz);
                    bool done = false;
                    if (ui0 != -1 && li1 != -1 && ui0 < li1) { blendInterval(fbi, ui0, li1, uh0, lh1, stepx, after, false); done = true; }
                    if (li0 != -1 && ui1 != -1 && li0 < ui1) { blendInterval(fbi, li0, ui1, lh0, uh1, stepx, befor, false); done = true; }
                    if (!done) {
                        if (ui0 != -1 && ui1 != -1 && ui0 < ui1) blendInterval(fbi, ui0, ui1, uh0, uh1, stepx, after, true);
                        if (li0 != -1 && li1 != -1 && li0 < li1) blendInterval(fbi, li0, li1, lh0, lh1, stepx, befor, true);
                    }
                }
                xstart = x1 + stepx;
            }
        }
    }
    return 0;
}
