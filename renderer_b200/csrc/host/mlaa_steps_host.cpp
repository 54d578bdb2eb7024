// mlaa_steps_host.cpp — the MLAA step functions (csrc/mlaa_steps.h) run in plain loops on the host.
//
// This is NOT a rendering path of the product (frames are filtered by csrc/cuda/mlaa_kernels.cu on the device): it exists so
// that the functions those kernels are made of - flags, line bounds, split heights, the batched in-place blends - are
// exercised by the CPU test suite against the oracle, in the job order the kernels keep: horizontal lines in 8-row blocks
// (even blocks, then odd blocks, rows of a block in order), then vertical lines in 8-column blocks (reference
// src/MLAA.cc:524-704).
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/b200render.h"
#include "../mlaa_steps.h"

namespace b200r { void set_global_error(const std::string& s); }

namespace {
using namespace b200r;

template <int BATCH>
void filter(uint32_t* fbi, int resX, int resY)
{
    const int sz = resX * resY;
    std::vector<uint32_t> flags((size_t)sz);
    uint32_t* fb0 = flags.data();
    // find fragments (MLAA.cc:437-503; mlaa_find_fragments_kernel)
    for (int y = 0; y < resY; y++)
        for (int x = 0; x < resX; x++) {
            const int ci = y * resX + x;
            const unsigned c = fbi[ci];
            const unsigned below = (y == resY - 1) ? c : fbi[ci + resX];
            const unsigned right = (x == resX - 1) ? c : fbi[ci + 1];
            fb0[ci] = c | (mlaa_differs(c, below) ? MLAA_HF : 0u) | (mlaa_differs(c, right) ? MLAA_VF : 0u);
        }
    for (int vertical = 0; vertical < 2; vertical++) {
        unsigned fc; int resx, resy, stepy, stepx;
        if (!vertical) { fc = MLAA_HF; resx = resX; resy = resY; stepy = resX; stepx = 1; }
        else { fc = MLAA_VF; resx = resY; resy = resX; stepy = 1; stepx = resX; }
        const int jobs = resy / 8 + ((resy % 8) ? 1 : 0);
        const int after = stepy;
        for (int yodd = 0; yodd < 2; yodd++) {
            const int count = yodd ? jobs - jobs / 2 : jobs / 2;        // MLAA.cc:545-552: the first half of the job list is the even blocks
            for (int job = 0; job < count; job++) {
                const int rfrst = (2 * job + yodd) * 8;
                int rlast = rfrst + 8;
                if (rlast >= resy) rlast = resy - 1;                    // the last row / column is never a block row (MLAA.cc:556-557)
                for (int row = rfrst; row < rlast; row++) {
                    const int yc = row * stepy;
                    const int befor = row ? -stepy : 0;
                    int lastEnd = -1;
                    for (int k = 0; k < resx; k++) {
                        const int x = yc + k * stepx;
                        if (!(fb0[x] & fc)) continue;
                        if (k > 0 && (fb0[x - stepx] & fc)) continue;   // not the first pixel of its run
                        int k1 = k;
                        while (k1 + 1 < resx && (fb0[yc + (k1 + 1) * stepx] & fc)) k1++;
                        if (k1 > lastEnd) lastEnd = k1;
                        const MlaaLineRec r = mlaa_line_bounds<BATCH>(fb0, fc, yc, x, yc + k1 * stepx, k1 - k + 1, stepx, befor, after, sz);
                        mlaa_line_blend<BATCH>(fbi, r, stepx, befor, after);
                    }
                    // the SSE scan quirk of the horizontal search (mlaa_blend_lines_kernel; oracle/port/mlaa_port.cpp explains it)
                    if (!vertical && lastEnd >= 0 && (lastEnd == resx - 4 || lastEnd == resx - 3)) {
                        const int base = yc + resx;
                        for (int q = 0; q < 4; q++)
                            if (fb0[base + q] & MLAA_HF) {
                                if (base + q + after < sz) mlaa_blend_one_cell(fbi, base + q, after);
                                break;
                            }
                    }
                }
            }
        }
    }
}
}  // namespace

extern "C" int b200r_selftest_mlaa_steps_host(uint32_t* frame_xrgb, uint32_t width, uint32_t height, int batched)
{
    if (!frame_xrgb || !width || !height || (width % 4) || (height % 8) || width > 16384 || height > 16384) {
        b200r::set_global_error("b200r_selftest_mlaa_steps_host: bad argument (needs width % 4 == 0 and height % 8 == 0)");
        return B200R_EINVAL;
    }
    if (batched) filter<8>(frame_xrgb, (int)width, (int)height);
    else filter<1>(frame_xrgb, (int)width, (int)height);
    return B200R_OK;
}
