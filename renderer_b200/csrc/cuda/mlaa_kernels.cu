// mlaa_kernels.cu — the reference's morphological anti-aliasing post filter (Intel MLAA 2009) on the device.
//
// Replaces MLAA(fbi, NULL, resX, resY) as called single-threaded from Screen::ShowScreen
// (reference src/Screen.h:133-134, src/MLAA.cc:374-714):
//   pass "find fragments" (MLAA.cc:437-503): H/V discontinuity flags into bits 31/30 of a scratch copy
//   blending (MLAA.cc:524-704): horizontal separation lines in 8-row blocks (even blocks, then odd blocks), then
//   vertical lines in 8-column blocks (even, odd); blending is IN PLACE on the frame, rows inside a block in order.
// What is parallel here and why it is still bit-identical to the serial job order of the reference:
//   * blocks of the same parity touch disjoint rows (block b reads/writes rows 8b..8b+8)        -> one CTA per block
//   * inside a block the 8 rows stay sequential (row y blends into row y+1)                      -> __syncthreads between rows
//   * separation lines of one row cover disjoint pixel ranges and only read/write inside them    -> one thread per line
//   * all decisions (flags, split heights) read the immutable scratch copy only.
// Float maths as in the reference (mixColor's float products + x86 byte truncation), -fmad=false.
#include <cuda.h>          // CUtensorMap (types only: the driver entry point is looked up at run time)

#include "device_types.cuh"
#include "rt_kernels.cuh"
#include "../mlaa_steps.h"

namespace b200r {
namespace {

constexpr unsigned HF = MLAA_HF, VF = MLAA_VF;
constexpr int BLEND_BATCH = 8;            // pixels of a blend run loaded together (mlaa_steps.h, mlaa_blendRun)

__device__ __forceinline__ bool differs(unsigned a, unsigned b) { return mlaa_differs(a, b); }

__global__ void mlaa_find_fragments_kernel(const uint32_t* __restrict__ fbi, uint32_t* __restrict__ fb0, int resX, int resY)
{
    // four pixels per thread (resX % 4 == 0 is a precondition of the filter, MLAA.cc:453-457): 16-byte loads/stores, and
    // no 64-bit division per pixel
    const int qx = resX >> 2;
    const int nq = qx * resY;
    for (int q = (int)(blockIdx.x * blockDim.x + threadIdx.x); q < nq; q += (int)(gridDim.x * blockDim.x)) {
        const int y = q / qx, x = (q - y * qx) << 2;
        const size_t ci = (size_t)y * resX + x;
        const uint4 c = *reinterpret_cast<const uint4*>(fbi + ci);
        const uint4 b = (y == resY - 1) ? c : *reinterpret_cast<const uint4*>(fbi + ci + resX);
        const unsigned r3 = (x + 3 == resX - 1) ? c.w : fbi[ci + 4];
        uint4 o;
        o.x = c.x | (differs(c.x, b.x) ? HF : 0u) | (differs(c.x, c.y) ? VF : 0u);
        o.y = c.y | (differs(c.y, b.y) ? HF : 0u) | (differs(c.y, c.z) ? VF : 0u);
        o.z = c.z | (differs(c.z, b.z) ? HF : 0u) | (differs(c.z, c.w) ? VF : 0u);
        o.w = c.w | (differs(c.w, b.w) ? HF : 0u) | (differs(c.w, r3) ? VF : 0u);
        *reinterpret_cast<uint4*>(fb0 + ci) = o;
    }
}

// One separation line [x0, x1] of row/column yc (the body of the while loop at MLAA.cc:565-699)
template <int BATCH>
__device__ void process_line(uint32_t* fbi, const uint32_t* __restrict__ fb0, unsigned fc, int yc, int x0, int x1, int len,
                             int stepx, int befor, int after, int sz)
{
    const MlaaLineRec r = mlaa_line_bounds<BATCH>(fb0, fc, yc, x0, x1, len, stepx, befor, after, sz);
    mlaa_line_blend<BATCH>(fbi, r, stepx, befor, after);
}

// One blending job = one 8-row (vertical == 0) or 8-column (vertical == 1) block; one CTA per job.
// `yodd` selects the parity half of the reference's job list (MLAA.cc:545-552).
template <int BATCH>
__global__ void __launch_bounds__(256)
mlaa_blend_kernel(uint32_t* fbi, const uint32_t* __restrict__ fb0, int resX, int resY, int vertical, int yodd)
{
    const int rows_per_job = 8;
    unsigned fc; int resx, resy, stepy, stepx;
    if (!vertical) { fc = HF; resx = resX; resy = resY; stepy = resX; stepx = 1; }
    else { fc = VF; resx = resY; resy = resX; stepy = 1; stepx = resX; }
    const int jobindex = (int)blockIdx.x;
    int yfrst = (2 * jobindex + yodd) * rows_per_job * stepy;
    int ylast = yfrst + rows_per_job * stepy;
    if (ylast >= resy * stepy) ylast = resy * stepy - stepy;
    int befor = yfrst ? -stepy : 0;
    const int after = stepy;
    const int sz = resX * resY;
    __shared__ int s_lastEnd;      // x1 (as a pixel index k along the row) of the right-most line of the current row

    for (int yc = yfrst; yc < ylast; yc += stepy, befor = -stepy) {
        if (threadIdx.x == 0) s_lastEnd = -1;
        __syncthreads();
        // every maximal run of flagged pixels in [yc, xend] is one separation line (findSeparationLine, :122-176)
        for (int k = (int)threadIdx.x; k < resx; k += (int)blockDim.x) {
            const int x = yc + k * stepx;
            if (!(fb0[x] & fc)) continue;
            if (k > 0 && (fb0[x - stepx] & fc)) continue;          // not the first pixel of its run
            int k1 = k;
            while (k1 + 1 < resx && (fb0[yc + (k1 + 1) * stepx] & fc)) k1++;
            atomicMax(&s_lastEnd, k1);
            process_line<BATCH>(fbi, fb0, fc, yc, x, yc + k1 * stepx, k1 - k + 1, stepx, befor, after, sz);
        }
        __syncthreads();
        // The SSE scan quirk of the horizontal search (see oracle/port/mlaa_port.cpp): when the last line of the row
        // ends 2 or 3 pixels before the end of the row, the search for the next line runs into the first four pixels
        // of the NEXT row and returns a flagged one as a one-pixel line. Applied after all lines of this row.
        if (!vertical && threadIdx.x == 0) {
            const int kEnd = s_lastEnd;
            if (kEnd >= 0 && (kEnd == resx - 4 || kEnd == resx - 3)) {
                const int base = yc + resx;                       // first pixel of the next row
                for (int q = 0; q < 4; q++)
                    if (fb0[base + q] & HF) {
                        if (base + q + after < sz) mlaa_blend_one_cell(fbi, base + q, after);   // (the reference would write out of bounds)
                        break;
                    }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Two-stage blending.  Everything process_line() decides - where the lines of a row are, their end points and split
// heights - reads the immutable scratch copy only, so it does not depend on the blending order: mlaa_lines_kernel finds
// ALL separation lines of the frame (both directions) at once, one thread per line, and stores a 32-byte record per
// line in its row's list. The ordered part (8-row blocks, rows in order, even blocks before odd ones) is then left with
// the in-place blends of the rows' records: no scan over the 98 % of pixels that carry no flag, and no serial walks
// along the lines inside the ordered loop (4K frame: 4 x mlaa_blend_kernel = 1.07 ms of a 1.19 ms frame before).
// ---------------------------------------------------------------------------------------------------------
using LineRec = MlaaLineRec;

// The first pixel of a separation line (gi < sz: horizontal line at pixel gi; else vertical at pixel gi - sz): find its end,
// its bounds and split heights, append the record to its row's (column's) list.
template <int BATCH>
__device__ __forceinline__ void line_record(const uint32_t* __restrict__ fb0, int resX, int resY, int gi, LineRec* __restrict__ recH,
                                            LineRec* __restrict__ recV, int* __restrict__ cntH, int* __restrict__ cntV,
                                            int* __restrict__ endH, int* __restrict__ endV, int capH, int capV)
{
    const int sz = resX * resY;
    const int vertical = gi >= sz;
    const int x = vertical ? gi - sz : gi;
    unsigned fc; int resx, stepy, stepx, row, k;
    if (!vertical) { fc = HF; resx = resX; stepy = resX; stepx = 1; row = x / resX; k = x - row * resX; }
    else { fc = VF; resx = resY; stepy = 1; stepx = resX; k = x / resX; row = x - k * resX; }
    const int yc = row * stepy;
    int k1 = k;
    while (k1 + 1 < resx && (fb0[yc + (k1 + 1) * stepx] & fc)) k1++;
    atomicMax(vertical ? &endV[row] : &endH[row], k1);
    const int befor = row ? -stepy : 0, after = stepy;
    const LineRec r = mlaa_line_bounds<BATCH>(fb0, fc, yc, x, yc + k1 * stepx, k1 - k + 1, stepx, befor, after, sz);
    const int slot = atomicAdd(vertical ? &cntV[row] : &cntH[row], 1);
    if (slot < (vertical ? capV : capH)) (vertical ? recV + (size_t)row * capV : recH + (size_t)row * capH)[slot] = r;
}

// is pixel item gi (see line_record) the first pixel of a separation line of a block row?
__device__ __forceinline__ bool is_line_start(const uint32_t* __restrict__ fb0, int resX, int resY, int gi)
{
    const int sz = resX * resY;
    const int vertical = gi >= sz;
    const int x = vertical ? gi - sz : gi;
    const unsigned fc = vertical ? VF : HF;
    if (!(fb0[x] & fc)) return false;                                   // 98 % of the pixels end here: no division yet
    int row, k, stepx, resy;
    if (!vertical) { row = x / resX; k = x - row * resX; stepx = 1; resy = resY; }
    else { k = x / resX; row = x - k * resX; stepx = resX; resy = resX; }
    if (row >= resy - 1) return false;                                  // the last row / column is never a block row (MLAA.cc:556-557)
    if (k > 0 && (fb0[x - stepx] & fc)) return false;                   // not the first pixel of its run
    return true;
}

// one pass over both orientations; the thread that finds a line start also walks it (B200R_MLAA_FULLSCAN: kept for A/B)
__global__ void __launch_bounds__(256)
mlaa_lines_kernel(const uint32_t* __restrict__ fb0, int resX, int resY, LineRec* __restrict__ recH, LineRec* __restrict__ recV,
                  int* __restrict__ cntH, int* __restrict__ cntV, int* __restrict__ endH, int* __restrict__ endV, int capH, int capV)
{
    const int total = 2 * resX * resY;
    for (int gi = (int)(blockIdx.x * blockDim.x + threadIdx.x); gi < total; gi += (int)(gridDim.x * blockDim.x))
        if (is_line_start(fb0, resX, resY, gi)) line_record<1>(fb0, resX, resY, gi, recH, recV, cntH, cntV, endH, endV, capH, capV);
}

// default: the scan only LISTS the line starts (warp-aggregated append) ...
__global__ void __launch_bounds__(256)
mlaa_line_starts_kernel(const uint32_t* __restrict__ fb0, int resX, int resY, int* __restrict__ list, int* __restrict__ listCount)
{
    const int total = 2 * resX * resY;
    const unsigned lane = threadIdx.x & 31u;
    const int stride = (int)(gridDim.x * blockDim.x);
    for (int g0 = (int)(blockIdx.x * blockDim.x + threadIdx.x) - (int)lane; g0 < total; g0 += stride) {    // warp-uniform trip count
        const int gi = g0 + (int)lane;
        const bool start = gi < total && is_line_start(fb0, resX, resY, gi);
        const unsigned m = __ballot_sync(0xffffffffu, start);
        if (!m) continue;
        int base = 0;
        if (lane == (unsigned)(__ffs(m) - 1)) base = atomicAdd(listCount, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (start) list[base + __popc(m & ((1u << lane) - 1u))] = gi;
    }
}
// ... and the walks along the lines run one line per thread, every lane busy
template <int BATCH>
__global__ void __launch_bounds__(128)
mlaa_line_records_kernel(const uint32_t* __restrict__ fb0, int resX, int resY, const int* __restrict__ list, const int* __restrict__ listCount,
                         LineRec* __restrict__ recH, LineRec* __restrict__ recV, int* __restrict__ cntH, int* __restrict__ cntV,
                         int* __restrict__ endH, int* __restrict__ endV, int capH, int capV)
{
    const int n = *listCount;
    for (int i = (int)(blockIdx.x * blockDim.x + threadIdx.x); i < n; i += (int)(gridDim.x * blockDim.x))
        line_record<BATCH>(fb0, resX, resY, list[i], recH, recV, cntH, cntV, endH, endV, capH, capV);
}

__global__ void mlaa_lines_reset_kernel(int* __restrict__ cnt, int* __restrict__ lastEnd, int n, int* __restrict__ listCount)
{
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i < n) { cnt[i] = 0; lastEnd[i] = -1; }
    if (i == 0) *listCount = 0;
}

// The ordered part: one CTA per 8-row (8-column) block of the given parity, rows in order, one thread per line record.
template <int BATCH>
__global__ void __launch_bounds__(256)
mlaa_blend_lines_kernel(uint32_t* fbi, const uint32_t* __restrict__ fb0, int resX, int resY, int vertical, int yodd,
                        const LineRec* __restrict__ rec, const int* __restrict__ cnt, const int* __restrict__ lastEnd, int cap)
{
    const int rows_per_job = 8;
    int resx, resy, stepy, stepx;
    if (!vertical) { resx = resX; resy = resY; stepy = resX; stepx = 1; }
    else { resx = resY; resy = resX; stepy = 1; stepx = resX; }
    const int jobindex = (int)blockIdx.x;
    const int rfrst = (2 * jobindex + yodd) * rows_per_job;
    int rlast = rfrst + rows_per_job;
    if (rlast >= resy) rlast = resy - 1;
    const int after = stepy;
    const int sz = resX * resY;
    for (int row = rfrst; row < rlast; row++) {
        const int befor = row ? -stepy : 0;
        const int n = min(cnt[row], cap);
        const LineRec* rr = rec + (size_t)row * cap;
        for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) {
            mlaa_line_blend<BATCH>(fbi, rr[i], stepx, befor, after);
        }
        __syncthreads();
        // the SSE scan quirk of the horizontal search (see mlaa_blend_kernel), applied after all lines of this row
        if (!vertical && threadIdx.x == 0) {
            const int kEnd = lastEnd[row];
            if (kEnd >= 0 && (kEnd == resx - 4 || kEnd == resx - 3)) {
                const int base = row * stepy + resx;
                for (int q = 0; q < 4; q++)
                    if (fb0[base + q] & HF) {
                        if (base + q + after < sz) mlaa_blend_one_cell(fbi, base + q, after);
                        break;
                    }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------
// Vertical blending with the strip staged in shared memory by TMA.
// A vertical separation line is walked DOWN a column: consecutive pixels are one row pitch apart (15 KB at 4K), so in
// mlaa_blend_lines_kernel every step of every line is its own 32-byte sector out of L2. One job only ever touches its 8 columns
// and one neighbour column on each side, so the CTA of a job pulls the whole 16-column strip (x0-4 .. x0+11, all rows:
// 64 bytes per row, 138 KB at 2160 rows) into shared memory with cp.async.bulk.tensor.2d boxes of 16 x 128 pixels that all
// complete on one mbarrier, runs the very same blends there (a record's pixel indices are re-based to the strip: row * 16 +
// column - x0), and writes the strip back with cp.async.bulk.tensor stores. Jobs of one parity own disjoint 16-column strips
// (the other columns of a strip are written back unchanged; the job at the left edge uses a 12-column strip from column 0),
// so the reference's job order - even jobs, odd jobs, columns of a
// job in order, one thread per line - is kept and the frame is bit-identical. Rows past the frame / columns left of it are
// zero-filled on load and clipped on store by the TMA unit.
// ---------------------------------------------------------------------------------------------------------
constexpr int STRIP_W = 16, STRIP_BOX_ROWS = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int strip_index(int g, int resX, int x0, int sw)
{
    const int y = g / resX;
    return y * sw + (g - y * resX) - x0;
}

template <int BATCH>
__global__ void __launch_bounds__(256)
mlaa_blend_vstrip_tma_kernel(const __grid_constant__ CUtensorMap frameMap, const __grid_constant__ CUtensorMap firstMap, int resX, int resY,
                             int yodd, const LineRec* __restrict__ rec, const int* __restrict__ cnt, int cap, int nBoxes)
{
    extern __shared__ __align__(128) uint32_t strip[];
    __shared__ __align__(8) unsigned long long bar;
    const int col0 = (2 * (int)blockIdx.x + yodd) * 8;            // first column of the job
    int colEnd = col0 + 8;
    if (colEnd >= resX) colEnd = resX - 1;                         // (rlast >= resy -> resy - 1, as in mlaa_blend_lines_kernel)
    // The job at the left edge has no columns to its left: its strip is columns 0..11 (a second tensor map with 12-pixel boxes) -
    // a TMA STORE must not start at a negative coordinate, and a 16-wide strip from 0 would overlap the next job's.
    const bool first = col0 == 0;
    const int x0 = first ? 0 : col0 - 4, sw = first ? STRIP_W - 4 : STRIP_W;
    const uint32_t barAddr = smem_u32(&bar), stripAddr = smem_u32(strip);
    const unsigned long long mapAddr = reinterpret_cast<unsigned long long>(first ? &firstMap : &frameMap);
    const uint32_t boxBytes = (uint32_t)sw * STRIP_BOX_ROWS * 4;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barAddr), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barAddr), "r"(boxBytes * (uint32_t)nBoxes) : "memory");
        for (int k = 0; k < nBoxes; k++)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(stripAddr + (uint32_t)k * boxBytes), "l"(mapAddr), "r"(barAddr), "r"(x0), "r"(k * STRIP_BOX_ROWS) : "memory");
    }
    {   // every thread waits for the strip (phase 0 of the barrier)
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(barAddr), "r"(0) : "memory");
    }
    for (int col = col0; col < colEnd; col++) {
        const int befor = col ? -1 : 0;
        const int n = min(cnt[col], cap);
        const LineRec* rr = rec + (size_t)col * cap;
        for (int i = (int)threadIdx.x; i < n; i += (int)blockDim.x) {
            LineRec r = rr[i];
            if (r.ui0 >= 0) r.ui0 = strip_index(r.ui0, resX, x0, sw);  // -1: no such end point
            if (r.ui1 >= 0) r.ui1 = strip_index(r.ui1, resX, x0, sw);  // -2: one-pixel line, -1: none
            if (r.li0 >= 0) r.li0 = strip_index(r.li0, resX, x0, sw);
            if (r.li1 >= 0) r.li1 = strip_index(r.li1, resX, x0, sw);
            mlaa_line_blend<BATCH>(strip, r, sw, befor, 1);
        }
        __syncthreads();
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the generic-proxy writes above, before the async-proxy reads below
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < nBoxes; k++)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(mapAddr), "r"(stripAddr + (uint32_t)k * boxBytes), "r"(x0), "r"(k * STRIP_BOX_ROWS) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (EncodeTiledFn)p;
        else cudaGetLastError();
    }
    return fn;
}

// The frame as a 2-D tensor of 32-bit pixels, boxes of 16 x 128. false: no TMA path for this frame (caller uses the plain kernel).
bool make_frame_map(CUtensorMap* map, uint32_t* d_frame, int resX, int resY, int boxW = STRIP_W)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || (reinterpret_cast<uintptr_t>(d_frame) & 15u) || (resX % 4)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)resX, (cuuint64_t)resY};
    const cuuint64_t strides[1] = {(cuuint64_t)resX * 4};
    const cuuint32_t box[2] = {(cuuint32_t)boxW, STRIP_BOX_ROWS};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_frame, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

size_t mlaa_lines_bytes(int resX, int resY)
{
    const size_t capH = (size_t)resX / 2 + 1, capV = (size_t)resY / 2 + 1;
    // records per row / column, their counts and right-most ends, then the list of line starts (every line covers >= 1 flagged
    // pixel of its orientation: at most resX*resY starts in all... per orientation at most half the pixels) + its counter
    return ((size_t)resY * capH + (size_t)resX * capV) * sizeof(LineRec) + 2 * ((size_t)resX + resY) * sizeof(int) +
           ((size_t)resX * resY + 2 * ((size_t)resX + resY) + 4) * sizeof(int);
}

cudaError_t launch_mlaa(uint32_t* d_frame, uint32_t* d_scratch, int resX, int resY, int numSMs, cudaStream_t st, int& launches,
                        void* d_lines, const Switches& sw)
{
    mlaa_find_fragments_kernel<<<numSMs * 4, 256, 0, st>>>(d_frame, d_scratch, resX, resY);
    const int n_hscan_jobs = (resY / 8) + ((resY % 8) ? 1 : 0);
    const int n_vscan_jobs = (resX / 8) + ((resX % 8) ? 1 : 0);
    // job list halves (MLAA.cc:545-552): the first scanjobs/2 jobs are the even blocks, the rest the odd blocks
    const int h0 = n_hscan_jobs / 2, h1 = n_hscan_jobs - h0, v0 = n_vscan_jobs / 2, v1 = n_vscan_jobs - v0;
    if (d_lines) {
        // two-stage path: all lines of the frame first (order-independent), then the ordered in-place blends
        const int capH = resX / 2 + 1, capV = resY / 2 + 1;
        LineRec* recH = reinterpret_cast<LineRec*>(d_lines);
        LineRec* recV = recH + (size_t)resY * capH;
        int* cntH = reinterpret_cast<int*>(recV + (size_t)resX * capV);
        int* cntV = cntH + resY; int* endH = cntV + resX; int* endV = endH + resY;
        int* listCount = endV + resX; int* list = listCount + 4;
        mlaa_lines_reset_kernel<<<(resX + resY + 255) / 256, 256, 0, st>>>(cntH, endH, resX + resY, listCount);   // cntH|cntV and endH|endV are contiguous
        // batched flag / pixel loads need runs shorter than a row/column by a wide margin (mlaa_steps.h); tiny frames walk step by step
        const bool batch = resX >= 4 * BLEND_BATCH && resY >= 4 * BLEND_BATCH && !sw.mlaa_nobatch;
        const bool fullScan = sw.mlaa_fullscan != 0;
        if (fullScan) mlaa_lines_kernel<<<numSMs * 8, 256, 0, st>>>(d_scratch, resX, resY, recH, recV, cntH, cntV, endH, endV, capH, capV);
        else {
            mlaa_line_starts_kernel<<<numSMs * 8, 256, 0, st>>>(d_scratch, resX, resY, list, listCount);
            auto records = batch ? mlaa_line_records_kernel<BLEND_BATCH> : mlaa_line_records_kernel<1>;
            records<<<numSMs * 8, 128, 0, st>>>(d_scratch, resX, resY, list, listCount, recH, recV, cntH, cntV, endH, endV, capH, capV);
            launches += 1;
        }
        auto blend = batch ? mlaa_blend_lines_kernel<BLEND_BATCH> : mlaa_blend_lines_kernel<1>;
        if (h0 > 0) blend<<<h0, 256, 0, st>>>(d_frame, d_scratch, resX, resY, 0, 0, recH, cntH, endH, capH);
        if (h1 > 0) blend<<<h1, 256, 0, st>>>(d_frame, d_scratch, resX, resY, 0, 1, recH, cntH, endH, capH);
        // vertical jobs: the 16-column strip of a job staged in shared memory by TMA when it fits (else / mlaa_no_tma: walked in L2)
        const int nBoxes = (resY + STRIP_BOX_ROWS - 1) / STRIP_BOX_ROWS;
        const size_t stripBytes = (size_t)nBoxes * STRIP_BOX_ROWS * STRIP_W * 4;
        CUtensorMap map, map12;
        if (!sw.mlaa_no_tma && stripBytes <= 200 * 1024 && resX >= 32 && make_frame_map(&map, d_frame, resX, resY) &&
            make_frame_map(&map12, d_frame, resX, resY, STRIP_W - 4)) {
            auto vstrip = batch ? mlaa_blend_vstrip_tma_kernel<BLEND_BATCH> : mlaa_blend_vstrip_tma_kernel<1>;
            static bool attr[2] = {false, false};
            if (!attr[batch ? 1 : 0]) {
                cudaError_t e = cudaFuncSetAttribute(vstrip, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
                if (e != cudaSuccess) return e;
                attr[batch ? 1 : 0] = true;
            }
            if (v0 > 0) vstrip<<<v0, 256, stripBytes, st>>>(map, map12, resX, resY, 0, recV, cntV, capV, nBoxes);
            if (v1 > 0) vstrip<<<v1, 256, stripBytes, st>>>(map, map12, resX, resY, 1, recV, cntV, capV, nBoxes);
        } else {
            if (v0 > 0) blend<<<v0, 256, 0, st>>>(d_frame, d_scratch, resX, resY, 1, 0, recV, cntV, endV, capV);
            if (v1 > 0) blend<<<v1, 256, 0, st>>>(d_frame, d_scratch, resX, resY, 1, 1, recV, cntV, endV, capV);
        }
        launches += 7;
        return cudaGetLastError();
    }
    if (h0 > 0) mlaa_blend_kernel<1><<<h0, 256, 0, st>>>(d_frame, d_scratch, resX, resY, 0, 0);
    if (h1 > 0) mlaa_blend_kernel<1><<<h1, 256, 0, st>>>(d_frame, d_scratch, resX, resY, 0, 1);
    if (v0 > 0) mlaa_blend_kernel<1><<<v0, 256, 0, st>>>(d_frame, d_scratch, resX, resY, 1, 0);
    if (v1 > 0) mlaa_blend_kernel<1><<<v1, 256, 0, st>>>(d_frame, d_scratch, resX, resY, 1, 1);
    launches += 5;
    return cudaGetLastError();
}

}  // namespace b200r
