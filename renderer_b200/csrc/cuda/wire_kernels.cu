// wire_kernels.cu — render mode 3: anti-aliased (Wu) wireframe, bit-exact with the reference's serial blending.
//
// Replaces Scene::renderWireframe (reference src/Rasterizers.cc:117-183) and the SDL_gfx-derived line code it calls:
//   my_aalineColor / _aalineColor   src/Wu.cc:1282-1512      _clipLine / _clipEncode   src/Wu.cc:949-1052
//   lineColor (alpha branch)        src/Wu.cc:1070-1255      hlineColor / vlineColor   src/Wu.cc:652-937
//   filledRectAlpha                 src/Wu.cc:320-575        _putPixelAlpha (32 bpp)   src/Wu.cc:163-193
// Every pixel operation of that code is "blend the line colour into the pixel with some alpha"
// (dst = dst + ((src - dst) * alpha >> 8), per channel in wrap-around uint32 arithmetic); the blend is not
// commutative, and the CPU draws triangle after triangle, line after line. The device formulation:
//   1. wire_lines_kernel<0>: one thread per triangle runs the SAME integer line algorithms and counts fragments
//      per pixel;                    2. exclusive scan of the counts;
//   3. wire_lines_kernel<1>: the same walk again, scattering {sequence number, alpha} into each pixel's segment;
//   4. wire_fold_kernel: one thread per touched pixel sorts its fragments by sequence number
//      (triangle, line, step) and folds the blends in exactly the reference's order.
// The colour passed down is SDL_MapRGB(200,200,200)=0x00C8C8C8 but decoded as 0xRRGGBBAA: R=0,G=200,B=200, alpha 200.
#include "device_types.cuh"
#include "rt_kernels.cuh"

namespace b200r {
namespace {

constexpr uint32_t kGreyPixel = 0x00C8C8C8u;
constexpr float kClip = 0.2f;

__device__ __forceinline__ uint32_t map_rgba(uint32_t color)
{
    return (((color >> 24) & 0xffu) << 16) | (((color >> 16) & 0xffu) << 8) | ((color >> 8) & 0xffu);
}

// per-channel blend of _putPixelAlpha / _filledRectAlpha (src/Wu.cc:170-190), no alpha channel on the surface
__device__ __forceinline__ uint32_t blend(uint32_t dc, uint32_t color, uint32_t alpha)
{
    const uint32_t Rm = 0x00FF0000u, Gm = 0x0000FF00u, Bm = 0x000000FFu;
    const uint32_t R = ((dc & Rm) + (((((color & Rm) - (dc & Rm)) >> 16) * alpha >> 8) << 16)) & Rm;
    const uint32_t G = ((dc & Gm) + (((((color & Gm) - (dc & Gm)) >> 8) * alpha >> 8) << 8)) & Gm;
    const uint32_t B = ((dc & Bm) + ((((color & Bm) - (dc & Bm)) * alpha >> 8))) & Bm;
    return R | G | B;
}

template <int PASS>
struct Emitter {
    int W, H, rowFirst, rowStep;
    uint32_t* counts; const uint32_t* offsets; uint2* frags; uint32_t capacity;
    uint32_t seqHi, k;
    __device__ __forceinline__ void put(int x, int y, uint32_t alpha)
    {
        const uint32_t kk = k++;
        if (y < rowFirst || ((y - rowFirst) % rowStep) != 0) return;
        const size_t p = (size_t)((y - rowFirst) / rowStep) * W + x;
        if (PASS == 0) atomicAdd(&counts[p], 1u);
        else {
            const uint32_t pos = offsets[p] + atomicSub(&counts[p], 1u) - 1u;
            if (pos < capacity) frags[pos] = make_uint2(seqHi, (kk << 8) | (alpha & 0xffu));
        }
    }
    // _putPixelAlpha: clip test, then blend (src/Wu.cc:59-60)
    __device__ __forceinline__ void pixel(short x, short y, uint32_t alpha)
    {
        if (x >= 0 && x <= W - 1 && y >= 0 && y <= H - 1) put(x, y, alpha);
    }
};

template <class E> __device__ __forceinline__ void pixelColorNolock(E& e, short x, short y, uint32_t color) { e.pixel(x, y, color & 0xffu); }
template <class E> __device__ __forceinline__ void pixelColorWeightNolock(E& e, short x, short y, uint32_t color, uint32_t weight)
{
    const uint32_t a = ((color & 0xffu) * weight) >> 8;
    e.pixel(x, y, a);
}

// filledRectAlpha (no clipping inside; callers clipped) - alpha 255 would still blend on this path
template <class E> __device__ void filledRectAlpha(E& e, short x1, short y1, short x2, short y2, uint32_t color)
{
    const uint32_t alpha = color & 0xffu;
    for (int y = y1; y <= y2; y++)
        for (int x = x1; x <= x2; x++) e.put(x, y, alpha);
}

template <class E> __device__ void hlineColor(E& e, short x1, short x2, short y, uint32_t color)
{
    if (x1 > x2) { const short t = x1; x1 = x2; x2 = t; }
    const short left = 0, right = (short)(e.W - 1), top = 0, bottom = (short)(e.H - 1);
    if (x2 < left) return;
    if (x1 > right) return;
    if ((y < top) || (y > bottom)) return;
    if (x1 < left) x1 = left;
    if (x2 > right) x2 = right;
    const int dx = x2 - x1;
    filledRectAlpha(e, x1, y, (short)(x1 + dx), y, color);
}

template <class E> __device__ void vlineColor(E& e, short x, short y1, short y2, uint32_t color)
{
    if (y1 > y2) { const short t = y1; y1 = y2; y2 = t; }
    const short left = 0, right = (short)(e.W - 1), top = 0, bottom = (short)(e.H - 1);
    if ((x < left) || (x > right)) return;
    if (y2 < top) return;
    if (y1 > bottom) return;
    if (y1 < top) y1 = top;
    if (y2 > bottom) y2 = bottom;
    const short h = (short)(y2 - y1);
    filledRectAlpha(e, x, y1, x, (short)(y1 + h), color);
}

__device__ __forceinline__ int clipEncode(short x, short y, short left, short top, short right, short bottom)
{
    int code = 0;
    if (x < left) code |= 1; else if (x > right) code |= 2;
    if (y < top) code |= 8; else if (y > bottom) code |= 4;
    return code;
}

// _clipLine (src/Wu.cc:990-1052): float slope, results cast to Sint16 the way x86 does it
__device__ int clipLine(int W, int H, short& x1, short& y1, short& x2, short& y2)
{
    const short left = 0, right = (short)(W - 1), top = 0, bottom = (short)(H - 1);
    int draw = 0;
    for (int guard = 0; guard < 64; guard++) {                 // the reference loops "while (1)"; 4 rounds suffice
        int code1 = clipEncode(x1, y1, left, top, right, bottom);
        const int code2 = clipEncode(x2, y2, left, top, right, bottom);
        if (!(code1 | code2)) { draw = 1; break; }
        else if (code1 & code2) break;
        else {
            if (!code1) {
                short t = x2; x2 = x1; x1 = t;
                t = y2; y2 = y1; y1 = t;
                code1 = code2;
            }
            float m;
            if (x2 != x1) m = (float)((int)y2 - (int)y1) / (float)((int)x2 - (int)x1); else m = 1.0f;
            if (code1 & 1) { y1 = (short)(y1 + (short)cvtt_x86((float)((int)left - (int)x1) * m)); x1 = left; }
            else if (code1 & 2) { y1 = (short)(y1 + (short)cvtt_x86((float)((int)right - (int)x1) * m)); x1 = right; }
            else if (code1 & 4) { if (x2 != x1) x1 = (short)(x1 + (short)cvtt_x86((float)((int)bottom - (int)y1) / m)); y1 = bottom; }
            else if (code1 & 8) { if (x2 != x1) x1 = (short)(x1 + (short)cvtt_x86((float)((int)top - (int)y1) / m)); y1 = top; }
        }
    }
    return draw;
}

// lineColor, alpha branch (src/Wu.cc:1206-1252)
template <class E> __device__ void lineColor(E& e, short x1, short y1, short x2, short y2, uint32_t color)
{
    if (!clipLine(e.W, e.H, x1, y1, x2, y2)) return;
    if (x1 == x2) {
        if (y1 < y2) { vlineColor(e, x1, y1, y2, color); return; }
        else if (y1 > y2) { vlineColor(e, x1, y2, y1, color); return; }
        else { pixelColorNolock(e, x1, y1, color); return; }
    }
    if (y1 == y2) {
        if (x1 < x2) { hlineColor(e, x1, x2, y1, color); return; }
        else if (x1 > x2) { hlineColor(e, x2, x1, y1, color); return; }
    }
    const int dx = x2 - x1, dy = y2 - y1;
    const int sx = (dx >= 0) ? 1 : -1, sy = (dy >= 0) ? 1 : -1;
    const int ax = abs(dx) << 1, ay = abs(dy) << 1;
    int x = x1, y = y1;
    if (ax > ay) {
        int d = ay - (ax >> 1);
        while (x != x2) {
            pixelColorNolock(e, (short)x, (short)y, color);
            if (d > 0 || (d == 0 && sx == 1)) { y += sy; d -= ax; }
            x += sx; d += ay;
        }
    } else {
        int d = ax - (ay >> 1);
        while (y != y2) {
            pixelColorNolock(e, (short)x, (short)y, color);
            if (d > 0 || ((d == 0) && (sy == 1))) { x += sx; d -= ay; }
            y += sy; d += ax;
        }
    }
    pixelColorNolock(e, (short)x, (short)y, color);
}

// _aalineColor(..., draw_endpoint = 1) (src/Wu.cc:1282-1495)
template <class E> __device__ void aalineColor(E& e, short x1, short y1, short x2, short y2, uint32_t color)
{
    if (!clipLine(e.W, e.H, x1, y1, x2, y2)) return;
    int xx0 = x1, yy0 = y1, xx1 = x2, yy1 = y2;
    if (yy0 > yy1) { int t = yy0; yy0 = yy1; yy1 = t; t = xx0; xx0 = xx1; xx1 = t; }
    int dx = xx1 - xx0, dy = yy1 - yy0;
    if (dx == 0) { vlineColor(e, x1, y1, y2, color); return; }
    else if (dy == 0) { hlineColor(e, x1, x2, y1, color); return; }
    else if (dx == dy) { lineColor(e, x1, y1, x2, y2, color); return; }
    int xdir;
    if (dx >= 0) xdir = 1; else { xdir = -1; dx = -dx; }
    uint32_t erracc = 0;
    pixelColorNolock(e, x1, y1, color);
    if (dy > dx) {
        const uint32_t erradj = (uint32_t)(((uint32_t)dx << 16) / (uint32_t)dy) << 16;
        int x0pxdir = xx0 + xdir;
        while (--dy) {
            const uint32_t erracctmp = erracc;
            erracc += erradj;
            if (erracc <= erracctmp) { xx0 = x0pxdir; x0pxdir += xdir; }
            yy0++;
            const uint32_t wgt = (erracc >> 24) & 255u;
            pixelColorWeightNolock(e, (short)xx0, (short)yy0, color, 255u - wgt);
            pixelColorWeightNolock(e, (short)x0pxdir, (short)yy0, color, wgt);
        }
    } else {
        const uint32_t erradj = (uint32_t)(((uint32_t)dy << 16) / (uint32_t)dx) << 16;
        int y0p1 = yy0 + 1;
        while (--dx) {
            const uint32_t erracctmp = erracc;
            erracc += erradj;
            if (erracc <= erracctmp) { yy0 = y0p1; y0p1++; }
            xx0 += xdir;
            const uint32_t wgt = (erracc >> 24) & 255u;
            pixelColorWeightNolock(e, (short)xx0, (short)yy0, color, 255u - wgt);
            pixelColorWeightNolock(e, (short)xx0, (short)y0p1, color, wgt);
        }
    }
    pixelColorNolock(e, x2, y2, color);
}

template <int PASS>
__global__ void __launch_bounds__(128)
wire_lines_kernel(DeviceScene sc, FrameParams fp, uint32_t* counts, const uint32_t* offsets, uint2* frags, uint32_t capacity)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= sc.n_tris) return;
    const V3 eye = mkv3(fp.eye[0], fp.eye[1], fp.eye[2]);
    const float4 t0 = __ldg(sc.rtris + 4 * (size_t)j + 0);
    const float4 t1 = __ldg(sc.rtris + 4 * (size_t)j + 1);
    const float4 t2 = __ldg(sc.rtris + 4 * (size_t)j + 2);
    if (dot3(eye - mkv3(t1.x, t1.y, t1.z), mkv3(t2.x, t2.y, t2.z)) < 0.f) return;
    const uint32_t vi[3] = {__float_as_uint(t0.x), __float_as_uint(t0.y), __float_as_uint(t0.z)};
    V3 c[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { const float4 p = __ldg(sc.rverts + 2 * (size_t)vi[k]); c[k] = transform3(mkv3(p.x, p.y, p.z), eye, fp.mv); }
    const int W = (int)fp.W, H = (int)fp.H;
    const float SD = (float)(H * 2);
    int sx[3], sy[3]; bool good[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        good[k] = c[k].z > kClip;
        sx[k] = cvtt_x86((float)(W / 2) + (SD * c[k].y) / c[k].z);
        sy[k] = cvtt_x86((float)(H / 2) - (SD * c[k].x) / c[k].z);
    }
    Emitter<PASS> e;
    e.W = W; e.H = H; e.rowFirst = (int)fp.row_first; e.rowStep = (int)fp.row_step;
    e.counts = counts; e.offsets = offsets; e.frags = frags; e.capacity = capacity;
    // line order of the reference (src/Rasterizers.cc:154-180): AB, AC, BC (those whose endpoints are in front)
    if (good[0]) {
        if (good[1]) {
            e.seqHi = j * 3u + 0u; e.k = 0; aalineColor(e, (short)sx[0], (short)sy[0], (short)sx[1], (short)sy[1], kGreyPixel);
            if (good[2]) {
                e.seqHi = j * 3u + 1u; e.k = 0; aalineColor(e, (short)sx[0], (short)sy[0], (short)sx[2], (short)sy[2], kGreyPixel);
                e.seqHi = j * 3u + 2u; e.k = 0; aalineColor(e, (short)sx[1], (short)sy[1], (short)sx[2], (short)sy[2], kGreyPixel);
            }
        } else if (good[2]) {
            e.seqHi = j * 3u + 1u; e.k = 0; aalineColor(e, (short)sx[0], (short)sy[0], (short)sx[2], (short)sy[2], kGreyPixel);
        }
    } else if (good[1] && good[2]) {
        e.seqHi = j * 3u + 2u; e.k = 0; aalineColor(e, (short)sx[1], (short)sy[1], (short)sx[2], (short)sy[2], kGreyPixel);
    }
}

// ---- exclusive scan of the per-pixel counts (n+1 outputs), three small kernels
constexpr int SCAN_BLOCK = 1024;
__global__ void scan_block_sums_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ blockSums, size_t n)
{
    __shared__ uint32_t s[32];
    const size_t i = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    uint32_t v = i < n ? in[i] : 0u;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = s[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) blockSums[blockIdx.x] = w;
    }
}
__global__ void scan_sums_kernel(uint32_t* blockSums, uint32_t nBlocks, uint32_t* total)
{
    // single thread block, sequential over chunks: nBlocks <= 65536
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nBlocks; base += SCAN_BLOCK) {
        __shared__ uint32_t s[SCAN_BLOCK];
        const uint32_t i = base + threadIdx.x;
        s[threadIdx.x] = i < nBlocks ? blockSums[i] : 0u;
        __syncthreads();
        for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
            const uint32_t t = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0u;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        const uint32_t incl = s[threadIdx.x], c = carry;
        const uint32_t excl = threadIdx.x ? s[threadIdx.x - 1] : 0u;
        __syncthreads();
        if (i < nBlocks) blockSums[i] = c + excl;                 // exclusive prefix of the block sums
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) carry = c + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
__global__ void scan_apply_kernel(const uint32_t* __restrict__ in, const uint32_t* __restrict__ blockOffsets,
                                  uint32_t* __restrict__ out, size_t n, const uint32_t* __restrict__ total)
{
    __shared__ uint32_t s[SCAN_BLOCK];
    const size_t i = (size_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
    const uint32_t v = i < n ? in[i] : 0u;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_BLOCK; o <<= 1) {
        const uint32_t t = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0u;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    if (i < n) out[i] = blockOffsets[blockIdx.x] + s[threadIdx.x] - v;
    if (i == 0) out[n] = *total;
}

// ---- per pixel: order the fragments as the serial reference produced them, fold the blends
__global__ void wire_fold_kernel(const uint32_t* __restrict__ offsets, uint2* __restrict__ frags, uint32_t capacity,
                                 uint32_t* __restrict__ out, size_t nPixels)
{
    const uint32_t mcolor = map_rgba(kGreyPixel);
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < nPixels; p += (size_t)gridDim.x * blockDim.x) {
        const uint32_t b = offsets[p], e = offsets[p + 1];
        if (e == b || e > capacity) continue;
        for (uint32_t i = b + 1; i < e; i++) {                    // insertion sort by (triangle*3+line, step)
            const uint2 key = frags[i];
            uint32_t jx = i;
            while (jx > b) {
                const uint2 q = frags[jx - 1];
                if (q.x < key.x || (q.x == key.x && (q.y >> 8) < (key.y >> 8))) break;
                frags[jx] = q; jx--;
            }
            frags[jx] = key;
        }
        uint32_t c = 0u;                                          // Screen::ClearScreen
        // (_putPixelAlpha stores the colour directly when alpha == 255; with the line colour's alpha of 200, and
        //  Wu weights (200*w)>>8 <= 199, that never happens, so every fragment is a blend)
        for (uint32_t i = b; i < e; i++) c = blend(c, mcolor, frags[i].y & 0xffu);
        out[p] = c;
    }
}

}  // namespace

// Pass A: count. Returns after enqueueing; the host reads *d_total (fragment count) to size the fragment buffer.
cudaError_t launch_wire_count(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, WireBuffers& wb, cudaStream_t st,
                              int& launches)
{
    const size_t px = (size_t)fp.W * fp.n_rows;
    cudaError_t e = cudaMemsetAsync(d_out, 0, px * 4, st);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(wb.counts, 0, px * 4, st);
    if (e != cudaSuccess) return e;
    wire_lines_kernel<0><<<(sc.n_tris + 127) / 128, 128, 0, st>>>(sc, fp, wb.counts, nullptr, nullptr, 0);
    const uint32_t nBlocks = (uint32_t)((px + SCAN_BLOCK - 1) / SCAN_BLOCK);
    scan_block_sums_kernel<<<nBlocks, SCAN_BLOCK, 0, st>>>(wb.counts, wb.blockSums, px);
    scan_sums_kernel<<<1, SCAN_BLOCK, 0, st>>>(wb.blockSums, nBlocks, wb.total);
    scan_apply_kernel<<<nBlocks, SCAN_BLOCK, 0, st>>>(wb.counts, wb.blockSums, wb.offsets, px, wb.total);
    launches += 4;
    return cudaGetLastError();
}

cudaError_t launch_wire_emit(const DeviceScene& sc, const FrameParams& fp, uint32_t* d_out, WireBuffers& wb, int numSMs,
                             cudaStream_t st, int& launches)
{
    const size_t px = (size_t)fp.W * fp.n_rows;
    wire_lines_kernel<1><<<(sc.n_tris + 127) / 128, 128, 0, st>>>(sc, fp, wb.counts, wb.offsets, reinterpret_cast<uint2*>(wb.frags), wb.capacity);
    wire_fold_kernel<<<numSMs * 8, 256, 0, st>>>(wb.offsets, reinterpret_cast<uint2*>(wb.frags), wb.capacity, d_out, px);
    launches += 2;
    return cudaGetLastError();
}

}  // namespace b200r
