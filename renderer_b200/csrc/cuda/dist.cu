// dist.cu — frames in flight and multi-GPU frame assembly behind the C-ABI (include/b200render.h (1b)).
//
// One b200r_pipeline per rank (one process - or one thread - per GPU). Pixels are independent, so rank r of P renders rows
// r, r+P, r+2P, ... of every frame (b200r_frame.row_first/row_step) into a packed shard; the scene is replicated and the
// Z-buffer never leaves a GPU. Frames are independent too (the reference's own benchmark loop, src/renderer.cc:491-606), so up
// to `depth` of them are in flight per rank: frame i uses slot i % depth = one render stream + one set of the renderer's
// scratch buffers + one shard + one assembled frame.
//
// Assembly of the finished rows into a scan-order frame on EVERY rank, two interchangeable ways:
//   B200R_ASSEMBLE_NCCL  one ncclAllGather of the packed shards per frame (the north star's collective; the parity reference),
//                        then a de-interleave kernel.
//   B200R_ASSEMBLE_PUSH  one kernel per frame and rank that stores the rank's rows straight into every rank's scan-order frame
//                        through NVLink peer mappings (cudaIpc handles between processes, plain peer access between threads),
//                        16 bytes per store, followed by ONE release-add per peer on that frame's arrival counter. The consumer
//                        stream waits for "P arrivals" with a one-thread polling kernel (acquire loads + __nanosleep). No second
//                        pass over the frame, no collective kernel competing with the persistent render kernels for SMs.
//                        Buffer reuse is flow-controlled the same way: a consumer acknowledges a slot to all producers, a producer
//                        waits for P acknowledgements of the slot's previous frame before it overwrites it.
// NCCL is loaded at run time (dlopen) and only when world > 1: the library has no link-time dependency on it.
#include <dlfcn.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "device_types.cuh"
#include "rt_kernels.cuh"

namespace b200r {
void set_global_error(const std::string& s);
int ctx_device(const b200r_ctx* ctx);
int ctx_sms(const b200r_ctx* ctx);
}
using namespace b200r;

namespace {

// ---- NCCL through dlopen: the six entry points the all-gather path and the bootstrap need
struct Id128 { char b[128]; };                 // ncclUniqueId (passed by value)
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

bool load_nccl(std::string& err)
{
    if (g_nccl.ok) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
    if (!g_nccl.lib) { err = std::string("cannot load libnccl.so.2: ") + dlerror(); return false; }
    auto sym = [&](const char* s) { return dlsym(g_nccl.lib, s); };
    g_nccl.GetUniqueId = (int (*)(void*))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, Id128, int))sym("ncclCommInitRank");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))sym("ncclAllGather");
    g_nccl.CommDestroy = (int (*)(void*))sym("ncclCommDestroy");
    g_nccl.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) { err = "libnccl lacks an entry point"; return false; }
    g_nccl.ok = true;
    return true;
}
constexpr int NCCL_INT8 = 0;      // ncclInt8 / ncclChar

// ---- stream memory operation of the driver API, without linking libcuda
typedef int (*StreamWaitValue32Fn)(cudaStream_t, unsigned long long, unsigned, unsigned);
StreamWaitValue32Fn g_waitValue = nullptr;
bool load_wait_value()
{
    if (g_waitValue) return true;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { cudaGetLastError(); return false; }
    g_waitValue = (StreamWaitValue32Fn)fn;
    return true;
}
constexpr unsigned WAIT_GEQ = 0x0;      // CU_STREAM_WAIT_VALUE_GEQ

constexpr int MAXD = B200R_MAX_FRAMES_IN_FLIGHT;
constexpr int MAXP = 16;

// Rank `rank` of P stores its packed rows (row k of the shard is screen row rank + k*P) into the scan-order frames of all P ranks,
// 16 bytes per store, rows dealt to CTAs; the last CTA to finish publishes the shard: one system-scope release-add per peer.
struct PushArgs {
    uint32_t* dst[MAXP];
    unsigned* arrive[MAXP];           // each peer's arrival counter of this slot
};
__global__ void __launch_bounds__(256)
push_rows_kernel(const uint32_t* __restrict__ shard, PushArgs a, int P, int rank, int W, int nRows, unsigned* __restrict__ blocksDone)
{
    const int vecPerRow = W >> 2;             // W % 4 == 0 (checked by the host)
    const size_t total = (size_t)nRows * vecPerRow;
    const uint4* src = reinterpret_cast<const uint4*>(shard);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i / vecPerRow), c = (int)(i % vecPerRow);
        const uint4 v = src[i];
        const size_t o = (size_t)(rank + k * P) * vecPerRow + c;
        for (int p = 0; p < P; p++) reinterpret_cast<uint4*>(a.dst[p])[o] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(blocksDone, 1u);
        if (done == gridDim.x - 1) {
            *blocksDone = 0u;
            __threadfence_system();
            for (int p = 0; p < P; p++) atomicAdd_system(a.arrive[p], 1u);
        }
    }
}

// Wait on a stream until a counter in THIS GPU's memory (written by peers with system-scope release-adds) has reached `target`.
// One thread polls with acquire loads and backs off with __nanosleep: it wakes within a microsecond of the last arrival. (The driver's
// stream memory operation - cuStreamWaitValue32, kept as B200R_PIPE_MEMOP=1 - holds no SM at all, but when the value is not there yet
// its polling backs off: measured up to ~0.8 ms of extra latency per frame on one rank of two with one frame in flight.)
__global__ void wait_counter_kernel(const unsigned* __restrict__ ctr, unsigned target)
{
    if (threadIdx.x == 0) {
        for (;;) {
            unsigned v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if ((int)(v - target) >= 0) break;
            __nanosleep(200);
        }
    }
}

struct SignalArgs { unsigned* word[MAXP]; };
__global__ void signal_peers_kernel(SignalArgs a, int P)
{
    if (threadIdx.x < (unsigned)P) atomicAdd_system(a.word[threadIdx.x], 1u);
}

// Measurement aid (b200r_pipeline_set_l2_flush): write a buffer larger than the L2 so that the frame behind it starts from a cold
// cache. The same bytes as a cudaMemsetAsync of the buffer, but from two small CTAs per SM instead of the runtime's 73 728-CTA fill
// kernel: the persistent render CTAs of the frames in flight own every register of an SM, so a fill grid of that size cannot run
// beside them - it displaces them for the whole 23 us the write takes (measured on one rank of 8 emulated, 8 frames in flight:
// 15 260 fps with cudaMemsetAsync, 23 370 fps without any flush) - while 16-byte stores from 256 threads per SM already saturate HBM.
__global__ void __launch_bounds__(128) l2_flush_kernel(uint4* __restrict__ p, size_t n16)
{
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const size_t stride = (size_t)gridDim.x * 128;
    size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) { p[i] = z; p[i + stride] = z; p[i + 2 * stride] = z; p[i + 3 * stride] = z; }
    for (; i < n16; i += stride) p[i] = z;
}

// prefetch a buffer into L2 (after a bench-mode L2 flush the scene would otherwise come back one 64-byte miss at a time)
__global__ void l2_prefetch_kernel(const char* p, size_t bytes)
{
    for (size_t o = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128; o < bytes; o += (size_t)gridDim.x * blockDim.x * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
}

struct PeerBlob {
    int pid, device;
    unsigned long long raw[MAXD + 1];             // full[0..D-1], counters
    cudaIpcMemHandle_t handle[MAXD + 1];
};

}  // namespace

struct b200r_pipeline {
    b200r_ctx* ctx = nullptr;
    int device = 0, sms = 0;
    uint32_t W = 0, H = 0, P = 1, rank = 0, D = 2, mode = 0, rps = 0;
    cudaStream_t rs[MAXD] = {}, push = nullptr, consume = nullptr;
    cudaEvent_t rendered[MAXD] = {}, pushed[MAXD] = {}, consumed[MAXD] = {}, gathered_ev[MAXD] = {};
    uint32_t* shard[MAXD] = {};
    uint32_t* gathered[MAXD] = {};
    uint32_t* full[MAXD] = {};
    unsigned* counters = nullptr;                 // [0..D) arrival counters, [D..2D) acknowledgement counters, [2D..3D) CTAs done
    uint32_t* peerFull[MAXP][MAXD] = {};
    unsigned* peerCounters[MAXP] = {};
    std::vector<void*> opened;                    // cudaIpcOpenMemHandle mappings to close
    void* nccl = nullptr;
    uint64_t submitted = 0;
    void* flushBuf = nullptr; size_t flushBytes = 0;
    bool flushMemset = false;                     // B200R_PIPE_FLUSH_MEMSET=1: the flush by cudaMemsetAsync (A/B)
    const char* prefetch[4] = {}; size_t prefetchBytes[4] = {};
    uint32_t launches = 0;
    bool timing = false;
    bool memop = false;                           // waits by cuStreamWaitValue32 instead of wait_counter_kernel
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> kev;      // per submitted frame: around this rank's render kernels (timing on)
    std::string err;
};

namespace {
int pfail(b200r_pipeline* p, int code, const std::string& msg)
{
    if (p) p->err = msg;
    set_global_error(msg);
    return code;
}
#define PCU(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return pfail(pipe, B200R_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)
#define PNCCL(call)                                                                                \
    do {                                                                                           \
        int r__ = (call);                                                                          \
        if (r__ != 0)                                                                              \
            return pfail(pipe, B200R_ECUDA, std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "NCCL error")); \
    } while (0)
}  // namespace

extern "C" {

int b200r_dist_unique_id(void* out128)
{
    std::string err;
    if (!out128) { set_global_error("b200r_dist_unique_id: NULL"); return B200R_EINVAL; }
    if (!load_nccl(err)) { set_global_error(err); return B200R_ESTATE; }
    if (g_nccl.GetUniqueId(out128) != 0) { set_global_error("ncclGetUniqueId failed"); return B200R_ECUDA; }
    return B200R_OK;
}

const char* b200r_pipeline_last_error(const b200r_pipeline* pipe) { return pipe ? pipe->err.c_str() : ""; }

void b200r_pipeline_destroy(b200r_pipeline* pipe)
{
    if (!pipe) return;
    cudaSetDevice(pipe->device);
    cudaDeviceSynchronize();
    for (void* m : pipe->opened) cudaIpcCloseMemHandle(m);
    if (pipe->nccl && g_nccl.ok) g_nccl.CommDestroy(pipe->nccl);
    for (uint32_t d = 0; d < pipe->D; d++) {
        cudaFree(pipe->shard[d]); cudaFree(pipe->gathered[d]); cudaFree(pipe->full[d]);
        if (pipe->rs[d]) cudaStreamDestroy(pipe->rs[d]);
        if (pipe->rendered[d]) cudaEventDestroy(pipe->rendered[d]);
        if (pipe->pushed[d]) cudaEventDestroy(pipe->pushed[d]);
        if (pipe->consumed[d]) cudaEventDestroy(pipe->consumed[d]);
        if (pipe->gathered_ev[d]) cudaEventDestroy(pipe->gathered_ev[d]);
    }
    cudaFree(pipe->counters); cudaFree(pipe->flushBuf);
    for (auto& e : pipe->kev) { cudaEventDestroy(e.first); cudaEventDestroy(e.second); }
    if (pipe->push) cudaStreamDestroy(pipe->push);
    if (pipe->consume) cudaStreamDestroy(pipe->consume);
    delete pipe;
}

int b200r_pipeline_create(b200r_ctx* ctx, uint32_t width, uint32_t height, uint32_t depth, uint32_t rank, uint32_t world,
                          const void* nccl_unique_id, uint32_t assemble, b200r_pipeline** out)
{
    b200r_pipeline* pipe = nullptr;
    if (!ctx || !out) return pfail(nullptr, B200R_EINVAL, "b200r_pipeline_create: NULL argument");
    *out = nullptr;
    if (!width || !height || depth < 1 || depth > (uint32_t)MAXD || world < 1 || world > (uint32_t)MAXP || rank >= world)
        return pfail(nullptr, B200R_EINVAL, "b200r_pipeline_create: bad size / depth / rank");
    if (world > 1 && !nccl_unique_id) return pfail(nullptr, B200R_EINVAL, "b200r_pipeline_create: world > 1 needs the NCCL unique id of b200r_dist_unique_id");
    if (assemble > B200R_ASSEMBLE_PUSH) return pfail(nullptr, B200R_EINVAL, "b200r_pipeline_create: unknown assembly mode");
    if (world > 1 && assemble == B200R_ASSEMBLE_PUSH && (width % 4)) return pfail(nullptr, B200R_EINVAL, "peer-push assembly needs width % 4 == 0");
    pipe = new b200r_pipeline();
    pipe->ctx = ctx; pipe->device = ctx_device(ctx); pipe->sms = ctx_sms(ctx);
    pipe->W = width; pipe->H = height; pipe->P = world; pipe->rank = rank; pipe->D = depth; pipe->mode = world > 1 ? assemble : 0;
    pipe->rps = (height + world - 1) / world;
    int rc = B200R_OK;
    auto bail = [&](int code) { b200r_pipeline_destroy(pipe); return code; };
#define PCU_B(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { pfail(nullptr, B200R_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); return bail(B200R_ECUDA); } } while (0)
    PCU_B(cudaSetDevice(pipe->device));
    int lo = 0, hi = 0;
    PCU_B(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // assembly streams at high priority: when a render CTA retires, pending push / gather CTAs get its SM before the next frame's
    // persistent CTAs do (those would hold it for a whole frame)
    PCU_B(cudaStreamCreateWithPriority(&pipe->push, cudaStreamNonBlocking, hi));
    PCU_B(cudaStreamCreateWithPriority(&pipe->consume, cudaStreamNonBlocking, hi));
    const size_t frameBytes = (size_t)width * height * 4, shardBytes = (size_t)width * pipe->rps * 4;
    for (uint32_t d = 0; d < depth; d++) {
        PCU_B(cudaStreamCreateWithFlags(&pipe->rs[d], cudaStreamNonBlocking));
        PCU_B(cudaEventCreateWithFlags(&pipe->rendered[d], cudaEventDisableTiming));
        PCU_B(cudaEventCreateWithFlags(&pipe->pushed[d], cudaEventDisableTiming));
        PCU_B(cudaEventCreateWithFlags(&pipe->consumed[d], cudaEventDisableTiming));
        PCU_B(cudaEventCreateWithFlags(&pipe->gathered_ev[d], cudaEventDisableTiming));
        PCU_B(cudaMalloc((void**)&pipe->full[d], frameBytes));
        PCU_B(cudaMemset(pipe->full[d], 0, frameBytes));
        if (world > 1) {
            PCU_B(cudaMalloc((void**)&pipe->shard[d], shardBytes));
            PCU_B(cudaMemset(pipe->shard[d], 0, shardBytes));
            if (pipe->mode == B200R_ASSEMBLE_NCCL) PCU_B(cudaMalloc((void**)&pipe->gathered[d], shardBytes * world));
        }
    }
    PCU_B(cudaMalloc((void**)&pipe->counters, 3 * MAXD * sizeof(unsigned)));
    PCU_B(cudaMemset(pipe->counters, 0, 3 * MAXD * sizeof(unsigned)));
    PCU_B(cudaDeviceSynchronize());
    if (world > 1) {
        std::string err;
        if (!load_nccl(err)) { pfail(nullptr, B200R_ESTATE, err); return bail(B200R_ESTATE); }
        Id128 id; memcpy(&id, nccl_unique_id, sizeof id);
        int r = g_nccl.CommInitRank(&pipe->nccl, (int)world, id, (int)rank);
        if (r != 0) { pfail(nullptr, B200R_ECUDA, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error")); return bail(B200R_ECUDA); }
        if (pipe->mode == B200R_ASSEMBLE_PUSH) {
            pipe->memop = getenv("B200R_PIPE_MEMOP") != nullptr;
            if (pipe->memop && !load_wait_value()) { pfail(nullptr, B200R_ESTATE, "cuStreamWaitValue32 is not available"); return bail(B200R_ESTATE); }
            // exchange where every rank's frames and counters live: one all-gather of a small blob (bootstrap only)
            PeerBlob mine; memset(&mine, 0, sizeof mine);
            mine.pid = (int)getpid(); mine.device = pipe->device;
            for (uint32_t d = 0; d <= depth; d++) {
                void* ptr = d < depth ? (void*)pipe->full[d] : (void*)pipe->counters;
                mine.raw[d] = (unsigned long long)ptr;
                PCU_B(cudaIpcGetMemHandle(&mine.handle[d], ptr));
            }
            PeerBlob* dBlobs = nullptr;
            PCU_B(cudaMalloc((void**)&dBlobs, sizeof(PeerBlob) * world));
            PCU_B(cudaMemcpy(dBlobs + rank, &mine, sizeof mine, cudaMemcpyHostToDevice));
            r = g_nccl.AllGather(dBlobs + rank, dBlobs, sizeof(PeerBlob), NCCL_INT8, pipe->nccl, pipe->push);
            if (r != 0) { cudaFree(dBlobs); pfail(nullptr, B200R_ECUDA, "ncclAllGather (bootstrap) failed"); return bail(B200R_ECUDA); }
            PCU_B(cudaStreamSynchronize(pipe->push));
            std::vector<PeerBlob> blobs(world);
            PCU_B(cudaMemcpy(blobs.data(), dBlobs, sizeof(PeerBlob) * world, cudaMemcpyDeviceToHost));
            cudaFree(dBlobs);
            for (uint32_t p = 0; p < world; p++) {
                for (uint32_t d = 0; d <= depth; d++) {
                    void* ptr = nullptr;
                    if (p == rank) ptr = (void*)mine.raw[d];
                    else if (blobs[p].pid == mine.pid) {          // another thread of this process: plain peer access
                        int can = 0;
                        PCU_B(cudaDeviceCanAccessPeer(&can, pipe->device, blobs[p].device));
                        if (!can) { pfail(nullptr, B200R_ESTATE, "peer access between the GPUs is not possible: use B200R_ASSEMBLE_NCCL"); return bail(B200R_ESTATE); }
                        cudaError_t e = cudaDeviceEnablePeerAccess(blobs[p].device, 0);
                        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { pfail(nullptr, B200R_ECUDA, cudaGetErrorString(e)); return bail(B200R_ECUDA); }
                        cudaGetLastError();
                        ptr = (void*)blobs[p].raw[d];
                    } else {
                        PCU_B(cudaIpcOpenMemHandle(&ptr, blobs[p].handle[d], cudaIpcMemLazyEnablePeerAccess));
                        pipe->opened.push_back(ptr);
                    }
                    if (d < depth) pipe->peerFull[p][d] = (uint32_t*)ptr; else pipe->peerCounters[p] = (unsigned*)ptr;
                }
            }
        }
    }
#undef PCU_B
    (void)rc;
    *out = pipe;
    return B200R_OK;
}

int b200r_pipeline_set_l2_flush(b200r_pipeline* pipe, uint64_t bytes)
{
    if (!pipe) return pfail(nullptr, B200R_EINVAL, "NULL pipeline");
    PCU(cudaSetDevice(pipe->device));
    if (pipe->flushBuf) { cudaFree(pipe->flushBuf); pipe->flushBuf = nullptr; }
    pipe->flushBytes = (size_t)bytes;
    pipe->flushMemset = getenv("B200R_PIPE_FLUSH_MEMSET") != nullptr;
    if (bytes) PCU(cudaMalloc(&pipe->flushBuf, (size_t)bytes));
    return B200R_OK;
}

int b200r_pipeline_set_prefetch(b200r_pipeline* pipe, uint32_t index, const void* dev_ptr, uint64_t bytes)
{
    if (!pipe || index >= 4) return pfail(pipe, B200R_EINVAL, "b200r_pipeline_set_prefetch: bad argument");
    pipe->prefetch[index] = (const char*)dev_ptr; pipe->prefetchBytes[index] = (size_t)bytes;
    return B200R_OK;
}

int b200r_pipeline_submit(b200r_pipeline* pipe, const b200r_frame* f, uint32_t* host_xrgb)
{
    if (!pipe || !f) return pfail(pipe, B200R_EINVAL, "b200r_pipeline_submit: NULL argument");
    if (f->width != pipe->W || f->height != pipe->H) return pfail(pipe, B200R_EINVAL, "b200r_pipeline_submit: frame size differs from the pipeline's");
    PCU(cudaSetDevice(pipe->device));
    const uint64_t i = pipe->submitted;
    const uint32_t d = (uint32_t)(i % pipe->D), P = pipe->P;
    const unsigned gen = (unsigned)(i / pipe->D);
    cudaStream_t rs = pipe->rs[d];
    b200r_frame fr = *f;
    if (P > 1) { fr.row_first = pipe->rank; fr.row_step = P; }      // (world 1: the caller's own row selection is kept - experiments)
    const bool mlaa = (fr.flags & B200R_F_MLAA) != 0;
    fr.flags &= ~(uint32_t)B200R_F_MLAA;                         // the filter needs neighbouring rows: it runs on the assembled frame
    // ---- render stream of the slot: the slot's previous frame must have left the buffers this frame writes
    if (i >= pipe->D) PCU(cudaStreamWaitEvent(rs, P > 1 ? pipe->pushed[d] : pipe->consumed[d], 0));
    if (pipe->flushBuf) {
        if (pipe->flushMemset) PCU(cudaMemsetAsync(pipe->flushBuf, 0, pipe->flushBytes, rs));      // measurement aid: evict the L2 before the frame ...
        else { l2_flush_kernel<<<pipe->sms * 2, 128, 0, rs>>>((uint4*)pipe->flushBuf, pipe->flushBytes / 16); PCU(cudaGetLastError()); pipe->launches += 1; }
        for (int k = 0; k < 4; k++)                                               // ... and pull the scene back in bulk, not miss by miss
            if (pipe->prefetch[k]) { l2_prefetch_kernel<<<pipe->sms, 256, 0, rs>>>(pipe->prefetch[k], pipe->prefetchBytes[k]); pipe->launches += 1; }
    }
    uint32_t* target = P > 1 ? pipe->shard[d] : pipe->full[d];
    if (pipe->timing) {
        cudaEvent_t a, b;
        PCU(cudaEventCreate(&a)); PCU(cudaEventCreate(&b));
        pipe->kev.push_back({a, b});
        PCU(cudaEventRecord(a, rs));
    }
    int rc = b200r_render_device_slot(pipe->ctx, &fr, target, rs, d);
    if (rc) return pfail(pipe, rc, b200r_last_error(pipe->ctx));
    if (pipe->timing) PCU(cudaEventRecord(pipe->kev.back().second, rs));
    uint32_t n = 0; b200r_last_launches(pipe->ctx, &n); pipe->launches += n;
    PCU(cudaEventRecord(pipe->rendered[d], rs));
    // ---- assembly
    cudaStream_t cs = pipe->consume;
    if (P > 1 && pipe->mode == B200R_ASSEMBLE_NCCL) {
        PCU(cudaStreamWaitEvent(pipe->push, pipe->rendered[d], 0));
        if (i >= pipe->D) PCU(cudaStreamWaitEvent(pipe->push, pipe->consumed[d], 0));      // gathered[d] / full[d] still being read
        const size_t shardBytes = (size_t)pipe->W * pipe->rps * 4;
        PNCCL(g_nccl.AllGather(pipe->shard[d], pipe->gathered[d], shardBytes, NCCL_INT8, pipe->nccl, pipe->push));
        PCU(cudaEventRecord(pipe->pushed[d], pipe->push));                                    // the shard may be rendered into again
        PCU(launch_deinterleave(pipe->gathered[d], pipe->full[d], pipe->W, pipe->H, P, pipe->sms, pipe->push));
        pipe->launches += 1;
        PCU(cudaEventRecord(pipe->gathered_ev[d], pipe->push));
        PCU(cudaStreamWaitEvent(cs, pipe->gathered_ev[d], 0));
    } else if (P > 1) {
        unsigned* cnt = pipe->counters;
        PCU(cudaStreamWaitEvent(pipe->push, pipe->rendered[d], 0));
        // every rank must have consumed the slot's previous frame before anyone overwrites it: P acknowledgements per generation
        if (gen > 0) {
            if (pipe->memop) { if (g_waitValue(pipe->push, (unsigned long long)(cnt + MAXD + d), P * gen, WAIT_GEQ) != 0) return pfail(pipe, B200R_ECUDA, "cuStreamWaitValue32 (acknowledgements) failed"); }
            else { wait_counter_kernel<<<1, 32, 0, pipe->push>>>(cnt + MAXD + d, P * gen); PCU(cudaGetLastError()); }
        }
        PushArgs a;
        for (uint32_t p = 0; p < P; p++) { a.dst[p] = pipe->peerFull[p][d]; a.arrive[p] = pipe->peerCounters[p] + d; }
        const int nRows = (int)((pipe->H - pipe->rank + P - 1) / P);
        push_rows_kernel<<<pipe->sms, 256, 0, pipe->push>>>(pipe->shard[d], a, (int)P, (int)pipe->rank, (int)pipe->W, nRows, cnt + 2 * MAXD + d);
        PCU(cudaGetLastError());
        pipe->launches += 1;
        PCU(cudaEventRecord(pipe->pushed[d], pipe->push));
        if (pipe->memop) { if (g_waitValue(cs, (unsigned long long)(cnt + d), P * (gen + 1), WAIT_GEQ) != 0) return pfail(pipe, B200R_ECUDA, "cuStreamWaitValue32 (arrivals) failed"); }
        else { wait_counter_kernel<<<1, 32, 0, cs>>>(cnt + d, P * (gen + 1)); PCU(cudaGetLastError()); }
    } else {
        PCU(cudaStreamWaitEvent(cs, pipe->rendered[d], 0));
    }
    // ---- consumer stream: post filter on the assembled frame, copy-out, release of the slot
    if (mlaa) {
        rc = b200r_mlaa_device(pipe->ctx, pipe->full[d], pipe->W, pipe->H, cs);
        if (rc) return pfail(pipe, rc, b200r_last_error(pipe->ctx));
    }
    if (host_xrgb) PCU(cudaMemcpyAsync(host_xrgb, pipe->full[d], (size_t)pipe->W * pipe->H * 4, cudaMemcpyDeviceToHost, cs));
    if (P > 1 && pipe->mode == B200R_ASSEMBLE_PUSH) {
        SignalArgs s;
        for (uint32_t p = 0; p < P; p++) s.word[p] = pipe->peerCounters[p] + MAXD + d;
        signal_peers_kernel<<<1, 32, 0, cs>>>(s, (int)P);
        PCU(cudaGetLastError());
    }
    PCU(cudaEventRecord(pipe->consumed[d], cs));
    pipe->submitted++;
    return B200R_OK;
}

int b200r_pipeline_slot_frame(b200r_pipeline* pipe, uint32_t slot, void** dev_xrgb)
{
    if (!pipe || !dev_xrgb || slot >= pipe->D) return pfail(pipe, B200R_EINVAL, "b200r_pipeline_slot_frame: bad argument");
    *dev_xrgb = pipe->full[slot];
    return B200R_OK;
}

int b200r_pipeline_fence(b200r_pipeline* pipe, void* cuda_stream, int pipeline_waits)
{
    if (!pipe || !cuda_stream) return pfail(pipe, B200R_EINVAL, "b200r_pipeline_fence: NULL argument");
    PCU(cudaSetDevice(pipe->device));
    cudaStream_t s = (cudaStream_t)cuda_stream;
    cudaEvent_t ev;
    PCU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    if (pipeline_waits) {            // nothing submitted from now on starts before the stream's current tail
        PCU(cudaEventRecord(ev, s));
        for (uint32_t d = 0; d < pipe->D; d++) PCU(cudaStreamWaitEvent(pipe->rs[d], ev, 0));
        PCU(cudaStreamWaitEvent(pipe->push, ev, 0));
        PCU(cudaStreamWaitEvent(pipe->consume, ev, 0));
    } else {                          // the stream waits for everything submitted so far
        for (uint32_t d = 0; d < pipe->D; d++) { PCU(cudaEventRecord(ev, pipe->rs[d])); PCU(cudaStreamWaitEvent(s, ev, 0)); }
        PCU(cudaEventRecord(ev, pipe->push)); PCU(cudaStreamWaitEvent(s, ev, 0));
        PCU(cudaEventRecord(ev, pipe->consume)); PCU(cudaStreamWaitEvent(s, ev, 0));
    }
    PCU(cudaEventDestroy(ev));
    return B200R_OK;
}

int b200r_pipeline_drain(b200r_pipeline* pipe)
{
    if (!pipe) return pfail(nullptr, B200R_EINVAL, "NULL pipeline");
    PCU(cudaSetDevice(pipe->device));
    for (uint32_t d = 0; d < pipe->D; d++) PCU(cudaStreamSynchronize(pipe->rs[d]));
    PCU(cudaStreamSynchronize(pipe->push));
    PCU(cudaStreamSynchronize(pipe->consume));
    return B200R_OK;
}

int b200r_pipeline_set_timing(b200r_pipeline* pipe, int enabled)
{
    if (!pipe) return pfail(nullptr, B200R_EINVAL, "NULL pipeline");
    pipe->timing = enabled != 0;
    return B200R_OK;
}

int b200r_pipeline_kernel_ms(b200r_pipeline* pipe, double* sum_ms, uint32_t* n_frames)
{
    if (!pipe || !sum_ms || !n_frames) return pfail(pipe, B200R_EINVAL, "NULL argument");
    PCU(cudaSetDevice(pipe->device));
    double sum = 0.0; uint32_t n = 0;
    for (auto& e : pipe->kev) {
        float ms = 0.f;
        PCU(cudaEventSynchronize(e.second));
        PCU(cudaEventElapsedTime(&ms, e.first, e.second));
        sum += ms; n++;
        cudaEventDestroy(e.first); cudaEventDestroy(e.second);
    }
    pipe->kev.clear();
    *sum_ms = sum; *n_frames = n;
    return B200R_OK;
}

int b200r_pipeline_launches(b200r_pipeline* pipe, uint32_t* n, int reset)
{
    if (!pipe || !n) return pfail(pipe, B200R_EINVAL, "NULL argument");
    *n = pipe->launches;
    if (reset) pipe->launches = 0;
    return B200R_OK;
}

}  // extern "C"
