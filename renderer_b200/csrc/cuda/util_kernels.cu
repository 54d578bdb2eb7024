// util_kernels.cu — small framebuffer helpers of the multi-GPU path.
#include "device_types.cuh"
#include "rt_kernels.cuh"

namespace b200r {

namespace {
// After the NCCL all-gather, shard s (of P) holds rows s, s+P, s+2P, ... packed; every shard is padded to
// rowsPerShard = ceil(H/P) rows. Scatter back into scan order with 16-byte accesses (W % 4 == 0) or words.
__global__ void deinterleave_kernel(const uint32_t* __restrict__ gathered, uint32_t* __restrict__ frame,
                                    uint32_t W, uint32_t H, uint32_t P, uint32_t rowsPerShard)
{
    const size_t total = (size_t)W * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / W), x = (uint32_t)(i % W);
        const uint32_t s = y % P, r = y / P;
        frame[i] = gathered[((size_t)s * rowsPerShard + r) * W + x];
    }
}
}  // namespace

cudaError_t launch_deinterleave(const uint32_t* gathered, uint32_t* frame, uint32_t W, uint32_t H, uint32_t P,
                                int numSMs, cudaStream_t stream)
{
    const uint32_t rowsPerShard = (H + P - 1) / P;
    deinterleave_kernel<<<numSMs * 4, 256, 0, stream>>>(gathered, frame, W, H, P, rowsPerShard);
    return cudaGetLastError();
}

}  // namespace b200r
