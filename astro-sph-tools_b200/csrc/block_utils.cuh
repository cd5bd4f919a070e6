// block_utils.cuh -- CTA-wide sum and exclusive prefix of one uint32 per thread (warp shuffles + shared memory)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ast {

__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t *smem /* >= 32 */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    uint32_t t = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0u;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) smem[0] = t;
    }
    __syncthreads();
    t = smem[0];
    __syncthreads();
    return t;
}

__device__ __forceinline__ uint64_t block_sum_u64(uint64_t v, uint64_t *smem /* >= 32 */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    uint64_t t = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0ull;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) smem[0] = t;
    }
    __syncthreads();
    t = smem[0];
    __syncthreads();
    return t;
}

// exclusive prefix of v over the block (thread order), total in *total
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t *smem /* >= 33 */, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < (int)(blockDim.x >> 5)) ? smem[lane] : 0u;
        uint32_t winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    uint32_t r = smem[warp] + inc - v;
    *total = smem[32];
    __syncthreads();
    return r;
}

}  // namespace ast
