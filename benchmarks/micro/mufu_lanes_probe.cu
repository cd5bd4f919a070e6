// mufu_lanes_probe.cu -- does a warp-wide MUFU.SQRT cost less when only some lanes are active?  (B200)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float fsqrt(float x) { float r; asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__global__ void __launch_bounds__(256) k(float *out, int iters, unsigned mask)
{
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 1.0f + threadIdx.x + i;
    const bool on = (mask >> (threadIdx.x & 31)) & 1u;
    if (on) {
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fsqrt(v[i]) + 1.0f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int grid = pr.multiProcessorCount * 8, iters = 20000;
    float *out; cudaMalloc(&out, sizeof(float) * grid * 256);
    const unsigned masks[] = {0xffffffffu, 0x0000ffffu, 0x000000ffu, 0x0000000fu, 0x00000001u, 0x55555555u, 0x11111111u, 0x00ff00ffu, 0x0f0f0f0fu};
    for (unsigned m : masks) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        k<<<grid, 256>>>(out, iters, m); cudaDeviceSynchronize();
        cudaEventRecord(a); k<<<grid, 256>>>(out, iters, m); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        printf("mask %08x: %8.3f ms  %6.2f SMSP-cycles per warp MUFU (+FADD)\n", m, ms, ms * 1e-3 * khz * 1e3 / iters / 16.0 / 8.0);
    }
    return cudaDeviceSynchronize() != cudaSuccess;
}
