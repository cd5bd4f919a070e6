// red_patch_probe.cu -- throughput of float64 reductions (RED.ADD.F64) to small patches of a 4096^2 map, the access pattern a
// warp-cooperative direct deposit of medium-footprint particles would have: one warp per particle, the lanes cover a w x w
// patch (row-major in the fast axis) around the particle, ~78 % of the lanes inside the disc add one double each.
// Particles in lattice order (consecutive warps hit neighbouring patches), 4 patches per lattice cell of 8 pixels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o red_patch_probe red_patch_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void patch_kernel(double *map, int npix, long long n_particles, int w, int per_row)
{
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_particles) return;
    // particle (i, j, k) of an n^3 lattice in C order projected along k: patch origin from (i, j) + a hash of k
    const long long cell = warp / per_row;                   // (i, j) column
    const unsigned hsh = (unsigned)(warp * 2654435761u) >> 16;
    const int n = npix / 8;
    const int ci = (int)(cell / n), cj = (int)(cell % n);
    const int x0 = (ci * 8 + (hsh & 7) - w / 2 + npix) % (npix - w), y0 = (cj * 8 + ((hsh >> 3) & 7) - w / 2 + npix) % (npix - w);
    const float r2max = 0.25f * w * w;
    for (int p = lane; p < w * w; p += 32) {
        const int dx = p / w, dy = p - dx * w;
        const float fx = dx - 0.5f * w + 0.5f, fy = dy - 0.5f * w + 0.5f;
        if (fx * fx + fy * fy < r2max) atomicAdd(map + (size_t)(x0 + dx) * npix + y0 + dy, 1.0e-3 * (double)(r2max - fx * fx - fy * fy));
    }
}

int main(int argc, char **argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 512, w = argc > 2 ? atoi(argv[2]) : 9;
    const int npix = 8 * n;
    const long long N = (long long)n * n * n;
    double *map;
    cudaMalloc(&map, sizeof(double) * (size_t)npix * npix);
    cudaMemset(map, 0, sizeof(double) * (size_t)npix * npix);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const long long threads = N * 32;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        patch_kernel<<<(unsigned)((threads + 255) / 256), 256>>>(map, npix, N, w, n);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double useful = 0; for (int p = 0; p < w * w; ++p) { int dx = p / w, dy = p % w; float fx = dx - 0.5f * w + 0.5f, fy = dy - 0.5f * w + 0.5f; if (fx * fx + fy * fy < 0.25f * w * w) useful += 1; }
        printf("n %d w %d rep %d: %.3f ms, %.3e patches/s, %.3e RED.F64/s (%g per patch) err=%s\n", n, w, rep, ms, N / (ms * 1e-3), N * useful / (ms * 1e-3), useful, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
