// presort.cuh -- optional spatial pre-ordering of the particles (shared by project2d.cu and grid3d.cu).
//
// The pipeline is correct for any particle order, but it is fast only when neighbouring particles in memory are neighbours in
// space: threads of a warp then take the same branches (similar h, same class), the (tile, particle) pairs come out almost
// sorted, and the record gathers of a tile list hit the same sectors.  Snapshot files usually are ordered like that (cell or
// Peano-Hilbert order); a randomly ordered set is not.  Measured on B200 (benchmarks/order_probe.py): config 2 in lattice
// order 34.2 ms, the same particles in random order 42.5 ms (sort 1.9 -> 5.4 ms, accumulate 29.2 -> 33.1 ms); config 4 (the
// S2 recipe shuffles its particles) 73.6 ms, ordered by brick 53.4 ms (binning 10.1 -> 3.7, emit 9.2 -> 3.7, sort 13.3 -> 5.3 ms).
// So: (1) a cheap SAMPLE decides whether the input is incoherent -- consecutive particles of every 64th block of 256 further
// than two cells apart; (2) if so, key = cell of the particle's own position -> stable radix sort -> gather of positions, h
// and weights into that order, and the pipeline runs on the copies.  Only the order of float additions changes.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "scan_sort.cuh"

namespace ast {

struct OrderGrid {
    double lo[3], inv_cell[3];      // cell index along axis c = floor((x[col[c]] - lo[c]) * inv_cell[c]), clamped to [0, n[c])
    int n[3], col[3];
    int dims;                       // 2 (projection: the two in-plane columns) or 3
};

__device__ __forceinline__ void order_cell(const OrderGrid &g, const double *__restrict__ pos, int64_t i, int c[3])
{
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c[k] = 0;
        if (k < g.dims) {
            const double t = floor((pos[3 * i + g.col[k]] - g.lo[k]) * g.inv_cell[k]);
            c[k] = t > 0.0 ? (t < (double)(g.n[k] - 1) ? (int)t : g.n[k] - 1) : 0;      // NaN -> 0
        }
    }
}

// counts[0] += sampled consecutive pairs, counts[1] += those more than two cells apart along some axis
static __global__ void __launch_bounds__(256) order_sample_kernel(const double *__restrict__ pos, int64_t n, OrderGrid g, int stride,
                                                                  unsigned long long *__restrict__ counts)
{
    const int64_t i = ((int64_t)blockIdx.x * stride) * 256 + threadIdx.x;
    bool ok = i + 1 < n, far = false;
    if (ok) {
        int a[3], b[3];
        order_cell(g, pos, i, a);
        order_cell(g, pos, i + 1, b);
        far = abs(a[0] - b[0]) > 2 || abs(a[1] - b[1]) > 2 || abs(a[2] - b[2]) > 2;
    }
    const unsigned bo = __ballot_sync(0xffffffffu, ok), bf = __ballot_sync(0xffffffffu, far);
    if ((threadIdx.x & 31) == 0 && bo) {
        atomicAdd(counts, (unsigned long long)__popc(bo));
        if (bf) atomicAdd(counts + 1, (unsigned long long)__popc(bf));
    }
}

static __global__ void __launch_bounds__(256) order_key_kernel(const double *__restrict__ pos, int64_t n, OrderGrid g,
                                                               uint64_t *__restrict__ elems)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int c[3];
    order_cell(g, pos, i, c);
    const uint32_t key = (uint32_t)((c[0] * g.n[1] + c[1]) * (g.dims == 3 ? g.n[2] : 1) + (g.dims == 3 ? c[2] : 0));
    elems[i] = ((uint64_t)key << 32) | (uint64_t)(uint32_t)i;
}

template <int NP>
static __global__ void __launch_bounds__(256) order_gather_kernel(const uint64_t *__restrict__ sorted, int64_t n, const double *__restrict__ pos,
                                                                  const double *__restrict__ h, const double *__restrict__ p0,
                                                                  const double *__restrict__ p1, double *__restrict__ spos,
                                                                  double *__restrict__ sh, double *__restrict__ sp0, double *__restrict__ sp1)
{
    const int64_t s = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (s >= n) return;
    const int64_t i = (int64_t)(uint32_t)sorted[s];
    spos[3 * s] = pos[3 * i]; spos[3 * s + 1] = pos[3 * i + 1]; spos[3 * s + 2] = pos[3 * i + 2];
    sh[s] = h[i];
    sp0[s] = p0[i];
    if (NP > 1) sp1[s] = p1[i];
}

struct PresortBuffers {
    uint64_t *ka, *kb;              // n each
    void *sort_ws;                  // sort_workspace_bytes(n)
    double *spos, *sh, *sprop[2];   // 3n, n, n, n
    unsigned long long *counts;     // 2
};

inline int presort_cell_count(const OrderGrid &g) { return g.n[0] * g.n[1] * (g.dims == 3 ? g.n[2] : 1); }

// Decides (mode 1 = by sample, mode 2 = always) and, if the input is to be reordered, fills the buffers and returns 1 through
// *done (the caller then uses B.spos / B.sh / B.sprop).  Synchronises the stream once in mode 1.
inline cudaError_t presort_particles(int mode, const OrderGrid &g, const double *pos, const double *h, const double *const *prop, int n_prop,
                                     int64_t n, const PresortBuffers &B, cudaStream_t s, int *done, int *launches)
{
    *done = 0;
    if (mode == 0 || n < 65536) return cudaSuccess;                 // (small sets: the pipeline is launch-bound anyway)
    if (mode == 1) {
        const int stride = 64;
        const int64_t nb = ((n + 255) / 256 + stride - 1) / stride;
        unsigned long long c[2] = { 0, 0 };
        cudaError_t e = cudaMemsetAsync(B.counts, 0, 2 * sizeof(unsigned long long), s);
        if (e != cudaSuccess) return e;
        order_sample_kernel<<<(unsigned)nb, 256, 0, s>>>(pos, n, g, stride, B.counts);
        e = cudaMemcpyAsync(c, B.counts, sizeof c, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return e;
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return e;
        if (launches) *launches += 1;
        if (c[0] == 0 || c[1] * 4 < c[0]) return cudaSuccess;       // fewer than a quarter of the neighbours are far apart: coherent
    }
    const unsigned nbk = (unsigned)((n + 255) / 256);
    order_key_kernel<<<nbk, 256, 0, s>>>(pos, n, g, B.ka);
    int in_b = 0, nl = 0, bits = 0;
    while (bits < 31 && (1ll << bits) < (long long)presort_cell_count(g)) ++bits;
    cudaError_t e = radix_sort_u64(B.ka, B.kb, n, 32, bits, B.sort_ws, s, &in_b, &nl);
    if (e != cudaSuccess) return e;
    const uint64_t *sorted = in_b ? B.kb : B.ka;
    if (n_prop > 1) order_gather_kernel<2><<<nbk, 256, 0, s>>>(sorted, n, pos, h, prop[0], prop[1], B.spos, B.sh, B.sprop[0], B.sprop[1]);
    else order_gather_kernel<1><<<nbk, 256, 0, s>>>(sorted, n, pos, h, prop[0], nullptr, B.spos, B.sh, B.sprop[0], nullptr);
    if (launches) *launches += 2 + nl;
    *done = 1;
    return cudaGetLastError();
}

}  // namespace ast
