// work_items.cuh -- splitting long tile / brick lists into work items (shared by project2d.cu and grid3d.cu)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace ast {

// K5b: work items.  A tile whose list is longer than seg_target entries is split into ceil(cnt / seg_target) segments, each
// accumulated by its own CTA (partial sums then go to the map with float64 atomics instead of a plain read-modify-write).
// This keeps every SM busy when few tiles hold most of the pairs: clustered particle sets, the slabs that index-sharded
// ranks and host batches deposit, and the tail of the last wave.
__device__ __forceinline__ uint32_t tile_segments(uint32_t cnt, uint32_t n_huge, uint32_t seg_target)
{
    if (cnt + n_huge == 0) return 0u;
    return cnt <= seg_target ? 1u : (cnt + seg_target - 1) / seg_target;
}
static __global__ void tile_segments_kernel(const uint32_t *__restrict__ tbeg, const uint32_t *__restrict__ tend, uint32_t n_huge,
                                     uint32_t seg_target, int ntiles, uint32_t *__restrict__ seg_off)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    seg_off[t] = t < ntiles ? tile_segments(tend[t] - tbeg[t], n_huge, seg_target) : 0u;
}

struct TileWork {
    int tile;
    uint32_t beg, cnt, n_huge;    // the tile's whole list: cnt sorted pairs from beg, then the n_huge large-h entries
    uint32_t first, step;         // this CTA takes the 32-entry batches starting at first, first + step, ...
    bool atomic_out;
};
// work item of this CTA (uniform over the CTA); false: nothing to do.  The segments of a split tile INTERLEAVE its batches
// (segment s of n takes batches s, s + n, ...): the list is in particle order, i.e. spatially ordered, so a contiguous piece
// of it would load the 8 warps (sub-tiles) of the CTA very unevenly (measured: +4.5 % on the kernel), a strided sample does not.
template <class A>
__device__ __forceinline__ bool resolve_work(const A &a, TileWork &w)
{
    const uint32_t b = blockIdx.x;
    if (b >= a.seg_off[a.ntiles]) return false;
    int lo = 0, hi = a.ntiles - 1;                 // first tile t with seg_off[t + 1] > b
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a.seg_off[mid + 1] > b) hi = mid; else lo = mid + 1;
    }
    const uint32_t s0 = a.seg_off[lo], nseg = a.seg_off[lo + 1] - s0, seg = b - s0;
    w.tile = lo;
    w.beg = a.tbeg[lo];
    w.cnt = a.tend[lo] - w.beg;
    w.n_huge = a.n_huge;
    w.first = 32u * seg;
    w.step = 32u * nseg;
    w.atomic_out = nseg > 1;
    return true;
}


// entries per segment so that the lists give about `waves` waves of work items over `slots` resident CTAs (AST_SEG_WAVES
// overrides the default 32: measured at config 2 of the 2-D path, 35.1 ms un-split, 34.4 / 34.0 / 33.9 ms at 16 / 32 / 128 waves)
inline uint32_t segment_target(int64_t n_pairs, int64_t slots)
{
    static const int waves_env = [] { const char *e = getenv("AST_SEG_WAVES"); const int w = e ? atoi(e) : 32; return w < 1 ? 1 : w; }();
    const int waves = waves_env;
    int64_t target = (n_pairs / (slots * waves) + 31) & ~(int64_t)31;
    return (uint32_t)(target < 1024 ? 1024 : (target > 65536 ? 65536 : target));
}

}  // namespace ast
