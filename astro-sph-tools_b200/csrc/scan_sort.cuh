// scan_sort.cuh -- hand-written device exclusive scan and stable LSD radix sort of 64-bit elements.
//
// The sort orders (tile key << 32 | particle) pairs by tile key.  It must be STABLE: within a tile the
// particles stay in emit order, which fixes the float summation order per pixel and makes the sorted
// permutation comparable bit-for-bit with the CPU oracle (oracle/sph_oracle.c orc_sort_pairs_stable).
//
// One pass = histogram (per block of 8192 elements) -> exclusive scan of the digit-major table ->
// scatter.  Inside the scatter kernel every warp owns a contiguous 1024-element slice of the block's
// chunk and ranks its elements with match.any, so the order of equal digits is the input order.
// HBM traffic per pass: 8 B read (histogram) + 8 B read (L2-resident re-read) + 8 B write per element.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ast {

constexpr int kScanThreads = 1024;
constexpr int kScanItems = 4;

// Single-block exclusive scan, in place.  n is small here (block sums, histogram tables).
template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_exclusive_kernel(T *data, int64_t n, T *total_out)
{
    __shared__ T warp_sum[32];
    __shared__ T carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += (int64_t)kScanThreads * kScanItems) {
        int64_t i0 = base + (int64_t)tid * kScanItems;
        T v[kScanItems];
        T s = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = (i0 + k < n) ? data[i0 + k] : (T)0;
            s += v[k];
        }
        T inc = s;                                   // inclusive scan of the thread sums inside the warp
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            T w = warp_sum[lane];
            T winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                T t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_sum[lane] = winc - w;               // exclusive offset of each warp
        }
        __syncthreads();
        T carry = carry_s;
        T excl = carry + warp_sum[warp] + (inc - s);
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (i0 + k < n) data[i0 + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (tid == kScanThreads - 1) carry_s = excl; // last thread holds carry + sum of this batch
        __syncthreads();
    }
    if (tid == 0 && total_out) *total_out = carry_s;
}

// two independent arrays scanned by one launch (block 0 -> a, block 1 -> b)
template <class T>
__global__ void __launch_bounds__(kScanThreads) scan2_exclusive_kernel(T *a, T *b, int64_t n)
{
    T *data = blockIdx.x == 0 ? a : b;
    __shared__ T warp_sum[32];
    __shared__ T carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += (int64_t)kScanThreads * kScanItems) {
        int64_t i0 = base + (int64_t)tid * kScanItems;
        T v[kScanItems];
        T s = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            v[k] = (i0 + k < n) ? data[i0 + k] : (T)0;
            s += v[k];
        }
        T inc = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            T w = warp_sum[lane];
            T winc = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                T t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            warp_sum[lane] = winc - w;
        }
        __syncthreads();
        T excl = carry_s + warp_sum[warp] + (inc - s);
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (i0 + k < n) data[i0 + k] = excl;
            excl += v[k];
        }
        __syncthreads();
        if (tid == kScanThreads - 1) carry_s = excl;
        __syncthreads();
    }
}

// ---- multi-block exclusive scan: per-block sums -> scan of the sums (single block) -> per-block scan + base.
// One block handles kScanTile consecutive elements.
constexpr int kScanTile = kScanThreads * kScanItems;      // 4096

template <class T>
__device__ __forceinline__ T block_reduce_sum(T v, T *smem32)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) smem32[warp] = v;
    __syncthreads();
    T t = (threadIdx.x < (blockDim.x >> 5)) ? smem32[threadIdx.x] : (T)0;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;                                            // valid in thread 0
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_block_sums_kernel(const T *__restrict__ data, int64_t n, T *__restrict__ sums)
{
    __shared__ T sm[32];
    const int64_t i0 = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) s += (i0 + k < n) ? data[i0 + k] : (T)0;
    T t = block_reduce_sum<T>(s, sm);
    if (threadIdx.x == 0) sums[blockIdx.x] = t;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(T *__restrict__ data, int64_t n, const T *__restrict__ base)
{
    __shared__ T warp_sum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * kScanTile + (int64_t)tid * kScanItems;
    T v[kScanItems];
    T s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (i0 + k < n) ? data[i0 + k] : (T)0;
        s += v[k];
    }
    T inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = warp_sum[lane];
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sum[lane] = winc - w;
    }
    __syncthreads();
    T excl = base[blockIdx.x] + warp_sum[warp] + (inc - s);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (i0 + k < n) data[i0 + k] = excl;
        excl += v[k];
    }
}

inline int64_t scan_num_blocks(int64_t n) { return (n + kScanTile - 1) / kScanTile; }
template <class T>
inline size_t scan_workspace_bytes(int64_t n) { return (size_t)(scan_num_blocks(n > 0 ? n : 1) + 1) * sizeof(T); }

// exclusive scan of data[0..n) in place; tmp: scan_workspace_bytes<T>(n); total_out (device, nullable) = sum of all
template <class T>
inline cudaError_t scan_exclusive(T *data, int64_t n, T *tmp, T *total_out, cudaStream_t s, int *launches = nullptr)
{
    if (n <= 0) return cudaSuccess;
    if (n <= 2 * kScanTile) {
        scan_exclusive_kernel<T><<<1, kScanThreads, 0, s>>>(data, n, total_out);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    const int64_t nb = scan_num_blocks(n);
    scan_block_sums_kernel<T><<<(unsigned)nb, kScanThreads, 0, s>>>(data, n, tmp);
    scan_exclusive_kernel<T><<<1, kScanThreads, 0, s>>>(tmp, nb, total_out);
    scan_apply_kernel<T><<<(unsigned)nb, kScanThreads, 0, s>>>(data, n, tmp);
    if (launches) *launches += 3;
    return cudaGetLastError();
}


// ---- two arrays, multi-block: blockIdx.y selects the array.  sums: 2 * scan_num_blocks(n) entries.
template <class T>
__global__ void __launch_bounds__(kScanThreads) scan2_block_sums_kernel(const T *__restrict__ a, const T *__restrict__ b, int64_t n,
                                                                        T *__restrict__ sums)
{
    __shared__ T sm[32];
    const T *data = blockIdx.y == 0 ? a : b;
    const int64_t i0 = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) s += (i0 + k < n) ? data[i0 + k] : (T)0;
    T t = block_reduce_sum<T>(s, sm);
    if (threadIdx.x == 0) sums[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
}

template <class T>
__global__ void __launch_bounds__(kScanThreads) scan2_apply_kernel(T *__restrict__ a, T *__restrict__ b, int64_t n,
                                                                   const T *__restrict__ base)
{
    __shared__ T warp_sum[32];
    T *data = blockIdx.y == 0 ? a : b;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i0 = (int64_t)blockIdx.x * kScanTile + (int64_t)tid * kScanItems;
    T v[kScanItems];
    T s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = (i0 + k < n) ? data[i0 + k] : (T)0;
        s += v[k];
    }
    T inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = warp_sum[lane];
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        warp_sum[lane] = winc - w;
    }
    __syncthreads();
    T excl = base[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] + warp_sum[warp] + (inc - s);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (i0 + k < n) data[i0 + k] = excl;
        excl += v[k];
    }
}

// exclusive scan of a[0..n) and b[0..n) in place.  tmp: 2 * scan_num_blocks(n) elements.
template <class T>
inline cudaError_t scan2_exclusive(T *a, T *b, int64_t n, T *tmp, cudaStream_t s, int *launches = nullptr)
{
    if (n <= 0) return cudaSuccess;
    const int64_t nb = scan_num_blocks(n);
    if (nb <= 2) {
        scan2_exclusive_kernel<T><<<2, kScanThreads, 0, s>>>(a, b, n);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    scan2_block_sums_kernel<T><<<dim3((unsigned)nb, 2), kScanThreads, 0, s>>>(a, b, n, tmp);
    scan2_exclusive_kernel<T><<<2, kScanThreads, 0, s>>>(tmp, tmp + nb, nb);
    scan2_apply_kernel<T><<<dim3((unsigned)nb, 2), kScanThreads, 0, s>>>(a, b, n, tmp);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
constexpr int kSortWarps = 8;
constexpr int kSortWarpItems = 1024;
constexpr int kSortBlockItems = kSortWarps * kSortWarpItems;   // 8192
constexpr int kSortThreads = kSortWarps * 32;                  // 256
constexpr int kRadixMaxBits = 9;                               // digits of up to 9 bits (512 buckets): 18 key bits = 2 passes
constexpr int kRadixMax = 1 << kRadixMaxBits;

inline int64_t sort_num_blocks(int64_t n) { return (n + kSortBlockItems - 1) / kSortBlockItems; }
inline size_t sort_table_entries(int64_t n) { return (size_t)sort_num_blocks(n > 0 ? n : 1) * kRadixMax; }
inline size_t sort_workspace_bytes(int64_t n)
{
    size_t t = (sort_table_entries(n) + 64) * sizeof(uint32_t);
    t = (t + 255) / 256 * 256;
    return t + scan_workspace_bytes<uint32_t>((int64_t)sort_table_entries(n));
}

// RADIX = number of buckets of this pass (256 or 512); thread t owns buckets t, t + 256, ...
template <int RADIX>
static __global__ void __launch_bounds__(kSortThreads)
sort_hist_kernel(const uint64_t *__restrict__ in, int64_t n, int shift, uint32_t mask, uint32_t *__restrict__ table,
                 int64_t nblocks)
{
    __shared__ uint32_t hist[RADIX];
    const int tid = threadIdx.x;
#pragma unroll
    for (int d = tid; d < RADIX; d += kSortThreads) hist[d] = 0;
    __syncthreads();
    const int64_t b = blockIdx.x;
    const int64_t beg = b * kSortBlockItems;
    const int64_t end = beg + kSortBlockItems < n ? beg + kSortBlockItems : n;
    for (int64_t i0 = beg; i0 < end; i0 += kSortThreads) {     // uniform trip count for the whole block
        int64_t i = i0 + tid;
        uint32_t d = i < end ? ((uint32_t)(in[i] >> shift) & mask) : 0xffffffffu;
        // aggregate equal digits inside the warp before touching shared memory
        unsigned peers = __match_any_sync(0xffffffffu, d);
        if (d != 0xffffffffu && (int)(__ffs(peers) - 1) == (tid & 31)) atomicAdd(&hist[d], (uint32_t)__popc(peers));
    }
    __syncthreads();
#pragma unroll
    for (int d = tid; d < RADIX; d += kSortThreads) table[(int64_t)d * nblocks + b] = hist[d];
}

template <int RADIX>
static __global__ void __launch_bounds__(kSortThreads)
sort_scatter_kernel(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, int64_t n, int shift, uint32_t mask,
                    const uint32_t *__restrict__ table, int64_t nblocks)
{
    __shared__ uint32_t wbase[kSortWarps][RADIX];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kSortWarps * RADIX; i += kSortThreads) (&wbase[0][0])[i] = 0;
    __syncthreads();
    const int64_t b = blockIdx.x;
    const int64_t wbeg = b * kSortBlockItems + (int64_t)warp * kSortWarpItems;
    // pass 1: per-warp digit counts of the warp's own contiguous slice
    for (int it = 0; it < kSortWarpItems / 32; ++it) {
        int64_t i = wbeg + it * 32 + lane;
        uint32_t d = i < n ? ((uint32_t)(in[i] >> shift) & mask) : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        if (d != 0xffffffffu && (__ffs(peers) - 1) == lane) wbase[warp][d] += (uint32_t)__popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // thread t owns digits t, t + 256, ...: global base of this block for the digit, then running offsets per warp
#pragma unroll
    for (int d = tid; d < RADIX; d += kSortThreads) {
        uint32_t run = table[(int64_t)d * nblocks + b];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            uint32_t c = wbase[w][d];
            wbase[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    // pass 2: stable scatter
    for (int it = 0; it < kSortWarpItems / 32; ++it) {
        int64_t i = wbeg + it * 32 + lane;
        uint64_t e = i < n ? in[i] : 0ull;
        uint32_t d = i < n ? ((uint32_t)(e >> shift) & mask) : 0xffffffffu;
        unsigned peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
        uint32_t base = 0;
        if (d != 0xffffffffu && leader == lane) {
            base = wbase[warp][d];
            wbase[warp][d] = base + (uint32_t)__popc(peers);
        }
        base = __shfl_sync(0xffffffffu, base, leader);
        if (d != 0xffffffffu) out[(int64_t)base + rank] = e;
        __syncwarp();
    }
}

// Sort n 64-bit elements by bits [bit_lo, bit_lo + n_bits).  Ping-pongs between a and b; returns which
// buffer holds the result through *result_in_b.  ws: sort_workspace_bytes(n).  As few passes as 9-bit digits allow
// (14 key bits: 7 + 7; 18 key bits: 9 + 9), 256 buckets whenever the digits of the pass fit 8 bits.
inline cudaError_t radix_sort_u64(uint64_t *a, uint64_t *b, int64_t n, int bit_lo, int n_bits, void *ws, cudaStream_t s,
                                  int *result_in_b, int *launches = nullptr)
{
    *result_in_b = 0;
    if (n <= 0 || n_bits <= 0) return cudaSuccess;
    int passes = (n_bits + kRadixMaxBits - 1) / kRadixMaxBits;
    int per = (n_bits + passes - 1) / passes;
    uint32_t *table = (uint32_t *)ws;
    size_t toff = ((sort_table_entries(n) + 64) * sizeof(uint32_t) + 255) / 256 * 256;
    uint32_t *scan_tmp = (uint32_t *)((char *)ws + toff);
    int64_t nb = sort_num_blocks(n);
    uint64_t *src = a, *dst = b;
    int done = 0;
    for (int p = 0; p < passes; ++p) {
        int bits = (n_bits - done) < per ? (n_bits - done) : per;
        uint32_t mask = (1u << bits) - 1u;
        int shift = bit_lo + done;
        const int radix = bits > 8 ? 512 : 256;
        if (radix == 512) sort_hist_kernel<512><<<(unsigned)nb, kSortThreads, 0, s>>>(src, n, shift, mask, table, nb);
        else sort_hist_kernel<256><<<(unsigned)nb, kSortThreads, 0, s>>>(src, n, shift, mask, table, nb);
        cudaError_t e = scan_exclusive<uint32_t>(table, nb * radix, scan_tmp, nullptr, s, launches);
        if (e != cudaSuccess) return e;
        if (radix == 512) sort_scatter_kernel<512><<<(unsigned)nb, kSortThreads, 0, s>>>(src, dst, n, shift, mask, table, nb);
        else sort_scatter_kernel<256><<<(unsigned)nb, kSortThreads, 0, s>>>(src, dst, n, shift, mask, table, nb);
        if (launches) *launches += 2;
        uint64_t *t = src; src = dst; dst = t;
        done += bits;
    }
    *result_in_b = (src == b);
    return cudaGetLastError();
}

}  // namespace ast
