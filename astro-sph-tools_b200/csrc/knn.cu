// knn.cu -- smoothing lengths by k nearest neighbours on a cell-linked list, sm_100a.
//
// Replaces the KDTree branch of SnapshotSWIFT.get_smoothing_lengths (io/SWIFT/_SnapshotSWIFT.py:62-83):
//   h_i = K-th smallest distance from particle i to all particles, i itself included,
// with scipy.spatial.cKDTree's float64 arithmetic so that distances are BIT-EQUAL to the reference's:
//   d = sqrt((dx*dx + dy*dy) + dz*dz), no FMA; periodic (boxsize): each delta is first wrapped by -+box when
//   |delta| > box/2  (SURVEY 8(a) A7).
//
//   1. cell key per particle -> (key << 32 | particle) -> stable radix sort (scan_sort.cuh)
//   2. gather positions into cell order (SoA), first/last particle of every cell
//   3. one thread per query: visit cells ring by ring (Chebyshev distance 0,1,2,...), keep the K smallest squared
//      distances in a max-heap, stop when the K-th is closer than anything that can still be outside the explored cube.
// Multi-GPU: positions are replicated (1024^3 x 24 B = 25.8 GB fits every 180 GB GPU), each rank answers the
// queries [q_begin, q_begin + q_count).
#include <string.h>

#include "ast_geom.h"
#include "common.cuh"
#include "scan_sort.cuh"

namespace ast {

struct KnnGrid {
    double lo[3], inv_cs[3], cs[3];
    double box, half_box;        // periodic if box > 0
    int G;
};

struct KnnArgs {
    KnnGrid g;
    const double *xs, *ys, *zs;  // cell order
    const uint32_t *sidx;        // cell order -> original index
    const uint32_t *cstart;      // ncell + 1: particles of cell c are [cstart[c], cstart[c+1]) in cell order
    const uint32_t *qlist;       // nullable: sorted positions of the queries
    const double *qpos;          // EXTERNAL queries: (nq,3) row-major positions, output row = query index
    int64_t nq, q_begin;
    const uint32_t *col_off;     // [G*G + 1] exclusive scan of the 32-query chunks per (cx, cy) column (warp-cooperative kernel)
    int64_t q_end;               // queries = particles with original index in [q_begin, q_end)
    int k;
    double *h_out;
    int32_t *idx_out;
    double *dist_out;
};

__host__ __device__ __forceinline__ int cell_coord(const KnnGrid &g, double x, int c)
{
    double t = floor((x - g.lo[c]) * g.inv_cs[c]);
    int i = t < 0.0 ? 0 : (t > (double)(g.G - 1) ? g.G - 1 : (int)t);
    return i;
}

__global__ void knn_key_kernel(const double *__restrict__ pos, int64_t n, KnnGrid g, uint64_t *__restrict__ elems)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = cell_coord(g, pos[3 * i], 0), cy = cell_coord(g, pos[3 * i + 1], 1), cz = cell_coord(g, pos[3 * i + 2], 2);
    const uint32_t key = ((uint32_t)cx * g.G + cy) * g.G + cz;
    elems[i] = ((uint64_t)key << 32) | (uint64_t)(uint32_t)i;
}

__global__ void knn_gather_kernel(const double *__restrict__ pos, const uint64_t *__restrict__ sorted, int64_t n,
                                  double *__restrict__ xs, double *__restrict__ ys, double *__restrict__ zs,
                                  uint32_t *__restrict__ sidx, uint32_t *__restrict__ cbeg,
                                  int64_t q_begin, int64_t q_end, uint32_t *__restrict__ qflag)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint64_t e = sorted[s];
    const uint32_t i = (uint32_t)e, key = (uint32_t)(e >> 32);
    xs[s] = pos[3 * (int64_t)i];
    ys[s] = pos[3 * (int64_t)i + 1];
    zs[s] = pos[3 * (int64_t)i + 2];
    sidx[s] = i;
    // per-cell counts from the run boundaries of the sorted keys (cbeg pre-zeroed): count = end - begin
    if (s == n - 1 || (uint32_t)(sorted[s + 1] >> 32) != key) atomicAdd(&cbeg[key], (uint32_t)(s + 1));
    if (s == 0 || (uint32_t)(sorted[s - 1] >> 32) != key) atomicSub(&cbeg[key], (uint32_t)s);
    if (qflag) qflag[s] = ((int64_t)i >= q_begin && (int64_t)i < q_end) ? 1u : 0u;
}

__global__ void knn_compact_kernel(const uint32_t *__restrict__ qexcl, const uint32_t *__restrict__ sidx, int64_t n, int64_t q_begin,
                                   int64_t q_end, uint32_t *__restrict__ qlist)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int64_t i = sidx[s];
    if (i >= q_begin && i < q_end) qlist[qexcl[s]] = (uint32_t)s;
}

// max-heap on (d2, idx): root = current K-th smallest.  The heap is a per-thread (local-memory) array on purpose: a
// shared-memory heap ([i * BLOCK + t], conflict-free) was measured 25 % SLOWER on B200 (61.2 vs 48.7 ms at 256^3, k = 48)
// because 48 KB of heaps per block leave 16 warps per SM instead of 57, and this kernel is latency-bound on divergent
// candidate gathers (ncu: long-scoreboard 18.9 stalls per issue), not on the 71 GB of local-memory DRAM traffic.
template <int KCAP, bool WANT_IDX>
struct Heap {
    double d[KCAP];
    uint32_t id[WANT_IDX ? KCAP : 1];
    int k;
    __device__ __forceinline__ bool less(double d2, uint32_t j, double e2, uint32_t l) const
    {
        return WANT_IDX ? (d2 < e2 || (d2 == e2 && j < l)) : (d2 < e2);
    }
    // 4-ary heap: children of p are 4p+1 .. 4p+4.  Three levels instead of six at k = 48, and the four child loads of a
    // level are independent (the binary version waited for two dependent local-memory loads per level: 46 % of all
    // stall samples of the query kernel, profiles/r01_v4_summary.md)
    __device__ __forceinline__ void replace_root(double d2, uint32_t j)
    {
        int p = 0;
        for (;;) {
            const int c = 4 * p + 1;
            if (c >= k) break;
            const int c1 = min(c + 1, k - 1), c2 = min(c + 2, k - 1), c3 = min(c + 3, k - 1);      // clamped: re-reads the last child
            const double e0 = d[c], e1 = d[c1], e2 = d[c2], e3 = d[c3];
            const uint32_t i0 = WANT_IDX ? id[c] : 0u, i1 = WANT_IDX ? id[c1] : 0u, i2 = WANT_IDX ? id[c2] : 0u, i3 = WANT_IDX ? id[c3] : 0u;
            int ma = c, mb = c2;
            double da = e0, db = e2;
            uint32_t ia = i0, ib = i2;
            if (less(e0, i0, e1, i1)) { ma = c1; da = e1; ia = i1; }
            if (less(e2, i2, e3, i3)) { mb = c3; db = e3; ib = i3; }
            if (less(da, ia, db, ib)) { ma = mb; da = db; ia = ib; }                                // largest child
            if (!less(d2, j, da, ia)) break;
            d[p] = da;
            if (WANT_IDX) id[p] = ia;
            p = ma;
        }
        d[p] = d2;
        if (WANT_IDX) id[p] = j;
    }
};

template <int KCAP, bool WANT_IDX, bool EXTERNAL>
__global__ void __launch_bounds__(128) knn_query_kernel(KnnArgs a)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nq) return;
    const int64_t s = EXTERNAL ? 0 : (a.qlist ? (int64_t)a.qlist[t] : t);
    const double x = EXTERNAL ? a.qpos[3 * t] : a.xs[s], y = EXTERNAL ? a.qpos[3 * t + 1] : a.ys[s],
                 z = EXTERNAL ? a.qpos[3 * t + 2] : a.zs[s];
    const KnnGrid &g = a.g;
    const int G = g.G;
    const bool per = g.box > 0.0;
    const int qc[3] = { cell_coord(g, x, 0), cell_coord(g, y, 1), cell_coord(g, z, 2) };
    const double xq[3] = { x, y, z };

    Heap<KCAP, WANT_IDX> hp;
    hp.k = a.k;
    for (int i = 0; i < a.k; ++i) {
        hp.d[i] = INFINITY;
        if (WANT_IDX) hp.id[i] = 0xffffffffu;
    }

    double kth = INFINITY;                                      // heap root (current K-th smallest d2) kept in a register
    // candidates j in [jb, je): contiguous particles of one or several consecutive cells of a z-column
    auto scan_range = [&](uint32_t jb, uint32_t je) {
        for (uint32_t j = jb; j < je; ++j) {
            double ex = a.xs[j] - x, ey = a.ys[j] - y, ez = a.zs[j] - z;
            if (per) {
                if (ex < -g.half_box) ex += g.box; else if (ex > g.half_box) ex -= g.box;
                if (ey < -g.half_box) ey += g.box; else if (ey > g.half_box) ey -= g.box;
                if (ez < -g.half_box) ez += g.box; else if (ez > g.half_box) ez -= g.box;
            }
            const double d2 = AST_DADD(AST_DADD(AST_DMUL(ex, ex), AST_DMUL(ey, ey)), AST_DMUL(ez, ez));
            if (WANT_IDX) {
                if (d2 <= kth) {                               // ties are ordered by index: look closer only then
                    const uint32_t oj = a.sidx[j];
                    if (hp.less(d2, oj, hp.d[0], hp.id[0])) { hp.replace_root(d2, oj); kth = hp.d[0]; }
                }
            } else if (d2 < kth) {
                hp.replace_root(d2, 0u);
                kth = hp.d[0];
            }
        }
    };
    // cells [z0, z1] (already inside [0, G)) of column (cx, cy): one contiguous particle range thanks to cstart
    auto scan_cells = [&](int cx, int cy, int z0, int z1) {
        const uint32_t base = ((uint32_t)cx * G + cy) * G;
        scan_range(a.cstart[base + z0], a.cstart[base + z1 + 1]);
    };
    for (int ring = 0;; ++ring) {
        for (int ox = -ring; ox <= ring; ++ox) {
            int cx = qc[0] + ox;
            if (per) { if (2 * abs(ox) > G || (2 * abs(ox) == G && ox < 0)) continue; cx = (cx % G + G) % G; }
            else if (cx < 0 || cx >= G) continue;
            for (int oy = -ring; oy <= ring; ++oy) {
                int cy = qc[1] + oy;
                if (per) { if (2 * abs(oy) > G || (2 * abs(oy) == G && oy < 0)) continue; cy = (cy % G + G) % G; }
                else if (cy < 0 || cy >= G) continue;
                if (abs(ox) == ring || abs(oy) == ring) {
                    // outer column of the shell: the whole z-run [qz - ring, qz + ring], visited as contiguous ranges
                    if (!per) {
                        scan_cells(cx, cy, max(qc[2] - ring, 0), min(qc[2] + ring, G - 1));
                    } else if (2 * ring + 1 >= G) {
                        scan_cells(cx, cy, 0, G - 1);                        // the run wraps onto itself: every cell once
                    } else {
                        const int z0 = qc[2] - ring, z1 = qc[2] + ring;
                        if (z0 < 0) { scan_cells(cx, cy, z0 + G, G - 1); scan_cells(cx, cy, 0, z1); }
                        else if (z1 >= G) { scan_cells(cx, cy, z0, G - 1); scan_cells(cx, cy, 0, z1 - G); }
                        else scan_cells(cx, cy, z0, z1);
                    }
                } else {
                    // inner column: only the two caps oz = -ring, +ring
                    for (int oz = -ring; oz <= ring; oz += 2 * ring) {
                        int cz = qc[2] + oz;
                        if (per) { if (2 * abs(oz) > G || (2 * abs(oz) == G && oz < 0)) continue; cz = (cz % G + G) % G; }
                        else if (cz < 0 || cz >= G) continue;
                        scan_cells(cx, cy, cz, cz);
                    }
                }
            }
        }
        // smallest possible distance to anything outside the explored cube of cells
        double dmin = INFINITY;
        bool all = true;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double frac = xq[c] - (g.lo[c] + (double)qc[c] * g.cs[c]);
            bool lo_open, hi_open;
            if (per) lo_open = hi_open = (2 * ring + 1 < G);
            else { lo_open = qc[c] - ring > 0; hi_open = qc[c] + ring < G - 1; }
            if (lo_open) dmin = fmin(dmin, frac + (double)ring * g.cs[c]);
            if (hi_open) dmin = fmin(dmin, (g.cs[c] - frac) + (double)ring * g.cs[c]);
            all = all && !lo_open && !hi_open;
        }
        if (all) break;
        const double safe = dmin - 1e-9 * g.cs[0];
        if (safe > 0.0 && kth < safe * safe) break;
    }

    const int64_t row = EXTERNAL ? t : (int64_t)a.sidx[s] - a.q_begin;
    if (a.h_out) a.h_out[row] = sqrt(hp.d[0]);
    if (WANT_IDX) {
        // heap-sort in place: ascending (d2, idx)
        const int k = a.k;
        for (int end = k - 1; end > 0; --end) {
            const double dd = hp.d[end];
            const uint32_t ii = hp.id[end];
            hp.d[end] = hp.d[0];
            hp.id[end] = hp.id[0];
            hp.k = end;
            hp.replace_root(dd, ii);
        }
        for (int i = 0; i < k; ++i) {
            if (a.idx_out) a.idx_out[row * k + i] = hp.d[i] < INFINITY ? (int32_t)hp.id[i] : -1;
            if (a.dist_out) a.dist_out[row * k + i] = sqrt(hp.d[i]);
        }
    }
}


// ---- warp-cooperative queries ---------------------------------------------------------------------------------------
// One warp answers 32 consecutive particles of one (cx, cy) column of cells (a run of cells along z, contiguous in the
// cell-ordered arrays).  All 32 queries walk the SAME candidates: for every ring the warp visits the new columns
// (cx + ox, cy + oy) over the z-range [zlo - ring, zhi + ring] of the chunk and the two new z-caps of the columns it already
// knows; each piece is one contiguous particle range (prefix cstart), loaded 32 candidates at a time into shared memory by
// the whole warp (coalesced) and then read by every lane as a broadcast.  No divergence in the distance loop, candidate
// loads shared by 32 queries; only the heap insert is per lane.  The thread-per-query kernel above evaluates ~250 candidates
// per query on 16 of 32 lanes and stalls on scattered gathers; this one evaluates ~860 per query in lock step (23 SASS
// instructions each) and, as measured, spends as many instructions again in the lock-step heap merges (ncu: 3.5e10 warp
// instructions, 80 ms at 256^3 against 2.8e10 and 46 ms), so it is NOT the default; it is kept, tested bit-equal to scipy,
// as the starting point for a selection scheme that does not pay a sift-down per accepted candidate.
// Arithmetic is scipy's, pair by pair: d2 = (ex*ex + ey*ey) + ez*ez with each delta wrapped by -+box when |delta| > box/2.
// When every pair of (chunk, candidate piece) provably takes the same wrap branch on an axis (cell offsets at least one cell
// away from box/2: 2 (|offset| + 2) <= G) the wrap is a per-piece constant shift added to the delta -- the same operation
// the branch would have performed -- otherwise the per-pair comparison is kept.
__global__ void knn_col_chunks_kernel(const uint32_t *__restrict__ cstart, int G, uint32_t *__restrict__ col_off)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col > G * G) return;
    col_off[col] = col < G * G ? (cstart[(int64_t)col * G + G] - cstart[(int64_t)col * G] + 31u) / 32u : 0u;
}

// number of chunks with at least one query of the subset [q_begin, q_end) (decides between the two query kernels)
__global__ void knn_active_chunks_kernel(const uint32_t *__restrict__ cstart, const uint32_t *__restrict__ col_off,
                                         const uint32_t *__restrict__ sidx, int G, int64_t q_begin, int64_t q_end,
                                         unsigned long long *__restrict__ count)
{
    const int col = blockIdx.x;
    const uint32_t b = cstart[(int64_t)col * G], e = cstart[(int64_t)col * G + G];
    unsigned long long mine = 0;
    for (uint32_t c = b + 32u * threadIdx.x; c < e; c += 32u * blockDim.x) {
        bool any = false;
        for (uint32_t j = c; j < e && j < c + 32u; ++j) any = any || ((int64_t)sidx[j] >= q_begin && (int64_t)sidx[j] < q_end);
        mine += any ? 1ull : 0ull;
    }
    if (mine) atomicAdd(count, mine);
}

// max-heap in SHARED memory, entry i of lane l at [i * 32 + l] (conflict-free); see knn_block_kernel
template <bool WANT_IDX>
struct SHeap {
    double *d;          // already offset by the lane
    uint32_t *id;
    int k;
    __device__ __forceinline__ double &D(int i) const { return d[i * 32]; }
    __device__ __forceinline__ uint32_t &I(int i) const { return id[i * 32]; }
    __device__ __forceinline__ bool less(double d2, uint32_t j, double e2, uint32_t l) const
    {
        return WANT_IDX ? (d2 < e2 || (d2 == e2 && j < l)) : (d2 < e2);
    }
    __device__ __forceinline__ void replace_root(double d2, uint32_t j)      // 4-ary, as Heap::replace_root
    {
        int p = 0;
        for (;;) {
            const int c = 4 * p + 1;
            if (c >= k) break;
            const int c1 = min(c + 1, k - 1), c2 = min(c + 2, k - 1), c3 = min(c + 3, k - 1);
            const double e0 = D(c), e1 = D(c1), e2 = D(c2), e3 = D(c3);
            const uint32_t i0 = WANT_IDX ? I(c) : 0u, i1 = WANT_IDX ? I(c1) : 0u, i2 = WANT_IDX ? I(c2) : 0u, i3 = WANT_IDX ? I(c3) : 0u;
            int ma = c, mb = c2;
            double da = e0, db = e2;
            uint32_t ia = i0, ib = i2;
            if (less(e0, i0, e1, i1)) { ma = c1; da = e1; ia = i1; }
            if (less(e2, i2, e3, i3)) { mb = c3; db = e3; ib = i3; }
            if (less(da, ia, db, ib)) { ma = mb; da = db; ia = ib; }
            if (!less(d2, j, da, ia)) break;
            D(p) = da;
            if (WANT_IDX) I(p) = ia;
            p = ma;
        }
        D(p) = d2;
        if (WANT_IDX) I(p) = j;
    }
};

constexpr int kPend = 16;          // pending-list slots per lane (flush when a lane holds more than kPend - 8)
__host__ __device__ inline size_t knn_block_warp_bytes(int k, bool want_idx)
{
    return (size_t)(k + kPend + 3) * 32 * sizeof(double) + (want_idx ? (size_t)(k + kPend + 1) * 32 * sizeof(uint32_t) : 0);
}

template <int WARPS, bool WANT_IDX, bool PER>
__global__ void __launch_bounds__(WARPS * 32) knn_block_kernel(KnnArgs a)
{
    // per warp: heap [k][32] doubles, pending [kPend][32], candidates x, y, z [32]; then (lists) the matching uint32 indices
    extern __shared__ __align__(16) unsigned char knn_smem[];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double *const wd = reinterpret_cast<double *>(knn_smem + (size_t)wib * knn_block_warp_bytes(a.k, WANT_IDX));
    double *const pend = wd + (size_t)a.k * 32 + lane;                    // pend[t * 32]
    double *const sx = wd + (size_t)(a.k + kPend) * 32, *const sy = sx + 32, *const sz = sy + 32;
    uint32_t *const wi = reinterpret_cast<uint32_t *>(sz + 32);
    uint32_t *const pend_id = wi + (size_t)a.k * 32 + lane;
    uint32_t *const si = wi + (size_t)(a.k + kPend) * 32;
    const uint32_t w = blockIdx.x * (uint32_t)WARPS + (uint32_t)wib;      // chunk index
    const KnnGrid &g = a.g;
    const int G = g.G, ncol = G * G;
    if (w >= a.col_off[ncol]) return;
    int lo = 0, hi = ncol - 1;                                            // first column with col_off[col + 1] > w
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a.col_off[mid + 1] > w) hi = mid; else lo = mid + 1;
    }
    const int col = lo, cx = col / G, cy = col - cx * G;
    const uint32_t cend = a.cstart[(int64_t)col * G + G];
    const uint32_t s0 = a.cstart[(int64_t)col * G] + 32u * (w - a.col_off[col]);
    const uint32_t s = s0 + (uint32_t)lane < cend ? s0 + (uint32_t)lane : s0;      // lanes past the end shadow the first query
    const uint32_t oi = a.sidx[s];
    const bool active = s0 + (uint32_t)lane < cend && (int64_t)oi >= a.q_begin && (int64_t)oi < a.q_end;
    if (!__any_sync(FULL, active)) return;
    const double x = a.xs[s], y = a.ys[s], z = a.zs[s];
    const int qc[3] = { cx, cy, cell_coord(g, z, 2) };
    const double xq[3] = { x, y, z };
    const int zlo = __reduce_min_sync(FULL, active ? qc[2] : G), zhi = __reduce_max_sync(FULL, active ? qc[2] : -1);
    const int zspan = zhi - zlo;

    SHeap<WANT_IDX> hp;
    hp.k = a.k;
    hp.d = wd + lane;
    hp.id = wi + lane;
    for (int i = 0; i < a.k; ++i) {
        hp.D(i) = INFINITY;
        if (WANT_IDX) hp.I(i) = 0xffffffffu;
    }
    double kth = active ? INFINITY : -INFINITY;                           // inactive lanes never accept a candidate

    // Selection in lock step: a candidate that beats the lane's current K-th distance is only APPENDED to a small per-lane
    // pending list (predicated store, no branch); the lists are merged into the heaps by all lanes together when one of them
    // could overflow with the next 32 candidates, and before every termination test.  Inserting straight into the heap would
    // run the divergent sift-down for almost every candidate (some lane of the 32 nearly always accepts).
    int cnt = 0;
    auto offer = [&](double d2, uint32_t oj) {
        if (WANT_IDX ? d2 <= kth : d2 < kth) {
            pend[cnt * 32] = d2;
            if (WANT_IDX) pend_id[cnt * 32] = oj;
            ++cnt;
        }
    };
    auto flush = [&]() {
        const int mx = __reduce_max_sync(FULL, cnt);
        for (int t = 0; t < mx; ++t) {
            if (t < cnt) {
                const double d2 = pend[t * 32];
                const uint32_t oj = WANT_IDX ? pend_id[t * 32] : 0u;
                if (hp.less(d2, oj, hp.D(0), WANT_IDX ? hp.I(0) : 0u)) hp.replace_root(d2, oj);
            }
        }
        cnt = 0;
        kth = active ? hp.D(0) : -INFINITY;
    };
    // candidates [jb, je): simple = the wrap of every pair is the constant shift (shx, shy, shz)
    auto scan_range = [&](uint32_t jb, uint32_t je, bool simple, double shx, double shy, double shz) {
        for (uint32_t j0 = jb; j0 < je; j0 += 32u) {
            const uint32_t j = j0 + (uint32_t)lane;
            __syncwarp();
            if (j < je) {
                sx[lane] = a.xs[j]; sy[lane] = a.ys[j]; sz[lane] = a.zs[j];
                if (WANT_IDX) si[lane] = a.sidx[j];
            }
            __syncwarp();
            const int m = (int)(je - j0 < 32u ? je - j0 : 32u);
            for (int t0 = 0; t0 < m; t0 += 8) {                            // at most 8 appends per lane between two flush tests
                const int t1 = min(t0 + 8, m);
                if (!PER || simple) {
                    for (int t = t0; t < t1; ++t) {
                        double ex = sx[t] - x, ey = sy[t] - y, ez = sz[t] - z;
                        if (PER) { ex = AST_DADD(ex, shx); ey = AST_DADD(ey, shy); ez = AST_DADD(ez, shz); }
                        offer(AST_DADD(AST_DADD(AST_DMUL(ex, ex), AST_DMUL(ey, ey)), AST_DMUL(ez, ez)), WANT_IDX ? si[t] : 0u);
                    }
                } else {
                    for (int t = t0; t < t1; ++t) {
                        double ex = sx[t] - x, ey = sy[t] - y, ez = sz[t] - z;
                        if (ex < -g.half_box) ex += g.box; else if (ex > g.half_box) ex -= g.box;
                        if (ey < -g.half_box) ey += g.box; else if (ey > g.half_box) ey -= g.box;
                        if (ez < -g.half_box) ez += g.box; else if (ez > g.half_box) ez -= g.box;
                        offer(AST_DADD(AST_DADD(AST_DMUL(ex, ex), AST_DMUL(ey, ey)), AST_DMUL(ez, ez)), WANT_IDX ? si[t] : 0u);
                    }
                }
                if (__any_sync(FULL, cnt > kPend - 8)) flush();
            }
        }
    };
    // raw z-cells [z0, z1] (may stick out of [0, G)) of column (ccx, ccy); at most G cells
    auto scan_z = [&](int ccx, int ccy, int z0, int z1, bool simple_xy, double shx, double shy, int ring) {
        const uint32_t base = ((uint32_t)ccx * G + ccy) * G;
        if (!PER) {
            z0 = max(z0, 0); z1 = min(z1, G - 1);
            if (z0 <= z1) scan_range(a.cstart[base + z0], a.cstart[base + z1 + 1], true, 0.0, 0.0, 0.0);
            return;
        }
        const bool simple = simple_xy && 2 * (zspan + ring + 2) <= G;
        if (z1 - z0 + 1 >= G) { scan_range(a.cstart[base], a.cstart[base + G], false, 0.0, 0.0, 0.0); return; }
        if (z0 < 0) {
            // cells z0+G .. min(z1,-1)+G lie above the chunk after wrapping: raw delta ~ +box -> the pair subtracts box
            const int zt = min(z1, -1);
            scan_range(a.cstart[base + z0 + G], a.cstart[base + zt + G + 1], simple, shx, shy, -g.box);
            if (z1 >= 0) scan_range(a.cstart[base], a.cstart[base + z1 + 1], simple, shx, shy, 0.0);
        } else if (z1 >= G) {
            const int zb = max(z0, G);
            if (z0 < G) scan_range(a.cstart[base + z0], a.cstart[base + G], simple, shx, shy, 0.0);
            scan_range(a.cstart[base + zb - G], a.cstart[base + z1 - G + 1], simple, shx, shy, g.box);
        } else {
            scan_range(a.cstart[base + z0], a.cstart[base + z1 + 1], simple, shx, shy, 0.0);
        }
    };

    for (int ring = 0;; ++ring) {
        const int len_prev = zspan + 1 + 2 * (ring - 1);                  // z-cells of a known column before this ring
        for (int ox = -ring; ox <= ring; ++ox) {
            int ccx = cx + ox;
            double shx = 0.0;
            if (PER) {
                if (2 * abs(ox) > G || (2 * abs(ox) == G && ox < 0)) continue;
                if (ccx < 0) { ccx += G; shx = -g.box; } else if (ccx >= G) { ccx -= G; shx = g.box; }
            } else if (ccx < 0 || ccx >= G) continue;
            for (int oy = -ring; oy <= ring; ++oy) {
                int ccy = cy + oy;
                double shy = 0.0;
                if (PER) {
                    if (2 * abs(oy) > G || (2 * abs(oy) == G && oy < 0)) continue;
                    if (ccy < 0) { ccy += G; shy = -g.box; } else if (ccy >= G) { ccy -= G; shy = g.box; }
                } else if (ccy < 0 || ccy >= G) continue;
                const bool simple_xy = 2 * (abs(ox) + 2) <= G && 2 * (abs(oy) + 2) <= G;
                if (abs(ox) == ring || abs(oy) == ring) {
                    scan_z(ccx, ccy, zlo - ring, zhi + ring, simple_xy, shx, shy, ring);          // new column: whole z-range
                } else if (!PER) {
                    scan_z(ccx, ccy, zlo - ring, zlo - ring, simple_xy, shx, shy, ring);          // known column: the two new caps
                    scan_z(ccx, ccy, zhi + ring, zhi + ring, simple_xy, shx, shy, ring);
                } else {
                    // periodic: a cap is new only while the known z-range has not closed on itself
                    if (len_prev + 1 <= G) scan_z(ccx, ccy, zlo - ring, zlo - ring, simple_xy, shx, shy, ring);
                    if (len_prev + 2 <= G) scan_z(ccx, ccy, zhi + ring, zhi + ring, simple_xy, shx, shy, ring);
                }
            }
        }
        flush();
        // per lane: smallest possible distance to anything outside the cube of cells explored around ITS OWN cell (the
        // warp has explored a superset of it)
        double dmin = INFINITY;
        bool all = true;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double frac = xq[c] - (g.lo[c] + (double)qc[c] * g.cs[c]);
            bool lo_open, hi_open;
            if (PER) lo_open = hi_open = (2 * ring + 1 < G);
            else { lo_open = qc[c] - ring > 0; hi_open = qc[c] + ring < G - 1; }
            if (lo_open) dmin = fmin(dmin, frac + (double)ring * g.cs[c]);
            if (hi_open) dmin = fmin(dmin, (g.cs[c] - frac) + (double)ring * g.cs[c]);
            all = all && !lo_open && !hi_open;
        }
        const double safe = dmin - 1e-9 * g.cs[0];
        const bool done = !active || all || (safe > 0.0 && kth < safe * safe);
        if (__all_sync(FULL, done)) break;
    }
    if (!active) return;
    const int64_t row = (int64_t)oi - a.q_begin;
    if (a.h_out) a.h_out[row] = sqrt(hp.D(0));
    if (WANT_IDX) {
        const int k = a.k;
        for (int end = k - 1; end > 0; --end) {                 // heap-sort in place: ascending (d2, idx)
            const double dd = hp.D(end);
            const uint32_t ii = hp.I(end);
            hp.D(end) = hp.D(0);
            hp.I(end) = hp.I(0);
            hp.k = end;
            hp.replace_root(dd, ii);
        }
        for (int i = 0; i < k; ++i) {
            if (a.idx_out) a.idx_out[row * k + i] = hp.D(i) < INFINITY ? (int32_t)hp.I(i) : -1;
            if (a.dist_out) a.dist_out[row * k + i] = sqrt(hp.D(i));
        }
    }
}

struct KnnLayout {
    int G;
    int64_t ncell, nq;
    uint64_t *ea, *eb;
    double *xs, *ys, *zs;
    uint32_t *sidx, *cbeg, *qflag, *qlist, *scan_tmp, *cell_tmp;   // cbeg: ncell + 1 counts -> exclusive scan = cstart
    uint32_t *col_off;                                             // G*G + 1 chunk offsets (warp-cooperative kernel)
    unsigned long long *active_chunks;
    void *sort_ws;
    size_t bytes;
};

static int knn_validate(const ast_knn_params *p)
{
    AST_REQUIRE(p != nullptr, "params is null");
    AST_REQUIRE(p->n >= 0 && p->n < (1ll << 31), "n out of range");
    AST_REQUIRE(p->k >= 1 && p->k <= 128, "k = %d not in [1, 128]", p->k);
    AST_REQUIRE(p->q_begin >= 0 && p->q_begin + (p->q_count > 0 ? p->q_count : 0) <= p->n, "query range outside [0, n)");
    if (!(p->box > 0.0))
        for (int c = 0; c < 3; ++c) AST_REQUIRE(p->hi[c] >= p->lo[c], "open box needs lo <= hi (extent of the positions)");
    return AST_OK;
}

static KnnLayout knn_layout(const ast_knn_params *p, void *ws)
{
    KnnLayout L;
    // mean particles per cell; measured on B200 (benchmarks/knn_probe.py, k = 48, 256^3): 2 per cell is twice as fast as
    // k/3 per cell (the explored cube of cells hugs the k-neighbour sphere more tightly; empty cells are cheap)
    const double m = p->cell_target > 0 ? p->cell_target : 2.0;
    double g = floor(cbrt((double)(p->n > 0 ? p->n : 1) / m));
    L.G = g < 1 ? 1 : (g > 1000 ? 1000 : (int)g);
    L.ncell = (int64_t)L.G * L.G * L.G;
    const int64_t n = p->n > 0 ? p->n : 1;
    L.nq = p->q_count > 0 ? p->q_count : p->n;
    Carver c(ws);
    L.ea = c.take<uint64_t>(n);
    L.eb = c.take<uint64_t>(n);
    L.xs = c.take<double>(n);
    L.ys = c.take<double>(n);
    L.zs = c.take<double>(n);
    L.sidx = c.take<uint32_t>(n);
    L.cbeg = c.take<uint32_t>(L.ncell + 1);
    L.cell_tmp = (uint32_t *)c.take<char>(scan_workspace_bytes<uint32_t>(L.ncell + 1));
    L.col_off = c.take<uint32_t>((int64_t)L.G * L.G + 1);
    L.active_chunks = c.take<unsigned long long>(1);
    L.qflag = c.take<uint32_t>(n);
    L.qlist = c.take<uint32_t>(n);
    L.scan_tmp = (uint32_t *)c.take<char>(scan_workspace_bytes<uint32_t>(n));
    L.sort_ws = c.take<char>(sort_workspace_bytes(n));
    L.bytes = c.bytes();
    return L;
}

template <int KCAP>
static void launch_query(const KnnArgs &a, bool want_idx, cudaStream_t s)
{
    const unsigned nb = (unsigned)((a.nq + 127) / 128);
    if (a.qpos) knn_query_kernel<KCAP, true, true><<<nb, 128, 0, s>>>(a);
    else if (want_idx) knn_query_kernel<KCAP, true, false><<<nb, 128, 0, s>>>(a);
    else knn_query_kernel<KCAP, false, false><<<nb, 128, 0, s>>>(a);
}

template <int WARPS, bool WANT_IDX, bool PER>
static cudaError_t launch_block_t(const KnnArgs &a, int64_t n, cudaStream_t s)
{
    const size_t smem = WARPS * knn_block_warp_bytes(a.k, WANT_IDX);
    cudaError_t e = cudaFuncSetAttribute(knn_block_kernel<WARPS, WANT_IDX, PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t max_chunks = n / 32 + (int64_t)a.g.G * a.g.G;          // sum over columns of ceil(count / 32) <= this
    knn_block_kernel<WARPS, WANT_IDX, PER><<<(unsigned)((max_chunks + WARPS - 1) / WARPS), WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}
template <int WARPS>
static cudaError_t launch_block_w(const KnnArgs &a, bool want_idx, int64_t n, cudaStream_t s)
{
    const bool per = a.g.box > 0.0;
    if (want_idx) return per ? launch_block_t<WARPS, true, true>(a, n, s) : launch_block_t<WARPS, true, false>(a, n, s);
    return per ? launch_block_t<WARPS, false, true>(a, n, s) : launch_block_t<WARPS, false, false>(a, n, s);
}
// warps per block by k: the shared-memory heaps ((k + 19) * 256 bytes per warp, 1.5x with lists) should leave >= 3 blocks per SM
static cudaError_t launch_block(const KnnArgs &a, bool want_idx, int64_t n, cudaStream_t s)
{
    return a.k <= 64 ? launch_block_w<4>(a, want_idx, n, s) : launch_block_w<2>(a, want_idx, n, s);
}

// builds the cell list of `pos` in the workspace (steps 1 and 2) and fills the grid / array part of KnnArgs
static int knn_build(const ast_knn_params *p, const double *pos, const KnnLayout &L, bool subset, int64_t q_begin, int64_t q_end,
                     cudaStream_t s, KnnArgs &a)
{
    KnnGrid g;
    g.G = L.G;
    g.box = p->box > 0.0 ? p->box : 0.0;
    g.half_box = 0.5 * g.box;
    for (int c = 0; c < 3; ++c) {
        const double lo = p->box > 0.0 ? 0.0 : p->lo[c];
        double ext = p->box > 0.0 ? p->box : (p->hi[c] - p->lo[c]);
        if (!(ext > 0.0)) ext = 1.0;                       // degenerate axis: every particle lands in cell 0
        g.lo[c] = lo;
        g.cs[c] = ext / (double)L.G;
        g.inv_cs[c] = (double)L.G / ext;
    }
    const int64_t n = p->n;
    const unsigned nb = (unsigned)((n + 255) / 256);
    knn_key_kernel<<<nb, 256, 0, s>>>(pos, n, g, L.ea);
    int in_b = 0;
    AST_CUDA_TRY(radix_sort_u64(L.ea, L.eb, n, 32, ceil_log2_u64((uint64_t)L.ncell), L.sort_ws, s, &in_b));
    const uint64_t *sorted = in_b ? L.eb : L.ea;
    AST_CUDA_TRY(cudaMemsetAsync(L.cbeg, 0, sizeof(uint32_t) * (L.ncell + 1), s));
    knn_gather_kernel<<<nb, 256, 0, s>>>(pos, sorted, n, L.xs, L.ys, L.zs, L.sidx, L.cbeg, q_begin, q_end,
                                         subset ? L.qflag : nullptr);
    AST_CUDA_TRY(scan_exclusive<uint32_t>(L.cbeg, L.ncell + 1, L.cell_tmp, nullptr, s));      // counts -> cstart
    if (subset) {
        AST_CUDA_TRY(scan_exclusive<uint32_t>(L.qflag, n, L.scan_tmp, nullptr, s));
        knn_compact_kernel<<<nb, 256, 0, s>>>(L.qflag, L.sidx, n, q_begin, q_end, L.qlist);
    }
    a.g = g;
    a.xs = L.xs; a.ys = L.ys; a.zs = L.zs; a.sidx = L.sidx; a.cstart = L.cbeg;
    a.qlist = subset ? L.qlist : nullptr;
    a.qpos = nullptr;
    a.k = p->k;
    AST_CUDA_TRY(cudaGetLastError());
    return AST_OK;
}

}  // namespace ast

using namespace ast;

extern "C" int ast_knn_workspace_bytes(const ast_knn_params *p, size_t *bytes)
{
    int rc = knn_validate(p);
    if (rc) return rc;
    AST_REQUIRE(bytes != nullptr, "bytes is null");
    *bytes = knn_layout(p, nullptr).bytes;
    return AST_OK;
}

extern "C" int ast_knn_h(const ast_knn_params *p, const double *pos, double *h_out, int32_t *idx_out, double *dist_out,
                         void *workspace, size_t workspace_bytes, void *stream)
{
    int rc = knn_validate(p);
    if (rc) return rc;
    if (p->n == 0) return AST_OK;
    AST_REQUIRE(pos && h_out, "null pointer");
    KnnLayout L = knn_layout(p, workspace);
    if (!workspace || workspace_bytes < L.bytes) {
        set_error("workspace too small: need %zu bytes, have %zu", L.bytes, workspace_bytes);
        return AST_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t n = p->n;
    const bool subset = p->q_count > 0 && p->q_count < n;
    const int64_t q_begin = subset ? p->q_begin : 0, q_end = subset ? p->q_begin + p->q_count : n;
    KnnArgs a;
    rc = knn_build(p, pos, L, subset, q_begin, q_end, s, a);
    if (rc) return rc;
    a.nq = q_end - q_begin;
    a.q_begin = q_begin;
    a.q_end = q_end;
    a.h_out = h_out; a.idx_out = idx_out; a.dist_out = dist_out;
    const bool want = idx_out != nullptr || dist_out != nullptr;
    // warp-cooperative kernel (opt-in, AST_KNN_WARP_COOPERATIVE): 32 consecutive particles of a column of cells per warp.
    // Measured on B200 (256^3, k = 48): 80 ms against 46 ms for one thread per query -- see the note in DESIGN.md section 7.
    // With a query subset it only makes sense if the subset is spatially coherent (index ranges of snapshot files are):
    // count the chunks that hold at least one query and fall back when more than 3x the ideal number would have to run.
    bool cooperative = (p->flags & AST_KNN_WARP_COOPERATIVE) != 0;
    if (cooperative) {
        const int ncol = L.G * L.G;
        knn_col_chunks_kernel<<<(unsigned)((ncol + 1 + 255) / 256), 256, 0, s>>>(L.cbeg, L.G, L.col_off);
        AST_CUDA_TRY(scan_exclusive<uint32_t>(L.col_off, (int64_t)ncol + 1, L.cell_tmp, nullptr, s));
        a.col_off = L.col_off;
        if (subset) {
            unsigned long long active = 0;
            AST_CUDA_TRY(cudaMemsetAsync(L.active_chunks, 0, sizeof(unsigned long long), s));
            knn_active_chunks_kernel<<<(unsigned)ncol, 32, 0, s>>>(L.cbeg, L.col_off, L.sidx, L.G, q_begin, q_end, L.active_chunks);
            AST_CUDA_TRY(cudaMemcpyAsync(&active, L.active_chunks, sizeof active, cudaMemcpyDeviceToHost, s));
            AST_CUDA_TRY(cudaStreamSynchronize(s));
            cooperative = (int64_t)active * 32 <= 3 * a.nq + 3 * 32;
        }
    }
    if (cooperative) {
        AST_CUDA_TRY(launch_block(a, want, n, s));
    } else {
        if (p->k <= 32) launch_query<32>(a, want, s);
        else if (p->k <= 48) launch_query<48>(a, want, s);
        else if (p->k <= 64) launch_query<64>(a, want, s);
        else launch_query<128>(a, want, s);
    }
    AST_CUDA_TRY(cudaGetLastError());
    return AST_OK;
}

// k nearest data points of every query point (separate query set): the reference's nearest-halo lookup
// KDTree(centres, boxsize).query(particles) (_scripts/find_nearest_haloes.py:207-215) generalised to k neighbours.
extern "C" int ast_knn_query(const ast_knn_params *p, const double *data_pos, const double *query_pos, int64_t n_query,
                             double *dist_out, int32_t *idx_out, void *workspace, size_t workspace_bytes, void *stream)
{
    int rc = knn_validate(p);
    if (rc) return rc;
    AST_REQUIRE(n_query >= 0, "n_query < 0");
    if (n_query == 0) return AST_OK;
    AST_REQUIRE(p->n > 0 && data_pos && query_pos && (dist_out || idx_out), "null pointer or empty data set");
    KnnLayout L = knn_layout(p, workspace);
    if (!workspace || workspace_bytes < L.bytes) {
        set_error("workspace too small: need %zu bytes, have %zu", L.bytes, workspace_bytes);
        return AST_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    KnnArgs a;
    rc = knn_build(p, data_pos, L, false, 0, p->n, s, a);
    if (rc) return rc;
    a.qpos = query_pos;
    a.nq = n_query;
    a.q_begin = 0;
    a.q_end = p->n;
    a.col_off = nullptr;
    a.h_out = nullptr; a.idx_out = idx_out; a.dist_out = dist_out;
    if (p->k <= 32) launch_query<32>(a, true, s);
    else if (p->k <= 48) launch_query<48>(a, true, s);
    else if (p->k <= 64) launch_query<64>(a, true, s);
    else launch_query<128>(a, true, s);
    AST_CUDA_TRY(cudaGetLastError());
    return AST_OK;
}
