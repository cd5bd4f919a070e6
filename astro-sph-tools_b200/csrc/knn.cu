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

__global__ void knn_key_kernel(const double *__restrict__ pos, int64_t n, KnnGrid g, const uint32_t *__restrict__ kept,
                               uint64_t *__restrict__ elems)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const int64_t i = kept ? (int64_t)kept[t] : t;               // n = number of kept particles when `kept` is given
    const int cx = cell_coord(g, pos[3 * i], 0), cy = cell_coord(g, pos[3 * i + 1], 1), cz = cell_coord(g, pos[3 * i + 2], 2);
    const uint32_t key = ((uint32_t)cx * g.G + cy) * g.G + cz;
    elems[t] = ((uint64_t)key << 32) | (uint64_t)(uint32_t)i;
}

// ---- reach-limited build for a query subset (multi-GPU: every rank answers an index range) ---------------------------
// Only particles within `margin` of the bounding box of the queries can be among their K nearest neighbours as long as the
// K-th distance stays below `margin`; the cell list is then built from those particles alone (the sort dominates the build,
// and it is replicated on every rank otherwise).  The caller verifies h <= margin for every query afterwards and widens.
struct KnnReach {
    double lo[3], len[3];        // per axis the queries lie in [lo, lo + len] (periodic: modulo box, the interval may wrap)
    double margin, box;          // box > 0: periodic
};
__device__ __forceinline__ bool knn_in_reach(const KnnReach &r, double x, double y, double z)
{
    const double q[3] = { x, y, z };
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double t = q[c] - r.lo[c], d;
        if (r.box > 0.0) {
            t -= floor(t / r.box) * r.box;                                   // position along the circle, from lo
            d = t <= r.len[c] ? 0.0 : fmin(t - r.len[c], r.box - t);
        } else {
            d = t < 0.0 ? -t : (t > r.len[c] ? t - r.len[c] : 0.0);
        }
        ok = ok && d <= r.margin;
    }
    return ok;
}
// periodic boxes: which of 64 equal bins per axis hold a query (the host takes the complement of the largest empty arc)
__global__ void knn_binmask_kernel(const double *__restrict__ pos, int64_t q_begin, int64_t q_end, double box, unsigned long long *__restrict__ masks)
{
    unsigned long long m[3] = { 0ull, 0ull, 0ull };
    for (int64_t i = q_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q_end; i += (int64_t)gridDim.x * blockDim.x)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int b = (int)floor(pos[3 * i + c] / box * 64.0);
            b = b < 0 ? 0 : (b > 63 ? 63 : b);
            m[c] |= 1ull << b;
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m[c] |= __shfl_xor_sync(0xffffffffu, m[c], o);
        if ((threadIdx.x & 31) == 0 && m[c]) atomicOr(&masks[c], m[c]);
    }
}
// per-block min / max of the query positions: out[block][6]
__global__ void __launch_bounds__(256) knn_bbox_kernel(const double *__restrict__ pos, int64_t q_begin, int64_t q_end, double *__restrict__ out)
{
    __shared__ double sm[6][8];
    double lo[3] = { INFINITY, INFINITY, INFINITY }, hi[3] = { -INFINITY, -INFINITY, -INFINITY };
    for (int64_t i = q_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < q_end; i += (int64_t)gridDim.x * blockDim.x)
#pragma unroll
        for (int c = 0; c < 3; ++c) { const double v = pos[3 * i + c]; lo[c] = fmin(lo[c], v); hi[c] = fmax(hi[c], v); }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fmin(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmax(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
#pragma unroll
        for (int c = 0; c < 3; ++c) { sm[c][warp] = lo[c]; sm[3 + c][warp] = hi[c]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sm[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) v = threadIdx.x < 3 ? fmin(v, sm[threadIdx.x][w]) : fmax(v, sm[threadIdx.x][w]);
        out[blockIdx.x * 6 + threadIdx.x] = v;
    }
}
__global__ void knn_reach_flag_kernel(const double *__restrict__ pos, int64_t n, KnnReach r, uint32_t *__restrict__ flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = knn_in_reach(r, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]) ? 1u : 0u;
}
__global__ void knn_reach_compact_kernel(const double *__restrict__ pos, int64_t n, KnnReach r, const uint32_t *__restrict__ excl,
                                         uint32_t *__restrict__ kept)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && knn_in_reach(r, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2])) kept[excl[i]] = (uint32_t)i;
}
// number of queries whose K-th distance exceeds the margin (their neighbourhood may be incomplete)
__global__ void knn_unsafe_kernel(const double *__restrict__ h, int64_t nq, double margin, unsigned long long *__restrict__ count)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool bad = i < nq && !(h[i] <= margin);
    const unsigned b = __ballot_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, (unsigned long long)__popc(b));
}

__global__ void knn_gather_kernel(const double *__restrict__ pos, const uint64_t *__restrict__ sorted, int64_t n,
                                  double *__restrict__ xs, double *__restrict__ ys, double *__restrict__ zs,
                                  uint32_t *__restrict__ sidx, uint32_t *__restrict__ cbeg,
                                  int64_t q_begin, int64_t q_end, uint32_t *__restrict__ qflag, const uint32_t *__restrict__ orig_flag,
                                  uint32_t flag_value)
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
    if (qflag) qflag[s] = orig_flag ? (orig_flag[i] == flag_value ? 1u : 0u) : (((int64_t)i >= q_begin && (int64_t)i < q_end) ? 1u : 0u);
}

__global__ void knn_compact_kernel(const uint32_t *__restrict__ qexcl, const uint32_t *__restrict__ sidx, int64_t n, int64_t q_begin,
                                   int64_t q_end, uint32_t *__restrict__ qlist, const uint32_t *__restrict__ orig_flag, uint32_t flag_value)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int64_t i = sidx[s];
    if (orig_flag ? orig_flag[i] == flag_value : (i >= q_begin && i < q_end)) qlist[qexcl[s]] = (uint32_t)s;
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


// ---- one thread per query, warp in lock step (default) ---------------------------------------------------------------
// Same traversal as knn_query_kernel (ring by ring around the query's own cell, whole z-runs of the new columns, two caps of
// the known ones) but the 32 queries of a warp -- neighbours in cell order -- walk it TOGETHER: every piece of the traversal
// is executed by all lanes for max-over-lanes iterations (a lane whose own range is shorter or empty idles), accepted
// candidates are appended to a per-lane pending list with a predicated store, and the lists are merged into the per-lane
// heaps by all lanes at the same points (when one list could overflow, and before each termination test).  In the diverging
// kernel 64 % of the stall samples sat in heap inserts executed by 11 of 32 lanes, each accept at its own moment (ncu source
// view, profiles/r01_v4_secondary_ncu.txt).  The periodic wrap of a piece is a constant shift whenever that is provably what
// every pair of the piece would do, else the per-pair comparison: a piece lies `off` cells from the query's cell along an
// axis (|off| <= ring); if it was reached without wrapping, every delta is below (ring + 1) cells, if it was reached by wrapping,
// every delta exceeds box - (ring + 1) cells in magnitude, and both are a full cell away from box/2 when 2 (ring + 2) <= G.
constexpr int kQPend = 16;        // 32 measured 3 % slower (staler thresholds, more local memory)

template <int KCAP, bool WANT_IDX, bool EXTERNAL, bool PER>
__global__ void __launch_bounds__(128) knn_lockstep_kernel(KnnArgs a)
{
    const unsigned FULL = 0xffffffffu;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < a.nq;                                 // invalid lanes shadow query 0 and never accept anything
    const int64_t tq = valid ? t : 0;
    const int64_t s = EXTERNAL ? 0 : (a.qlist ? (int64_t)a.qlist[tq] : tq);
    const double x = EXTERNAL ? a.qpos[3 * tq] : a.xs[s], y = EXTERNAL ? a.qpos[3 * tq + 1] : a.ys[s],
                 z = EXTERNAL ? a.qpos[3 * tq + 2] : a.zs[s];
    const KnnGrid &g = a.g;
    const int G = g.G;
    const double box = g.box, half_box = g.half_box;
    const double *__restrict__ xs = a.xs, *__restrict__ ys = a.ys, *__restrict__ zs = a.zs;
    const uint32_t *__restrict__ cstart = a.cstart;
    const int qc[3] = { cell_coord(g, x, 0), cell_coord(g, y, 1), cell_coord(g, z, 2) };
    const double xq[3] = { x, y, z };

    Heap<KCAP, WANT_IDX> hp;
    hp.k = a.k;
    for (int i = 0; i < a.k; ++i) {
        hp.d[i] = INFINITY;
        if (WANT_IDX) hp.id[i] = 0xffffffffu;
    }
    double kth = valid ? INFINITY : -INFINITY;
    double pend[kQPend];
    uint32_t pend_id[WANT_IDX ? kQPend : 1];
    int cnt = 0;
    bool done = !valid;

    auto flush = [&]() {
        const int mx = __reduce_max_sync(FULL, cnt);
        for (int i = 0; i < mx; ++i) {
            if (i < cnt) {
                const double d2 = pend[i];
                const uint32_t oj = WANT_IDX ? pend_id[i] : 0u;
                if (hp.less(d2, oj, hp.d[0], WANT_IDX ? hp.id[0] : 0u)) hp.replace_root(d2, oj);
            }
        }
        cnt = 0;
        kth = valid ? hp.d[0] : -INFINITY;
    };
    // candidates [jb, je) of THIS lane (empty for a lane that has nothing to visit here); all lanes iterate together
    auto scan_range = [&](uint32_t jb, uint32_t je, bool simple, double shx, double shy, double shz) {
        const int len = (int)(je - jb);
        const int maxlen = __reduce_max_sync(FULL, len);
        for (int t0 = 0; t0 < maxlen; t0 += 8) {
            const int t1 = min(t0 + 8, maxlen);
            for (int tt = t0; tt < t1; ++tt) {
                if (tt < len) {
                    const uint32_t j = jb + (uint32_t)tt;
                    double ex = xs[j] - x, ey = ys[j] - y, ez = zs[j] - z;
                    if (PER) {
                        if (simple) {
                            ex = AST_DADD(ex, shx); ey = AST_DADD(ey, shy); ez = AST_DADD(ez, shz);
                        } else {                                   // scipy's per-pair wrap, branch-free (x + 0.0 is x)
                            ex = AST_DADD(ex, ex < -half_box ? box : (ex > half_box ? -box : 0.0));
                            ey = AST_DADD(ey, ey < -half_box ? box : (ey > half_box ? -box : 0.0));
                            ez = AST_DADD(ez, ez < -half_box ? box : (ez > half_box ? -box : 0.0));
                        }
                    }
                    const double d2 = AST_DADD(AST_DADD(AST_DMUL(ex, ex), AST_DMUL(ey, ey)), AST_DMUL(ez, ez));
                    if (WANT_IDX ? d2 <= kth : d2 < kth) {
                        pend[cnt] = d2;
                        if (WANT_IDX) pend_id[cnt] = a.sidx[j];
                        ++cnt;
                    }
                }
            }
            if (__any_sync(FULL, cnt > kQPend - 8)) flush();
        }
    };
    // raw z-cells [z0, z1] (may stick out of [0, G)) of column (ccx, ccy); `on` = this lane visits the piece at all.
    // Always the same number of scan_range calls for every lane (two), so the warp stays converged.
    auto scan_z = [&](bool on, int ccx, int ccy, int z0, int z1, bool simple, double shx, double shy) {
        const uint32_t base = ((uint32_t)ccx * G + ccy) * G;
        uint32_t b0 = 0, e0 = 0, b1 = 0, e1 = 0;
        double sh0 = 0.0, sh1 = 0.0;
        bool smp = simple;
        if (on) {
            if (!PER) {
                z0 = max(z0, 0); z1 = min(z1, G - 1);
                if (z0 <= z1) { b0 = cstart[base + z0]; e0 = cstart[base + z1 + 1]; }
            } else if (z1 - z0 + 1 >= G) {
                b0 = cstart[base]; e0 = cstart[base + G]; smp = false;          // the run closes on itself: every cell once
            } else if (z0 < 0) {
                const int zt = min(z1, -1);
                b0 = cstart[base + z0 + G]; e0 = cstart[base + zt + G + 1]; sh0 = -box;
                if (z1 >= 0) { b1 = cstart[base]; e1 = cstart[base + z1 + 1]; }
            } else if (z1 >= G) {
                const int zb = max(z0, G);
                if (z0 < G) { b0 = cstart[base + z0]; e0 = cstart[base + G]; }
                b1 = cstart[base + zb - G]; e1 = cstart[base + z1 - G + 1]; sh1 = box;
            } else {
                b0 = cstart[base + z0]; e0 = cstart[base + z1 + 1];
            }
        }
        // `smp` may differ between lanes only through the closed-run case, which depends on ring and G alone: uniform
        scan_range(b0, e0, smp, shx, shy, sh0);
        if (PER) scan_range(b1, e1, smp, shx, shy, sh1);
    };

    for (int ring = 0;; ++ring) {
        const bool simple = PER && 2 * (ring + 2) <= G;
        for (int ox = -ring; ox <= ring; ++ox) {
            int ccx = qc[0] + ox;
            double shx = 0.0;
            bool okx = !done;
            if (PER) {
                if (2 * abs(ox) > G || (2 * abs(ox) == G && ox < 0)) continue;          // uniform: depends on ox and G only
                if (ccx < 0) { ccx += G; shx = -box; } else if (ccx >= G) { ccx -= G; shx = box; }
            } else if (ccx < 0 || ccx >= G) { okx = false; ccx = 0; }
            for (int oy = -ring; oy <= ring; ++oy) {
                int ccy = qc[1] + oy;
                double shy = 0.0;
                bool on = okx;
                if (PER) {
                    if (2 * abs(oy) > G || (2 * abs(oy) == G && oy < 0)) continue;
                    if (ccy < 0) { ccy += G; shy = -box; } else if (ccy >= G) { ccy -= G; shy = box; }
                } else if (ccy < 0 || ccy >= G) { on = false; ccy = 0; }
                if (abs(ox) == ring || abs(oy) == ring) {
                    scan_z(on, ccx, ccy, qc[2] - ring, qc[2] + ring, simple, shx, shy);           // new column: whole z-run
                } else if (!PER) {
                    scan_z(on, ccx, ccy, qc[2] - ring, qc[2] - ring, simple, shx, shy);           // known column: two new caps
                    scan_z(on, ccx, ccy, qc[2] + ring, qc[2] + ring, simple, shx, shy);
                } else {
                    // periodic: a cap is new only while the known z-run (2 ring - 1 cells) has not closed on itself
                    if (2 * ring <= G) scan_z(on, ccx, ccy, qc[2] - ring, qc[2] - ring, simple, shx, shy);
                    if (2 * ring + 1 <= G) scan_z(on, ccx, ccy, qc[2] + ring, qc[2] + ring, simple, shx, shy);
                }
            }
        }
        flush();
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
        done = done || all || (safe > 0.0 && kth < safe * safe);
        if (done) kth = -INFINITY;                                // a finished lane accepts nothing more (its heap is final)
        if (__all_sync(FULL, done)) break;
    }
    if (!valid) return;
    const int64_t row = EXTERNAL ? t : (int64_t)a.sidx[s] - a.q_begin;
    if (a.h_out) a.h_out[row] = sqrt(hp.d[0]);
    if (WANT_IDX) {
        const int k = a.k;
        for (int end = k - 1; end > 0; --end) {                 // heap-sort in place: ascending (d2, idx)
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

}  // namespace ast
#include "knn_select.cuh"
namespace ast {

// failed queries of the selection kernel -> ordered list of cell-ordered positions (flags already scanned exclusively)
__global__ void knn_fail_compact_kernel(const uint32_t *__restrict__ excl, int64_t n, const uint32_t *__restrict__ total, uint32_t *__restrict__ qlist)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const uint32_t e = excl[s], nx = s + 1 < n ? excl[s + 1] : *total;
    if (nx != e) qlist[e] = (uint32_t)s;
}

// Pending queries leave for another grid when this one does not suit their surroundings (multi-level search).  On a grid sized
// for the mean density a query inside a halo looks at thousands of candidates per cell of its 27-cell neighbourhood (NFW-clustered
// 256^3: a third of the queries made 88 % of all candidate evaluations), and a query in a void walks four or five rings of
// near-empty cells.  n27 = particles in the 3^3 cells around the query's own cell (clipped at the grid faces, no wrap: it is a
// cost estimate, not part of the result): n27 >= thr_hi -> level[particle] = 1 (fine grid), n27 < thr_lo -> 2 (coarse grid).
__global__ void knn_level_split_kernel(uint32_t *__restrict__ qflag, const uint64_t *__restrict__ sorted, const uint32_t *__restrict__ cstart,
                                       const uint32_t *__restrict__ sidx, int64_t n, int G, uint32_t thr_hi, uint32_t thr_lo,
                                       uint32_t *__restrict__ level, unsigned long long *__restrict__ count)
{
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t lv = 0;
    if (s < n && qflag[s]) {
        const uint32_t key = (uint32_t)(sorted[s] >> 32);
        const int cz = (int)(key % (uint32_t)G), cy = (int)((key / (uint32_t)G) % (uint32_t)G), cx = (int)(key / ((uint32_t)G * (uint32_t)G));
        const int z0 = max(cz - 1, 0), z1 = min(cz + 1, G - 1);
        uint32_t n27 = 0;
        for (int x = max(cx - 1, 0); x <= min(cx + 1, G - 1); ++x)
            for (int y = max(cy - 1, 0); y <= min(cy + 1, G - 1); ++y) {
                const uint32_t base = ((uint32_t)x * G + y) * G;
                n27 += cstart[base + z1 + 1] - cstart[base + z0];
            }
        lv = n27 >= thr_hi ? 1u : (n27 < thr_lo ? 2u : 0u);
        if (lv) { qflag[s] = 0u; level[sidx[s]] = lv; }
    }
    const unsigned b1 = __ballot_sync(0xffffffffu, lv == 1u), b2 = __ballot_sync(0xffffffffu, lv == 2u);
    if ((threadIdx.x & 31) == 0) {
        if (b1) atomicAdd(count, (unsigned long long)__popc(b1));
        if (b2) atomicAdd(count + 1, (unsigned long long)__popc(b2));
    }
}

// Order a query list by the crowding of the queries' surroundings (one class per factor 4 of the 27-cell count), keeping the
// cell order inside each class: the 32 queries of a lock-step warp then need rings and candidate counts of similar size (ncu on
// the clustered set: 16 of 32 threads per instruction when warps mix sparse and crowded queries).  NFW-clustered 256^3: 76.5 ms
// unordered, 72.9 with a class per factor 2, 71.6 per factor 4, 77.6 per factor sqrt(2) (finer classes scatter the warps).
__global__ void knn_qorder_key_kernel(const uint32_t *__restrict__ qlist, int64_t nq, const double *__restrict__ xs, const double *__restrict__ ys,
                                      const double *__restrict__ zs, KnnGrid g, const uint32_t *__restrict__ cstart, uint64_t *__restrict__ keys)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq) return;
    const uint32_t s = qlist[t];
    const int G = g.G;
    const int cx = cell_coord(g, xs[s], 0), cy = cell_coord(g, ys[s], 1), cz = cell_coord(g, zs[s], 2);
    const int z0 = max(cz - 1, 0), z1 = min(cz + 1, G - 1);
    uint32_t n27 = 0;
    for (int x = max(cx - 1, 0); x <= min(cx + 1, G - 1); ++x)
        for (int y = max(cy - 1, 0); y <= min(cy + 1, G - 1); ++y) {
            const uint32_t base = ((uint32_t)x * G + y) * G;
            n27 += cstart[base + z1 + 1] - cstart[base + z0];
        }
    const uint32_t cls = (uint32_t)min((31 - __clz((int)(n27 | 1u))) / 2, 15);
    keys[t] = ((uint64_t)cls << 32) | (uint64_t)s;
}
__global__ void knn_qorder_unpack_kernel(const uint64_t *__restrict__ keys, int64_t nq, uint32_t *__restrict__ qlist)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nq) qlist[t] = (uint32_t)keys[t];
}

struct KnnLayout {
    int G2, G0;                                                    // fine / coarse grid of the multi-level search (0: none)
    uint32_t *dense;                                               // per particle index: 1 = the query moved to the fine grid, 2 = to the coarse one
    int G;
    int64_t ncell, nq;
    uint64_t *ea, *eb;
    double *xs, *ys, *zs;
    uint32_t *sidx, *cbeg, *qflag, *qlist, *scan_tmp, *cell_tmp;   // cbeg: ncell + 1 counts -> exclusive scan = cstart
    uint32_t *kept;                                                // reach-limited build: original indices of the kept particles
    double *bbox_part;                                             // 256 x 6 partial min / max of the query positions
    unsigned long long *counter;                                   // kept total / unsafe-query count
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
    // k/3 per cell for the ring traversal (the explored cube of cells hugs the k-neighbour sphere more tightly; empty cells are
    // cheap).  The selection kernel stages the 8^3 cells around a block in a buffer of 1152 candidates: at 2.0 per cell a region
    // holds ~1030, at 2.2 most regions overflow and fall back (benchmarks/knn_celltarget_probe.py, whole call at 256^3:
    // 1.5: 14.4 ms, 1.6: 14.3, 1.75: 14.4 with 113 fallbacks, 1.9: 15.0, 2.0: 15.9, 2.2: 26.4, 2.5: 36.5).  1.75 keeps 28 % of
    // head room for denser regions and sweeps an eighth fewer candidates.
    const double m = p->cell_target > 0 ? p->cell_target : 1.75;
    double g = floor(cbrt((double)(p->n > 0 ? p->n : 1) / m));
    L.G = g < 1 ? 1 : (g > 1000 ? 1000 : (int)g);
    L.ncell = (int64_t)L.G * L.G * L.G;
    // fine grid for the queries in dense surroundings: cells up to 4 times smaller per axis, at most 1000^3 of them
    static const int fine_max = env_int("AST_KNN_FINE_FACTOR", 4);
    const int fine = L.G >= 16 ? (fine_max * L.G <= 1000 ? fine_max : 1000 / L.G) : 0;
    L.G2 = fine >= 2 ? fine * L.G : 0;
    L.G0 = L.G >= 16 ? L.G / 2 : 0;
    const int64_t ncell_max = L.G2 ? (int64_t)L.G2 * L.G2 * L.G2 : L.ncell;
    const int64_t n = p->n > 0 ? p->n : 1;
    L.nq = p->q_count > 0 ? p->q_count : p->n;
    Carver c(ws);
    L.ea = c.take<uint64_t>(n);
    L.eb = c.take<uint64_t>(n);
    L.xs = c.take<double>(n);
    L.ys = c.take<double>(n);
    L.zs = c.take<double>(n);
    L.sidx = c.take<uint32_t>(n);
    L.cbeg = c.take<uint32_t>(ncell_max + 1);
    L.cell_tmp = (uint32_t *)c.take<char>(scan_workspace_bytes<uint32_t>(ncell_max + 1));
    L.dense = c.take<uint32_t>(n);
    L.qflag = c.take<uint32_t>(n);
    L.qlist = c.take<uint32_t>(n);
    L.scan_tmp = (uint32_t *)c.take<char>(scan_workspace_bytes<uint32_t>(n));
    L.kept = c.take<uint32_t>(n);
    L.bbox_part = c.take<double>(256 * 6);
    L.counter = c.take<unsigned long long>(6);
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

template <int KCAP>
static void launch_lockstep(const KnnArgs &a, bool want_idx, cudaStream_t s)
{
    const unsigned nb = (unsigned)((a.nq + 127) / 128);
    const bool per = a.g.box > 0.0;
#define AST_LS(W, E) do { if (per) knn_lockstep_kernel<KCAP, W, E, true><<<nb, 128, 0, s>>>(a); \
                         else knn_lockstep_kernel<KCAP, W, E, false><<<nb, 128, 0, s>>>(a); } while (0)
    if (a.qpos) AST_LS(true, true);
    else if (want_idx) AST_LS(true, false);
    else AST_LS(false, false);
#undef AST_LS
}

// builds the cell list of `pos` in the workspace (steps 1 and 2) and fills the grid / array part of KnnArgs
// (kept != nullptr: from the n_build particles listed there only)
static int knn_build(const ast_knn_params *p, const double *pos, const KnnLayout &L, bool subset, int64_t q_begin, int64_t q_end,
                     cudaStream_t s, KnnArgs &a, int64_t n_build = -1, const uint32_t *kept = nullptr, const uint32_t *orig_flag = nullptr,
                     const uint64_t **sorted_out = nullptr, uint32_t flag_value = 1u)
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
    const int64_t n = n_build >= 0 ? n_build : p->n;
    const unsigned nb = (unsigned)((n + 255) / 256);
    knn_key_kernel<<<nb, 256, 0, s>>>(pos, n, g, kept, L.ea);
    int in_b = 0;
    AST_CUDA_TRY(radix_sort_u64(L.ea, L.eb, n, 32, ceil_log2_u64((uint64_t)L.ncell), L.sort_ws, s, &in_b));
    const uint64_t *sorted = in_b ? L.eb : L.ea;
    AST_CUDA_TRY(cudaMemsetAsync(L.cbeg, 0, sizeof(uint32_t) * (L.ncell + 1), s));
    knn_gather_kernel<<<nb, 256, 0, s>>>(pos, sorted, n, L.xs, L.ys, L.zs, L.sidx, L.cbeg, q_begin, q_end,
                                         subset ? L.qflag : nullptr, orig_flag, flag_value);
    AST_CUDA_TRY(scan_exclusive<uint32_t>(L.cbeg, L.ncell + 1, L.cell_tmp, nullptr, s));      // counts -> cstart
    if (subset) {
        AST_CUDA_TRY(scan_exclusive<uint32_t>(L.qflag, n, L.scan_tmp, nullptr, s));
        knn_compact_kernel<<<nb, 256, 0, s>>>(L.qflag, L.sidx, n, q_begin, q_end, L.qlist, orig_flag, flag_value);
    }
    if (sorted_out) *sorted_out = sorted;
    a.g = g;
    a.xs = L.xs; a.ys = L.ys; a.zs = L.zs; a.sidx = L.sidx; a.cstart = L.cbeg;
    a.qlist = subset ? L.qlist : nullptr;
    a.qpos = nullptr;
    a.k = p->k;
    AST_CUDA_TRY(cudaGetLastError());
    return AST_OK;
}

// sorts a.qlist (a.nq entries) by crowding class; L.ea / L.eb are free once the cell list is built
static int order_queries(const KnnLayout &L, const KnnArgs &a, cudaStream_t s)
{
    static const bool on = env_flag("AST_KNN_QORDER", true);
    if (!on || !a.qlist || a.nq < 4096) return AST_OK;
    const unsigned nb = (unsigned)((a.nq + 255) / 256);
    knn_qorder_key_kernel<<<nb, 256, 0, s>>>(a.qlist, a.nq, a.xs, a.ys, a.zs, a.g, a.cstart, L.ea);
    int in_b = 0;
    AST_CUDA_TRY(radix_sort_u64(L.ea, L.eb, a.nq, 32, 4, L.sort_ws, s, &in_b));
    knn_qorder_unpack_kernel<<<nb, 256, 0, s>>>(in_b ? L.eb : L.ea, a.nq, const_cast<uint32_t *>(a.qlist));
    AST_KERNEL_CHECK(s, "knn_qorder");
    return AST_OK;
}

// ---- selection fast path (h only) --------------------------------------------------------------------------------------
static bool select_usable(const ast_knn_params *p, const KnnLayout &L, bool want_lists, int64_t n_build)
{
    static const bool enabled = env_flag("AST_KNN_SELECT", true);
    if (!enabled || want_lists || (p->flags & (AST_KNN_DIVERGING | AST_KNN_NO_SELECT))) return false;
    if (L.G < 2 * (kSelBS + 2 * kSelR) || n_build < 4096) return false;        // wrapped region pieces must stay disjoint / minimum image
    // the K-th neighbour has to lie within R = 2 cells (+ the query's offset in its cell) for the answer to be verifiable: with m
    // particles per occupied cell it sits at (K / (4.19 m))^(1/3) cells.  Beyond ~2.1 nearly every query would fall through to
    // the lock-step kernel after a wasted sweep (K > 67 at the default 1.75 per cell); a caller's cell_target below that is the slab
    // decomposition's "1.75 per occupied cell", above it a true density (whose regions then exceed the staging buffer anyway)
    const double m_occ = p->cell_target > 1.75 ? p->cell_target : 1.75;
    if (cbrt((double)p->k / (4.19 * m_occ)) > 2.1) return false;
    if (!(p->box > 0.0)) {
        double lo = INFINITY, hi = 0.0;
        for (int c = 0; c < 3; ++c) { const double e = p->hi[c] - p->lo[c]; if (!(e > 0.0)) return false; lo = e < lo ? e : lo; hi = e > hi ? e : hi; }
        if (hi > 1.5 * lo) return false;                                           // the float32 error bound assumes near-cubic cells
    }
    return true;
}

// runs the selection kernel for every query; *need_lockstep = some queries were flagged (a.qlist / a.nq then describe them)
static int launch_select(const ast_knn_params *p, const KnnLayout &L, KnnArgs &a, int64_t n_build, const uint64_t *sorted, cudaStream_t s,
                         bool *need_lockstep, int64_t *n_dense, int64_t *n_sparse)
{
    SelParams sp;
    double cs_min = a.g.cs[0] < a.g.cs[1] ? a.g.cs[0] : a.g.cs[1];
    cs_min = cs_min < a.g.cs[2] ? cs_min : a.g.cs[2];
    const double dref = ((double)kSelR + 0.5) * cs_min;
    for (int c = 0; c < 3; ++c) sp.sc[c] = a.g.cs[c] / dref;
    sp.dref2 = dref * dref;
    // a sphere of R cells holds 4.19 R^3 / W^3 of the region's particles when they are spread evenly: demand 0.9 K of them;
    // a region that does not fit the staging buffer is so dense that the ring traversal (27 cells, not 512) is the cheaper search
    const double wreg = (double)(kSelBS + 2 * kSelR);
    sp.m_min = (uint32_t)(0.9 * (double)p->k * wreg * wreg * wreg / (4.19 * kSelR * kSelR * kSelR));
    sp.m_max = kSelChunk;
    sp.fail = L.qflag;
    const int nbk = (L.G + kSelBS - 1) / kSelBS;
    const unsigned nblocks = (unsigned)nbk * nbk * nbk;
    const size_t smem = sizeof(SelShared);
    // (set on every call: the attribute belongs to the current device, and a process may drive several)
    AST_CUDA_TRY(cudaFuncSetAttribute(knn_select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AST_CUDA_TRY(cudaFuncSetAttribute(knn_select_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (a.g.box > 0.0) knn_select_kernel<true><<<nblocks, kSelThreads, smem, s>>>(a, sp);
    else knn_select_kernel<false><<<nblocks, kSelThreads, smem, s>>>(a, sp);
    AST_KERNEL_CHECK(s, "knn_select_kernel");
    uint32_t n_fail = 0;
    unsigned long long moved[2] = { 0, 0 };
    uint32_t *total = reinterpret_cast<uint32_t *>(L.counter + 2);
    // measured on the NFW-clustered 256^3 set (ms per call): one level 103.4; fine grid for n27 >= 60: 96.3, 100: 86.6, 200: 78.7,
    // 400: 76.5, 1000: 76.7 (cells 2 / 3 / 4 times smaller: 79.0 / 74.7 / 77.2).  A COARSE grid (cells twice as large) for the queries
    // of near-empty surroundings loses: n27 < 20: 88.4, 40: 87.7, 150: 93.7 -- walking rings of empty cells is cheaper than the
    // eightfold candidates of every occupied one; off unless AST_KNN_SPARSE_N27 is set.
    static const int thr_hi = env_int("AST_KNN_DENSE_N27", 400), thr_lo = env_int("AST_KNN_SPARSE_N27", 0);
    const bool multi = (L.G2 > 0 && thr_hi > 0) || (L.G0 > 0 && thr_lo > 0);
    if (multi) {
        AST_CUDA_TRY(cudaMemsetAsync(L.dense, 0, sizeof(uint32_t) * (size_t)(p->n > 0 ? p->n : 1), s));
        AST_CUDA_TRY(cudaMemsetAsync(L.counter + 3, 0, 2 * sizeof(unsigned long long), s));
        knn_level_split_kernel<<<(unsigned)((n_build + 255) / 256), 256, 0, s>>>(L.qflag, sorted, a.cstart, a.sidx, n_build, L.G,
                                                                                 L.G2 > 0 && thr_hi > 0 ? (uint32_t)thr_hi : 0xffffffffu,
                                                                                 L.G0 > 0 && thr_lo > 0 ? (uint32_t)thr_lo : 0u, L.dense, L.counter + 3);
        AST_KERNEL_CHECK(s, "knn_level_split_kernel");
    }
    AST_CUDA_TRY(scan_exclusive<uint32_t>(L.qflag, n_build, L.scan_tmp, total, s));
    AST_CUDA_TRY(cudaMemcpyAsync(&n_fail, total, sizeof n_fail, cudaMemcpyDeviceToHost, s));
    if (multi) AST_CUDA_TRY(cudaMemcpyAsync(moved, L.counter + 3, sizeof moved, cudaMemcpyDeviceToHost, s));
    AST_CUDA_TRY(cudaStreamSynchronize(s));
    *n_dense = (int64_t)moved[0];
    *n_sparse = (int64_t)moved[1];
    *need_lockstep = n_fail > 0;
    if (n_fail) {
        knn_fail_compact_kernel<<<(unsigned)((n_build + 255) / 256), 256, 0, s>>>(L.qflag, n_build, total, L.qlist);
        a.qlist = L.qlist;
        a.nq = (int64_t)n_fail;
    }
    static const bool verbose = env_flag("AST_KNN_VERBOSE", false);
    if (verbose) fprintf(stderr, "[ast_knn_h] selection kernel: %u queries left to the lock-step kernel on this grid (G = %d), %llu to the fine grid (G = %d), %llu to the coarse grid (G = %d)\n", n_fail, L.G, moved[0], L.G2, moved[1], L.G0);
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
    const int64_t nq = q_end - q_begin;
    const bool want = idx_out != nullptr || dist_out != nullptr;
    // Reach-limited build (query subsets only, i.e. the per-rank call of a multi-GPU job): build the cell list from the
    // particles within `margin` of the region the queries occupy, answer, and verify that every K-th distance stayed
    // below the margin; otherwise double the margin and repeat (it ends with the full build).  AST_KNN_FULL_BUILD skips it.
    bool limited = subset && !(p->flags & AST_KNN_FULL_BUILD) && nq * 2 <= n;
    double margin_cells = 4.0;
    for (;;) {
        int64_t n_build = -1;
        double margin = 0.0;
        if (limited) {
            KnnReach r;
            r.box = p->box > 0.0 ? p->box : 0.0;
            double cs_max = 0.0;
            for (int c = 0; c < 3; ++c) {
                double ext = p->box > 0.0 ? p->box : (p->hi[c] - p->lo[c]);
                if (!(ext > 0.0)) ext = 1.0;
                cs_max = ext / (double)L.G > cs_max ? ext / (double)L.G : cs_max;
            }
            margin = r.margin = margin_cells * cs_max;
            const unsigned nbq = (unsigned)((nq + 255) / 256 < 256 ? (nq + 255) / 256 : 256);
            bool whole = true;                                   // the reach covers the whole box on every axis
            if (r.box > 0.0) {
                unsigned long long masks[3];
                AST_CUDA_TRY(cudaMemsetAsync(L.counter, 0, 4 * sizeof(unsigned long long), s));
                knn_binmask_kernel<<<nbq, 256, 0, s>>>(pos, q_begin, q_end, r.box, L.counter + 1);
                AST_CUDA_TRY(cudaMemcpyAsync(masks, L.counter + 1, sizeof masks, cudaMemcpyDeviceToHost, s));
                AST_CUDA_TRY(cudaStreamSynchronize(s));
                const double w = r.box / 64.0;
                for (int c = 0; c < 3; ++c) {
                    // largest circular run of empty bins; the queries lie in its complement
                    int best_len = 0, best_start = 0;
                    for (int b0 = 0; b0 < 64; ++b0) {
                        if ((masks[c] >> b0) & 1ull) continue;
                        int len = 0;
                        while (len < 64 && !((masks[c] >> ((b0 + len) & 63)) & 1ull)) ++len;
                        if (len > best_len) { best_len = len; best_start = b0; }
                    }
                    if (best_len == 0) { r.lo[c] = 0.0; r.len[c] = r.box; }
                    else { r.lo[c] = (double)((best_start + best_len) & 63) * w; r.len[c] = (double)(64 - best_len) * w; }
                    whole = whole && r.len[c] + 2.0 * r.margin >= r.box;
                }
            } else {
                double part[256 * 6];
                knn_bbox_kernel<<<nbq, 256, 0, s>>>(pos, q_begin, q_end, L.bbox_part);
                AST_CUDA_TRY(cudaMemcpyAsync(part, L.bbox_part, sizeof(double) * 6 * nbq, cudaMemcpyDeviceToHost, s));
                AST_CUDA_TRY(cudaStreamSynchronize(s));
                for (int c = 0; c < 3; ++c) {
                    double lo = part[c], hi = part[3 + c];
                    for (unsigned b = 1; b < nbq; ++b) { lo = part[6 * b + c] < lo ? part[6 * b + c] : lo; hi = part[6 * b + 3 + c] > hi ? part[6 * b + 3 + c] : hi; }
                    r.lo[c] = lo; r.len[c] = hi - lo;
                    whole = whole && lo - r.margin <= p->lo[c] && hi + r.margin >= p->hi[c];
                }
            }
            if (whole) {
                limited = false;
            } else {
                const unsigned nbn = (unsigned)((n + 255) / 256);
                uint32_t total = 0;
                knn_reach_flag_kernel<<<nbn, 256, 0, s>>>(pos, n, r, L.qflag);
                AST_CUDA_TRY(scan_exclusive<uint32_t>(L.qflag, n, L.scan_tmp, reinterpret_cast<uint32_t *>(L.counter), s));
                knn_reach_compact_kernel<<<nbn, 256, 0, s>>>(pos, n, r, L.qflag, L.kept);
                AST_CUDA_TRY(cudaMemcpyAsync(&total, L.counter, sizeof total, cudaMemcpyDeviceToHost, s));
                AST_CUDA_TRY(cudaStreamSynchronize(s));
                if ((int64_t)total * 10 > n * 8) limited = false;          // nearly everything is in reach: nothing to gain
                else n_build = (int64_t)total;
            }
        }
        KnnArgs a;
        const int64_t n_fast = limited ? n_build : n;
        const bool fast = select_usable(p, L, want, n_fast);
        const uint64_t *sorted = nullptr;
        rc = knn_build(p, pos, L, subset && !fast, q_begin, q_end, s, a, limited ? n_build : -1, limited ? L.kept : nullptr, nullptr, &sorted);
        if (rc) return rc;
        a.nq = nq;
        a.q_begin = q_begin;
        a.h_out = h_out; a.idx_out = idx_out; a.dist_out = dist_out;
        if (p->flags & AST_KNN_DIVERGING) {
            if (p->k <= 32) launch_query<32>(a, want, s);
            else if (p->k <= 48) launch_query<48>(a, want, s);
            else if (p->k <= 64) launch_query<64>(a, want, s);
            else launch_query<128>(a, want, s);
        } else {
            bool run_lockstep = true;
            int64_t n_moved[2] = { 0, 0 };
            if (fast) {
                // selection kernel over blocks of cells (knn_select.cuh); what it cannot verify goes to the lock-step kernel,
                // on this grid or -- queries in dense / near-empty surroundings -- on a finer / coarser one
                rc = launch_select(p, L, a, n_fast, sorted, s, &run_lockstep, &n_moved[0], &n_moved[1]);
                if (rc) return rc;
            }
            if (run_lockstep) {
                if (fast) { rc = order_queries(L, a, s); if (rc) return rc; }
                if (p->k <= 32) launch_lockstep<32>(a, want, s);
                else if (p->k <= 48) launch_lockstep<48>(a, want, s);
                else if (p->k <= 64) launch_lockstep<64>(a, want, s);
                else launch_lockstep<128>(a, want, s);
            }
            for (int lv = 0; lv < 2; ++lv) {
                if (n_moved[lv] <= 0) continue;
                // another level: the same particles on cells L.G2 / L.G times smaller (lv 0) or twice as large (lv 1); the buffers
                // of the previous level are free again once its kernels have run (same stream); queries = particles flagged lv + 1
                KnnLayout Lx = L;
                Lx.G = lv == 0 ? L.G2 : L.G0;
                Lx.ncell = (int64_t)Lx.G * Lx.G * Lx.G;
                KnnArgs ax;
                rc = knn_build(p, pos, Lx, true, q_begin, q_end, s, ax, limited ? n_build : -1, limited ? L.kept : nullptr, L.dense, nullptr,
                               (uint32_t)(lv + 1));
                if (rc) return rc;
                ax.nq = n_moved[lv];
                ax.q_begin = q_begin;
                ax.h_out = h_out; ax.idx_out = idx_out; ax.dist_out = dist_out;
                rc = order_queries(Lx, ax, s);
                if (rc) return rc;
                if (p->k <= 32) launch_lockstep<32>(ax, want, s);
                else if (p->k <= 48) launch_lockstep<48>(ax, want, s);
                else if (p->k <= 64) launch_lockstep<64>(ax, want, s);
                else launch_lockstep<128>(ax, want, s);
            }
        }
        AST_CUDA_TRY(cudaGetLastError());
        if (!limited) break;
        unsigned long long unsafe = 0;
        AST_CUDA_TRY(cudaMemsetAsync(L.counter, 0, sizeof(unsigned long long), s));
        knn_unsafe_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, s>>>(h_out, nq, margin, L.counter);
        AST_CUDA_TRY(cudaMemcpyAsync(&unsafe, L.counter, sizeof unsafe, cudaMemcpyDeviceToHost, s));
        AST_CUDA_TRY(cudaStreamSynchronize(s));
        if (unsafe == 0) break;
        margin_cells *= 2.0;                                     // some neighbourhood may reach beyond the margin: widen
    }
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
    a.h_out = nullptr; a.idx_out = idx_out; a.dist_out = dist_out;
    if (p->flags & AST_KNN_DIVERGING) {
        if (p->k <= 32) launch_query<32>(a, true, s);
        else if (p->k <= 48) launch_query<48>(a, true, s);
        else if (p->k <= 64) launch_query<64>(a, true, s);
        else launch_query<128>(a, true, s);
    } else {
        if (p->k <= 32) launch_lockstep<32>(a, true, s);
        else if (p->k <= 48) launch_lockstep<48>(a, true, s);
        else if (p->k <= 64) launch_lockstep<64>(a, true, s);
        else launch_lockstep<128>(a, true, s);
    }
    AST_CUDA_TRY(cudaGetLastError());
    return AST_OK;
}
