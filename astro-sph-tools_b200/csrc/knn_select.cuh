// knn_select.cuh -- K-th neighbour distance by SELECTION instead of a heap (the h-only path of ast_knn_h), sm_100a.
//
// The smoothing length needs only the K-th smallest distance, not the neighbour list, so nothing has to be kept sorted.
// Round 1's kernel kept a 48-entry heap per query in local memory: 49 GB of DRAM traffic for 0.5 GB of input and half of
// all issue slots idle.  Here a CTA takes a BLOCK of BS^3 cells; its queries are the particles of the block,
// and all of them see the same candidates: the particles of the block grown by R cells on every side, staged ONCE per CTA
// in shared memory as float4 (v, |v|^2), v = float32 coordinates relative to the region's centre in units of the reach
// Dref = (R + 1/2) cells, so that s = |v_c - v_q|^2 = |v_c|^2 - 2 v_c.v_q + |v_q|^2 is three FFMA and one saturating FADD
// and s = 1 is the reach.  A warp takes up to 32 queries of one QUADRANT of the block (2 x 2 columns of cells, full z-run) and
// sweeps only the columns within R cells of that quadrant (6 x 6 of the 8 x 8 staged columns = 6 contiguous runs of the staged
// list) twice, with broadcast LDS.128 reads, fully converged, branch-free, no per-thread memory except its own shared-memory
// column; after the one barrier behind the staging the warps of a CTA run independently:
//   pass A  s -> histogram of 32 bins over [0, 1) (one shared-memory atomic per candidate inside the reach); the bin b* in
//           which the running count reaches K gives the float band  [b*/32 - delta, (b*+1)/32 + delta];
//   pass B  s again: candidates inside the band are remembered (a handful per query);
//   exact   the remembered candidates are re-evaluated in float64 with scipy's arithmetic (raw coordinates, per-pair
//           periodic wrap, (dx*dx + dy*dy) + dz*dz without FMA) and classified against the float64 edges TL / TH:
//           below TL -> counted, inside [TL, TH] -> kept in a small exact list.  With count = (histogram count below bin b*,
//           corrected for the remembered candidates) : if  count < K <= count + len(list)  the answer is the (K - count)-th
//           smallest of the list -- the same bits as a full float64 search, because the partition {< TL} / [TL, TH] / {> TH}
//           is exact: float32 decides only candidates further than delta from an edge, and delta bounds the float32 error
//           (below).  The answer is accepted if it lies closer than the nearest face of the region (nothing outside can
//           beat it) -- the termination test of the ring traversal.
// A query that fails any check (sparse neighbourhood, over-full bin in a dense cell, K-th beyond the reach) is flagged and
// answered by knn_lockstep_kernel afterwards; the fast path never returns an unverified value.  Blocks whose region holds too
// few particles for K neighbours inside the reach, or so many that the ring traversal (which looks at 27 cells, not 512) is
// cheaper, are flagged without being swept.
//
// delta: |v| <= (W / 2) cs / Dref <= 2.4 (W <= 8 cells, cell anisotropy limited to 1.5 by the caller), |v|^2 <= 17.3.  A staged
// coordinate is off by <= 2^-25 * 2.4 = 7e-8, which moves s by <= 2 * 1.7 * 7e-8 * 6 = 1.5e-6; |v|^2 and the three FFMA and the
// FADD round intermediates below 35: 5 * 2^-25 * 35 = 5.2e-6.  Total < 7e-6 = 2.2e-4 bins; delta = 4e-3 bins = 1.25e-4 is
// eighteen times that.  A candidate lands inside a delta margin with probability ~1e-2 of a bin population.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ast {

constexpr int kSelWarps = 6;
constexpr int kSelThreads = 32 * kSelWarps;
constexpr int kSelNB = 32;            // histogram bins over s in [0, 1); row 32 collects everything beyond the reach
constexpr int kSelChunk = 1152;       // staged candidates (a region that holds more goes to the lock-step kernel)
constexpr int kSelList = 16;          // remembered / exact candidates per query
constexpr int kSelBS = 4, kSelR = 2;  // block of 4^3 cells, 2 cells of reach: 8 x 8 columns, quadrants see 6 x 6 of them
constexpr int kSelW = kSelBS + 2 * kSelR;
constexpr int kSelSeg = 2 * kSelW * kSelW;
constexpr float kSelDelta = 4e-3f / (float)kSelNB;   // in units of s
static_assert(kSelSeg <= kSelThreads, "one thread per segment");

struct SelParams {
    double sc[3];                     // cell units -> reach units per axis (cs / Dref)
    double dref2;                     // Dref^2: s -> squared distance
    uint32_t m_min, m_max;            // blocks whose region holds fewer / more particles go to the lock-step kernel unswept
    uint32_t *fail;                   // per cell-ordered particle: 1 = a query the fast path could not answer
};

struct SelShared {
    float4 cand[kSelChunk];
    uint32_t candj[kSelChunk];
    union WarpArea {                  // the histogram is dead once the thresholds are known
        unsigned hist[(kSelNB + 1) * 32];
        double exact[kSelList * 32];
    } wa[kSelWarps];
    unsigned short pend[kSelWarps][kSelList * 32];
    uint32_t seg_begin[kSelSeg], seg_off[kSelSeg + 1];
    uint32_t qcol_begin[16], qcol_cnt[16];          // the block's own columns, quadrant-major
    uint32_t warp_tot[kSelWarps];
};
static_assert(kSelList * sizeof(double) <= (kSelNB + 1) * sizeof(unsigned), "exact list must fit over the histogram");

template <bool PER>
__device__ __forceinline__ float4 sel_stage(const KnnGrid &g, const SelParams &sp, const int rlo[3], const int W[3], double x, double y, double z)
{
    const double q[3] = { x, y, z };
    float v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double u = (q[c] - g.lo[c]) * g.inv_cs[c] - (double)rlo[c];
        if (PER) { if (u < 0.0) u += (double)g.G; else if (u >= (double)g.G) u -= (double)g.G; }
        v[c] = (float)((u - 0.5 * (double)W[c]) * sp.sc[c]);
    }
    return make_float4(v[0], v[1], v[2], fmaf(v[2], v[2], fmaf(v[1], v[1], v[0] * v[0])));
}

template <bool PER>
__global__ void __launch_bounds__(kSelThreads) knn_select_kernel(KnnArgs a, SelParams sp)
{
    extern __shared__ __align__(16) unsigned char sel_smem_raw[];
    SelShared &S = *reinterpret_cast<SelShared *>(sel_smem_raw);
    constexpr int BS = kSelBS, R = kSelR;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const KnnGrid &g = a.g;
    const int G = g.G;
    const uint32_t *__restrict__ cstart = a.cstart;
    const int nbk = (G + BS - 1) / BS;
    // block coordinates, z fastest
    int b = blockIdx.x;
    const int kz = b % nbk; b /= nbk;
    const int ky = b % nbk; b /= nbk;
    const int kx = b;
    const int b0[3] = { kx * BS, ky * BS, kz * BS };
    const int b1[3] = { min(b0[0] + BS, G) - 1, min(b0[1] + BS, G) - 1, min(b0[2] + BS, G) - 1 };      // inclusive
    const int rlo[3] = { b0[0] - R, b0[1] - R, b0[2] - R };
    const int W[3] = { b1[0] - b0[0] + 1 + 2 * R, b1[1] - b0[1] + 1 + 2 * R, b1[2] - b0[2] + 1 + 2 * R };

    // ---- segment table: one z-run (two pieces when it wraps) per region column, column index = rx * 8 + ry ---------------
    {
        const int col = tid >> 1, piece = tid & 1;
        const int rx = col / kSelW, ry = col % kSelW;
        uint32_t sb = 0, se = 0;
        if (tid < kSelSeg && rx < W[0] && ry < W[1]) {
            int cx = rlo[0] + rx, cy = rlo[1] + ry;
            bool on = true;
            if (PER) { cx = cx < 0 ? cx + G : (cx >= G ? cx - G : cx); cy = cy < 0 ? cy + G : (cy >= G ? cy - G : cy); }
            else on = cx >= 0 && cx < G && cy >= 0 && cy < G;
            if (on) {
                const uint32_t base = ((uint32_t)cx * G + cy) * G;
                int z0 = rlo[2], z1 = b1[2] + R;
                if (!PER) {
                    z0 = max(z0, 0); z1 = min(z1, G - 1);
                    if (piece == 0 && z0 <= z1) { sb = cstart[base + z0]; se = cstart[base + z1 + 1]; }
                } else if (z0 < 0) {                     // cells [z0 + G, G) then [0, z1]   (G >= 2 W: the pieces are disjoint)
                    if (piece == 0) { sb = cstart[base + z0 + G]; se = cstart[base + G]; }
                    else { sb = cstart[base]; se = cstart[base + z1 + 1]; }
                } else if (z1 >= G) {                    // cells [z0, G) then [0, z1 - G]
                    if (piece == 0) { sb = cstart[base + z0]; se = cstart[base + G]; }
                    else { sb = cstart[base]; se = cstart[base + z1 - G + 1]; }
                } else if (piece == 0) { sb = cstart[base + z0]; se = cstart[base + z1 + 1]; }
            }
        }
        const uint32_t len = se - sb;
        uint32_t inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) S.warp_tot[warp] = inc;
        if (tid < kSelSeg) S.seg_begin[tid] = sb;
        // the block's own columns, quadrant-major: entry (qx * 2 + qy) * 4 + ax * 2 + ay  <->  column (2 qx + ax, 2 qy + ay)
        if (tid >= kSelThreads - 16) {
            const int e = tid - (kSelThreads - 16);
            const int ix = 2 * (e >> 3) + ((e >> 1) & 1), iy = 2 * ((e >> 2) & 1) + (e & 1);
            uint32_t qb = 0, qe = 0;
            if (b0[0] + ix <= b1[0] && b0[1] + iy <= b1[1]) {
                const uint32_t base = ((uint32_t)(b0[0] + ix) * G + (b0[1] + iy)) * G;
                qb = cstart[base + b0[2]]; qe = cstart[base + b1[2] + 1];
            }
            S.qcol_begin[e] = qb;
            S.qcol_cnt[e] = qe - qb;
        }
        __syncthreads();
        if (tid < kSelSeg) {
            uint32_t off = inc - len;
            for (int w = 0; w < warp; ++w) off += S.warp_tot[w];
            S.seg_off[tid] = off;
            if (tid == kSelSeg - 1) S.seg_off[kSelSeg] = off + len;
        }
        __syncthreads();
    }
    const uint32_t M = S.seg_off[kSelSeg];
    uint32_t qcnt[4];                                            // queries per quadrant
    uint32_t Q = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) { qcnt[d] = S.qcol_cnt[4 * d] + S.qcol_cnt[4 * d + 1] + S.qcol_cnt[4 * d + 2] + S.qcol_cnt[4 * d + 3]; Q += qcnt[d]; }
    if (Q == 0) return;
    const bool swept = M >= sp.m_min && M <= sp.m_max;          // else: too sparse for K neighbours inside the reach / too dense to sweep
    const double *__restrict__ xs = a.xs, *__restrict__ ys = a.ys, *__restrict__ zs = a.zs;
    if (swept) {
        for (uint32_t i = tid; i < M; i += kSelThreads) {
            int lo = 0, hi = kSelSeg;                            // last segment with seg_off <= i
#pragma unroll
            for (int it = 0; it < 7; ++it) { const int mid = (lo + hi) >> 1; if (S.seg_off[mid] <= i) lo = mid; else hi = mid; }
            const uint32_t j = S.seg_begin[lo] + (i - S.seg_off[lo]);
            S.cand[i] = sel_stage<PER>(g, sp, rlo, W, xs[j], ys[j], zs[j]);
            S.candj[i] = j;
        }
    }
    __syncthreads();                                             // the only barrier after the tables: warps run on their own from here

    const double box = g.box, half_box = g.half_box;
    const int K = a.k;
    unsigned char *hcol = reinterpret_cast<unsigned char *>(S.wa[warp].hist + lane);
    double *ecol = S.wa[warp].exact + lane;
    unsigned short *pcol = S.pend[warp] + lane;
    const int n_items = (int)((qcnt[0] + 31) / 32 + (qcnt[1] + 31) / 32 + (qcnt[2] + 31) / 32 + (qcnt[3] + 31) / 32);
    for (int item = warp; item < n_items; item += kSelWarps) {
        // item -> (quadrant, chunk of 32 queries)
        int qd = 0, ch = item;
#pragma unroll
        for (int d = 0; d < 3; ++d) { const int ni = (int)((qcnt[d] + 31) / 32); if (qd == d && ch >= ni) { ch -= ni; qd = d + 1; } }
        const int qx = qd >> 1, qy = qd & 1;
        const uint32_t qi = (uint32_t)ch * 32 + lane;
        const bool in_block = qi < qcnt[qd];
        uint32_t s = 0;
        if (in_block) {
            uint32_t r = qi;
            int e = 4 * qd;
            while (r >= S.qcol_cnt[e]) { r -= S.qcol_cnt[e]; ++e; }
            s = S.qcol_begin[e] + r;
        }
        int64_t row = 0;
        bool active = false;
        if (in_block) {
            row = (int64_t)a.sidx[s] - a.q_begin;
            active = row >= 0 && row < a.nq;
        }
        if (!swept) {
            if (in_block) sp.fail[s] = active ? 1u : 0u;
            continue;
        }
        if (!__any_sync(0xffffffffu, active)) {                  // nothing to answer here (cells that hold ghosts only)
            if (in_block) sp.fail[s] = 0u;
            continue;
        }
        const double x = active ? xs[s] : 0.0, y = active ? ys[s] : 0.0, z = active ? zs[s] : 0.0;
        const float4 qv = sel_stage<PER>(g, sp, rlo, W, x, y, z);
        const float m2x = -2.f * qv.x, m2y = -2.f * qv.y, m2z = -2.f * qv.z, q2 = qv.w;
        // distance to the nearest open face of the swept sub-region: the quadrant's cells grown by R in x and y, the block's in z
        double safe2 = INFINITY;
        if (active) {
            const double xq[3] = { x, y, z };
            const int s0[3] = { b0[0] + 2 * qx, b0[1] + 2 * qy, b0[2] };
            const int s1[3] = { min(s0[0] + 1, b1[0]), min(s0[1] + 1, b1[1]), b1[2] };
            double dmin = INFINITY;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int qc = cell_coord(g, xq[c], c);
                const double frac = xq[c] - (g.lo[c] + (double)qc * g.cs[c]);
                const int below = qc - s0[c] + R, above = s1[c] + R - qc;             // whole cells between the query's cell and the faces
                const bool lo_open = PER || s0[c] - R > 0, hi_open = PER || s1[c] + R < G - 1;
                if (lo_open) dmin = fmin(dmin, frac + (double)below * g.cs[c]);
                if (hi_open) dmin = fmin(dmin, (g.cs[c] - frac) + (double)above * g.cs[c]);
            }
            const double safe = dmin - 1e-9 * g.cs[0];
            safe2 = dmin < INFINITY ? (safe > 0.0 ? safe * safe : 0.0) : INFINITY;
        }
        // the six runs of the staged list this quadrant sweeps: columns rx in [2 qx, 2 qx + 6), ry in [2 qy, 2 qy + 6)
        const int run0 = 2 * ((2 * qx) * kSelW + 2 * qy);         // seg_off index of run r: run0 + 2 kSelW r, its end 12 entries on
        __syncwarp();                                            // the previous item's exact lists are dead
#pragma unroll
        for (int i = 0; i <= kSelNB; ++i) *reinterpret_cast<unsigned *>(hcol + i * 128) = 0u;

        // ---- pass A: histogram of s.  bin = floor(32 s): adding 2^11 with round-toward-zero leaves floor(s * 2^12) in the
        // mantissa, so the masked bits 7..12 are already the byte offset bin * 128 of the lane's column (bin 32 when s = 1)
        {
            auto deposit = [&](const float4 c) {
                const float sd = __saturatef(fmaf(m2x, c.x, fmaf(m2y, c.y, fmaf(m2z, c.z, c.w))) + q2);
                const unsigned off = (__float_as_uint(__fadd_rz(sd, 2048.f)) & (63u << 7));       // floor(32 s) * 128, 32 when s = 1
                atomicAdd(reinterpret_cast<unsigned *>(hcol + off), 1u);
            };
#pragma unroll 1
            for (int r = 0; r < 6; ++r) {                        // (rolled: the unrolled kernel stalled on instruction fetch, ncu no_inst 27 %)
                int i = (int)S.seg_off[run0 + 2 * kSelW * r];
                const int end = (int)S.seg_off[run0 + 2 * kSelW * r + 12];
                for (; i + 4 <= end; i += 4) {                   // four loads in flight before the first atomic (which orders them);
                    const float4 c0 = S.cand[i], c1 = S.cand[i + 1], c2 = S.cand[i + 2], c3 = S.cand[i + 3];      // eight: 79 registers, 2 % slower
                    deposit(c0); deposit(c1); deposit(c2); deposit(c3);
                }
                for (; i < end; ++i) deposit(S.cand[i]);
            }
        }
        // bin in which the running count reaches K
        int bstar = kSelNB, cnt = 0;
        {
            int run = 0;
            bool found = false;
#pragma unroll
            for (int i = 0; i < kSelNB; ++i) {
                const int hcount = (int)*reinterpret_cast<unsigned *>(hcol + i * 128);
                if (!found && run + hcount >= K) { found = true; bstar = i; cnt = run; }
                run += hcount;
            }
        }
        bool ok = active && bstar < kSelNB;
        // bin b holds s in [b / 32, (b + 1) / 32); cnt = candidates whose float s lies below the bin
        const float elo = (float)bstar * (1.f / kSelNB);
        const float midq = ((float)bstar + 0.5f) * (1.f / kSelNB) - q2, hw = 0.5f / kSelNB + kSelDelta;      // band |s - mid| <= hw, tested before q2 is added
        const double TL = (double)bstar * (1.0 / kSelNB) * sp.dref2, TH = (double)(bstar + 1) * (1.0 / kSelNB) * sp.dref2;
        int np = 0, ne = 0;
        __syncwarp();                                            // every histogram column has been read: the exact lists may overwrite them

        // ---- pass B: remember the candidates of the band -----------------------------------------------------------------
        {
            auto in_band = [&](const float4 c) {
                return fabsf(fmaf(m2x, c.x, fmaf(m2y, c.y, fmaf(m2z, c.z, c.w))) - midq) <= hw;
            };
            auto remember = [&](int i) { pcol[min(np, kSelList - 1) * 32] = (unsigned short)i; ++np; };
#pragma unroll 1
            for (int r = 0; r < 6; ++r) {
                int i = (int)S.seg_off[run0 + 2 * kSelW * r];
                const int end = (int)S.seg_off[run0 + 2 * kSelW * r + 12];
                for (; i + 16 <= end; i += 16) {                 // 16 candidates -> one hit mask; the few hits are stored afterwards
                    unsigned mask = 0u;
#pragma unroll
                    for (int k = 0; k < 16; k += 4) {
                        const float4 c0 = S.cand[i + k], c1 = S.cand[i + k + 1], c2 = S.cand[i + k + 2], c3 = S.cand[i + k + 3];
                        if (in_band(c0)) mask |= 1u << k;
                        if (in_band(c1)) mask |= 2u << k;
                        if (in_band(c2)) mask |= 4u << k;
                        if (in_band(c3)) mask |= 8u << k;
                    }
                    while (mask) {
                        remember(i + __ffs((int)mask) - 1);
                        mask &= mask - 1u;
                    }
                }
                for (; i + 4 <= end; i += 4) {                   // tail of the run: groups of four, then single candidates
                    const float4 c0 = S.cand[i], c1 = S.cand[i + 1], c2 = S.cand[i + 2], c3 = S.cand[i + 3];
                    unsigned mask = (in_band(c0) ? 1u : 0u) | (in_band(c1) ? 2u : 0u) | (in_band(c2) ? 4u : 0u) | (in_band(c3) ? 8u : 0u);
                    while (mask) {
                        remember(i + __ffs((int)mask) - 1);
                        mask &= mask - 1u;
                    }
                }
                for (; i < end; ++i)
                    if (in_band(S.cand[i])) remember(i);
            }
        }
        // ---- exact float64 evaluation of the remembered candidates, all lanes together ----------------------------------------
        bool ovf = np > kSelList;
        np = min(np, kSelList);
        {
            const int mx = __reduce_max_sync(0xffffffffu, ok ? np : 0);
            for (int i = 0; i < mx; ++i) {
                if (ok && i < np) {
                    const int ci = pcol[i * 32];
                    const float4 c = S.cand[ci];
                    const float sd = __saturatef(fmaf(m2x, c.x, fmaf(m2y, c.y, fmaf(m2z, c.z, c.w))) + q2);      // as in pass A
                    cnt -= sd < elo ? 1 : 0;                     // pass A counted it below the bin; the float64 test decides
                    const uint32_t j = S.candj[ci];
                    double ex = xs[j] - x, ey = ys[j] - y, ez = zs[j] - z;
                    if (PER) {                                   // scipy's per-pair wrap, branch-free (x + 0.0 is x)
                        ex = AST_DADD(ex, ex < -half_box ? box : (ex > half_box ? -box : 0.0));
                        ey = AST_DADD(ey, ey < -half_box ? box : (ey > half_box ? -box : 0.0));
                        ez = AST_DADD(ez, ez < -half_box ? box : (ez > half_box ? -box : 0.0));
                    }
                    const double d2 = AST_DADD(AST_DADD(AST_DMUL(ex, ex), AST_DMUL(ey, ey)), AST_DMUL(ez, ez));
                    if (d2 < TL) ++cnt;
                    else if (d2 <= TH) { ecol[ne * 32] = d2; ++ne; }      // ne <= np <= kSelList
                }
            }
        }
        // ---- the (K - cnt)-th smallest of the exact list ------------------------------------------------------------------
        const int want = K - cnt;                                // 1-based rank inside the list
        ok = ok && !ovf && want >= 1 && want <= ne;
        double ans = INFINITY;
        {
            const int mx = __reduce_max_sync(0xffffffffu, ok ? ne : 0);
            for (int i = 0; i < mx; ++i) {
                const double ei = i < ne ? ecol[i * 32] : INFINITY;
                int rank = 0;
                for (int j = 0; j < mx; ++j) {
                    const double ej = j < ne ? ecol[j * 32] : INFINITY;
                    rank += (ej < ei || (ej == ei && j < i)) ? 1 : 0;
                }
                if (ok && i < ne && rank == want - 1) ans = ei;
            }
        }
        ok = ok && ans < safe2;
        if (in_block) sp.fail[s] = (active && !ok) ? 1u : 0u;
        if (ok) a.h_out[row] = sqrt(ans);
    }
}

}  // namespace ast
