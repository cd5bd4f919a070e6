// slabroute.cu -- packing kernel of the multi-GPU k-NN exchange (slabs along x + ghost zones).
//
// Every rank holds an arbitrary set of particles (the reference's per-rank read, io/EAGLE/_SnapshotEAGLE.py:120-130) and has
// to send each of them to the rank that OWNS its slab and, as a ghost, to every rank whose slab lies within the ghost width
// w of it.  The torch version of this step walks the particle array once per destination rank (remainder, where, two
// nonzero, a cat and a gather: ~10 passes of n_local per rank of the job); here it is two passes in all:
//   count  one ballot per (kind, destination) and warp -> per-block counts, table [kind][destination][block]
//   scan   one exclusive scan over the flattened table = first send row of every (kind, destination, block)
//   write  the same ballots again -> rank inside the warp / block -> send row; the particle's position goes straight into
//          the send buffer and its index into src_index (the way h finds back home)
// Send buffer layout: [owned -> rank 0 | owned -> rank 1 | ... | ghosts -> rank 0 | ghosts -> rank 1 | ...], inside each piece
// ascending particle index (blocks, warps and lanes are walked in order), which is also what the torch version produces.
// Two all-to-all calls (owned, then ghosts) then leave the receiver with [owned from all | ghosts from all] without a reorder.
#include "common.cuh"
#include "scan_sort.cuh"

namespace ast {

constexpr int kRouteMaxWorld = 32;

struct RouteArgs {
    const double *pos;            // (n, 3)
    const int64_t *owner;         // (n) owning rank of every particle (from the global histogram plan)
    int64_t n;
    int world, periodic, covers_all;
    double length, w;
    double b_lo[kRouteMaxWorld], b_hi[kRouteMaxWorld];
};

// bit g of the result: the particle goes to rank g as a ghost
__device__ __forceinline__ uint32_t route_ghosts(const RouteArgs &a, double x, int own)
{
    uint32_t m = 0;
    for (int g = 0; g < a.world; ++g) {
        if (g == own) continue;
        double d;
        if (a.periodic) {
            double t = fmod(x - a.b_lo[g], a.length);
            if (t < 0.0) t += a.length;
            const double seg = a.b_hi[g] - a.b_lo[g];
            d = t <= seg ? 0.0 : fmin(t - seg, a.length - t);
        } else {
            d = fmax(fmax(a.b_lo[g] - x, x - a.b_hi[g]), 0.0);
        }
        if (d <= a.w || a.covers_all) m |= 1u << g;
    }
    return m;
}

// table[(kind * world + g) * nblocks + block], kind 0 = owned, 1 = ghost
template <bool WRITE>
__global__ void __launch_bounds__(256) slab_route_kernel(RouteArgs a, uint64_t *__restrict__ table, int64_t nblocks,
                                                         double *__restrict__ send, int64_t *__restrict__ src_index)
{
    __shared__ uint32_t wcount[2 * kRouteMaxWorld][8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t i = (int64_t)blockIdx.x * 256 + tid;
    const bool valid = i < a.n;
    int own = -1;
    uint32_t gm = 0;
    double x = 0.0, y = 0.0, z = 0.0;
    if (valid) {
        x = a.pos[3 * i];
        own = (int)a.owner[i];
        if (own < 0 || own >= a.world) own = -1;                       // (never from _slab_plan; such a particle would be dropped)
        gm = route_ghosts(a, x, own);
        if (WRITE) { y = a.pos[3 * i + 1]; z = a.pos[3 * i + 2]; }
    }
    const unsigned lt = (1u << lane) - 1u;
    for (int g = 0; g < a.world; ++g) {
        const unsigned bo = __ballot_sync(0xffffffffu, own == g), bg = __ballot_sync(0xffffffffu, (gm >> g) & 1u);
        if (lane == 0) { wcount[g][warp] = __popc(bo); wcount[a.world + g][warp] = __popc(bg); }
    }
    __syncthreads();
    if (!WRITE) {
        if (tid < 2 * a.world) {
            uint32_t s = 0;
            for (int w = 0; w < 8; ++w) s += wcount[tid][w];
            table[(int64_t)tid * nblocks + blockIdx.x] = s;
        }
    }
    // (all lanes stay to the end: the ballots below are full-warp)
    for (int g = 0; WRITE && g < a.world; ++g) {
        const bool is_own = own == g, is_ghost = (gm >> g) & 1u;
        const unsigned bo = __ballot_sync(0xffffffffu, is_own), bg = __ballot_sync(0xffffffffu, is_ghost);
        if (!is_own && !is_ghost) continue;
        const int piece = is_own ? g : a.world + g;
        uint32_t before = 0;
        for (int w = 0; w < warp; ++w) before += wcount[piece][w];
        const int64_t row = (int64_t)table[(int64_t)piece * nblocks + blockIdx.x] + before + __popc((is_own ? bo : bg) & lt);
        send[3 * row] = x; send[3 * row + 1] = y; send[3 * row + 2] = z;
        src_index[row] = i;
    }
}

// counts[kind * world + g] = table entry of the first block of the NEXT piece - first block of this piece
__global__ void slab_route_counts_kernel(const uint64_t *__restrict__ table, const uint64_t *__restrict__ total, int64_t nblocks, int pieces,
                                         int64_t *__restrict__ counts)
{
    const int p = threadIdx.x;
    if (p >= pieces) return;
    const uint64_t b = table[(int64_t)p * nblocks], e = p + 1 < pieces ? table[(int64_t)(p + 1) * nblocks] : *total;
    counts[p] = (int64_t)(e - b);
}

}  // namespace ast

using namespace ast;

static int route_validate(const ast_slab_route_params *p)
{
    AST_REQUIRE(p != nullptr, "params is null");
    AST_REQUIRE(p->n >= 0, "n < 0");
    AST_REQUIRE(p->world >= 1 && p->world <= kRouteMaxWorld, "world = %d not in [1, %d]", p->world, kRouteMaxWorld);
    AST_REQUIRE(p->w >= 0.0, "ghost width < 0");
    AST_REQUIRE(!p->periodic || p->length > 0.0, "periodic routing needs length > 0");
    return AST_OK;
}

static int64_t route_blocks(int64_t n) { return n > 0 ? (n + 255) / 256 : 1; }

extern "C" int ast_slab_route_workspace_bytes(const ast_slab_route_params *p, size_t *bytes)
{
    int rc = route_validate(p);
    if (rc) return rc;
    AST_REQUIRE(bytes != nullptr, "bytes is null");
    const int64_t entries = 2 * (int64_t)p->world * route_blocks(p->n);
    Carver c(nullptr);
    c.take<uint64_t>(entries + 1);
    c.take<char>(scan_workspace_bytes<uint64_t>(entries + 1));
    *bytes = c.bytes();
    return AST_OK;
}

static RouteArgs route_args(const ast_slab_route_params *p, const double *pos, const int64_t *owner)
{
    RouteArgs a;
    memset(&a, 0, sizeof a);
    a.pos = pos; a.owner = owner; a.n = p->n; a.world = p->world; a.periodic = p->periodic; a.covers_all = p->covers_all;
    a.length = p->length; a.w = p->w;
    for (int g = 0; g < p->world; ++g) { a.b_lo[g] = p->bounds[g]; a.b_hi[g] = p->bounds[g + 1]; }
    return a;
}

extern "C" int ast_slab_route_count(const ast_slab_route_params *p, const double *pos, const int64_t *owner, int64_t *counts,
                                    void *workspace, size_t workspace_bytes, void *stream)
{
    int rc = route_validate(p);
    if (rc) return rc;
    AST_REQUIRE(counts != nullptr && (p->n == 0 || (pos && owner)), "null pointer");
    size_t need = 0;
    ast_slab_route_workspace_bytes(p, &need);
    if (!workspace || workspace_bytes < need) { set_error("workspace too small: need %zu bytes, have %zu", need, workspace_bytes); return AST_EWORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t nb = route_blocks(p->n), entries = 2 * (int64_t)p->world * nb;
    Carver c(workspace);
    uint64_t *table = c.take<uint64_t>(entries + 1);
    uint64_t *tmp = (uint64_t *)c.take<char>(scan_workspace_bytes<uint64_t>(entries + 1));
    RouteArgs a = route_args(p, pos, owner);
    AST_CUDA_TRY(cudaMemsetAsync(table, 0, sizeof(uint64_t) * (entries + 1), s));
    if (p->n > 0) slab_route_kernel<false><<<(unsigned)nb, 256, 0, s>>>(a, table, nb, nullptr, nullptr);
    AST_KERNEL_CHECK(s, "slab_route_kernel<count>");
    AST_CUDA_TRY(scan_exclusive<uint64_t>(table, entries, tmp, table + entries, s));      // total behind the table
    slab_route_counts_kernel<<<1, 64, 0, s>>>(table, table + entries, nb, 2 * p->world, counts);
    AST_KERNEL_CHECK(s, "slab_route_counts_kernel");
    return AST_OK;
}

extern "C" int ast_slab_route_write(const ast_slab_route_params *p, const double *pos, const int64_t *owner, double *send,
                                    int64_t *src_index, void *workspace, size_t workspace_bytes, void *stream)
{
    int rc = route_validate(p);
    if (rc) return rc;
    if (p->n == 0) return AST_OK;
    AST_REQUIRE(pos && owner && send && src_index, "null pointer");
    size_t need = 0;
    ast_slab_route_workspace_bytes(p, &need);
    if (!workspace || workspace_bytes < need) { set_error("workspace too small: need %zu bytes, have %zu", need, workspace_bytes); return AST_EWORKSPACE; }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t nb = route_blocks(p->n);
    Carver c(workspace);
    uint64_t *table = c.take<uint64_t>(2 * (int64_t)p->world * nb + 1);      // scanned by ast_slab_route_count on the same workspace
    RouteArgs a = route_args(p, pos, owner);
    slab_route_kernel<true><<<(unsigned)nb, 256, 0, s>>>(a, table, nb, send, src_index);
    AST_KERNEL_CHECK(s, "slab_route_kernel<write>");
    return AST_OK;
}
