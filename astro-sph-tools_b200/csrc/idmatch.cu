// idmatch.cu -- particle-ID matching on the GPU: for every target ID the index of the equal source ID.
//
// Replaces the sort / intersect / searchsorted arithmetic of the reference's ArrayReorder family
// (tools/_ArrayReorder.py:744-768 ArrayReorder_2.create -> np.intersect1d(..., assume_unique=True, return_indices=True);
// :988-1038 ArrayReorder.create -> argsort + np.isin; used by io/EAGLE/_CatalogueSUBFIND.py:292-295).  Integer work: the
// result is bit-exact by construction.  Open-addressing hash table (capacity = power of two >= 2 n_src) built with 64-bit
// atomicCAS, then one probe sequence per target; duplicate source IDs resolve to the smallest source index.
#include "common.cuh"

namespace ast {

constexpr long long kEmptyKey = (long long)0x8000000000000000ull;      // INT64_MIN is not a valid particle ID

__device__ __forceinline__ uint64_t mix64(uint64_t x)                   // splitmix64 finaliser
{
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27; x *= 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

__global__ void idmatch_init_kernel(long long *keys, long long *vals, int64_t cap)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) { keys[i] = kEmptyKey; vals[i] = 0x7fffffffffffffffll; }
}

__global__ void idmatch_insert_kernel(const long long *__restrict__ ids, const uint8_t *__restrict__ filter, int64_t n,
                                      long long *keys, long long *vals, uint64_t mask)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (filter && !filter[i])) return;
    const long long id = ids[i];
    uint64_t slot = mix64((uint64_t)id) & mask;
    for (;;) {
        const long long old = (long long)atomicCAS((unsigned long long *)&keys[slot], (unsigned long long)kEmptyKey, (unsigned long long)id);
        if (old == kEmptyKey || old == id) { atomicMin(&vals[slot], (long long)i); return; }
        slot = (slot + 1) & mask;
    }
}

__global__ void idmatch_lookup_kernel(const long long *__restrict__ ids, const uint8_t *__restrict__ filter, int64_t n,
                                      const long long *__restrict__ keys, const long long *__restrict__ vals, uint64_t mask,
                                      long long *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long r = -1;
    if (!filter || filter[i]) {
        const long long id = ids[i];
        uint64_t slot = mix64((uint64_t)id) & mask;
        for (;;) {
            const long long k = keys[slot];
            if (k == id) { r = vals[slot]; break; }
            if (k == kEmptyKey) break;
            slot = (slot + 1) & mask;
        }
    }
    out[i] = r;
}

// out[j] = src[index[j]] (rows of row_bytes bytes) where index[j] >= 0; other rows are left untouched
__global__ void gather_rows_kernel(const uint8_t *__restrict__ src, int64_t row_bytes, const long long *__restrict__ index, int64_t n_out,
                                   uint8_t *__restrict__ out)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t j = t / row_bytes, b = t - j * row_bytes;
    if (j >= n_out) return;
    const long long s = index[j];
    if (s >= 0) out[t] = src[s * row_bytes + b];
}

static int64_t idmatch_capacity(int64_t n)
{
    int64_t cap = 1024;
    while (cap < 2 * n) cap <<= 1;
    return cap;
}

}  // namespace ast

using namespace ast;

extern "C" int ast_match_ids_workspace_bytes(int64_t n_source, size_t *bytes)
{
    AST_REQUIRE(bytes != nullptr && n_source >= 0, "bad argument");
    *bytes = (size_t)idmatch_capacity(n_source) * 16 + 512;
    return AST_OK;
}

extern "C" int ast_match_ids(const int64_t *source_ids, int64_t n_source, const uint8_t *source_filter, const int64_t *target_ids,
                             int64_t n_target, const uint8_t *target_filter, int64_t *source_index_of_target, void *workspace,
                             size_t workspace_bytes, void *stream)
{
    AST_REQUIRE(n_source >= 0 && n_target >= 0, "negative length");
    AST_REQUIRE(n_target == 0 || (target_ids && source_index_of_target), "null pointer");
    AST_REQUIRE(n_source == 0 || source_ids, "null pointer");
    const int64_t cap = idmatch_capacity(n_source);
    if (!workspace || workspace_bytes < (size_t)cap * 16) {
        set_error("workspace too small: need %zu bytes", (size_t)cap * 16 + 512);
        return AST_EWORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    long long *keys = (long long *)workspace, *vals = keys + cap;
    idmatch_init_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, s>>>(keys, vals, cap);
    if (n_source > 0)
        idmatch_insert_kernel<<<(unsigned)((n_source + 255) / 256), 256, 0, s>>>((const long long *)source_ids, source_filter, n_source, keys,
                                                                              vals, (uint64_t)(cap - 1));
    if (n_target > 0)
        idmatch_lookup_kernel<<<(unsigned)((n_target + 255) / 256), 256, 0, s>>>((const long long *)target_ids, target_filter, n_target, keys,
                                                                              vals, (uint64_t)(cap - 1), (long long *)source_index_of_target);
    AST_CUDA_TRY(cudaGetLastError());
    return AST_OK;
}

extern "C" int ast_gather_rows(const void *src, int64_t row_bytes, const int64_t *index, int64_t n_out, void *out, void *stream)
{
    AST_REQUIRE(row_bytes > 0 && n_out >= 0, "bad argument");
    if (n_out == 0) return AST_OK;
    AST_REQUIRE(index && out, "null pointer");                 // src may be null when the source is empty (every index is -1)
    const int64_t total = n_out * row_bytes;
    gather_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const uint8_t *)src, row_bytes,
                                                                                          (const long long *)index, n_out, (uint8_t *)out);
    AST_CUDA_TRY(cudaGetLastError());
    return AST_OK;
}
