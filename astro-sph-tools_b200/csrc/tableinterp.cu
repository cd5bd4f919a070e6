// tableinterp.cu -- N-D linear table interpolation on the GPU and the fused per-particle weight.
//
// Replaces IonisationTableBase.__call__ / evaluate_at_redshift of the reference (data_structures/_IonisationTable.py:44-58:
// scipy RegularGridInterpolator(method="linear", bounds_error=False, fill_value=-inf) over (log10 nH, log10 T, z); table
// loaded at io/ionisation_tables/_HM01.py:73-97).  The arithmetic is scipy's, restated operation by operation so the result
// is bit-equal in float64 (no FMA contraction: every product/sum goes through __dmul_rn/__dadd_rn):
//   interval  i_d = clamp(#{g_d <= x_d} - 1, 0, n_d - 2)                (scipy _rgi_cython.find_indices)
//   y_d = (x_d - g_d[i_d]) / (g_d[i_d+1] - g_d[i_d])
//   value = 0 + sum over corners in itertools.product order (dimension 0 slowest, lower corner first) of
//           table[corner] * (((1 * w_0) * w_1) ...),  w_d = 1 - y_d (lower) or y_d (upper)      (_rgi.py _evaluate_linear)
//   (2-D tables: scipy's compiled fast path evaluate_linear_2d, ((v * w_0) * w_1) summed left to right)
//   any x_d outside [g_d[0], g_d[-1]] -> fill_value;  any NaN coordinate -> NaN (applied last)   (_rgi.py __call__)
// The fused form out = base * 10^value is the per-particle weight of an ion column-density map (element mass x ion
// fraction): it never leaves the device between the table and the deposition kernels.
#include <string.h>

#include "common.cuh"

namespace ast {

struct TabArgs {
    int ndim, fixed_dim, pow10;
    int shape[AST_TABLE_MAX_DIM];
    int axis_off[AST_TABLE_MAX_DIM];          // offset of each axis in the shared-memory copy
    const double *axes[AST_TABLE_MAX_DIM];
    const double *cols[AST_TABLE_MAX_DIM];
    int64_t strides[AST_TABLE_MAX_DIM];
    const double *table, *base;
    double fill, fixed_value;
    double *out;
    int64_t n;
    int smem_axes;                            // 1: axes are staged in shared memory
};

template <int NDIM>
__global__ void __launch_bounds__(256) table_interp_kernel(TabArgs a)
{
    extern __shared__ double s_axes[];
    if (a.smem_axes) {
        for (int d = 0; d < NDIM; ++d)
            for (int j = threadIdx.x; j < a.shape[d]; j += blockDim.x) s_axes[a.axis_off[d] + j] = a.axes[d][j];
        __syncthreads();
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    int idx[NDIM];
    double y[NDIM];
    bool nan = false, oob = false;
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
        const double x = d == a.fixed_dim ? a.fixed_value : a.cols[d][i * a.strides[d]];
        const double *g = a.smem_axes ? s_axes + a.axis_off[d] : a.axes[d];
        const int n = a.shape[d];
        nan = nan || (x != x);
        oob = oob || (x < g[0]) || (x > g[n - 1]);
        int lo = 0, hi = n;                   // first index with g[index] > x  (NaN: every comparison false -> n)
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (g[mid] > x) hi = mid; else lo = mid + 1;
        }
        int k = lo - 1;
        k = k < 0 ? 0 : (k > n - 2 ? n - 2 : k);
        idx[d] = k;
        y[d] = __ddiv_rn(__dsub_rn(x, g[k]), __dsub_rn(g[k + 1], g[k]));
    }
    double value = 0.0;
    if (NDIM == 2) {
        // scipy's compiled fast path for 2-D float64 tables (_rgi_cython.evaluate_linear_2d) associates differently:
        // ((v * w0) * w1) per corner, summed left to right without the leading zero
        const double *t = a.table + (int64_t)idx[0] * a.shape[1] + idx[1];
        const double u0 = __dsub_rn(1.0, y[0]), u1 = __dsub_rn(1.0, y[1]);
        value = __dmul_rn(__dmul_rn(__ldg(t), u0), u1);
        value = __dadd_rn(value, __dmul_rn(__dmul_rn(__ldg(t + 1), u0), y[1]));
        value = __dadd_rn(value, __dmul_rn(__dmul_rn(__ldg(t + a.shape[1]), y[0]), u1));
        value = __dadd_rn(value, __dmul_rn(__dmul_rn(__ldg(t + a.shape[1] + 1), y[0]), y[1]));
    } else
#pragma unroll
    for (int c = 0; c < (1 << NDIM); ++c) {
        double w = 1.0;
        int64_t off = 0;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) {
            const int up = (c >> (NDIM - 1 - d)) & 1;               // dimension 0 varies slowest
            w = __dmul_rn(w, up ? y[d] : __dsub_rn(1.0, y[d]));
            off = off * a.shape[d] + idx[d] + up;
        }
        value = __dadd_rn(value, __dmul_rn(__ldg(a.table + off), w));
    }
    if (oob) value = a.fill;
    if (nan) value = __longlong_as_double(0x7ff8000000000000ll);
    if (a.pow10) value = exp10(value);
    if (a.base) value = __dmul_rn(a.base[i], value);
    a.out[i] = value;
}

}  // namespace ast

using namespace ast;

extern "C" int ast_table_interp(const ast_table_params *t, const double *const *x_cols, const int64_t *x_strides, int64_t n,
                                const double *base, double *out, void *stream)
{
    AST_REQUIRE(t != nullptr, "table params are null");
    AST_REQUIRE(t->ndim >= 1 && t->ndim <= AST_TABLE_MAX_DIM, "ndim must be 1..%d, got %d", AST_TABLE_MAX_DIM, t->ndim);
    AST_REQUIRE(t->table != nullptr, "table is null");
    AST_REQUIRE(t->fixed_dim >= -1 && t->fixed_dim < t->ndim, "fixed_dim %d out of range", t->fixed_dim);
    AST_REQUIRE(n >= 0, "n < 0");
    AST_REQUIRE(n == 0 || (x_cols && x_strides && out), "null coordinate / output pointer");
    TabArgs a;
    memset(&a, 0, sizeof a);
    int total = 0;
    for (int d = 0; d < t->ndim; ++d) {
        AST_REQUIRE(t->shape[d] >= 2, "dimension %d has %d grid points, linear interpolation needs >= 2", d, t->shape[d]);
        AST_REQUIRE(t->axes[d] != nullptr, "axes[%d] is null", d);
        AST_REQUIRE(n == 0 || d == t->fixed_dim || x_cols[d] != nullptr, "x_cols[%d] is null", d);
        a.shape[d] = t->shape[d];
        a.axes[d] = t->axes[d];
        a.axis_off[d] = total;
        total += t->shape[d];
        a.cols[d] = (n > 0 && d != t->fixed_dim) ? x_cols[d] : nullptr;
        a.strides[d] = (n > 0 && d != t->fixed_dim) ? x_strides[d] : 0;
    }
    if (n == 0) return AST_OK;
    a.ndim = t->ndim; a.fixed_dim = t->fixed_dim; a.pow10 = (t->flags & AST_TABLE_POW10) ? 1 : 0;
    a.table = t->table; a.base = base; a.fill = t->fill_value; a.fixed_value = t->fixed_value; a.out = out; a.n = n;
    a.smem_axes = total <= 4096 ? 1 : 0;
    const size_t smem = a.smem_axes ? sizeof(double) * (size_t)total : 0;
    cudaStream_t s = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n + 255) / 256);
    switch (t->ndim) {
    case 1: table_interp_kernel<1><<<grid, 256, smem, s>>>(a); break;
    case 2: table_interp_kernel<2><<<grid, 256, smem, s>>>(a); break;
    case 3: table_interp_kernel<3><<<grid, 256, smem, s>>>(a); break;
    default: table_interp_kernel<4><<<grid, 256, smem, s>>>(a); break;
    }
    AST_KERNEL_CHECK(s, "table_interp_kernel");
    return AST_OK;
}
