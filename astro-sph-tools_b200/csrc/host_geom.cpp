// host_geom.cpp -- compiles the product's float64 index math (ast_geom.h) for the HOST so that the CPU test
// suite can compare it bit-for-bit with the oracle without a GPU.  Not part of libastsph_b200.so.
#include <vector>

#include "../../include/astro_sph_b200.h"
#include "ast_geom.h"

using namespace ast;

extern "C" void hostgeom_bbox_cls2d(const double *pos, const double *h, int64_t n, int axis, int nx, int ny, double x_min,
                                    double x_max, double y_min, double y_max, int n_img, const double *shift_a,
                                    const double *shift_b, int64_t small_max_px, int64_t huge_min_tiles, int32_t *bbox,
                                    uint8_t *cls)
{
    int ac, bc;
    plane_columns(axis, ac, bc);
    Axis1 ax = make_axis(x_min, x_max, nx), ay = make_axis(y_min, y_max, ny);
    for (int m = 0; m < n_img; ++m)
        for (int64_t i = 0; i < n; ++i) {
            double pa = pos[3 * i + ac] + shift_a[m], pb = pos[3 * i + bc] + shift_b[m];
            Bin2 b = classify2<AST_TILE>(ax, ay, pa, pb, h[i], radius2(h[i]), small_max_px, huge_min_tiles);
            int64_t j = (int64_t)m * n + i;
            bbox[4 * j] = b.bb.x0; bbox[4 * j + 1] = b.bb.x1; bbox[4 * j + 2] = b.bb.y0; bbox[4 * j + 3] = b.bb.y1;
            cls[j] = (uint8_t)b.cls;
        }
}

// pairs in emit order (particle, image, tx, ty); returns the count, writes at most cap elements
extern "C" int64_t hostgeom_pairs2d(const double *pos, const double *h, int64_t n, int axis, int nx, int ny, double x_min,
                                    double x_max, double y_min, double y_max, int n_img, const double *shift_a,
                                    const double *shift_b, int64_t small_max_px, int64_t huge_min_tiles, uint64_t *pairs,
                                    int64_t cap)
{
    int ac, bc;
    plane_columns(axis, ac, bc);
    Axis1 ax = make_axis(x_min, x_max, nx), ay = make_axis(y_min, y_max, ny);
    const int nty = (ny + AST_TILE - 1) / AST_TILE;
    const int img_shift = n_img == 1 ? 0 : 4;
    int64_t g = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int m = 0; m < n_img; ++m) {
            double pa = pos[3 * i + ac] + shift_a[m], pb = pos[3 * i + bc] + shift_b[m];
            double R2 = radius2(h[i]);
            Bin2 b = classify2<AST_TILE>(ax, ay, pa, pb, h[i], R2, small_max_px, huge_min_tiles);
            if (b.cls != CLS_TILED) continue;
            for_each_tile2<AST_TILE>(ax, ay, pa, pb, R2, b, nty, [&](uint32_t key) {
                if (g < cap) pairs[g] = ((uint64_t)((key << img_shift) | (uint32_t)m) << 32) | (uint64_t)(uint32_t)i;
                ++g;
            });
        }
    // large-h split: the member tiles of the large-h entries (list order) follow the tiled pairs; the device gives every
    // entry a warp (huge_tiles_kernel), the membership test and the order are those of for_each_tile2
    for (int64_t i = 0; i < n; ++i)
        for (int m = 0; m < n_img; ++m) {
            double pa = pos[3 * i + ac] + shift_a[m], pb = pos[3 * i + bc] + shift_b[m];
            double R2 = radius2(h[i]);
            Bin2 b = classify2<AST_TILE>(ax, ay, pa, pb, h[i], R2, small_max_px, huge_min_tiles);
            if (b.cls != CLS_HUGE) continue;
            for_each_tile2<AST_TILE>(ax, ay, pa, pb, R2, b, nty, [&](uint32_t key) {
                if (g < cap) pairs[g] = ((uint64_t)((key << img_shift) | (uint32_t)m) << 32) | (uint64_t)(uint32_t)i;
                ++g;
            });
        }
    return g;
}
