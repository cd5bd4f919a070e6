"""CPU check of the float32 error bound the k-NN selection kernel relies on (csrc/knn_select.cuh, header comment): the kernel
classifies a candidate by its float32 squared distance s only when s lies further than delta = 4e-3 / 32 from an edge of the
band; everything closer is re-evaluated in float64.  That is exact as long as |s_float32 - s_true| stays below delta / 2.  Here
the kernel's float32 expression is restated in numpy over the worst geometry it accepts (region 8 cells wide, cell anisotropy
1.5, reach 2.5 cells) and compared with float64."""
import numpy as np


def staged(u, W, sc):
    """sel_stage: v = float32((u - W/2) * sc) per axis, w = |v|^2 accumulated in float32"""
    v = ((u - 0.5 * W) * sc).astype(np.float32)
    w = (v[:, 0] * v[:, 0]).astype(np.float32)
    w = (v[:, 1] * v[:, 1] + w).astype(np.float32)
    w = (v[:, 2] * v[:, 2] + w).astype(np.float32)
    return v, w


def test_float32_squared_distance_error_stays_far_below_delta():
    rng = np.random.default_rng(2024)
    delta = 4e-3 / 32
    worst = 0.0
    for aniso in (1.0, 1.25, 1.5):
        cs = np.array([aniso, 1.0, 1.0])                       # cell sizes; Dref = 2.5 * min(cs)
        sc = cs / 2.5                                          # cell units -> reach units
        W = np.array([8.0, 8.0, 8.0])
        n = 400000
        uc = rng.uniform(0.0, 8.0, (n, 3))                     # candidates anywhere in the staged region
        uq = rng.uniform(2.0, 6.0, (n, 3))                     # queries inside the block
        vc, wc = staged(uc, W, sc)
        vq, wq = staged(uq, W, sc)
        m2 = (np.float32(-2.0) * vq).astype(np.float32)
        # s = fma(m2x, cx, fma(m2y, cy, fma(m2z, cz, w_c))) + w_q ; numpy rounds after every multiply as well (a looser model)
        t = (m2[:, 2] * vc[:, 2] + wc).astype(np.float32)
        t = (m2[:, 1] * vc[:, 1] + t).astype(np.float32)
        t = (m2[:, 0] * vc[:, 0] + t).astype(np.float32)
        s32 = np.clip((t + wq).astype(np.float32), 0.0, 1.0).astype(np.float64)
        true = np.clip((((uc - uq) * sc) ** 2).sum(axis=1), 0.0, 1.0)
        near = true < 1.0                                      # only candidates inside the reach are ever binned
        worst = max(worst, float(np.abs(s32[near] - true[near]).max()))
    assert worst < delta / 4, worst                            # measured ~4e-6; the header derives < 7e-6
