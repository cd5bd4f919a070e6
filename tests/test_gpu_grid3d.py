"""-m gpu: 3-D voxel gridding (ast_grid3d / ast_bin3d through the C ABI) against the CPU oracle.  The reference has
no 3-D function (EXTENSION, BASELINE.json config 4); the oracle restates the 2-D rules in 3-D.  Index work bit-exact,
grids within 1e-5 relative L2 and 1e-6 in the total."""
import ctypes as C

import numpy as np
import pytest

from conftest import rel_l2, random_cloud

pytestmark = pytest.mark.gpu


def gpu_grid(pos, h, prop, size, lo, hi, kernel="cubic_spline_3d", periodic=False, box=None, **kw):
    import torch
    from astro_sph_tools_b200.tools.projections import Gridder3D
    g = Gridder3D(**kw)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    out = g.grid(d(pos), d(h), d(prop), size, lo, hi, kernel, periodic, box)
    torch.cuda.synchronize()
    return out.cpu().numpy(), g.last_stats


def gpu_bin3d(pos, h, size, lo, hi, periodic=False, box=None, small=64, huge=512, cap=1 << 22):
    import torch
    from astro_sph_tools_b200 import _lib
    from astro_sph_tools_b200.tools.projections import Gridder3D
    g = Gridder3D(pair_capacity=cap, huge_capacity=1 << 18, small_max_vox=small, huge_min_bricks=huge)
    n = len(h); n_img = 27 if periodic else 1
    p = g.params(n, size, lo, hi, "cubic_spline_3d", periodic, box)
    ws = g.workspace(p)
    pos_d = torch.from_numpy(np.ascontiguousarray(pos)).cuda(); h_d = torch.from_numpy(np.ascontiguousarray(h)).cuda()
    bbox = torch.empty((n_img * n, 6), dtype=torch.int32, device="cuda")
    cls = torch.empty(n_img * n, dtype=torch.uint8, device="cuda")
    ps = torch.zeros(cap, dtype=torch.int64, device="cuda"); hg = torch.zeros(1 << 18, dtype=torch.int64, device="cuda")
    counts = (C.c_int64 * 2)()
    _lib.check(g.lib.ast_bin3d(C.byref(p), _lib.ptr(pos_d), _lib.ptr(h_d), _lib.ptr(bbox), _lib.ptr(cls), _lib.ptr(ps), _lib.ptr(hg),
                               counts, _lib.ptr(ws), C.c_size_t(ws.numel()), _lib.stream_ptr()))
    torch.cuda.synchronize()
    return dict(bbox=bbox.cpu().numpy(), cls=cls.cpu().numpy().reshape(n_img, n), sorted=ps.cpu().numpy()[:counts[0]].view(np.uint64),
                huge=hg.cpu().numpy()[:counts[1]].view(np.uint64))


def check(g, ref):
    assert g.shape == ref.shape
    assert rel_l2(g, ref) <= 1e-5
    assert abs(g.sum() - ref.sum()) <= 1e-6 * np.abs(ref).sum()


@pytest.mark.parametrize("periodic", [False, True])
def test_bin3d_bit_exact(oracle, periodic):
    rng = np.random.default_rng(31)
    pos = rng.uniform(-0.05, 1.05, (3000, 3))
    d = 1.0 / 40
    on = rng.random(3000) < 0.3
    pos[on] = np.round(pos[on] / d) * d                                     # exactly on voxel corners
    h = rng.choice([0.2 * d, 0.5 * d, d, 2 * d, 3.7 * d, 9 * d], 3000)
    h[::19] = 0.0; h[3::23] = np.nan
    box = (1.0, 1.0, 1.0) if periodic else None
    size, lo, hi = (40, 48, 33), (0.0, 0.0, 0.0), (1.0, 1.2, 0.825)
    o = oracle.bin3d(pos, h, size, lo, hi, small_max_vox=20, huge_min_bricks=60, periodic=periodic, box=box)
    ob = oracle.bbox3d(pos, h, size, lo, hi, periodic=periodic, box=box)
    g = gpu_bin3d(pos, h, size, lo, hi, periodic, box, small=20, huge=60)
    assert np.array_equal(g["bbox"], ob)
    assert np.array_equal(g["cls"], o["cls"])
    assert np.array_equal(g["sorted"], o["sorted"])
    assert np.array_equal(g["huge"], o["huge"])
    assert len(o["sorted"]) > 1000 and len(o["huge"]) > 0 and set(np.unique(o["cls"])) == {0, 1, 2, 3}


PATHS = {"default": {}, "all_direct": dict(small_max_vox=1 << 40), "all_bricks": dict(small_max_vox=1, huge_min_bricks=1 << 40),
         "all_global": dict(small_max_vox=1, huge_min_bricks=0)}


@pytest.mark.parametrize("kernel", ["cubic_spline_3d", "wendland_c2_3d"])
@pytest.mark.parametrize("path", list(PATHS))
def test_grid_vs_oracle(oracle, kernel, path):
    pos, h, prop = random_cloud(41, 3000, L=1.0, h_lo=0.0, h_hi=0.09, signed=True)
    size, lo, hi = (48, 40, 56), (0.0, 0.1, 0.0), (1.0, 0.9, 1.0)
    ref = oracle.grid3d(pos, h, prop, size, lo, hi, kernel=kernel)
    g, st = gpu_grid(pos, h, prop, size, lo, hi, kernel=kernel, **PATHS[path])
    check(g, ref)
    if path == "all_direct":
        assert st["n_pairs"] == 0 and st["n_huge"] == 0


def test_periodic_grid_conserves_mass(oracle):
    from astro_sph_tools_b200 import synthetic
    s = synthetic.s1(12, k=32)
    ref = oracle.grid3d(s["pos"], s["h"], s["mass"], (32, 32, 32), (0, 0, 0), (1, 1, 1), periodic=True, box=(1.0, 1.0, 1.0))
    g, _ = gpu_grid(s["pos"], s["h"], s["mass"], (32, 32, 32), (0, 0, 0), (1, 1, 1), periodic=True, box=1.0)
    check(g, ref)
    assert abs(g.sum() / 32 ** 3 - s["mass"].sum()) < 2e-3 * s["mass"].sum()      # 3-D normalised kernel, ~9-voxel support


def test_create_grid_host_api_and_rounds(oracle):
    from astro_sph_tools_b200.tools.projections import create_grid, wendland_c2_kernel_3d
    pos, h, prop = random_cloud(43, 2000, L=1.0, h_lo=0.02, h_hi=0.1)
    ref = oracle.grid3d(pos, h, prop, (40, 40, 40), (0, 0, 0), (1, 1, 1), kernel="wendland_c2_3d")
    g = create_grid(pos, h, prop, (40, 40, 40), 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, kernel_func=wendland_c2_kernel_3d)
    check(g, ref)
    g2, st = gpu_grid(pos, h, prop, (40, 40, 40), (0, 0, 0), (1, 1, 1), kernel="wendland_c2_3d", pair_capacity=3000)
    assert st["n_rounds"] > 2
    check(g2, ref)
    empty, _ = gpu_grid(np.zeros((0, 3)), np.zeros(0), np.zeros(0), (8, 8, 8), (0, 0, 0), (1, 1, 1))
    assert not empty.any()


def test_grid_against_the_literal_numpy_restatement():
    """independent pin (tests/literal3d.py): the reference's per-pixel rule carried to three dimensions, one voxel at a time,
    with the reference's own compiled kernel function when oracle/_ref travelled to this box"""
    from literal3d import grid3d_literal
    rng = np.random.default_rng(4)
    pos = rng.uniform(-0.05, 1.05, (800, 3))
    h = rng.choice([0.01, 0.03, 0.06, 0.15], 800)
    prop = rng.normal(size=800)
    size, lo, hi = (20, 17, 23), (0.0, 0.0, 0.1), (1.0, 0.9, 1.0)
    ref = grid3d_literal(pos, h, prop, size, lo, hi)
    for kw in ({}, dict(small_max_vox=1 << 40), dict(small_max_vox=1, huge_min_bricks=1 << 40), dict(small_max_vox=1, huge_min_bricks=0)):
        g, _ = gpu_grid(pos, h, prop, size, lo, hi, **kw)
        assert rel_l2(g, ref) <= 1e-5 and abs(g.sum() - ref.sum()) <= 1e-6 * np.abs(ref).sum()


def test_grid_capacities_are_windows_and_weights_keep_float64_range(oracle):
    """(high) a large-h list longer than huge_capacity used to fail AFTER the binning kernel had deposited the
    few-voxel particles -- and the retry added them twice in accumulate mode; (medium) float32 weights overflow / flush for
    weights like 1e48 or 1e-59.  Both windows are walked in passes now; the brick path sums in units of 2^E."""
    import torch
    from astro_sph_tools_b200.tools.projections import Gridder3D
    rng = np.random.default_rng(17)
    n = 4000
    pos = rng.uniform(0.0, 1.0, (n, 3))
    h = rng.choice([0.004, 0.02, 0.05, 0.12], n, p=[0.4, 0.3, 0.2, 0.1])          # direct, bricks, large-h list
    prop = rng.uniform(0.5, 1.5, n)
    size, lo, hi = (48, 48, 48), (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    ref = oracle.grid3d(pos, h, prop, size, lo, hi)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    g = Gridder3D(pair_capacity=2000, huge_capacity=16, huge_min_bricks=8)
    base = torch.full(size, 2.0, dtype=torch.float64, device="cuda")
    out = g.grid(d(pos), d(h), d(prop), size, lo, hi, out=base, accumulate=True)
    st = g.last_stats
    assert st["n_huge"] > 16 and st["n_rounds"] > 4 and st["n_pairs"] > 2000
    got = out.cpu().numpy() - 2.0
    assert rel_l2(got, ref) <= 1e-5 and abs(got.sum() - ref.sum()) <= 1e-6 * np.abs(ref).sum()
    for scale in (1e45, 1e-60, 1e300):
        for kw in ({}, dict(small_max_vox=1, huge_min_bricks=1 << 40), dict(small_max_vox=1, huge_min_bricks=0)):
            got, _ = gpu_grid(pos, h, prop * scale, size, lo, hi, **kw)
            assert np.isfinite(got).all()
            assert rel_l2(got / scale, ref) <= 1e-5 and abs((got / scale).sum() - ref.sum()) <= 1e-6 * np.abs(ref).sum()


@pytest.mark.parametrize("seed", range(12))
def test_random_grids_through_random_windows(oracle, seed):
    """random clouds, random grid shapes, periodic or not, random pair / large-h windows and class thresholds: the large-h split
    (a warp per entry enumerates its bricks) and the rounds over both windows give the oracle's grid"""
    rng = np.random.default_rng(500 + seed)
    n = int(rng.integers(200, 2500))
    size = tuple(int(v) for v in rng.integers(9, 40, 3))
    lo = tuple(rng.uniform(-0.1, 0.2, 3)); hi = tuple(np.array(lo) + rng.uniform(0.6, 1.1, 3))
    pos = rng.uniform(0.0, 1.0, (n, 3))
    vox = max((hi[c] - lo[c]) / size[c] for c in range(3))
    h = np.exp(rng.uniform(np.log(0.1 * vox), np.log(8 * vox), n))
    h[rng.random(n) < 0.03] = 0.0
    prop = rng.normal(size=n)
    periodic = bool(rng.random() < 0.4)
    box = (1.0, 1.0, 1.0) if periodic else None
    kw = dict(pair_capacity=int(rng.choice([300, 5000, 200_000])), huge_capacity=int(rng.choice([1, 5, 500])),
              huge_min_bricks=int(rng.choice([0, 2, 30, 512])), small_max_vox=int(rng.choice([1, 27, 200])))
    ref = oracle.grid3d(pos, h, prop, size, lo, hi, periodic=periodic, box=box)
    g, st = gpu_grid(pos, h, prop, size, lo, hi, periodic=periodic, box=box, **kw)
    if np.abs(ref).sum() == 0:
        assert not g.any()
        return
    tol = (1e-5, 1e-6) if kw["small_max_vox"] >= 8 else (1e-4, 2e-5)       # sub-voxel particles forced through the brick path
    assert rel_l2(g, ref) <= tol[0], (seed, kw, st)
    assert abs(g.sum() - ref.sum()) <= tol[1] * np.abs(ref).sum(), (seed, kw, st)
