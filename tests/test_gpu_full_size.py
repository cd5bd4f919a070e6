"""-m gpu: BASELINE.json's full sizes through size-independent properties (the oracle cannot run 256^3 in seconds):
config 2 (256^3 particles -> 2048^2, two weight fields) -- one-pass maps equal single-field passes, shards add up,
order does not matter, the deposited total equals the analytic sum; k-NN at 128^3 against scipy on a sample."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def config2():
    import torch
    from astro_sph_tools_b200 import synthetic
    from astro_sph_tools_b200.tools.smoothing import compute_smoothing_lengths_device
    pos, rng = synthetic.s1_positions(256)
    N = len(pos)
    pos_d = torch.from_numpy(pos).cuda()
    h_d = compute_smoothing_lengths_device(pos_d, 48, box_size=1.0)
    m_d = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
    T_d = torch.from_numpy(10.0 ** rng.uniform(4.0, 7.0, N)).cuda()
    return pos_d, h_d, m_d, m_d * T_d


def test_config2_properties(config2):
    import torch
    from astro_sph_tools_b200 import CoordinateAxes
    from astro_sph_tools_b200.tools.projections import Projector2D
    pos, h, m, mT = config2
    N = pos.shape[0]
    eng = Projector2D()
    args = ((2048, 2048), CoordinateAxes.Z, (0.0, 1.0, 0.0, 1.0), "cubic_spline_3d", True, 1.0)
    both = eng.project(pos, h, [m, mT], *args).clone()
    st = dict(eng.last_stats)
    assert st["n_pairs"] > 100e6 and st["n_rounds"] == 1
    # 1. each map of the one-pass pair equals its own single-field pass
    assert rel_l2(eng.project(pos, h, m, *args).cpu().numpy(), both[0].cpu().numpy()) < 1e-12
    assert rel_l2(eng.project(pos, h, mT, *args).cpu().numpy(), both[1].cpu().numpy()) < 1e-12
    # 2. particles shard by index: the partial maps of 3 unequal shards add up to the full map
    acc = torch.zeros_like(both)
    for lo, hi in ((0, N // 5), (N // 5, N // 2), (N // 2, N)):
        acc += eng.project(pos[lo:hi].contiguous(), h[lo:hi].contiguous(), [m[lo:hi].contiguous(), mT[lo:hi].contiguous()], *args)
    assert rel_l2(acc.cpu().numpy(), both.cpu().numpy()) < 1e-7
    # 3. order invariance (the sort has to undo a random permutation)
    perm = torch.randperm(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    shuffled = eng.project(pos[perm].contiguous(), h[perm].contiguous(), [m[perm].contiguous(), mT[perm].contiguous()], *args)
    assert rel_l2(shuffled.cpu().numpy(), both.cpu().numpy()) < 1e-7
    # 4. total: with the reference's 3-D normalised cubic spline and periodic images, sum(img) * A_pix -> 0.7 * sum(A_i / h_i)
    #    (SURVEY 8(a) A4); the pixel lattice samples each 36-pixel footprint finely enough for 1e-4
    total = float(both[0].sum()) / 2048 ** 2
    expect = 0.7 * float((m / h).sum())
    assert abs(total - expect) < 1e-4 * expect
    # 5. a small pair capacity (several rounds over the pair window) gives the same maps
    small = Projector2D(pair_capacity=40_000_000)
    again = small.project(pos, h, [m, mT], *args)
    assert small.last_stats["n_rounds"] >= 3 and rel_l2(again.cpu().numpy(), both.cpu().numpy()) < 1e-7


def test_knn_128cubed_sample_against_scipy():
    import torch
    from scipy.spatial import cKDTree
    from astro_sph_tools_b200 import synthetic
    from astro_sph_tools_b200.tools.smoothing import compute_smoothing_lengths_device
    pos, _ = synthetic.s1_positions(128)
    h = compute_smoothing_lengths_device(torch.from_numpy(pos).cuda(), 48, box_size=1.0).cpu().numpy()
    sel = np.random.default_rng(2).choice(len(pos), 20000, replace=False)
    ref = cKDTree(pos, boxsize=1.0).query(pos[sel], k=48, workers=-1)[0][:, 47]
    assert np.array_equal(h[sel], ref)
    # sortedness-free invariant: h is the K-th distance, so exactly >= K particles lie within h (self included)
    assert h.min() > 0 and np.isfinite(h).all()


# ---- BASELINE.json configs at FULL size against the oracle: the GPU deposits the whole set, the oracle a window of the map
# (or a sub-volume of the grid) fed every particle that can touch it -- pixel-for-pixel the same computation as the full map.
def _window_check(gpu_crop, ref):
    assert gpu_crop.shape == ref.shape and np.array_equal(gpu_crop != 0, ref != 0)
    assert rel_l2(gpu_crop, ref) <= 1e-5                                   # north_star: maps within 1e-5 relative L2
    assert abs(gpu_crop.sum() - ref.sum()) <= 1e-6 * abs(ref.sum())       # and totals within 1e-6


def test_config2_full_size_window_against_oracle(config2, oracle):
    import bench
    from astro_sph_tools_b200 import CoordinateAxes
    from astro_sph_tools_b200.tools.projections import Projector2D
    pos, h, m, mT = config2
    npix, wp = 2048, 256
    out = Projector2D().project(pos, h, [m, mT], (npix, npix), CoordinateAxes.Z, (0.0, 1.0, 0.0, 1.0))
    p0, lo, hi = bench.window_bounds(npix, wp)
    P, H = pos.cpu().numpy(), h.cpu().numpy()
    sel = bench.select_column(P, lo, hi, 2.0 * H.max())
    ref = oracle.project2d(P[sel], H[sel], np.stack([m.cpu().numpy()[sel], mT.cpu().numpy()[sel]]), (wp, wp), 2, lo, hi, lo, hi)
    crop = out[:, p0:p0 + wp, p0:p0 + wp].cpu().numpy()
    _window_check(crop[0], ref[0])
    _window_check(crop[1], ref[1])


@pytest.fixture(scope="module")
def config3():
    """S1 512^3 = 134 217 728 particles on ONE GPU (BASELINE.json configs[2] / the north_star target size), h = d_48"""
    import torch
    from astro_sph_tools_b200 import synthetic
    from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
    n = 512
    pos_d = torch.empty((n ** 3, 3), dtype=torch.float64, device="cuda")
    keep = []                                                             # host copy of a slab around x = 0.5 for the checks
    for i0, i1, blk in synthetic.s1_blocks(n):
        pos_d[i0:i1].copy_(torch.from_numpy(blk))
        sel = (blk[:, 0] > 0.42) & (blk[:, 0] < 0.62) & (blk[:, 1] > 0.42) & (blk[:, 1] < 0.62)
        keep.append((np.nonzero(sel)[0] + i0, blk[sel]))
    idx = np.concatenate([k[0] for k in keep]); col = np.ascontiguousarray(np.concatenate([k[1] for k in keep]))
    sol = SmoothingLengthSolver()
    h_d = sol.solve(pos_d, 48, 1.0)
    sol._ws = None
    torch.cuda.empty_cache()
    return pos_d, h_d, idx, col


def test_config3_full_size_window_against_oracle(config3, oracle):
    import torch
    import bench
    from astro_sph_tools_b200 import CoordinateAxes
    from astro_sph_tools_b200.tools.projections import Projector2D
    pos_d, h_d, idx, col = config3
    N, npix, wp = pos_d.shape[0], 4096, 256
    m_d = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
    eng = Projector2D()
    out = eng.project(pos_d, h_d, m_d, (npix, npix), CoordinateAxes.Z, (0.0, 1.0, 0.0, 1.0))
    assert eng.last_stats["n_pairs"] > 1.0e9
    p0, lo, hi = bench.window_bounds(npix, wp)
    h_col = h_d[torch.from_numpy(idx).cuda()].cpu().numpy()
    sel = bench.select_column(col, lo, hi, 2.0 * h_col.max())
    assert lo - 2.0 * h_col.max() > 0.42 and hi + 2.0 * h_col.max() < 0.62       # the kept slab holds every contributor
    ref = oracle.project2d(col[sel], h_col[sel], np.full(int(sel.sum()), 1.0 / N), (wp, wp), 2, lo, hi, lo, hi)
    _window_check(out[p0:p0 + wp, p0:p0 + wp].cpu().numpy(), ref)
    # deposited total of the whole map: sum(img) * A_pix = 0.7 * sum(m / h) for the 3-D-normalised cubic spline;
    # non-periodic projection loses what the particles near the box faces deposit outside the map
    total = float(out.sum()) / npix ** 2
    expect = 0.7 * float((m_d / h_d).sum())
    assert 0.0 < expect - total < 0.05 * expect
    del out, eng
    torch.cuda.empty_cache()


def test_more_than_2_to_the_32_pairs(config3, oracle):
    """the 134 M particles with 2.5 x h: 5e9 (tile, particle) pairs -- beyond 32 bits.  Global emit indices are 64-bit, the pair
    window (2^30) is indexed with 32 bits; five rounds.  Window of the map against the oracle."""
    import torch
    import bench
    from astro_sph_tools_b200 import CoordinateAxes
    from astro_sph_tools_b200.tools.projections import Projector2D
    pos_d, h_d, idx, col = config3
    N, npix, wp = pos_d.shape[0], 4096, 128
    m_d = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
    h_big = h_d * 2.5
    eng = Projector2D()
    out = eng.project(pos_d, h_big, m_d, (npix, npix), CoordinateAxes.Z, (0.0, 1.0, 0.0, 1.0))
    st = eng.last_stats
    assert st["n_pairs"] > (1 << 32) and st["n_rounds"] >= 5
    p0, lo, hi = bench.window_bounds(npix, wp)
    h_col = h_big[torch.from_numpy(idx).cuda()].cpu().numpy()
    sel = bench.select_column(col, lo, hi, 2.0 * h_col.max())
    assert lo - 2.0 * h_col.max() > 0.42 and hi + 2.0 * h_col.max() < 0.62
    ref = oracle.project2d(col[sel], h_col[sel], np.full(int(sel.sum()), 1.0 / N), (wp, wp), 2, lo, hi, lo, hi)
    _window_check(out[p0:p0 + wp, p0:p0 + wp].cpu().numpy(), ref)
    del out, eng
    torch.cuda.empty_cache()


def test_config5_knn_512cubed_subvolume_against_scipy(config3):
    """smoothing lengths of the 512^3 set (what configs[4] computes per GPU) against scipy on a sub-volume with margin"""
    import torch
    from scipy.spatial import cKDTree
    from astro_sph_tools_b200 import synthetic
    pos_d, h_d, idx, col = config3
    h_cap = 1.25 * synthetic.s1_h_lattice_estimate(512, 48)
    inner = (col[:, 0] > 0.5) & (col[:, 0] < 0.54) & (col[:, 1] > 0.5) & (col[:, 1] < 0.54) & (col[:, 2] > 0.9)   # wraps in z
    assert 0.5 - h_cap > 0.42 and 0.54 + h_cap < 0.62 and inner.sum() > 20000
    ref = cKDTree(col, boxsize=[0.0, 0.0, 1.0]).query(col[inner], k=48, workers=-1)[0][:, 47]
    got = h_d[torch.from_numpy(idx[inner]).cuda()].cpu().numpy()
    assert ref.max() <= h_cap                                             # the column holds every neighbour
    assert np.array_equal(got, ref)


def test_config4_nfw_full_size_subvolume_against_oracle(oracle):
    """BASELINE.json configs[3]: 256^3 NFW-clustered particles -> 512^3 voxels (periodic), a 40^3-voxel sub-volume against the oracle"""
    import torch
    import bench
    from astro_sph_tools_b200.tools.projections import Gridder3D
    from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
    n, ng, wv = 256, 512, 40
    N = n ** 3
    dev_ = torch.device("cuda")
    pos_d = bench.nfw_positions_device(torch, dev_, N)
    sol = SmoothingLengthSolver()
    h_d = sol.solve(pos_d, 48, 1.0)
    sol._ws = None
    m_d = torch.full((N,), 1.0 / N, dtype=torch.float64, device="cuda")
    g = Gridder3D()
    out = g.grid(pos_d, h_d, m_d, (ng,) * 3, (0, 0, 0), (1, 1, 1), periodic=True, box=1.0)
    assert g.last_stats["n_pairs"] > 1e8
    assert abs(float(out.sum()) / ng ** 3 - 1.0) < 2e-3                  # periodic: the mass is conserved up to the voxel sampling of h << voxel
    pos, h = pos_d.cpu().numpy(), h_d.cpu().numpy()
    # the sub-volume around the densest voxel's neighbourhood that lies in the box interior
    v0 = ng // 2 - wv // 2
    lo, hi = v0 / ng, (v0 + wv) / ng
    sel = np.all((pos + 2.0 * h[:, None] > lo) & (pos - 2.0 * h[:, None] < hi), axis=1)
    assert lo - 2 * h[sel].max() > 0.0 and hi + 2 * h[sel].max() < 1.0 and sel.sum() > 5000
    ref = oracle.grid3d(pos[sel], h[sel], np.full(int(sel.sum()), 1.0 / N), (wv,) * 3, (lo,) * 3, (hi,) * 3)
    _window_check(out[v0:v0 + wv, v0:v0 + wv, v0:v0 + wv].cpu().numpy(), ref)


def test_config1_as_worded_wendland_c2_periodic(oracle):
    """BASELINE.json configs[0]: 64^3 particles, periodic box, Wendland C2 surface density, 512 x 512 -- the whole map
    against the oracle, and mass conservation (2-D normalised kernel + periodic images: sum(img) * A_pix = sum(m))"""
    from astro_sph_tools_b200 import CoordinateAxes, synthetic
    from astro_sph_tools_b200.tools.projections import create_image, wendland_c2_kernel
    from astro_sph_tools_b200.tools.smoothing import compute_smoothing_lengths
    pos, _ = synthetic.s1_positions(64)
    h = compute_smoothing_lengths(pos, 48, box_size=1.0)
    m = np.full(len(h), 1.0 / len(h))
    img = create_image(pos, h, m, (512, 512), 32, CoordinateAxes.Z, 0.0, 1.0, 0.0, 1.0, kernel_func=wendland_c2_kernel, periodic=True, box_size=1.0)
    ref = oracle.project2d(pos, h, m, (512, 512), 2, 0.0, 1.0, 0.0, 1.0, kernel="wendland_c2_2d", periodic=True, box=(1.0, 1.0))
    _window_check(img, ref)
    assert abs(img.sum() / 512 ** 2 - 1.0) < 1e-6
