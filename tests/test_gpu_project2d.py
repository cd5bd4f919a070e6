"""-m gpu: 2-D projection maps from the CUDA path (through the C ABI) against the reference's golden maps and the
CPU oracle.  Tolerances are BASELINE.json's: relative L2 <= 1e-5 per map, total deposited weight within 1e-6."""
import numpy as np
import pytest

from conftest import golden_cases, load_golden, rel_l2, random_cloud

pytestmark = pytest.mark.gpu

REL_L2 = 1e-5      # north_star: "maps within 1e-5 relative L2 of the reference"
REL_SUM = 1e-6     # north_star: "total deposited mass conserved to 1e-6"


def check(m, ref):
    assert m.shape == ref.shape and m.dtype == np.float64
    assert rel_l2(m, ref) <= REL_L2
    assert abs(m.sum() - ref.sum()) <= REL_SUM * np.abs(ref).sum()


# the three execution paths: direct deposit only, tiled only, global large-h list only, and the default mix
PATHS = {"default": {}, "all_direct": dict(small_max_px=1 << 40), "all_tiled": dict(small_max_px=1, huge_min_tiles=1 << 40),
         "all_global": dict(small_max_px=1, huge_min_tiles=0)}


@pytest.mark.parametrize("name", golden_cases())
@pytest.mark.parametrize("path", list(PATHS))
def test_golden_maps(name, path):
    from gpu_util import gpu_project
    g = load_golden(name)
    n = int(g["npix"])
    m, st = gpu_project(g["pos"], g["h"], g["prop"], (n, n), int(g["axis"]), tuple(g["bounds"]), kernel=str(g["kernel"]),
                        **PATHS[path])
    check(m, g["ref_map"])
    if path == "all_direct":
        assert st["n_pairs"] == 0 and st["n_huge"] == 0
    if path == "all_global":
        assert st["n_huge"] > 0 and st["n_pairs"] > 0          # every tiled image goes through the large-h split kernel


def test_create_image_drop_in_signature():
    """the public function with the reference's positional signature, numpy in / numpy out"""
    from astro_sph_tools_b200 import CoordinateAxes
    from astro_sph_tools_b200.tools.projections import create_image, quartic_spline_kernel
    g = load_golden("s1_n16_p128_cubic_z")
    img = create_image(g["pos"], g["h"], g["prop"], (128, 128), 32, CoordinateAxes.Z, 0.0, 1.0, 0.0, 1.0)
    check(img, g["ref_map"])
    img2 = create_image(g["pos"], g["h"], g["prop"], (128, 128), 50, CoordinateAxes.Z, 0.0, 1.0, 0.0, 1.0,
                        kernel_func=quartic_spline_kernel)
    assert rel_l2(img, img2) < 1e-12                       # chunk_size does not change the result (float64 atomics reorder sums)
    g = load_golden("cloud_axis0")
    img = create_image(g["pos"], g["h"], g["prop"], (64, 64), 50, CoordinateAxes.X, *g["bounds"])
    check(img, g["ref_map"])


def test_periodic_wendland_config1_shape(oracle):
    from astro_sph_tools_b200 import CoordinateAxes
    from astro_sph_tools_b200.tools.projections import create_image, wendland_c2_kernel
    g = load_golden("s1_n16_p128_wc2_periodic")
    img = create_image(g["base_pos"], g["base_h"], g["base_prop"], (128, 128), 32, CoordinateAxes.Z, 0.0, 1.0, 0.0, 1.0,
                       kernel_func=wendland_c2_kernel, periodic=True, box_size=1.0)
    check(img, g["ref_map"])
    assert abs(img.sum() / 128 ** 2 - g["base_prop"].sum()) <= 1e-6 * g["base_prop"].sum()     # mass conservation


def test_arbitrary_python_kernel_func_is_served_by_the_table(oracle):
    """kernel_func is an arbitrary Python callable in the reference (_projector.py:86): the golden map below was produced by
    the reference with this very callable; here the callable is tabulated on the host and interpolated on the device"""
    import sys, os
    from astro_sph_tools_b200 import CoordinateAxes
    from astro_sph_tools_b200.tools.projections import create_image, create_grid
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    from gen_golden import wendland_c2_2d
    g = load_golden("s1_n16_p128_wc2_periodic")
    img = create_image(g["pos"], g["h"], g["prop"], (128, 128), 32, CoordinateAxes.Z, 0.0, 1.0, 0.0, 1.0, kernel_func=wendland_c2_2d)
    check(img, g["ref_map"])
    # a kernel that is not built in: truncated Gaussian, 2-D normalised; oracle = the literal definition in numpy
    gauss = lambda r, h: np.exp(-(r / h) ** 2) / (np.pi * h ** 2)
    pos, h, prop = random_cloud(5, 300, h_lo=0.05, h_hi=0.9)
    img = create_image(pos, h, prop, (48, 48), 16, CoordinateAxes.Z, 0.0, 10.0, 0.0, 10.0, kernel_func=gauss)
    X = (np.arange(48) * (10.0 / 48))
    ref = np.zeros((48, 48))
    for p_, h_, a_ in zip(pos, h, prop):
        r2 = (p_[0] - X[:, None]) ** 2 + (p_[1] - X[None, :]) ** 2
        ref += np.where(r2 < (2 * h_) ** 2, a_ * gauss(np.sqrt(r2), h_), 0.0)
    check(img, ref)
    # and in 3-D
    top = lambda r, h: (1.0 - 0.5 * r / h) * 3.0 / (4 * np.pi * h ** 3)
    grid = create_grid(pos[:, :3] / 10.0, h / 10.0, prop, (24, 24, 24), 0.0, 1.0, 0.0, 1.0, 0.0, 1.0, kernel_func=top)
    from literal3d import grid3d_literal                      # the reference's per-pixel rule carried to 3-D, in numpy, with this callable
    check(grid, grid3d_literal(pos[:, :3] / 10.0, h / 10.0, prop, (24, 24, 24), (0.0,) * 3, (1.0,) * 3, kernel_func=top))


def test_two_properties_one_pass():
    from astro_sph_tools_b200 import CoordinateAxes
    from astro_sph_tools_b200.tools.projections import create_images
    a, b = load_golden("s1_n12_p96_mass"), load_golden("s1_n12_p96_massT")
    maps = create_images(a["pos"], a["h"], [a["prop"], b["prop"]], (96, 96), 32, CoordinateAxes.Z, 0.0, 1.0, 0.0, 1.0)
    check(maps[0], a["ref_map"])
    check(maps[1], b["ref_map"])


@pytest.mark.parametrize("kernel", ["cubic_spline_3d", "wendland_c2_2d", "wendland_c2_3d", "cubic_spline_2d"])
@pytest.mark.parametrize("shape,bounds", [((100, 100), (0.0, 10.0, 0.0, 10.0)), ((70, 45), (2.0, 9.0, -1.0, 6.5)),
                                          ((33, 129), (0.0, 10.0, 0.0, 10.0))])
def test_random_clouds_vs_oracle(oracle, kernel, shape, bounds):
    import zlib
    from gpu_util import gpu_project
    seed = zlib.crc32(repr((kernel, shape)).encode()) % 1000
    pos, h, prop = random_cloud(seed, 5000, h_lo=0.0, h_hi=1.3, signed=True)
    h[::50] = 6.0                                          # a few very large ones
    ref = oracle.project2d(pos, h, prop, shape, 2, *bounds, kernel=kernel)
    m, _ = gpu_project(pos, h, prop, shape, 2, bounds, kernel=kernel)
    check(m, ref)
    # forcing every particle through the tile kernel / the global list: the float32 tile-relative coordinates need
    # h of the order of a pixel or more (the default classes guarantee it: bbox area > 16 pixels), so drop sub-pixel h here
    big = h > 0.5 * max((bounds[1] - bounds[0]) / shape[0], (bounds[3] - bounds[2]) / shape[1])
    ref = oracle.project2d(pos[big], h[big], prop[big], shape, 2, *bounds, kernel=kernel)
    for path in ("all_tiled", "all_global"):
        m, _ = gpu_project(pos[big], h[big], prop[big], shape, 2, bounds, kernel=kernel, **PATHS[path])
        check(m, ref)


def test_multiple_rounds_small_pair_capacity(oracle):
    from gpu_util import gpu_project
    pos, h, prop = random_cloud(11, 6000, h_hi=1.0)
    ref = oracle.project2d(pos, h, prop, (128, 128), 2, 0.0, 10.0, 0.0, 10.0)
    m, st = gpu_project(pos, h, prop, (128, 128), 2, (0.0, 10.0, 0.0, 10.0), pair_capacity=2000)
    assert st["n_rounds"] > 3
    check(m, ref)


def test_huge_capacity_grows(oracle):
    from gpu_util import gpu_project
    pos, h, prop = random_cloud(12, 3000, h_lo=2.0, h_hi=4.0)
    ref = oracle.project2d(pos, h, prop, (256, 256), 2, 0.0, 10.0, 0.0, 10.0)
    m, st = gpu_project(pos, h, prop, (256, 256), 2, (0.0, 10.0, 0.0, 10.0), huge_min_tiles=4, huge_capacity=16)
    assert st["n_huge"] > 16
    check(m, ref)


def test_edge_cases(oracle):
    from gpu_util import gpu_project
    # empty input
    m, st = gpu_project(np.zeros((0, 3)), np.zeros(0), np.zeros(0), (16, 16), 2, (0.0, 1.0, 0.0, 1.0))
    assert m.shape == (16, 16) and not m.any()
    # everything outside the window / non-positive / non-finite h
    pos = np.array([[50.0, 50.0, 0.0], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5], [0.5, 0.5, 0.5], [np.nan, 0.5, 0.5]])
    h = np.array([1.0, 0.0, -1.0, np.nan, 0.3])
    m, _ = gpu_project(pos, h, np.ones(5), (16, 16), 2, (0.0, 1.0, 0.0, 1.0))
    assert not m.any()
    # one particle exactly on a pixel corner: value A/(pi h^3) at that pixel (reference known answer)
    g = load_golden("single_particle")
    m, _ = gpu_project(g["pos"], g["h"], g["prop"], (10, 10), 2, (0.0, 10.0, 0.0, 10.0))
    assert np.unravel_index(m.argmax(), m.shape) == (3, 7)
    assert m.max() == pytest.approx(2.0 / (np.pi * 0.6 ** 3), rel=1e-6)
    assert sorted(map(tuple, np.argwhere(m != 0))) == [(2, 7), (3, 6), (3, 7), (3, 8), (4, 7)]
    # 1x1 image, and a particle much larger than the whole map
    pos, h, prop = random_cloud(13, 50, h_lo=20.0, h_hi=40.0)
    ref = oracle.project2d(pos, h, prop, (1, 1), 2, 0.0, 10.0, 0.0, 10.0)
    m, _ = gpu_project(pos, h, prop, (1, 1), 2, (0.0, 10.0, 0.0, 10.0))
    check(m, ref)


def test_batched_host_pipeline_matches_single_batch(oracle):
    """project_host streams the particles in batches (H2D of batch b+1 overlaps the deposit of batch b)"""
    from astro_sph_tools_b200.tools.projections import Projector2D
    pos, h, prop = random_cloud(15, 7001, h_hi=0.9, signed=True)
    ref = oracle.project2d(pos, h, np.stack([prop, 2 * prop]), (80, 80), 2, 0.0, 10.0, 0.0, 10.0)
    eng = Projector2D()
    one = eng.project_host(pos, h, [prop, 2 * prop], (80, 80), 2, (0.0, 10.0, 0.0, 10.0))
    ramped = eng.project_host(pos, h, [prop, 2 * prop], (80, 80), 2, (0.0, 10.0, 0.0, 10.0), batch_particles=1000)
    assert eng.last_stats["n_batches"] == 10            # 219 + 438 particles first, then batches of 876
    assert rel_l2(ramped, ref) <= 1e-5
    many = eng.project_host(pos, h, [prop, 2 * prop], (80, 80), 2, (0.0, 10.0, 0.0, 10.0), batch_particles=1000, ramp=False)
    assert eng.last_stats["n_batches"] == 8
    check(one[0], ref[0]); check(many[1], ref[1])
    assert rel_l2(many, one) < 1e-7           # only the float32 partial-sum grouping differs between batchings


def test_subpixel_regime_exact_support(oracle):
    """HBM-bound regime: supports below one pixel, particles on pixel corners; the direct-deposit path applies the exact
    float64 mask, so the set of non-zero pixels must equal the oracle's contributor mask"""
    from gpu_util import gpu_project
    rng = np.random.default_rng(77)
    n, npix = 50000, 200
    d = 10.0 / npix
    pos = rng.uniform(-0.2, 10.2, (n, 3))
    on = rng.random(n) < 0.3
    pos[on] = np.round(pos[on] / d) * d
    h = rng.choice([0.05 * d, 0.2 * d, 0.25 * d, 0.4 * d, 0.4999 * d, 0.5 * d, 0.7 * d], n)
    prop = rng.uniform(0.5, 1.5, n)
    for periodic, box in ((False, None), (True, (10.0, 10.0))):
        ref = oracle.project2d(pos, h, prop, (npix, npix), 2, 0.0, 10.0, 0.0, 10.0, periodic=periodic, box=box)
        cnt = oracle.contrib_count2d(pos, h, (npix, npix), 2, 0.0, 10.0, 0.0, 10.0, periodic=periodic, box=box)
        m, st = gpu_project(pos, h, prop, (npix, npix), 2, (0.0, 10.0, 0.0, 10.0), periodic=periodic, box=box)
        assert st["n_pairs"] == 0
        check(m, ref)
        assert not (m != 0)[cnt == 0].any()                     # never deposits outside the reference mask
        # the float32 shape function gives exactly 0 for q within ~1e-7 of 2 (e.g. h = d/2 particles sitting on a pixel
        # corner: their four neighbours are at r = 2h up to rounding); the reference deposits W ~ 1e-45 there
        assert not ((m == 0) & (ref > 1e-25 * ref.max())).any()


def test_properties_linearity_permutation(oracle):
    from gpu_util import gpu_project
    pos, h, prop = random_cloud(14, 4000, h_hi=0.9, signed=True)
    b = (0.0, 10.0, 0.0, 10.0)
    m1, _ = gpu_project(pos, h, prop, (96, 96), 2, b)
    m2, _ = gpu_project(pos, h, 3.0 * prop, (96, 96), 2, b)
    assert rel_l2(m2, 3.0 * m1) < 1e-6
    perm = np.random.default_rng(0).permutation(len(h))
    m3, _ = gpu_project(pos[perm], h[perm], prop[perm], (96, 96), 2, b)
    assert rel_l2(m3, m1) < 1e-6
    # axis permutation equivalence: projecting along X of (x,y,z) == along Z of (y,z,x)
    mx, _ = gpu_project(pos, h, prop, (96, 96), 0, b)
    mz, _ = gpu_project(np.ascontiguousarray(pos[:, [1, 2, 0]]), h, prop, (96, 96), 2, b)
    assert rel_l2(mx, mz) < 1e-12                          # only the order of the float64 atomic adds differs
    # sharding: sum of the maps of two halves == map of the whole (linear in particles)
    ma, _ = gpu_project(pos[:2000], h[:2000], prop[:2000], (96, 96), 2, b)
    mb, _ = gpu_project(pos[2000:], h[2000:], prop[2000:], (96, 96), 2, b)
    assert rel_l2(ma + mb, m1) < 1e-6


def test_kernel_functions_match_oracle(oracle):
    from astro_sph_tools_b200.tools import projections as P
    rng = np.random.default_rng(3)
    r = rng.uniform(0, 3, 10000); h = rng.uniform(0.5, 2.0, 10000)
    for fn, name in ((P.quartic_spline_kernel, "cubic_spline_3d"), (P.wendland_c2_kernel, "wendland_c2_2d"),
                     (P.wendland_c2_kernel_3d, "wendland_c2_3d"), (P.cubic_spline_kernel_2d, "cubic_spline_2d")):
        got = fn(r, h)
        ref = oracle.kernel_eval(name, r, h)
        assert np.allclose(got, ref, rtol=1e-13, atol=0)
        assert np.array_equal(got == 0, ref == 0)
    with pytest.raises(ValueError, match="Buffer dtype mismatch, expected 'double' but got 'float'"):
        P.quartic_spline_kernel(r.astype(np.float32), h.astype(np.float32))


def test_full_size_config1_properties(oracle):
    """config-1 size (64^3 -> 512^2, h = d_48): compared with the OpenMP oracle and the reference's summary values"""
    import os
    from astro_sph_tools_b200 import synthetic
    from conftest import GOLDEN_DIR
    from gpu_util import gpu_project
    s = synthetic.s1(64, k=48)
    m, st = gpu_project(s["pos"], s["h"], s["mass"], (512, 512), 2, (0.0, 1.0, 0.0, 1.0))
    ref = oracle.project2d(s["pos"], s["h"], s["mass"], (512, 512), 2, 0.0, 1.0, 0.0, 1.0)
    check(m, ref)
    f = os.path.join(GOLDEN_DIR, "s1_n64_p512_cubic_z_summary.npz")
    if os.path.exists(f):
        g = np.load(f)
        assert m.sum() / 512 ** 2 == pytest.approx(float(g["sumA"]), rel=1e-6)
        assert rel_l2(m[::8, ::8], g["sample"]) <= REL_L2
