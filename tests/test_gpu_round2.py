"""-m gpu: round-2 additions to the 2-D path, each against the CPU oracle --
  * capacities never fail after the direct deposits have started (host batches + accumulate mode + tiny huge_capacity),
  * weights far outside float32 range (the tile path stores mantissa + exponent, sums in units of 2^E),
  * the large-h split (a warp per large-h particle image emits its (tile, particle) pairs) on a big image, against the
    all-tiled path whose per-block pair counts exceed 2^20,
  * any pair window gives the same map.
"""
import numpy as np
import pytest

from conftest import rel_l2, random_cloud
from gpu_util import dev, gpu_project

pytestmark = pytest.mark.gpu


def check(img, ref, tol_l2=1e-5, tol_tot=1e-6):
    assert img.shape == ref.shape
    assert rel_l2(img, ref) <= tol_l2
    assert abs(img.sum() - ref.sum()) <= tol_tot * np.abs(ref).sum()


def mixed_cloud(seed, n, L=10.0):
    """tiny (direct deposit), medium (tiled) and very large (large-h list) supports in one set"""
    rng = np.random.default_rng(seed)
    pos = rng.uniform(0, L, (n, 3))
    h = rng.choice([0.004 * L, 0.02 * L, 0.05 * L, 0.2 * L], n, p=[0.4, 0.3, 0.2, 0.1])
    prop = rng.uniform(0.5, 1.5, n)
    return pos, h, prop


def test_host_batches_with_tiny_huge_capacity_do_not_double_count(oracle):
    """the direct deposits of the binning kernel used to be added twice when a batch in accumulate mode
    overflowed huge_capacity and the call was retried.  Capacities are windows now: no retry, no failure."""
    from astro_sph_tools_b200.tools.projections import Projector2D
    pos, h, prop = mixed_cloud(3, 6000)
    eng = Projector2D(huge_capacity=16, huge_min_tiles=4, pair_capacity=30000)
    img = eng.project_host(pos, h, prop, (256, 256), 2, (0.0, 10.0, 0.0, 10.0), batch_particles=1500)
    assert eng.last_stats["n_batches"] >= 4
    ref = oracle.project2d(pos, h, prop, (256, 256), 2, 0.0, 10.0, 0.0, 10.0)
    check(img, ref)
    # one device call, accumulate=True on top of an existing map, same tiny windows
    import torch
    base = torch.full((256, 256), 3.0, dtype=torch.float64, device="cuda")
    out = eng.project(dev(pos), dev(h), dev(prop), (256, 256), 2, (0.0, 10.0, 0.0, 10.0), out=base, accumulate=True)
    st = eng.last_stats
    assert st["n_huge"] > 16 and st["n_rounds"] > 4            # several large-h windows, several pair rounds each
    check(out.cpu().numpy() - 3.0, ref)


@pytest.mark.parametrize("scale", [1e45, 1e-60, 1e300])
def test_weights_outside_float32_range(oracle, scale):
    """a luminosity in erg/s times 1/(pi h^3) is ~1e48, Msun over cm^3 ~1e-59; the reference is float64"""
    pos, h, prop = mixed_cloud(8, 3000)
    ref = oracle.project2d(pos, h, prop, (192, 192), 2, 0.0, 10.0, 0.0, 10.0)
    for kw in ({}, dict(small_max_px=1, huge_min_tiles=1 << 40), dict(small_max_px=1, huge_min_tiles=0)):
        img, _ = gpu_project(pos, h, prop * scale, (192, 192), 2, (0.0, 10.0, 0.0, 10.0), **kw)
        assert np.isfinite(img).all()
        check(img / scale, ref)
    # two fields of very different magnitude in one pass keep their own exponents
    both, _ = gpu_project(pos, h, [prop * scale, prop * 1e-3], (192, 192), 2, (0.0, 10.0, 0.0, 10.0))
    check(both[0] / scale, ref)
    check(both[1] / 1e-3, ref)


def test_zero_and_signed_weights_keep_their_sign_and_zero(oracle):
    pos, h, _ = mixed_cloud(9, 2000)
    prop = np.random.default_rng(1).normal(size=2000)
    prop[::3] = 0.0
    img, _ = gpu_project(pos, h, prop, (128, 128), 2, (0.0, 10.0, 0.0, 10.0), small_max_px=1)
    ref = oracle.project2d(pos, h, prop, (128, 128), 2, 0.0, 10.0, 0.0, 10.0)
    assert rel_l2(img, ref) <= 1e-5
    zero, _ = gpu_project(pos, h, np.zeros(2000), (128, 128), 2, (0.0, 10.0, 0.0, 10.0), small_max_px=1)
    assert not zero.any()


def test_large_h_split_on_a_big_image(oracle):
    """300 particles whose supports cover ~5000 tiles each of a 4096^2 map.  Default: they go through the warp-per-particle
    split kernel.  huge_min_tiles = 2^40 forces them through the thread-per-particle enumeration, where one block of 256
    particles holds > 2^20 pairs (the packed 20-bit block count used to overflow).  Same map either way."""
    rng = np.random.default_rng(12)
    n = 300
    pos = rng.uniform(0.35, 0.65, (n, 3))
    h = rng.uniform(0.15, 0.175, n)
    prop = rng.uniform(0.5, 1.5, n)
    size, b = (4096, 4096), (0.0, 1.0, 0.0, 1.0)
    ref = oracle.project2d(pos, h, prop, size, 2, *b)
    split, st = gpu_project(pos, h, prop, size, 2, b)
    assert st["n_huge"] == n and st["n_pairs"] > (1 << 20)
    check(split, ref)
    tiled, st2 = gpu_project(pos, h, prop, size, 2, b, huge_min_tiles=1 << 40)
    assert st2["n_huge"] == 0 and st2["n_pairs"] == st["n_pairs"]
    check(tiled, ref)
    # a mix: many ordinary particles plus a few large ones, periodic images on
    pos2, h2, prop2 = mixed_cloud(5, 20000, L=1.0)
    ref2 = oracle.project2d(pos2, h2, prop2, (512, 512), 2, *b, periodic=True, box=(1.0, 1.0))
    img2, st3 = gpu_project(pos2, h2, prop2, (512, 512), 2, b, periodic=True, box=(1.0, 1.0), huge_min_tiles=16)
    assert st3["n_huge"] > 1000
    check(img2, ref2)
    # the same through small windows: many large-h windows, several pair rounds in each
    img3, st4 = gpu_project(pos2, h2, prop2, (512, 512), 2, b, periodic=True, box=(1.0, 1.0), huge_min_tiles=16, huge_capacity=64,
                            pair_capacity=50_000)
    assert st4["n_huge"] == st3["n_huge"] and st4["n_pairs"] == st3["n_pairs"] and st4["n_rounds"] > st4["n_huge"] // 64
    check(img3, ref2)


def test_large_h_pairs_bit_exact(oracle):
    """index work of the split: the pairs of the large-h entries follow the tiled pairs, entry by entry in list order, tiles
    in emit order -- array_equal with the oracle, before and after the stable sort"""
    from gpu_util import gpu_bin2d
    pos, h, _ = mixed_cloud(21, 5000)
    for periodic, box in ((False, None), (True, (10.0, 10.0))):
        o = oracle.bin2d(pos, h, (320, 320), 2, 0.0, 10.0, 0.0, 10.0, tile=32, small_max_px=16, huge_min_tiles=6, periodic=periodic, box=box)
        g = gpu_bin2d(pos, h, (320, 320), 2, (0.0, 10.0, 0.0, 10.0), periodic, box, 16, 6)
        assert len(o["huge"]) > 300 and np.array_equal(g["huge"], o["huge"])
        assert np.array_equal(g["pairs"], o["pairs"]) and np.array_equal(g["sorted"], o["sorted"])


def test_rounds_over_a_small_pair_window_equal_one_round():
    """any pair_capacity gives the same map: the pairs are walked through the window in rounds (random particle order, so every
    round touches every tile)"""
    import torch
    from astro_sph_tools_b200 import synthetic
    from astro_sph_tools_b200.tools.projections import Projector2D
    s = synthetic.s1(64, k=48, h_mode="uniform")
    perm = np.random.default_rng(2).permutation(len(s["h"]))
    pos, h, m = dev(s["pos"][perm]), dev(s["h"][perm]), dev(s["mass"][perm])
    args = ((512, 512), 2, (0.0, 1.0, 0.0, 1.0))
    one = Projector2D()
    a = one.project(pos, h, m, *args).clone()
    many = Projector2D(pair_capacity=300_000)
    b = many.project(pos, h, m, *args)
    torch.cuda.synchronize()
    assert one.last_stats["n_rounds"] == 1 and many.last_stats["n_rounds"] >= 7 and one.last_stats["n_pairs"] == many.last_stats["n_pairs"]
    assert rel_l2(b.cpu().numpy(), a.cpu().numpy()) < 1e-6


def test_spatial_pre_ordering_gives_the_same_map(oracle):
    """a randomly ordered set: AST_FLAG_ORDER_AUTO detects it from a sample and projects a copy ordered by tile; lattice order
    is left alone; 'always' / 'never' are honoured; the map is the oracle's either way (only the order of additions changes)"""
    import torch
    from astro_sph_tools_b200 import synthetic
    from astro_sph_tools_b200.tools.projections import Projector2D, Gridder3D
    s = synthetic.s1(48, k=48, h_mode="uniform")
    rng = np.random.default_rng(5)
    perm = rng.permutation(len(s["h"]))
    h = s["h"] * rng.uniform(0.05, 1.5, len(s["h"]))
    props = [s["mass"], s["mass"] * rng.uniform(1, 100, len(h))]
    ref = oracle.project2d(s["pos"], h, np.stack(props), (384, 384), 2, 0.0, 1.0, 0.0, 1.0)
    eng = Projector2D()
    args = ((384, 384), 2, (0.0, 1.0, 0.0, 1.0))
    lat = eng.project(dev(s["pos"]), dev(h), [dev(q) for q in props], *args, presort="auto").cpu().numpy()
    assert eng.last_stats["reordered"] is False
    rnd_in = (dev(s["pos"][perm]), dev(h[perm]), [dev(q[perm]) for q in props])
    rnd = eng.project(*rnd_in, *args, presort="auto").cpu().numpy()
    assert eng.last_stats["reordered"] is True
    again = eng.project(*rnd_in, *args).cpu().numpy()                      # device-resident default: 'never'
    assert eng.last_stats["reordered"] is False
    never = eng.project(*rnd_in, *args, presort="never").cpu().numpy()
    assert eng.last_stats["reordered"] is False
    always = eng.project(dev(s["pos"]), dev(h), [dev(q) for q in props], *args, presort="always").cpu().numpy()
    assert eng.last_stats["reordered"] is True
    for m in (lat, rnd, again, never, always):
        for k in range(2):
            check(m[k], ref[k])
    # 3-D
    g = Gridder3D()
    ref3 = oracle.grid3d(s["pos"], h, s["mass"], (40, 40, 40), (0, 0, 0), (1, 1, 1), periodic=True, box=(1.0, 1.0, 1.0))
    a3 = g.grid(dev(s["pos"][perm]), dev(h[perm]), dev(s["mass"][perm]), (40, 40, 40), (0, 0, 0), (1, 1, 1), periodic=True, box=1.0).cpu().numpy()
    assert g.last_stats["reordered"] is True
    b3 = g.grid(dev(s["pos"]), dev(h), dev(s["mass"]), (40, 40, 40), (0, 0, 0), (1, 1, 1), periodic=True, box=1.0).cpu().numpy()
    assert g.last_stats["reordered"] is False
    for m in (a3, b3):
        assert rel_l2(m, ref3) <= 1e-5 and abs(m.sum() - ref3.sum()) <= 1e-6 * np.abs(ref3).sum()


@pytest.mark.parametrize("periodic,world,wfrac", [(True, 4, 0.03), (True, 8, 0.2), (False, 3, 0.05), (True, 2, 0.6), (False, 5, 0.0)])
def test_slab_packing_kernel_equals_the_torch_route(periodic, world, wfrac):
    """ast_slab_route_count / _write (the exchange step of the multi-GPU k-NN) against the torch index-op definition of the same
    routing (distributed._route_torch, which the gloo tests run on the CPU): identical counts, send rows and source indices,
    for wrapped ghost zones, ghost zones wider than a slab, no ghost zone, and a width that covers the whole box"""
    import torch
    from astro_sph_tools_b200 import distributed as astd
    rng = np.random.default_rng(world * 7 + int(periodic))
    n = 70001
    length, lo = 2.5, (0.0 if periodic else -1.0)
    pos = rng.uniform(lo, lo + length, (n, 3))
    pos[:, 0] = lo + length * rng.beta(0.7, 1.3, n)                     # uneven along x: slabs of unequal width
    pos_d = torch.from_numpy(pos).cuda()
    x = pos_d[:, 0].contiguous()
    edges = np.sort(rng.choice(np.arange(1, 4096), world - 1, replace=False))
    bounds = torch.from_numpy(lo + length * np.concatenate([[0], edges, [4096]]) / 4096.0).cuda()
    bins = torch.clamp(((x - lo) / length * 4096).floor().long(), 0, 4095)
    owner = torch.bucketize(bins, torch.from_numpy(edges).cuda(), right=True)
    w = wfrac * length
    covers_all = periodic and w >= 0.5 * length
    ref_send, ref_src, ref_counts = astd._route_torch(pos_d, x, owner, bounds[:-1], bounds[1:], w, length, periodic, covers_all, world)
    send, src, counts = astd._SlabRouter(pos_d.device).route(pos_d, owner, bounds, w, length, periodic, covers_all, world)
    torch.cuda.synchronize()
    assert torch.equal(counts, ref_counts)
    assert torch.equal(src, ref_src) and torch.equal(send, ref_send)
    assert int(counts[0].sum()) == n                                     # every particle is owned exactly once


def test_slab_packing_kernel_edge_cases():
    """no particles at all; 32 destination ranks (the limit); every particle owned by one rank"""
    import torch
    from astro_sph_tools_b200 import distributed as astd
    dev = torch.device("cuda")
    router = astd._SlabRouter(dev)
    bounds = torch.linspace(0.0, 1.0, 5, dtype=torch.float64, device=dev)
    send, src, counts = router.route(torch.empty((0, 3), dtype=torch.float64, device=dev), torch.empty(0, dtype=torch.int64, device=dev),
                                     bounds, 0.1, 1.0, True, False, 4)
    assert send.shape == (0, 3) and src.numel() == 0 and int(counts.sum()) == 0
    rng = np.random.default_rng(3)
    pos = torch.from_numpy(rng.uniform(0, 1, (5000, 3))).to(dev)
    x = pos[:, 0].contiguous()
    b32 = torch.linspace(0.0, 1.0, 33, dtype=torch.float64, device=dev)
    owner = torch.clamp((x * 32).floor().long(), 0, 31)
    ref = astd._route_torch(pos, x, owner, b32[:-1], b32[1:], 0.01, 1.0, True, False, 32)
    got = router.route(pos, owner, b32, 0.01, 1.0, True, False, 32)
    torch.cuda.synchronize()
    assert torch.equal(got[2], ref[2]) and torch.equal(got[1], ref[1]) and torch.equal(got[0], ref[0])
    one = torch.zeros(5000, dtype=torch.int64, device=dev)
    b1 = torch.tensor([0.0, 1.0], dtype=torch.float64, device=dev)
    send, src, counts = router.route(pos, one, b1, 0.3, 1.0, True, False, 1)
    assert counts.tolist() == [[5000], [0]] and torch.equal(src, torch.arange(5000, device=dev)) and torch.equal(send, pos)
