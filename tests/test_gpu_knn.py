"""-m gpu: smoothing-length k-NN (ast_knn_h through the Python host layer) against scipy.spatial.cKDTree, the
reference's actual arithmetic.  Distances must be BIT-EQUAL; neighbour indices equal wherever distances are distinct
(inside groups of exactly equal distance scipy's order is traversal dependent)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gpu_knn(pos, k, box=None, lists=False, **kw):
    import torch
    from astro_sph_tools_b200.tools.smoothing import SmoothingLengthSolver
    sol = SmoothingLengthSolver(cell_target=kw.pop("cell_target", 0.0))
    pos_d = torch.from_numpy(np.ascontiguousarray(pos)).cuda()
    res = sol.solve(pos_d, k, box, want_neighbours=lists, want_distances=lists, **kw)
    torch.cuda.synchronize()
    if lists:
        return tuple(t.cpu().numpy() for t in res)
    return res.cpu().numpy()


@pytest.mark.parametrize("n,k,box", [(20000, 32, None), (30000, 48, 1.0), (5000, 1, None), (4097, 64, 1.0), (3000, 100, None)])
def test_h_bit_equal_to_scipy(oracle, n, k, box):
    rng = np.random.default_rng(n + k)
    pos = rng.uniform(0, 1.0, (n, 3))
    h_ref = oracle.knn_scipy(pos, k, box, workers=-1)[0]
    h = gpu_knn(pos, k, box)
    assert np.array_equal(h, h_ref)


def test_reference_default_api(oracle):
    """get_smoothing_lengths(positions): K = 32, self included, non-periodic == the reference's KDTree branch"""
    from astro_sph_tools_b200.tools.smoothing import get_smoothing_lengths, compute_smoothing_lengths
    pos = np.random.default_rng(0).normal(size=(8000, 3)) * [1.0, 3.0, 0.2] + 50.0      # anisotropic, offset extent
    h = get_smoothing_lengths(pos)
    assert np.array_equal(h, oracle.knn_scipy(pos, 32)[0])
    h48 = compute_smoothing_lengths(np.mod(pos, 7.0), k=48, box_size=7.0)
    assert np.array_equal(h48, oracle.knn_scipy(np.mod(pos, 7.0), 48, 7.0)[0])
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        get_smoothing_lengths(pos.astype(np.float32))


def test_clustered_and_lattice_sets(oracle):
    from astro_sph_tools_b200 import synthetic
    pos, _ = synthetic.s2_positions(40000, 1.0, n_haloes=8, seed=3)
    assert np.array_equal(gpu_knn(pos, 48, 1.0), oracle.knn_scipy(pos, 48, 1.0, workers=-1)[0])
    s = synthetic.s1(24, k=48)                       # S1 recipe: h from scipy inside the generator
    assert np.array_equal(gpu_knn(s["pos"], 48, 1.0), s["h"])


def test_neighbour_lists_and_ties(oracle):
    rng = np.random.default_rng(9)
    pos = rng.uniform(0, 1, (6000, 3))
    h, idx, dist = gpu_knn(pos, 16, None, lists=True)
    h_ref, d_ref, i_ref = oracle.knn_scipy(pos, 16)
    assert np.array_equal(dist, d_ref) and np.array_equal(h, h_ref)
    assert np.array_equal(idx, i_ref.astype(np.int32))
    # exact duplicates: distances still bit-equal, index SETS equal per group of equal distance
    pos2 = np.concatenate([pos[:2000], pos[:2000], pos[:2000]])
    h, idx, dist = gpu_knn(pos2, 8, None, lists=True)
    h_ref, d_ref, i_ref = oracle.knn_scipy(pos2, 8)
    assert np.array_equal(dist, d_ref)
    for r in range(0, 6000, 97):
        for d in np.unique(d_ref[r, :-1][d_ref[r, :-1] < d_ref[r, -1]]):          # complete tie groups only
            assert set(idx[r][dist[r] == d]) == set(i_ref[r][d_ref[r] == d])


def test_edge_cases(oracle):
    pos = np.random.default_rng(1).uniform(0, 1, (10, 3))
    assert np.all(np.isinf(gpu_knn(pos, 32)))                       # fewer points than k: inf, like scipy
    one = gpu_knn(pos[:1], 1)
    assert one.shape == (1,) and one[0] == 0.0
    flat = pos.copy(); flat[:, 2] = 0.25                            # degenerate extent along z
    assert np.array_equal(gpu_knn(flat, 4), oracle.knn_scipy(flat, 4)[0])


def test_query_slices_and_cell_size_invariance(oracle):
    rng = np.random.default_rng(12)
    pos = rng.uniform(0, 1, (25000, 3))
    full = gpu_knn(pos, 48, 1.0)
    # multi-GPU decomposition: every rank answers a slice of the queries against all positions
    parts = [gpu_knn(pos, 48, 1.0, q_begin=lo, q_count=hi - lo) for lo, hi in ((0, 6250), (6250, 12500), (12500, 25000))]
    assert np.array_equal(np.concatenate(parts), full)
    for ct in (2.0, 7.0, 40.0, 500.0):
        assert np.array_equal(gpu_knn(pos, 48, 1.0, cell_target=ct), full)


def test_nearest_halo_lookup_separate_query_set(oracle):
    """KDTree(centres, boxsize=L).query(particles): the reference's nearest-halo CLI (_scripts/find_nearest_haloes.py:207-215)"""
    from scipy.spatial import cKDTree
    from astro_sph_tools_b200.tools.smoothing import nearest_neighbours
    rng = np.random.default_rng(21)
    centres = rng.uniform(0, 25.0, (700, 3)); parts = rng.uniform(0, 25.0, (40000, 3))
    d_ref, i_ref = cKDTree(centres, boxsize=25.0).query(parts, workers=-1)
    d, i = nearest_neighbours(centres, parts, box_size=25.0)
    assert d.shape == (40000,) and np.array_equal(d, d_ref) and np.array_equal(i, i_ref)
    # open box, k = 5, queries partly outside the extent of the data
    parts2 = rng.uniform(-5.0, 30.0, (5000, 3))
    d_ref, i_ref = cKDTree(centres).query(parts2, k=5)
    d, i = nearest_neighbours(centres, parts2, k=5)
    assert np.array_equal(d, d_ref) and np.array_equal(i, i_ref)
    # fewer data points than k: scipy pads with inf / n
    d, i = nearest_neighbours(centres[:3], parts2[:10], k=5)
    assert np.all(np.isinf(d[:, 3:])) and np.all(i[:, 3:] == -1) and np.array_equal(d[:, :3], cKDTree(centres[:3]).query(parts2[:10], k=3)[0])


@pytest.mark.parametrize("n,k,box,ct", [(30000, 48, 1.0, 0.0), (30000, 48, None, 0.0), (2000, 8, 1.0, 60.0), (700, 16, 1.0, 90.0),
                                         (900, 4, 1.0, 0.5), (5000, 48, 3.0, 8.0), (64, 8, 1.0, 0.0)])
def test_all_query_kernels_agree_with_scipy(oracle, n, k, box, ct):
    """all query kernels (selection over blocks of cells + lock-step remainder = default, lock-step thread per query alone,
    diverging thread per query); small grids (G = 2..6) take
    the per-pair wrap path, large ones the constant-shift path"""
    rng = np.random.default_rng(n * 7 + k)
    pos = rng.uniform(0, box or 1.0, (n, 3))
    ref = oracle.knn_scipy(pos, k, box, workers=-1)[0]
    for kernel in ("select", "lockstep", "diverging"):
        assert np.array_equal(gpu_knn(pos, k, box, cell_target=ct, kernel=kernel), ref), kernel


def test_neighbour_lists_periodic_clustered_all_kernels(oracle):
    from astro_sph_tools_b200 import synthetic
    pos, _ = synthetic.s2_positions(30000, 1.0, n_haloes=5, seed=11)
    h_ref, d_ref, i_ref = oracle.knn_scipy(pos, 24, 1.0, workers=-1)
    distinct = np.all(np.diff(d_ref, axis=1) > 0, axis=1)           # rows without exactly tied distances
    assert distinct.mean() > 0.9
    for kernel in ("lockstep", "diverging"):
        h, idx, dist = gpu_knn(pos, 24, 1.0, lists=True, kernel=kernel)
        assert np.array_equal(h, h_ref) and np.array_equal(dist, d_ref), kernel
        assert np.array_equal(idx[distinct], i_ref[distinct]), kernel


def test_scattered_and_coherent_query_subsets(oracle):
    """index ranges of a spatially ordered set and of a shuffled set both give the slice of the full answer"""
    from astro_sph_tools_b200 import synthetic
    pos, rng = synthetic.s1_positions(28)                            # lattice order: index ranges are slabs
    full = oracle.knn_scipy(pos, 48, 1.0, workers=-1)[0]
    n = len(pos)
    for lo, hi in ((0, n // 8), (n // 3, n // 2), (n - 100, n)):
        assert np.array_equal(gpu_knn(pos, 48, 1.0, q_begin=lo, q_count=hi - lo), full[lo:hi])
    perm = rng.permutation(n)
    shuffled = np.ascontiguousarray(pos[perm])
    assert np.array_equal(gpu_knn(shuffled, 48, 1.0, q_begin=1000, q_count=2500), full[perm][1000:3500])


@pytest.mark.parametrize("box", [1.0, None])
def test_reach_limited_build_for_query_subsets(oracle, box):
    """a query subset (the per-rank call of a multi-GPU job) builds its cell list from the particles within reach only and
    widens the reach until every K-th distance is covered: same bits as the full build, for slabs, for a slab that wraps
    around the periodic box, for a scattered subset and for a clustered set with isolated particles"""
    from astro_sph_tools_b200 import synthetic
    pos, rng = synthetic.s1_positions(40)                            # lattice order: index ranges are x-slabs
    if box is None:
        pos = pos * 3.0 + 5.0
    n = len(pos)
    full = oracle.knn_scipy(pos, 48, box, workers=-1)[0]
    for lo, hi in ((0, n // 8), (3 * n // 8, n // 2), (7 * n // 8, n), (n // 2 - 50, n // 2 + 50)):
        got = gpu_knn(pos, 48, box, q_begin=lo, q_count=hi - lo)
        assert np.array_equal(got, full[lo:hi]), (lo, hi)
        assert np.array_equal(gpu_knn(pos, 48, box, q_begin=lo, q_count=hi - lo, full_build=True), got)
    shuffled = np.ascontiguousarray(pos[rng.permutation(n)])         # scattered subset: everything is in reach, full build
    full_s = oracle.knn_scipy(shuffled, 48, box, workers=-1)[0]
    assert np.array_equal(gpu_knn(shuffled, 48, box, q_begin=100, q_count=n // 10), full_s[100:100 + n // 10])
    # clustered set sorted along x, plus far-away stragglers whose 48th neighbour lies well beyond the first margin
    cl, _ = synthetic.s2_positions(30000, 1.0, n_haloes=6, seed=5)
    cl = cl[np.argsort(cl[:, 0])]
    if box is None:
        cl = np.concatenate([cl, np.array([[5.0, 5.0, 5.0], [-3.0, 0.5, 0.5]])])
    ref = oracle.knn_scipy(cl, 48, box, workers=-1)
    for lo, hi in ((0, 3000), (12000, 15000), (len(cl) - 2500, len(cl))):
        h, idx, dist = gpu_knn(cl, 48, box, lists=True, q_begin=lo, q_count=hi - lo)
        assert np.array_equal(h, ref[0][lo:hi]) and np.array_equal(dist, ref[1][lo:hi]), (lo, hi)


@pytest.mark.parametrize("kind,n,k,box", [("s1", 40, 48, 1.0), ("s1", 33, 48, None), ("uniform", 60000, 48, 1.0), ("uniform", 50000, 32, None),
                                          ("uniform", 50000, 100, 1.0), ("s2", 60000, 48, 1.0), ("s2", 60000, 16, None),
                                          ("aniso", 50000, 48, None), ("dup", 40000, 48, 1.0), ("uniform", 60000, 1, 1.0), ("uniform", 60000, 2, None),
                                          ("s1", 36, 64, 1.0), ("s1", 36, 77, 1.0), ("uniform", 300000, 48, 1.0)])
def test_selection_kernel_bit_equal_to_scipy(oracle, kind, n, k, box):
    """the selection kernel (float32 histogram + exact float64 edge band, knn_select.cuh) with its lock-step remainder against
    scipy on jittered lattices, uniform, clustered (most queries fall back), anisotropic extents (1.4 : 1 : 1, still on the
    fast path), and a set with exactly duplicated points (tied distances across the K-th)"""
    from astro_sph_tools_b200 import synthetic
    rng = np.random.default_rng(n + k)
    if kind == "s1":
        pos, _ = synthetic.s1_positions(n)
    elif kind == "s2":
        pos, _ = synthetic.s2_positions(n, 1.0, n_haloes=6, seed=3)
    else:
        pos = rng.uniform(0, 1.0, (n, 3))
    if kind == "aniso":
        pos = pos * np.array([1.4, 1.0, 1.0])
    if kind == "dup":
        pos[n // 2:] = pos[:n - n // 2]                            # every point twice: distances tie pairwise
    if box is None and kind in ("s1", "s2"):
        pos = pos * 2.0 - 7.0
    ref = oracle.knn_scipy(pos, k, box, workers=-1)[0]
    got = gpu_knn(pos, k, box, kernel="select")
    assert np.array_equal(got, ref)
    lo, cnt = len(pos) // 3, len(pos) // 4                          # a query slice through the fast path
    assert np.array_equal(gpu_knn(pos, k, box, kernel="select", q_begin=lo, q_count=cnt, full_build=True), ref[lo:lo + cnt])


def test_selection_kernel_gates_and_boundaries(oracle):
    """inputs at the edges of what the selection path accepts: points exactly on the faces of an open box, a set barely above
    the 4096-particle threshold, an extent ratio just inside and just outside the 1.5 limit, and a mix of a dense clump with a
    sparse field (fine grid for the clump, lock-step for the voids, selection for the rest)"""
    rng = np.random.default_rng(77)
    pos = rng.uniform(0, 1, (40000, 3))
    pos[:200] = np.round(pos[:200])                                     # on the corners / faces of [0, 1]^3
    assert np.array_equal(gpu_knn(pos, 48, None), oracle.knn_scipy(pos, 48, None, workers=-1)[0])
    small = rng.uniform(0, 1, (4100, 3))
    assert np.array_equal(gpu_knn(small, 16, 1.0), oracle.knn_scipy(small, 16, 1.0, workers=-1)[0])
    for ratio in (1.49, 1.51):
        q = rng.uniform(0, 1, (50000, 3)) * np.array([ratio, 1.0, 1.0])
        assert np.array_equal(gpu_knn(q, 32, None), oracle.knn_scipy(q, 32, None, workers=-1)[0]), ratio
    mix = np.concatenate([rng.uniform(0, 1, (60000, 3)), 0.5 + 0.004 * rng.normal(size=(40000, 3))]) % 1.0
    assert np.array_equal(gpu_knn(mix, 48, 1.0), oracle.knn_scipy(mix, 48, 1.0, workers=-1)[0])
    lo, cnt = 55000, 20000                                             # a query slice across the clump, reach-limited build
    assert np.array_equal(gpu_knn(mix, 48, 1.0, q_begin=lo, q_count=cnt), oracle.knn_scipy(mix, 48, 1.0, workers=-1)[0][lo:lo + cnt])
