"""-m gpu: index work of the CUDA path through the C ABI, bit-exact against the CPU oracle."""
import numpy as np
import pytest

from conftest import random_cloud
from test_index_work_cpu import adversarial, CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,bit_lo,n_bits", [(1, 32, 12), (100, 32, 3), (8192, 32, 14), (8193, 0, 20), (300001, 32, 17),
                                             (50000, 32, 0), (70000, 40, 9)])
def test_radix_sort_stable(n, bit_lo, n_bits):
    from gpu_util import gpu_sort
    rng = np.random.default_rng(n)
    e = rng.integers(0, 2 ** 63, n, dtype=np.int64).view(np.uint64)
    out = gpu_sort(e, bit_lo, n_bits)
    key = ((e >> np.uint64(bit_lo)) & np.uint64((1 << n_bits) - 1)).astype(np.int64)
    assert np.array_equal(out, e[np.argsort(key, kind="stable")])


def test_radix_sort_few_distinct_keys():
    from gpu_util import gpu_sort
    rng = np.random.default_rng(5)
    n = 200000
    e = ((rng.integers(0, 7, n).astype(np.uint64) * np.uint64(37)) << np.uint64(32)) | np.arange(n, dtype=np.uint64)
    out = gpu_sort(e, 32, 9)
    assert np.array_equal(out, e[np.argsort((e >> np.uint64(32)).astype(np.int64), kind="stable")])


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("periodic", [False, True])
def test_bin2d_bit_exact(oracle, case, periodic):
    from gpu_util import gpu_bin2d
    pos, h = adversarial(case["seed"] + 20, case["n"], case["npix"][0], case["bounds"][0], case["bounds"][1])
    box = (case["bounds"][1] - case["bounds"][0], case["bounds"][3] - case["bounds"][2]) if periodic else None
    small, huge = 9, 2
    o = oracle.bin2d(pos, h, case["npix"], case["axis"], *case["bounds"], tile=32, small_max_px=small, huge_min_tiles=huge,
                     periodic=periodic, box=box)
    ob = oracle.bbox2d(pos, h, case["npix"], case["axis"], *case["bounds"], periodic=periodic, box=box)
    g = gpu_bin2d(pos, h, case["npix"], case["axis"], case["bounds"], periodic, box, small, huge)
    assert np.array_equal(g["bbox"], ob)
    assert np.array_equal(g["cls"], o["cls"])
    assert np.array_equal(g["pairs"], o["pairs"])          # emit order
    assert np.array_equal(g["sorted"], o["sorted"])        # stable sort permutation
    assert np.array_equal(g["huge"], o["huge"])
    assert len(o["pairs"]) > 100 and len(o["huge"]) > 0


def test_bin2d_large_s1(oracle):
    """SPH-realistic set: 32^3 lattice, h = d_48, 256^2 map: ~10 tiles per particle"""
    from astro_sph_tools_b200 import synthetic
    from gpu_util import gpu_bin2d
    s = synthetic.s1(32, k=48)
    o = oracle.bin2d(s["pos"], s["h"], (256, 256), 2, 0.0, 1.0, 0.0, 1.0)
    g = gpu_bin2d(s["pos"], s["h"], (256, 256), 2, (0.0, 1.0, 0.0, 1.0))
    assert np.array_equal(g["cls"], o["cls"])
    assert np.array_equal(g["sorted"], o["sorted"]) and len(o["sorted"]) > 100000


@pytest.mark.parametrize("periodic", [False, True])
def test_contributor_mask_bit_exact(oracle, periodic):
    from gpu_util import gpu_contrib_count
    pos, h = adversarial(99, 3000, 80, 0.0, 10.0)
    box = (10.0, 10.0) if periodic else None
    ref = oracle.contrib_count2d(pos, h, (80, 80), 2, 0.0, 10.0, 0.0, 10.0, periodic=periodic, box=box)
    got = gpu_contrib_count(pos, h, (80, 80), 2, (0.0, 10.0, 0.0, 10.0), periodic, box)
    assert np.array_equal(got, ref) and ref.sum() > 1000
