"""-m gpu: the ingestion shim: unit-carrying arrays, in-place page-locking, the float32 transfer path and
the snapshot accessor driver give exactly what create_image gives on the plain float64 arrays."""
import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


class FakeUnytArray(np.ndarray):
    """stands in for unyt.unyt_array (an ndarray subclass carrying .units); unyt is not installed in this image"""
    def __new__(cls, a, units):
        o = np.asarray(a).view(cls)
        o.units = units
        return o


class FakeSnapshot:
    """the four accessors of the reference's SnapshotBase that feed the hot path (_SnapshotBase.py:599,618,708,889)"""
    def __init__(self, n, seed=0):
        r = np.random.default_rng(seed)
        self.pos = r.uniform(0, 1, (n, 3)); self.h = r.uniform(0.004, 0.03, n)
        self.m = r.uniform(0.5, 1.5, n); self.T = 10 ** r.uniform(4, 7, n)
    def get_positions(self, pt, use_proper_units=False): return FakeUnytArray(self.pos, "Mpc")
    def get_smoothing_lengths(self, pt, use_proper_units=False): return FakeUnytArray(self.h, "Mpc")
    def get_masses(self, pt): return FakeUnytArray(self.m, "Msun")
    def get_temperatures(self, pt): return FakeUnytArray(self.T, "K")


ARGS = ((192, 192), 32, 2, 0.0, 1.0, 0.0, 1.0)


def test_snapshot_maps_equals_create_image_on_plain_arrays(oracle):
    from astro_sph_tools_b200.tools.projections import create_image, snapshot_maps
    s = FakeSnapshot(50_000)
    maps = snapshot_maps(s, "gas", (192, 192), 2, 0.0, 1.0, 0.0, 1.0)
    m_ref = create_image(s.pos, s.h, s.m, *ARGS)
    mT_ref = create_image(s.pos, s.h, s.m * s.T, *ARGS)
    # (not bitwise: directly deposited particles use float64 atomics, whose order differs from run to run)
    assert rel_l2(maps["mass"], m_ref) <= 1e-13
    nz = m_ref != 0
    assert np.array_equal(maps["mass"] != 0, nz)
    assert rel_l2(maps["temperature"][nz], (mT_ref / np.where(nz, m_ref, 1))[nz]) <= 1e-12
    orc = oracle.project2d(s.pos, s.h, s.m, (192, 192), 2, 0.0, 1.0, 0.0, 1.0)
    assert rel_l2(maps["mass"], orc) <= 1e-5


def test_pinned_in_place_and_batched_float32_path():
    from astro_sph_tools_b200.tools.projections import create_image, pinned, default_projector
    s = FakeSnapshot(300_000, seed=3)
    ref = create_image(s.pos, s.h, s.m, *ARGS)
    with pinned(s.pos, s.h, s.m):
        assert rel_l2(create_image(s.pos, s.h, s.m, *ARGS), ref) <= 1e-13
    p32, h32, m32 = s.pos.astype(np.float32), s.h.astype(np.float32), s.m.astype(np.float32)
    want = create_image(p32.astype(np.float64), h32.astype(np.float64), m32.astype(np.float64), *ARGS)
    with pytest.raises(ValueError, match="expected 'double' but got 'float'"):
        create_image(p32, h32, m32, *ARGS)
    assert rel_l2(create_image(p32, h32, m32, *ARGS, allow_float32=True), want) <= 1e-13
    # batched (copy/compute overlapped) route with float32 staging
    eng = default_projector()
    got = eng.project_host(p32, h32, m32, (192, 192), 2, (0.0, 1.0, 0.0, 1.0), batch_particles=1 << 16)
    assert eng.last_stats["n_batches"] > 1
    assert rel_l2(got, want) <= 1e-6           # batches regroup the float32 partial sums inside the tile kernel
