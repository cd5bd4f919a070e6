"""Property tests (hypothesis) of the float64 index work on the CPU: for arbitrary windows, image sizes, positions and
smoothing lengths the product's geometry (csrc/ast_geom.h built for the host) equals the oracle's brute-force
definitions bit for bit: canonical bbox, classes and the (tile, particle) pair list with exact tile culling."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

from test_index_work_cpu import hostgeom, product_bbox_cls, product_pairs  # noqa: F401  (fixture + helpers)

finite = dict(allow_nan=False, allow_infinity=False)


@st.composite
def scene(draw):
    nx = draw(st.integers(1, 90)); ny = draw(st.integers(1, 90))
    x0 = draw(st.floats(-1e3, 1e3, **finite)); wx = draw(st.floats(1e-3, 1e3, **finite))
    y0 = draw(st.floats(-1e3, 1e3, **finite)); wy = draw(st.floats(1e-3, 1e3, **finite))
    n = draw(st.integers(1, 40))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    dx, dy = wx / nx, wy / ny
    pos = np.empty((n, 3))
    pos[:, 0] = x0 + rng.uniform(-0.3, 1.3, n) * wx
    pos[:, 1] = y0 + rng.uniform(-0.3, 1.3, n) * wy
    pos[:, 2] = rng.uniform(-1, 1, n)
    snap = rng.random(n) < 0.4                                   # sit exactly on computed sample points
    pos[snap, 0] = x0 + np.round((pos[snap, 0] - x0) / dx) * dx
    pos[snap, 1] = y0 + np.round((pos[snap, 1] - y0) / dy) * dy
    h = rng.choice([0.1, 0.25, 0.5, 0.75, 1.0, 1.5, 3.0, 17.0, 60.0], n) * rng.choice([dx, dy], n) * \
        rng.choice([1.0, 1.0 + 2e-16, 1.0 - 2e-16, 0.9999], n)
    small = draw(st.sampled_from([0, 4, 16, 100]))
    huge = draw(st.sampled_from([0, 3, 256]))
    axis = draw(st.sampled_from([0, 1, 2]))
    return dict(pos=pos, h=h, size=(nx, ny), bounds=(x0, x0 + wx, y0, y0 + wy), small=small, huge=huge, axis=axis)


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(sc=scene())
def test_geometry_equals_brute_force_definitions(oracle, hostgeom, sc):
    # with axis != 2 the in-plane columns change; permute so that the generated in-plane coordinates stay in the window
    cols = {0: (1, 2), 1: (0, 2), 2: (0, 1)}[sc["axis"]]
    pos = np.zeros_like(sc["pos"])
    pos[:, cols[0]] = sc["pos"][:, 0]; pos[:, cols[1]] = sc["pos"][:, 1]
    pos = np.ascontiguousarray(pos)
    args = (pos, sc["h"], sc["size"], sc["axis"]) + tuple(sc["bounds"])
    brute_bbox = oracle.bbox2d(*args, brute=True)
    assert np.array_equal(oracle.bbox2d(*args), brute_bbox)
    o = oracle.bin2d(*args, tile=32, small_max_px=sc["small"], huge_min_tiles=sc["huge"], brute=True, sort=False)
    bbox, cls = product_bbox_cls(hostgeom, oracle, pos, sc["h"], sc["size"], sc["axis"], sc["bounds"], sc["small"], sc["huge"])
    assert np.array_equal(bbox, brute_bbox)
    assert np.array_equal(cls, o["cls"])
    pairs = product_pairs(hostgeom, oracle, pos, sc["h"], sc["size"], sc["axis"], sc["bounds"], sc["small"], sc["huge"])
    assert np.array_equal(pairs, o["pairs"])
